"""`-m gpu`: the tcgen05 scoring path (B200REC_SCORE_TC).  The bar: the fused
tensor-core candidate pass + fp32 re-rank returns EXACTLY what the exact CUDA-core
kernel returns (which tests/test_gpu_parity.py pins bit-exactly to the C oracle)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from recsys_pytorch_b200 import engine  # noqa: E402
from recsys_pytorch_b200._lib import SCORE_EXACT, SCORE_TC  # noqa: E402


def _tables(rng, nu, ni, d, dev, std=1.0):
    U = engine.alloc_table(nu, d, dev, std=0.0); V = engine.alloc_table(ni, d, dev, std=0.0)
    U[:, :d] = torch.from_numpy((rng.standard_normal((nu, d)) * std).astype(np.float32)).to(dev)
    V[:, :d] = torch.from_numpy((rng.standard_normal((ni, d)) * std).astype(np.float32)).to(dev)
    return U, V


def _mask(rng, nu, ni, lo, hi, dev, heavy=()):
    rows = []
    for u in range(nu):
        n = hi * 6 if u in heavy else int(rng.integers(lo, hi))
        rows.append(np.sort(rng.choice(ni, min(n, ni - 200), replace=False)).astype(np.int32))
    indptr = np.zeros(nu + 1, np.int64); indptr[1:] = np.cumsum([len(r) for r in rows])
    return engine.DeviceCSR(torch.from_numpy(indptr).to(dev), torch.from_numpy(np.concatenate(rows)).to(dev), (nu, ni))


@pytest.mark.parametrize("d", [128, 64, 50, 256, 200, 150, 8])
def test_tc_raw_scores_are_the_bf16_gemm(dev, d):
    """TMA tiles + UMMA descriptors + TMEM read-back: the candidate pass computes bf16(U) . bf16(V)^T."""
    rng = np.random.default_rng(d)
    nu, ni = 300, 1000                      # ragged in both tile dimensions
    U, V = _tables(rng, nu, ni, d, dev)
    users = torch.from_numpy(rng.permutation(nu).astype(np.int32)).to(dev)
    got, su, sv, perm = engine.debug_tc_scores(U, V, d, users)
    assert sorted(perm.tolist()) == list(range(ni))                       # a permutation of the items ...
    norms = torch.linalg.norm(V[:, :d], dim=1)[perm]
    # ... by descending norm, to the resolution of the 16-bit sort key (sign, exponent, 7 mantissa bits); catalogues
    # of >= 8*17 tiles are visited head | stratified sample | rest instead (score_tc.cu::reorder_kernel)
    assert bool((norms[:-1] >= norms[1:] * (1 - 2.0 ** -7 - 1e-6)).all())
    Uh = (U[users.long(), :d] * su).to(torch.float16).to(torch.float64) / su   # power-of-two rescale + fp16 rounding
    Vh = (V[:, :d] * sv).to(torch.float16).to(torch.float64) / sv
    ref = (Uh @ Vh.T).to(torch.float32)
    scale = float(torch.linalg.norm(Uh, dim=1).max() * torch.linalg.norm(Vh, dim=1).max())
    err = float((got - ref).abs().max())
    assert err < 2e-6 * scale, f"max |tc - fp16 gemm| = {err} (scale {scale})"
    # and the rigorous bound the candidate filter relies on: |S~ - S| <= c |u| |v|
    exact = (U[users.long(), :d].double() @ V[:, :d].double().T)
    bound = (2.0 ** -10 * 1.05 + d * 2.4e-7) * torch.linalg.norm(U[users.long(), :d].double(), dim=1)[:, None] \
        * torch.linalg.norm(V[:, :d].double(), dim=1)[None, :]
    assert bool(((got.double() - exact).abs() <= bound).all())


@pytest.mark.parametrize("d,k,nu,ni,std", [(128, 10, 300, 3000, 1.0), (128, 100, 513, 20000, 1.0), (64, 100, 256, 9000, 0.1),
                                            (50, 10, 100, 1682, 1.0), (256, 100, 260, 5000, 0.05), (32, 5, 943, 1682, 1.0)])
def test_tc_equals_exact(dev, d, k, nu, ni, std):
    rng = np.random.default_rng(d + k)
    U, V = _tables(rng, nu, ni, d, dev, std)
    mask = _mask(rng, nu, ni, 0, 60, dev)
    users = torch.from_numpy(rng.permutation(nu).astype(np.int32)).to(dev)
    ie, se = engine.score_topk(U, V, d, users, mask, k, algo=SCORE_EXACT)
    it, st = engine.score_topk(U, V, d, users, mask, k, algo=SCORE_TC)
    assert torch.equal(it, ie), f"{(it != ie).sum().item()} of {it.numel()} ids differ"
    assert torch.equal(st, se)                      # same k-ordered fp32 FMA chain in the re-rank: bit-exact
    it2, _ = engine.score_topk(U, V, d, users, None, k, algo=SCORE_TC)
    ie2, _ = engine.score_topk(U, V, d, users, None, k, algo=SCORE_EXACT)
    assert torch.equal(it2, ie2)


def test_tc_equals_exact_heavy_tailed_norms_and_few_rows(dev):
    """Popular items with 10x norms (what a trained model looks like), K=100, and a user count small enough
    that the exact kernel itself runs in its item-split + merge form."""
    rng = np.random.default_rng(11)
    nu, ni, d, k = 700, 30000, 128, 100
    U, V = _tables(rng, nu, ni, d, dev, 0.1)
    V *= torch.exp(torch.randn(ni, 1, device=dev) * 0.8)
    mask = _mask(rng, nu, ni, 0, 100, dev)
    users = torch.from_numpy(rng.permutation(nu).astype(np.int32)).to(dev)
    ie, se = engine.score_topk(U, V, d, users, mask, k, algo=SCORE_EXACT)
    it, st = engine.score_topk(U, V, d, users, mask, k, algo=SCORE_TC)
    assert torch.equal(it, ie) and torch.equal(st, se)
    few = users[:5].contiguous()                                   # 5 rows -> hundreds of item splits
    i5, s5 = engine.score_topk(U, V, d, few, mask, k, algo=SCORE_EXACT)
    assert torch.equal(i5, ie[:5]) and torch.equal(s5, se[:5])


def test_tc_heavy_mask_rows_and_ties_fall_back_to_exact(dev):
    """Rows whose K + #masked exceeds the candidate buffer, and rows drowning in exact ties
    (integer tables: bf16 is exact, thousands of equal scores), are re-done by the exact kernel."""
    rng = np.random.default_rng(3)
    nu, ni, d, k = 300, 6000, 64, 100
    U = engine.alloc_table(nu, d, dev, std=0.0); V = engine.alloc_table(ni, d, dev, std=0.0)
    U[:, :d] = torch.from_numpy(rng.integers(-1, 2, (nu, d)).astype(np.float32)).to(dev)
    V[:, :d] = torch.from_numpy(rng.integers(-1, 2, (ni, d)).astype(np.float32)).to(dev)
    mask = _mask(rng, nu, ni, 0, 60, dev, heavy=set(range(0, nu, 7)))
    users = torch.arange(nu, dtype=torch.int32, device=dev)
    ie, se = engine.score_topk(U, V, d, users, mask, k, algo=SCORE_EXACT)
    it, st = engine.score_topk(U, V, d, users, mask, k, algo=SCORE_TC)
    assert torch.equal(it, ie) and torch.equal(st, se)


def test_tc_tiny_catalogue_and_rows_with_fewer_than_k_unmasked_items(dev):
    """a catalogue no larger than the candidate buffer, and users whose unmasked items number fewer than k
    (the reference then returns masked -inf ids) - the tensor-core path must hand such rows to the exact
    kernel instead of emitting the list sentinel; result identical to B200REC_SCORE_EXACT."""
    rng = np.random.default_rng(12)
    for (nu, ni, d, k) in ((70, 300, 64, 50), (64, 512, 128, 100), (33, 2000, 128, 100)):
        U = engine.alloc_table(nu, d, dev, 1.0); V = engine.alloc_table(ni, d, dev, 1.0)
        rows = []
        for u in range(nu):
            if u % 5 == 0:                                         # leaves only k - 3 unmasked items
                keep = rng.choice(ni, k - 3, replace=False)
                rows.append(np.setdiff1d(np.arange(ni), keep).astype(np.int32))
            else:
                rows.append(np.sort(rng.choice(ni, rng.integers(0, 40), replace=False)).astype(np.int32))
        indptr = np.zeros(nu + 1, np.int64); indptr[1:] = np.cumsum([len(r) for r in rows])
        mask = engine.DeviceCSR(torch.from_numpy(indptr).to(dev), torch.from_numpy(np.concatenate(rows)).to(dev), (nu, ni))
        users = torch.arange(nu, dtype=torch.int32, device=dev)
        ie, se = engine.score_topk(U, V, d, users, mask, k, algo=SCORE_EXACT)
        it, st = engine.score_topk(U, V, d, users, mask, k, algo=SCORE_TC)
        assert int(it.max()) < ni and int(it.min()) >= 0
        assert torch.equal(it, ie) and torch.equal(st, se)


def test_tc_large_catalogue_properties(dev):
    """BASELINE configs[1] item count (100k), top-100: sorted by (score desc, id asc), no masked item,
    identical to the exact kernel on a sample of rows."""
    rng = np.random.default_rng(5)
    nu, ni, d, k = 2048, 100_000, 128, 100
    U, V = _tables(rng, nu, ni, d, dev, 0.1)
    mask = _mask(rng, nu, ni, 10, 120, dev)
    users = torch.arange(nu, dtype=torch.int32, device=dev)
    it, st = engine.score_topk(U, V, d, users, mask, k, algo=SCORE_TC)
    s = st.cpu().numpy(); i = it.cpu().numpy()
    assert (np.diff(s, axis=1) <= 0).all()
    ties = np.diff(s, axis=1) == 0
    assert (np.diff(i, axis=1)[ties] > 0).all()
    ip, ix = mask.indptr.cpu().numpy(), mask.indices.cpu().numpy()
    for u in range(0, nu, 97):
        assert not set(i[u]) & set(ix[ip[u]:ip[u + 1]])
    sample = users[::16].contiguous()
    ie, se = engine.score_topk(U, V, d, sample, mask, k, algo=SCORE_EXACT)
    assert torch.equal(it[::16], ie) and torch.equal(st[::16], se)


def test_tc_more_users_than_one_launch_holds(dev):
    """The TC path works through the users in chunks of 148*256*2 = 75 776 rows (cfg5 has 10M): the second chunk
    re-uses the candidate lists, filters and the fp16 user tile of the first."""
    rng = np.random.default_rng(77)
    nu, ni, d, k = 75_776 + 300, 2_000, 32, 10
    U, V = _tables(rng, nu, ni, d, dev, 0.3)
    deg = rng.integers(0, 12, nu)
    indptr = np.zeros(nu + 1, np.int64); indptr[1:] = np.cumsum(deg)
    cols = np.empty(int(indptr[-1]), np.int32)
    for u in range(nu):                                     # sorted unique columns per row
        cols[indptr[u]:indptr[u + 1]] = np.sort(rng.choice(ni, deg[u], replace=False))
    mask = engine.DeviceCSR(torch.from_numpy(indptr).to(dev), torch.from_numpy(cols).to(dev), (nu, ni))
    users = torch.from_numpy(rng.permutation(nu).astype(np.int32)).to(dev)
    it, st = engine.score_topk(U, V, d, users, mask, k, algo=SCORE_TC)
    tail = torch.arange(75_776 - 200, nu, device=dev)       # rows on both sides of the chunk boundary
    ie, se = engine.score_topk(U, V, d, users[tail].contiguous(), mask, k, algo=SCORE_EXACT)
    assert torch.equal(it[tail], ie) and torch.equal(st[tail], se)
    head = torch.arange(0, 500, device=dev)
    ie, se = engine.score_topk(U, V, d, users[head].contiguous(), mask, k, algo=SCORE_EXACT)
    assert torch.equal(it[head], ie) and torch.equal(st[head], se)
