"""`-m "not gpu"`: the CSR SpMM kernels of csrc/spmm.cu (LightGCN / NGCF propagation: `torch.sparse.mm` of
models/LightGCN.py:196 with the layer mean of :198-200 fused) executed ON THE HOST by the SIMT emulator of
tests/simt_host.py from their own source text - sub-warp groups that leave and loop independently (partial-mask
shuffles), every row-width class, the running-mean accumulation, and the long-row split form with its deterministic
partial sums - against scipy and the propagated tables the reference itself produced on ml-100k."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.timeout(1200)


@pytest.fixture(scope="module")
def simt(tmp_path_factory):
    from tests.simt_host import build_spmm
    return build_spmm(str(tmp_path_factory.mktemp("simt_spmm")))


def _spmm(simt, A, X, d, Y=None, acc=None, acc_scale=1.0, acc_init=0, seg_len=0, grid=3):
    indptr = np.ascontiguousarray(A.indptr, np.int64); indices = np.ascontiguousarray(A.indices, np.int32)
    values = np.ascontiguousarray(A.data, np.float32)
    P = lambda a: a.ctypes.data if a is not None else None
    plan = dict(seg_begin=None, seg_end=None, n_seg=0, long_rows=None, long_seg_ptr=None, n_long=0, partial=None)
    if seg_len:
        import torch
        from recsys_pytorch_b200 import engine
        pl = engine.spmm_plan(torch.from_numpy(indptr), seg_len=seg_len)        # the product's own host logic
        if pl.n_seg:
            plan = dict(seg_begin=pl.seg_begin.numpy(), seg_end=pl.seg_end.numpy(), n_seg=pl.n_seg,
                        long_rows=pl.long_rows.numpy(), long_seg_ptr=pl.long_seg_ptr.numpy(), n_long=pl.n_long,
                        partial=np.full(pl.n_seg * ((d + 3) // 4 * 4), np.nan, np.float32))
    rc = simt.emu_spmm(P(indptr), P(indices), P(values), A.shape[0], P(X), X.shape[1], d, P(Y), Y.shape[1] if Y is not None else 0,
                       P(acc), acc.shape[1] if acc is not None else 0, acc_scale, acc_init, seg_len, P(plan["seg_begin"]),
                       P(plan["seg_end"]), plan["n_seg"], P(plan["long_rows"]), P(plan["long_seg_ptr"]), plan["n_long"],
                       P(plan["partial"]), grid)
    return rc, plan["n_seg"]


def _graph(rng, n, max_deg, heavy=0):
    rows = []
    for r in range(n):
        deg = int(rng.integers(0, max_deg))
        if r < heavy:
            deg = int(rng.integers(300, min(n, 900)))                           # a few very long rows (popular items)
        rows.append(np.sort(rng.choice(n, deg, replace=False)))
    indptr = np.zeros(n + 1, np.int64); indptr[1:] = np.cumsum([len(r) for r in rows])
    indices = np.concatenate(rows).astype(np.int32)
    vals = rng.random(len(indices)).astype(np.float32)
    return sp.csr_matrix((vals, indices, indptr), shape=(n, n))


@pytest.mark.parametrize("d", [4, 8, 16, 30, 64, 128, 200])
def test_spmm_every_row_width_class_against_scipy(simt, d):
    """Y = A X and the running accumulation acc += s A X for every (G, CPL) the dispatcher distinguishes; empty rows,
    ragged row count, padded leading dimensions."""
    rng = np.random.default_rng(d)
    n = 203
    A = _graph(rng, n, 25)
    ld = (d + 3) // 4 * 4 + 4                                                   # a leading dimension larger than d
    X = np.zeros((n, ld), np.float32); X[:, :d] = rng.standard_normal((n, d))
    Y = np.full((n, ld), 7.0, np.float32)
    acc = np.zeros((n, ld), np.float32); acc[:, :d] = rng.standard_normal((n, d))
    acc0 = acc.copy()
    rc, _ = _spmm(simt, A, X, d, Y=Y, acc=acc, acc_scale=0.25)
    d4 = (d + 3) // 4
    G = 1
    while G < d4 and G < 32:
        G <<= 1
    assert rc == G * 100 + (d4 + G - 1) // G                                    # the intended instantiation ran
    want = (A.astype(np.float64) @ X[:, :d].astype(np.float64))
    np.testing.assert_allclose(Y[:, :d], want, rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(acc[:, :d], acc0[:, :d] + 0.25 * want, rtol=2e-5, atol=2e-6)
    assert (Y[:, d4 * 4:] == 7.0).all()                                         # nothing written beyond the row's float4s


def test_spmm_long_row_split_is_deterministic_and_equals_the_plain_form(simt):
    """Rows longer than seg_len are cut into segments (engine.spmm_plan), one sub-group per segment, partials added in
    order: same result as scipy, the same bits on every run, and within summation-order distance of the plain kernel."""
    rng = np.random.default_rng(1)
    n, d = 1000, 64
    A = _graph(rng, n, 12, heavy=5)
    X = rng.standard_normal((n, d)).astype(np.float32)
    Y1, Y2, Y3 = (np.zeros((n, d), np.float32) for _ in range(3))
    _, nseg = _spmm(simt, A, X, d, Y=Y1, seg_len=256)
    assert nseg >= 10
    _spmm(simt, A, X, d, Y=Y2, seg_len=256, grid=1)                            # another schedule: identical bits
    _spmm(simt, A, X, d, Y=Y3)                                                  # plain form
    assert np.array_equal(Y1, Y2)
    want = A.astype(np.float64) @ X.astype(np.float64)
    np.testing.assert_allclose(Y1, want, rtol=2e-5, atol=1e-5)
    np.testing.assert_allclose(Y3, want, rtol=2e-5, atol=1e-5)


def test_lightgcn_propagation_matches_the_reference_golden(simt, golden):
    """models/LightGCN.py:174-202 on ml-100k: three SpMM launches with the layer mean fused as a running accumulation
    (first layer: acc = s (E0 + A E0)) reproduce the propagated tables of the reference (tests/golden/lightgcn_ml100k.npz)."""
    g, ml = golden["lightgcn_ml100k"], golden["ml100k"]
    nu, ni, d, L = int(ml["num_users"]), int(ml["num_items"]), 16, 3
    A = sp.csr_matrix((g["adj_vals"], (g["adj_rows"], g["adj_cols"])), shape=(nu + ni, nu + ni)).tocsr()
    A.sort_indices()
    E0 = np.concatenate([g["U0"], g["V0"]]).astype(np.float32)
    out = np.zeros_like(E0)
    bufs = [np.zeros_like(E0), np.zeros_like(E0)]
    s, X = 1.0 / (L + 1), E0
    for layer in range(L):                                                      # lightgcn.py::propagate
        Y = bufs[layer & 1]
        _spmm(simt, A, X, d, Y=Y, acc=out, acc_scale=s, acc_init=1 if layer == 0 else 0, grid=8)
        X = Y
    np.testing.assert_allclose(out[:nu], g["prop_U"], rtol=2e-5, atol=1e-8)
    np.testing.assert_allclose(out[nu:], g["prop_V"], rtol=2e-5, atol=1e-8)
