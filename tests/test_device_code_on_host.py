"""`-m "not gpu"`: the integer / index device functions of the CUDA sources (counter RNG, positive / negative sampling
with rejection against the sorted CSR row, shard ranges, the 64-bit ranking key) compiled FOR THE HOST from their own
source text (tests/host_device_code.py) and run against the Python mirrors the GPU parity tests rely on
(oracle/bpr_oracle.py).  The `-m gpu` suite checks the same mirror against the kernels on the device; this closes the
loop on a box without a GPU."""
import ctypes as C

import numpy as np
import pytest

from oracle import bpr_oracle as O
from recsys_pytorch_b200 import _lib


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    from tests.host_device_code import build
    return build(str(tmp_path_factory.mktemp("hostdev")))


def _csr(rng, nu, ni, lo, hi):
    rows = [np.sort(rng.choice(ni, int(rng.integers(lo, hi)), replace=False)).astype(np.int32) for _ in range(nu)]
    indptr = np.zeros(nu + 1, np.int64); indptr[1:] = np.cumsum([len(r) for r in rows])
    return indptr, (np.concatenate(rows) if indptr[-1] else np.zeros(0, np.int32)).astype(np.int32)


def _fetch(host, users, indptr, indices, ni, seed, step, pos=None, neg=None, item_range=None):
    a = _lib.BprArgs()
    B = len(users)
    users = np.ascontiguousarray(users, np.int32)
    a.users, a.B, a.num_items = users.ctypes.data, B, ni
    keep = [users]
    for name, arr in (("pos", pos), ("neg", neg)):
        if arr is not None:
            arr = np.ascontiguousarray(arr, np.int32); keep.append(arr)
            setattr(a, name, arr.ctypes.data)
    a.csr_indptr, a.csr_indices = indptr.ctypes.data, indices.ctypes.data
    a.seed, a.step = seed, step
    op, on = np.full(B, -7, np.int32), np.full(B, -7, np.int32)
    a.out_pos, a.out_neg = op.ctypes.data, on.ctypes.data
    if item_range is not None:
        a.item_lo, a.item_hi = item_range
    v, u, i, j = (np.zeros(B, np.int32) for _ in range(4))
    host.host_fetch_triples(C.byref(a), C.c_void_p(v.ctypes.data), C.c_void_p(u.ctypes.data), C.c_void_p(i.ctypes.data),
                            C.c_void_p(j.ctypes.data))
    return v.astype(bool), u, i, j, op, on


def test_counter_rng_matches_the_mirror(host):
    rng = np.random.default_rng(0)
    for _ in range(2000):
        seed, step, idx = (int(x) for x in rng.integers(0, 2**63, 3))
        draw = int(rng.integers(0, 300))
        assert host.host_rng_u32(seed, step, idx, draw) == O.rng_u32(seed, step, idx, draw)
    idx = rng.integers(0, 2**40, 500).astype(np.uint64)
    vec = O.rng_u32_vec(2020, 7, idx, 3)
    assert [host.host_rng_u32(2020, 7, int(t), 3) for t in idx] == [int(x) for x in vec]


@pytest.mark.parametrize("seed", range(4))
def test_device_sampler_source_equals_the_mirror(host, seed):
    """fetch_triple (csrc/sampler.cuh) for a whole batch == oracle sample_triple / sample_triples_vec: same positives,
    same negatives, every negative a non-positive, skipped triples reported as -1."""
    rng = np.random.default_rng(seed)
    nu, ni, B = 400, int(rng.integers(50, 3000)), 1500
    indptr, indices = _csr(rng, nu, ni, 1, min(40, ni // 2))
    users = rng.integers(0, nu, B)
    valid, u, i, j, op, on = _fetch(host, users, indptr, indices, ni, 2020 + seed, 5 + seed)
    pos, neg = O.sample_triples_vec(2020 + seed, 5 + seed, users, indptr, indices, ni)
    assert valid.all() and np.array_equal(u, users)
    assert np.array_equal(i, pos) and np.array_equal(j, neg) and np.array_equal(op, pos) and np.array_equal(on, neg)
    for t in range(0, B, 7):
        row = indices[indptr[users[t]]:indptr[users[t] + 1]]
        assert i[t] in row and j[t] not in row and 0 <= j[t] < ni
        assert (int(i[t]), int(j[t])) == O.sample_triple(2020 + seed, 5 + seed, t, int(users[t]), indptr, indices, ni)
    # given positives: only the negative is drawn (and it does not depend on the positive)
    v2, _, i2, j2, _, _ = _fetch(host, users, indptr, indices, ni, 2020 + seed, 5 + seed, pos=pos[::-1].copy())
    assert v2.all() and np.array_equal(i2, pos[::-1]) and np.array_equal(j2, neg)
    # given both: nothing is sampled, nothing is written back
    v3, _, i3, j3, op3, on3 = _fetch(host, users, indptr, indices, ni, 1, 1, pos=pos, neg=neg)
    assert v3.all() and np.array_equal(i3, pos) and np.array_equal(j3, neg) and (op3 == -7).all() and (on3 == -7).all()


def test_device_sampler_edge_cases(host):
    """A user without positives emits no triple; a user whose positives cover the whole sampling range is skipped after
    64 rejected draws (never trained positive-vs-positive); item-sharded ranges keep only the owned triples
    and draw the negative from the shard."""
    ni = 64
    rows = [np.arange(ni, dtype=np.int32), np.zeros(0, np.int32), np.arange(0, ni, 2, dtype=np.int32)]
    indptr = np.array([0, ni, ni, ni + ni // 2], np.int64)
    indices = np.concatenate(rows)
    users = np.array([0, 1, 2] * 50, np.int32)
    valid, _, i, j, op, on = _fetch(host, users, indptr, indices, ni, 9, 1)
    assert not valid[0::3].any() and (on[0::3] == -1).all() and (op[0::3] == -1).all()      # owns everything
    assert not valid[1::3].any()                                                          # no positives
    assert valid[2::3].all() and (i[2::3] % 2 == 0).all() and (j[2::3] % 2 == 1).all()
    for t in range(2, 150, 3):
        assert (int(i[t]), int(j[t])) == O.sample_triple(9, 1, t, 2, indptr, indices, ni)
    # item-sharded: rank range [16, 48)
    valid, _, i, j, _, on = _fetch(host, users, indptr, indices, ni, 9, 1, item_range=(16, 48))
    own = valid[2::3]
    assert own.any() and not own.all()
    assert ((i[2::3][own] >= 16) & (i[2::3][own] < 48)).all() and ((j[2::3][own] >= 16) & (j[2::3][own] < 48)).all()
    assert (j[2::3][own] % 2 == 1).all() and (on[2::3][~own] == -1).all()
    for t in np.arange(2, 150, 3)[own]:
        p, n = O.sample_triple(9, 1, int(t), 2, indptr, indices, ni, item_bounds=[0, 16, 48, 64])
        assert (int(i[t]), int(j[t])) == (p, n)


def test_head_plus_shard_negative_sampler(host):
    """sample_neg2 (replicated head [0, head) U one tail shard, csrc/p2p.cu's router) against the mirror."""
    rng = np.random.default_rng(3)
    ni, head = 500, 40
    bounds = [head, 150, 320, ni]
    indptr, indices = _csr(rng, 200, ni, 1, 60)
    for t in range(600):
        user = int(rng.integers(0, 200))
        row = np.ascontiguousarray(indices[indptr[user]:indptr[user + 1]])
        r = int(rng.integers(0, 3))
        j = C.c_int(0)
        ok = host.host_sample_neg2(row.ctypes.data, len(row), head, bounds[r], bounds[r + 1] - bounds[r], 11, 4, t, C.byref(j))
        # mirror: a positive inside shard r makes sample_triple use exactly that union
        neg = 0
        for tr in range(64):
            q = O.rng_u32(11, 4, t, 1 + tr) * (head + bounds[r + 1] - bounds[r]) >> 32
            neg = q if q < head else bounds[r] + (q - head)
            k = int(np.searchsorted(row, neg))
            if not (k < len(row) and int(row[k]) == neg):
                break
        else:
            neg = -1
        assert (j.value if ok else -1) == neg
        if ok:
            assert j.value not in row and (j.value < head or bounds[r] <= j.value < bounds[r + 1])


def test_ranking_key_orders_by_score_desc_then_id_asc(host):
    """make_key / key_id / key_score (csrc/common.cuh): the 64-bit key every top-k list sorts by.  Larger key == better
    == (score descending, item id ascending) - the documented tie order; -inf (masked) sorts below every
    finite score; the round trip is exact."""
    rng = np.random.default_rng(5)
    scores = np.concatenate([rng.standard_normal(300).astype(np.float32) * 10, np.float32([0.0, 1e-38, -1e-38, 3e38, -3e38]),
                             np.float32([-np.inf, np.inf]), np.round(rng.standard_normal(200)).astype(np.float32)])
    ids = rng.integers(0, 2**31 - 1, len(scores)).astype(np.int32)
    keys = np.array([host.host_make_key(float(s), int(i)) for s, i in zip(scores, ids)], np.uint64)
    for k, s, i in zip(keys, scores, ids):
        assert host.host_key_id(int(k)) == int(i) and np.float32(host.host_key_score(int(k))) == s
    order = np.argsort(keys)[::-1]                                   # best first
    # score desc, id asc - on the IEEE total order: the key ranks -0.0 just below +0.0 (a float compare calls them equal;
    # a dot product accumulated from +0.0 never yields -0.0, so only caller-supplied score blocks can show the difference)
    want = np.lexsort((ids, np.signbit(scores) & (scores == 0), -scores.astype(np.float64)))
    assert np.array_equal(scores[order], scores[want]) and np.array_equal(ids[order], ids[want])
    assert np.signbit(scores).any() and (scores == 0).sum() >= 2     # the -0.0 / +0.0 case is exercised


# ---- the metric kernels of csrc/metrics.cu, thread by thread on the host ------------------------------------------------
@pytest.fixture(scope="module")
def host_metrics(tmp_path_factory):
    from tests.host_device_code import build_metrics
    return build_metrics(str(tmp_path_factory.mktemp("hostmetrics")))


@pytest.mark.parametrize("seed", range(6))
def test_metric_kernels_source_is_bitwise_the_reference_cpp(host_metrics, seed):
    """holdout_kernel / loo_kernel (one thread per user) executed on the host from their own source == the reference's
    holdout.h / loo.h (oracle/_ref when built, else the C restatement, which the differential tests pin to it), bit for
    bit, incl. the row_ids indirection the sharded evaluation uses."""
    import os
    from tests.util import OracleC
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref_path = os.path.join(root, "oracle", "_ref", "libref_eval.so")
    if os.path.exists(ref_path):
        from oracle.ref_harness import RefNative
        ref = RefNative()
    else:
        ref = OracleC(os.path.join(root, "oracle", "liboracle.so"))
    rng = np.random.default_rng(40 + seed)
    users, items, max_k = int(rng.integers(1, 400)), int(rng.integers(80, 4000)), int(rng.integers(1, 60))
    ks = np.array(sorted(set([1, max_k] + rng.integers(1, max_k + 1, 3).tolist())), np.int32)
    n_truth_rows = users + 5                                                   # truth CSR has more rows than scored users
    truths = []
    for _ in range(n_truth_rows):
        truths.append(np.sort(rng.choice(items, int(rng.integers(1, 2 * max_k + 2)), replace=False)).astype(np.int32))
    row_ids = rng.permutation(n_truth_rows)[:users].astype(np.int32)
    topk = np.stack([rng.choice(items, max_k, replace=False) for _ in range(users)]).astype(np.int32)
    for r in range(0, users, 2):                                               # plant hits
        t = truths[row_ids[r]]
        m = min(len(t), max_k, 3)
        topk[r, rng.choice(max_k, m, replace=False)] = t[:m]
        if len(set(topk[r].tolist())) < max_k:                                 # keep rankings duplicate-free
            topk[r] = rng.choice(items, max_k, replace=False)
    tptr = np.zeros(n_truth_rows + 1, np.int64); tptr[1:] = np.cumsum([len(t) for t in truths])
    tidx = np.concatenate(truths).astype(np.int32)
    P = lambda a: C.c_void_p(a.ctypes.data)
    out = np.zeros((users, 3 * len(ks)), np.float32)
    host_metrics.host_holdout(P(topk), users, max_k, P(row_ids), P(tptr), P(tidx), P(ks), len(ks), P(out))
    np.testing.assert_array_equal(out, ref.holdout(topk, [truths[r] for r in row_ids], ks))
    np.testing.assert_array_equal(out, O.holdout_metrics(topk, [truths[r] for r in row_ids], ks))
    out = np.zeros((users, 2 * len(ks)), np.float32)
    host_metrics.host_loo(P(topk), users, max_k, P(row_ids), P(tptr), P(tidx), P(ks), len(ks), P(out))
    np.testing.assert_array_equal(out, ref.loo(topk, [truths[r][:1] for r in row_ids], ks))
    # without the indirection: row r of the truth CSR belongs to scored row r
    out2 = np.zeros((users, 3 * len(ks)), np.float32)
    host_metrics.host_holdout(P(topk), users, max_k, None, P(tptr), P(tidx), P(ks), len(ks), P(out2))
    np.testing.assert_array_equal(out2, ref.holdout(topk, truths[:users], ks))


# ---- P2P work order / shard lookup, fp16 rescale ---------------------------------------------------------------------------
@pytest.fixture(scope="module")
def host_misc(tmp_path_factory):
    from tests.host_device_code import build_misc
    return build_misc(str(tmp_path_factory.mktemp("hostmisc")))


def test_p2p_owner_lookup_matches_host_bounds_logic(host_misc):
    from recsys_pytorch_b200.p2p import owner_from_bounds
    rng = np.random.default_rng(1)
    for W in (1, 2, 3, 4, 8, 16):
        n = 5000
        cuts = np.sort(rng.choice(np.arange(1, n), W - 1, replace=False)) if W > 1 else np.zeros(0, int)
        bounds = np.concatenate([[0], cuts, [n]]).astype(np.int32)
        ids = np.concatenate([bounds[:-1], np.maximum(bounds[1:] - 1, 0), rng.integers(0, n, 300)])
        want = owner_from_bounds(ids, bounds.tolist())
        got = [host_misc.host_owner_of(C.c_void_p(bounds.ctypes.data), W, int(i)) for i in ids]
        assert got == want.tolist()
        assert all(bounds[r] <= i < bounds[r + 1] for i, r in zip(ids, got))


@pytest.mark.parametrize("round_robin", [0, 1])
def test_p2p_chunk_order_visits_every_chunk_exactly_once(host_misc, round_robin):
    """chunk_of (csrc/p2p.cu): the dynamic chunk counter c -> (source slot, offset) must be a bijection onto the chunks of
    every source, for both visiting orders, any world size, empty sources included - a slip here would drop or repeat
    32 triples silently.  pref / m are computed the way the kernel's thread 0 does."""
    rng = np.random.default_rng(7 + round_robin)
    for case in range(300):
        W = int(rng.choice([1, 2, 3, 4, 5, 8, 16]))
        cnt = rng.integers(0, 400, W)
        if case % 5 == 0:
            cnt[rng.integers(0, W)] = 0
        if case % 7 == 0:
            cnt[:] = 0
        chunks = (cnt + 31) >> 5
        pref = np.zeros(W + 1, np.int32); pref[1:] = np.cumsum(chunks)
        R, nl = int(pref[W - 1]), int(pref[W] - pref[W - 1])
        m = min(nl, R // (W - 1)) if W > 1 else 0
        if round_robin:
            m = int(chunks.min())
        seen = set()
        k, off = C.c_int(0), C.c_int(0)
        for c in range(int(pref[W])):
            host_misc.host_chunk_of(c, W, m, C.c_void_p(pref.ctypes.data), round_robin, C.byref(k), C.byref(off))
            assert 0 <= k.value < W and off.value % 32 == 0 and 0 <= off.value < 32 * chunks[k.value], (case, c, k.value, off.value)
            seen.add((k.value, off.value))
        assert len(seen) == int(pref[W])                       # no chunk twice -> with the range check: every chunk once
        if not round_robin and W > 1 and m > 0:                # a local chunk (slot W-1) after every W-1 remote chunks
            host_misc.host_chunk_of(W - 1, W, m, C.c_void_p(pref.ctypes.data), 0, C.byref(k), C.byref(off))
            assert k.value == W - 1 and off.value == 0


def test_fp16_rescale_is_an_exact_power_of_two_inside_the_fp16_range(host_misc):
    """pow2_scale (csrc/score_tc.cu): the table-wide scale in front of the fp16 candidate pass is a power of two (the
    rescale is exact in fp32) that maps max|x| into [2^13, 2^14) - far from fp16 overflow (65504) with the headroom of a
    d = 256 dot product in the exponent; degenerate tables (all zero, inf, nan) get scale 1."""
    rng = np.random.default_rng(2)
    vals = np.concatenate([np.float32(2.0) ** rng.integers(-60, 60, 200), rng.standard_normal(300).astype(np.float32) * 7,
                           np.float32([1e-30, 3e30, 65504.0, 1.0, 0.5, 16383.9, 16384.0])])
    for v in np.abs(vals).astype(np.float32):
        if v == 0:
            continue
        s = np.float32(host_misc.host_pow2_scale(int(np.float32(v).view(np.uint32))))
        mant, ex = np.frexp(s)
        assert mant == 0.5                                      # exact power of two
        if 2.0 ** -86 <= float(v) < 2.0 ** 114:                 # the shift is clamped to +-100 outside this range
            assert 2.0 ** 13 <= float(v) * float(s) < 2.0 ** 14
        else:
            assert abs(int(ex) - 1) == 100 and (float(v) * float(s) < 2.0 ** 14 or float(v) >= 2.0 ** 114)
        assert np.float32(v) * s / s == np.float32(v)           # scaling and unscaling is lossless
    for bits in (0, 0x7F800000, 0x7FC00000):                    # 0, +inf, nan
        assert host_misc.host_pow2_scale(bits) == 1.0


def test_tc_visiting_order_is_a_permutation_of_the_catalogue(host_misc):
    """reorder_kernel (csrc/score_tc.cu): head | stratified sample | rest must place every rank of the descending-norm
    order exactly once - with the (H, S, stride) the host derives in score_topk_tc_impl, at the edge sizes (the first
    catalogue that gets a sample, one more, primes, tile multiples +-1) - and keep (item, norm) pairs together; each of
    the three segments stays in descending-norm order."""
    rng = np.random.default_rng(0)
    for tile in (256, 128):
        H, S_full = tile, 16 * tile
        edge = 8 * (H + S_full)
        sizes = [1, 2, tile - 1, tile, tile + 1, edge - 1, edge, edge + 1, edge + S_full - 1, 2 * edge + 17, 100_003, 99_991,
                 3 * S_full * 7 + H, 250_000] + rng.integers(1, 300_000, 6).tolist()
        for n in sizes:
            sample = n >= 8 * (H + S_full)
            S = S_full if sample else 0
            stride = (n - H) // S_full if sample else 1
            perm = rng.permutation(n).astype(np.int32)                               # item id at rank r
            norms = np.sort(rng.random(n).astype(np.float32))[::-1].copy()           # descending
            perm_out = np.full(n, -1, np.int32); norm_out = np.full(n, -1.0, np.float32)
            P = lambda a: C.c_void_p(a.ctypes.data)
            host_misc.host_reorder(P(perm), P(norms.view(np.uint32)), n, H, S, stride, P(perm_out), P(norm_out))
            assert np.array_equal(np.sort(perm_out), np.arange(n)), (tile, n)        # a permutation: nothing lost, nothing twice
            rank_of = np.empty(n, np.int64); rank_of[perm] = np.arange(n)
            assert np.array_equal(norm_out, norms[rank_of[perm_out]])                # norms travel with their items
            r = rank_of[perm_out]                                                     # rank visited at each position
            assert np.array_equal(r[:min(H, n)], np.arange(min(H, n)))                # head: the H highest norms, in order
            if sample:
                assert np.array_equal(r[H:H + S], H + stride * np.arange(S))          # every stride-th rank of the rest
                assert np.all(np.diff(r[H + S:]) > 0)                                 # rest: still descending norm
            else:
                assert np.array_equal(r, np.arange(n))
