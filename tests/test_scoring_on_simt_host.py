"""`-m "not gpu"`: the exact scoring + masked top-K kernels (csrc/score_exact.cu, csrc/topk_list.cuh) - the exactness
anchor of the whole scoring path, also the fallback and re-rank reference of the tensor-core kernel - executed ON THE HOST
by the SIMT emulator of tests/simt_host.py from their own source text.  Expectation, as on the device
(tests/test_gpu_parity.py::test_score_topk_exact_bitwise_vs_oracle): ids AND scores bit-identical to the C oracle
(oracle/eval_oracle.c: k-ascending fp32 FMA chain, ties by item id ascending).  The tcgen05 candidate pass cannot be
emulated and stays with the `-m gpu` suite."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.timeout(1200)


@pytest.fixture(scope="module")
def simt(tmp_path_factory):
    from tests.simt_host import build_score
    return build_score(str(tmp_path_factory.mktemp("simt_score")))


def _tables(rng, nu, ni, d, std=1.0):
    ld = (d + 3) // 4 * 4
    U = np.zeros((nu, ld), np.float32); V = np.zeros((ni, ld), np.float32)
    U[:, :d] = rng.standard_normal((nu, d)) * std; V[:, :d] = rng.standard_normal((ni, d)) * std
    return U, V, ld


def _mask(rng, nu, ni, lo, hi):
    rows = [np.sort(rng.choice(ni, int(rng.integers(lo, hi)), replace=False)).astype(np.int32) for _ in range(nu)]
    indptr = np.zeros(nu + 1, np.int64); indptr[1:] = np.cumsum([len(r) for r in rows])
    return indptr, np.concatenate(rows).astype(np.int32)


def _run(simt, U, V, ld, d, users, ni, mask, k, splits=1, dense=False):
    P = lambda a: a.ctypes.data if a is not None else None
    users = np.ascontiguousarray(users, np.int32)
    n = len(users)
    idx = np.full((n, max(k, 1)), -5, np.int32); sc = np.full((n, max(k, 1)), np.nan, np.float32)
    dn = np.full((n, ni), np.nan, np.float32) if dense else None
    used = simt.emu_score_topk_exact(P(U), P(V), ld, d, P(users), n, ni, P(mask[0]) if mask else None,
                                     P(mask[1]) if mask else None, k, P(idx), P(sc), P(dn), splits)
    return idx, sc, dn, used


@pytest.mark.parametrize("d,k,nu,ni,splits", [(32, 10, 70, 333, 1), (128, 5, 64, 200, 1), (50, 100, 20, 450, 1),
                                              (7, 1, 9, 65, 1), (64, 10, 40, 700, 4), (20, 300, 18, 420, 1),
                                              (20, 300, 18, 420, 2)])
def test_exact_scoring_kernel_is_bitwise_the_oracle(simt, oracle_c, d, k, nu, ni, splits):
    rng = np.random.default_rng(d * 1000 + k)
    U, V, ld = _tables(rng, nu, ni, d)
    mask = _mask(rng, nu, ni, 0, 30)
    users = rng.permutation(nu)[: nu - 3]                                   # a ragged, permuted subset of the rows
    idx, sc, _, used = _run(simt, U, V, ld, d, users, ni, mask, k, splits)
    assert used == (splits if splits > 1 else 1)
    ref_idx, ref_sc = oracle_c.score_topk(U, V, d, users, ni, mask[0], mask[1], k)
    np.testing.assert_array_equal(idx, ref_idx)
    np.testing.assert_array_equal(sc, ref_sc)                               # the same fp32 FMA chain: bit for bit
    for r, u in enumerate(users[::7]):                                      # masked items never appear
        assert not np.intersect1d(idx[7 * r], mask[1][mask[0][u]:mask[0][u + 1]]).size


def test_exact_scoring_ties_and_rows_with_fewer_unmasked_items_than_k(simt, oracle_c):
    """Quantised tables (many equal scores: order must be id ascending) and users who own nearly the whole catalogue (the
    tail of the list is then filled with masked -inf items in id order, exactly like the oracle)."""
    rng = np.random.default_rng(4)
    nu, ni, d, k = 12, 90, 8, 20
    ld = 8
    U = np.round(rng.standard_normal((nu, ld))).astype(np.float32); V = np.round(rng.standard_normal((ni, ld))).astype(np.float32)
    rows = [np.sort(rng.choice(ni, ni - int(rng.integers(0, 25)), replace=False)).astype(np.int32) for _ in range(nu)]
    indptr = np.zeros(nu + 1, np.int64); indptr[1:] = np.cumsum([len(r) for r in rows])
    mask = (indptr, np.concatenate(rows))
    idx, sc, _, _ = _run(simt, U, V, ld, d, np.arange(nu), ni, mask, k)
    ref_idx, ref_sc = oracle_c.score_topk(U, V, d, np.arange(nu), ni, mask[0], mask[1], k)
    np.testing.assert_array_equal(idx, ref_idx)
    np.testing.assert_array_equal(sc, ref_sc)
    assert np.isinf(sc).any()                                               # the fewer-than-k case is exercised
    nomask_idx, nomask_sc, _, _ = _run(simt, U, V, ld, d, np.arange(nu), ni, None, k)
    r2_idx, r2_sc = oracle_c.score_topk(U, V, d, np.arange(nu), ni, None, None, k)
    np.testing.assert_array_equal(nomask_idx, r2_idx)
    same = nomask_sc[:, 1:] == nomask_sc[:, :-1]
    assert same.any() and np.all(nomask_idx[:, 1:][same] > nomask_idx[:, :-1][same])      # ties: smaller id first


def test_predict_dense_contract(simt):
    """k = 0 + dense output = models/MF.py:109-132: the [n, I] block of scores with -inf at the user's train positives."""
    rng = np.random.default_rng(8)
    nu, ni, d = 30, 150, 24
    U, V, ld = _tables(rng, nu, ni, d)
    mask = _mask(rng, nu, ni, 1, 20)
    users = np.array([3, 3, 17, 0, 29], np.int32)                           # duplicates allowed
    _, _, dense, _ = _run(simt, U, V, ld, d, users, ni, mask, 0, dense=True)
    want = np.zeros((len(users), ni), np.float32)
    for r, u in enumerate(users):
        acc = np.zeros(ni, np.float32)
        for kk in range(d):                                                 # k-ascending fp32 FMA chain of the oracle
            acc = np.float32(np.float64(U[u, kk]) * V[:, kk].astype(np.float64) + acc.astype(np.float64))
        want[r] = acc
        want[r, mask[1][mask[0][u]:mask[0][u + 1]]] = -np.inf
    np.testing.assert_allclose(dense, want, rtol=1e-6, atol=1e-6)
    assert np.array_equal(np.isinf(dense), np.isinf(want))


def test_topk_rows_kernel_vs_oracle(simt, oracle_c):
    """Device-resident c_top_k_array_index (func.h:22-31): one warp per row, ties by id, -inf entries, k up to the width."""
    rng = np.random.default_rng(6)
    S = np.round(rng.standard_normal((11, 170)) * 4).astype(np.float32) + np.float32(0.0)   # + 0.0: no -0.0 (DESIGN, known limits)
    S[:, rng.choice(170, 20, replace=False)] = -np.inf
    for k in (1, 10, 33, 170):
        out = np.zeros((11, k), np.int32)
        simt.emu_topk_rows(S.ctypes.data, 170, 11, 170, k, out.ctypes.data)
        np.testing.assert_array_equal(out, oracle_c.topk(S, k))


# ---- the CUDA-core half of the tensor-core path (csrc/score_tc.cu): exact re-rank of candidate lists, mask filters --------
@pytest.mark.parametrize("staged", [0, 1])
@pytest.mark.parametrize("d", [20, 128, 200])
def test_rerank_kernels_turn_candidate_lists_into_the_exact_topk(simt, oracle_c, d, staged):
    """rerank_kernel / rerank_staged_kernel: given candidate lists that CONTAIN the exact top-k (what the tcgen05 pass
    guarantees) plus anything else - masked items, duplicates of the approximate-score word, the maybe-masked flag bit -
    the output is the exact masked top-k, bit-identical to the C oracle; rows flagged cnt = -1 and rows with fewer than k
    unmasked candidates go to the redo list (the exact kernel finishes them)."""
    rng = np.random.default_rng(d + staged)
    nu, ni, k, n_rows = 60, 500, 10, 37
    U, V, ld = _tables(rng, nu, ni, d)
    mask = _mask(rng, nu, ni, 0, 40)
    users = rng.permutation(nu)[:n_rows].astype(np.int32)
    order = rng.permutation(ni).astype(np.int32)                            # item at visiting position p
    pos_of = np.empty(ni, np.int64); pos_of[order] = np.arange(ni)
    ref_idx, ref_sc = oracle_c.score_topk(U, V, d, users, ni, mask[0], mask[1], k)
    S = U[users, :d].astype(np.float64) @ V[:, :d].astype(np.float64).T
    cand = np.zeros((n_rows, 512), np.uint64); cnt = np.zeros(n_rows, np.int32)
    starved = 5                                                             # this row gets fewer than k unmasked candidates
    for r in range(n_rows):
        best = np.argsort(-S[r])[: k + 25]                                  # incl. masked ones: the re-rank must drop them
        extra = rng.choice(ni, 120, replace=False)
        items = np.unique(np.concatenate([best, extra]))
        if r == starved:
            m = mask[1][mask[0][users[r]]:mask[0][users[r] + 1]]
            items = np.concatenate([m, np.setdiff1d(np.arange(ni), m)[: k - 3]])
        rng.shuffle(items)
        words = pos_of[items].astype(np.uint64) | (rng.integers(0, 2, len(items)).astype(np.uint64) << np.uint64(31)) \
            | (rng.integers(0, 2**32, len(items)).astype(np.uint64) << np.uint64(32))
        cand[r, :len(items)] = words; cnt[r] = len(items)
    overflow = [3, 20]
    cnt[overflow] = -1
    oi = np.full((n_rows, k), -5, np.int32); os_ = np.full((n_rows, k), np.nan, np.float32)
    redo = np.full(n_rows, -1, np.int32); redo_n = np.zeros(1, np.int32)
    P = lambda a: a.ctypes.data
    simt.emu_rerank(P(U), P(V), ld, d, P(users), n_rows, k, P(mask[0]), P(mask[1]), P(order), P(cand), P(cnt), P(oi), P(os_),
                    P(redo), P(redo_n), staged)
    assert sorted(redo[:redo_n[0]].tolist()) == sorted(overflow + [starved])
    done = np.setdiff1d(np.arange(n_rows), overflow + [starved])
    np.testing.assert_array_equal(oi[done], ref_idx[done])
    np.testing.assert_array_equal(os_[done], ref_sc[done])                  # same k-ascending fp32 FMA chain: bit for bit
    assert (oi[overflow + [starved]] == -5).all()                           # redo rows are left to the exact kernel


def test_mask_filters_have_no_false_negatives(simt):
    """bloom_kernel: per scored row an EXACT bitmap of visiting positions 0..63 and a 2048-bit filter over all positions
    of the row's masked items - the candidate pass may call an item 'certainly unmasked' only when its bit is clear, so a
    set of masked positions must never read clear (no false negatives); the bitmap has no false positives either."""
    rng = np.random.default_rng(3)
    nu, ni, n_rows = 50, 3000, 21
    mask = _mask(rng, nu, ni, 0, 200)
    users = rng.permutation(nu)[:n_rows].astype(np.int32)
    inv_perm = rng.permutation(ni).astype(np.int32)                         # item id -> visiting position
    wide = np.full((n_rows, 33), 0xFFFFFFFFFFFFFFFF, np.uint64)             # the kernel must clear its rows first
    P = lambda a: a.ctypes.data
    simt.emu_bloom(P(users), n_rows, P(mask[0]), P(mask[1]), P(inv_perm), P(wide))
    for r, u in enumerate(users):
        pos = inv_perm[mask[1][mask[0][u]:mask[0][u + 1]]].astype(np.uint32)
        exact = 0
        for p_ in pos[pos < 64]:
            exact |= 1 << int(p_)
        assert int(wide[r, 0]) == exact                                     # exact head bitmap
        h = ((pos.astype(np.uint64) * np.uint64(2654435761)) & np.uint64(0xFFFFFFFF)) >> np.uint64(21)
        for hh in h:
            assert (int(wide[r, 1 + (int(hh) >> 6)]) >> (int(hh) & 63)) & 1  # every masked position is flagged
        bits = sum(bin(int(w)).count("1") for w in wide[r, 1:])
        assert bits <= len(pos)                                             # and nothing else was set
