"""`-m "not gpu"`: the tensor-core scoring path of csrc/score_tc.cu with everything but the tensor core executed ON THE
HOST from the kernels' own source text (tests/simt_host.py): the pre-pass kernels (row norms and max |x|, the visiting
order head | stratified sample | rest, the exact power-of-two rescale + fp16 rounding, per-tile norm bounds, the mask
filters), the WHOLE epilogue of the N = 256 ping-pong kernel (per-tile threshold tau - e_t, FMNMX max tree, branch-free
append, register bootstrap of tau, warp-cooperative threshold raise by bisection, append budget / overflow hand-off) and
the exact fp32 re-rank.  The one thing replaced is the accumulator: `tcgen05.ld` reads the fp16 x fp16 -> fp32 products
from a numpy GEMM of the very fp16 operands the pre-pass produced, instead of from TMEM.  The host orchestration below
restates score_topk_tc_impl step by step.

What this pins on a box without a GPU is the EXACTNESS ARGUMENT of the path: whatever the approximate pass
keeps, the final top-k must be bit-identical to the exact oracle.  The MMA / TMA / mbarrier choreography, TMEM addressing
and all performance questions are hardware and stay with the `-m gpu` suite (tests/test_gpu_tc.py)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.timeout(1500)
P = lambda a: a.ctypes.data if a is not None else None
KBM, KCAND = 256, 512


@pytest.fixture(scope="module")
def simt(tmp_path_factory):
    from tests.simt_host import build_score, build_tc
    d = str(tmp_path_factory.mktemp("simt_tc"))
    return build_tc(d), build_score(d)


def tc_score_topk(simt, U, V, ld, d, users, ni, mask, k, adversary=None):
    """score_topk_tc_impl (one launch of <= 75,776 rows) on the emulator.  Returns (idx, scores, stats).
    `adversary` (a numpy Generator): every approximate score is moved by a random amount of up to 90 % of the error bound
    e(u,i) = c |u| |v_i| the exactness argument allows the fp16 pass - far more than real rounding ever does."""
    tc, ex = simt
    users = np.ascontiguousarray(users, np.int32)
    nr = len(users)
    up = lambda x, a: (x + a - 1) // a * a
    TILE = 256 if d <= 128 else 128                                 # tc_layout: N = 256 ping-pong kernel for d <= 128, else N = 128
    dpad, items_pad, nr_pad = up(d, 64), up(ni, TILE), up(nr, KBM)
    n_tiles = items_pad // TILE
    # ---- item side, once per call ----
    vnorm = np.zeros(items_pad, np.float32); scal = np.zeros(64, np.uint32)
    tc.emu_row_stats(P(V), ld, d, None, ni, P(vnorm), P(scal[0:1]))
    key16 = vnorm[:ni].view(np.uint32) >> 16                        # cub::DeviceRadixSort::SortPairsDescending, bits [16, 32)
    perm = np.argsort(-key16.astype(np.int64), kind="stable").astype(np.int32)
    vnorm_sorted = np.ascontiguousarray(vnorm[:ni][perm])
    H, S = TILE, 16 * TILE
    sample = ni >= 8 * (H + S)
    stride = (ni - H) // S if sample else 1
    order = np.zeros(items_pad, np.int32); norm_final = np.zeros(items_pad, np.float32)
    tc.emu_reorder(P(perm), P(vnorm_sorted.view(np.uint32)), ni, H, S if sample else 0, stride, P(order), P(norm_final))
    inv_perm = np.zeros(items_pad, np.int32)
    tc.emu_inverse_perm(P(order), ni, P(inv_perm))
    vh = np.zeros((items_pad, dpad), np.float16); scales = np.zeros(2, np.float32)       # [scale_v, scale_u]
    tc.emu_to_f16(P(V), ld, d, None, P(order), ni, items_pad, dpad, P(scal[0:1]), P(vh), P(scales[0:1]))
    tnorm = np.zeros(n_tiles, np.float32)
    tc.emu_tile_norm(P(norm_final), ni, n_tiles, TILE, P(tnorm))
    # ---- user side ----
    unorm = np.zeros(nr_pad, np.float32)
    tc.emu_row_stats(P(U), ld, d, P(users), nr, P(unorm), P(scal[1:2]))
    uh = np.zeros((nr_pad, dpad), np.float16)
    tc.emu_to_f16(P(U), ld, d, P(users), None, nr, nr_pad, dpad, P(scal[1:2]), P(uh), P(scales[1:2]))
    wide = None
    if mask is not None:
        wide = np.zeros((nr_pad, 33), np.uint64)
        tc.emu_bloom(P(users), nr, P(mask[0]), P(mask[1]), P(inv_perm), P(wide))
    # ---- the tensor core's job: fp16 operands, exact products, fp32 accumulation ----
    acc = np.ascontiguousarray(uh.astype(np.float32) @ vh.astype(np.float32).T)          # [nr_pad, items_pad]
    if adversary is not None:
        c = np.float32(0.0009765625 * 1.05 + d * 2.4e-7)                                  # tc_epilogue: cu = c |u| s_u s_v
        bound = c * unorm[:, None] * norm_final[None, :] * scales[1] * scales[0]
        acc = np.ascontiguousarray(acc + adversary.uniform(-0.9, 0.9, acc.shape).astype(np.float32) * bound, np.float32)
    cand = np.zeros((nr_pad, KCAND), np.uint64); cnt = np.full(nr_pad, -7, np.int32)
    tc.emu_tc_epilogue(nr, ni, n_tiles, k, d, P(users), P(mask[0]) if mask else None, P(mask[1]) if mask else None, P(inv_perm),
                       P(wide), P(unorm), P(tnorm), P(scales[1:2]), P(scales[0:1]), P(cand), P(cnt), P(acc), items_pad, 0 if TILE == 256 else 1)
    # ---- exact fp32 re-rank of the candidates; rows on the redo list go to the exact kernel ----
    oi = np.full((nr, k), -5, np.int32); os_ = np.full((nr, k), np.nan, np.float32)
    redo = np.full(nr, -1, np.int32); redo_n = np.zeros(1, np.int32)
    ex.emu_rerank(P(U), P(V), ld, d, P(users), nr, k, P(mask[0]) if mask else None, P(mask[1]) if mask else None, P(order), P(cand),
                  P(cnt), P(oi), P(os_), P(redo), P(redo_n), 0)
    rows = np.sort(redo[:redo_n[0]])
    if len(rows):
        ri = np.zeros((len(rows), k), np.int32); rs = np.zeros((len(rows), k), np.float32)
        ex.emu_score_topk_exact(P(U), P(V), ld, d, P(np.ascontiguousarray(users[rows])), len(rows), ni, P(mask[0]) if mask else None,
                                P(mask[1]) if mask else None, k, P(ri), P(rs), None, 1)
        oi[rows], os_[rows] = ri, rs
    stats = dict(cnt=cnt[:nr].copy(), redo=rows, scales=scales.copy(), sample=sample, n_tiles=n_tiles,
                 pairs=nr * ni, kept=int(np.clip(cnt[:nr], 0, None).sum()))
    return oi, os_, stats


def _tables(rng, nu, ni, d, item_norm_sigma=0.0, user_norm_sigma=0.0, std=0.1):
    ld = (d + 3) // 4 * 4
    U = np.zeros((nu, ld), np.float32); V = np.zeros((ni, ld), np.float32)
    U[:, :d] = rng.standard_normal((nu, d)) * std * np.exp(user_norm_sigma * rng.standard_normal((nu, 1)))
    V[:, :d] = rng.standard_normal((ni, d)) * std * np.exp(item_norm_sigma * rng.standard_normal((ni, 1)))
    return U, V, ld


def _mask(rng, nu, ni, lo, hi):
    rows = [np.sort(rng.choice(ni, int(rng.integers(lo, hi)), replace=False)).astype(np.int32) for _ in range(nu)]
    indptr = np.zeros(nu + 1, np.int64); indptr[1:] = np.cumsum([len(r) for r in rows])
    return indptr, np.concatenate(rows).astype(np.int32)


@pytest.mark.parametrize("d,k,ni,sigma", [(64, 10, 5000, 0.5), (128, 10, 3000, 0.0), (32, 100, 4000, 0.5), (20, 1, 1500, 1.0),
                                          (64, 10, 36000, 0.7)])
def test_tc_path_equals_the_exact_oracle(simt, oracle_c, d, k, ni, sigma):
    """Trained-like tables (log-normal item norms), iso-norm random tables (the adversarial case for a norm-ordered sweep),
    K = 100 (no register bootstrap), K = 1, and a catalogue large enough for the stratified sample: ids AND scores equal the
    C oracle bit for bit, and the approximate pass really pruned (a fraction of the catalogue reaches the re-rank)."""
    rng = np.random.default_rng(d * 7 + k)
    nu = 300
    U, V, ld = _tables(rng, nu, ni, d, item_norm_sigma=sigma, user_norm_sigma=0.3)
    mask = _mask(rng, nu, ni, 0, 40)
    users = rng.permutation(nu)[:270 if ni < 10000 else 150]               # 2 CTAs, the second ragged (1 for the big catalogue)
    idx, sc, st = tc_score_topk(simt, U, V, ld, d, users, ni, mask, k)
    ref_idx, ref_sc = oracle_c.score_topk(U, V, d, users, ni, mask[0], mask[1], k)
    np.testing.assert_array_equal(idx, ref_idx)
    np.testing.assert_array_equal(sc, ref_sc)
    assert st["sample"] == (ni >= 8 * (256 + 4096))                         # d <= 128 here: tile 256
    done = st["cnt"] >= 0
    assert (st["cnt"][done] >= k).all() and st["cnt"].max() < 512          # lists hold at least k, never overflow the slots
    assert st["kept"] < 0.25 * st["pairs"]                                 # the fp16 pass pruned the catalogue
    mant, _ = np.frexp(st["scales"])
    assert (mant == 0.5).all()                                             # both rescales are exact powers of two


def test_tc_path_hard_rows_are_handed_to_the_exact_kernel_and_still_exact(simt, oracle_c):
    """Users anti-aligned with the popularity direction (scores RISE along the norm-ordered sweep), quantised tables (ties
    everywhere), users who own most of a small catalogue (fewer than k unmasked items): whatever the candidate pass does with
    such rows - long lists, raises, overflow to the redo list - the result stays the exact oracle's."""
    rng = np.random.default_rng(11)
    nu, ni, d, k = 256, 2600, 32, 10
    ld = 32
    pop = np.zeros(d, np.float32); pop[0] = 1.0
    V = np.zeros((ni, ld), np.float32); U = np.zeros((nu, ld), np.float32)
    norms = np.exp(0.8 * rng.standard_normal(ni)).astype(np.float32)
    V[:, :d] = rng.standard_normal((ni, d)) * 0.05 + norms[:, None] * pop   # item norm ~ popularity direction
    U[:, :d] = rng.standard_normal((nu, d)) * 0.05
    U[:64, 0] -= 1.0                                                        # anti-aligned: best items have the SMALLEST norms
    U[64:128, 0] += 1.0                                                     # aligned: top-k found in the head tile
    U[128:192] = np.round(U[128:192] * 8) / 8                               # quantised users against
    V[::3] = np.round(V[::3] * 4) / 4                                       # quantised items: many exactly equal scores
    rows = []
    for u in range(nu):
        n_own = ni - int(rng.integers(3, 9)) if u >= 248 else int(rng.integers(0, 60))   # the last 8 users own nearly everything
        rows.append(np.sort(rng.choice(ni, n_own, replace=False)).astype(np.int32))
    indptr = np.zeros(nu + 1, np.int64); indptr[1:] = np.cumsum([len(r) for r in rows])
    mask = (indptr, np.concatenate(rows))
    users = np.arange(nu, dtype=np.int32)
    idx, sc, st = tc_score_topk(simt, U, V, ld, d, users, ni, mask, k)
    ref_idx, ref_sc = oracle_c.score_topk(U, V, d, users, ni, mask[0], mask[1], k)
    np.testing.assert_array_equal(idx, ref_idx)
    np.testing.assert_array_equal(sc, ref_sc)
    redo = set(st["redo"].tolist())
    assert set(range(248, 256)) <= redo                                    # fewer than k unmasked items: exact kernel
    # anti-aligned users beat their own threshold along the whole sweep (this catalogue is too small for the stratified
    # sample that gives them a threshold early): the append budget hands them over instead of dragging their warp along
    assert len(redo & set(range(0, 64))) >= 32
    assert not redo & set(range(64, 128))                                  # aligned users: done in the candidate pass
    assert len(redo & set(range(192, 248))) <= 2                           # ordinary users too


def test_tc_path_without_a_mask(simt, oracle_c):
    rng = np.random.default_rng(2)
    nu, ni, d, k = 130, 2100, 48, 20
    U, V, ld = _tables(rng, nu, ni, d, item_norm_sigma=0.6)
    users = np.arange(nu, dtype=np.int32)
    idx, sc, st = tc_score_topk(simt, U, V, ld, d, users, ni, None, k)
    ref_idx, ref_sc = oracle_c.score_topk(U, V, d, users, ni, None, None, k)
    np.testing.assert_array_equal(idx, ref_idx)
    np.testing.assert_array_equal(sc, ref_sc)
    assert len(st["redo"]) == 0


@pytest.mark.parametrize("k", [10, 100])
def test_tc_path_stays_exact_under_an_adversarial_accumulator(simt, oracle_c, k):
    """Soundness of the pruning itself: move EVERY approximate score by a random amount of up to 0.9 e(u,i) - the bound the
    thresholds are built on (tau - e_t, the raise's lower / upper bounds, the bootstrap) - three orders of magnitude more
    than fp16 rounding of these tables produces.  The final top-k is still the exact oracle's, bit for bit."""
    rng = np.random.default_rng(k)
    nu, ni, d = 256, 4000, 64
    U, V, ld = _tables(rng, nu, ni, d, item_norm_sigma=0.5, user_norm_sigma=0.3)
    mask = _mask(rng, nu, ni, 0, 40)
    users = np.arange(nu, dtype=np.int32)
    idx, sc, st = tc_score_topk(simt, U, V, ld, d, users, ni, mask, k, adversary=np.random.default_rng(99))
    ref_idx, ref_sc = oracle_c.score_topk(U, V, d, users, ni, mask[0], mask[1], k)
    np.testing.assert_array_equal(idx, ref_idx)
    np.testing.assert_array_equal(sc, ref_sc)
    assert (st["cnt"][st["cnt"] >= 0] >= k).all()


def test_tc_path_exact_across_twelve_octaves_of_row_norms(simt, oracle_c):
    """One power-of-two scale serves the whole table, so small rows use the low end of fp16's range.  With item AND user
    norms spread log-uniformly over 2^-12 .. 1 of the largest (every element still an fp16 NORMAL number after the rescale)
    the relative bound e = c |u| |v| holds and the result is the exact oracle's - including users / items whose scores are
    4096 times smaller than their neighbours'.  (Below ~2^-27 of the table's largest element a non-zero row turns
    fp16-subnormal and the bound no longer covers it: DESIGN, known limits.)"""
    rng = np.random.default_rng(12)
    nu, ni, d, k = 256, 4000, 64, 10
    ld = 64
    U = (rng.standard_normal((nu, d)) * 0.1).astype(np.float32) * np.exp2(-12 * rng.random((nu, 1))).astype(np.float32)
    V = (rng.standard_normal((ni, d)) * 0.1).astype(np.float32) * np.exp2(-12 * rng.random((ni, 1))).astype(np.float32)
    mask = _mask(rng, nu, ni, 0, 30)
    users = np.arange(nu, dtype=np.int32)
    idx, sc, st = tc_score_topk(simt, U, V, ld, d, users, ni, mask, k)
    ref_idx, ref_sc = oracle_c.score_topk(U, V, d, users, ni, mask[0], mask[1], k)
    np.testing.assert_array_equal(idx, ref_idx)
    np.testing.assert_array_equal(sc, ref_sc)
    small = np.linalg.norm(U, axis=1) < 2.0 ** -8 * np.linalg.norm(U, axis=1).max()
    assert small.sum() > 20 and (st["cnt"][small] >= k).all()                # small users were served by the fp16 pass too


@pytest.mark.parametrize("d,k,ni", [(200, 10, 3000), (256, 100, 2500), (136, 1, 17500)])
def test_tc_path_wide_rows_n128_kernel(simt, oracle_c, d, k, ni):
    """d > 128 takes the N = 128 kernel (tc_candidate_kernel): 128-item tiles, a pair of accumulators per stage and two
    stages - same epilogue template with PP = false.  (17,500 items: the first size with a stratified sample at this tile.)"""
    rng = np.random.default_rng(d + k)
    nu = 200
    U, V, ld = _tables(rng, nu, ni, d, item_norm_sigma=0.5, user_norm_sigma=0.3)
    mask = _mask(rng, nu, ni, 0, 40)
    users = rng.permutation(nu)[:170]
    idx, sc, st = tc_score_topk(simt, U, V, ld, d, users, ni, mask, k)
    ref_idx, ref_sc = oracle_c.score_topk(U, V, d, users, ni, mask[0], mask[1], k)
    np.testing.assert_array_equal(idx, ref_idx)
    np.testing.assert_array_equal(sc, ref_sc)
    assert st["n_tiles"] == -(-ni // 128) and st["sample"] == (ni >= 8 * (128 + 2048))
    assert st["kept"] < 0.3 * st["pairs"]
