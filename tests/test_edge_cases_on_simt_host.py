"""`-m "not gpu"`: edge cases of the kernels on the SIMT emulator (tests/simt_host.py) - the shapes a real run rarely
produces and the device suite does not enumerate: batches smaller than a warp, users without positives or owning the
whole catalogue inside the fused sampling kernels, ranks that receive no triples, catalogues smaller than one tile,
k at its maximum, empty candidate lists, matrices without nonzeros."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import bpr_oracle as O
from recsys_pytorch_b200._lib import F_ITEM_DELTA, F_USERS_UNIQUE, SINK_UPDATE
from tests.test_kernels_on_simt_host import FAST, GROUP8, GROUP16, LDG, Step, _problem

pytestmark = pytest.mark.timeout(1500)
P = lambda a: a.ctypes.data if a is not None else None


@pytest.fixture(scope="module")
def simt(tmp_path_factory):
    from tests.simt_host import build
    return build(str(tmp_path_factory.mktemp("simt_edge")))


@pytest.mark.parametrize("kind", [LDG, FAST, GROUP8, GROUP16])
@pytest.mark.parametrize("B", [1, 3, 31, 33])
def test_batches_around_the_warp_size(simt, kind, B):
    U0, V0, u, i, j = _problem(B, 64, 200, 128, B, std=0.3, unique_users=True, unique_items=True)
    s = Step(simt, U0, V0, 128, u, i, j, lr=0.5, reg=0.01, flags=F_USERS_UNIQUE, kind=kind, chunk=0 if kind != LDG else 4)
    Ur, Vr, lref = O.sgd_step(U0, V0, u, i, j, 0.5, 0.01)
    np.testing.assert_allclose(s.U, Ur, rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(s.V, Vr, rtol=2e-5, atol=2e-6)
    assert abs(s.loss[0] / B - float(lref)) < 2e-5


@pytest.mark.parametrize("kind", [LDG, FAST, GROUP8, GROUP16])
def test_sampling_kernels_skip_users_without_a_valid_triple(simt, kind):
    """In-kernel sampling: a user without positives emits no triple, a user who owns the whole catalogue is skipped after
    64 rejected draws - their rows must not move, they are reported as -1, the loss counts only real triples, and the
    rest of the batch is the oracle step on the triples the host mirror predicts."""
    rng = np.random.default_rng(kind)
    nu, ni, d = 90, 70, 128
    rows = [np.sort(rng.choice(ni, int(rng.integers(1, 20)), replace=False)).astype(np.int32) for _ in range(nu)]
    for u in range(0, nu, 7):
        rows[u] = np.zeros(0, np.int32)                                      # no positives
    for u in range(3, nu, 11):
        rows[u] = np.arange(ni, dtype=np.int32)                              # owns everything: no negative exists
    indptr = np.zeros(nu + 1, np.int64); indptr[1:] = np.cumsum([len(r) for r in rows])
    indices = np.concatenate(rows)
    U0 = (rng.standard_normal((nu, d)) * 0.3).astype(np.float32); V0 = (rng.standard_normal((ni, d)) * 0.3).astype(np.float32)
    users = rng.permutation(nu)
    s = Step(simt, U0, V0, d, users, csr=(indptr, indices), lr=1.0, reg=0.0, flags=F_USERS_UNIQUE | F_ITEM_DELTA, seed=5, step=2,
             kind=kind, inv_batch=1.0 / nu)
    dead = np.array([len(rows[u]) in (0, ni) for u in users])
    assert dead.sum() >= 15
    assert (s.out_pos[dead] == -1).all() and (s.out_neg[dead] == -1).all()
    live = ~dead
    drawn = [O.sample_triple(5, 2, int(t), int(users[t]), indptr, indices, ni) for t in np.flatnonzero(live)]
    pos, neg = np.array([p_ for p_, _ in drawn]), np.array([n_ for _, n_ in drawn])
    assert np.array_equal(s.out_pos[live], pos) and np.array_equal(s.out_neg[live], neg)
    assert np.array_equal(s.U[users[dead]], U0[users[dead]])                 # untouched rows
    dU, dV, _, _ = O.bpr_grads(U0, V0, users[live], pos, neg)                # scaled by 1/len(live): rescale to 1/nu
    f = np.float32(live.sum() / nu)
    np.testing.assert_allclose(s.U, U0 - f * dU, rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(V0 + s.gV, V0 - f * dV, rtol=2e-5, atol=2e-6)
    lref, _ = O.bpr_loss(U0, V0, users[live], pos, neg)
    assert abs(s.loss[0] / live.sum() - float(lref)) < 2e-5


# ---- P2P: ranks without work ------------------------------------------------------------------------------------------------
def test_p2p_ranks_that_receive_nothing(tmp_path_factory):
    """Only one rank routes triples, and all of them to one owner: every other rank launches its step over empty outbox
    segments (n_chunks = 0), the owner pulls from a single source - both visiting orders."""
    from tests.simt_host import build_p2p
    from tests.test_p2p_on_simt_host import _make, _tables
    from recsys_pytorch_b200._lib import F_P2P_ROUND_ROBIN
    simt = build_p2p(str(tmp_path_factory.mktemp("simt_p2p_edge")))
    for flags in (F_USERS_UNIQUE, F_USERS_UNIQUE | F_P2P_ROUND_ROBIN):
        W, nu, ni, d = 4, 128, 400, 128
        ranks, U0, V0, _, _, ib, ub = _make(W, nu, ni, d, seed=3)
        rng = np.random.default_rng(0)
        B = 20
        ul = rng.permutation(ub[1] - ub[0])[:B].astype(np.int32)               # rank 0's users
        items = rng.permutation(np.arange(ib[2], ib[3]))                        # every positive AND negative owned by rank 2
        pi, pj = items[:B].astype(np.int32), items[B:2 * B].astype(np.int32)
        ranks[0].route(simt, ul, 1, pos=pi, neg=pj)
        for r in ranks[1:]:
            r.route(simt, np.zeros(0, np.int32), 1, pos=np.zeros(0, np.int32), neg=np.zeros(0, np.int32))
        for r in ranks:
            r.step(simt, ranks, B, 16, flags=flags)
        assert [int(r.n_processed[0]) for r in ranks] == [0, 0, B, 0]
        Ur, Vr, _ = O.sgd_step(U0, V0, ul + ub[0], pi, pj, 0.9, 0.01)
        U, V = _tables(ranks, d)
        np.testing.assert_allclose(U, Ur, rtol=2e-5, atol=2e-6)
        np.testing.assert_allclose(V, Vr, rtol=2e-5, atol=2e-6)


# ---- scoring ----------------------------------------------------------------------------------------------------------------
def test_exact_scoring_tiny_shapes(tmp_path_factory, oracle_c):
    """One user, d = 1, catalogues smaller than one tile, k equal to the catalogue size."""
    from tests.simt_host import build_score
    simt = build_score(str(tmp_path_factory.mktemp("simt_score_edge")))
    rng = np.random.default_rng(1)
    for nu, ni, d, k in ((1, 5, 1, 5), (3, 63, 3, 63), (2, 64, 5, 1), (65, 65, 2, 7), (1, 300, 40, 256)):
        ld = (d + 3) // 4 * 4
        U = np.zeros((nu, ld), np.float32); V = np.zeros((ni, ld), np.float32)
        U[:, :d] = rng.standard_normal((nu, d)); V[:, :d] = rng.standard_normal((ni, d))
        users = np.arange(nu, dtype=np.int32)
        idx = np.zeros((nu, k), np.int32); sc = np.zeros((nu, k), np.float32)
        simt.emu_score_topk_exact(P(U), P(V), ld, d, P(users), nu, ni, None, None, k, P(idx), P(sc), None, 1)
        ref_idx, ref_sc = oracle_c.score_topk(U, V, d, users, ni, None, None, k)
        np.testing.assert_array_equal(idx, ref_idx)
        np.testing.assert_array_equal(sc, ref_sc)


def test_tc_path_tiny_catalogue_and_largest_k(tmp_path_factory, oracle_c):
    """A catalogue smaller than the 64-item bootstrap chunk, one barely larger than a tile, and k = 128 (the largest the
    tensor-core path accepts: kCand / 4): still the exact oracle's answer."""
    from tests.simt_host import build_score, build_tc
    from tests.test_tc_path_on_simt_host import _mask, _tables, tc_score_topk
    dd = str(tmp_path_factory.mktemp("simt_tc_edge"))
    simt = (build_tc(dd), build_score(dd))
    rng = np.random.default_rng(4)
    for ni, d, k, nu in ((40, 16, 10, 70), (300, 64, 128, 40), (257, 8, 32, 33), (1000, 128, 33, 20)):
        U, V, ld = _tables(rng, nu, ni, d, item_norm_sigma=0.4)
        mask = _mask(rng, nu, ni, 0, min(20, ni // 3))
        users = np.arange(nu, dtype=np.int32)
        idx, sc, st = tc_score_topk(simt, U, V, ld, d, users, ni, mask, k)
        ref_idx, ref_sc = oracle_c.score_topk(U, V, d, users, ni, mask[0], mask[1], k)
        np.testing.assert_array_equal(idx, ref_idx)
        np.testing.assert_array_equal(sc, ref_sc)


def test_rerank_with_empty_and_short_candidate_lists(tmp_path_factory):
    from tests.simt_host import build_score
    simt = build_score(str(tmp_path_factory.mktemp("simt_rr_edge")))
    nu, ni, d, k = 9, 50, 8, 5
    rng = np.random.default_rng(0)
    U = rng.standard_normal((nu, d)).astype(np.float32); V = rng.standard_normal((ni, d)).astype(np.float32)
    order = np.arange(ni, dtype=np.int32)
    cand = np.zeros((nu, 512), np.uint64); cnt = np.zeros(nu, np.int32)
    cand[1, :3] = [4, 9, 11]; cnt[1] = 3                                       # fewer candidates than k
    cand[2, :5] = [0, 1, 2, 3, 4]; cnt[2] = 5                                  # exactly k
    oi = np.full((nu, k), -5, np.int32); os_ = np.zeros((nu, k), np.float32)
    redo = np.full(nu, -1, np.int32); redo_n = np.zeros(1, np.int32)
    users = np.arange(nu, dtype=np.int32)
    simt.emu_rerank(P(U), P(V), d, d, P(users), nu, k, None, None, P(order), P(cand), P(cnt), P(oi), P(os_), P(redo), P(redo_n), 0)
    assert sorted(redo[:redo_n[0]].tolist()) == [0, 1, 3, 4, 5, 6, 7, 8]      # everything but the row with k candidates
    s = U[2] @ V[:5].T
    assert oi[2].tolist() == np.argsort(-s, kind="stable").tolist()


def test_spmm_degenerate_matrices(tmp_path_factory):
    from tests.simt_host import build_spmm
    from tests.test_spmm_on_simt_host import _spmm
    simt = build_spmm(str(tmp_path_factory.mktemp("simt_spmm_edge")))
    rng = np.random.default_rng(0)
    for n, d in ((1, 8), (3, 64), (40, 16)):
        X = rng.standard_normal((n, d)).astype(np.float32)
        Z = sp.csr_matrix((n, n), dtype=np.float32)                            # no nonzeros at all
        Y = np.full((n, d), 3.0, np.float32); acc = np.ones((n, d), np.float32)
        _spmm(simt, Z, X, d, Y=Y, acc=acc, acc_scale=0.5, acc_init=1)
        assert not Y.any()
        np.testing.assert_allclose(acc, 0.5 * X, rtol=1e-6)                    # acc_init: acc = s (X + A X)
        I = sp.identity(n, dtype=np.float32, format="csr")
        Y2 = np.zeros((n, d), np.float32)
        _spmm(simt, I, X, d, Y=Y2)
        assert np.array_equal(Y2, X)
