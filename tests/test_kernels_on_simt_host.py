"""`-m "not gpu"`: the fused BPR step kernels of csrc/bpr_step.cu executed ON THE HOST by the SIMT emulator of
tests/simt_host.py (the kernels' own source text; lanes are threads, warp intrinsics are rendezvous) against the numpy
oracle.  Same comparisons and tolerances as tests/test_gpu_parity.py makes on the device: rtol 2e-5 for fp32 training
arithmetic (summation order), bit-exact ids.  What the emulator cannot tell - the hardware memory model, the TMA / tcgen05
paths, performance - stays with the `-m gpu` suite."""
import ctypes as C

import numpy as np
import pytest

from oracle import bpr_oracle as O
from recsys_pytorch_b200 import _lib
from recsys_pytorch_b200._lib import F_ITEM_DELTA, F_USERS_UNIQUE, SINK_GRAD, SINK_NONE, SINK_STAGE, SINK_UPDATE

LDG, FAST, GROUP8, GROUP16 = 0, 1, 2, 3
pytestmark = pytest.mark.timeout(900)          # a lost rendezvous in the emulator must fail, not hang


@pytest.fixture(scope="module")
def simt(tmp_path_factory):
    from tests.simt_host import build
    return build(str(tmp_path_factory.mktemp("simt")))


def _pad(W):
    ld = (W.shape[1] + 3) // 4 * 4
    out = np.zeros((W.shape[0], ld), np.float32)
    out[:, :W.shape[1]] = W
    return out


class Step:
    """One emulated launch: keeps every buffer alive and exposes the results."""

    def __init__(self, simt, U, V, d, users, pos=None, neg=None, csr=None, lr=0.0, reg=0.0, sink=SINK_UPDATE, flags=0,
                 seed=0, step=0, want_loss=True, want_x=False, inv_batch=0.0, kind=LDG, chunk=0, grid=3):
        self.U, self.V = _pad(U), _pad(V)
        ld, B = self.U.shape[1], len(users)
        a = _lib.BprArgs()
        self.keep = [np.ascontiguousarray(users, np.int32)]
        a.U, a.V, a.ld, a.d = self.U.ctypes.data, self.V.ctypes.data, ld, d
        a.num_users, a.num_items, a.users, a.B = U.shape[0], V.shape[0], self.keep[0].ctypes.data, B
        for name, arr in (("pos", pos), ("neg", neg)):
            if arr is not None:
                self.keep.append(np.ascontiguousarray(arr, np.int32)); setattr(a, name, self.keep[-1].ctypes.data)
        if csr is not None:
            self.keep += [np.ascontiguousarray(csr[0], np.int64), np.ascontiguousarray(csr[1], np.int32)]
            a.csr_indptr, a.csr_indices = self.keep[-2].ctypes.data, self.keep[-1].ctypes.data
        a.seed, a.step, a.lr, a.reg, a.sink, a.flags, a.inv_batch = seed, step, lr, reg, sink, flags, inv_batch
        self.out_pos, self.out_neg = np.full(B, -9, np.int32), np.full(B, -9, np.int32)
        if pos is None or neg is None:
            a.out_pos, a.out_neg = self.out_pos.ctypes.data, self.out_neg.ctypes.data
        self.loss = np.zeros(1, np.float64)
        if want_loss:
            a.loss_sum = self.loss.ctypes.data
        self.x = np.zeros(B, np.float32)
        if want_x:
            a.x_out = self.x.ctypes.data
        self.stage = np.zeros((B, 3, ld), np.float32)
        self.gU, self.gV = np.zeros_like(self.U), np.zeros_like(self.V)
        if sink == SINK_STAGE:
            a.stage = self.stage.ctypes.data
        if sink == SINK_GRAD or (flags & F_ITEM_DELTA):
            a.gU, a.gV = self.gU.ctypes.data, self.gV.ctypes.data
        assert simt.emu_bpr_step(C.addressof(a), kind, chunk, grid) == 0
        self.a = a

    def apply(self, simt, grid=2):
        P = lambda x: x.ctypes.data
        simt.emu_bpr_apply(P(self.U), P(self.V), self.U.shape[1], self.a.users, self.a.pos, self.a.neg, self.a.B,
                           P(self.stage), grid)


def _problem(seed, nu, ni, d, B, std=0.5, unique_users=False, unique_items=False):
    rng = np.random.default_rng(seed)
    U0 = (rng.standard_normal((nu, d)) * std).astype(np.float32)
    V0 = (rng.standard_normal((ni, d)) * std).astype(np.float32)
    u = rng.permutation(nu)[:B] if unique_users else rng.integers(0, nu, B)
    if unique_items:
        items = rng.permutation(ni)[:2 * B]; i, j = items[:B], items[B:]
    else:
        i, j = rng.integers(0, ni, B), rng.integers(0, ni, B)
    return U0, V0, u, i, j


# ---- the generic kernel: every row width class, every sink, ragged batch, duplicates --------------------------------------
@pytest.mark.parametrize("d,chunk", [(4, 32), (8, 8), (20, 32), (32, 4), (50, 16), (64, 32), (100, 8), (128, 4), (128, 32),
                                     (200, 32), (256, 4), (400, 32)])
def test_ldg_kernel_exact_step_all_widths(simt, d, chunk):
    """tests/test_gpu_parity.py::test_exact_step_all_widths on the emulator: SINK_STAGE + bpr_apply == oracle.sgd_step
    (all gradients from pre-step weights), loss, score differences; ragged B, duplicate users and items, pad columns 0."""
    nu, ni, B = 97, 61, 300 + d % 7
    U0, V0, u, i, j = _problem(d, nu, ni, d, B)
    s = Step(simt, U0, V0, d, u, i, j, lr=0.7, reg=0.02, sink=SINK_STAGE, want_x=True, chunk=chunk)
    s.apply(simt)
    Ur, Vr, lref = O.sgd_step(U0, V0, u, i, j, 0.7, 0.02)
    np.testing.assert_allclose(s.U[:, :d], Ur, rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(s.V[:, :d], Vr, rtol=2e-5, atol=2e-6)
    assert abs(s.loss[0] / B - float(lref)) < 2e-5 * max(1.0, float(lref))
    _, _, _, x = O.bpr_grads(U0, V0, u, i, j)
    np.testing.assert_allclose(s.x, x, rtol=1e-5, atol=2e-6)
    assert not s.U[:, d:].any() and not s.V[:, d:].any()


@pytest.mark.parametrize("d", [8, 64, 128])
def test_ldg_kernel_gradient_and_forward_sinks(simt, d):
    U0, V0, u, i, j = _problem(3 * d, 50, 40, d, 200, std=1.0)
    s = Step(simt, U0, V0, d, u, i, j, reg=0.05, sink=SINK_GRAD)
    dU, dV, _, _ = O.bpr_grads(U0, V0, u, i, j, 0.05)
    np.testing.assert_allclose(s.gU[:, :d], dU, rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(s.gV[:, :d], dV, rtol=2e-5, atol=1e-7)
    assert np.array_equal(s.U[:, :d], U0) and np.array_equal(s.V[:, :d], V0)         # SINK_GRAD leaves the tables alone
    n = Step(simt, U0, V0, d, u, i, j, sink=SINK_NONE)
    assert abs(n.loss[0] - s.loss[0]) < 1e-9 and np.array_equal(n.U[:, :d], U0)


# ---- the three in-place training kernels (generic, warp-per-row, sub-warp groups = the default) ----------------------------
@pytest.mark.parametrize("kind,chunk", [(LDG, 32), (LDG, 4), (FAST, 4), (FAST, 32), (GROUP8, 0), (GROUP16, 0)])
@pytest.mark.parametrize("uniq", [0, F_USERS_UNIQUE])
def test_fused_update_without_collisions_is_the_exact_step(simt, kind, chunk, uniq):
    """With no id shared between triples the one-kernel in-place step IS the exact step (d = 128, ragged batch)."""
    nu, ni, d, B = 700, 1300, 128, 601
    U0, V0, u, i, j = _problem(kind * 10 + chunk, nu, ni, d, B, std=0.3, unique_users=True, unique_items=True)
    s = Step(simt, U0, V0, d, u, i, j, lr=0.9, reg=0.01, sink=SINK_UPDATE, flags=uniq, kind=kind, chunk=chunk)
    Ur, Vr, lref = O.sgd_step(U0, V0, u, i, j, 0.9, 0.01)
    np.testing.assert_allclose(s.U, Ur, rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(s.V, Vr, rtol=2e-5, atol=2e-6)
    assert abs(s.loss[0] / B - float(lref)) < 2e-5


@pytest.mark.parametrize("kind", [LDG, FAST, GROUP8, GROUP16])
def test_item_delta_variant_is_exact_under_item_collisions(simt, kind):
    """F_ITEM_DELTA (user-sharded multi-GPU layout): user rows in place (unique users), item deltas into a separate
    buffer from PRE-step item rows - exact whatever the collisions: V0 + dV == oracle, V itself untouched; the global
    batch size of a sharded step comes in through inv_batch."""
    nu, ni, d, B = 500, 40, 128, 333
    U0, V0, u, i, j = _problem(50 + kind, nu, ni, d, B, std=0.3, unique_users=True)
    s = Step(simt, U0, V0, d, u, i, j, lr=2.0, reg=0.01, sink=SINK_UPDATE, flags=F_USERS_UNIQUE | F_ITEM_DELTA, kind=kind,
             inv_batch=1.0 / B)
    Ur, Vr, _ = O.sgd_step(U0, V0, u, i, j, 2.0, 0.01)
    np.testing.assert_allclose(s.U, Ur, rtol=2e-5, atol=2e-6)
    assert np.array_equal(s.V, V0)
    np.testing.assert_allclose(V0 + s.gV, Vr, rtol=2e-5, atol=2e-6)
    h = Step(simt, U0, V0, d, u, i, j, lr=2.0, reg=0.01, sink=SINK_UPDATE, flags=F_USERS_UNIQUE | F_ITEM_DELTA, kind=kind,
             inv_batch=0.5 / B)                                                    # twice the global batch: half the step
    np.testing.assert_allclose(h.gV, 0.5 * s.gV, rtol=2e-5, atol=1e-7)


@pytest.mark.parametrize("kind", [LDG, FAST, GROUP8])
def test_fused_update_with_collisions_is_close(simt, kind):
    """Hogwild inside a launch (tests/test_gpu_parity.py::test_fused_step_with_collisions_is_close): colliding rows may
    read partially-updated weights; the deviation from the exact step stays second order in the step."""
    nu, ni, d, B = 500, 300, 128, 2048                       # the device test's regime: per-triple step lr / B = 0.005
    U0, V0, u, i, j = _problem(70 + kind, nu, ni, d, B, std=0.3)
    s = Step(simt, U0, V0, d, u, i, j, lr=10.0, reg=0.0, sink=SINK_UPDATE, kind=kind)
    Ur, Vr, _ = O.sgd_step(U0, V0, u, i, j, 10.0, 0.0)
    step = max(np.abs(Ur - U0).max(), np.abs(Vr - V0).max())
    dev = max(np.abs(s.U - Ur).max(), np.abs(s.V - Vr).max())
    assert step > 1e-2 and 0 < dev < 0.05 * step, (step, dev)


# ---- on-device sampling inside the kernels ---------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", [LDG, FAST, GROUP8])
def test_sampling_step_equals_given_triples_step(simt, kind):
    """The kernels draw (pos, neg) themselves: the triples they report are the host mirror's, and the tables equal the
    step on those triples given explicitly (tests/test_gpu_parity.py::test_fused_sampling_step_equals_given_triples)."""
    rng = np.random.default_rng(kind)
    nu, ni, d, B = 400, 900, 128, 350
    rows = [np.sort(rng.choice(ni, int(rng.integers(1, 25)), replace=False)).astype(np.int32) for _ in range(nu)]
    indptr = np.zeros(nu + 1, np.int64); indptr[1:] = np.cumsum([len(r) for r in rows])
    indices = np.concatenate(rows)
    U0 = (rng.standard_normal((nu, d)) * 0.3).astype(np.float32); V0 = (rng.standard_normal((ni, d)) * 0.3).astype(np.float32)
    users = rng.permutation(nu)[:B]
    s = Step(simt, U0, V0, d, users, csr=(indptr, indices), lr=1.5, reg=0.01, sink=SINK_UPDATE,
             flags=F_USERS_UNIQUE | F_ITEM_DELTA, seed=2020, step=7, kind=kind)
    pos, neg = O.sample_triples_vec(2020, 7, users, indptr, indices, ni)
    assert np.array_equal(s.out_pos, pos) and np.array_equal(s.out_neg, neg)
    for t in range(0, B, 11):
        assert pos[t] in rows[users[t]] and neg[t] not in rows[users[t]]
    Ur, Vr, lref = O.sgd_step(U0, V0, users, pos, neg, 1.5, 0.01)
    np.testing.assert_allclose(s.U, Ur, rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(V0 + s.gV, Vr, rtol=2e-5, atol=2e-6)
    assert abs(s.loss[0] / B - float(lref)) < 2e-5


# ---- optimisers and the reference's own training trajectories, replayed by the kernels on the host ----------------------------
P = lambda a: a.ctypes.data if a is not None else None


def _grad_step(simt, U, V, d, u, i, j, reg=0.0):
    s = Step(simt, U[:, :d], V[:, :d], d, u, i, j, reg=reg, sink=SINK_GRAD)
    return s.gU, s.gV, s.loss[0] / len(u)


def test_tiny_forward_and_dense_adam_trajectory_match_the_reference(simt, golden):
    """tests/golden/tiny_bpr.npz: models/MF.py forward scores, then the reference's optimiser as-is (torch.optim.Adam
    lr=1e-3, MF.py:30) for 3 steps with duplicate ids - SINK_GRAD kernel + adam_dense_kernel."""
    g = golden["tiny_bpr"]
    U, V = _pad(g["U0"]), _pad(g["V0"])
    u0, i0, j0 = (np.ascontiguousarray(g[k][0], np.int32) for k in ("users", "pos", "neg"))
    out = np.zeros(16, np.float32)
    simt.emu_mf_forward(P(U), P(V), 8, 8, P(u0), P(i0), 16, P(out))
    np.testing.assert_allclose(out, g["pos_scores"], rtol=1e-6, atol=1e-6)
    st = [np.zeros_like(U), np.zeros_like(U), np.zeros_like(V), np.zeros_like(V)]
    for b in range(3):
        gU, gV, _ = _grad_step(simt, U, V, 8, g["users"][b], g["pos"][b], g["neg"][b])
        simt.emu_adam_dense(P(U), P(gU), P(st[0]), P(st[1]), U.size, 1e-3, 0.9, 0.999, 1e-8, b + 1)
        simt.emu_adam_dense(P(V), P(gV), P(st[2]), P(st[3]), V.size, 1e-3, 0.9, 0.999, 1e-8, b + 1)
    np.testing.assert_allclose(U[:, :8], g["adam_U"], rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(V[:, :8], g["adam_V"], rtol=1e-4, atol=2e-6)


def test_lazy_adam_kernel_matches_the_reference_sparse_adam(simt, golden):
    """adam_rows_kernel (one claimed update per touched row, scratch zeroed row by row) vs the reference MF
    driven by torch.optim.SparseAdam, 6 steps with duplicate users / items."""
    g, t = golden["tiny_bpr"], golden["tiny_lazy_adam"]
    U, V = _pad(g["U0"]), _pad(g["V0"])
    gU, gV = np.zeros_like(U), np.zeros_like(V)
    st = [np.zeros_like(U), np.zeros_like(U), np.zeros_like(V), np.zeros_like(V)]
    sU, sV = np.zeros(U.shape[0], np.int32), np.zeros(V.shape[0], np.int32)
    s = 0
    for rep in range(2):
        for b in range(3):
            u, i, j = (np.ascontiguousarray(g[k][b], np.int32) for k in ("users", "pos", "neg"))
            dU, dV, loss = _grad_step(simt, U, V, 8, u, i, j)
            gU += dU; gV += dV                                              # the kernel accumulates into the persistent scratch
            assert abs(loss - float(t["loss"][s])) < 1e-5
            for W, gr, m, v, stamp, ids in ((U, gU, st[0], st[1], sU, u), (V, gV, st[2], st[3], sV, i), (V, gV, st[2], st[3], sV, j)):
                simt.emu_adam_rows(P(W), P(gr), P(m), P(v), P(stamp), W.shape[1], P(ids), len(ids), 1e-3, 0.9, 0.999, 1e-8, s + 1)
            assert not gU.any() and not gV.any()                            # scratch left zero
            np.testing.assert_allclose(U[:, :8], t["U"][s], rtol=1e-4, atol=2e-6)
            np.testing.assert_allclose(V[:, :8], t["V"][s], rtol=1e-4, atol=2e-6)
            s += 1


@pytest.mark.parametrize("tag", ["sgd", "adam"])
def test_ml100k_reference_trajectory_replayed_by_the_kernels(simt, golden, oracle_c, tag):
    """BASELINE configs[0] on the host: the reference's own ml-100k run (main.py sequence, d = 32, B = 256, its sampler's
    recorded batches) replayed through the kernels - SGD swap (stage + apply, per-occurrence L2) and Adam as-is (dense
    gradient + Adam sweep): per-batch losses and the final tables; same tolerances as the device test."""
    g = golden["ml100k"]
    U, V = _pad(g[f"{tag}_U0"]), _pad(g[f"{tag}_V0"])
    st = [np.zeros_like(U), np.zeros_like(U), np.zeros_like(V), np.zeros_like(V)]
    losses, off = [], 0
    for t, n in enumerate(g[f"{tag}_blen"], 1):
        sl = slice(off, off + int(n)); off += int(n)
        u, i, j = g[f"{tag}_bu"][sl], g[f"{tag}_bi"][sl], g[f"{tag}_bj"][sl]
        if tag == "sgd":
            s = Step(simt, U[:, :32], V[:, :32], 32, u, i, j, lr=float(g["sgd_lr"]), reg=float(g["sgd_reg"]), sink=SINK_STAGE)
            s.apply(simt)
            U, V = s.U, s.V
            losses.append(s.loss[0] / int(n))
        else:
            gU, gV, loss = _grad_step(simt, U, V, 32, u, i, j)
            simt.emu_adam_dense(P(U), P(gU), P(st[0]), P(st[1]), U.size, 1e-3, 0.9, 0.999, 1e-8, t)
            simt.emu_adam_dense(P(V), P(gV), P(st[2]), P(st[3]), V.size, 1e-3, 0.9, 0.999, 1e-8, t)
            losses.append(loss)
    if tag == "adam":
        np.testing.assert_allclose(losses, g["adam_losses"], rtol=2e-5)
        for got, ref in ((U[:, :32], g["adam_U"]), (V[:, :32], g["adam_V"])):
            bad = ~np.isclose(got, ref, rtol=2e-4, atol=2e-5)              # Adam amplifies ulp-level sigmoid saturation
            assert bad.mean() < 0.01 and np.abs(got - ref).max() < 12 * 1e-3
    else:
        np.testing.assert_allclose(U[:, :32], g["sgd_U"], rtol=2e-4, atol=2e-5)
        np.testing.assert_allclose(V[:, :32], g["sgd_V"], rtol=2e-4, atol=2e-5)
    # ... and ends at the reference's NDCG@10 (BASELINE: within 1e-4).  Scoring / metrics through the C oracle here: the
    # scoring and metric kernels are checked bit for bit against it in tests/test_scoring_on_simt_host.py and
    # tests/test_device_code_on_host.py
    nu, ni = int(g["num_users"]), int(g["num_items"])
    idx, sc = oracle_c.score_topk(np.ascontiguousarray(U[:, :32]), np.ascontiguousarray(V[:, :32]), 32, np.arange(nu), ni,
                                  g["train_indptr"], g["train_indices"], 10)
    np.testing.assert_allclose(sc, g[f"{tag}_top10_scores"], rtol=2e-4, atol=2e-4)
    assert (idx == g[f"{tag}_top10"]).mean() > 0.99
    truths = [g["valid_indices"][g["valid_indptr"][r]:g["valid_indptr"][r + 1]] for r in range(nu)]
    ndcg10 = float(O.mean_f32(oracle_c.holdout(idx, truths, [5, 10])[:, 5]))
    assert abs(ndcg10 - float(g[f"{tag}_NDCG@10"][-1])) < 1e-4


def test_dense_helpers_of_the_multi_gpu_layouts(simt):
    """sgd_dense, rows_add, delta_diff / delta_apply, snap_apply, add_clear: elementwise kernels against numpy."""
    rng = np.random.default_rng(0)
    n = 4 * 777
    W, g_ = rng.standard_normal(n).astype(np.float32), rng.standard_normal(n).astype(np.float32)
    W0 = W.copy()
    simt.emu_sgd_dense(P(W), P(g_), n, 0.3)
    np.testing.assert_array_equal(W, W0 - np.float32(0.3) * g_)
    snap = rng.standard_normal(n).astype(np.float32)
    wire, own = np.zeros(n, np.float32), np.zeros(n, np.float32)
    simt.emu_delta_diff(P(W), P(snap), P(wire), P(own), n)
    assert np.array_equal(wire, W - snap) and np.array_equal(own, wire)
    total = wire + rng.standard_normal(n).astype(np.float32)                # what the all-reduce would return
    W1 = W.copy()
    simt.emu_delta_apply(P(W1), P(total), P(own), n)
    np.testing.assert_array_equal(W1, W + (total - own))
    W2 = np.zeros(n, np.float32)
    simt.emu_snap_apply(P(W2), P(snap), P(total), 0.25, n)
    np.testing.assert_allclose(W2, snap + np.float32(0.25) * total, rtol=1e-6, atol=1e-7)
    d_ = total.copy(); W3 = W.copy()
    simt.emu_add_clear(P(W3), P(d_), n)
    assert np.array_equal(W3, W + total) and not d_.any()
    T = rng.standard_normal((50, 12)).astype(np.float32); T0 = T.copy()
    ids = rng.integers(0, 50, 80).astype(np.int32)                           # duplicates: atomics
    delta = rng.standard_normal((80, 16)).astype(np.float32)
    simt.emu_rows_add(P(T), 12, P(ids), 80, P(delta), 16, 0.5)
    want = T0.astype(np.float64)
    np.add.at(want, ids, 0.5 * delta[:, :12].astype(np.float64))
    np.testing.assert_allclose(T, want, rtol=1e-5, atol=1e-6)


# ---- the asynchronous-gather variants of the step (F_TMA_GATHER: cp.async.bulk rows + mbarriers; F_ASYNC_GATHER: cp.async ring)
# The copy engines are not emulated: a copy completes when it is issued (one valid schedule) and the barriers are no-ops -
# what runs from the source is the staging layout, the slot / phase bookkeeping and all the arithmetic.
ASYNC, TMA = 4, 5


@pytest.mark.parametrize("d,chunk", [(8, 32), (32, 8), (50, 32), (64, 4), (128, 32), (200, 16), (256, 32), (400, 8)])
def test_tma_variant_exact_step_all_widths(simt, d, chunk):
    nu, ni, B = 97, 61, 250 + d % 5
    U0, V0, u, i, j = _problem(1000 + d, nu, ni, d, B)
    s = Step(simt, U0, V0, d, u, i, j, lr=0.7, reg=0.02, sink=SINK_STAGE, want_x=True, kind=TMA, chunk=chunk)
    s.apply(simt)
    Ur, Vr, lref = O.sgd_step(U0, V0, u, i, j, 0.7, 0.02)
    np.testing.assert_allclose(s.U[:, :d], Ur, rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(s.V[:, :d], Vr, rtol=2e-5, atol=2e-6)
    assert abs(s.loss[0] / B - float(lref)) < 2e-5 * max(1.0, float(lref))
    g = Step(simt, U0, V0, d, u, i, j, reg=0.02, sink=SINK_GRAD, kind=TMA, chunk=chunk)
    dU, dV, _, _ = O.bpr_grads(U0, V0, u, i, j, 0.02)
    np.testing.assert_allclose(g.gU[:, :d], dU, rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(g.gV[:, :d], dV, rtol=2e-5, atol=1e-7)


@pytest.mark.parametrize("kind,d", [(TMA, 128), (ASYNC, 128), (ASYNC, 256)])
@pytest.mark.parametrize("uniq", [0, F_USERS_UNIQUE])
def test_async_gather_variants_in_place_update(simt, kind, d, uniq):
    """smoke() and the device suite run the TMA variant beside the default: in-place update without collisions == exact
    step; with in-kernel sampling the reported triples are the host mirror's."""
    nu, ni, B = 500, 1100, 389
    U0, V0, u, i, j = _problem(kind + d, nu, ni, d, B, std=0.3, unique_users=True, unique_items=True)
    s = Step(simt, U0, V0, d, u, i, j, lr=0.9, reg=0.01, sink=SINK_UPDATE, flags=uniq, kind=kind, chunk=32)
    Ur, Vr, lref = O.sgd_step(U0, V0, u, i, j, 0.9, 0.01)
    np.testing.assert_allclose(s.U[:, :d], Ur, rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(s.V[:, :d], Vr, rtol=2e-5, atol=2e-6)
    assert abs(s.loss[0] / B - float(lref)) < 2e-5
    rng = np.random.default_rng(3)
    rows = [np.sort(rng.choice(ni, int(rng.integers(1, 25)), replace=False)).astype(np.int32) for _ in range(nu)]
    indptr = np.zeros(nu + 1, np.int64); indptr[1:] = np.cumsum([len(r) for r in rows])
    indices = np.concatenate(rows)
    users = rng.permutation(nu)[:B]
    t = Step(simt, U0, V0, d, users, csr=(indptr, indices), lr=0.0, sink=SINK_UPDATE, flags=F_USERS_UNIQUE, seed=2020, step=3,
             kind=kind, chunk=32)
    pos, neg = O.sample_triples_vec(2020, 3, users, indptr, indices, ni)
    assert np.array_equal(t.out_pos, pos) and np.array_equal(t.out_neg, neg)
