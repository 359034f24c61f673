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
