"""Runs, on the host SIMT emulator (tests/simt_host.py: the REAL kernel sources compiled for the host), the assertions of
the three `-m gpu` tests that were written or re-stated after the last GPU call of round 2 - same inputs, same sizes:

    python tools/emulate_gpu_assertions.py            (about 2 minutes on 8 cores)

  cfg2shape   tests/test_zz_gpu_cfg2shape.py                                          d = 128, one 65,536-triple batch
  step_diff   tests/test_dist.py::test_step_diff_world1_equals_delta_buffer_step      in-place vs delta-buffer step
  stale       tests/test_dist.py::test_step_overlapped_matches_one_step_stale_oracle  item delta applied one step late
"""
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import bpr_oracle as O  # noqa: E402
from recsys_pytorch_b200._lib import F_ITEM_DELTA, F_USERS_UNIQUE, SINK_GRAD, SINK_NONE, SINK_STAGE  # noqa: E402
from tests import simt_host  # noqa: E402
from tests.test_kernels_on_simt_host import GROUP8, Step  # noqa: E402


def cfg2shape(simt):
    from oracle.make_golden_cfg2shape import B, D, LR, inputs
    g = np.load(os.path.join(ROOT, "tests", "golden", "cfg2shape_bpr.npz"))
    U0, V0, u, i, j, su, si = inputs(int(g["seed"]))
    s = Step(simt, U0, V0, D, u, i, j, sink=SINK_NONE, want_x=True, grid=8)
    assert abs(s.loss[0] / B - float(g["loss"])) < 5e-6
    np.testing.assert_allclose(s.x, g["x"], rtol=1e-5, atol=2e-6)
    s = Step(simt, U0, V0, D, u, i, j, sink=SINK_GRAD, want_loss=False, grid=8)
    for got, ref in ((s.gU[su][:, :D], g["dU_rows"]), (s.gV[si][:, :D], g["dV_rows"])):
        np.testing.assert_allclose(got, ref, rtol=2e-5, atol=5e-5 * np.abs(ref).max())
        print("  gradient rows: max |dev| / scale = %.2e (allowed 5e-5)" % (np.abs(got - ref).max() / np.abs(ref).max()))
    assert abs(np.abs(s.gU.astype(np.float64)).sum() / float(g["dU_abs_sum"]) - 1) < 1e-5
    assert abs(np.abs(s.gV.astype(np.float64)).sum() / float(g["dV_abs_sum"]) - 1) < 1e-5
    s = Step(simt, U0, V0, D, u, i, j, lr=float(LR), sink=SINK_STAGE, want_loss=False, grid=8)
    s.apply(simt, grid=8)
    np.testing.assert_allclose(s.U[su][:, :D], g["U_rows"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(s.V[si][:, :D], g["V_rows"], rtol=2e-5, atol=2e-6)


def _csr(rng, nu, ni, lo, hi):
    rows = [np.sort(rng.choice(ni, rng.integers(lo, hi), replace=False)).astype(np.int32) for _ in range(nu)]
    indptr = np.zeros(nu + 1, np.int64); indptr[1:] = np.cumsum([len(r) for r in rows])
    return indptr, np.concatenate(rows)


def step_diff(simt):
    rng = np.random.default_rng(11)
    nu, ni, d, B = 3000, 500, 128, 1024
    csr = _csr(rng, nu, ni, 2, 20)
    r2 = np.random.default_rng(4)
    U0 = (r2.standard_normal((nu, d)) * 0.1).astype(np.float32); V0 = (r2.standard_normal((ni, d)) * 0.1).astype(np.float32)
    Ua, Va, Ub, Vb = U0.copy(), V0.copy(), U0.copy(), V0.copy()
    kw = dict(csr=csr, lr=2.0, reg=0.01, seed=4, kind=GROUP8, inv_batch=1.0 / B, want_loss=False)
    for s in range(4):
        users = rng.permutation(nu)[:B]
        a = Step(simt, Ua, Va, d, users, flags=F_USERS_UNIQUE | F_ITEM_DELTA, step=s + 1, **kw)      # UserShardedBPR.step
        Ua, Va = a.U, a.V + a.gV
        b = Step(simt, Ub, Vb, d, users, flags=F_USERS_UNIQUE, step=s + 1, **kw)                     # .step_diff at world 1
        Ub, Vb = b.U, b.V
    moved = np.abs(Va - V0).max()
    dv, du = np.abs(Vb - Va).max(), np.abs(Ub - Ua).max()
    print("  moved %.2e; |b - a|: V %.2e U %.2e (allowed %.2e)" % (moved, dv, du, 0.05 * moved))
    assert moved > 1e-3 and dv < 0.05 * moved and du < 0.05 * moved
    old = np.allclose(Vb, Va, rtol=1e-4, atol=1e-6) and np.allclose(Ub, Ua, rtol=1e-4, atol=1e-6)
    print("  the tolerance this test had before (rtol 1e-4, atol 1e-6) would %s" % ("pass" if old else "FAIL: second-order Hogwild term"))


def stale(simt):
    rng = np.random.default_rng(21)
    nu, ni, d, B, LR = 3000, 700, 128, 1024, 300.0
    csr = _csr(rng, nu, ni, 2, 20)
    r2 = np.random.default_rng(4)
    U0 = (r2.standard_normal((nu, d)) * 0.1).astype(np.float32); V0 = (r2.standard_normal((ni, d)) * 0.1).astype(np.float32)
    U, V, pending, batches = U0.copy(), V0.copy(), None, []
    for s in range(5):
        users = rng.permutation(nu)[:B]
        a = Step(simt, U, V, d, users, csr=csr, lr=LR, reg=0.01, flags=F_USERS_UNIQUE | F_ITEM_DELTA, seed=4, step=s + 1,
                 kind=GROUP8, inv_batch=1.0 / B, want_loss=False)
        U = a.U
        if pending is not None:
            V = V + pending                                                   # the delta of the previous step lands now
        pending = a.gV.copy()
        batches.append((users, a.out_pos.copy(), a.out_neg.copy()))
    V = V + pending
    Ur, Vr = O.sgd_steps_stale_items(U0, V0, batches, LR, 0.01)
    print("  max |dev|: U %.2e V %.2e (allowed 5e-6 + 2e-5 rel)" % (np.abs(U - Ur).max(), np.abs(V - Vr).max()))
    np.testing.assert_allclose(U, Ur, rtol=2e-5, atol=5e-6)
    np.testing.assert_allclose(V, Vr, rtol=2e-5, atol=5e-6)


if __name__ == "__main__":
    simt = simt_host.build(tempfile.mkdtemp())
    for fn in (step_diff, stale, cfg2shape):
        t0 = time.time()
        print(fn.__name__)
        fn(simt)
        print("  holds (%.0f s)" % (time.time() - t0))
