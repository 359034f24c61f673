"""`-m gpu` (sorted last: added after the round's last GPU call): the CUDA step at the row width and batch regime of
BASELINE configs[1] - d = 128, ONE 65,536-triple batch whose Zipf head puts 7,164 triples on one item row - against what
the reference itself computed from the same inputs (tests/golden/cfg2shape_bpr.npz, oracle/make_golden_cfg2shape.py;
tests/test_oracle_cpu.py pins the numpy oracle to the same file).  Exact-semantics sinks only (all gradients from pre-step
weights, like autograd); the one-kernel Hogwild sink is bounded in tests/test_gpu_parity.py."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_cuda_step_matches_reference_at_cfg2_row_width(golden, dev):
    from oracle.make_golden_cfg2shape import B, D, LR, inputs
    from recsys_pytorch_b200 import engine
    from recsys_pytorch_b200._lib import SINK_GRAD, SINK_NONE, SINK_STAGE
    g = golden["cfg2shape_bpr"]
    U0, V0, u, i, j, su, si = inputs(int(g["seed"]))
    U = engine.alloc_table(U0.shape[0], D, dev, std=0.0); U[:, :D] = torch.from_numpy(U0).to(dev)
    V = engine.alloc_table(V0.shape[0], D, dev, std=0.0); V[:, :D] = torch.from_numpy(V0).to(dev)
    tu, ti, tj = (torch.from_numpy(a.astype(np.int32)).to(dev) for a in (u, i, j))
    # forward (models/MF.py:38-42) and the loss (:99-105)
    x = (engine.mf_forward(U, V, D, tu, ti) - engine.mf_forward(U, V, D, tu, tj)).cpu().numpy()
    np.testing.assert_allclose(x, g["x"], rtol=1e-5, atol=2e-6)
    loss = torch.zeros(1, dtype=torch.float64, device=dev)
    xk = torch.empty(B, device=dev)
    engine.bpr_step(U, V, D, tu, ti, tj, sink=SINK_NONE, loss_sum=loss, x_out=xk)
    assert abs(loss.item() / B - float(g["loss"])) < 5e-6
    np.testing.assert_allclose(xk.cpu().numpy(), g["x"], rtol=1e-5, atol=2e-6)
    # backward (:67): dense gradient rows; fp32 vector atomics in arrival order vs the reference's fp32 accumulation
    gU, gV = torch.zeros_like(U), torch.zeros_like(V)
    engine.bpr_step(U, V, D, tu, ti, tj, sink=SINK_GRAD, gU=gU, gV=gV)
    for got, ref in ((gU.cpu().numpy()[su][:, :D], g["dU_rows"]), (gV.cpu().numpy()[si][:, :D], g["dV_rows"])):
        np.testing.assert_allclose(got, ref, rtol=2e-5, atol=5e-5 * np.abs(ref).max())
    assert abs(float(gU.double().abs().sum()) / float(g["dU_abs_sum"]) - 1) < 1e-5              # the unsampled rows too
    assert abs(float(gV.double().abs().sum()) / float(g["dV_abs_sum"]) - 1) < 1e-5
    assert np.array_equal(U.cpu().numpy()[:, :D], U0)                                           # tables untouched so far
    # optimiser step (:68 with the SGD swap): stage + apply = all gradients from pre-step weights
    stage = torch.empty((B, 3, U.shape[1]), dtype=torch.float32, device=dev)
    engine.bpr_step(U, V, D, tu, ti, tj, lr=float(LR), reg=0.0, sink=SINK_STAGE, stage=stage)
    engine.bpr_apply(U, V, tu, ti, tj, stage)
    np.testing.assert_allclose(U.cpu().numpy()[su][:, :D], g["U_rows"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(V.cpu().numpy()[si][:, :D], g["V_rows"], rtol=2e-5, atol=2e-6)
