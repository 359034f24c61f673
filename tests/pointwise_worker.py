"""Device check of b200rec_pointwise_step against the oracle and the reference's golden vectors; run in its own process
by tests/test_gpu_pointwise.py (green on a B200 since the round-1 driver run; its own process so that a fault cannot take
the suite's CUDA context with it)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def table(a, dev):
    from recsys_pytorch_b200 import engine
    t = engine.alloc_table(a.shape[0], a.shape[1], dev, std=0.0)
    t[:, :a.shape[1]] = torch.from_numpy(a).to(dev)
    return t


def main():
    from recsys_pytorch_b200 import engine
    from recsys_pytorch_b200._lib import SINK_GRAD, SINK_NONE, SINK_UPDATE
    from oracle import bpr_oracle as O
    dev = torch.device("cuda:0")
    g = np.load(os.path.join(ROOT, "tests", "golden", "tiny_pointwise.npz"))
    t = np.load(os.path.join(ROOT, "tests", "golden", "tiny_bpr.npz"))
    for lf in ("ce", "mse"):                                  # 1. reference golden: loss + autograd gradients (duplicates)
        sc = float(g[f"{lf}_scale"])
        U0, V0 = (t["U0"] * sc).astype(np.float32), (t["V0"] * sc).astype(np.float32)
        d = U0.shape[1]
        U, V = table(U0, dev), table(V0, dev)
        u, i = (torch.from_numpy(g[k][0].astype(np.int32)).to(dev) for k in ("users", "items"))
        r = torch.from_numpy(g["ratings"][0]).to(dev)
        gU, gV = torch.zeros_like(U), torch.zeros_like(V)
        loss = torch.zeros(1, dtype=torch.float64, device=dev)
        engine.pointwise_step(U, V, d, u, i, r, loss_func=lf, sink=SINK_GRAD, gU=gU, gV=gV, loss_sum=loss)
        B = u.numel()
        assert abs(loss.item() / B - float(g[f"{lf}_loss"])) <= 2e-5 * max(1.0, abs(float(g[f"{lf}_loss"]))), (lf, loss.item() / B)
        np.testing.assert_allclose(gU.cpu().numpy()[:, :d], g[f"{lf}_dU"], rtol=2e-5, atol=2e-6)
        np.testing.assert_allclose(gV.cpu().numpy()[:, :d], g[f"{lf}_dV"], rtol=2e-5, atol=2e-6)
        assert torch.equal(U.cpu()[:, :d], torch.from_numpy(U0)) and torch.equal(V.cpu()[:, :d], torch.from_numpy(V0))
        loss.zero_()
        engine.pointwise_step(U, V, d, u, i, r, loss_func=lf, sink=SINK_NONE, loss_sum=loss)
        assert abs(loss.item() / B - float(g[f"{lf}_loss"])) <= 2e-5 * max(1.0, abs(float(g[f"{lf}_loss"])))
    rng = np.random.default_rng(3)                             # 2. every row width, in-place SGD update, no repeated rows
    for d in (128, 64, 50, 200, 8, 256):
        nu, ni, B = 700, 900, 512
        U0 = (rng.standard_normal((nu, d)) * 0.3).astype(np.float32); V0 = (rng.standard_normal((ni, d)) * 0.3).astype(np.float32)
        u = rng.permutation(nu)[:B]; i = rng.permutation(ni)[:B]
        y = (rng.random(B) < 0.4).astype(np.float32)
        for lf in ("ce", "mse"):
            U, V = table(U0, dev), table(V0, dev)
            lr, reg = 0.7, 0.01
            engine.pointwise_step(U, V, d, torch.from_numpy(u.astype(np.int32)).to(dev), torch.from_numpy(i.astype(np.int32)).to(dev),
                                  torch.from_numpy(y).to(dev), loss_func=lf, lr=lr, reg=reg, sink=SINK_UPDATE)
            dU, dV, _, _ = O.pointwise_grads(U0, V0, u, i, y, lf)
            rU = np.zeros_like(U0); rV = np.zeros_like(V0)
            np.add.at(rU, u, U0[u]); np.add.at(rV, i, V0[i])                         # per-occurrence L2
            Ur = U0 - lr * (dU + reg / B * rU); Vr = V0 - lr * (dV + reg / B * rV)
            np.testing.assert_allclose(U.cpu().numpy()[:, :d], Ur, rtol=2e-5, atol=2e-6)
            np.testing.assert_allclose(V.cpu().numpy()[:, :d], Vr, rtol=2e-5, atol=2e-6)
            assert float(U.cpu()[:, d:].abs().sum()) == 0.0                       # pad columns stay zero
    torch.cuda.synchronize()
    print("POINTWISE_OK")


if __name__ == "__main__":
    main()
