"""Generate tests/golden/cfg2shape_bpr.npz by RUNNING THE REFERENCE (models/MF.py forward / process_one_batch + autograd)
at the row width and batch regime of BASELINE configs[1]: d = 128 and ONE 65,536-triple batch with heavy id collisions
(SURVEY 8(c) golden (iii)).  TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden_cfg2shape          # from the repo root

The tables are not stored: `inputs(seed)` below regenerates them (numpy Generator streams are reproducible), the golden
holds what the reference computed from them - the 65,536 score differences, the loss, the dense-gradient rows of a fixed
sample of users / items, and those rows after one step of the SGD swap - 0.6 MB instead of 13 MB.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

OUT = os.path.join(ROOT, "tests", "golden")
NU, NI, D, B, SEED, LR = 20_000, 5_000, 128, 65_536, 31, 40.0
N_SAMPLE = 384


def inputs(seed=SEED):
    """(U0, V0, users, pos, neg, sample_users, sample_items) - shared by the generator and the tests."""
    rng = np.random.default_rng(seed)
    U0 = (rng.standard_normal((NU, D)) * 0.1).astype(np.float32)
    V0 = (rng.standard_normal((NI, D)) * 0.1).astype(np.float32)
    users = rng.integers(0, NU, B).astype(np.int64)                       # ~3.3 triples per user, ~26 per item
    zipf = 1.0 / np.arange(1, NI + 1); zipf /= zipf.sum()
    pos = rng.choice(NI, B, p=zipf).astype(np.int64)                      # Zipf head: the hottest item ~7,000 times
    neg = rng.integers(0, NI, B).astype(np.int64)
    su = np.unique(np.concatenate([users[:N_SAMPLE // 2], rng.integers(0, NU, N_SAMPLE // 2)]))
    si = np.unique(np.concatenate([np.arange(16), pos[:N_SAMPLE // 2], rng.integers(0, NI, N_SAMPLE // 2)]))
    return U0, V0, users, pos, neg, su, si


def main():
    import torch
    from oracle import ref_harness
    ref = ref_harness.load()
    U0, V0, users, pos, neg, su, si = inputs()
    ds = types.SimpleNamespace(num_users=NU, num_items=NI)
    m = ref.MF(ds, {"hidden_dim": D, "pointwise": False, "loss_func": "ce"}, torch.device("cpu"))
    with torch.no_grad():
        m.user_embedding.weight.copy_(torch.from_numpy(U0)); m.item_embedding.weight.copy_(torch.from_numpy(V0))
    m.optimizer = torch.optim.SGD(m.parameters(), lr=LR)                  # SURVEY H1 swap (the reference's Adam is covered elsewhere)
    u, i, j = (torch.from_numpy(a) for a in (users, pos, neg))
    with torch.no_grad():
        x = (m.forward(u, i) - m.forward(u, j)).numpy().copy()           # models/MF.py:99-105
    m.optimizer.zero_grad()
    loss = m.process_one_batch(u, i, j); loss.backward()                   # models/MF.py:64-67
    dU = m.user_embedding.weight.grad.numpy().copy(); dV = m.item_embedding.weight.grad.numpy().copy()
    m.optimizer.step()                                                     # :68
    out = dict(seed=np.int64(SEED), lr=np.float32(LR), x=x.astype(np.float32), loss=np.float32(loss.item()),
               sample_users=su.astype(np.int32), sample_items=si.astype(np.int32),
               dU_rows=dU[su], dV_rows=dV[si],
               U_rows=m.user_embedding.weight.detach().numpy()[su].copy(),
               V_rows=m.item_embedding.weight.detach().numpy()[si].copy(),
               dU_abs_sum=np.float64(np.abs(dU.astype(np.float64)).sum()), dV_abs_sum=np.float64(np.abs(dV.astype(np.float64)).sum()))
    path = os.path.join(OUT, "cfg2shape_bpr.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; loss", float(loss.item()), "hottest item count", int(np.bincount(pos).max()))


if __name__ == "__main__":
    main()
