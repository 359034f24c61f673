"""A few launches of the fused BPR step at cfg2 for ncu (MODE=sample|given, STEPS, env B200REC_ABL / B200REC_STEP_VARIANT)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recsys_pytorch_b200 import _lib, engine, synthetic
dev = torch.device("cuda:0")
NU, NI, D, B = 1_000_000, int(os.environ.get("NI", 100_000)), 128, 1_000_000
train, _ = synthetic.make_interactions(NU, NI, seed=2020, device=dev)
g = torch.Generator(device=dev); g.manual_seed(1)
U = engine.alloc_table(NU, D, dev, 0.01, g); V = engine.alloc_table(NI, D, dev, 0.01, g)
perms = [torch.randperm(NU, device=dev, generator=g)[:B].to(torch.int32).contiguous() for _ in range(2)]
trip = [engine.sample_triples(u, train, 7, k + 1) for k, u in enumerate(perms)]
loss = torch.zeros(1, dtype=torch.float64, device=dev)
mode = os.environ.get("MODE", "given")
for s in range(int(os.environ.get("STEPS", 6))):
    if mode == "sample":
        engine.bpr_step(U, V, D, perms[s % 2], csr=train, lr=0.0, reg=1e-4, flags=_lib.F_USERS_UNIQUE, seed=7, step=s + 1, loss_sum=loss)
    else:
        engine.bpr_step(U, V, D, perms[s % 2], trip[s % 2][0], trip[s % 2][1], lr=0.0, reg=1e-4, flags=_lib.F_USERS_UNIQUE, loss_sum=loss)
torch.cuda.synchronize()
print("done")
