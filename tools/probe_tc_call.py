"""One score_topk(TC) call on the trained bench model, for launch lists (ncu) and whole-call timing."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recsys_pytorch_b200 import engine, synthetic, _lib
from recsys_pytorch_b200.mf import MF

dev = torch.device("cuda")
nu, ni, d = int(os.environ.get("NU", 1_000_000)), int(os.environ.get("NI", 100_000)), int(os.environ.get("D", 128))
ne, k = int(os.environ.get("NE", 37888)), int(os.environ.get("K", 10))
train, target = synthetic.make_interactions(nu, ni, seed=2020, device=dev)
ds = types.SimpleNamespace(num_users=nu, num_items=ni, train_data=train, valid_input=train, valid_target=target,
                           protocol="holdout", dataname="synthetic")
hp = {"hidden_dim": d, "pointwise": False, "loss_func": "ce", "optimizer": "sgd", "lr": 0.05 * 1_000_000, "reg": 1e-4,
      "init_std": 0.01, "gather": "ldg", "seed": 2020, "score_algo": "tc"}
model = MF(ds, hp, dev)
g = torch.Generator(device=dev); g.manual_seed(2020)
B = min(nu, 1_000_000)
perm = torch.randperm(nu, device=dev, generator=g)[:B].to(torch.int32).contiguous()
for s in range(int(os.environ.get("TRAIN_STEPS", 23))):
    engine.bpr_step(model.U, model.V, d, perm, csr=train, lr=hp["lr"], reg=hp["reg"], sink=_lib.SINK_UPDATE,
                    flags=_lib.GATHER_FLAGS["ldg"] | _lib.F_USERS_UNIQUE, seed=2020, step=s + 1)
torch.cuda.synchronize()
users = torch.arange(ne, dtype=torch.int32, device=dev)
mask = train if os.environ.get("MASK", "1") == "1" else None
for rep in range(int(os.environ.get("REPS", 3))):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    it, sc = engine.score_topk(model.U, model.V, d, users, mask, k, algo=_lib.SCORE_TC)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"score_topk TC {ne}x{ni} d={d} k={k}: {ms:.3f} ms  {ne * ni / ms / 1e6:.1f} Gpairs/s", flush=True)
