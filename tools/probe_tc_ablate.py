"""Ablation probe of the tcgen05 candidate kernel (diagnostics, not a bench value).

Trains the bench model for a few steps (same config as bench.py), then times the candidate kernel alone
(CUDA events inside the library, B200REC_TC_TIME=1) with parts of it switched off:
  ablate 0  full kernel                      1  filter but never append
         2  TMEM loads only (no filter)      4  no TMEM loads (MMA + TMA pipeline only)
        12  no TMEM loads and no TMA (MMA issue rate alone)
for every kernel variant selected by B200REC_TC_KERNEL.
"""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recsys_pytorch_b200 import engine, synthetic, _lib
from recsys_pytorch_b200.mf import MF

dev = torch.device("cuda")
nu, ni, d = int(os.environ.get("NU", 1_000_000)), int(os.environ.get("NI", 100_000)), int(os.environ.get("D", 128))
ne = int(os.environ.get("NE", 32768))
k = int(os.environ.get("K", 10))
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
ref, _ = engine.score_topk(model.U, model.V, d, users, train, k, algo=_lib.SCORE_EXACT)
os.environ["B200REC_TC_TIME"] = os.environ.get("TCTIME", "1")
for kern in os.environ.get("KERNELS", "0,1,2").split(","):
    os.environ["B200REC_TC_KERNEL"] = kern
    for mask in (train, None):
        for abl in os.environ.get("ABLATE", "0,0,1,2,4,12").split(","):
            os.environ["B200REC_TC_ABLATE"] = abl
            sys.stderr.write(f"kernel={kern} mask={'yes' if mask is not None else 'no'} ablate={abl}: "); sys.stderr.flush()
            try:
                it, _ = engine.score_topk(model.U, model.V, d, users, mask, k, algo=_lib.SCORE_TC)
                torch.cuda.synchronize()
                if abl == "0" and mask is not None:
                    sys.stderr.write(f"    equal to exact: {torch.equal(it, ref)}\n")
            except Exception as e:  # noqa: BLE001
                sys.stderr.write(f"FAILED {e}\n")
        if kern != "0":
            break
