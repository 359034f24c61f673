"""Group-kernel variants of the fused BPR step at cfg2 (B200REC_STEP_VARIANT="G,PF"), sample / given triples."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recsys_pytorch_b200 import _lib, engine, synthetic
dev = torch.device("cuda:0")
NU, NI, D, B = 1_000_000, int(os.environ.get("NI", 100_000)), 128, 1_000_000
train, _ = synthetic.make_interactions(NU, NI, seed=2020, device=dev)
g = torch.Generator(device=dev); g.manual_seed(1)
U = engine.alloc_table(NU, D, dev, 0.01, g); V = engine.alloc_table(NI, D, dev, 0.01, g)
perms = [torch.randperm(NU, device=dev, generator=g)[:B].to(torch.int32).contiguous() for _ in range(4)]
trip = [engine.sample_triples(u, train, 7, k + 1) for k, u in enumerate(perms)]
loss = torch.zeros(1, dtype=torch.float64, device=dev)


def run(mode, steps=20, lr=0.0, uniq=True, with_loss=True, hints=False):
    fl = (_lib.F_USERS_UNIQUE if uniq else 0) | (_lib.F_L2_HINTS if hints else 0)
    def one(s):
        if mode == "sample":
            engine.bpr_step(U, V, D, perms[s % 4], csr=train, lr=lr, reg=1e-4, flags=fl, seed=7, step=s + 1, loss_sum=loss if with_loss else None)
        else:
            engine.bpr_step(U, V, D, perms[s % 4], trip[s % 4][0], trip[s % 4][1], lr=lr, reg=1e-4, flags=fl, loss_sum=loss if with_loss else None)
    for s in range(5):
        one(s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(steps):
        one(s)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


for var in os.environ.get("VARIANTS", "0,0 32,2 16,1 16,2 8,0 8,1 4,0").split():
    os.environ["B200REC_STEP_VARIANT"] = var
    r = {m: run(m) for m in ("sample", "given")}
    r["sample_nouniq"] = run("sample", uniq=False)
    r["sample_h"] = run("sample", hints=True); r["given_h"] = run("given", hints=True)
    print("variant %-6s sample %.4f ms (%.2f G/s)  given %.4f ms (%.2f G/s)  sample,red-users %.4f ms" %
          (var, r["sample"], B / r["sample"] / 1e6, r["given"], B / r["given"] / 1e6, r["sample_nouniq"]), flush=True)
    print("               +hints: sample %.4f ms  given %.4f ms" % (r["sample_h"], r["given_h"]), flush=True)
