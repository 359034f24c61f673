"""Where does the fused BPR step's time go?  (round 2, VERDICT weak-4)
Run on a B200:   B200REC_LIB=gpurun_build/libb200rec_abl.so python tools/probe_bpr_ablate.py
Times the cfg2 step (1M users x 100k items, d=128, B=1M) with parts of the fast kernel switched off (B200REC_ABL bits,
bpr_step.cu), with given / position-sorted triples, and with the item table pinned in L2."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recsys_pytorch_b200 import _lib, engine, synthetic

dev = torch.device("cuda:0")
NU, NI, D, B = 1_000_000, 100_000, 128, 1_000_000
train, _ = synthetic.make_interactions(NU, NI, seed=2020, device=dev)
g = torch.Generator(device=dev); g.manual_seed(1)
U = engine.alloc_table(NU, D, dev, 0.01, g); V = engine.alloc_table(NI, D, dev, 0.01, g)
perms = [torch.randperm(NU, device=dev, generator=g)[:B].to(torch.int32).contiguous() for _ in range(4)]
trip = []
for k, u in enumerate(perms):
    p, n = engine.sample_triples(u, train, 7, k + 1)
    o = torch.argsort(p.long() * NI + n.long())
    o2 = torch.argsort(n.long())
    trip.append(dict(u=u, p=p, n=n, us=u[o].contiguous(), ps=p[o].contiguous(), ns=n[o].contiguous(),
                     un=u[o2].contiguous(), pn=p[o2].contiguous(), nn=n[o2].contiguous()))
loss = torch.zeros(1, dtype=torch.float64, device=dev)
FL = _lib.F_USERS_UNIQUE


def run(mode, steps=20, flags=FL, lr=0.0):
    def one(s):
        t = trip[s % 4]
        if mode == "sample":
            engine.bpr_step(U, V, D, t["u"], csr=train, lr=lr, reg=1e-4, flags=flags, seed=7, step=s + 1, loss_sum=loss)
        elif mode == "given":
            engine.bpr_step(U, V, D, t["u"], t["p"], t["n"], lr=lr, reg=1e-4, flags=flags, loss_sum=loss)
        elif mode == "sorted_pos":
            engine.bpr_step(U, V, D, t["us"], t["ps"], t["ns"], lr=lr, reg=1e-4, flags=flags, loss_sum=loss)
        elif mode == "sorted_neg":
            engine.bpr_step(U, V, D, t["un"], t["pn"], t["nn"], lr=lr, reg=1e-4, flags=flags, loss_sum=loss)
    for s in range(5):
        one(s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(steps):
        one(s)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


out = {}
abl = "abl" in os.environ.get("B200REC_LIB", "")
for mode in ("sample", "given", "sorted_pos", "sorted_neg"):
    os.environ["B200REC_ABL"] = "0"
    out[mode] = run(mode)
    out[mode + "+hints"] = run(mode, flags=FL | _lib.F_L2_HINTS)
engine.l2_persist(V, 1.0)
for mode in ("sample", "given", "sorted_pos"):
    out[mode + "+persistV"] = run(mode)
engine.l2_persist(None)
if abl:
    names = {1: "item RED->st", 2: "no item writes", 4: "no user write", 6: "no writes at all", 8: "no neg row",
             16: "no pos row", 24: "no item rows", 28: "user row read only", 9: "no neg row, pos RED->st",
             32: "noop-bit (cost of the ablation branch itself)"}
    for bits, nm in names.items():
        os.environ["B200REC_ABL"] = str(bits)
        for mode in ("given", "sorted_pos"):
            out["%s | %s" % (mode, nm)] = run(mode)
    os.environ["B200REC_ABL"] = "0"
for k, v in out.items():
    print("%-60s %.4f ms  %.2f G triples/s" % (k, v, B / v / 1e6))
print(json.dumps(out))
