"""Single-GPU timing of the pieces of a user-sharded step (not a bench value)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recsys_pytorch_b200 import engine, synthetic, _lib
from recsys_pytorch_b200.dist import UserShardedBPR

dev = torch.device("cuda")
nu, ni, d, B = 1_000_000, 100_000, 128, 1_000_000
train, _ = synthetic.make_interactions(nu, ni, seed=1, device=dev)
tr = UserShardedBPR(nu, ni, d, train, 0, 1, dev, lr=5e4, reg=1e-4)
g = torch.Generator(device=dev); g.manual_seed(0)
users = torch.randperm(nu, device=dev, generator=g).to(torch.int32)


def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


k = [0]
def kern_delta():
    k[0] += 1
    engine.bpr_step(tr.U, tr.V, d, users, csr=train, lr=tr.lr, reg=tr.reg, flags=_lib.F_ITEM_DELTA | _lib.F_USERS_UNIQUE,
                    seed=1, step=k[0], gV=tr.dV)
def kern_inplace_generic():
    k[0] += 1
    engine.bpr_step(tr.U, tr.V, d, users, csr=train, lr=tr.lr, reg=tr.reg, flags=_lib.F_GENERIC | _lib.F_USERS_UNIQUE, seed=1, step=k[0])
def kern_inplace_fast():
    k[0] += 1
    engine.bpr_step(tr.U, tr.V, d, users, csr=train, lr=tr.lr, reg=tr.reg, flags=_lib.F_USERS_UNIQUE, seed=1, step=k[0])
def kern_delta_hints():
    k[0] += 1
    engine.bpr_step(tr.U, tr.V, d, users, csr=train, lr=tr.lr, reg=tr.reg,
                    flags=_lib.F_ITEM_DELTA | _lib.F_USERS_UNIQUE | _lib.F_L2_HINTS, seed=1, step=k[0], gV=tr.dV)
def kern_inplace_fast_hints():
    k[0] += 1
    engine.bpr_step(tr.U, tr.V, d, users, csr=train, lr=tr.lr, reg=tr.reg, flags=_lib.F_USERS_UNIQUE | _lib.F_L2_HINTS, seed=1, step=k[0])
print("kernel, item deltas -> dV  : %.3f ms" % timeit(kern_delta))
print("kernel, deltas -> dV, hints: %.3f ms" % timeit(kern_delta_hints))
print("kernel, in place fast+hints: %.3f ms" % timeit(kern_inplace_fast_hints))
print("kernel, in place (generic) : %.3f ms" % timeit(kern_inplace_generic))
print("kernel, in place (fast)    : %.3f ms" % timeit(kern_inplace_fast))
print("dV.zero_()                 : %.3f ms" % timeit(lambda: tr.dV.zero_()))
print("apply V += dV              : %.3f ms" % timeit(lambda: tr.apply_item_delta(tr.dV)))
print("zero + kernel(dV) + apply  : %.3f ms" % timeit(lambda: (tr.dV.zero_(), kern_delta(), tr.apply_item_delta(tr.dV))))
