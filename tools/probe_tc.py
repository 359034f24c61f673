"""Timing probe for the scoring kernels (not a bench value): random tables vs a trained model."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recsys_pytorch_b200 import engine, synthetic, _lib

dev = torch.device("cuda")
nu, ni, d = int(os.environ.get("NU", 32768)), int(os.environ.get("NI", 100000)), int(os.environ.get("D", 128))


def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


g = torch.Generator(device=dev); g.manual_seed(0)
train, _ = synthetic.make_interactions(nu, ni, seed=1, device=dev)
users = torch.arange(nu, dtype=torch.int32, device=dev)
for tag, std_v in (("random N(0,.1)", None), ("item norms lognormal x(0.3..3)", 0.5)):
    U = engine.alloc_table(nu, d, dev, 0.1, g); V = engine.alloc_table(ni, d, dev, 0.1, g)
    if std_v:
        V *= torch.exp(torch.randn(ni, 1, device=dev, generator=g) * std_v)
    for k in (10, 100):
        for mask in (None, train):
            only = os.environ.get("ONLY")          # e.g. ONLY="item:10:yes"
            if only and only != "%s:%d:%s" % (tag[:4], k, "yes" if mask is not None else "no"):
                continue
            os.environ["B200REC_TC_STATS"] = "1"
            it, _ = engine.score_topk(U, V, d, users, mask, k, algo=_lib.SCORE_TC)
            os.environ.pop("B200REC_TC_STATS")
            t_tc = timeit(lambda: engine.score_topk(U, V, d, users, mask, k, algo=_lib.SCORE_TC))
            t_ex = timeit(lambda: engine.score_topk(U, V, d, users, mask, k, algo=_lib.SCORE_EXACT), 1)
            ie, _ = engine.score_topk(U, V, d, users, mask, k, algo=_lib.SCORE_EXACT)
            print(f"{tag} k={k} mask={'yes' if mask is not None else 'no'}: tc {t_tc:.2f} ms ({nu*ni/t_tc/1e6:.1f} Gpairs/s) "
                  f"exact {t_ex:.2f} ms  equal={torch.equal(it, ie)}", flush=True)
