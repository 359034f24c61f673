"""cfg5 slice (BASELINE configs[4]: predict() U.V^T + top-100, 10M x 1M, d in {64,128,256}): ONE chunk of 75 776 users
against the full 1M-item catalogue per d, random-init tables plus a popularity-skewed variant.  Diagnostics."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recsys_pytorch_b200 import engine, synthetic, _lib
dev = torch.device("cuda")
nu, ni, k = int(os.environ.get("NU", 75776)), int(os.environ.get("NI", 1_000_000)), int(os.environ.get("K", 100))
train, _ = synthetic.make_interactions(nu, ni, seed=7, device=dev)
users = torch.arange(nu, dtype=torch.int32, device=dev)
g = torch.Generator(device=dev); g.manual_seed(0)
for d in (64, 128, 256):
    for tag in ("random-init N(0,0.01)", "norms lognormal(0.5)"):
        U = engine.alloc_table(nu, d, dev, 0.01, g); V = engine.alloc_table(ni, d, dev, 0.01, g)
        if tag.startswith("norms"):
            V *= torch.exp(torch.randn(ni, 1, device=dev, generator=g) * 0.5)
        os.environ["B200REC_TC_STATS"] = "1"
        engine.score_topk(U, V, d, users, train, k, algo=_lib.SCORE_TC); torch.cuda.synchronize()
        os.environ.pop("B200REC_TC_STATS")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); it, _ = engine.score_topk(U, V, d, users, train, k, algo=_lib.SCORE_TC); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        # spot check 256 rows against the exact kernel
        sub = users[:256].contiguous()
        ie, _ = engine.score_topk(U, V, d, sub, train, k, algo=_lib.SCORE_EXACT)
        print(f"d={d} k={k} {tag}: {ms:.2f} ms  {nu * ni / ms / 1e6:.0f} Gpairs/s  {2 * d * nu * ni / ms / 1e9:.0f} TFLOP/s  "
              f"equal(256 rows)={torch.equal(it[:256], ie)}", flush=True)
        del U, V
