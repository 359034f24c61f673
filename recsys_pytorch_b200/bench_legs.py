"""Secondary workloads of bench.py (BASELINE configs[3] and [4]) - reported as extra keys of the ONE JSON line.

cfg5  predict() U.V^T + top-100 at 10M users x 1M items, d in {64, 128, 256}: users are independent, so N GPUs each
      take 10M/N of them (strong scaling).  Timed on a bounded slice schedule - `chunks` launches' worth of users per
      rank (75,776 rows per tensor-core launch) through the public `engine.score_topk` call, whole call including the
      item-side pre-pass, the exact fp32 re-rank and the CSR mask - and projected to the full 10M users.
cfg4  LightGCN 1M x 100k, 3 layers, d=64: the CSR SpMM propagation (HBM roofline, algorithmic bytes
      nnz(A)(8+4d) + N 4d per layer) and one full training step.
"""
from __future__ import annotations

import time
import types

import numpy as np
import torch


def _median_ms(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    out = []
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1))
    return float(np.median(out))


def cfg5_leg(dev, rank, world, bf16_tf, peak_src, chunks=8, k=100, num_users=10_000_000, num_items=1_000_000,
             dims=(64, 128, 256), small=False):
    from . import engine, synthetic
    from ._lib import SCORE_TC
    import torch.distributed as dist
    if small:
        num_users, num_items, chunks = 400_000, 50_000, 1
    rows = 75_776 * chunks
    per_rank_users = num_users // world
    n = min(rows, per_rank_users)
    # the slice's mask rows: the synthetic interaction generator at the full catalogue width
    train, _ = synthetic.make_interactions(n, num_items, seed=77 + rank, device=dev, item_seed=2020)
    users = torch.arange(n, dtype=torch.int32, device=dev)
    g = torch.Generator(device=dev); g.manual_seed(5 + rank)
    gi = torch.Generator(device=dev); gi.manual_seed(5)
    res = {}
    for d in dims:
        # "trained-like" tables: N(0, 0.1) directions with log-normal row norms (a trained MF model's item norms follow
        # popularity; random iso-norm tables are the adversarial case for the norm-ordered candidate pass)
        U = engine.alloc_table(n, d, dev, 0.1, g)
        V = engine.alloc_table(num_items, d, dev, 0.1, gi)
        V[:, :d] *= torch.exp(0.5 * torch.randn(num_items, 1, device=dev, generator=gi))
        U[:, :d] *= torch.exp(0.3 * torch.randn(n, 1, device=dev, generator=g))
        ms = _median_ms(lambda: engine.score_topk(U, V, d, users, train, k, algo=SCORE_TC, want_scores=False))
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        pairs = float(n) * num_items * world                 # all ranks score their slice in that time
        tf = pairs * 2 * d / (ms * 1e-3) / 1e12
        res["d=%d" % d] = {"ms_slice": ms, "scored_pairs_per_sec": pairs / (ms * 1e-3),
                           "tflops": tf, "frac_of_bf16_peak": tf / (bf16_tf * world),
                           "projected_full_sweep_s": (num_users / world) / n * ms * 1e-3}
        del U, V
    return {"workload": "predict() U.V^T + top-%d, %d users x %d items (BASELINE configs[4]), users sharded over %d GPU(s)"
                        % (k, num_users, num_items, world),
            "k": k, "slice_users_per_rank": n, "slice": "%d tensor-core launches of 75,776 rows per rank, whole score_topk "
            "call (item pre-pass + tcgen05 candidate pass + exact fp32 re-rank + CSR mask), median of 3" % chunks,
            "tables": "synthetic, log-normal row norms", "mask_nnz_per_user": train.nnz / max(n, 1),
            "peak_tflops_per_gpu": bf16_tf, "peak_source": peak_src, "results": res}


def cfg4_leg(dev, hbm_gbs, peak_src, nu=1_000_000, ni=100_000, d=64, L=3, B=1_000_000, small=False):
    from . import synthetic
    from .lightgcn import LightGCN
    if small:
        nu, ni, B = 50_000, 20_000, 50_000
    train, target = synthetic.make_interactions(nu, ni, seed=2020, device=dev)
    ds = types.SimpleNamespace(num_users=nu, num_items=ni, train_data=train, valid_input=train, valid_target=target,
                               protocol="holdout", dataname="synthetic")
    m = LightGCN(ds, {"emb_dim": d, "num_layers": L, "optimizer": "adam", "lr": 1e-3, "score_algo": "tc"}, dev)
    t0 = time.perf_counter(); m.Graph = m.getSparseGraph(train); torch.cuda.synchronize(); t_graph = time.perf_counter() - t0
    nnzA, N = int(m.Graph[1].numel()), nu + ni
    ms_prop = _median_ms(lambda: m.propagate(m.E0, m.out), reps=5, warm=2)
    g = torch.Generator(device=dev); g.manual_seed(0)
    users = torch.randperm(nu, device=dev, generator=g)[:B].to(torch.int32).contiguous()
    key = [0]

    def step():
        key[0] += 1
        m.train_batch(users, csr=train, step_key=key[0], users_unique=True)
    ms_step = _median_ms(step, reps=3, warm=1)
    bytes_layer = nnzA * (8 + 4 * d) + N * 4 * d              # no-reuse gather model
    achieved = bytes_layer * L / (ms_prop * 1e-3) / 1e9
    return {"workload": "LightGCN synthetic %dx%d L=%d d=%d (BASELINE configs[3])" % (nu, ni, L, d), "nnz_adj": nnzA,
            "graph_build_s": t_graph, "propagate_ms": ms_prop, "ms_per_layer": ms_prop / L,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_gbs, "unit": "GB/s", "frac": achieved / hbm_gbs,
                         "traffic": None, "peak_source": peak_src, "kernel": "spmm_csr_kernel + spmm_segment_kernel",
                         "bytes_per_layer": bytes_layer, "compulsory_bytes_per_layer": nnzA * 8 + 2 * N * 4 * d},
            "train_step_ms": ms_step, "train_triples_per_s": B / (ms_step * 1e-3), "batch_triples": B,
            "step_anatomy": "propagate fwd (3 SpMM) + fused BPR step (SINK_GRAD) + propagate bwd (3 SpMM) + dense Adam"}
