#!/usr/bin/env python
"""bench.py - BPR-MF hot path on B200 (BASELINE.json metric: BPR triples/s for
training, scored pairs/s + NDCG@10 for evaluation).

    python bench.py --gpus 1 --steps 20 --warmup 3          # this repo's CUDA path
    python bench.py --impl reference --steps 5 --warmup 1   # the reference's CPU path (torch restatement)
    torchrun --nproc-per-node N bench.py --gpus N ...       # one rank per GPU

A *step* is one pass of the hot path over one batch of synthetic triples at the
configuration BASELINE.json's metric is quoted on (configs[1]: BPRMF synthetic
1M users x 100k items, d=128): ONE fused kernel launch that samples (pos, neg) on
the device for a batch of B users, gathers the three rows, and scatters the SGD
update.  Inputs are larger than L2 (the step walks B user rows of 512 B = 512 MB at
B=1M, L2 is 126 MB), so no L2 flush is needed between timed iterations.

Prints ONE JSON line (rank 0).  `value` is device-resident throughput; `e2e` is the
same metric through the public plugin API (`MF.train_batch`) with the batch's user
ids coming from pinned HOST memory and the loss read back to the host every step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

CFG = dict(num_users=1_000_000, num_items=100_000, d=128, batch=1_000_000, seed=2020, lr=0.05 * 1_000_000, reg=1e-4,
           lr_per_triple=0.05,                            # lr / batch: what one triple moves its rows by (same for every N)
           init_std=0.01, eval_users=37_888, eval_k=10)   # 148 SMs x 256 rows: one full wave of the scoring kernel
FALLBACK_HBM_GBS = 6650.0


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            return float(j["hbm_gbs"]), float(j.get("bf16_tflops", 1590.0)), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, 1590.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe).

    nvidia-smi needs ~0.1-0.3 s to deliver its first row and samples every 20 ms, while K steps of this kernel last a
    few milliseconds; `timed_under_load` therefore keeps the GPU running the very same step (learning rate 0) before
    and after the timed K steps, and only rows stamped inside that busy window are used."""

    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t_lo=None, t_hi=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        # a row is printed right after it is sampled: keep the rows that arrived inside the busy window
        rows = [r for (t, r) in self.rows if len(r) >= 8 and (t_lo is None or t_lo <= t <= t_hi)]

        def num(x):
            try:
                return float(x)
            except ValueError:
                return None
        sm = [v for v in (num(r[1]) for r in rows) if v is not None]
        mx = [v for v in (num(r[2]) for r in rows) if v is not None]
        pw = [v for v in (num(r[3]) for r in rows) if v is not None]
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(pw) if pw else None,
                "window": "GPU kept busy with the same step (lr=0) before and after the timed K steps; rows inside it"}


def dist_env():
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def timed_region(fn, steps, world):
    """barrier + sync | K steps between CUDA events on the launching stream | sync + barrier; max over ranks."""
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(steps):
        fn(s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        dist.barrier()
    return ms


def timed_under_load(step, idle_step, steps, world, gpu_index, rank=0, pre_steps=900, post_steps=250):
    """`timed_region(step, K)` bracketed by `pre_steps` / `post_steps` launches of `idle_step` (the same kernel on the
    same inputs with learning rate 0: identical traffic, tables unchanged) so that the clock sampler sees the GPU under
    this load on both sides of the timed region.  The counts are fixed so every rank issues the same collectives."""
    clocks = ClockSampler(gpu_index)
    if rank == 0:
        clocks.start()
    torch.cuda.synchronize()
    t_lo = time.time()
    for s in range(pre_steps):
        idle_step(s)
    ms = timed_region(step, steps, world)
    for s in range(post_steps):
        idle_step(s)
    torch.cuda.synchronize()
    t_hi = time.time()
    return ms, (clocks.stop(t_lo + 0.02, t_hi) if rank == 0 else None)


# ------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import types
    from recsys_pytorch_b200 import _lib, engine, synthetic
    from recsys_pytorch_b200.evaluation import Evaluator
    from recsys_pytorch_b200.mf import MF

    rank, world, local = dist_env()
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: the engine has no CPU path"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        # NCCL announces its version on STDOUT when the first communicator comes up; the contract is ONE JSON
        # line on stdout, so that banner is sent to stderr
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    c = dict(CFG)
    if args.small:
        c.update(num_users=50_000, num_items=20_000, batch=50_000, eval_users=4096)
    hbm_gbs, bf16_tf, peak_src = measured_peaks()
    args.bf16_tf = bf16_tf

    if world > 1:
        c["small"] = bool(args.small)
        if args.layout == "p2p":
            from recsys_pytorch_b200 import p2p
            return p2p.bench_p2p(args, c, rank, world, dev, timed_region, timed_under_load, hbm_gbs, peak_src)
        from recsys_pytorch_b200 import dist as bdist
        return bdist.bench_multi_gpu(args, c, rank, world, dev, timed_region, timed_under_load, hbm_gbs, peak_src)

    train, target = synthetic.make_interactions(c["num_users"], c["num_items"], seed=c["seed"], device=dev)
    ds = types.SimpleNamespace(num_users=c["num_users"], num_items=c["num_items"], train_data=train,
                               valid_input=train, valid_target=target, protocol="holdout", dataname="synthetic")
    hp = {"hidden_dim": c["d"], "pointwise": False, "loss_func": "ce", "optimizer": "sgd", "lr": c["lr"],
          "reg": c["reg"], "init_std": c["init_std"], "gather": args.gather, "seed": c["seed"],
          "score_algo": args.score_algo}
    model = MF(ds, hp, dev)
    B, d, ld = c["batch"], c["d"], model.U.shape[1]
    g = torch.Generator(device=dev); g.manual_seed(c["seed"])
    n_perm = 4
    perms = [torch.randperm(c["num_users"], device=dev, generator=g)[:B].to(torch.int32).contiguous() for _ in range(n_perm)]
    loss = torch.zeros(1, dtype=torch.float64, device=dev)
    flags = _lib.GATHER_FLAGS[args.gather] | _lib.F_USERS_UNIQUE

    def step_dev(s):
        engine.bpr_step(model.U, model.V, d, perms[s % n_perm], csr=train, lr=c["lr"], reg=c["reg"],
                        sink=_lib.SINK_UPDATE, flags=flags, seed=c["seed"], step=s + 1, loss_sum=loss)

    scratch_loss = torch.zeros(1, dtype=torch.float64, device=dev)

    def step_idle(s):      # same kernel, same traffic, learning rate 0 (tables unchanged): clock-sampling window only
        engine.bpr_step(model.U, model.V, d, perms[s % n_perm], csr=train, lr=0.0, reg=c["reg"],
                        sink=_lib.SINK_UPDATE, flags=flags, seed=c["seed"], step=100000 + s, loss_sum=scratch_loss)

    for s in range(args.warmup):
        step_dev(s)
    counted = [0, 0]

    def step_counted(s):   # the library counts every kernel it launches: read the counter at both ends of the region
        if s == 0:
            counted[0] = _lib.launch_count()
        step_dev(s)
        if s == args.steps - 1:
            counted[1] = _lib.launch_count()
    ms, clk = timed_under_load(step_counted, step_idle, args.steps, world, local)
    launches = counted[1] - counted[0]
    # median of 3 timed repeats of the K steps (the timed region is ~8 ms: one repeat is at the mercy of a stray stall)
    repeats_ms = [ms] + [timed_region(lambda s, r=r: step_dev(args.steps * (r + 1) + s), args.steps, world) for r in range(2)]
    ms = float(np.median(repeats_ms))
    triples_per_s = B * args.steps / (ms * 1e-3)
    ms_per_step = ms / args.steps

    # ---- e2e: plugin API, every step's user ids come from pinned host memory (H2D inside the timed region) and
    # every step's loss is read back to the host (D2H); MF.train_batch_async double-buffers the copies so the
    # H2D of step s+1 and the read-back of step s-1 overlap the kernel of step s ----
    host_perms = [p.cpu().pin_memory() for p in perms]
    e2e_losses = []
    pend = [None]

    def step_e2e(s):
        h = model.train_batch_async(host_perms[s % n_perm], csr=train, step_key=1000 + s, users_unique=True)
        if pend[0] is not None:
            e2e_losses.append(pend[0].loss())                                      # host reads step s-1's result
        pend[0] = h

    def drain():
        if pend[0] is not None:
            e2e_losses.append(pend[0].loss()); pend[0] = None

    for s in range(min(args.warmup, 3)):
        step_e2e(s)
    drain()
    e2e_losses.clear()
    ms_e2e = timed_region(lambda s: (step_e2e(s), drain() if s == args.steps - 1 else None), args.steps, world)
    assert len(e2e_losses) == args.steps and all(np.isfinite(e2e_losses))
    e2e = {"value": B * args.steps / (ms_e2e * 1e-3), "unit": "triples/s", "h2d_bytes_per_step": B * 4,
           "d2h_bytes_per_step": 8, "ms_per_step": ms_e2e / args.steps,
           "api": "recsys_pytorch_b200.mf.MF.train_batch_async (users from pinned host memory; every step's loss "
                  "read back, one step late)"}

    # ---- evaluation leg: fused score + mask + top-k + holdout metrics on a user sample ----
    ev_users = torch.arange(c["eval_users"], dtype=torch.int32, device=dev)
    ev = Evaluator(train, _SubsetTarget(target, c["eval_users"]), protocol="holdout", ks=[c["eval_k"]])
    ev.evaluate(model)                                                              # warm-up
    t_evals, t_devs = [], []
    for _ in range(5):                                                              # median of 5
        torch.cuda.synchronize(); t0 = time.perf_counter()
        scores = ev.evaluate(model)
        torch.cuda.synchronize(); t_evals.append(time.perf_counter() - t0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); idx, _ = model.predict_topk_device(ev_users, train, c["eval_k"]); e1.record(); torch.cuda.synchronize()
        t_devs.append(e0.elapsed_time(e1) * 1e-3)
    t_eval, t_dev = float(np.median(t_evals)), float(np.median(t_devs))
    pairs = c["eval_users"] * c["num_items"]
    eval_leg = {"scored_pairs_per_sec": pairs / t_dev, "e2e_pairs_per_sec": pairs / t_eval,
                "ndcg@%d" % c["eval_k"]: float(scores["NDCG@%d" % c["eval_k"]]), "users": c["eval_users"],
                "k": c["eval_k"], "algo": args.score_algo,
                "flops_per_pair": 2 * d,
                # whole predict_topk call (pre-pass + tcgen05 candidate kernel + fp32 re-rank) against the measured
                # dense bf16 peak; the candidate kernel's own tensor-pipe % is in profiles/ (ncu)
                "tensor_roofline": {"bound": "tensor", "achieved": pairs * 2 * d / t_dev / 1e12, "peak": bf16_tf,
                                    "unit": "TFLOP/s", "frac": pairs * 2 * d / t_dev / 1e12 / bf16_tf,
                                    "peak_source": peak_src}}

    # ---- roofline of the dominant kernel (fused BPR step): algorithmic bytes = 24d+8 per triple ----
    bytes_per_triple = 24 * d + 8
    achieved = bytes_per_triple * B / (ms_per_step * 1e-3) / 1e9
    # roofline.traffic: DRAM bytes of ONE launch from the committed `ncu --set full` capture - valid only for the kernel
    # that capture profiled, so it is tied to the kernel this run actually dispatched to
    traffic, traffic_note = None, None
    ran = _lib.last_step_kernel()
    prof = os.path.join(ROOT, "profiles", "bpr_step_dram_bytes.json")
    if os.path.exists(prof):
        pj = json.load(open(prof))
        if pj.get("kernel") == ran:
            traffic, traffic_note = pj.get("dram_bytes_per_launch"), pj.get("source")
        else:
            traffic_note = "no capture of %s committed (profiles/bpr_step_dram_bytes.json is for %s)" % (ran, pj.get("kernel"))
            # loud, not fatal: the line keeps its measured numbers, `traffic` stays null and says why
            sys.stderr.write("bench.py: WARNING " + traffic_note + " - re-capture the default kernel with ncu\n")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_gbs, "unit": "GB/s", "frac": achieved / hbm_gbs,
                "traffic": traffic, "traffic_source": traffic_note, "peak_source": peak_src,
                "bytes_per_triple": bytes_per_triple, "kernel": ran,
                "note": "the step is bound by the SM<->L2 fabric (3.04 GB of sectors per launch = the algorithmic bytes; "
                        "~7.8 TB/s), not by HBM: the item rows hit L2, DRAM sees `traffic` (profiles/r02_bpr_ablation.md)"}

    cpu_base = None
    if not args.no_cpu:
        # the (u,i,j) batches the device sampler itself draws, read back for the CPU arm
        try:
            gpu_triples = {}
            for bsz in (65_536 if not args.small else 8192, 256):
                lst = []
                for b in range(4):
                    u = perms[b % n_perm][:bsz].contiguous()
                    p_, n_ = engine.sample_triples(u, train, c["seed"], 9000 + b)
                    lst.append(tuple(t.cpu().long() for t in (u, p_, torch.clamp(n_, min=0))))
                gpu_triples[bsz] = lst
            cpu_base = cpu_baseline_leg(c, args, gpu_triples)
        except Exception as e:                                       # pragma: no cover - reported, never fatal
            cpu_base = {"value": None, "unit": "triples/s", "cores": None, "kind": "port", "error": repr(e)[:300]}
        # scores and loss of the SAME (u,i,j) batch on the trained tables, CUDA kernel vs the
        # reference's torch CPU ops (models/MF.py:38-42,99-105)
        try:
            u_, p_, n_ = gpu_triples[max(gpu_triples)][0]
            xg = torch.empty(u_.numel(), dtype=torch.float32, device=dev)
            lg = torch.zeros(1, dtype=torch.float64, device=dev)
            engine.bpr_step(model.U, model.V, d, *(t.to(dev, torch.int32) for t in (u_, p_, n_)), sink=_lib.SINK_NONE,
                            loss_sum=lg, x_out=xg)
            Uc, Vc = model.U[:, :d].cpu(), model.V[:, :d].cpu()
            xc = torch.sum(Uc[u_] * Vc[p_], 1) - torch.sum(Uc[u_] * Vc[n_], 1)
            lc = float(-torch.sigmoid(xc).log().mean())
            fin = lambda v: float(v) if np.isfinite(float(v)) else None   # json has no inf / nan
            cpu_base["parity_same_triples"] = {
                "triples": int(u_.numel()), "max_abs_err_score_diff": fin((xg.cpu() - xc).abs().max()),
                "max_abs_score_diff": fin(xc.abs().max()), "loss_gpu": fin(lg.item() / u_.numel()), "loss_cpu": fin(lc),
                "abs_err_loss": fin(abs(lg.item() / u_.numel() - lc)),
                "what": "x = s(u,i) - s(u,j) and -mean log sigmoid(x) of one device-sampled batch on the tables as trained "
                        "by this run: fused kernel (forward-only sink) vs torch CPU fp32"}
        except Exception as e:                                       # pragma: no cover
            if isinstance(cpu_base, dict):
                cpu_base["parity_same_triples"] = {"error": repr(e)[:300]}
    # secondary workloads: a failure there is reported under its key, it never costs the headline line
    legs = {}
    if not args.no_legs:
        from recsys_pytorch_b200 import bench_legs
        del model, ev
        for name, leg in (("lightgcn_cfg4", lambda: bench_legs.cfg4_leg(dev, hbm_gbs, peak_src, small=args.small)),
                          ("eval_cfg5", lambda: bench_legs.cfg5_leg(dev, 0, 1, bf16_tf, peak_src, small=args.small))):
            torch.cuda.empty_cache()
            try:
                legs[name] = leg()
            except Exception as e:                                   # pragma: no cover
                legs[name] = {"error": repr(e)[:300]}
    out = {"metric": "BPR triples/sec (train)", "value": triples_per_s, "unit": "triples/s", "n_gpus": 1,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "BPRMF synthetic %dx%d d=%d, 1xB200 fused kernel (BASELINE configs[1])" %
                      (c["num_users"], c["num_items"], d), "batch_triples": B, "optimizer": "sgd+l2", "lr": c["lr"],
                      "reg": c["reg"], "sampler": "on-device uniform negative vs CSR", "gather": args.gather,
                      "l2_policy": "inputs larger than L2 (512 MB of user rows per step)", "nnz_train": train.nnz,
                      "timing": "median of 3 repeats of the K steps", "repeats_ms": repeats_ms},
           "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_base,
           "eval": eval_leg, "final_loss": float(loss.item()) / (B * (3 * args.steps + args.warmup))}
    out.update(legs)
    print(json.dumps(out))


class _SubsetTarget:
    """First n rows of a DeviceCSR presented as the dict-free target the Evaluator accepts."""

    def __new__(cls, csr, n):
        import scipy.sparse as sp
        indptr = csr.indptr[: n + 1].cpu().numpy()
        indices = csr.indices[: int(indptr[-1])].cpu().numpy()
        return sp.csr_matrix((np.ones(len(indices), np.float32), indices, indptr), shape=(n, csr.shape[1]))


# ------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the reference's torch CPU path restated (oracle/torch_port.py)
# ------------------------------------------------------------------------------------------------
def _host_threads():
    """All host cores for the CPU arm - torchrun exports OMP_NUM_THREADS=1, which made round 1's N>1 reference
    numbers 7.5x too slow."""
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        pass
    torch.set_num_threads(max(n, 1))
    return torch.get_num_threads()


def _cpu_csr(c, n_rows, seed):
    """Synthetic interactions (same recipe as the GPU arm) for `n_rows` users, on the host."""
    from recsys_pytorch_b200 import synthetic
    tr, tg = synthetic.make_interactions_raw(n_rows, c["num_items"], seed=seed, device="cpu")
    return ((tr[0].numpy(), tr[1].numpy()), (tg[0].numpy(), tg[1].numpy()))


def _cpu_batches(c, csr, batch, n_batches, seed):
    """(u, i, j) batches drawn by the HOST MIRROR of the device sampler (oracle.bpr_oracle.sample_triples_vec: the same
    counter RNG, bit-identical to the kernel's draws for the same CSR, tests/test_gpu_parity.py): users are random rows
    of the full-size table, their positives / negatives come from the CSR rows of a host-generated user sample."""
    from oracle import bpr_oracle as O
    indptr, indices = csr
    n_rows = len(indptr) - 1
    rng = np.random.default_rng(seed)
    out = []
    for b in range(n_batches):
        rows = rng.permutation(n_rows)[:batch] if batch <= n_rows else rng.integers(0, n_rows, batch)
        pos, neg = O.sample_triples_vec(c["seed"], b + 1, rows, indptr, indices, c["num_items"])
        neg = np.where(neg < 0, 0, neg)
        users = rng.permutation(c["num_users"])[:batch]
        out.append(tuple(torch.from_numpy(np.ascontiguousarray(a, np.int64)) for a in (users, pos, neg)))
    return out


def _cpu_train_legs(c, batches_by_size, steps, warmup, table_sizes):
    """reference step (MF.py:64-68) timed as-is (dense Adam) and with the parity optimiser (SGD swap), at
    every batch size given.  Returns {leg: triples/s}."""
    from oracle import torch_port as TP
    nu, ni = table_sizes
    legs = {}
    for opt in ("adam", "sgd"):
        model = TP.RefMF(nu, ni, c["d"], init_std=c["init_std"], optimizer=opt, lr=1e-3 if opt == "adam" else 0.05)
        for batch, batches in batches_by_size.items():
            sec = TP.time_train(model, batches[:steps + warmup], warmup)
            legs["%s_b%d" % (opt, batch)] = {"triples_per_s": batch / sec, "s_per_step": sec}
        del model
    return legs


def cpu_lightgcn_leg(small=False, nu=25_000, ni=100_000, d=64, L=3):
    """BASELINE.md section 3 row 4 on a bounded sample: the reference's LightGCN propagation (models/LightGCN.py:174-202,
    restated on the same torch ops: L x torch.sparse.mm + stack + mean; autograd backward) on the adjacency of the FIRST
    `nu` users of the cfg4 recipe over the full 100k-item catalogue, d = 64, L = 3."""
    from oracle import torch_port as TP
    from recsys_pytorch_b200 import synthetic
    if small:
        nu, ni = 5_000, 20_000
    (tp, ti), _ = synthetic.make_interactions_raw(nu, ni, seed=2020, device="cpu")
    t0 = time.perf_counter()
    G = TP.lightgcn_graph(tp.numpy(), ti.numpy(), nu, ni)
    t_graph = time.perf_counter() - t0
    fw, fb = TP.time_lightgcn(G, d, L, reps=2)
    nnzA, N = int(G._nnz()), nu + ni
    bytes_layer = nnzA * (8 + 4 * d) + N * 4 * d                      # same algorithmic model as the GPU leg
    return {"sample": "adjacency of the first %d users x %d items of the cfg4 recipe (nnz %d), d=%d, L=%d; median of 2"
                      % (nu, ni, nnzA, d, L),
            "graph_build_s": t_graph, "propagate_s": fw, "s_per_layer": fw / L, "propagate_fwd_bwd_s": fw + fb,
            "achieved_gbs": bytes_layer * L / fw / 1e9, "ns_per_nnz_layer": fw / L / nnzA * 1e9}


def cpu_cfg5_leg(small=False, num_items=1_000_000, d=128, k=100, n_chunks=4, chunk_users=128):
    """BASELINE.md section 3 row 5 on a bounded sample: the reference's predict() cannot allocate [U, I] at 10M x 1M, so
    its chunked restatement is timed - predict_batch_users (models/MF.py:109-112) + in-chunk -inf mask (:130) + the
    reference's own C++ top-k (func.h:12-31) + holdout metrics (holdout.h:20-103) - on `n_chunks` chunks of
    `chunk_users` users against the full 1M-item catalogue, d = 128, K = 100."""
    from oracle import torch_port as TP
    from recsys_pytorch_b200 import synthetic
    if small:
        num_items, n_chunks = 50_000, 2
    n = n_chunks * chunk_users
    (tp, ti), (vp, vi) = synthetic.make_interactions_raw(n, num_items, seed=77, device="cpu", item_seed=2020)
    tp, ti, vp, vi = (x.numpy() for x in (tp, ti, vp, vi))
    fns, kind = TP.native_eval_lib()
    model = TP.RefMF(n, num_items, d, init_std=0.1)
    TP.eval_chunk(model, np.arange(min(8, n)), tp, ti, vp, vi, k, fns)                   # warm-up (page in, thread pool)
    t0 = time.perf_counter(); pairs = 0
    for ch in range(n_chunks):
        pairs += TP.eval_chunk(model, np.arange(ch * chunk_users, (ch + 1) * chunk_users), tp, ti, vp, vi, k, fns)[0]
    sec = time.perf_counter() - t0
    return {"sample": "%d chunks of %d users x %d items, d=%d, K=%d (chunked predict + mask + top-k + holdout metrics)"
                      % (n_chunks, chunk_users, num_items, d, k),
            "scored_pairs_per_sec": pairs / sec, "seconds": sec, "native": kind,
            "projected_full_sweep_s": 10_000_000 * float(num_items) / (pairs / sec)}


def cpu_baseline_leg(c, args, gpu_triples=None, steps=3, warmup=1):
    """Bounded sample on the host cores: the reference's own step (dense autograd grads + dense Adam over all U+I
    rows, models/MF.py:64-68) at the same table sizes, on the (u, i, j) batches the GPU engine itself sampled
    (`gpu_triples`: {batch_size: [(u,i,j) int64 CPU tensors]})."""
    t_all = time.perf_counter()
    cores = _host_threads()
    legs = _cpu_train_legs(c, gpu_triples, steps, warmup, (c["num_users"], c["num_items"]))
    head = legs["adam_b65536"] if "adam_b65536" in legs else next(iter(legs.values()))
    if not getattr(args, "no_legs", False):
        try:
            legs["lightgcn_cfg4_sample"] = cpu_lightgcn_leg(small=bool(args.small))
        except Exception as e:                                       # pragma: no cover
            legs["lightgcn_cfg4_sample"] = {"error": repr(e)[:300]}
        try:
            legs["eval_cfg5_sample"] = cpu_cfg5_leg(small=bool(args.small))
        except Exception as e:                                       # pragma: no cover
            legs["eval_cfg5_sample"] = {"error": repr(e)[:300]}
    return {"value": head["triples_per_s"], "unit": "triples/s", "cores": cores, "kind": "port", "legs": legs,
            "sample": "%d timed steps per leg of the reference step (MF.py:64-68 restated on torch CPU: dense autograd "
                      "grads + dense optimiser sweep over %d rows) on the very (u,i,j) batches the device sampler "
                      "produced in this run; legs: Adam as-is and the SGD parity swap at B=256 and B=65,536; value = "
                      "Adam at B=65,536; %.1f s total" % (steps, c["num_users"] + c["num_items"],
                                                            time.perf_counter() - t_all)}


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    from oracle import torch_port as TP
    cores = _host_threads()
    c = dict(CFG)
    if args.small:
        c.update(num_users=50_000, num_items=20_000, batch=50_000, eval_users=4096)
    big = 65_536 if not args.small else 8192
    (tr, tg) = _cpu_csr(c, 2 * big, seed=c["seed"])
    n_b = args.steps + args.warmup
    batches = {big: _cpu_batches(c, tr, big, n_b, 1), 256: _cpu_batches(c, tr, 256, n_b, 2)}
    legs = _cpu_train_legs(c, batches, args.steps, args.warmup, (c["num_users"], c["num_items"]))
    head = legs["adam_b%d" % big]
    try:
        legs["lightgcn_cfg4_sample"] = cpu_lightgcn_leg(small=bool(args.small))
    except Exception as e:                                           # pragma: no cover
        legs["lightgcn_cfg4_sample"] = {"error": repr(e)[:300]}
    try:
        legs["eval_cfg5_sample"] = cpu_cfg5_leg(small=bool(args.small))
    except Exception as e:                                           # pragma: no cover
        legs["eval_cfg5_sample"] = {"error": repr(e)[:300]}
    # evaluation leg: the chunked restatement (MF.py:109-112 + in-chunk -inf mask + func.h top-k + holdout.h) on 8
    # chunks of 1024 users
    fns, kind = TP.native_eval_lib()
    model = TP.RefMF(c["num_users"], c["num_items"], c["d"], init_std=c["init_std"])
    n_chunks, cu = 8, 1024
    t0 = time.perf_counter(); pairs = 0
    for ch in range(n_chunks):
        users = np.arange(ch * cu, (ch + 1) * cu)
        n, _ = TP.eval_chunk(model, users, tr[0], tr[1], tg[0], tg[1], c["eval_k"], fns)
        pairs += n
    t_eval = time.perf_counter() - t0
    sample = ("%d timed steps x %d triples at %dx%d d=%d (dense autograd + dense Adam, MF.py:64-68) on batches drawn by "
              "the host mirror of the device sampler; legs: Adam / SGD swap at B=256 and B=%d"
              % (args.steps, big, c["num_users"], c["num_items"], c["d"], big))
    out = {"impl": "reference", "metric": "BPR triples/sec (train)", "value": head["triples_per_s"], "unit": "triples/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["s_per_step"] * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "BPRMF synthetic %dx%d d=%d (BASELINE configs[1]), reference CPU path" %
                      (c["num_users"], c["num_items"], c["d"]), "batch_triples": big, "optimizer": "adam (MF.py:30)"},
           "cpu_baseline": {"value": head["triples_per_s"], "unit": "triples/s", "cores": cores, "kind": "port",
                            "sample": sample, "legs": legs},
           "e2e": {"value": head["triples_per_s"], "unit": "triples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "eval": {"scored_pairs_per_sec": pairs / t_eval, "native": kind, "users": n_chunks * cu, "chunks": n_chunks,
                    "k": c["eval_k"]},
           "gpu_launches": 0}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--gather", default="ldg", choices=["ldg", "async", "tma", "generic"])
    ap.add_argument("--score-algo", dest="score_algo", default="tc", choices=["exact", "tc"])
    ap.add_argument("--layout", default="p2p", choices=["p2p", "item_sharded", "user_sharded"],
                    help="N>1: p2p (default - item table sharded by item-id range as north_star asks AND user table "
                         "sharded by user-id range, user rows through NVSwitch peer memory inside the fused step; the "
                         "fastest measured at BASELINE configs[2]: 5.28 G triples/s on 8 GPUs), user_sharded (users "
                         "sharded, item table replicated, dense item delta all-reduced one step late: 3.91 G at N=8, "
                         "2.67 G at N=2; also reports the p2p layout under 'item_sharded_p2p') or item_sharded (north_star "
                         "exactly as written: replicated user table + NCCL all-reduce of the user-delta buffer: 0.61 G). "
                         "N=8 is BASELINE configs[2] (10M x 1M) in every layout")
    ap.add_argument("--nccl-north-star", dest="nccl_north_star", action="store_true",
                    help="N>1 user_sharded: also measure the north_star layout as written (NCCL all-reduce of user deltas)")
    ap.add_argument("--no-secondary", dest="no_secondary", action="store_true")
    ap.add_argument("--shape", default="cfg3", choices=["cfg3", "cfg2"],
                    help="N>1 item_sharded / user_sharded: cfg3 = 1.25M users + 125k items per GPU (N=8: 10M x 1M), "
                         "cfg2 = N x 1M users over 100k items (round 1's scaling shape)")
    ap.add_argument("--head", type=int, default=None,
                    help="N>1 p2p: number of most-popular items replicated on every rank (default 0 = pure range "
                         "sharding of the whole catalogue; experimental, see p2p.default_head)")
    ap.add_argument("--wire", default="fp32", choices=["fp32", "bf16"],
                    help="N>1 user_sharded: dtype of the all-reduced item-delta buffer")
    ap.add_argument("--exchange", default="auto", choices=["auto", "diff", "buffer"],
                    help="N>1 user_sharded: 'diff' = kernel updates the item replica in place and the difference is "
                         "all-reduced; 'buffer' = kernel accumulates item deltas in a separate dense buffer; 'auto' = "
                         "diff at N=2 (measured 3.44 vs 3.18 G triples/s), buffer at N>=4 (6.20 vs 5.23 G at N=4)")
    ap.add_argument("--head-reduce", dest="head_reduce", default="sum", choices=["mean", "sum"],
                    help="N>1 p2p with a replicated head: how the per-rank differences of the head rows are combined")
    ap.add_argument("--small", action="store_true", help="tiny shapes for a quick functional run (NOT a bench value)")
    ap.add_argument("--no-cpu", dest="no_cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-legs", dest="no_legs", action="store_true",
                    help="skip the secondary workloads (LightGCN cfg4 at N=1, scoring sweep cfg5 at every N)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
