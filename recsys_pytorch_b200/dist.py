"""Multi-GPU BPR-MF training: one process per GPU, torch.distributed (NCCL over
NVLink 5 / NVSwitch) for the plumbing, the fused sm_100a step kernel for the math.

The reference has no distributed code at all; the correctness oracle
of everything here is the single-device result on the triples the ranks actually
used (tests/test_dist_*.py).

Two layouts:

``item_sharded``  (BASELINE north_star)  the item table is partitioned by contiguous
    item-id range, I/G rows per GPU; the user table is replicated.  Every rank walks
    the same user batch [B]; the rank whose range holds the sampled positive owns the
    triple and draws the negative from its own range.  It updates its item rows in
    place and writes the user delta row into a batch-aligned [B, ld] buffer (zeros
    for triples it does not own) -> ONE all-reduce(sum) of that buffer per step ->
    every rank applies the identical update to its user replica.
    Wire cost: 4*ld bytes per triple per GPU - this, not HBM, bounds the layout
    (SURVEY H9): at d=128 NVLink 5 caps the whole box near 1.0-1.8 G triples/s.

``user_sharded``  (the transpose; SURVEY section 8(e) "route to >= 6x")  users are partitioned
    by id range (U/G rows per GPU, never exchanged); the item table is replicated.
    Each rank trains on its own users; item-row deltas accumulate in a dense
    [I, ld] buffer -> ONE all-reduce(sum) of it per step -> V += dV on every rank.
    Wire cost: 4*ld*I bytes per step, independent of the batch size.
"""
from __future__ import annotations

import json
import os
import time

import numpy as np
import torch
import torch.distributed as dist

from . import engine
from ._lib import F_ITEM_DELTA, F_ITEM_DELTA_BF16, F_TMA_GATHER, F_USERS_UNIQUE, SINK_UPDATE


def shard_range(n: int, world: int, rank: int):
    """Contiguous id range [lo, hi) of `rank` when n ids are split over `world` ranks."""
    return (n * rank) // world, (n * (rank + 1)) // world


def owner_of(ids, n: int, world: int):
    """Rank owning each id under shard_range (vectorised; numpy or torch)."""
    # smallest r with ids < n*(r+1)//world  ==  ((ids+1)*world - 1) // n   for 0 <= ids < n
    return ((ids + 1) * world - 1) // n


def allreduce_sum(t):
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def allgather_rows(shard, total_rows: int, world: int, rank: int):
    """Full [total_rows, ld] table from its contiguous row shards (shard_range), identical on every rank: ONE
    all-gather, used once per evaluation to give every GPU the whole item table (SURVEY section 8(e) "Scoring").
    Shards may differ by one row: they are padded to the largest shard for the collective."""
    ld = shard.shape[1]
    sizes = [shard_range(total_rows, world, r)[1] - shard_range(total_rows, world, r)[0] for r in range(world)]
    assert shard.shape[0] == sizes[rank], "shard does not match shard_range"
    if world == 1 or not dist.is_initialized():
        return shard
    mx = max(sizes)
    send = shard if sizes[rank] == mx else torch.cat([shard, shard.new_zeros((mx - sizes[rank], ld))])
    parts = [shard.new_empty((mx, ld)) for _ in range(world)]
    dist.all_gather(parts, send.contiguous())
    return torch.cat([parts[r][:sizes[r]] for r in range(world)])


def evaluate_user_shard(U, V, d, users, mask, truth, ks, protocol="holdout", row_ids=None, algo=None):
    """Scoring + masked top-K + metrics for THIS rank's users (no collective on the data path: users are independent),
    then one tiny all-reduce of the metric sums.  `users` index rows of U; `mask` / `truth` are DeviceCSRs whose rows
    are addressed by `row_ids` (default: `users`).  Returns ({metric@k: global mean}, n_users_global)."""
    from ._lib import SCORE_TC
    ks = sorted(ks)
    metrics = ("Prec", "Recall", "NDCG") if protocol == "holdout" else ("HR", "NDCG")
    fn = engine.holdout_metrics if protocol == "holdout" else engine.loo_metrics
    sums = torch.zeros(len(metrics) * len(ks) + 1, dtype=torch.float64, device=U.device)
    if users.numel() > 0:
        idx, _ = engine.score_topk(U, V, d, users, mask, max(ks), algo=SCORE_TC if algo is None else algo)
        rows = fn(idx, truth, ks, row_ids=users if row_ids is None else row_ids)
        sums[:-1] = rows.double().sum(0)
        sums[-1] = users.numel()
    allreduce_sum(sums)
    n = int(sums[-1].item())
    means = (sums[:-1] / max(n, 1)).cpu().numpy()
    nk = len(ks)
    return {"%s@%d" % (m, k): float(means[i * nk + j]) for i, m in enumerate(metrics) for j, k in enumerate(ks)}, n


class ItemShardedBPR:
    """north_star layout.  `train` is the GLOBAL DeviceCSR of positives (replicated)."""

    def __init__(self, num_users, num_items, d, train, rank, world, device, lr=0.05, reg=0.0, init_std=0.01,
                 seed=2020, gather="ldg"):
        self.num_users, self.num_items, self.d = num_users, num_items, d
        self.rank, self.world, self.device = rank, world, device
        self.lr, self.reg, self.seed = lr, reg, seed
        self.train = train
        self.lo, self.hi = shard_range(num_items, world, rank)
        g = torch.Generator(device=device); g.manual_seed(seed)
        self.U = engine.alloc_table(num_users, d, device, init_std, g)            # identical replica on every rank
        gi = torch.Generator(device=device); gi.manual_seed(seed * 7919 + 1 + rank)
        self.V = engine.alloc_table(self.hi - self.lo, d, device, init_std, gi)   # rows [lo, hi)
        self.flags = (F_TMA_GATHER if gather == "tma" else 0)
        self._udelta = None

    def local_compute(self, users, step_key, loss_sum=None, out_pos=None, out_neg=None):
        """Fused step on the triples this rank owns: item rows updated in place; returns the
        batch-aligned [B, ld] user-delta buffer (zero rows for triples owned elsewhere)."""
        B = users.numel()
        ld = self.U.shape[1]
        if self._udelta is None or self._udelta.shape[0] != B:
            self._udelta = torch.empty((B, ld), dtype=torch.float32, device=self.device)
        self._udelta.zero_()
        engine.bpr_step(self.U, self.V, self.d, users, csr=self.train, lr=self.lr, reg=self.reg, sink=SINK_UPDATE,
                        flags=self.flags, seed=self.seed, step=step_key, loss_sum=loss_sum, out_pos=out_pos,
                        out_neg=out_neg, item_range=(self.lo, self.hi), udelta=self._udelta)
        return self._udelta

    def apply_user_delta(self, users, udelta):
        engine.rows_add(self.U, users, udelta)

    def step(self, users, step_key, loss_sum=None, out_pos=None, out_neg=None):
        """users: int32 [B] - the SAME tensor content on every rank."""
        ud = self.local_compute(users, step_key, loss_sum, out_pos, out_neg)
        allreduce_sum(ud)                                                          # the ONE collective of the step
        self.apply_user_delta(users, ud)

    def evaluate(self, eval_users, truth, ks, protocol="holdout"):
        """Evaluation: ONE all-gather of the item shards, then every rank scores its contiguous slice of
        `eval_users` (global ids, same tensor on every rank) against the full table; metric sums are all-reduced."""
        V_full = allgather_rows(self.V, self.num_items, self.world, self.rank)
        lo, hi = shard_range(eval_users.numel(), self.world, self.rank)
        return evaluate_user_shard(self.U, V_full, self.d, eval_users[lo:hi].contiguous(), self.train, truth, ks, protocol)


class UserShardedBPR:
    """transpose layout.  `train_local` holds the CSR rows of this rank's users (local row ids)."""

    def __init__(self, num_users, num_items, d, train_local, rank, world, device, lr=0.05, reg=0.0, init_std=0.01,
                 seed=2020, gather="ldg", wire_dtype="fp32"):
        """wire_dtype='bf16': the item-delta buffer (what crosses NVLink and what the kernel's vector atomics hit in
        L2) is bf16 - half the bytes; each atomic add rounds to 8 mantissa bits, so the per-step item update carries
        ~1e-2 relative noise (SGD-level).  'fp32' (default) is exact and is what the equivalence tests use."""
        self.num_users, self.num_items, self.d = num_users, num_items, d
        self.wire_bf16 = str(wire_dtype).lower() == "bf16"
        self.rank, self.world, self.device = rank, world, device
        self.lr, self.reg, self.seed = lr, reg, seed
        self.train = train_local
        self.lo, self.hi = shard_range(num_users, world, rank)
        gu = torch.Generator(device=device); gu.manual_seed(seed * 7919 + 1 + rank)
        self.U = engine.alloc_table(self.hi - self.lo, d, device, init_std, gu)   # rows [lo, hi)
        g = torch.Generator(device=device); g.manual_seed(seed)
        self.V = engine.alloc_table(num_items, d, device, init_std, g)            # identical replica on every rank
        self.dV = torch.zeros_like(self.V, dtype=torch.bfloat16 if self.wire_bf16 else torch.float32)
        self.flags = (F_TMA_GATHER if gather == "tma" else 0) | F_ITEM_DELTA | (F_ITEM_DELTA_BF16 if self.wire_bf16 else 0)

    def local_compute(self, users_local, step_key, global_batch, loss_sum=None, users_unique=True, out_pos=None,
                      out_neg=None):
        """Fused step on this rank's users: user rows updated in place; returns the dense [I, ld]
        item-delta buffer."""
        self.dV.zero_()
        engine.bpr_step(self.U, self.V, self.d, users_local, csr=self.train, lr=self.lr, reg=self.reg,
                        sink=SINK_UPDATE, flags=self.flags | (F_USERS_UNIQUE if users_unique else 0), seed=self.seed,
                        step=step_key * self.world + self.rank, loss_sum=loss_sum, gV=self.dV, out_pos=out_pos,
                        out_neg=out_neg, inv_batch=1.0 / float(global_batch))
        return self.dV

    def apply_item_delta(self, dV):
        if self.wire_bf16:
            engine.add_bf16(self.V, dV)
        else:
            engine.sgd_dense(self.V, dV, -1.0)                                     # V += dV on every replica

    def step(self, users_local, step_key, global_batch, loss_sum=None, users_unique=True, out_pos=None, out_neg=None):
        """users_local: int32 [B_local] local row ids of this rank's batch."""
        dV = self.local_compute(users_local, step_key, global_batch, loss_sum, users_unique, out_pos, out_neg)
        allreduce_sum(dV)                                                          # the ONE collective of the step
        self.apply_item_delta(dV)

    # ---- same step with the collective hidden behind the next step's kernel --------------------------
    def step_overlapped(self, users_local, step_key, global_batch, loss_sum=None, users_unique=True, out_pos=None,
                        out_neg=None):
        """The all-reduce of step s runs on a side stream while the fused kernel of step s+1 executes; the
        reduced item delta is applied one step late (item rows are one step stale - bounded-staleness SGD;
        `flush()` applies the last delta).  User rows are never stale."""
        if not hasattr(self, "_ov"):
            self._ov = dict(bufs=[self.dV, torch.zeros_like(self.dV)], comm=torch.cuda.Stream(device=self.device),
                            ev_done=[torch.cuda.Event(), torch.cuda.Event()], ev_k=torch.cuda.Event(), n=0,
                            pending=None)
        ov = self._ov
        cur = ov["n"] & 1
        buf = ov["bufs"][cur]
        main = torch.cuda.current_stream(self.device)
        buf.zero_()
        engine.bpr_step(self.U, self.V, self.d, users_local, csr=self.train, lr=self.lr, reg=self.reg,
                        sink=SINK_UPDATE, flags=self.flags | (F_USERS_UNIQUE if users_unique else 0), seed=self.seed,
                        step=step_key * self.world + self.rank, loss_sum=loss_sum, gV=buf, out_pos=out_pos,
                        out_neg=out_neg, inv_batch=1.0 / float(global_batch))
        ov["ev_k"].record(main)
        with torch.cuda.stream(ov["comm"]):
            ov["comm"].wait_event(ov["ev_k"])
            allreduce_sum(buf)
            ov["ev_done"][cur].record(ov["comm"])
        if ov["pending"] is not None:                                              # delta of the previous step
            main.wait_event(ov["ev_done"][ov["pending"]])
            self.apply_item_delta(ov["bufs"][ov["pending"]])
        ov["pending"] = cur
        ov["n"] += 1

    # ---- update in place, exchange the difference ------------------------------------------------------
    def step_diff(self, users_local, step_key, global_batch, loss_sum=None, users_unique=True, out_pos=None, out_neg=None):
        """Same overlapped schedule, but the fused kernel updates this rank's item replica IN PLACE (the single-GPU
        fast path: a separate 51 MB delta buffer next to V does not fit L2 together with 1 GB of streaming user rows and
        costs the kernel +25 %), and what crosses NVLink is the difference it made:
            snapshot = V;  kernel(V);  d = V - snapshot;  all_reduce(d) on the side stream;
            one step later  V += sum(d) - d_own          (the other ranks' contribution).
        The arithmetic is the same sum of per-triple deltas, but a replica adds its own deltas triple by triple and the
        others' as one dense sum, so replicas agree to fp32 rounding (not bit for bit); `resync()` makes them identical
        again (broadcast of rank 0's table), `flush()` applies the last pending difference."""
        if not hasattr(self, "_df"):
            z = lambda: torch.zeros_like(self.V)
            self._df = dict(snap=torch.empty_like(self.V), wire=[z(), z()], own=[z(), z()],
                            comm=torch.cuda.Stream(device=self.device), ev_done=[torch.cuda.Event(), torch.cuda.Event()],
                            ev_k=torch.cuda.Event(), n=0, pending=None)
        df = self._df
        cur = df["n"] & 1
        main = torch.cuda.current_stream(self.device)
        df["snap"].copy_(self.V)
        engine.bpr_step(self.U, self.V, self.d, users_local, csr=self.train, lr=self.lr, reg=self.reg,
                        sink=SINK_UPDATE, flags=(self.flags & ~(F_ITEM_DELTA | F_ITEM_DELTA_BF16)) |
                        (F_USERS_UNIQUE if users_unique else 0), seed=self.seed,
                        step=step_key * self.world + self.rank, loss_sum=loss_sum, out_pos=out_pos, out_neg=out_neg,
                        inv_batch=1.0 / float(global_batch))
        engine.delta_diff(self.V, df["snap"], df["wire"][cur], df["own"][cur])
        df["ev_k"].record(main)
        with torch.cuda.stream(df["comm"]):
            df["comm"].wait_event(df["ev_k"])
            allreduce_sum(df["wire"][cur])
            df["ev_done"][cur].record(df["comm"])
        if df["pending"] is not None:
            main.wait_event(df["ev_done"][df["pending"]])
            engine.delta_apply(self.V, df["wire"][df["pending"]], df["own"][df["pending"]])
        df["pending"] = cur
        df["n"] += 1

    def resync(self):
        """Make the item replicas bit-identical again (step_diff lets them differ by fp32 rounding)."""
        self.flush()
        if dist.is_initialized() and self.world > 1:
            dist.broadcast(self.V, src=0)

    def evaluate(self, eval_users_local, truth_local, ks, protocol="holdout"):
        """Evaluation: the item table is already replicated, every rank scores its own users (local row
        ids of this rank's shard; `truth_local` rows are local too); metric sums are all-reduced."""
        self.flush()
        return evaluate_user_shard(self.U, self.V, self.d, eval_users_local, self.train, truth_local, ks, protocol)

    def flush(self):
        ov = getattr(self, "_ov", None)
        if ov and ov["pending"] is not None:
            torch.cuda.current_stream(self.device).wait_event(ov["ev_done"][ov["pending"]])
            self.apply_item_delta(ov["bufs"][ov["pending"]])
            ov["pending"] = None
        df = getattr(self, "_df", None)
        if df and df["pending"] is not None:
            torch.cuda.current_stream(self.device).wait_event(df["ev_done"][df["pending"]])
            engine.delta_apply(self.V, df["wire"][df["pending"]], df["own"][df["pending"]])
            df["pending"] = None


# ------------------------------------------------------------------------------------------------
# bench.py --gpus N (N > 1)
# ------------------------------------------------------------------------------------------------
def bench_multi_gpu(args, c, rank, world, dev, timed_region, timed_under_load, hbm_gbs, peak_src):
    """The two NCCL-collective layouts (kept beside the default P2P layout of p2p.py for comparison).  Weak scaling:
    per-GPU work is fixed (B_local = c['batch'] triples per rank per step).  --shape cfg3 (default): every GPU brings
    1.25M users and 125k items (N = 8 is BASELINE configs[2], 10M x 1M); --shape cfg2: round 1's shape, N x 1M users
    over a fixed 100k-item catalogue.  The per-triple step lr/B_glob is the same for every N."""
    from . import _lib, synthetic
    from .p2p import PER_GPU
    d = c["d"]
    B_local = c["batch"]
    if getattr(args, "shape", "cfg3") == "cfg3" and not c.get("small"):
        nu, ni = PER_GPU["users"] * world, PER_GPU["items"] * world
    else:
        nu, ni = c["num_users"] * world, c["num_items"]
    c = dict(c)
    c["lr"] = c["lr_per_triple"] * B_local * world          # r1 kept lr fixed while 1/B_glob shrank: user step ~ 1/N
    layout = args.layout
    loss = torch.zeros(1, dtype=torch.float64, device=dev)
    secondary = None
    if layout == "user_sharded" and getattr(args, "nccl_north_star", False):
        # the north_star layout exactly as written (replicated user table, NCCL all-reduce of the [B, ld] user deltas),
        # measured in the same run on request (it needs the whole CSR and user table on every rank)
        secondary = _bench_item_sharded(args, c, rank, world, dev, timed_region, nu, ni, max(3, args.steps // 10))
    if layout == "item_sharded":
        # replicated CSR + replicated user table; every rank walks the same global batch of N*B_local users
        train, _ = synthetic.make_interactions(nu, ni, seed=c["seed"], device=dev)
        tr = ItemShardedBPR(nu, ni, d, train, rank, world, dev, lr=c["lr"], reg=c["reg"], init_std=c["init_std"],
                            seed=c["seed"], gather=args.gather)
        g = torch.Generator(device=dev); g.manual_seed(c["seed"])
        B_glob = B_local * world
        perms = [torch.randperm(nu, device=dev, generator=g)[:B_glob].to(torch.int32).contiguous() for _ in range(2)]

        def step(s):
            tr.step(perms[s % 2], s + 1, loss_sum=loss)
        wire = 4 * tr.U.shape[1]
        coll = "all_reduce(sum) of the [B, ld] fp32 user-delta buffer (%d MiB) per step" % (B_glob * wire >> 20)
    else:
        ulo, uhi = shard_range(nu, world, rank)
        train, target = synthetic.make_interactions(uhi - ulo, ni, seed=c["seed"] + 1 + rank, device=dev,
                                                    item_seed=c["seed"])
        tr = UserShardedBPR(nu, ni, d, train, rank, world, dev, lr=c["lr"], reg=c["reg"], init_std=c["init_std"],
                            seed=c["seed"], gather=args.gather, wire_dtype=getattr(args, "wire", "fp32"))
        g = torch.Generator(device=dev); g.manual_seed(c["seed"] + rank)
        perms = [torch.randperm(uhi - ulo, device=dev, generator=g)[:B_local].to(torch.int32).contiguous() for _ in range(2)]
        B_glob = B_local * world

        sync_mode = os.environ.get("B200REC_DIST_SYNC", "0") == "1"
        exchange = getattr(args, "exchange", "auto")
        if exchange == "auto":      # measured: diff wins at N=2, the delta buffer at N=4 (profiles/r01_run30/31)
            exchange = "diff" if world <= 2 else "buffer"
        if tr.wire_bf16 and exchange == "diff":
            exchange = "buffer"

        def step(s):
            if sync_mode:
                tr.step(perms[s % 2], s + 1, B_glob, loss_sum=loss)
            elif exchange == "diff":
                tr.step_diff(perms[s % 2], s + 1, B_glob, loss_sum=loss)
            else:
                tr.step_overlapped(perms[s % 2], s + 1, B_glob, loss_sum=loss)
        if exchange == "diff" and not sync_mode:
            coll = ("item replica updated in place by the fused kernel; all_reduce(sum) of the dense [I, ld] fp32 "
                    "difference V - snapshot (%d MiB) per step on a side stream, overlapped with the next step's kernel; "
                    "V += sum - own one step later (item rows one step stale, replicas equal to fp32 rounding)"
                    % (tr.V.numel() * 4 >> 20))
        else:
            coll = ("all_reduce(sum) of the dense [I, ld] %s item-delta buffer (%d MiB) per step, on a side stream, "
                    "overlapped with the next step's kernel (item rows one step stale)"
                    % ("bf16" if tr.wire_bf16 else "fp32", tr.dV.numel() * tr.dV.element_size() >> 20))

    for s in range(args.warmup):
        step(s)
    counted = [0, 0]

    def step_counted(s):
        if s == 0:
            counted[0] = _lib.launch_count()
        step(s)
        if s == args.steps - 1:
            counted[1] = _lib.launch_count()

    def step_idle(s):      # same step (kernel + collective) with learning rate 0: clock-sampling window only
        lr = tr.lr
        tr.lr = 0.0
        try:
            step(100000 + s)
        finally:
            tr.lr = lr
    ms, clk = timed_under_load(step_counted, step_idle, args.steps, world, dev.index or 0, rank=rank, pre_steps=500,
                               post_steps=150)
    launches = counted[1] - counted[0]
    reps = [ms] + [timed_region(lambda s: step(1000 * (r + 1) + s), args.steps, world) for r in range(2)]
    ms = float(np.median(reps))                               # median of 3 timed repeats of the K steps
    # e2e: same step with the batch's user ids arriving from pinned host memory and the loss read back
    host = [p.cpu().pin_memory() for p in perms]

    def step_e2e(s):
        u = host[s % 2].to(dev, non_blocking=True)
        loss.zero_()
        if layout == "item_sharded":
            tr.step(u, 1000 + s, loss_sum=loss)
        elif exchange == "diff" and not sync_mode:
            tr.step_diff(u, 1000 + s, B_glob, loss_sum=loss)
        else:
            tr.step_overlapped(u, 1000 + s, B_glob, loss_sum=loss)
        return float(loss.item())

    for s in range(2):
        step_e2e(s)
    ms_e2e = timed_region(step_e2e, args.steps, world)
    # evaluation leg: users are independent - every rank scores the same number of its OWN users against
    # the replicated item table (weak scaling), no collective on the data path, one all-reduce of the metric sums
    eval_leg = None
    if layout == "user_sharded":
        n_ev = min(int(c["eval_users"]), uhi - ulo)
        ev_users = torch.arange(n_ev, dtype=torch.int32, device=dev)
        tr.evaluate(ev_users, target, [c["eval_k"]])                                  # warm-up
        res = {}

        def ev_step(s):
            res["scores"], res["n"] = tr.evaluate(ev_users, target, [c["eval_k"]])
        ms_ev = timed_region(ev_step, 3, world) / 3
        eval_leg = {"scored_pairs_per_sec": float(res["n"]) * ni / (ms_ev * 1e-3), "ms": ms_ev,
                    "ndcg@%d" % c["eval_k"]: res["scores"]["NDCG@%d" % c["eval_k"]], "users": res["n"],
                    "k": c["eval_k"], "algo": "tc", "flops_per_pair": 2 * d,
                    "note": "whole evaluate() per rank incl. metrics and the metric all-reduce; users sharded, item "
                            "table replicated"}
    # the item-sharded P2P layout and the scoring sweep, measured in the same run (failures are reported, not fatal)
    p2p_res, cfg5 = None, None
    if layout == "user_sharded" and not getattr(args, "no_secondary", False):
        del tr, train
        torch.cuda.empty_cache()
        try:
            from . import p2p as _p2p
            p2p_res = _p2p.measure_p2p(c, rank, world, dev, timed_region, steps=max(5, args.steps // 2))
        except Exception as e:                                       # pragma: no cover
            p2p_res = {"error": repr(e)[:300]} if rank == 0 else None
    if not getattr(args, "no_legs", False):
        try:
            from . import bench_legs
            torch.cuda.empty_cache()
            cfg5 = bench_legs.cfg5_leg(dev, rank, world, float(getattr(args, "bf16_tf", 1590.0)), peak_src,
                                       small=bool(c.get("small")))
        except Exception as e:                                       # pragma: no cover
            cfg5 = {"error": repr(e)[:300]}
    if rank == 0:
        val = B_glob * args.steps / (ms * 1e-3)
        bpt = 24 * d + 8
        per_gpu = bpt * B_local / (ms / args.steps * 1e-3) / 1e9
        out = {"metric": "BPR triples/sec (train)", "value": val, "unit": "triples/s", "n_gpus": world,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": {"workload": "BPRMF synthetic %dx%d d=%d, %s over %d GPUs" % (nu, ni, d, layout, world),
                          "batch_triples": B_glob, "per_gpu_triples": B_local, "optimizer": "sgd+l2",
                          "parallelism": "%s%d" % (layout, world), "collective": coll, "gather": args.gather,
                          "exchange": (exchange if layout == "user_sharded" else None), "lr_per_triple": c["lr_per_triple"],
                          "l2_policy": "inputs larger than L2", "timing": "median of 3 repeats of the K steps",
                          "repeats_ms": reps},
               "clocks": clk,
               "e2e": {"value": B_glob * args.steps / (ms_e2e * 1e-3), "unit": "triples/s",
                       "h2d_bytes_per_step": int(perms[0].numel() * 4), "d2h_bytes_per_step": 8,
                       "ms_per_step": ms_e2e / args.steps},
               "gpu_launches": int(launches),
               "roofline": {"bound": "hbm", "achieved": per_gpu, "peak": hbm_gbs, "unit": "GB/s",
                            "frac": per_gpu / hbm_gbs, "traffic": None, "peak_source": peak_src,
                            "note": "per-GPU algorithmic bytes of the step kernel over the WHOLE step time "
                                    "(collective included)"},
               "cpu_baseline": None}
        if eval_leg is not None:
            out["eval"] = eval_leg
        if secondary is not None:
            out["north_star_item_sharded_nccl"] = secondary
        if p2p_res is not None:
            out["item_sharded_p2p"] = p2p_res
        if cfg5 is not None:
            out["eval_cfg5"] = cfg5
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


def _bench_item_sharded(args, c, rank, world, dev, timed_region, nu, ni, steps):
    """BASELINE north_star layout (item table sharded by id range, user table replicated, ONE all-reduce of the
    [B, ld] user-delta buffer per step) at the same shape, for the record."""
    from . import synthetic
    train, _ = synthetic.make_interactions(nu, ni, seed=c["seed"], device=dev)
    tr = ItemShardedBPR(nu, ni, c["d"], train, rank, world, dev, lr=c["lr"], reg=c["reg"], init_std=c["init_std"],
                        seed=c["seed"])
    g = torch.Generator(device=dev); g.manual_seed(c["seed"])
    B_glob = c["batch"] * world
    perms = [torch.randperm(nu, device=dev, generator=g)[:B_glob].to(torch.int32).contiguous() for _ in range(2)]
    for s in range(2):
        tr.step(perms[s % 2], s + 1)
    ms = timed_region(lambda s: tr.step(perms[s % 2], s + 3), steps, world)
    res = {"value": B_glob * steps / (ms * 1e-3), "unit": "triples/s", "steps": steps, "ms_per_step": ms / steps,
           "batch_triples": B_glob,
           "collective": "all_reduce(sum) of the [B, ld] fp32 user-delta buffer (%d MiB) per step" % (B_glob * 4 * tr.U.shape[1] >> 20)}
    del tr, train, perms
    torch.cuda.empty_cache()
    return res
