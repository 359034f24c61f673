"""Per-step GPU timeline of the user-sharded overlapped step (torchrun, N ranks): where do kernel, all-reduce and
apply sit relative to each other?  Diagnostics only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from recsys_pytorch_b200 import engine, synthetic, _lib
from recsys_pytorch_b200.dist import UserShardedBPR, shard_range, allreduce_sum
from recsys_pytorch_b200._lib import SINK_UPDATE, F_USERS_UNIQUE

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
nu, ni, d, B = 1_000_000 * world, 100_000, 128, 1_000_000
ulo, uhi = shard_range(nu, world, rank)
train, _ = synthetic.make_interactions(uhi - ulo, ni, seed=2020 + rank, device=dev)
tr = UserShardedBPR(nu, ni, d, train, rank, world, dev, lr=0.05 * B, reg=1e-4, init_std=0.01, seed=2020)
g = torch.Generator(device=dev); g.manual_seed(rank)
perm = torch.randperm(uhi - ulo, device=dev, generator=g)[:B].to(torch.int32).contiguous()
main = torch.cuda.current_stream(dev); comm = torch.cuda.Stream(device=dev)
bufs = [tr.dV, torch.zeros_like(tr.dV)]
E = lambda: torch.cuda.Event(enable_timing=True)
n_steps = 12
ev = [dict(k0=E(), k1=E(), a0=E(), a1=E(), p0=E(), p1=E()) for _ in range(n_steps)]
done = [torch.cuda.Event(), torch.cuda.Event()]
reserve = int(os.environ.get("RESERVE", "0"))
pending = None
dist.barrier(); torch.cuda.synchronize()
t0 = E(); t0.record(main)
for s in range(n_steps):
    cur = s & 1; buf = bufs[cur]
    buf.zero_()
    ev[s]["k0"].record(main)
    engine.bpr_step(tr.U, tr.V, d, perm, csr=train, lr=tr.lr, reg=tr.reg, sink=SINK_UPDATE, flags=tr.flags | F_USERS_UNIQUE,
                    seed=2020, step=s * world + rank, gV=buf, inv_batch=1.0 / (B * world))
    ev[s]["k1"].record(main)
    with torch.cuda.stream(comm):
        comm.wait_event(ev[s]["k1"])
        ev[s]["a0"].record(comm)
        allreduce_sum(buf)
        ev[s]["a1"].record(comm)
        done[cur].record(comm)
    if os.environ.get("SYNC", "0") == "1":
        main.wait_event(done[cur])            # no overlap: the next kernel waits for this step's all-reduce
    if pending is not None:
        main.wait_event(done[pending])
        ev[s]["p0"].record(main)
        tr.apply_item_delta(bufs[pending])
        ev[s]["p1"].record(main)
    pending = cur
torch.cuda.synchronize()
if rank == 0:
    for s in range(4, n_steps):
        e = ev[s]
        f = lambda x: t0.elapsed_time(x)
        print(f"step {s}: kernel [{f(e['k0']):7.3f} {f(e['k1']):7.3f}] ({e['k0'].elapsed_time(e['k1']):.3f})  "
              f"allreduce [{f(e['a0']):7.3f} {f(e['a1']):7.3f}] ({e['a0'].elapsed_time(e['a1']):.3f})  "
              f"apply(s-1) [{f(e['p0']):7.3f} {f(e['p1']):7.3f}] ({e['p0'].elapsed_time(e['p1']):.3f})")
dist.barrier(); dist.destroy_process_group()
