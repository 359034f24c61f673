"""Where does the P2P step's time go with many peers?  Times the step kernel alone (static outboxes, lr = 0) under
torchrun for: source order (round-robin / sequential), kernel structure (B200REC_P2P_VARIANT), and with the peer read /
peer write of the user row switched off (measurement flags)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from recsys_pytorch_b200 import p2p
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
c = dict(d=128, batch=1_000_000, seed=2020, lr=0.05, reg=1e-4, init_std=0.01, small=False)
B = c["batch"]
m, train, target, hist = p2p.build_rank(c, rank, world, dev, lr=0.0, max_batch=B, head=0)
n_loc = m.uhi - m.ulo
g = torch.Generator(device=dev); g.manual_seed(rank)
perms = [torch.randperm(n_loc, device=dev, generator=g)[:B].to(torch.int32).contiguous() for _ in range(2)]
m.route(perms[0], 77); m._n += 1; m.route(perms[1], 78); m._n += 1; m.barrier()


LAST = [0.0]


def timed(n=6):
    dist.barrier() if world > 1 else None
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(n):
        m.compute(B * world)
    e1.record(); torch.cuda.synchronize()
    LAST[0] = e0.elapsed_time(e1) / n
    t = torch.tensor([LAST[0]], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def stage_ids_locally():
    """diagnostic: bulk-copy every inbound outbox segment (ids + count) into local memory and point the step at the copies"""
    from recsys_pytorch_b200._lib import lib, check, current_stream
    import ctypes as C
    keep = []
    for b in range(2):
        a = m._step_args[b]
        for s_ in range(world):
            for name in ("in_u", "in_i", "in_j"):
                loc = torch.empty(m.cap, dtype=torch.int32, device=dev)
                check(lib().b200rec_peer_copy(loc.data_ptr(), getattr(a, name)[s_], m.cap * 4, current_stream()))
                getattr(a, name)[s_] = loc.data_ptr(); keep.append(loc)
            loc = torch.empty(1, dtype=torch.int32, device=dev)
            check(lib().b200rec_peer_copy(loc.data_ptr(), a.in_cnt[s_], 4, current_stream()))
            a.in_cnt[s_] = loc.data_ptr(); keep.append(loc)
    torch.cuda.synchronize()
    return keep


RR, NOW, NOR, SEQ = 256, 512, 1024, 2048
for var in ("16", "0"):
    os.environ["B200REC_P2P_VARIANT"] = var
    for name, fl in (("one peer + local interleaved", 0), ("round-robin all", RR), ("rr no-uwrite", RR | NOW),
                     ("rr no-uread no-uwrite", RR | NOW | NOR)):
        if var == "0" and (fl & (NOW | NOR)):
            continue
        m.extra_flags = fl
        timed(2)
        ms = timed()
        print("rank %d variant %-3s %-26s max %.3f ms own %.3f ms" % (rank, var, name, ms, LAST[0]), flush=True)
print("rank %d items in shard %d  processed %d" % (rank, m.ihi - m.ilo, int(m.n_processed.item())), flush=True)
if world > 1:
    dist.barrier(); m.close(); dist.destroy_process_group()
