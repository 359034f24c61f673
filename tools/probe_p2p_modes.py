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


def timed(n=6):
    dist.barrier() if world > 1 else None
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(n):
        m.compute(B * world)
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


SEQ, NOW, NOR = 256, 512, 1024
for var in ("16", "0"):
    os.environ["B200REC_P2P_VARIANT"] = var
    for name, fl in (("round-robin", 0), ("sequential", SEQ), ("rr no-uwrite", NOW), ("rr no-uread no-uwrite", NOW | NOR),
                     ("seq no-uwrite", SEQ | NOW), ("seq no-uread no-uwrite", SEQ | NOW | NOR)):
        if var == "0" and (fl & (NOW | NOR)):
            continue
        m.extra_flags = fl
        timed(2)
        ms = timed()
        if rank == 0:
            print("variant %-3s %-26s %.3f ms" % (var, name, ms), flush=True)
if world > 1:
    dist.barrier(); m.close(); dist.destroy_process_group()
