"""Phase timing of the P2P-fused step (route / barrier / step kernel) under torchrun (or single process, world=1)."""
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
m, train, target, hist = p2p.build_rank(c, rank, world, dev, lr=0.05 * B * world, max_batch=B,
                                        head=int(os.environ["HEAD"]) if "HEAD" in os.environ else None)
n_loc = m.uhi - m.ulo
g = torch.Generator(device=dev); g.manual_seed(rank)
perms = [torch.randperm(n_loc, device=dev, generator=g)[:B].to(torch.int32).contiguous() for _ in range(2)]
loss = torch.zeros(1, dtype=torch.float64, device=dev)


def timed(fn, n=10):
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(n): fn(k)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for s in range(4):
    m.step(perms[s % 2], s + 1, B * world, loss_sum=loss)
res = {}
res["full_step"] = timed(lambda k: m.step(perms[k % 2], 10 + k, B * world, loss_sum=loss), 20)
res["route"] = timed(lambda k: m.route(perms[k % 2], 50 + k))
res["barrier"] = timed(lambda k: m.barrier())
def _sync(k):
    m._head_dirty = bool(m.head)
    m.sync_head()
res["sync_head"] = timed(_sync)
# step kernel alone on the (static) outbox content of both parities
m.route(perms[0], 77); m._n += 1; m.route(perms[1], 78); m._n += 1; m.barrier()
m.lr = 0.0
res["compute"] = timed(lambda k: m.compute(B * world, loss_sum=loss))
res["compute_noloss"] = timed(lambda k: m.compute(B * world))
res["n_processed"] = int(m.n_processed.item())
res["route+compute (no barrier)"] = timed(lambda k: (m.route(perms[k % 2], 90 + k), m.compute(B * world, loss_sum=loss)))
print("rank %d %s" % (rank, json.dumps(res)), flush=True)
if world > 1:
    dist.barrier(); m.close(); dist.destroy_process_group()
