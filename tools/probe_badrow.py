import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recsys_pytorch_b200 import engine, synthetic, _lib
from recsys_pytorch_b200.mf import MF
dev = torch.device("cuda")
nu, ni, d = 1_000_000, 100_000, 128
train, target = synthetic.make_interactions(nu, ni, seed=2020, device=dev)
ds = types.SimpleNamespace(num_users=nu, num_items=ni, train_data=train, valid_input=train, valid_target=target,
                           protocol="holdout", dataname="synthetic")
hp = {"hidden_dim": d, "pointwise": False, "loss_func": "ce", "optimizer": "sgd", "lr": 0.05 * 1_000_000, "reg": 1e-4,
      "init_std": 0.01, "gather": "ldg", "seed": 2020, "score_algo": "tc"}
model = MF(ds, hp, dev)
g = torch.Generator(device=dev); g.manual_seed(2020)
perm = torch.randperm(nu, device=dev, generator=g)[:1_000_000].to(torch.int32).contiguous()
for s in range(23):
    engine.bpr_step(model.U, model.V, d, perm, csr=train, lr=hp["lr"], reg=hp["reg"], sink=_lib.SINK_UPDATE,
                    flags=_lib.GATHER_FLAGS["ldg"] | _lib.F_USERS_UNIQUE, seed=2020, step=s + 1)
torch.cuda.synchronize()
U, V = model.U[:, :d], model.V[:, :d]
vn = V.norm(dim=1)
order = torch.argsort(vn, descending=True)
print("item norm quantiles:", torch.quantile(vn, torch.tensor([0., .01, .1, .5, .9, .99, 1.], device=dev)).tolist())
un = U[:37888].norm(dim=1)
print("user norm quantiles:", torch.quantile(un, torch.tensor([0., .01, .1, .5, .9, .99, 1.], device=dev)).tolist())
for row in (506, 793, 0, 1):
    s = (V[order] @ U[row]).double().cpu()
    n_sorted = vn[order].double().cpu()
    c = 2.0 ** -10 * 1.05 + d * 2.4e-7
    e = c * float(un[row]) * n_sorted
    # running 10th best of L = s - e
    import heapq
    heap = []; hits = 0; hit_pos = []
    for i in range(ni):
        L = float(s[i] - e[i]); H = float(s[i] + e[i])
        tau = heap[0] if len(heap) >= 10 else -1e30
        if H >= tau:
            hits += 1; hit_pos.append(i)
        if len(heap) < 10: heapq.heappush(heap, L)
        elif L > heap[0]: heapq.heapreplace(heap, L)
    hp_ = torch.tensor(hit_pos)
    print(f"row {row}: |u|={float(un[row]):.4f} deg={int(train.indptr[row+1]-train.indptr[row])} ideal-streaming hits={hits} "
          f"hit position quantiles={[int(x) for x in torch.quantile(hp_.double(), torch.tensor([0.,.25,.5,.75,1.]).double()).tolist()]} "
          f"score min/median/max={float(s.min()):.4g}/{float(s.median()):.4g}/{float(s.max()):.4g} final tau={heap[0]:.4g} "
          f"e first/mid/last={float(e[0]):.3g}/{float(e[ni//2]):.3g}/{float(e[-1]):.3g}")
    # correlation between score and norm
    print("   corr(score, norm) =", float(torch.corrcoef(torch.stack([s, n_sorted]))[0, 1]), " top10 positions:", sorted(torch.topk(s, 10).indices.tolist()))
