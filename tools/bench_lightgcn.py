"""cfg4 measurement (BASELINE configs[3]): LightGCN synthetic 1M x 100k, 3 layers, d=64 - sparse propagation
(csrc/spmm.cu) + the shared BPR step / scoring kernels.  Prints one JSON line; not the headline bench (bench.py)."""
import json, os, sys, types, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recsys_pytorch_b200 import engine, synthetic, _lib
from recsys_pytorch_b200.lightgcn import LightGCN

dev = torch.device("cuda")
nu, ni, d, L = int(os.environ.get("NU", 1_000_000)), int(os.environ.get("NI", 100_000)), int(os.environ.get("D", 64)), 3
B = int(os.environ.get("B", 1_000_000))
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
train, target = synthetic.make_interactions(nu, ni, seed=2020, device=dev)
ds = types.SimpleNamespace(num_users=nu, num_items=ni, train_data=train, valid_input=train, valid_target=target,
                           protocol="holdout", dataname="synthetic")
m = LightGCN(ds, {"emb_dim": d, "num_layers": L, "optimizer": "adam", "lr": 1e-3, "score_algo": "tc"}, dev)
t0 = time.perf_counter(); m.Graph = m.getSparseGraph(train); torch.cuda.synchronize(); t_graph = time.perf_counter() - t0
nnzA = int(m.Graph[1].numel()); N = nu + ni


def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


ms_prop = timeit(lambda: m.propagate(m.E0, m.out))
g = torch.Generator(device=dev); g.manual_seed(0)
users = torch.randperm(nu, device=dev, generator=g)[:B].to(torch.int32).contiguous()
k = [0]
def step():
    k[0] += 1
    m.train_batch(users, csr=train, step_key=k[0], users_unique=True)
ms_step = timeit(step, n=5, warm=2)
m.update_lightgcn_embedding()
ev_users = torch.arange(37888, dtype=torch.int32, device=dev)
ms_eval = timeit(lambda: m.predict_topk_device(ev_users, train, 10), n=5, warm=2)
bytes_layer = nnzA * (8 + 4 * d) + N * 4 * d            # no-reuse gather model
compulsory = nnzA * 8 + 2 * N * 4 * d
out = {"workload": "LightGCN synthetic %dx%d L=%d d=%d (BASELINE configs[3])" % (nu, ni, L, d), "nnz_adj": nnzA,
       "graph_build_s": t_graph, "propagate_ms": ms_prop, "ms_per_layer": ms_prop / L,
       "algorithmic_gb_per_layer": bytes_layer / 1e9, "achieved_gbs": bytes_layer * L / (ms_prop * 1e-3) / 1e9,
       "compulsory_gbs": compulsory * L / (ms_prop * 1e-3) / 1e9, "hbm_peak_gbs": peaks["hbm_gbs"],
       "frac_of_hbm_peak": bytes_layer * L / (ms_prop * 1e-3) / 1e9 / peaks["hbm_gbs"],
       "train_step_ms": ms_step, "train_triples_per_s": B / (ms_step * 1e-3), "batch_triples": B,
       "step_anatomy": "propagate fwd (3 SpMM) + fused BPR step (SINK_GRAD) + propagate bwd (3 SpMM) + dense Adam",
       "eval_ms_37888_users_k10": ms_eval, "eval_pairs_per_s": 37888 * ni / (ms_eval * 1e-3)}
print(json.dumps(out))
