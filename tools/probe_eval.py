import os, sys, types, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from recsys_pytorch_b200 import engine, synthetic, _lib
from recsys_pytorch_b200.mf import MF
from recsys_pytorch_b200.evaluation import Evaluator
import bench
dev = torch.device("cuda")
c = dict(bench.CFG)
train, target = synthetic.make_interactions(c["num_users"], c["num_items"], seed=c["seed"], device=dev)
ds = types.SimpleNamespace(num_users=c["num_users"], num_items=c["num_items"], train_data=train, valid_input=train,
                           valid_target=target, protocol="holdout", dataname="synthetic")
hp = {"hidden_dim": c["d"], "pointwise": False, "loss_func": "ce", "optimizer": "sgd", "lr": c["lr"], "reg": c["reg"],
      "init_std": c["init_std"], "gather": "ldg", "seed": c["seed"], "score_algo": "tc"}
model = MF(ds, hp, dev)
ev = Evaluator(train, bench._SubsetTarget(target, c["eval_users"]), protocol="holdout", ks=[c["eval_k"]])
ev.evaluate(model); torch.cuda.synchronize()
for _ in range(3):
    t0 = time.perf_counter(); ev.evaluate(model); torch.cuda.synchronize(); print("evaluate: %.2f ms" % ((time.perf_counter() - t0) * 1e3))
pr = cProfile.Profile(); pr.enable(); ev.evaluate(model); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)

def T(label, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("%-28s host %.3f ms, +sync %.3f ms" % (label, (t1 - t0) * 1e3, (t2 - t1) * 1e3)); return r
users = ev._dev["users"]; truth = ev._truth_device(dev)
for rep in range(2):
    idx = T("predict_topk_device", lambda: model.predict_topk_device(users, train, 10))[0]
    rows = T("holdout_metrics", lambda: engine.holdout_metrics(idx, truth, [10], row_ids=users))
    T("column_means", lambda: engine.column_means(rows))
    T("keys->array", lambda: np.array(list(ev.eval_target.keys())))
    T("empty", lambda: None)
    T("evaluate", lambda: ev.evaluate(model))
