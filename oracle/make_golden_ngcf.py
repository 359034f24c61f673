"""tests/golden/ngcf_ml100k.npz: the reference's NGCF (models/NGCF.py) on its own ml-100k split with node_dropout =
mess_dropout = 0 (deterministic), L=2, d=16: parameters, propagated tables (`_ngcf_embedding` :182-221), one BPR batch
(`process_one_batch` :126-132): loss, autograd gradients of every parameter, and the parameters after one Adam step
(:46).  TEST INFRASTRUCTURE ONLY (needs /root/reference).      python -m oracle.make_golden_ngcf"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness  # noqa: E402


def main():
    import torch
    ref = ref_harness.load()
    from models.NGCF import NGCF
    ref.set_random_seed(2020)
    ds = ref_harness.ml100k_dataset(ref)
    hp = {"emb_dim": 16, "num_layers": 2, "node_dropout": 0.0, "mess_dropout": 0.0, "split": False, "num_folds": 100,
          "graph_dir": os.path.join(ref_harness.WORK, "graph_ngcf"), "reg": 1e-4}
    m = NGCF(ds, hp, torch.device("cpu"))
    tr = ds.train_data.tocsr(); tr.sort_indices()
    m.Graph = m.getSparseGraph(tr)
    out = {"U0": m.user_embedding.weight.detach().numpy().copy(), "V0": m.item_embedding.weight.detach().numpy().copy()}
    for k in range(2):
        for nm in ("W_gc", "b_gc", "W_bi", "b_bi"):
            out["%s_%d" % (nm, k)] = m.weight_dict["%s_%d" % (nm, k)].detach().numpy().copy()
    m.train()
    m.update_ngcf_embedding()
    out["prop_U"] = m.user_embeddings.detach().numpy().copy()
    out["prop_V"] = m.item_embeddings.detach().numpy().copy()
    rng = np.random.default_rng(5)
    u = rng.integers(0, ds.num_users, 256); i = rng.integers(0, ds.num_items, 256); j = rng.integers(0, ds.num_items, 256)
    out.update(users=u.astype(np.int32), pos=i.astype(np.int32), neg=j.astype(np.int32))
    m.optimizer.zero_grad()
    ls = m.process_one_batch(*(torch.from_numpy(a) for a in (u, i, j))); ls.backward()
    out["loss"] = np.float32(ls.item())
    out["dU0"] = m.user_embedding.weight.grad.numpy().copy(); out["dV0"] = m.item_embedding.weight.grad.numpy().copy()
    for k in range(2):
        for nm in ("W_gc", "b_gc", "W_bi", "b_bi"):
            out["d%s_%d" % (nm, k)] = m.weight_dict["%s_%d" % (nm, k)].grad.numpy().copy()
    m.optimizer.step()                                               # Adam(lr=1e-3), NGCF.py:46
    out["adam_U"] = m.user_embedding.weight.detach().numpy().copy(); out["adam_V"] = m.item_embedding.weight.detach().numpy().copy()
    for k in range(2):
        for nm in ("W_gc", "b_gc", "W_bi", "b_bi"):
            out["adam_%s_%d" % (nm, k)] = m.weight_dict["%s_%d" % (nm, k)].detach().numpy().copy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ngcf_ml100k.npz"), **out)
    print("ngcf ok, loss", out["loss"])


if __name__ == "__main__":
    main()
