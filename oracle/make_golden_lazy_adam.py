"""tests/golden/tiny_lazy_adam.npz: the reference's MF (models/MF.py) with its embeddings switched to sparse
gradients and its optimiser swapped for torch.optim.SparseAdam(lr=1e-3) - the row-wise Adam of SURVEY section 8(f)
rank 1 - on the batches of tiny_bpr.npz.  TEST INFRASTRUCTURE ONLY.   python -m oracle.make_golden_lazy_adam"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness  # noqa: E402


def main():
    import torch
    ref = ref_harness.load()
    g = np.load(os.path.join(ROOT, "tests", "golden", "tiny_bpr.npz"))
    ds = types.SimpleNamespace(num_users=50, num_items=40)
    m = ref.MF(ds, {"hidden_dim": 8, "pointwise": False, "loss_func": "ce"}, torch.device("cpu"))
    with torch.no_grad():
        m.user_embedding.weight.copy_(torch.from_numpy(g["U0"])); m.item_embedding.weight.copy_(torch.from_numpy(g["V0"]))
    m.user_embedding.sparse = True; m.item_embedding.sparse = True      # nn.Embedding(sparse=True) gradients
    m.optimizer = torch.optim.SparseAdam(list(m.parameters()), lr=1e-3)
    losses, Us, Vs = [], [], []
    for rep in range(2):                       # 6 steps: the 3 batches twice (rows re-touched after a gap)
        for b in range(3):
            u, i, j = (torch.from_numpy(g[k][b]) for k in ("users", "pos", "neg"))
            m.optimizer.zero_grad()
            ls = m.process_one_batch(u, i, j); ls.backward(); m.optimizer.step()
            losses.append(ls.item())
            Us.append(m.user_embedding.weight.detach().numpy().copy()); Vs.append(m.item_embedding.weight.detach().numpy().copy())
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "tiny_lazy_adam.npz"), U=np.stack(Us), V=np.stack(Vs),
                        loss=np.array(losses, np.float32))
    print("tiny_lazy_adam ok", losses)


if __name__ == "__main__":
    main()
