"""tests/golden/tiny_pointwise.npz: the reference's MF (models/MF.py) in its POINTWISE mode (hparams['pointwise']=True,
MF.py:49-52,101-102) on a tiny case with duplicate ids - forward, loss ('ce' =
binary_cross_entropy_with_logits and 'mse'), autograd gradients, three dense-Adam steps (MF.py:30) - and the batches
the reference's PointwiseGenerator (data/generators.py:43-136) emits after np.random.seed.  TEST INFRASTRUCTURE ONLY.
    python -m oracle.make_golden_pointwise"""
import os
import sys
import types

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness  # noqa: E402


def main():
    import torch
    ref = ref_harness.load()
    from data.generators import PointwiseGenerator
    g = np.load(os.path.join(ROOT, "tests", "golden", "tiny_bpr.npz"))
    U0, V0 = g["U0"], g["V0"]
    nu, ni = U0.shape[0], V0.shape[0]
    rng = np.random.default_rng(12)
    out = {}
    users = np.stack([np.r_[rng.integers(0, nu, 12), [3, 3, 7, 7]] for _ in range(3)]).astype(np.int64)   # duplicates
    items = np.stack([np.r_[rng.integers(0, ni, 12), [5, 5, 5, 9]] for _ in range(3)]).astype(np.int64)
    ratings = (rng.random((3, 16)) < 0.5).astype(np.float32)
    out.update(users=users, items=items, ratings=ratings)
    ds = types.SimpleNamespace(num_users=nu, num_items=ni)
    for lf in ("ce", "mse"):
        m = ref.MF(ds, {"hidden_dim": U0.shape[1], "pointwise": True, "loss_func": lf}, torch.device("cpu"))
        with torch.no_grad():
            m.user_embedding.weight.copy_(torch.from_numpy(U0 * (4.0 if lf == "ce" else 1.0)))   # 'ce': saturating logits too
            m.item_embedding.weight.copy_(torch.from_numpy(V0 * (4.0 if lf == "ce" else 1.0)))
        u, i, r = torch.from_numpy(users[0]), torch.from_numpy(items[0]), torch.from_numpy(ratings[0])
        m.optimizer.zero_grad()
        loss = m.process_one_batch(u, i, r); loss.backward()
        out[f"{lf}_scale"] = np.float32(4.0 if lf == "ce" else 1.0)
        out[f"{lf}_scores"] = m.forward(u, i).detach().numpy()
        out[f"{lf}_loss"] = np.float32(loss.item())
        out[f"{lf}_dU"] = m.user_embedding.weight.grad.numpy().copy()
        out[f"{lf}_dV"] = m.item_embedding.weight.grad.numpy().copy()
        losses = []
        for b in range(3):                                                        # Adam as-is (MF.py:30), from the same start
            u, i, r = torch.from_numpy(users[b]), torch.from_numpy(items[b]), torch.from_numpy(ratings[b])
            if b > 0:
                m.optimizer.zero_grad()
                ls = m.process_one_batch(u, i, r); ls.backward()
            else:
                ls = loss
            m.optimizer.step()
            losses.append(ls.item())
        out[f"{lf}_adam_U"] = m.user_embedding.weight.detach().numpy().copy()
        out[f"{lf}_adam_V"] = m.item_embedding.weight.detach().numpy().copy()
        out[f"{lf}_adam_loss"] = np.array(losses, np.float32)
    # the reference generator's batches (2 epochs) on a small interaction matrix
    R = sp.random(30, 25, density=0.15, random_state=4, format="csr", dtype=np.float64); R.data[:] = 1.0
    out.update(R_indptr=R.indptr.astype(np.int64), R_indices=R.indices.astype(np.int32), R_shape=np.array(R.shape))
    np.random.seed(99)
    gen = PointwiseGenerator(R, return_rating=True, num_negatives=1, batch_size=32, shuffle=True, device=torch.device("cpu"))
    bu, bi, br, lens = [], [], [], []
    for _ in range(2):
        for (a, b, c) in gen:
            bu.append(a.numpy()); bi.append(b.numpy()); br.append(c.numpy()); lens.append(len(a))
    out.update(gen_users=np.concatenate(bu), gen_items=np.concatenate(bi), gen_ratings=np.concatenate(br),
               gen_lens=np.array(lens, np.int32), gen_num_batches=np.int32(len(gen)))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "tiny_pointwise.npz"), **out)
    print("tiny_pointwise ok", {k: float(out[k]) for k in ("ce_loss", "mse_loss")}, lens[:4])


if __name__ == "__main__":
    main()
