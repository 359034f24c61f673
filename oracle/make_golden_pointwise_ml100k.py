"""tests/golden/ml100k_pointwise.npz: the reference's MF in POINTWISE mode (models/MF.py:49-52,63-68,101-102) on its own
ml-100k split, driven by its own PointwiseGenerator (data/generators.py:43-136) after set_random_seed(2020): the first
STEPS batches exactly as the reference emits them, the per-batch losses and the tables after those dense-Adam steps.
TEST INFRASTRUCTURE ONLY (needs /root/reference).      python -m oracle.make_golden_pointwise_ml100k"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness  # noqa: E402

STEPS, D = 6, 32


def main():
    import torch
    ref = ref_harness.load()
    from data.generators import PointwiseGenerator
    out = {}
    for lf in ("ce", "mse"):
        ref.set_random_seed(2020)
        ds = ref_harness.ml100k_dataset(ref)
        m = ref.MF(ds, {"hidden_dim": D, "pointwise": True, "loss_func": lf}, torch.device("cpu"))
        if lf == "mse":                                   # N(0,1) tables give |x| ~ 6: keep the squared error in fp32 range
            with torch.no_grad():
                m.user_embedding.weight.mul_(0.3); m.item_embedding.weight.mul_(0.3)
        out[f"{lf}_U0"] = m.user_embedding.weight.detach().numpy().copy()
        out[f"{lf}_V0"] = m.item_embedding.weight.detach().numpy().copy()
        gen = PointwiseGenerator(ds.train_data, return_rating=True, num_negatives=1, batch_size=256, shuffle=True,
                                 device=torch.device("cpu"))                      # MF.py:49-52
        bu, bi, br, lens, losses = [], [], [], [], []
        for b, (u, i, r) in enumerate(gen):                                       # MF.py:63-68
            if b == STEPS:
                break
            m.optimizer.zero_grad()
            loss = m.process_one_batch(u, i, r)
            loss.backward()
            m.optimizer.step()
            bu.append(u.numpy()); bi.append(i.numpy()); br.append(r.numpy()); lens.append(len(u)); losses.append(loss.item())
        out[f"{lf}_users"] = np.concatenate(bu).astype(np.int32); out[f"{lf}_items"] = np.concatenate(bi).astype(np.int32)
        out[f"{lf}_ratings"] = np.concatenate(br).astype(np.float32); out[f"{lf}_lens"] = np.array(lens, np.int32)
        out[f"{lf}_losses"] = np.array(losses, np.float32)
        out[f"{lf}_U"] = m.user_embedding.weight.detach().numpy().copy()
        out[f"{lf}_V"] = m.item_embedding.weight.detach().numpy().copy()
        out["num_batches_per_epoch"] = np.int32(len(gen))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ml100k_pointwise.npz"), **out)
    print("ml100k_pointwise ok", {k: out[k].tolist() for k in ("ce_losses", "mse_losses", "ce_lens")})


if __name__ == "__main__":
    main()
