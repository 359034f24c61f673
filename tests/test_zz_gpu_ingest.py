"""`-m gpu` (sorted last on purpose): the on-device ingest (recsys_pytorch_b200/dataset.py::ingest_device)
on CUDA tensors against the same array code on CPU tensors, which tests/test_dataset_cpu.py pins to the host mirror of the
reference's UIRTDataset.  With the time-based split (split_random=False) nothing is random: every CSR must be identical."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_ingest_device_on_cuda_equals_cpu(dev):
    from recsys_pytorch_b200.dataset import ingest_device
    rng = np.random.default_rng(2)
    nu_raw, ni_raw = 3000, 900
    deg = rng.integers(1, 60, nu_raw)
    uu = np.repeat(rng.permutation(100_000)[:nu_raw], deg)
    ii = np.concatenate([rng.choice(np.arange(7000, 7000 + ni_raw), n, replace=False) for n in deg])
    tt = rng.permutation(len(uu)).astype(np.float64)                    # distinct timestamps: a unique time order
    perm = rng.permutation(len(uu)); uu, ii, tt = uu[perm], ii[perm], tt[perm]
    for protocol in ("holdout", "leave_one_out"):
        kw = dict(min_item_per_user=10, min_user_per_item=3, protocol=protocol, valid_ratio=0.1, test_ratio=0.2, leave_k=2,
                  split_random=False)
        c = ingest_device(torch.from_numpy(uu), torch.from_numpy(ii), torch.from_numpy(tt), **kw)
        g = ingest_device(torch.from_numpy(uu).to(dev), torch.from_numpy(ii).to(dev), torch.from_numpy(tt).to(dev), **kw)
        assert (g.num_users, g.num_items) == (c.num_users, c.num_items)
        assert torch.equal(g.raw_users.cpu(), c.raw_users) and torch.equal(g.raw_items.cpu(), c.raw_items)
        for name in ("train_data", "valid_target", "test_target"):
            assert g.parts[name][0].is_cuda and torch.equal(g.parts[name][0].cpu(), c.parts[name][0]), name
            assert torch.equal(g.parts[name][1].cpu(), c.parts[name][1]), name
        csr = g.device_csr("train_data")                                  # what the engine consumes
        assert csr.shape == (c.num_users, c.num_items) and csr.nnz == int(c.parts["train_data"][1].numel())
    # random split on the device: same law (held-out count per user), different draws
    kw["split_random"] = True
    g = ingest_device(torch.from_numpy(uu).to(dev), torch.from_numpy(ii).to(dev), torch.from_numpy(tt).to(dev), seed=3, **kw)
    for name in ("train_data", "valid_target", "test_target"):
        assert torch.equal(torch.diff(g.parts[name][0]).cpu(), torch.diff(c.parts[name][0])), name
