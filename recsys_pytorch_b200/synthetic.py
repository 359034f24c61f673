"""Synthetic interaction matrices of the shapes BASELINE.json names (the reference
has no generator; recipe from SURVEY.md section 8(d)):

per-user degree = clip(round(lognormal(mu=3.5, sigma=0.8)), 10, 1000); items drawn
from Zipf(alpha=1) over a random permutation of item ids (duplicates inside a user
removed, so realised degrees are slightly lower); CSR with sorted int32 columns and
int64 indptr; holdout: ceil(20%) of each user's items become validation targets
(every user keeps >= 1 train and >= 1 target item).

Runs on the GPU with torch (data plumbing, not the measured path).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .engine import DeviceCSR


def _csr_from_keys(keys, num_users, num_items):
    keys = torch.unique(keys)                      # sorted, de-duplicated (user*I + item)
    rows = torch.div(keys, num_items, rounding_mode="floor")
    cols = (keys - rows * num_items).to(torch.int32)
    counts = torch.bincount(rows, minlength=num_users)
    indptr = torch.zeros(num_users + 1, dtype=torch.int64, device=keys.device)
    indptr[1:] = torch.cumsum(counts, 0)
    return indptr, cols, rows


def make_interactions(num_users, num_items, seed=2020, device="cuda", **kw):
    """Returns (train: DeviceCSR, target: DeviceCSR); see `make_interactions_raw` for the recipe and options."""
    (tp, ti), (vp, vi) = make_interactions_raw(num_users, num_items, seed=seed, device=device, **kw)
    return DeviceCSR(tp, ti, (num_users, num_items)), DeviceCSR(vp, vi, (num_users, num_items))


def make_interactions_raw(num_users, num_items, seed=2020, device="cuda", mu=3.5, sigma=0.8, dmin=10, dmax=1000,
                          alpha=1.0, holdout_frac=0.2, chunk_users=2_000_000, item_seed=None):
    """Returns ((train_indptr int64, train_indices int32), (target_indptr, target_indices)) as plain tensors on
    `device` (which may be the CPU: the host arm of bench.py builds its sample of the dataset with the same recipe).  `item_seed` (multi-GPU shards of ONE dataset): the popularity
    order of the catalogue comes from its own generator, so ranks that draw different users (`seed`) still agree on
    which items are popular."""
    device = torch.device(device)
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    deg = torch.exp(torch.randn(num_users, generator=g, device=device) * sigma + mu).round().clamp_(dmin, min(dmax, num_items // 2)).long()
    ranks = torch.arange(1, num_items + 1, device=device, dtype=torch.float64)
    cdf = torch.cumsum(ranks.pow(-alpha), 0)
    cdf = (cdf / cdf[-1]).to(torch.float32)
    if item_seed is None:
        item_perm = torch.randperm(num_items, generator=g, device=device)
    else:
        gi = torch.Generator(device=device)
        gi.manual_seed(item_seed)
        item_perm = torch.randperm(num_items, generator=gi, device=device)
    tr_ptr, tr_idx, va_ptr, va_idx = [], [], [], []
    base_tr = base_va = 0
    for u0 in range(0, num_users, chunk_users):
        u1 = min(u0 + chunk_users, num_users)
        d = deg[u0:u1]
        owner = torch.repeat_interleave(torch.arange(u1 - u0, device=device), d)
        r = torch.rand(owner.numel(), generator=g, device=device)
        items = item_perm[torch.searchsorted(cdf, r).clamp_(max=num_items - 1)]
        indptr, cols, rows = _csr_from_keys(owner * num_items + items, u1 - u0, num_items)
        # holdout: a random ceil(frac) subset of each row -> target, rest -> train
        rdeg = indptr[1:] - indptr[:-1]
        n_tgt = torch.clamp(torch.ceil(rdeg.double() * holdout_frac).long(), min=1)
        n_tgt = torch.minimum(n_tgt, torch.clamp(rdeg - 1, min=0))
        prio = torch.rand(cols.numel(), generator=g, device=device)
        order = torch.argsort(rows.double() + prio.double() * 0.999999)       # random order inside each row
        pos_in_row = torch.arange(cols.numel(), device=device) - indptr[rows[order]]
        is_tgt = torch.zeros(cols.numel(), dtype=torch.bool, device=device)
        is_tgt[order] = pos_in_row < n_tgt[rows[order]]
        for mask, ptrs, idxs, which in ((~is_tgt, tr_ptr, tr_idx, 0), (is_tgt, va_ptr, va_idx, 1)):
            c = torch.bincount(rows[mask], minlength=u1 - u0)
            p = torch.cumsum(c, 0)
            base = base_tr if which == 0 else base_va
            ptrs.append(p + base)
            idxs.append(cols[mask])
            if which == 0:
                base_tr += int(p[-1])
            else:
                base_va += int(p[-1])
    z = torch.zeros(1, dtype=torch.int64, device=device)
    return ((torch.cat([z] + tr_ptr).contiguous(), torch.cat(tr_idx).contiguous()),
            (torch.cat([z] + va_ptr).contiguous(), torch.cat(va_idx).contiguous()))


def to_scipy(csr: DeviceCSR):
    import scipy.sparse as sp
    indptr = csr.indptr.cpu().numpy()
    indices = csr.indices.cpu().numpy()
    return sp.csr_matrix((np.ones(len(indices), np.float32), indices, indptr), shape=csr.shape)
