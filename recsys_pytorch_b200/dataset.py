"""Data ingest -> CSR: mirror of the reference's `UIRTDataset` (data/dataset.py:12-199) and
its per-user split (data/preprocess.py:9-90) without the pandas group-by loops.

Two split engines:

``split='reference'``  reproduces the reference draw for draw: users in ascending id order, each user's rows ordered by
    ``argsort(timestamp, kind='quicksort')`` (what ``DataFrame.sort_values`` runs, preprocess.py:64), the held-out
    positions drawn with ``np.random.choice(n, k, replace=False)`` (preprocess.py:68) from numpy's GLOBAL generator - so
    after ``np.random.seed(s)`` the train / valid / test matrices equal the reference's bit for bit
    (tests/test_dataset_cpu.py against tests/golden/ml100k.npz).  One cheap numpy call per user instead of half a dozen
    DataFrame operations.

``split='device'``     the same split law, vectorised with torch on any device (one random key per interaction, one
    sort): for catalogues where a Python loop over users is the bottleneck (10M users at cfg3/cfg5).  Same held-out
    COUNT per user as the reference (``ceil(ratio * n_u)`` / ``leave_k``), different draws.

Reference quirks kept: with ``generalization='weak'`` the FIRST split uses ``valid_ratio`` and becomes the
*test* matrix, the SECOND uses ``test_ratio`` and becomes the *valid* matrix (preprocess.py:14-15); ``'strong'`` and
``binarize_threshold > 0`` are broken upstream (dataset.py:87, :49) and are refused here.  The CSV cache the reference
writes next to its input (dataset.py:185-189) is not reproduced.
"""
from __future__ import annotations

import math
import os

import numpy as np
import scipy.sparse as sp
import torch

__all__ = ["UIRTDataset", "DeviceIngest", "ingest_device", "split_by_user_reference", "split_by_user_device"]


def _read_uirt(path, sep):
    """dataset.py:115-127: columns user, item, rating, timestamp; missing rating / timestamp columns become ones."""
    import pandas as pd
    df = pd.read_csv(path, sep=sep, header=None, names=["user", "item", "rating", "timestamp"], engine="c" if len(sep) == 1 else "python")
    n = len(df)
    u = df["user"].to_numpy(np.int64)
    i = df["item"].to_numpy(np.int64)
    r = df["rating"].to_numpy(np.float64)
    t = df["timestamp"].to_numpy(np.float64)
    if n and np.isnan(r[0]):
        r = np.ones(n)
    if n and np.isnan(t[0]):
        t = np.ones(n)
    return u, i, r, t


def _held_out_count(n, ratio):
    return int(math.ceil(ratio * n)) if isinstance(ratio, float) else int(ratio)     # preprocess.py:59-62


def split_by_user_reference(u, t, ratio, split_random=True):
    """preprocess.py:52-90 on arrays.  `u` (new user ids) and `t` (timestamps) list the interactions in the order the
    reference's DataFrame holds them.  Returns (keep_rows, held_rows): row numbers into the input, each in the order the
    reference concatenates them (ascending user, timestamp-sorted inside a user).  Consumes numpy's global RNG exactly
    like the reference."""
    order = np.argsort(u, kind="stable")                 # groupby('user'): ascending key, rows in original order
    us = u[order]
    bounds = np.flatnonzero(np.r_[True, us[1:] != us[:-1], True])
    keep, held = [], []
    for b in range(len(bounds) - 1):
        rows = order[bounds[b]:bounds[b + 1]]
        rows = rows[np.argsort(t[rows], kind="quicksort")]          # group.sort_values(by='timestamp')
        n = len(rows)
        k = _held_out_count(n, ratio)
        idx = np.ones(n, dtype=bool)
        if split_random:
            idx[np.random.choice(n, k, replace=False)] = False
        else:
            idx[-k:] = False                                       # (k == 0 holds out everything, as upstream: idx[-0:])
        keep.append(rows[idx]); held.append(rows[~idx])
    cat = lambda parts: np.concatenate(parts) if parts else np.zeros(0, np.int64)
    return cat(keep), cat(held)


def split_by_user_device(u, t, ratio, split_random=True, generator=None):
    """The same split law, vectorised (torch, any device): every user keeps n - k rows, k = ceil(ratio * n) or `ratio`
    when it is an int; random: the k rows with the smallest random keys are held out; time: the k latest.
    Returns boolean mask `held` over the interactions."""
    u = torch.as_tensor(u)
    dev = u.device
    n = u.numel()
    if n == 0:
        return torch.zeros(0, dtype=torch.bool, device=dev)
    if split_random:
        key = torch.rand(n, device=dev, generator=generator, dtype=torch.float64)
    else:
        tt = torch.as_tensor(t, device=dev).double()
        key = -(torch.argsort(torch.argsort(tt, stable=True), stable=True).double())   # latest first; ties by position
    # sort by (user, key): a stable sort by key followed by a stable sort by user
    o1 = torch.argsort(key, stable=True)
    o2 = torch.argsort(u[o1], stable=True)
    order = o1[o2]
    us = u[order]
    num_users = int(us.max().item()) + 1
    deg = torch.bincount(us, minlength=num_users)
    start = torch.cumsum(deg, 0) - deg
    rank = torch.arange(n, device=dev) - start[us]
    if isinstance(ratio, float):
        k = torch.ceil(deg.double() * ratio).long()
    else:
        k = torch.full_like(deg, int(ratio))
    held_sorted = rank < k[us]
    held = torch.zeros(n, dtype=torch.bool, device=dev)
    held[order] = held_sorted
    return held


class DeviceIngest:
    """Result of `ingest_device`: everything stays on the device the arrays came from.  `parts[name]` = (indptr int64
    [U+1], indices int32 sorted per row) for name in train_data / valid_target / test_target; `raw_users` / `raw_items`
    map new id -> original id (the inverse of the reference's user2id / item2id, dataset.py:155-162)."""

    def __init__(self, num_users, num_items, raw_users, raw_items, parts, protocol):
        self.num_users, self.num_items = int(num_users), int(num_items)
        self.raw_users, self.raw_items, self.parts, self.protocol = raw_users, raw_items, parts, protocol

    def device_csr(self, name):
        from . import engine
        if name == "valid_input":
            name = "train_data"
        indptr, indices = self.parts[name]
        return engine.DeviceCSR(indptr, indices, (self.num_users, self.num_items))

    def to_scipy(self, name):
        indptr, indices = self.parts["train_data" if name == "valid_input" else name]
        return sp.csr_matrix((np.ones(indices.numel()), indices.cpu().numpy(), indptr.cpu().numpy()),
                             shape=(self.num_users, self.num_items))


def _csr_from_pairs(u, i, num_users, num_items):
    """Sorted, de-duplicated CSR (implicit ones) from (row, col) pairs - torch sort / unique / bincount on any device."""
    key = torch.unique(u * num_items + i)                          # sorted; utils/types.py:5-11 + sum_duplicates
    rows = torch.div(key, num_items, rounding_mode="floor")
    indptr = torch.zeros(num_users + 1, dtype=torch.int64, device=u.device)
    indptr[1:] = torch.cumsum(torch.bincount(rows, minlength=num_users), 0)
    return indptr.contiguous(), (key - rows * num_items).to(torch.int32).contiguous()


def ingest_device(users, items, timestamps=None, min_item_per_user=0, min_user_per_item=0, protocol="holdout",
                  valid_ratio=0.1, test_ratio=0.2, leave_k=1, split_random=True, seed=1234, device=None):
    """The whole ingest of data/dataset.py:129-199 + data/preprocess.py:9-90 as device array code: ONE
    pass of the user filter, user ids fixed BEFORE the item filter (:133-158), ONE pass of the item filter, dense re-id
    in ascending raw-id order, the two per-user splits (the first - sized by valid_ratio / leave_k - becomes TEST, the
    second VALID: quirk Q6), CSR with sorted int32 columns.  `users` / `items`: raw integer ids of the interactions
    (any integer tensor / array); implicit feedback (ratings are ones).  No pandas, no Python loop over users; the same
    code runs on CPU tensors (tests) and CUDA tensors (10M-user catalogues)."""
    dev = torch.device(device) if device is not None else (users.device if isinstance(users, torch.Tensor) else torch.device("cpu"))
    u = torch.as_tensor(users).to(dev, torch.int64)
    i = torch.as_tensor(items).to(dev, torch.int64)
    t = torch.as_tensor(timestamps).to(dev, torch.float64) if timestamps is not None else torch.ones(u.numel(), dtype=torch.float64, device=dev)
    _, inv, cnt = torch.unique(u, return_inverse=True, return_counts=True)
    keep = cnt[inv] >= min_item_per_user
    u, i, t = u[keep], i[keep], t[keep]
    raw_users = torch.unique(u)                                     # user2id is built BEFORE the item filter (:155-158)
    _, inv, cnt = torch.unique(i, return_inverse=True, return_counts=True)
    keep = cnt[inv] >= min_user_per_item
    u, i, t = u[keep], i[keep], t[keep]
    raw_items = torch.unique(i)
    nu, ni = int(raw_users.numel()), int(raw_items.numel())
    u = torch.searchsorted(raw_users, u)
    i = torch.searchsorted(raw_items, i)
    first, second = (leave_k, leave_k) if protocol == "leave_one_out" else (valid_ratio, test_ratio)
    g = torch.Generator(device=dev); g.manual_seed(int(seed))
    held1 = split_by_user_device(u, t, first, split_random, g)
    rest = torch.nonzero(~held1).flatten()
    held2 = split_by_user_device(u[rest], t[rest], second, split_random, g)
    tr, va = rest[~held2], rest[held2]
    parts = {"train_data": _csr_from_pairs(u[tr], i[tr], nu, ni), "valid_target": _csr_from_pairs(u[va], i[va], nu, ni),
             "test_target": _csr_from_pairs(u[held1], i[held1], nu, ni)}
    return DeviceIngest(nu, ni, raw_users, raw_items, parts, protocol)


class UIRTDataset:
    """Same constructor and attributes as data/dataset.py:12-41 (weak generalisation): `num_users`, `num_items`,
    `train_data`, `valid_target`, `test_target`, `valid_input`, `test_input` (scipy CSR, float64 ones), `user2id`,
    `item2id`, `protocol`, `dataname`; plus `device_csr(name, device)` for the engine."""

    def __init__(self, data_path, dataname=None, separator=",", binarize_threshold=0.0, implicit=True,
                 min_item_per_user=0, min_user_per_item=0, protocol="holdout", generalization="weak", holdout_users=0.1,
                 valid_ratio=0.1, test_ratio=0.2, leave_k=1, split_random=True, cache_dir="cache", seed=1234,
                 split="reference", device=None):
        if generalization != "weak":
            raise NotImplementedError("generalization='strong' is broken upstream (data/dataset.py:87 assigns to a "
                                      "setter-less property); only 'weak' is live")
        if implicit and binarize_threshold > 0:
            raise NotImplementedError("binarize_threshold > 0 hits an undefined attribute upstream (data/dataset.py:49)")
        if protocol not in ("holdout", "leave_one_out"):
            raise ValueError(f"{protocol} is not a valid protocol.")
        if split not in ("reference", "device"):
            raise ValueError("split must be 'reference' or 'device'")
        self.data_path = str(data_path)
        self.dataname = dataname if dataname is not None else os.path.basename(os.path.dirname(os.path.abspath(self.data_path)))
        self.separator, self.implicit = separator, implicit
        self.min_item_per_user, self.min_user_per_item = min_item_per_user, min_user_per_item
        self.protocol, self.generalization, self.holdout_users = protocol, generalization, holdout_users
        self.valid_ratio, self.test_ratio, self.leave_k, self.split_random = valid_ratio, test_ratio, leave_k, split_random
        self.seed, self.cache_dir, self.split = seed, cache_dir, split

        u, i, r, t = _read_uirt(self.data_path, separator)
        # dataset.py:133-152: ONE pass of each filter, users first
        ids, inv, cnt = np.unique(u, return_inverse=True, return_counts=True)
        keep = cnt[inv] >= min_item_per_user
        u, i, r, t = u[keep], i[keep], r[keep], t[keep]
        raw_users = np.unique(u)                                   # user2id is built BEFORE the item filter (:155-158)
        ids, inv, cnt = np.unique(i, return_inverse=True, return_counts=True)
        keep = cnt[inv] >= min_user_per_item
        u, i, r, t = u[keep], i[keep], r[keep], t[keep]
        raw_items = np.unique(i)
        self.user2id = {int(x): n for n, x in enumerate(raw_users)}
        self.item2id = {int(x): n for n, x in enumerate(raw_items)}
        self.num_users, self.num_items = len(raw_users), len(raw_items)
        u = np.searchsorted(raw_users, u).astype(np.int64)
        i = np.searchsorted(raw_items, i).astype(np.int64)
        if implicit:
            r = np.ones(len(r))

        # preprocess.py:12-19 (Q6: the first split - sized by valid_ratio / leave_k - becomes TEST, the second VALID)
        first, second = (leave_k, leave_k) if protocol == "leave_one_out" else (valid_ratio, test_ratio)
        if split == "reference":
            rest, test_rows = split_by_user_reference(u, t, first, split_random)
            keep2, held2 = split_by_user_reference(u[rest], t[rest], second, split_random)
            train_rows, valid_rows = rest[keep2], rest[held2]
        else:
            dev = torch.device(device) if device is not None else torch.device("cpu")
            g = torch.Generator(device=dev); g.manual_seed(int(seed))
            ud, td = torch.from_numpy(u).to(dev), torch.from_numpy(t).to(dev)
            held1 = split_by_user_device(ud, td, first, split_random, g)
            rest_t = torch.nonzero(~held1).flatten()
            held2 = split_by_user_device(ud[rest_t], td[rest_t], second, split_random, g)
            test_rows = torch.nonzero(held1).flatten().cpu().numpy()
            train_rows, valid_rows = rest_t[~held2].cpu().numpy(), rest_t[held2].cpu().numpy()

        shape = (self.num_users, self.num_items)
        mk = lambda rows: sp.csr_matrix((r[rows], (u[rows], i[rows])), shape=shape)      # utils/types.py:5-11
        self.train_data, self.valid_target, self.test_target = mk(train_rows), mk(valid_rows), mk(test_rows)
        self.train_users = self.valid_users = self.test_users = list(np.unique(u[train_rows]))

    @property
    def valid_input(self):                                           # dataset.py:239-243
        return self.train_data

    @property
    def test_input(self):                                            # dataset.py:246-250
        return self.train_data + self.valid_target

    @property
    def num_train_users(self):
        return len(self.train_users)

    def device_csr(self, name, device):
        """`train_data` / `valid_target` / `test_target` / `valid_input` / `test_input` as an engine DeviceCSR."""
        from . import engine
        m = getattr(self, name).tocsr()
        m.sum_duplicates(); m.sort_indices()
        return engine.DeviceCSR.from_scipy(m, torch.device(device))
