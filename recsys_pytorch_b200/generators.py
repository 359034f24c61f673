"""Pairwise (BPR) batch generators - host-side mirror of the reference's
`data/generators.py::PairwiseGenerator` (:138-224).

Same constructor keywords and iteration contract (`__len__` = #batches,
`__iter__` yields `(users, pos, neg)` int64 tensors on `device`), two samplers:

* ``sampler='device'`` (default): the B200 path.  One triple per user per epoch
  (what the reference emits with ``num_positives_per_user=1``), users walked in a
  fresh on-device permutation each epoch; the positive is drawn uniformly from the
  user's CSR row and the negative uniformly over non-positives by rejection
  against the CSR row, inside the CUDA library (counter RNG keyed by
  (seed, step, triple)).  This implements the BPR that generators.py:178-189
  intends; it deliberately does NOT reproduce quirk Q2 (generators.py:184 draws
  the "positive" from the whole catalogue) nor Q3 (triples frozen at
  construction) - see DESIGN.md.
* ``sampler='reference'``: host numpy restatement of generators.py:168-224 call
  for call (same np.random draws in the same order), used for end-to-end parity
  at ml-100k size.  O(U*I) on the CPU, like the reference.
"""
from __future__ import annotations

import numpy as np
import torch

from . import engine


class PairwiseGenerator:
    def __init__(self, input_matrix, as_numpy=False, num_positives_per_user=-1, num_negatives=1, batch_size=32,
                 shuffle=True, device=None, sampler="device", seed=2020):
        self.input_matrix = input_matrix
        self.num_positives_per_user = num_positives_per_user
        self.num_negatives = num_negatives
        self.as_numpy = as_numpy
        self.batch_size = int(batch_size)
        self.shuffle = shuffle
        self.device = torch.device(device) if device is not None else torch.device("cuda")
        self.sampler = sampler
        self.seed = int(seed)
        self.epoch = 0
        if num_negatives != 1:
            raise NotImplementedError("num_negatives != 1 is never used on the BPR-MF path (models/MF.py:55)")
        self._construct()

    # ------------------------------------------------------------------ #
    def _construct(self):
        num_users, num_items = self.input_matrix.shape
        if self.sampler == "reference":
            self._data = self._sample_reference()
            self._num_data = len(self._data[0])
            self.users_unique = 0 < self.num_positives_per_user == 1
        elif self.sampler == "device":
            if self.num_positives_per_user != 1:
                raise NotImplementedError("device sampler emits one triple per user per epoch "
                                          "(num_positives_per_user=1, models/MF.py:55)")
            if isinstance(self.input_matrix, engine.DeviceCSR):
                self.csr = self.input_matrix
            else:
                self.csr = engine.DeviceCSR.from_scipy(self.input_matrix, self.device)
            deg = self.csr.indptr[1:] - self.csr.indptr[:-1]
            # users without positives emit no triple (generators.py:186-189)
            self._active_users = torch.nonzero(deg > 0).flatten().to(torch.int32)
            self._num_data = int(self._active_users.numel())
            self.users_unique = True
        else:
            raise ValueError(f"unknown sampler {self.sampler!r}")

    def _sample_reference(self):
        """generators.py:168-201 restated: identical np.random call sequence."""
        mat = self.input_matrix.tocsr()
        num_users, num_items = mat.shape
        users, positives, negatives = [], [], []
        for u in range(num_users):
            u_pos = mat.indices[mat.indptr[u]:mat.indptr[u + 1]]
            prob = np.ones(num_items)
            prob[u_pos] = 0.0
            prob = prob / sum(prob)          # python sum, as generators.py:180 (bit-identical p vector)
            n_pos = len(u_pos)
            if 0 < self.num_positives_per_user < n_pos:
                pos_s = np.random.choice(num_items, size=self.num_positives_per_user, replace=False)   # Q2
                neg_s = np.random.choice(num_items, size=self.num_positives_per_user, replace=False, p=prob)
            else:
                pos_s = u_pos
                neg_s = np.random.choice(num_items, size=n_pos, replace=False, p=prob)
            users += [u] * len(neg_s)
            positives += pos_s.tolist()
            negatives += neg_s.tolist()
        return np.array(users), np.array(positives), np.array(negatives)

    # ------------------------------------------------------------------ #
    def __len__(self):
        return int(np.ceil(self._num_data / self.batch_size))

    def iter_device(self):
        """B200 path: yields (users_int32, step_key) - the fused step samples pos/neg itself."""
        self.epoch += 1
        g = torch.Generator(device=self.device)
        g.manual_seed(self.seed * 1000003 + self.epoch)
        if self.shuffle:
            perm = torch.randperm(self._num_data, device=self.device, generator=g)
            users = self._active_users[perm]
        else:
            users = self._active_users
        nb = len(self)
        for b, st in enumerate(range(0, self._num_data, self.batch_size)):
            yield users[st:st + self.batch_size], self.epoch * nb + b

    def __iter__(self):
        if self.sampler == "device":
            for users, step in self.iter_device():
                pos, neg = engine.sample_triples(users, self.csr, self.seed, step)
                if self.as_numpy:
                    yield users.cpu().numpy(), pos.cpu().numpy(), neg.cpu().numpy()
                else:
                    yield users.long(), pos.long(), neg.long()
            return
        perm = np.random.permutation(self._num_data) if self.shuffle else np.arange(self._num_data)
        for st in range(0, self._num_data, self.batch_size):
            idx = perm[st:min(st + self.batch_size, self._num_data)]
            bu, bp, bn = self._data[0][idx], self._data[1][idx], self._data[2][idx]
            if not self.as_numpy:
                bu = torch.tensor(bu, dtype=torch.long, device=self.device)
                bp = torch.tensor(bp, dtype=torch.long, device=self.device)
                bn = torch.tensor(bn, dtype=torch.long, device=self.device)
            yield bu, bp, bn


class PointwiseGenerator:
    """Host mirror of `data/generators.py::PointwiseGenerator` (:43-136) - the batch source of the reference's
    POINTWISE MF mode (models/MF.py:49-52).  Same constructor keywords and iteration
    contract: `(users, items, ratings)` (int64, int64, float32 tensors on `device`) when `return_rating`, else
    `(users, items)`.  Reproduces the reference draw for draw (same np.random call sequence), including its quirk:
    `sample_negatives` ignores the batch it is handed and draws `num_negatives` unobserved items for EVERY user of the
    matrix, so each batch carries `num_users * num_negatives` extra zero-rated rows (generators.py:79-101,121-126).
    Host side only (numpy): the fused device step for this mode is the next kernel to write."""

    def __init__(self, input_matrix, return_rating=True, as_numpy=False, negative_sample=True, num_negatives=1,
                 batch_size=32, shuffle=True, device=None):
        self.input_matrix = input_matrix.tocsr()
        self.return_rating, self.as_numpy = return_rating, as_numpy
        self.negative_sample, self.num_negatives = negative_sample, num_negatives
        self.batch_size, self.shuffle = int(batch_size), shuffle
        self.device = torch.device(device) if device is not None else torch.device("cpu")
        m = self.input_matrix
        self.users = np.repeat(np.arange(m.shape[0]), np.diff(m.indptr))        # :58-70 without the per-user loop
        self.items = m.indices.astype(np.int64).copy()
        self.ratings = m.data.astype(np.float64).copy() if return_rating else np.array([])
        self._num_data = len(self.users)

    def sample_negatives(self, users=None):
        """generators.py:79-101: one exact-uniform draw over the non-positives per user, for ALL users."""
        m = self.input_matrix
        num_users, num_items = m.shape
        out_u, out_i = [], []
        for u in range(num_users):
            prob = np.ones(num_items)
            prob[m.indices[m.indptr[u]:m.indptr[u + 1]]] = 0.0
            prob = prob / sum(prob)                      # python sum, as generators.py:90 (bit-identical p vector)
            neg = np.random.choice(num_items, size=self.num_negatives, replace=False, p=prob)
            out_u += [u] * len(neg)
            out_i += neg.tolist()
        out_u, out_i = np.array(out_u), np.array(out_i)
        return out_u, out_i, np.zeros_like(out_u)

    def __len__(self):
        return int(np.ceil(self._num_data / self.batch_size))

    def __iter__(self):
        perm = np.random.permutation(self._num_data) if self.shuffle else np.arange(self._num_data)
        for st in range(0, self._num_data, self.batch_size):
            idx = perm[st:min(st + self.batch_size, self._num_data)]
            bu, bi = self.users[idx], self.items[idx]
            if not self.return_rating:
                if self.as_numpy:
                    yield bu, bi
                else:
                    yield (torch.tensor(bu, dtype=torch.long, device=self.device),
                           torch.tensor(bi, dtype=torch.long, device=self.device))
                continue
            br = self.ratings[idx]
            if self.negative_sample and self.num_negatives > 0:
                nu_, ni_, nr_ = self.sample_negatives(bu)
                bu, bi, br = np.concatenate((bu, nu_)), np.concatenate((bi, ni_)), np.concatenate((br, nr_))
            if self.as_numpy:
                yield bu, bi, br
            else:
                yield (torch.tensor(bu, dtype=torch.long, device=self.device),
                       torch.tensor(bi, dtype=torch.long, device=self.device),
                       torch.tensor(br, dtype=torch.float32, device=self.device))
