"""CPU (PyTorch) restatement of the reference's BPR-MF training step and chunked
evaluation, used ONLY as the timed CPU baseline (`bench.py` cpu_baseline leg and
`--impl reference`) - /root/reference itself does not exist on the GPU box, so
the reference's own Python cannot travel; this file restates its call sequence on
the same torch ops.  TEST/BENCH INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows (paths relative to /root/reference):
  models/MF.py:14-30    two nn.Embedding tables (N(0,1)), torch.optim.Adam(lr=1e-3)
  models/MF.py:38-42    forward = sum(mul(user_emb, item_emb), 1)
  models/MF.py:63-68    zero_grad -> process_one_batch -> backward -> optimizer.step
  models/MF.py:99-107   loss = -sigmoid(pos - neg).log().mean()
  models/MF.py:109-112  predict_batch_users = user_latent @ item_latent.T
  models/MF.py:130      -inf at the user's train positives (applied per chunk here:
                        the dense [U,I] float64 matrix of :117 cannot be allocated at
                        benchmark sizes)
  evaluation/backend/cython/include/func.h:22-31, holdout.h:20-103 via oracle/_ref
  (the reference's own C++) when that library is present, else oracle/liboracle.so.
"""
from __future__ import annotations

import ctypes as C
import os
import time

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))


class RefMF(nn.Module):
    def __init__(self, num_users, num_items, hidden_dim, init_std=1.0, optimizer="adam", lr=1e-3):
        super().__init__()
        self.user_embedding = nn.Embedding(num_users, hidden_dim)       # MF.py:23
        self.item_embedding = nn.Embedding(num_items, hidden_dim)       # MF.py:24
        if init_std != 1.0:
            with torch.no_grad():
                self.user_embedding.weight.mul_(init_std); self.item_embedding.weight.mul_(init_std)
        if optimizer == "adam":
            self.optimizer = torch.optim.Adam(self.parameters(), lr=lr)  # MF.py:30
        else:
            self.optimizer = torch.optim.SGD(self.parameters(), lr=lr)   # SURVEY H1 parity swap

    def forward(self, user_ids, item_ids):                               # MF.py:38-42
        return torch.sum(torch.mul(self.user_embedding(user_ids), self.item_embedding(item_ids)), 1)

    def process_one_batch(self, users, items, negs):                     # MF.py:99-107
        pos = self.forward(users, items)
        neg = self.forward(users, negs)
        return -torch.sigmoid(pos - neg).log().mean()

    def train_step(self, users, items, negs):                            # MF.py:64-68
        self.optimizer.zero_grad()
        loss = self.process_one_batch(users, items, negs)
        loss.backward()
        self.optimizer.step()
        return loss

    def predict_batch_users(self, user_ids):                             # MF.py:109-112
        return self.user_embedding(user_ids) @ self.item_embedding.weight.data.T


def native_eval_lib():
    """(lib, kind): the reference's own C++ (oracle/_ref) if built, else the C restatement."""
    ref = os.path.join(HERE, "_ref", "libref_eval.so")
    if os.path.exists(ref):
        lib = C.CDLL(ref)
        return (lib.ref_top_k_array_index, lib.ref_evaluate_holdout), "reference"
    lib = C.CDLL(os.path.join(HERE, "liboracle.so"))
    return (lib.oracle_top_k_array_index, lib.oracle_evaluate_holdout), "port"


def eval_chunk(model, users, mask_indptr, mask_indices, truth_indptr, truth_indices, k, fns):
    """predict_batch_users -> -inf mask -> top-k -> holdout metrics for one chunk of users.
    Returns (pairs scored, metric rows fp32 [n, 3])."""
    topk_fn, holdout_fn = fns
    with torch.no_grad():
        S = model.predict_batch_users(torch.from_numpy(users.astype(np.int64))).numpy()
    for r, u in enumerate(users):
        S[r, mask_indices[mask_indptr[u]:mask_indptr[u + 1]]] = -np.inf
    S = np.ascontiguousarray(S, np.float32)                              # evaluator.py:37 astype(float32)
    top = np.zeros((len(users), k), np.int32)
    topk_fn(C.c_void_p(S.ctypes.data), C.c_int(S.shape[1]), C.c_int(S.shape[0]), C.c_int(k), C.c_void_p(top.ctypes.data))
    truths = [np.ascontiguousarray(truth_indices[truth_indptr[u]:truth_indptr[u + 1]], np.int32) for u in users]
    ptrs = (C.c_void_p * len(truths))(*[t.ctypes.data for t in truths])
    lens = np.array([len(t) for t in truths], np.int32)
    ks = np.array([k], np.int32)
    rows = np.zeros((len(users), 3), np.float32)
    holdout_fn(C.c_int(len(users)), C.c_void_p(top.ctypes.data), C.c_int(k), C.c_void_p(ks.ctypes.data), C.c_int(1), ptrs,
               C.c_void_p(lens.ctypes.data), C.c_void_p(rows.ctypes.data))
    return S.shape[0] * S.shape[1], rows


def time_train(model, batches, warmup=1):
    """Seconds per step over `batches` (list of (u,i,j) int64 tensors) after `warmup` untimed steps."""
    for b in batches[:warmup]:
        model.train_step(*b)
    t0 = time.perf_counter()
    for b in batches[warmup:]:
        model.train_step(*b)
    return (time.perf_counter() - t0) / max(len(batches) - warmup, 1)
