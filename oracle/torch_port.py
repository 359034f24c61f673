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


# ---------------------------------------------------------------------------------------------------------------
# LightGCN propagation (BASELINE configs[3]) - same torch ops as the reference
# ---------------------------------------------------------------------------------------------------------------
def lightgcn_graph(indptr, indices, num_users, num_items):
    """The tensor models/LightGCN.py:228-267 ends with: A_hat = D^-1/2 [[0,R],[R^T,0]] D^-1/2 as a coalesced fp32
    torch sparse COO matrix of size (U+I)^2 (:259-266 `_convert_sp_mat_to_sp_tensor` + `.coalesce()`), built here
    from the train CSR without the scipy dok/lil detour."""
    indptr = np.asarray(indptr, np.int64); indices = np.asarray(indices, np.int64)
    rows = np.repeat(np.arange(num_users, dtype=np.int64), np.diff(indptr))
    deg_u = np.diff(indptr).astype(np.float64)
    deg_i = np.bincount(indices, minlength=num_items).astype(np.float64)
    with np.errstate(divide="ignore"):
        du, di = np.power(deg_u, -0.5), np.power(deg_i, -0.5)          # :249-250
    du[np.isinf(du)] = 0.0; di[np.isinf(di)] = 0.0
    v = ((du[rows].astype(np.float32) * np.float32(1.0)) * di[indices].astype(np.float32)).astype(np.float32)
    r = np.concatenate([rows, indices + num_users]); c = np.concatenate([indices + num_users, rows])
    n = num_users + num_items
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")                                # torch's "sparse invariant checks" notice
        g = torch.sparse_coo_tensor(torch.from_numpy(np.stack([r, c])), torch.from_numpy(np.concatenate([v, v])), (n, n))
        return g.coalesce()


def lightgcn_embedding(graph, all_emb, num_layers):
    """models/LightGCN.py:174-202 (node_dropout = 0, split = False): L sparse.mm's, stack, mean over layers."""
    embs = [all_emb]
    for _ in range(num_layers):
        all_emb = torch.sparse.mm(graph, all_emb)                      # :196
        embs.append(all_emb)
    return torch.mean(torch.stack(embs, dim=1), dim=1)                 # :198-200


def time_lightgcn(graph, d, num_layers, reps=3):
    """(seconds per forward propagation, seconds per forward + backward) - the reference's autograd walks the same L
    sparse.mm's backwards; median of `reps` after one warm-up."""
    n = graph.shape[0]
    E = (torch.randn(n, d) * 0.01).requires_grad_(True)
    fw, fb = [], []
    for r in range(reps + 1):
        t0 = time.perf_counter()
        with torch.no_grad():
            lightgcn_embedding(graph, E, num_layers)
        t1 = time.perf_counter()
        out = lightgcn_embedding(graph, E, num_layers)
        out.sum().backward()
        t2 = time.perf_counter()
        E.grad = None
        if r:
            fw.append(t1 - t0); fb.append(t2 - t1)
    return float(np.median(fw)), float(np.median(fb))
