"""Evaluation - host-side mirror of the reference's `evaluation/` package.

* `Evaluator` has the constructor and `evaluate(model, mean=True)` contract of
  `evaluation/evaluator.py:10-55`.  When the model offers `predict_topk_device`
  (the B200 MF / LightGCN do) scoring, masking, top-k, the per-user metrics and
  their means all stay on the GPU and only `len(metrics)*len(ks)` floats come
  back; otherwise it runs the reference's own sequence (`model.predict` -> dense
  float64 matrix -> `predict_topk` -> `eval_func`) with the GPU drop-ins below.
* `predict_topk_func`, `eval_func_router`, `HOLDOUT_METRICS`, `LOO_METRICS` are the
  module-level names of `evaluation/backend/__init__.py:1-29`, bound to the
  HOST-buffer C-ABI entry points that replace `c_top_k_array_index`,
  `evaluate_holdout`, `evaluate_loo` (include/b200rec.h group 1).
* `Statistics` mirrors `utils/stats.py:3-39`.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from typing import Iterable

import numpy as np
import torch

from . import _lib, engine
from ._lib import check

HOLDOUT_METRICS = ["Prec", "Recall", "NDCG"]     # evaluation/backend/__init__.py:1
LOO_METRICS = ["HR", "NDCG"]                     # evaluation/backend/__init__.py:2


class Statistics:
    """utils/stats.py:3-39 (list-append accumulator; mean = np.mean(history, dtype=float32))."""

    def __init__(self, name="AVG"):
        self.name = name
        self.history = []
        self.sum = 0
        self.cnt = 0

    def update(self, val):
        if isinstance(val, (list, np.ndarray)):
            val = [float(v) for v in val]
            self.history += val
            self.sum += sum(val)
            self.cnt += len(val)
        elif isinstance(val, (int, float, np.integer, np.floating)):
            self.history.append(float(val))
            self.sum += val
            self.cnt += 1
        else:
            raise TypeError("'val' should be float, int or list of them.")

    @property
    def mean(self):
        return np.mean(self.history, dtype=np.float32)

    @property
    def std(self):
        return np.std(self.history, dtype=np.float32)

    @property
    def mean_std(self):
        return self.mean, self.std

    def __repr__(self):
        return "%s: mean=%.4f, std=%.4f" % (self.name, self.mean, self.std)


def sparse_to_dict(sparse):
    """utils/types.py:13-21: CSR -> {row: indices} for ALL rows."""
    if isinstance(sparse, dict):
        return sparse
    return {i: sparse.indices[sparse.indptr[i]:sparse.indptr[i + 1]] for i in range(sparse.shape[0])}


# --------------------------------------------------------------------------- #
# drop-ins for the reference's native layer (host numpy in / out)
# --------------------------------------------------------------------------- #
def predict_topk_b200(scores, max_k):
    """evaluation/backend/cython/func.pyx:12-25 `predict_topk_cy` contract:
    scores fp32 [U,I] C-contiguous -> int32 [U,max_k], best first."""
    scores = np.ascontiguousarray(scores, dtype=np.float32)
    users_num, rank_len = scores.shape
    topk = np.zeros([users_num, max_k], dtype=np.int32)
    check(_lib.lib().b200rec_top_k_array_index(scores.ctypes.data, rank_len, users_num, int(max_k), topk.ctypes.data))
    return topk


def _truth_table(target):
    arrs = [np.ascontiguousarray(target[u], dtype=np.int32) for u in target]
    ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
    lens = np.array([len(a) for a in arrs], dtype=np.int32)
    return arrs, ptrs, lens


def compute_holdout(topk, target, metrics_num, Ks):
    """evaluation/backend/cython/holdout_func.pyx:12-49 contract -> fp32 [U, 3*len(Ks)]."""
    topk = np.ascontiguousarray(topk, dtype=np.int32)
    Ks = np.ascontiguousarray(Ks, dtype=np.int32)
    arrs, ptrs, lens = _truth_table(target)
    results = np.zeros([len(arrs), metrics_num * len(Ks)], dtype=np.float32)
    check(_lib.lib().b200rec_evaluate_holdout(len(arrs), topk.ctypes.data, int(max(Ks)), Ks.ctypes.data, len(Ks),
                                              C.cast(ptrs, C.c_void_p), lens.ctypes.data, results.ctypes.data))
    return results


def compute_loo(topk, target, metrics_num, Ks):
    """evaluation/backend/cython/loo_func.pyx:11-42 contract -> fp32 [U, 2*len(Ks)]."""
    topk = np.ascontiguousarray(topk, dtype=np.int32)
    Ks = np.ascontiguousarray(Ks, dtype=np.int32)
    arrs, ptrs, _ = _truth_table(target)
    results = np.zeros([len(arrs), metrics_num * len(Ks)], dtype=np.float32)
    check(_lib.lib().b200rec_evaluate_loo(len(arrs), topk.ctypes.data, int(max(Ks)), Ks.ctypes.data, len(Ks),
                                          C.cast(ptrs, C.c_void_p), results.ctypes.data))
    return results


def _cumulate(results, metrics, ks):
    cum = OrderedDict()
    for i, metric in enumerate(metrics):
        cum[metric] = {}
        for j, k in enumerate(ks):
            st = Statistics("%s@%d" % (metric, k))
            st.update(results[:, i * len(ks) + j].astype(np.float64))   # vectorised holdout.py:23-27
            cum[metric][k] = st
    return cum


def compute_holdout_metrics_b200(pred, target, ks):
    """evaluation/backend/cython/holdout.py:14-29 contract."""
    return _cumulate(compute_holdout(pred.astype(np.int32), target, len(HOLDOUT_METRICS), np.array(ks, np.int32)),
                     HOLDOUT_METRICS, ks)


def compute_loo_metrics_b200(pred, target, ks):
    """evaluation/backend/cython/loo.py:15-27 contract."""
    return _cumulate(compute_loo(pred.astype(np.int32), target, len(LOO_METRICS), np.array(ks, np.int32)),
                     LOO_METRICS, ks)


eval_func_router = {"leave_one_out": compute_loo_metrics_b200, "holdout": compute_holdout_metrics_b200}
predict_topk_func = predict_topk_b200


# --------------------------------------------------------------------------- #
class Evaluator:
    def __init__(self, eval_input, eval_target, protocol, ks, eval_batch_size=1024):
        self.top_k = sorted(list(ks)) if isinstance(ks, Iterable) else [ks]
        self.max_k = max(self.top_k)
        self.batch_size = eval_batch_size
        self.eval_input = eval_input
        self._target_csr = eval_target if not isinstance(eval_target, dict) else None
        self.eval_target = sparse_to_dict(eval_target)
        self.protocol = protocol
        self._register_eval_func()
        self._dev = {}
        self._eval_users = None

    def _register_eval_func(self):
        self.eval_func = eval_func_router[self.protocol]
        self.predict_topk = predict_topk_func

    def _truth_device(self, device):
        if "truth" not in self._dev:
            if self._target_csr is not None:
                indptr = np.ascontiguousarray(self._target_csr.indptr, np.int64)
                indices = np.ascontiguousarray(self._target_csr.indices, np.int32)
                shape = self._target_csr.shape
            else:
                keys = list(self.eval_target.keys())
                lens = [len(self.eval_target[u]) for u in keys]
                indptr = np.zeros(max(keys) + 2, np.int64)
                indptr[np.asarray(keys) + 1] = lens
                indptr = np.cumsum(indptr)
                order = np.argsort(keys, kind="stable")
                indices = np.concatenate([np.asarray(self.eval_target[keys[o]], np.int32) for o in order]) \
                    if keys else np.zeros(0, np.int32)
                shape = (len(indptr) - 1, int(indices.max()) + 1 if len(indices) else 1)
            self._dev["truth"] = engine.DeviceCSR(torch.from_numpy(indptr).to(device),
                                                  torch.from_numpy(indices).to(device), shape)
        return self._dev["truth"]

    def evaluate(self, model, mean=True):
        model.eval()
        if self._eval_users is None or len(self._eval_users) != len(self.eval_target):   # evaluator.py:34, cached
            self._eval_users = np.array(list(self.eval_target.keys()))
        eval_users = self._eval_users
        if hasattr(model, "predict_topk_device"):
            return self._evaluate_fused(model, eval_users, mean)
        # reference sequence, evaluation/evaluator.py:35-48, with the GPU drop-ins
        output = model.predict(eval_users, self.eval_input, self.batch_size)
        pred = self.predict_topk(output.astype(np.float32), self.max_k)
        score_cumulator = self.eval_func(pred, self.eval_target, self.top_k)
        scores = {}
        for metric in score_cumulator:
            for k in score_cumulator[metric]:
                st = score_cumulator[metric][k]
                scores["%s@%d" % (metric, k)] = st.mean if mean else st.history
        return scores

    def _evaluate_fused(self, model, eval_users, mean):
        device = model.device
        if "users" not in self._dev or self._dev["users"].numel() != len(eval_users):
            self._dev["users"] = torch.from_numpy(eval_users.astype(np.int32)).to(device)
        users = self._dev["users"]
        truth = self._truth_device(device)
        metrics = HOLDOUT_METRICS if self.protocol == "holdout" else LOO_METRICS
        fn = engine.holdout_metrics if self.protocol == "holdout" else engine.loo_metrics
        nk = len(self.top_k)
        sums = np.zeros(len(metrics) * nk, np.float64)
        hist = []
        # users per fused call: 8 launches' worth of the tensor-core kernel (148 SMs x 256 rows x 2 waves = 75,776 rows per
        # launch, csrc/score_tc.cu): whole waves, and the item-side pre-pass of a call is shared by all of them
        chunk = 75_776 * 8
        for st in range(0, len(eval_users), chunk):
            u = users[st:st + chunk]
            idx, _ = model.predict_topk_device(u, self.eval_input, self.max_k)
            rows = fn(idx, truth, self.top_k, row_ids=u)
            if mean:
                sums += engine.column_means(rows) * u.numel()
            else:
                hist.append(rows.cpu().numpy())
        scores = {}
        if mean:
            means = (sums / max(len(eval_users), 1)).astype(np.float32)
            for i, metric in enumerate(metrics):
                for j, k in enumerate(self.top_k):
                    scores["%s@%d" % (metric, k)] = means[i * nk + j]
        else:
            allrows = np.concatenate(hist, 0)
            for i, metric in enumerate(metrics):
                for j, k in enumerate(self.top_k):
                    scores["%s@%d" % (metric, k)] = allrows[:, i * nk + j].astype(np.float64).tolist()
        return scores
