"""Import the Python reference from /root/reference (read-only) with the three
arithmetic-neutral shims of SURVEY.md section 8(c).  TEST INFRASTRUCTURE ONLY; usable
only in the build container (the GPU box has no /root/reference).

Shims: (1) np.int/np.float aliases for utils/stats.py:15; (2) stub `neptune`
and bypass OmegaConf (config.py:4) by building hparams / exp_config by hand;
(3) writable working copy for everything the reference writes next to its
inputs (dataset cache data/dataset.py:211-227, ./graph LightGCN.py:38-39).
"""
from __future__ import annotations

import ctypes
import os
import shutil
import sys
import types

import numpy as np

REF = os.environ.get("B200REC_REFERENCE", "/root/reference")
WORK = os.environ.get("B200REC_REF_WORK", "/tmp/b200rec_refwork")
HERE = os.path.dirname(os.path.abspath(__file__))


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "models", "MF.py"))


def load():
    """Returns a namespace with the reference classes."""
    if not available():
        raise RuntimeError(f"reference not present at {REF}")
    np.int = int          # shim 1
    np.float = float
    sys.modules.setdefault("neptune", types.ModuleType("neptune"))  # shim 2a
    if REF not in sys.path:
        sys.path.insert(0, REF)
    os.makedirs(WORK, exist_ok=True)
    os.chdir(WORK)        # shim 3: relative writes (./graph, ./saves) land here
    import torch
    torch.set_num_threads(1)
    from data.dataset import UIRTDataset
    from data.generators import PairwiseGenerator
    from evaluation.evaluator import Evaluator
    from evaluation.backend.python.func import predict_topk_py
    from evaluation.backend.python.holdout import compute_holdout_metrics_py
    from evaluation.backend.python.loo import compute_loo_metrics_py
    from models.MF import MF
    from models.LightGCN import LightGCN
    from utils.general import set_random_seed
    return types.SimpleNamespace(**locals())


def ml100k_dataset(ref):
    """config.py:7-23 defaults on a writable copy of datasets/ml-100k/u.data."""
    dst = os.path.join(WORK, "ml-100k")
    os.makedirs(dst, exist_ok=True)
    if not os.path.exists(os.path.join(dst, "u.data")):
        shutil.copy(os.path.join(REF, "datasets", "ml-100k", "u.data"), dst)
    return ref.UIRTDataset(
        data_path=os.path.join(dst, "u.data"), dataname="ml-100k", separator="\t",
        binarize_threshold=0.0, implicit=True, min_item_per_user=10, min_user_per_item=1,
        protocol="holdout", generalization="weak", holdout_users=600,
        valid_ratio=0.1, test_ratio=0.2, leave_k=1, split_random=True)


def exp_config(num_epochs=10, batch_size=256):
    """config.py:35-47 without OmegaConf (shim 2b)."""
    return types.SimpleNamespace(num_epochs=num_epochs, batch_size=batch_size, verbose=0,
                                 test_from=1, test_step=1)


class RefNative:
    """ctypes view of oracle/_ref/libref_eval.so (the reference's own C++)."""

    def __init__(self):
        path = os.path.join(HERE, "_ref", "libref_eval.so")
        self.lib = ctypes.CDLL(path)

    def topk(self, scores, k):
        scores = np.ascontiguousarray(scores, np.float32)
        out = np.zeros((scores.shape[0], k), np.int32)
        self.lib.ref_top_k_array_index(
            scores.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(scores.shape[1]),
            ctypes.c_int(scores.shape[0]), ctypes.c_int(k), out.ctypes.data_as(ctypes.c_void_p))
        return out

    @staticmethod
    def _truth_table(truths):
        arrs = [np.ascontiguousarray(t, np.int32) for t in truths]
        ptrs = (ctypes.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
        lens = np.array([len(a) for a in arrs], np.int32)
        return arrs, ptrs, lens

    def holdout(self, topk, truths, ks):
        topk = np.ascontiguousarray(topk, np.int32)
        ks = np.ascontiguousarray(ks, np.int32)
        arrs, ptrs, lens = self._truth_table(truths)
        out = np.zeros((len(arrs), 3 * len(ks)), np.float32)
        self.lib.ref_evaluate_holdout(
            ctypes.c_int(len(arrs)), topk.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(topk.shape[1]),
            ks.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(len(ks)), ptrs,
            lens.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p))
        return out

    def loo(self, topk, truths, ks):
        topk = np.ascontiguousarray(topk, np.int32)
        ks = np.ascontiguousarray(ks, np.int32)
        arrs, ptrs, _ = self._truth_table(truths)
        out = np.zeros((len(arrs), 2 * len(ks)), np.float32)
        self.lib.ref_evaluate_loo(
            ctypes.c_int(len(arrs)), topk.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(topk.shape[1]),
            ks.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(len(ks)), ptrs,
            out.ctypes.data_as(ctypes.c_void_p))
        return out
