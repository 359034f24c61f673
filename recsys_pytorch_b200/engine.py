"""Thin tensor-level wrappers over the C ABI (include/b200rec.h).

Everything here takes CUDA tensors, passes raw device pointers + the current
torch stream to libb200rec.so, and returns CUDA tensors.  PyTorch is used for
memory and streams only; no torch op computes anything on the hot path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import (B200RecError, BprArgs, F_TMA_GATHER, F_USERS_UNIQUE, SCORE_EXACT, SCORE_TC, SINK_GRAD,
                   SINK_STAGE, SINK_UPDATE, check, current_stream, ptr, require_cuda)

__all__ = ["pointwise_step", "DeviceCSR", "padded_dim", "alloc_table", "mf_forward", "bpr_step", "bpr_apply", "sample_triples",
           "sgd_dense", "adam_dense", "score_topk", "predict_dense", "topk_rows", "holdout_metrics", "loo_metrics",
           "column_means", "spmm_csr", "spmm_plan", "SpmmPlan"]


def padded_dim(d: int) -> int:
    """Row stride (floats): rows are 16-byte multiples so they move as float4 / TMA bulk."""
    return (int(d) + 3) // 4 * 4


def alloc_table(rows: int, d: int, device, std: float = 1.0, generator=None):
    """fp32 [rows, ld] storage, N(0, std) in the first d columns (nn.Embedding's
    default init is N(0,1), models/MF.py:23-24), zeros in the pad columns."""
    ld = padded_dim(d)
    store = torch.zeros((rows, ld), dtype=torch.float32, device=device)
    if std > 0:
        store[:, :d].normal_(0.0, std, generator=generator)
    return store


class DeviceCSR:
    """int64 indptr + int32 sorted column indices on the device (implicit 1.0 values)."""

    def __init__(self, indptr, indices, shape):
        self.indptr = require_cuda(indptr, "indptr", torch.int64)
        self.indices = require_cuda(indices, "indices", torch.int32)
        self.shape = (int(shape[0]), int(shape[1]))

    @classmethod
    def from_scipy(cls, mat, device):
        mat = mat.tocsr()
        if not mat.has_sorted_indices:
            mat = mat.copy()
            mat.sort_indices()
        return cls(torch.from_numpy(np.ascontiguousarray(mat.indptr, np.int64)).to(device),
                   torch.from_numpy(np.ascontiguousarray(mat.indices, np.int32)).to(device), mat.shape)

    @property
    def nnz(self):
        return int(self.indices.numel())


def _i32(t, name):
    return require_cuda(t, name, torch.int32)


def mf_forward(U, V, d, users, items):
    """models/MF.py:38-42."""
    require_cuda(U, "U", torch.float32); require_cuda(V, "V", torch.float32)
    out = torch.empty(users.numel(), dtype=torch.float32, device=U.device)
    check(_lib.lib().b200rec_mf_forward(ptr(U), ptr(V), U.shape[1], d, ptr(_i32(users, "users")),
                                        ptr(_i32(items, "items")), users.numel(), ptr(out), current_stream()))
    return out


def bpr_step(U, V, d, users, pos=None, neg=None, csr: DeviceCSR | None = None, lr=0.0, reg=0.0, sink=SINK_UPDATE,
             flags=0, seed=0, step=0, loss_sum=None, x_out=None, out_pos=None, out_neg=None, stage=None, gU=None,
             gV=None, item_range=None, udelta=None, inv_batch=0.0):
    """models/MF.py:63-68 as one fused kernel (see b200rec_bpr_step)."""
    require_cuda(U, "U", torch.float32); require_cuda(V, "V", torch.float32)
    a = BprArgs()
    a.U, a.V, a.ld, a.d = ptr(U), ptr(V), U.shape[1], d
    a.num_users, a.num_items = U.shape[0], V.shape[0]
    a.users = ptr(_i32(users, "users"))
    a.pos = ptr(_i32(pos, "pos")) if pos is not None else None
    a.neg = ptr(_i32(neg, "neg")) if neg is not None else None
    a.B = users.numel()
    if csr is not None:
        a.csr_indptr, a.csr_indices = ptr(csr.indptr), ptr(csr.indices)
    a.seed, a.step = int(seed) & (2**64 - 1), int(step) & (2**64 - 1)
    a.out_pos = ptr(_i32(out_pos, "out_pos")) if out_pos is not None else None
    a.out_neg = ptr(_i32(out_neg, "out_neg")) if out_neg is not None else None
    a.lr, a.reg, a.sink, a.flags = float(lr), float(reg), int(sink), int(flags)
    a.stage = ptr(require_cuda(stage, "stage", torch.float32)) if stage is not None else None
    a.gU = ptr(require_cuda(gU, "gU", torch.float32)) if gU is not None else None
    a.gV = ptr(require_cuda(gV, "gV")) if gV is not None else None        # fp32, or bf16 with F_ITEM_DELTA_BF16
    a.loss_sum = ptr(require_cuda(loss_sum, "loss_sum", torch.float64)) if loss_sum is not None else None
    a.x_out = ptr(require_cuda(x_out, "x_out", torch.float32)) if x_out is not None else None
    if item_range is not None:
        a.item_lo, a.item_hi = int(item_range[0]), int(item_range[1])
    a.udelta = ptr(require_cuda(udelta, "udelta", torch.float32)) if udelta is not None else None
    a.inv_batch = float(inv_batch)
    check(_lib.lib().b200rec_bpr_step(C.byref(a), current_stream()))


def pointwise_step(U, V, d, users, items, ratings, loss_func="ce", lr=0.0, reg=0.0, sink=SINK_UPDATE, gU=None, gV=None,
                   loss_sum=None, inv_batch=0.0):
    """models/MF.py:63-68 in pointwise mode (MF.py:101-102) as one fused kernel (see b200rec_pointwise_step)."""
    require_cuda(U, "U", torch.float32); require_cuda(V, "V", torch.float32)
    ratings = require_cuda(ratings, "ratings", torch.float32)
    check(_lib.lib().b200rec_pointwise_step(
        ptr(U), ptr(V), U.shape[1], d, ptr(_i32(users, "users")), ptr(_i32(items, "items")), ptr(ratings),
        users.numel(), 1 if loss_func == "mse" else 0, float(lr), float(reg), int(sink),
        ptr(require_cuda(gU, "gU", torch.float32)) if gU is not None else None,
        ptr(require_cuda(gV, "gV", torch.float32)) if gV is not None else None,
        ptr(require_cuda(loss_sum, "loss_sum", torch.float64)) if loss_sum is not None else None,
        float(inv_batch), current_stream()))


def bpr_apply(U, V, users, pos, neg, stage):
    check(_lib.lib().b200rec_bpr_apply(ptr(U), ptr(V), U.shape[1], ptr(_i32(users, "users")), ptr(_i32(pos, "pos")),
                                       ptr(_i32(neg, "neg")), users.numel(), ptr(stage), current_stream()))


def rows_add(W, ids, delta, scale=1.0):
    """W[ids[t]] += scale * delta[t] (vector atomics)."""
    check(_lib.lib().b200rec_rows_add(ptr(W), W.shape[1], ptr(_i32(ids, "ids")), ids.numel(), ptr(delta),
                                      delta.shape[-1], float(scale), current_stream()))


def sample_triples(users, csr: DeviceCSR, seed, step):
    """data/generators.py:168-201 on the device; returns (pos, neg) int32."""
    users = _i32(users, "users")
    pos = torch.empty_like(users); neg = torch.empty_like(users)
    check(_lib.lib().b200rec_sample_triples(ptr(users), users.numel(), ptr(csr.indptr), ptr(csr.indices),
                                            csr.shape[1], int(seed) & (2**64 - 1), int(step) & (2**64 - 1),
                                            ptr(pos), ptr(neg), current_stream()))
    return pos, neg


def add_bf16(param, delta_bf16):
    """param += float(delta) for a torch.bfloat16 delta of the same shape (user-sharded exchange buffer)."""
    check(_lib.lib().b200rec_add_bf16(ptr(param), ptr(delta_bf16), param.numel(), current_stream()))


def snap_apply(W, snapshot, d_sum, scale=1.0):
    """W = snapshot + scale * d_sum (dense, same shape)."""
    check(_lib.lib().b200rec_snap_apply(ptr(W), ptr(snapshot), ptr(d_sum), float(scale), W.numel(), current_stream()))


def add_clear(W, delta):
    """W += delta; delta = 0 (dense, same shape)."""
    check(_lib.lib().b200rec_add_clear(ptr(W), ptr(delta), W.numel(), current_stream()))


def delta_diff(W, snapshot, d_wire, d_own):
    """d_wire = d_own = W - snapshot (dense, same shape)."""
    check(_lib.lib().b200rec_delta_diff(ptr(W), ptr(snapshot), ptr(d_wire), ptr(d_own), W.numel(), current_stream()))


def delta_apply(W, d_sum, d_own):
    """W += d_sum - d_own."""
    check(_lib.lib().b200rec_delta_apply(ptr(W), ptr(d_sum), ptr(d_own), W.numel(), current_stream()))


def l2_persist(table, hit_ratio=1.0):
    """Keep `table` resident in L2 for kernels launched on the current stream (None clears the window)."""
    if table is None:
        check(_lib.lib().b200rec_l2_persist(None, 0, 0.0, current_stream()))
    else:
        check(_lib.lib().b200rec_l2_persist(ptr(table), table.numel() * table.element_size(), float(hit_ratio),
                                            current_stream()))


def sgd_dense(param, grad, lr):
    check(_lib.lib().b200rec_sgd_dense(ptr(param), ptr(grad), param.numel(), float(lr), current_stream()))


def adam_dense(param, grad, exp_avg, exp_avg_sq, step, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8):
    check(_lib.lib().b200rec_adam_dense(ptr(param), ptr(grad), ptr(exp_avg), ptr(exp_avg_sq), param.numel(),
                                        float(lr), float(beta1), float(beta2), float(eps), int(step),
                                        current_stream()))


def adam_rows(W, grad, exp_avg, exp_avg_sq, stamp, ids, step, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.SparseAdam on the rows ids[] (see b200rec_adam_rows); zeroes the used grad rows."""
    check(_lib.lib().b200rec_adam_rows(ptr(W), ptr(grad), ptr(exp_avg), ptr(exp_avg_sq), ptr(_i32(stamp, "stamp")),
                                       W.shape[1], ptr(_i32(ids, "ids")), ids.numel(), float(lr), float(beta1),
                                       float(beta2), float(eps), int(step), current_stream()))


def score_topk(U, V, d, users, mask: DeviceCSR | None, k, algo=SCORE_EXACT, want_scores=True):
    """models/MF.py:109-132 + func.h:12-31 fused: (idx int32 [n,k], score fp32 [n,k])."""
    require_cuda(U, "U", torch.float32); require_cuda(V, "V", torch.float32)
    users = _i32(users, "users")
    n, ni = users.numel(), V.shape[0]
    idx = torch.empty((n, k), dtype=torch.int32, device=U.device)
    sc = torch.empty((n, k), dtype=torch.float32, device=U.device) if want_scores else None
    ws_bytes = int(_lib.lib().b200rec_score_topk_workspace(n, ni, d, k, algo))
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=U.device)
    check(_lib.lib().b200rec_score_topk(ptr(U), ptr(V), U.shape[1], d, ptr(users), n, ni,
                                        ptr(mask.indptr) if mask is not None else None,
                                        ptr(mask.indices) if mask is not None else None, k, ptr(idx), ptr(sc),
                                        ptr(ws), ws_bytes, algo, current_stream()))
    return idx, sc


def debug_tc_scores(U, V, d, users):
    """Test hook: raw fp16 tensor-core scores of the candidate pass.  Returns (scores fp32 [n, I] in
    ORIGINAL item order and unscaled, scale_u, scale_v, perm) - the kernel itself works on items sorted by
    descending norm and on power-of-two rescaled tables."""
    users = _i32(users, "users")
    n, ni = users.numel(), V.shape[0]
    import os
    tile = 256 if (d <= 128 and os.environ.get("B200REC_TC_KERNEL", "1") != "0") else 128   # as score_tc.cu::tc_layout
    rp, ip = (n + 255) // 256 * 256, (ni + tile - 1) // tile * tile
    out = torch.zeros((rp, ip), dtype=torch.float32, device=U.device)
    ws_bytes = int(_lib.lib().b200rec_score_topk_workspace(n, ni, d, 1, SCORE_TC))
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=U.device)
    check(_lib.lib().b200rec_debug_tc_scores(ptr(U), ptr(V), U.shape[1], d, ptr(users), n, ni, ptr(out), ptr(ws),
                                             ws_bytes, current_stream()))
    off_perm, off_scales = C.c_int64(0), C.c_int64(0)
    check(_lib.lib().b200rec_debug_tc_layout(n, ni, d, C.byref(off_perm), C.byref(off_scales)))
    base = (-ws.data_ptr()) % 1024
    perm = ws[base + off_perm.value: base + off_perm.value + 4 * ni].view(torch.int32).long()
    scales = ws[base + off_scales.value: base + off_scales.value + 8].view(torch.float32)
    sv, su = float(scales[0]), float(scales[1])
    res = torch.empty((n, ni), dtype=torch.float32, device=U.device)
    res[:, perm] = out[:n, :ni] / (su * sv)
    return res, su, sv, perm


def predict_dense(U, V, d, users, mask: DeviceCSR | None):
    """models/MF.py:109-130 for a chunk of users: fp32 [n, I] with -inf at mask nonzeros."""
    users = _i32(users, "users")
    out = torch.empty((users.numel(), V.shape[0]), dtype=torch.float32, device=U.device)
    check(_lib.lib().b200rec_predict_dense(ptr(U), ptr(V), U.shape[1], d, ptr(users), users.numel(), V.shape[0],
                                           ptr(mask.indptr) if mask is not None else None,
                                           ptr(mask.indices) if mask is not None else None, ptr(out),
                                           current_stream()))
    return out


def topk_rows(scores, k):
    scores = require_cuda(scores, "scores", torch.float32)
    out = torch.empty((scores.shape[0], k), dtype=torch.int32, device=scores.device)
    check(_lib.lib().b200rec_topk_rows(ptr(scores), scores.shape[1], scores.shape[0], scores.shape[1], k, ptr(out),
                                       current_stream()))
    return out


def _metrics(fn, width, topk, truth: DeviceCSR, ks, row_ids):
    topk = _i32(topk, "topk")
    ks_arr = (C.c_int * len(ks))(*[int(k) for k in ks])
    out = torch.empty((topk.shape[0], width * len(ks)), dtype=torch.float32, device=topk.device)
    check(fn(ptr(topk), topk.shape[0], topk.shape[1], ptr(_i32(row_ids, "row_ids")) if row_ids is not None else None,
             ptr(truth.indptr), ptr(truth.indices), ks_arr, len(ks), ptr(out), current_stream()))
    return out


def holdout_metrics(topk, truth: DeviceCSR, ks, row_ids=None):
    """holdout.h:20-103 on the device: fp32 [n, 3*len(ks)] = [Prec.., Recall.., NDCG..]."""
    return _metrics(_lib.lib().b200rec_holdout_metrics, 3, topk, truth, ks, row_ids)


def loo_metrics(topk, truth: DeviceCSR, ks, row_ids=None):
    """loo.h:20-85 on the device: fp32 [n, 2*len(ks)] = [HR.., NDCG..]."""
    return _metrics(_lib.lib().b200rec_loo_metrics, 2, topk, truth, ks, row_ids)


def column_means(mat):
    mat = require_cuda(mat, "mat", torch.float32)
    out = np.zeros(mat.shape[1], np.float64)
    check(_lib.lib().b200rec_column_means(ptr(mat), mat.shape[0], mat.shape[1], out.ctypes.data, current_stream()))
    return out


class SpmmPlan:
    """Segments of the long rows of a CSR matrix (see b200rec_spmm_csr_split); depends on indptr only."""

    def __init__(self, indptr, seg_len=256):
        dev = indptr.device
        deg = indptr[1:] - indptr[:-1]
        long_rows = torch.nonzero(deg > seg_len).flatten()
        self.seg_len = int(seg_len)
        self.n_long = int(long_rows.numel())
        self.n_seg = 0
        self._partial = None
        if self.n_long == 0:
            return
        nseg = (deg[long_rows] + seg_len - 1) // seg_len
        seg_ptr = torch.zeros(self.n_long + 1, dtype=torch.int64, device=dev)
        seg_ptr[1:] = torch.cumsum(nseg, 0)
        self.n_seg = int(seg_ptr[-1].item())
        owner = torch.repeat_interleave(torch.arange(self.n_long, device=dev), nseg)
        idx = torch.arange(self.n_seg, device=dev) - seg_ptr[owner]
        row = long_rows[owner]
        self.seg_begin = (indptr[row] + idx * seg_len).contiguous()
        self.seg_end = torch.minimum(self.seg_begin + seg_len, indptr[row + 1]).contiguous()
        self.long_rows = long_rows.to(torch.int32).contiguous()
        self.long_seg_ptr = seg_ptr.to(torch.int32).contiguous()

    def partial(self, d, device):
        need = self.n_seg * ((d + 3) // 4 * 4)
        if self._partial is None or self._partial.numel() < need:
            self._partial = torch.empty(max(need, 1), dtype=torch.float32, device=device)
        return self._partial


def spmm_plan(indptr, seg_len=256):
    return SpmmPlan(indptr, seg_len)


def spmm_csr(indptr, indices, values, X, d, Y=None, acc=None, acc_scale=1.0, acc_init=False, plan=None):
    """models/LightGCN.py:196: Y = A X; optionally acc += acc_scale * A X, or with
    acc_init acc = acc_scale * (X + A X) (start of the running layer mean, :198-200).
    `plan` (spmm_plan(indptr)) spreads rows longer than plan.seg_len over many sub-groups."""
    n_rows = indptr.numel() - 1
    if plan is not None and plan.n_seg > 0:
        check(_lib.lib().b200rec_spmm_csr_split(ptr(indptr), ptr(indices), ptr(values), n_rows, ptr(X), X.shape[1], d,
                                                ptr(Y), Y.shape[1] if Y is not None else 0, ptr(acc),
                                                acc.shape[1] if acc is not None else 0, float(acc_scale),
                                                int(bool(acc_init)), plan.seg_len, ptr(plan.seg_begin), ptr(plan.seg_end),
                                                plan.n_seg, ptr(plan.long_rows), ptr(plan.long_seg_ptr), plan.n_long,
                                                ptr(plan.partial(d, X.device)), current_stream()))
        return Y
    check(_lib.lib().b200rec_spmm_csr(ptr(indptr), ptr(indices), ptr(values), n_rows, ptr(X), X.shape[1], d,
                                      ptr(Y), Y.shape[1] if Y is not None else 0, ptr(acc),
                                      acc.shape[1] if acc is not None else 0, float(acc_scale), int(bool(acc_init)),
                                      current_stream()))
    return Y
