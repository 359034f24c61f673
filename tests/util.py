"""ctypes view of oracle/liboracle.so (the plain-C restatement) for the tests."""
import ctypes as C

import numpy as np


class OracleC:
    def __init__(self, path):
        self.lib = C.CDLL(path)

    def topk(self, scores, k):
        scores = np.ascontiguousarray(scores, np.float32)
        out = np.zeros((scores.shape[0], k), np.int32)
        self.lib.oracle_top_k_array_index(C.c_void_p(scores.ctypes.data), C.c_int(scores.shape[1]),
                                          C.c_int(scores.shape[0]), C.c_int(k), C.c_void_p(out.ctypes.data))
        return out

    @staticmethod
    def _truth(truths):
        arrs = [np.ascontiguousarray(t, np.int32) for t in truths]
        ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
        lens = np.array([len(a) for a in arrs], np.int32)
        return arrs, ptrs, lens

    def holdout(self, topk, truths, ks):
        topk = np.ascontiguousarray(topk, np.int32); ks = np.ascontiguousarray(ks, np.int32)
        arrs, ptrs, lens = self._truth(truths)
        out = np.zeros((len(arrs), 3 * len(ks)), np.float32)
        self.lib.oracle_evaluate_holdout(C.c_int(len(arrs)), C.c_void_p(topk.ctypes.data), C.c_int(topk.shape[1]),
                                         C.c_void_p(ks.ctypes.data), C.c_int(len(ks)), ptrs,
                                         C.c_void_p(lens.ctypes.data), C.c_void_p(out.ctypes.data))
        return out

    def loo(self, topk, truths, ks):
        topk = np.ascontiguousarray(topk, np.int32); ks = np.ascontiguousarray(ks, np.int32)
        arrs, ptrs, _ = self._truth(truths)
        out = np.zeros((len(arrs), 2 * len(ks)), np.float32)
        self.lib.oracle_evaluate_loo(C.c_int(len(arrs)), C.c_void_p(topk.ctypes.data), C.c_int(topk.shape[1]),
                                     C.c_void_p(ks.ctypes.data), C.c_int(len(ks)), ptrs, C.c_void_p(out.ctypes.data))
        return out

    def score_topk(self, U, V, d, users, num_items, indptr, indices, k):
        U = np.ascontiguousarray(U, np.float32); V = np.ascontiguousarray(V, np.float32)
        users = np.ascontiguousarray(users, np.int32)
        idx = np.zeros((len(users), k), np.int32); sc = np.zeros((len(users), k), np.float32)
        scratch = np.zeros(num_items, np.float32)
        ip = np.ascontiguousarray(indptr, np.int64) if indptr is not None else None
        ix = np.ascontiguousarray(indices, np.int32) if indices is not None else None
        self.lib.oracle_score_topk_chunk(
            C.c_void_p(U.ctypes.data), C.c_void_p(V.ctypes.data), C.c_int(d), C.c_int(U.shape[1]),
            C.c_void_p(users.ctypes.data), C.c_int(len(users)), C.c_int(num_items),
            C.c_void_p(ip.ctypes.data) if ip is not None else None,
            C.c_void_p(ix.ctypes.data) if ix is not None else None,
            C.c_int(k), C.c_void_p(idx.ctypes.data), C.c_void_p(sc.ctypes.data), C.c_void_p(scratch.ctypes.data))
        return idx, sc


def truths_from_csr(indptr, indices):
    return [indices[indptr[u]:indptr[u + 1]] for u in range(len(indptr) - 1)]
