// On-device triple sampling shared by the fused BPR step (bpr_step.cu) and the multi-GPU router (p2p.cu).
// Semantics of data/generators.py:168-201: positive uniform in the user's CSR row (the intended BPR of :188),
// negative uniform over the non-positives of an id range by rejection against the sorted row (:178-189 for n=1).
// Host mirror: oracle/bpr_oracle.py::sample_triple (same counter RNG, same draw order).
#pragma once
#include "common.cuh"

namespace b200 {

constexpr uint32_t kNegTries = 64;

__device__ __forceinline__ int sample_pos(const int32_t *row, uint32_t deg, uint64_t seed, uint64_t step, uint64_t t) {
    return row[(uint32_t)(((uint64_t)rng_u32(seed, step, t, 0) * deg) >> 32)];
}

// uniform over [n_lo, n_lo + n_cnt) minus the row; false when no non-positive was found in kNegTries draws (a user
// whose positives cover (almost) the whole range: the triple is then skipped, never trained positive-vs-positive)
__device__ __forceinline__ bool sample_neg(const int32_t *row, uint32_t deg, uint32_t n_lo, uint32_t n_cnt,
                                           uint64_t seed, uint64_t step, uint64_t t, int &j) {
    for (uint32_t tries = 0; tries < kNegTries; ++tries) {
        j = (int)(n_lo + (uint32_t)(((uint64_t)rng_u32(seed, step, t, 1 + tries) * (uint64_t)n_cnt) >> 32));
        uint32_t l = 0, r = deg;  // lower_bound in the sorted row
        while (l < r) {
            const uint32_t m = (l + r) >> 1;
            if (row[m] < j) l = m + 1; else r = m;
        }
        if (!(l < deg && row[l] == j)) return true;
    }
    return false;
}

// same over the union [0, cnt0) U [lo1, lo1 + cnt1): the replicated head of the catalogue plus one tail shard
// (csrc/p2p.cu).  cnt0 == 0 makes exactly the draws of sample_neg(row, deg, lo1, cnt1, ...).
__device__ __forceinline__ bool sample_neg2(const int32_t *row, uint32_t deg, uint32_t cnt0, uint32_t lo1, uint32_t cnt1,
                                            uint64_t seed, uint64_t step, uint64_t t, int &j) {
    for (uint32_t tries = 0; tries < kNegTries; ++tries) {
        const uint32_t r = (uint32_t)(((uint64_t)rng_u32(seed, step, t, 1 + tries) * (uint64_t)(cnt0 + cnt1)) >> 32);
        j = (int)(r < cnt0 ? r : lo1 + (r - cnt0));
        uint32_t l = 0, h = deg;
        while (l < h) {
            const uint32_t m = (l + h) >> 1;
            if (row[m] < j) l = m + 1; else h = m;
        }
        if (!(l < deg && row[l] == j)) return true;
    }
    return false;
}

// lane-parallel triple fetch / sampling for the single-table kernels
__device__ __forceinline__ void fetch_triple(const b200rec_bpr_args &a, int64_t t, bool &valid, int &u, int &i,
                                             int &j) {
    u = i = j = 0;
    if (!valid) return;
    u = a.users[t];
    const bool do_pos = (a.pos == nullptr), do_neg = (a.neg == nullptr);
    if (!do_pos) i = a.pos[t];
    if (!do_neg) j = a.neg[t];
    if (do_pos || do_neg) {
        const int64_t lo = a.csr_indptr[u], hi = a.csr_indptr[u + 1];
        const uint32_t deg = (uint32_t)(hi - lo);
        if (deg == 0 && do_pos) {
            valid = false;  // generators.py:186-189: a user without positives emits no triple
        } else {
            const int32_t *row = a.csr_indices + lo;
            if (do_pos) i = sample_pos(row, deg, a.seed, a.step, (uint64_t)t);
            // item-sharded layout: only the rank owning the positive processes the triple,
            // and it draws the negative from its own id range
            const bool sharded = a.item_hi > a.item_lo;
            const uint32_t n_lo = sharded ? (uint32_t)a.item_lo : 0u;
            const uint32_t n_cnt = sharded ? (uint32_t)(a.item_hi - a.item_lo) : (uint32_t)a.num_items;
            if (sharded && (i < a.item_lo || i >= a.item_hi)) valid = false;
            if (do_neg && valid) valid = sample_neg(row, deg, n_lo, n_cnt, a.seed, a.step, (uint64_t)t, j);
        }
        if (a.out_pos) a.out_pos[t] = valid ? i : -1;
        if (a.out_neg) a.out_neg[t] = valid ? j : -1;
    }
}

}  // namespace b200
