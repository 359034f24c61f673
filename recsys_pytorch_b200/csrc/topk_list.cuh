// Warp-cooperative sorted top-K list (descending 64-bit keys, see make_key()).
// Semantics of evaluation/backend/cython/include/func.h:12-20 (partial_sort_copy
// by score desc) with the tie order fixed to (score desc, id asc).
#pragma once
#include "common.cuh"

namespace b200 {

// insert key c (known to beat rk[k-1]) into the descending list rk[0..k)
__device__ __forceinline__ void topk_list_insert(uint64_t *rk, int k, uint64_t c, int lane) {
    int pos = 0;
    for (int base = 0; base < k; base += 32) {
        const int p = base + lane;
        const unsigned b = __ballot_sync(0xffffffffu, (p < k) && (rk[p] > c));
        pos += __popc(b);
        if (b != 0xffffffffu) break;
    }
    for (int base = ((k - 1) >> 5) << 5; base >= 0; base -= 32) {  // shift the tail down, top segment first
        const int p = base + lane;
        uint64_t nv = 0;
        bool w = false;
        if (p < k && p > pos) { nv = rk[p - 1]; w = true; }
        else if (p == pos) { nv = c; w = true; }
        __syncwarp();
        if (w) rk[p] = nv;
        __syncwarp();
        if (base <= pos) break;
    }
}

// every lane offers one candidate (valid when ok); survivors are inserted in lane order
__device__ __forceinline__ void topk_list_offer(uint64_t *rk, int k, uint64_t key, bool ok, int lane) {
    unsigned pend = __ballot_sync(0xffffffffu, ok && key > rk[k - 1]);
    while (pend) {
        const int src = __ffs(pend) - 1;
        pend &= pend - 1;
        const uint64_t c = __shfl_sync(0xffffffffu, key, src);
        if (c > rk[k - 1]) topk_list_insert(rk, k, c, lane);
        __syncwarp();
    }
}

}  // namespace b200
