/* Plain-C restatement of the reference's native evaluation layer and of the
 * chunked scoring loop.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py):
 * linked by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg,
 * never by the product.
 *
 * Follows (paths relative to /root/reference):
 *   evaluation/backend/cython/include/func.h:12-31     top-k by score desc
 *   evaluation/backend/cython/include/holdout.h:20-103 Prec/Recall/NDCG
 *   evaluation/backend/cython/include/loo.h:20-85      HR/NDCG leave-one-out
 *   models/MF.py:109-112,130                           U[users] @ V^T, -inf mask
 * Tie order in func.h is unspecified (std::partial_sort_copy); this file and the
 * CUDA engine both use (score desc, item id asc).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* binary min-heap on (score asc, id desc) == worst-of-the-best at the root */
static int worse(float sa, int ia, float sb, int ib) {
    return (sa < sb) || (sa == sb && ia > ib);
}

static void sift_down(float *hs, int *hi, int n, int p) {
    for (;;) {
        int l = 2 * p + 1, r = l + 1, m = p;
        if (l < n && worse(hs[l], hi[l], hs[m], hi[m])) m = l;
        if (r < n && worse(hs[r], hi[r], hs[m], hi[m])) m = r;
        if (m == p) return;
        float ts = hs[p]; hs[p] = hs[m]; hs[m] = ts;
        int ti = hi[p]; hi[p] = hi[m]; hi[m] = ti;
        p = m;
    }
}

/* func.h:12-20 c_top_k_index: indices of the top_k largest ratings, descending */
void oracle_top_k_index(const float *ratings, int rating_len, int top_k, int *result) {
    float *hs = (float *)malloc(sizeof(float) * (size_t)top_k);
    int *hi = (int *)malloc(sizeof(int) * (size_t)top_k);
    int n = 0;
    for (int i = 0; i < rating_len; ++i) {
        float s = ratings[i];
        if (n < top_k) {
            hs[n] = s; hi[n] = i; ++n;
            if (n == top_k)
                for (int p = n / 2 - 1; p >= 0; --p) sift_down(hs, hi, n, p);
        } else if (worse(hs[0], hi[0], s, i)) {
            hs[0] = s; hi[0] = i;
            sift_down(hs, hi, n, 0);
        }
    }
    if (n < top_k)
        for (int p = n / 2 - 1; p >= 0; --p) sift_down(hs, hi, n, p);
    /* pop ascending-worst into the tail -> result is best-first */
    for (int m = n; m > 0; --m) {
        result[m - 1] = hi[0];
        hs[0] = hs[m - 1]; hi[0] = hi[m - 1];
        sift_down(hs, hi, m - 1, 0);
    }
    free(hs); free(hi);
}

/* func.h:22-31 c_top_k_array_index */
void oracle_top_k_array_index(const float *scores, int columns_num, int rows_num,
                              int max_k, int *rankings) {
    for (int i = 0; i < rows_num; ++i)
        oracle_top_k_index(scores + (size_t)columns_num * i, columns_num, max_k,
                           rankings + (size_t)max_k * i);
}

static int in_truth(const int *truth, int n, int v) {
    for (int i = 0; i < n; ++i) if (truth[i] == v) return 1;
    return 0;
}

/* holdout.h:20-103 evaluate_holdout.  `float` accumulators, double terms. */
void oracle_evaluate_holdout(int users_num, const int *rankings, int max_k,
                             const int *Ks, int K_len, int **ground_truths,
                             const int *ground_truths_num, float *results) {
    for (int uid = 0; uid < users_num; ++uid) {
        const int *cur = rankings + (size_t)uid * max_k;
        const int *truth = ground_truths[uid];
        int truth_len = ground_truths_num[uid];
        float *res = results + (size_t)uid * 3 * K_len;
        float hits = 0, iDCG = 0, DCG = 0;
        for (int i = 0; i < max_k; ++i) {
            if (in_truth(truth, truth_len, cur[i])) {
                hits += 1;
                DCG += 1.0 / log2(i + 2);
            }
            if (i < truth_len) iDCG += 1.0 / log2(i + 2);
            for (int j = 0; j < K_len; ++j)
                if (Ks[j] == i + 1) {
                    res[j] = hits / (float)Ks[j];
                    res[K_len + j] = hits / truth_len;
                    res[2 * K_len + j] = DCG / iDCG;
                }
        }
    }
}

/* loo.h:20-85 evaluate_loo */
void oracle_evaluate_loo(int users_num, const int *rankings, int max_k,
                         const int *Ks, int K_len, int **ground_truths,
                         float *results) {
    for (int uid = 0; uid < users_num; ++uid) {
        const int *cur = rankings + (size_t)uid * max_k;
        int truth = ground_truths[uid][0];
        float *res = results + (size_t)uid * 2 * K_len;
        int hit_at_k = max_k + 1;
        for (int i = 0; i < max_k; ++i)
            if (cur[i] == truth) { hit_at_k = i + 1; break; }
        for (int j = 0; j < K_len; ++j) {
            if (Ks[j] >= hit_at_k) {
                res[j] = 1.0;
                res[K_len + j] = 1 / log2(hit_at_k + 1);
            } else {
                res[j] = 0.0;
                res[K_len + j] = 0.0;
            }
        }
    }
}

/* models/MF.py:109-112 + :130 + func.h, for one chunk of users, without the
 * dense float64 [U,I] container: fp32 dot in k order, -inf on the
 * user's train positives, top-k.  `scratch` holds num_items floats. */
void oracle_score_topk_chunk(const float *U, const float *V, int d, int ld,
                             const int *users, int n_users, int num_items,
                             const int64_t *mask_indptr, const int *mask_indices,
                             int k, int *out_idx, float *out_score, float *scratch) {
    for (int r = 0; r < n_users; ++r) {
        const float *u = U + (size_t)users[r] * ld;
        for (int it = 0; it < num_items; ++it) {
            const float *v = V + (size_t)it * ld;
            float acc = 0.f;
            for (int c = 0; c < d; ++c) acc = fmaf(u[c], v[c], acc);
            scratch[it] = acc;
        }
        if (mask_indptr)
            for (int64_t p = mask_indptr[users[r]]; p < mask_indptr[users[r] + 1]; ++p)
                scratch[mask_indices[p]] = -INFINITY;
        oracle_top_k_index(scratch, num_items, k, out_idx + (size_t)r * k);
        for (int c = 0; c < k; ++c)
            out_score[(size_t)r * k + c] = scratch[out_idx[(size_t)r * k + c]];
    }
}
