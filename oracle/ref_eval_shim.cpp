// extern "C" doorway onto the UNMODIFIED reference headers (C++ linkage,
// header-only).  Compiled by oracle/Makefile with
//   -I/root/reference/evaluation/backend/cython/include
// into oracle/_ref/libref_eval.so.  No reference source is copied: the three
// #includes below resolve into /root/reference at build time.
// TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
#include "func.h"      // c_top_k_array_index   (func.h:22)
#include "holdout.h"   // evaluate_holdout      (holdout.h:20)
#include "loo.h"       // evaluate_loo          (loo.h:20)

extern "C" {
void ref_top_k_array_index(float *scores, int columns_num, int rows_num, int max_k, int *rankings) {
    c_top_k_array_index(scores, columns_num, rows_num, max_k, rankings);
}
void ref_evaluate_holdout(int users_num, int *rankings, int max_k, int *Ks, int K_len,
                          int **ground_truths, int *ground_truths_num, float *results) {
    evaluate_holdout(users_num, rankings, max_k, Ks, K_len, ground_truths, ground_truths_num, results);
}
void ref_evaluate_loo(int users_num, int *rankings, int max_k, int *Ks, int K_len,
                      int **ground_truths, float *results) {
    evaluate_loo(users_num, rankings, max_k, Ks, K_len, ground_truths, results);
}
}
