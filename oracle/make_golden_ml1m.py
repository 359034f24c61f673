"""Generate tests/golden/ml1m_dataset.npz by RUNNING THE REFERENCE's UIRTDataset (data/dataset.py:12-250 +
data/preprocess.py:9-90) on its own datasets/ml-1m/ratings.dat in this container.  TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden_ml1m          # from the repo root

Second real dataset for the ingest row: 6040 users x 3706 items, 1 000 209 ratings, '::' separator, many
equal timestamps inside a user (the time-based split depends on how pandas orders ties).  Two configurations:
  A  holdout / weak, split_random=True after set_random_seed(2020)          (config.py:7-23 defaults)
  B  leave_one_out, leave_k=2, split_random=False (time split), min_user_per_item=20 (the item filter bites)
Per configuration the train / valid / test matrices the reference built (CSR, indices int16: 3706 items fit).
"""
from __future__ import annotations

import os
import shutil
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
CONFIGS = {
    "A": dict(min_item_per_user=10, min_user_per_item=1, protocol="holdout", valid_ratio=0.1, test_ratio=0.2, leave_k=1,
              split_random=True),
    "B": dict(min_item_per_user=10, min_user_per_item=20, protocol="leave_one_out", valid_ratio=0.1, test_ratio=0.2,
              leave_k=2, split_random=False),
}


def main():
    ref = ref_harness.load()
    out = {}
    for tag, kw in CONFIGS.items():
        dst = os.path.join(ref_harness.WORK, "ml-1m-" + tag)               # own cache dir per configuration
        os.makedirs(dst, exist_ok=True)
        if os.path.isdir(os.path.join(dst, "cache")):
            shutil.rmtree(os.path.join(dst, "cache"))                       # dataset.py:211-227 would reload a stale split
        if not os.path.exists(os.path.join(dst, "ratings.dat")):
            shutil.copy(os.path.join(ref_harness.REF, "datasets", "ml-1m", "ratings.dat"), dst)
        ref.set_random_seed(2020)                                           # utils/general.py:31
        ds = ref.UIRTDataset(data_path=os.path.join(dst, "ratings.dat"), dataname="ml-1m", separator="::",
                             binarize_threshold=0.0, implicit=True, generalization="weak", holdout_users=600, **kw)
        out[tag + "_num_users"], out[tag + "_num_items"] = np.int64(ds.num_users), np.int64(ds.num_items)
        for name in ("train_data", "valid_target", "test_target"):
            m = getattr(ds, name).tocsr().copy()
            m.sum_duplicates(); m.sort_indices()
            assert m.shape[1] < 32768 and (m.data == 1).all()
            out["%s_%s_indptr" % (tag, name)] = m.indptr.astype(np.int32)
            out["%s_%s_indices" % (tag, name)] = m.indices.astype(np.int16)
        print(tag, ds.num_users, ds.num_items, ds.train_data.nnz, ds.valid_target.nnz, ds.test_target.nnz)
    np.savez_compressed(os.path.join(OUT, "ml1m_dataset.npz"), **out)
    print("wrote", os.path.join(OUT, "ml1m_dataset.npz"), os.path.getsize(os.path.join(OUT, "ml1m_dataset.npz")), "bytes")


if __name__ == "__main__":
    main()
