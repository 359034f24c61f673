"""B200REC_RERANK=2 (rerank_staged_kernel, csrc/score_tc.cu) must give exactly the default result; own process so that
a fault in this opt-in path cannot take the suite's CUDA context with it (tests/test_gpu_experimental.py; green on a
B200 since the round-1 driver run)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["B200REC_RERANK"] = "2"


def main():
    from recsys_pytorch_b200 import engine, synthetic
    from recsys_pytorch_b200._lib import SCORE_EXACT, SCORE_TC
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(1)
    for (nu, ni, d, k) in ((600, 5000, 128, 10), (513, 20000, 64, 100), (300, 3000, 50, 10), (260, 5000, 256, 100), (100, 1682, 32, 5)):
        U = engine.alloc_table(nu, d, dev, std=0.0); V = engine.alloc_table(ni, d, dev, std=0.0)
        U[:, :d] = torch.from_numpy(rng.standard_normal((nu, d)).astype(np.float32)).to(dev)
        V[:, :d] = torch.from_numpy((rng.standard_normal((ni, d)) * np.exp(rng.standard_normal((ni, 1)) * 0.5)).astype(np.float32)).to(dev)
        mask, _ = synthetic.make_interactions(nu, ni, seed=nu, device=dev, dmax=60)
        users = torch.from_numpy(rng.permutation(nu).astype(np.int32)).to(dev)
        it, st = engine.score_topk(U, V, d, users, mask, k, algo=SCORE_TC)
        ie, se = engine.score_topk(U, V, d, users, mask, k, algo=SCORE_EXACT)
        assert torch.equal(it, ie) and torch.equal(st, se), (nu, ni, d, k, int((it != ie).sum()))
    torch.cuda.synchronize()
    print("RERANK_STAGED_OK")


if __name__ == "__main__":
    main()
