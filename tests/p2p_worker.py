"""2-process / 2-GPU check of the P2P-fused layout with REAL CUDA-IPC peer mappings (launched by tests/test_p2p.py via
torch.distributed.run).  Every rank builds the same seeded global problem, keeps its shard, runs sampled steps, and
rank 0 replays the union of the triples with the numpy oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from oracle import bpr_oracle as O
    from recsys_pytorch_b200 import engine
    from recsys_pytorch_b200.p2p import P2PShardedBPR, balanced_item_bounds, uniform_bounds
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    nu, ni, d = 4096, 3000, 128
    rng = np.random.default_rng(0)
    U0 = (rng.standard_normal((nu, d)) * 0.3).astype(np.float32)
    V0 = (rng.standard_normal((ni, d)) * 0.3).astype(np.float32)
    pop = 1.0 / rng.permutation(np.arange(1, ni + 1)); pop /= pop.sum()
    rows = [np.unique(rng.choice(ni, size=rng.integers(2, 20), p=pop)).astype(np.int32) for _ in range(nu)]
    indptr = np.zeros(nu + 1, np.int64); indptr[1:] = np.cumsum([len(r) for r in rows])
    indices = np.concatenate(rows)
    head = 256                                             # replicated head rows [0, 256) + range-sharded tail
    ib = balanced_item_bounds(torch.from_numpy(np.bincount(indices, minlength=ni)), world, head)
    ub = uniform_bounds(nu, world)
    lo, hi = ub[rank], ub[rank + 1]
    csr = engine.DeviceCSR(torch.from_numpy(indptr[lo:hi + 1] - indptr[lo]).to(dev),
                           torch.from_numpy(indices[indptr[lo]:indptr[hi]]).to(dev), (hi - lo, ni))
    m = P2PShardedBPR(nu, ni, d, csr, rank, world, dev, ib, ub, lr=8.0, reg=0.0, init_std=0.0, seed=5, max_batch=hi - lo,
                      head=head)
    m.U[:, :d] = torch.from_numpy(U0[lo:hi]).to(dev)
    m.V[:, :d] = torch.from_numpy(V0[ib[rank]:ib[rank + 1]]).to(dev)
    m.Vh[:, :d] = torch.from_numpy(V0[:head]).to(dev)
    m.connect()
    Uc, Vc = U0, V0
    prng = np.random.default_rng(100 + rank)
    loss = torch.zeros(1, dtype=torch.float64, device=dev)
    for step in (1, 2, 3):
        B = (hi - lo) - 5
        ul = torch.from_numpy(prng.permutation(hi - lo)[:B].astype(np.int32)).to(dev)
        dp, dn = torch.empty_like(ul), torch.empty_like(ul)
        m.route(ul, step, dbg_pos=dp, dbg_neg=dn)
        m.sync_head()
        m.compute(B * world, loss_sum=loss)
        m.sync_head()                                      # apply this step's head delta so the tables can be compared
        m.barrier()
        torch.cuda.synchronize()
        mine = (ul.cpu().numpy() + lo, dp.cpu().numpy(), dn.cpu().numpy(), m.U.cpu().numpy()[:, :d], m.V.cpu().numpy()[:, :d],
                int(m.n_processed.item()), m.Vh.cpu().numpy()[:, :d])
        allm = [None] * world
        dist.all_gather_object(allm, mine)
        if rank == 0:
            gu, gi, gj = (np.concatenate([a[k] for a in allm]) for k in range(3))
            assert sum(a[5] for a in allm) == len(gu)
            Ur, Vr, _ = O.sgd_step(Uc, Vc, gu, gi, gj, 8.0, 0.0)
            U = np.concatenate([a[3] for a in allm]); V = np.concatenate([allm[0][6]] + [a[4] for a in allm])
            assert all(np.array_equal(a[6], allm[0][6]) for a in allm)          # head replicas identical
            stepsz = max(np.abs(Ur - Uc).max(), np.abs(Vr - Vc).max())
            dev_ = max(np.abs(U - Ur).max(), np.abs(V - Vr).max())
            assert stepsz > 2e-3 and dev_ < 0.05 * stepsz, (step, stepsz, dev_)
            Uc, Vc = U, V
        bc = [Uc, Vc]
        dist.broadcast_object_list(bc, src=0)
        Uc, Vc = bc
    full = m.gather_items()
    assert np.array_equal(full.cpu().numpy()[:, :d], Vc)
    # fixed-triple mode across shards: user on this rank, negative anywhere
    items = np.random.default_rng(7).permutation(ni)
    B = min(512, ni // (2 * world), hi - lo)
    ul = torch.from_numpy(prng.permutation(hi - lo)[:B].astype(np.int32)).to(dev)
    pi = items[rank * 2 * B: rank * 2 * B + B].astype(np.int32); pj = items[rank * 2 * B + B: (rank + 1) * 2 * B].astype(np.int32)
    m.lr, m.reg = 0.9, 0.01
    m.route(ul, 9, pos=torch.from_numpy(pi).to(dev), neg=torch.from_numpy(pj).to(dev))
    m.sync_head(); m.compute(B * world); m.sync_head(); m.barrier()
    torch.cuda.synchronize()
    allm = [None] * world
    dist.all_gather_object(allm, (ul.cpu().numpy() + lo, pi, pj, m.U.cpu().numpy()[:, :d], m.V.cpu().numpy()[:, :d],
                                  m.Vh.cpu().numpy()[:, :d]))
    if rank == 0:
        gu, gi, gj = (np.concatenate([a[k] for a in allm]) for k in range(3))
        Ur, Vr, _ = O.sgd_step(Uc, Vc, gu, gi, gj, 0.9, 0.01)
        np.testing.assert_allclose(np.concatenate([a[3] for a in allm]), Ur, rtol=2e-5, atol=2e-6)
        np.testing.assert_allclose(np.concatenate([allm[0][5]] + [a[4] for a in allm]), Vr, rtol=2e-5, atol=2e-6)
        print("P2P_WORKER_OK")
    dist.barrier()
    m.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
