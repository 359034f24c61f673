"""torchrun worker for tests/test_dist.py::test_torchrun_two_gpus_replicas_stay_identical."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from recsys_pytorch_b200 import synthetic
    from recsys_pytorch_b200.dist import ItemShardedBPR, UserShardedBPR, shard_range
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    nu, ni, d, B = 20000, 8000, 128, 8192
    train, _ = synthetic.make_interactions(nu, ni, seed=1, device=dev)
    tr = ItemShardedBPR(nu, ni, d, train, rank, world, dev, lr=1.0, reg=0.001, init_std=0.1, seed=3)
    g = torch.Generator(device=dev); g.manual_seed(11)
    for s in range(5):
        users = torch.randperm(nu, device=dev, generator=g)[:B].to(torch.int32)
        tr.step(users, s + 1)
    chk = tr.U.double().sum().reshape(1)
    both = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(both, chk)
    assert all(torch.equal(b, both[0]) for b in both), "user replicas diverged"
    ulo, uhi = shard_range(nu, world, rank)
    trl, _ = synthetic.make_interactions(uhi - ulo, ni, seed=2 + rank, device=dev)
    tu = UserShardedBPR(nu, ni, d, trl, rank, world, dev, lr=1.0, reg=0.001, init_std=0.1, seed=3)
    for s in range(5):
        users = torch.randperm(uhi - ulo, device=dev, generator=g)[:4096].to(torch.int32)
        tu.step(users, s + 1, 4096 * world)
    chk = tu.V.double().sum().reshape(1)
    dist.all_gather(both, chk)
    assert all(torch.equal(b, both[0]) for b in both), "item replicas diverged"
    for s in range(6):                                   # collective overlapped with the next step's kernel
        users = torch.randperm(uhi - ulo, device=dev, generator=g)[:4096].to(torch.int32)
        tu.step_overlapped(users, 100 + s, 4096 * world)
    tu.flush()
    torch.cuda.synchronize()
    chk = tu.V.double().sum().reshape(1)
    dist.all_gather(both, chk)
    assert all(torch.equal(b, both[0]) for b in both), "item replicas diverged (overlapped steps)"
    # the overlapped schedule against its one-step-stale numpy oracle on the union of the ranks' triples (rank 0 replays)
    import numpy as np
    from oracle import bpr_oracle as O
    nu_s, ni_s = 4000, 1200
    ulo_s, uhi_s = shard_range(nu_s, world, rank)
    trs, _ = synthetic.make_interactions(uhi_s - ulo_s, ni_s, seed=40 + rank, device=dev)
    ts = UserShardedBPR(nu_s, ni_s, d, trs, rank, world, dev, lr=300.0, reg=0.01, init_std=0.1, seed=13)
    U_init, V_init = ts.U.cpu().numpy()[:, :d], ts.V.cpu().numpy()[:, :d]
    rec = []
    for s in range(4):
        users = torch.randperm(uhi_s - ulo_s, device=dev, generator=g)[:1024].to(torch.int32)
        op, on = torch.empty_like(users), torch.empty_like(users)
        ts.step_overlapped(users, 300 + s, 1024 * world, out_pos=op, out_neg=on)
        rec.append((users.cpu().numpy() + ulo_s, op.cpu().numpy(), on.cpu().numpy()))
    ts.flush(); torch.cuda.synchronize()
    allr = [None] * world
    dist.all_gather_object(allr, (rec, U_init, ts.U.cpu().numpy()[:, :d], ts.V.cpu().numpy()[:, :d]))
    if rank == 0:
        batches = [tuple(np.concatenate([allr[r][0][s][k] for r in range(world)]) for k in range(3)) for s in range(4)]
        Ur, Vr = O.sgd_steps_stale_items(np.concatenate([a[1] for a in allr]), V_init, batches, 300.0, 0.01)
        np.testing.assert_allclose(np.concatenate([a[2] for a in allr]), Ur, rtol=2e-5, atol=5e-6)
        for a in allr:
            np.testing.assert_allclose(a[3], Vr, rtol=2e-5, atol=5e-6)
    # "update in place, exchange the difference": same SGD sums as the delta-buffer schedule up to fp32 rounding
    ta = UserShardedBPR(nu, ni, d, trl, rank, world, dev, lr=1.0, reg=0.001, init_std=0.1, seed=9)
    tb = UserShardedBPR(nu, ni, d, trl, rank, world, dev, lr=1.0, reg=0.001, init_std=0.1, seed=9)
    gg = torch.Generator(device=dev); gg.manual_seed(5 + rank)
    for s in range(6):
        users = torch.randperm(uhi - ulo, device=dev, generator=gg)[:4096].to(torch.int32)
        ta.step_overlapped(users, 200 + s, 4096 * world)
        tb.step_diff(users, 200 + s, 4096 * world)
    ta.flush(); tb.flush()
    torch.cuda.synchronize()
    # buffer schedule: gradients on pre-step item rows; diff schedule: Hogwild inside the launch -> second-order difference
    V_init = UserShardedBPR(nu, ni, d, trl, rank, world, dev, lr=1.0, reg=0.001, init_std=0.1, seed=9).V
    moved = float((ta.V - V_init).abs().max())
    assert moved > 1e-4 and float((ta.V - tb.V).abs().max()) < 0.05 * moved, (moved, float((ta.V - tb.V).abs().max()))
    assert float((ta.U - tb.U).abs().max()) < 0.05 * moved
    mx = tb.V.abs().max()
    peers = [torch.zeros_like(tb.V) for _ in range(world)]
    dist.all_gather(peers, tb.V)
    assert all(float((q - peers[0]).abs().max()) <= 1e-5 * float(mx) for q in peers), "replicas drifted beyond rounding"
    tb.resync()
    chk = tb.V.double().sum().reshape(1)
    dist.all_gather(both, chk)
    assert all(torch.equal(b, both[0]) for b in both), "resync did not make the replicas identical"
    # evaluation: item-sharded = one all-gather of the item shards + user slices; user-sharded = local users
    from recsys_pytorch_b200 import engine
    from recsys_pytorch_b200.dist import allgather_rows, evaluate_user_shard
    _, truth = synthetic.make_interactions(nu, ni, seed=1, device=dev)      # any CSR over the global users works as truth
    ev_users = torch.arange(4000, dtype=torch.int32, device=dev)
    got, n = tr.evaluate(ev_users, truth, [10])
    assert n == 4000
    V_full = allgather_rows(tr.V, ni, world, rank)
    idx, _ = engine.score_topk(tr.U, V_full, d, ev_users, train, 10)
    ref = float(engine.holdout_metrics(idx, truth, [10], row_ids=ev_users)[:, 2].double().mean())
    assert abs(got["NDCG@10"] - ref) < 1e-6, (got, ref)
    _, truth_l = synthetic.make_interactions(uhi - ulo, ni, seed=2 + rank, device=dev)
    ev_l = torch.arange(min(3000, uhi - ulo), dtype=torch.int32, device=dev)
    got_u, n_u = tu.evaluate(ev_l, truth_l, [10])
    assert n_u == world * ev_l.numel() and 0.0 <= got_u["NDCG@10"] <= 1.0
    dist.barrier()
    if rank == 0:
        print("DIST_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
