"""Generate tests/golden/*.npz by RUNNING THE REFERENCE ITSELF in this container
(CPU, torch fp32, 1 thread).  TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden          # from the repo root

Upstream ships no golden vectors - these files are the parity pin.
Each block says which reference code produced it (paths relative to
/root/reference).
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _set_weights(model, U0, V0):
    import torch
    with torch.no_grad():
        model.user_embedding.weight.copy_(torch.from_numpy(U0))
        model.item_embedding.weight.copy_(torch.from_numpy(V0))


def _loss_with_reg(model, u, i, j, reg):
    """reference loss (models/MF.py:99-107) + the engine's per-occurrence L2
    (oracle.bpr_oracle.bpr_grads docstring).  reg=0 -> reference loss exactly."""
    loss = model.process_one_batch(u, i, j)
    if reg:
        ue, ie = model.embeddings(u, i)
        _, je = model.embeddings(u, j)
        loss = loss + 0.5 * reg * (ue.pow(2).sum(1) + ie.pow(2).sum(1) + je.pow(2).sum(1)).mean()
    return loss


def tiny_bpr(ref):
    """(i) U=50, I=40, d=8, B=16 with duplicate users AND items.
    models/MF.py forward/process_one_batch + autograd + {SGD swap, Adam as-is}."""
    import torch
    rng = np.random.default_rng(7)
    nu, ni, d, B = 50, 40, 8, 16
    ds = types.SimpleNamespace(num_users=nu, num_items=ni)
    hp = {"hidden_dim": d, "pointwise": False, "loss_func": "ce"}
    U0 = rng.standard_normal((nu, d)).astype(np.float32)
    V0 = rng.standard_normal((ni, d)).astype(np.float32)
    batches = []
    for _ in range(3):
        u = rng.integers(0, nu, B); u[3] = u[0]; u[7] = u[0]
        i = rng.integers(0, ni, B); i[5] = i[1]
        j = rng.integers(0, ni, B); j[2] = i[1]; j[9] = j[4]
        batches.append((u.astype(np.int64), i.astype(np.int64), j.astype(np.int64)))
    out = dict(U0=U0, V0=V0, users=np.stack([b[0] for b in batches]),
               pos=np.stack([b[1] for b in batches]), neg=np.stack([b[2] for b in batches]))

    m = ref.MF(ds, hp, torch.device("cpu")); _set_weights(m, U0, V0)
    u, i, j = (torch.from_numpy(a) for a in batches[0])
    out["pos_scores"] = m.forward(u, i).detach().numpy()
    out["neg_scores"] = m.forward(u, j).detach().numpy()
    m.optimizer.zero_grad()
    loss = m.process_one_batch(u, i, j); loss.backward()
    out["loss"] = np.float32(loss.item())
    out["dU"] = m.user_embedding.weight.grad.numpy().copy()
    out["dV"] = m.item_embedding.weight.grad.numpy().copy()

    for tag, lr, reg in (("sgd", 0.5, 0.0), ("sgdreg", 0.5, 0.05)):
        m = ref.MF(ds, hp, torch.device("cpu")); _set_weights(m, U0, V0)
        m.optimizer = torch.optim.SGD(m.parameters(), lr=lr)       # SURVEY H1 swap
        losses = []
        for (u, i, j) in batches:
            u, i, j = (torch.from_numpy(a) for a in (u, i, j))
            m.optimizer.zero_grad()
            ls = _loss_with_reg(m, u, i, j, reg); ls.backward(); m.optimizer.step()
            losses.append(ls.item())
        out[f"{tag}_lr"], out[f"{tag}_reg"] = np.float32(lr), np.float32(reg)
        out[f"{tag}_U"] = m.user_embedding.weight.detach().numpy().copy()
        out[f"{tag}_V"] = m.item_embedding.weight.detach().numpy().copy()
        out[f"{tag}_loss"] = np.array(losses, np.float32)

    m = ref.MF(ds, hp, torch.device("cpu")); _set_weights(m, U0, V0)   # Adam as-is (MF.py:30)
    losses = []
    for (u, i, j) in batches:
        u, i, j = (torch.from_numpy(a) for a in (u, i, j))
        m.optimizer.zero_grad(); ls = m.process_one_batch(u, i, j); ls.backward(); m.optimizer.step()
        losses.append(ls.item())
    out["adam_U"] = m.user_embedding.weight.detach().numpy().copy()
    out["adam_V"] = m.item_embedding.weight.detach().numpy().copy()
    out["adam_loss"] = np.array(losses, np.float32)
    np.savez_compressed(os.path.join(OUT, "tiny_bpr.npz"), **out)
    print("tiny_bpr ok, loss", out["loss"])


def _fit_recorded(ref, model, gen, evaluator, epochs, reg=0.0):
    """models/MF.py:59-95 loop, restated so the batches can be recorded."""
    import torch
    rec_u, rec_i, rec_j, losses, scores = [], [], [], [], []
    for _ in range(epochs):
        model.train()
        for (bu, bp, bn) in gen:
            rec_u.append(bu.numpy().copy()); rec_i.append(bp.numpy().copy()); rec_j.append(bn.numpy().copy())
            model.optimizer.zero_grad()
            ls = _loss_with_reg(model, bu, bp, bn, reg)
            ls.backward(); model.optimizer.step()
            losses.append(ls.item())
        scores.append(evaluator.evaluate(model))
    return rec_u, rec_i, rec_j, losses, scores


def _pack_batches(lst):
    lens = np.array([len(a) for a in lst], np.int32)
    return np.concatenate(lst).astype(np.int32), lens


def ml100k(ref, native):
    """(ii) main.py:30-70 sequence on ml-100k, d=32, B=256, seed 2020, ks=[5,10]."""
    import torch
    ref.set_random_seed(2020)
    ds = ref_harness.ml100k_dataset(ref)
    tr = ds.train_data.tocsr(); tr.sort_indices()
    va = ds.valid_target.tocsr(); va.sort_indices()
    ev = ref.Evaluator(ds.valid_input, ds.valid_target, protocol=ds.protocol, ks=[5, 10])
    hp = {"hidden_dim": 32, "pointwise": False, "loss_func": "ce"}
    out = dict(num_users=ds.num_users, num_items=ds.num_items,
               train_indptr=tr.indptr.astype(np.int64), train_indices=tr.indices.astype(np.int32),
               valid_indptr=va.indptr.astype(np.int64), valid_indices=va.indices.astype(np.int32))

    for tag in ("adam", "sgd"):
        ref.set_random_seed(2020)
        m = ref.MF(ds, hp, torch.device("cpu"))
        U0 = m.user_embedding.weight.detach().numpy().copy()
        V0 = m.item_embedding.weight.detach().numpy().copy()
        reg = 0.0
        if tag == "sgd":
            m.optimizer = torch.optim.SGD(m.parameters(), lr=2.0); reg = 0.01
            out["sgd_lr"], out["sgd_reg"] = np.float32(2.0), np.float32(reg)
        gen = ref.PairwiseGenerator(tr, num_negatives=1, num_positives_per_user=1,
                                    batch_size=256, shuffle=True, device=torch.device("cpu"))
        ru, ri, rj, losses, scores = _fit_recorded(ref, m, gen, ev, epochs=3, reg=reg)
        out[f"{tag}_U0"], out[f"{tag}_V0"] = U0, V0
        out[f"{tag}_bu"], out[f"{tag}_blen"] = _pack_batches(ru)
        out[f"{tag}_bi"], _ = _pack_batches(ri)
        out[f"{tag}_bj"], _ = _pack_batches(rj)
        out[f"{tag}_losses"] = np.array(losses, np.float32)
        out[f"{tag}_U"] = m.user_embedding.weight.detach().numpy().copy()
        out[f"{tag}_V"] = m.item_embedding.weight.detach().numpy().copy()
        for k in scores[0]:
            out[f"{tag}_{k}"] = np.array([s[k] for s in scores], np.float32)
        # final-epoch top-10 and per-user metric rows through the reference's OWN C++ (oracle/_ref)
        pred = m.predict(np.arange(ds.num_users), ds.valid_input, 1024).astype(np.float32)
        top = native.topk(pred, 10)
        out[f"{tag}_top10"] = top
        out[f"{tag}_top10_scores"] = np.take_along_axis(pred, top.astype(np.int64), 1)
        truths = [va.indices[va.indptr[u]:va.indptr[u + 1]] for u in range(ds.num_users)]
        out[f"{tag}_metric_rows"] = native.holdout(top, truths, [5, 10])
        print(tag, "NDCG@10 per epoch", out[f"{tag}_NDCG@10"],
              "C++ mean", out[f"{tag}_metric_rows"][:, 5].mean(dtype=np.float32))
    np.savez_compressed(os.path.join(OUT, "ml100k.npz"), **out)
    return ds


def eval_blocks(ref, native):
    """(iv) top-K / holdout / LOO through the reference's C++ (oracle/_ref,
    func.h / holdout.h / loo.h) and its numpy twins (backend/python/*.py)."""
    rng = np.random.default_rng(11)
    S = rng.standard_normal((48, 3000)).astype(np.float32)
    S[5, rng.choice(3000, 2990, replace=False)] = -np.inf      # fewer than K unmasked (H6)
    S[6, :] = np.round(S[6, :], 1)                              # heavy ties
    top_cy = native.topk(S, 100)
    top_py = ref.predict_topk_py(S, 100)
    out = dict(scores=S, top100_cpp=top_cy, top100_py=top_py.astype(np.int32))

    ks = [1, 5, 10, 50]
    topk = np.stack([rng.permutation(400)[:50] for _ in range(200)]).astype(np.int32)
    truths = [np.sort(rng.choice(400, rng.integers(1, 60), replace=False)).astype(np.int32) for _ in range(200)]
    tptr = np.zeros(201, np.int64); tptr[1:] = np.cumsum([len(t) for t in truths])
    out.update(m_topk=topk, m_truth_indptr=tptr, m_truth_indices=np.concatenate(truths), m_ks=np.array(ks, np.int32))
    out["holdout_cpp"] = native.holdout(topk, truths, ks)
    tgt = {u: truths[u] for u in range(200)}
    cum = ref.compute_holdout_metrics_py(topk, tgt, ks)
    out["holdout_py"] = np.stack([[cum[mn][k].history[u] for mn in ("Prec", "Recall", "NDCG") for k in ks]
                                  for u in range(200)]).astype(np.float64)
    out["holdout_py_mean"] = np.array([cum[mn][k].mean for mn in ("Prec", "Recall", "NDCG") for k in ks], np.float32)
    out["loo_cpp"] = native.loo(topk, truths, ks)
    cum = ref.compute_loo_metrics_py(topk, {u: truths[u][:1] for u in range(200)}, ks)
    out["loo_py"] = np.stack([[cum[mn][k].history[u] for mn in ("HR", "NDCG") for k in ks]
                              for u in range(200)]).astype(np.float64)
    np.savez_compressed(os.path.join(OUT, "eval_blocks.npz"), **out)
    print("eval_blocks ok; cpp==py top100 rows:",
          int((top_cy == top_py).all(1).sum()), "/ 48")


def lightgcn(ref, ds):
    """(v) models/LightGCN.py getSparseGraph :228-267, _lightgcn_embedding
    :174-202, process_one_batch :117-123 on ml-100k, L=3, d=16."""
    import torch
    ref.set_random_seed(2020)
    hp = {"emb_dim": 16, "num_layers": 3, "node_dropout": 0.0, "split": False, "num_folds": 100,
          "graph_dir": os.path.join(ref_harness.WORK, "graph"), "reg": 1e-4}
    m = ref.LightGCN(ds, hp, torch.device("cpu"))
    tr = ds.train_data.tocsr(); tr.sort_indices()
    m.Graph = m.getSparseGraph(tr)
    g = m.Graph.coalesce()
    out = dict(adj_nnz=g.values().numel(), adj_rows=g.indices()[0].numpy().astype(np.int32),
               adj_cols=g.indices()[1].numpy().astype(np.int32), adj_vals=g.values().numpy())
    U0 = m.user_embedding.weight.detach().numpy().copy(); V0 = m.item_embedding.weight.detach().numpy().copy()
    out.update(U0=U0, V0=V0)
    m.update_lightgcn_embedding()
    out["prop_U"] = m.user_embeddings.detach().numpy().copy()
    out["prop_V"] = m.item_embeddings.detach().numpy().copy()
    rng = np.random.default_rng(3)
    u = rng.integers(0, ds.num_users, 256); i = rng.integers(0, ds.num_items, 256); j = rng.integers(0, ds.num_items, 256)
    out.update(users=u.astype(np.int32), pos=i.astype(np.int32), neg=j.astype(np.int32))
    m.optimizer = torch.optim.SGD(m.parameters(), lr=10.0)
    m.optimizer.zero_grad()
    ls = m.process_one_batch(*(torch.from_numpy(a) for a in (u, i, j))); ls.backward()
    out["loss"] = np.float32(ls.item())
    out["dU0"] = m.user_embedding.weight.grad.numpy().copy()
    out["dV0"] = m.item_embedding.weight.grad.numpy().copy()
    m.optimizer.step()
    out["sgd_lr"] = np.float32(10.0)
    out["sgd_U"] = m.user_embedding.weight.detach().numpy().copy()
    out["sgd_V"] = m.item_embedding.weight.detach().numpy().copy()
    np.savez_compressed(os.path.join(OUT, "lightgcn_ml100k.npz"), **out)
    print("lightgcn ok, loss", out["loss"], "adj nnz", out["adj_nnz"])


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = ref_harness.load()
    native = ref_harness.RefNative()
    tiny_bpr(ref)
    eval_blocks(ref, native)
    ds = ml100k(ref, native)
    lightgcn(ref, ds)


if __name__ == "__main__":
    main()
