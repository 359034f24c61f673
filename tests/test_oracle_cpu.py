"""`-m "not gpu"`: the oracle is pinned against the golden vectors produced by the
reference itself (oracle/make_golden.py); host logic; the C-ABI library loads and
exports every symbol include/b200rec.h declares."""
import os
import re

import numpy as np
import pytest

from oracle import bpr_oracle as O
from tests.util import truths_from_csr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- BPR arithmetic vs reference autograd / optimisers (tests/golden/tiny_bpr.npz) ----
def test_oracle_forward_loss_grads(golden):
    g = golden["tiny_bpr"]
    U0, V0 = g["U0"], g["V0"]
    u, i, j = g["users"][0], g["pos"][0], g["neg"][0]
    np.testing.assert_allclose(O.forward_scores(U0, V0, u, i), g["pos_scores"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(O.forward_scores(U0, V0, u, j), g["neg_scores"], rtol=1e-6, atol=1e-6)
    loss, _ = O.bpr_loss(U0, V0, u, i, j)
    assert abs(float(loss) - float(g["loss"])) < 1e-6
    dU, dV, _, _ = O.bpr_grads(U0, V0, u, i, j)
    np.testing.assert_allclose(dU, g["dU"], rtol=1e-5, atol=1e-7)      # duplicates included
    np.testing.assert_allclose(dV, g["dV"], rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("tag", ["sgd", "sgdreg"])
def test_oracle_sgd_trajectory(golden, tag):
    g = golden["tiny_bpr"]
    U, V = g["U0"].copy(), g["V0"].copy()
    for b in range(3):
        U, V, _ = O.sgd_step(U, V, g["users"][b], g["pos"][b], g["neg"][b], float(g[f"{tag}_lr"]), float(g[f"{tag}_reg"]))
    np.testing.assert_allclose(U, g[f"{tag}_U"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(V, g[f"{tag}_V"], rtol=2e-5, atol=2e-6)


def test_oracle_dense_adam(golden):
    g = golden["tiny_bpr"]
    U, V = g["U0"].copy(), g["V0"].copy()
    opt = O.DenseAdam([U.shape, V.shape])
    for b in range(3):
        dU, dV, _, _ = O.bpr_grads(U, V, g["users"][b], g["pos"][b], g["neg"][b])
        U, V = opt.step([U, V], [dU, dV])
    np.testing.assert_allclose(U, g["adam_U"], rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(V, g["adam_V"], rtol=1e-4, atol=2e-6)


def test_oracle_ml100k_sgd_replay(golden):
    """Replaying the recorded reference batches through the oracle reproduces the reference's tables."""
    g = golden["ml100k"]
    U, V = g["sgd_U0"].copy(), g["sgd_V0"].copy()
    off = 0
    for n in g["sgd_blen"]:
        sl = slice(off, off + n); off += n
        U, V, _ = O.sgd_step(U, V, g["sgd_bu"][sl], g["sgd_bi"][sl], g["sgd_bj"][sl], float(g["sgd_lr"]), float(g["sgd_reg"]))
    np.testing.assert_allclose(U, g["sgd_U"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(V, g["sgd_V"], rtol=1e-4, atol=1e-5)


# ---- evaluation layer vs the reference's own C++ (oracle/_ref) and numpy twins ----
def test_oracle_topk_vs_reference(golden, oracle_c):
    g = golden["eval_blocks"]
    S = g["scores"]
    mine = oracle_c.topk(S, 100)
    np.testing.assert_array_equal(mine, O.topk_desc(S, 100))            # C and numpy restatements agree
    for name in ("top100_cpp", "top100_py"):
        ref = g[name].astype(np.int64)
        # same score at every rank (tie order is unspecified upstream)
        np.testing.assert_array_equal(np.take_along_axis(S, mine.astype(np.int64), 1), np.take_along_axis(S, ref, 1))
    tie_free = [r for r in range(S.shape[0]) if r not in (5, 6)]
    np.testing.assert_array_equal(mine[tie_free], g["top100_cpp"][tie_free])


def test_oracle_metrics_vs_reference(golden, oracle_c):
    g = golden["eval_blocks"]
    truths = truths_from_csr(g["m_truth_indptr"], g["m_truth_indices"])
    ks = g["m_ks"]
    np.testing.assert_array_equal(oracle_c.holdout(g["m_topk"], truths, ks), g["holdout_cpp"])   # bit-exact vs holdout.h
    np.testing.assert_array_equal(O.holdout_metrics(g["m_topk"], truths, ks), g["holdout_cpp"])
    np.testing.assert_allclose(g["holdout_cpp"], g["holdout_py"], rtol=1e-6, atol=1e-7)           # python twin
    np.testing.assert_array_equal(oracle_c.loo(g["m_topk"], truths, ks), g["loo_cpp"])
    np.testing.assert_array_equal(O.loo_metrics(g["m_topk"], truths, ks), g["loo_cpp"])
    np.testing.assert_allclose(g["loo_cpp"], g["loo_py"], rtol=1e-6, atol=1e-7)


def test_oracle_ml100k_eval(golden, oracle_c):
    g = golden["ml100k"]
    nu, ni = int(g["num_users"]), int(g["num_items"])
    for tag in ("adam", "sgd"):
        idx, sc = oracle_c.score_topk(g[f"{tag}_U"], g[f"{tag}_V"], 32, np.arange(nu), ni, g["train_indptr"],
                                      g["train_indices"], 10)
        np.testing.assert_allclose(sc, g[f"{tag}_top10_scores"], rtol=1e-5, atol=1e-5)
        assert (idx == g[f"{tag}_top10"]).mean() > 0.999
        truths = truths_from_csr(g["valid_indptr"], g["valid_indices"])
        rows = oracle_c.holdout(idx, truths, [5, 10])
        ndcg10 = O.mean_f32(rows[:, 5])
        assert abs(float(ndcg10) - float(g[f"{tag}_NDCG@10"][-1])) < 1e-6


def test_oracle_lightgcn(golden):
    g = golden["lightgcn_ml100k"]
    ml = golden["ml100k"]
    import scipy.sparse as sp
    nu, ni = int(ml["num_users"]), int(ml["num_items"])
    R = sp.csr_matrix((np.ones(len(ml["train_indices"]), np.float32), ml["train_indices"], ml["train_indptr"]), shape=(nu, ni))
    A = O.lightgcn_adj(R)
    ref = sp.csr_matrix((g["adj_vals"], (g["adj_rows"], g["adj_cols"])), shape=A.shape)
    assert A.nnz == int(g["adj_nnz"])
    assert abs(A - ref).max() < 1e-7
    out = O.lightgcn_propagate(A, np.concatenate([g["U0"], g["V0"]]), 3)
    np.testing.assert_allclose(out[:nu], g["prop_U"], rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(out[nu:], g["prop_V"], rtol=1e-4, atol=1e-7)


# ---- host logic ---------------------------------------------------------------
def test_build_norm_adj_on_cpu_tensors_matches_reference_graph(golden):
    """lightgcn.build_norm_adj is plain torch array code: run on CPU tensors it must give the adjacency the reference's
    getSparseGraph built (models/LightGCN.py:228-267; golden from oracle/make_golden.py::lightgcn) - same pattern, values
    within 2 ulp (pow(-0.5) rounding).  The `-m gpu` twin runs the same function on CUDA tensors."""
    import types
    import torch
    from recsys_pytorch_b200.lightgcn import build_norm_adj
    g, ml = golden["lightgcn_ml100k"], golden["ml100k"]
    nu, ni = int(ml["num_users"]), int(ml["num_items"])
    csr = types.SimpleNamespace(indptr=torch.from_numpy(ml["train_indptr"].astype(np.int64)),
                                indices=torch.from_numpy(ml["train_indices"].astype(np.int32)), shape=(nu, ni),
                                nnz=int(len(ml["train_indices"])))
    indptr, cols, vals = build_norm_adj(csr)
    assert indptr.dtype == torch.int64 and cols.dtype == torch.int32 and vals.dtype == torch.float32
    assert cols.numel() == int(g["adj_nnz"])
    order = np.lexsort((g["adj_cols"], g["adj_rows"]))
    np.testing.assert_array_equal(cols.numpy(), g["adj_cols"][order])
    np.testing.assert_allclose(vals.numpy(), g["adj_vals"][order], rtol=5e-7, atol=0)
    np.testing.assert_array_equal(np.diff(indptr.numpy()), np.bincount(g["adj_rows"], minlength=nu + ni))
    # symmetric: the backward pass reuses the forward propagation on that ground
    import scipy.sparse as sp
    A = sp.csr_matrix((vals.numpy(), cols.numpy(), indptr.numpy()), shape=(nu + ni, nu + ni))
    assert abs(A - A.T).max() == 0.0


def test_reference_sampler_restatement(golden):
    """sampler='reference' replays data/generators.py:168-224 draw for draw (same numpy seed -> same batches)."""
    import random
    import scipy.sparse as sp
    import torch
    from recsys_pytorch_b200.generators import PairwiseGenerator
    g = golden["ml100k"]
    nu, ni = int(g["num_users"]), int(g["num_items"])
    R = sp.csr_matrix((np.ones(len(g["train_indices"])), g["train_indices"], g["train_indptr"]), shape=(nu, ni))
    random.seed(2020); np.random.seed(2020); torch.manual_seed(2020)        # utils/general.py:31-38
    gen = PairwiseGenerator(R, num_negatives=1, num_positives_per_user=1, batch_size=256, shuffle=True,
                            device="cpu", sampler="reference")
    assert len(gen) == 4
    bu, bi, bj = [], [], []
    for _ in range(3):
        for (u, i, j) in gen:
            bu.append(u.numpy()); bi.append(i.numpy()); bj.append(j.numpy())
    np.testing.assert_array_equal(np.concatenate(bu), g["adam_bu"])
    np.testing.assert_array_equal(np.concatenate(bi), g["adam_bi"])
    np.testing.assert_array_equal(np.concatenate(bj), g["adam_bj"])


def test_sampler_mirror_properties():
    rng = np.random.default_rng(5)
    ni = 200
    rows = [np.sort(rng.choice(ni, rng.integers(1, 150), replace=False)).astype(np.int32) for _ in range(50)]
    indptr = np.zeros(51, np.int64); indptr[1:] = np.cumsum([len(r) for r in rows])
    indices = np.concatenate(rows)
    for t in range(400):
        u = t % 50
        p, n = O.sample_triple(9, 3, t, u, indptr, indices, ni)
        assert p in rows[u] and n not in rows[u] and 0 <= n < ni


def test_statistics_and_cpu_refusal():
    from recsys_pytorch_b200.evaluation import Statistics
    st = Statistics("x"); st.update(1.0); st.update([2.0, 3.0]); st.update(np.float32(4.0))
    assert st.cnt == 4 and abs(float(st.mean) - 2.5) < 1e-7
    import types
    import torch
    from recsys_pytorch_b200 import B200RecError
    from recsys_pytorch_b200.mf import MF
    with pytest.raises(B200RecError):        # no CPU fallback: a CPU device is refused loudly
        MF(types.SimpleNamespace(num_users=4, num_items=4), {"hidden_dim": 8}, torch.device("cpu"))


# ---- C ABI ---------------------------------------------------------------------
def test_cabi_exports_every_declared_symbol(built_lib):
    from recsys_pytorch_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "b200rec.h")).read()
    declared = set(re.findall(r"\b(b200rec_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(built_lib, name), f"{name} declared in include/b200rec.h but not exported"
    assert declared == set(_lib.EXPORTS)
    assert built_lib.b200rec_version() >= 100


def test_cabi_argument_errors_without_gpu(built_lib):
    """Argument validation and the no-device refusal run on a CPU-only box (no compute)."""
    from recsys_pytorch_b200 import _lib
    import ctypes as C
    assert built_lib.b200rec_bpr_step(None, None) == _lib.EINVAL
    assert b"NULL" in built_lib.b200rec_last_error()
    sc = np.zeros((2, 8), np.float32); out = np.zeros((2, 4), np.int32)
    assert built_lib.b200rec_top_k_array_index(sc.ctypes.data, 8, 2, 9, out.ctypes.data) == _lib.EINVAL   # k > cols
    assert built_lib.b200rec_pointwise_step(None, None, 8, 8, None, None, None, 4, 0, 0.1, 0.0, 0, None, None, None, 0.0,
                                            None) == _lib.EINVAL
    assert built_lib.b200rec_spmm_csr_split(None, None, None, 4, None, 8, 8, None, 8, None, 8, 1.0, 0, 256, None, None, 0,
                                            None, None, 0, None, None) == _lib.EINVAL
    assert built_lib.b200rec_delta_diff(None, None, None, None, 8, None) == _lib.EINVAL
    import torch
    if not torch.cuda.is_available():
        rc = built_lib.b200rec_top_k_array_index(sc.ctypes.data, 8, 2, 4, out.ctypes.data)
        assert rc == _lib.ECUDA and b"no CPU fallback" in built_lib.b200rec_last_error()


def test_torch_port_reproduces_reference_adam(golden):
    """oracle/torch_port.py (the timed CPU baseline) IS the reference's step: same tables after 3 Adam steps."""
    import torch
    from oracle.torch_port import RefMF
    g = golden["tiny_bpr"]
    torch.set_num_threads(1)
    m = RefMF(50, 40, 8)
    with torch.no_grad():
        m.user_embedding.weight.copy_(torch.from_numpy(g["U0"])); m.item_embedding.weight.copy_(torch.from_numpy(g["V0"]))
    losses = []
    for b in range(3):
        u, i, j = (torch.from_numpy(g[k][b]) for k in ("users", "pos", "neg"))
        losses.append(float(m.train_step(u, i, j)))
    np.testing.assert_allclose(m.user_embedding.weight.detach().numpy(), g["adam_U"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(m.item_embedding.weight.detach().numpy(), g["adam_V"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(losses, g["adam_loss"], rtol=1e-6)


def test_oracle_sparse_adam(golden):
    """Row-wise (lazy) Adam restatement vs the reference MF driven by torch.optim.SparseAdam (6 steps)."""
    g, t = golden["tiny_bpr"], golden["tiny_lazy_adam"]
    U, V = g["U0"].copy(), g["V0"].copy()
    opt = O.SparseAdam([U.shape, V.shape])
    s = 0
    for rep in range(2):
        for b in range(3):
            u, i, j = g["users"][b], g["pos"][b], g["neg"][b]
            dU, dV, _, _ = O.bpr_grads(U, V, u, i, j)
            U, V = opt.step([U, V], [dU, dV], [u, np.concatenate([i, j])])
            np.testing.assert_allclose(U, t["U"][s], rtol=1e-4, atol=2e-6)
            np.testing.assert_allclose(V, t["V"][s], rtol=1e-4, atol=2e-6)
            s += 1


# ---- pointwise MF mode: oracle + host generator pinned to the reference ----
@pytest.mark.parametrize("lf", ["ce", "mse"])
def test_oracle_pointwise_loss_grads_and_adam(golden, lf):
    g = golden["tiny_pointwise"]
    sc = float(g[f"{lf}_scale"])
    t = golden["tiny_bpr"]
    U, V = (t["U0"] * sc).astype(np.float32), (t["V0"] * sc).astype(np.float32)
    u, i, r = g["users"][0], g["items"][0], g["ratings"][0]
    loss, x = O.pointwise_loss(U, V, u, i, r, lf)
    np.testing.assert_allclose(x, g[f"{lf}_scores"], rtol=2e-6, atol=2e-6)
    assert abs(float(loss) - float(g[f"{lf}_loss"])) <= 2e-6 * max(1.0, abs(float(g[f"{lf}_loss"])))
    dU, dV, _, _ = O.pointwise_grads(U, V, u, i, r, lf)
    np.testing.assert_allclose(dU, g[f"{lf}_dU"], rtol=2e-5, atol=2e-7)
    np.testing.assert_allclose(dV, g[f"{lf}_dV"], rtol=2e-5, atol=2e-7)
    opt = O.DenseAdam([U.shape, V.shape])                      # the reference's optimiser as-is (MF.py:30)
    for b in range(3):
        dU, dV, _, _ = O.pointwise_grads(U, V, g["users"][b], g["items"][b], g["ratings"][b], lf)
        U, V = opt.step([U, V], [dU, dV])
    np.testing.assert_allclose(U, g[f"{lf}_adam_U"], rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(V, g[f"{lf}_adam_V"], rtol=1e-4, atol=2e-6)


def test_pointwise_generator_mirror_reproduces_reference_batches(golden):
    import scipy.sparse as sp
    from recsys_pytorch_b200.generators import PointwiseGenerator
    g = golden["tiny_pointwise"]
    R = sp.csr_matrix((np.ones(len(g["R_indices"])), g["R_indices"], g["R_indptr"]), shape=tuple(g["R_shape"]))
    np.random.seed(99)
    gen = PointwiseGenerator(R, return_rating=True, num_negatives=1, batch_size=32, shuffle=True, as_numpy=True)
    assert len(gen) == int(g["gen_num_batches"])
    bu, bi, br, lens = [], [], [], []
    for _ in range(2):
        for (a, b, c) in gen:
            bu.append(a); bi.append(b); br.append(c); lens.append(len(a))
    np.testing.assert_array_equal(lens, g["gen_lens"])
    np.testing.assert_array_equal(np.concatenate(bu), g["gen_users"])
    np.testing.assert_array_equal(np.concatenate(bi), g["gen_items"])
    np.testing.assert_array_equal(np.concatenate(br).astype(np.float32), g["gen_ratings"])


def test_stale_item_oracle_reduces_to_sgd_step_and_lags_one_step():
    """oracle/bpr_oracle.py::sgd_steps_stale_items (the overlapped multi-GPU schedule): one batch == sgd_step; with two
    batches the second step's gradients see U_1 but still V_0."""
    rng = np.random.default_rng(0)
    U = rng.standard_normal((30, 8)).astype(np.float32); V = rng.standard_normal((20, 8)).astype(np.float32)
    b1 = (rng.permutation(30)[:10], rng.integers(0, 20, 10), rng.integers(0, 20, 10))
    b2 = (rng.permutation(30)[:10], rng.integers(0, 20, 10), rng.integers(0, 20, 10))
    a = O.sgd_steps_stale_items(U, V, [b1], 0.5, 0.01); c = O.sgd_step(U, V, *b1, 0.5, 0.01)
    assert np.array_equal(a[0], c[0]) and np.array_equal(a[1], c[1])
    U2, V2 = O.sgd_steps_stale_items(U, V, [b1, b2], 0.5, 0.01)
    dU2, dV2, _, _ = O.bpr_grads(c[0], V, *b2, 0.01)                 # gradients of step 2 at (U_1, V_0)
    dU1, dV1, _, _ = O.bpr_grads(U, V, *b1, 0.01)
    np.testing.assert_allclose(U2, c[0] - np.float32(0.5) * dU2, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(V2, V - np.float32(0.5) * dV1 - np.float32(0.5) * dV2, rtol=1e-6, atol=1e-6)


def test_vectorised_sampler_mirror_equals_scalar_mirror():
    rng = np.random.default_rng(4)
    nu, ni = 300, 90
    rows = [np.sort(rng.choice(ni, rng.integers(1, 70), replace=False)).astype(np.int32) for _ in range(nu)]
    rows[7] = np.arange(ni, dtype=np.int32)                          # a user who has everything: neg = -1
    indptr = np.zeros(nu + 1, np.int64); indptr[1:] = np.cumsum([len(r) for r in rows])
    indices = np.concatenate(rows)
    users = rng.permutation(nu)[:200]
    users[3] = 7
    p, n = O.sample_triples_vec(11, 5, users, indptr, indices, ni)
    for t, u in enumerate(users):
        assert (int(p[t]), int(n[t])) == O.sample_triple(11, 5, t, int(u), indptr, indices, ni)
    assert n[3] == -1


def test_pointwise_oracle_replays_reference_ml100k_trajectory(golden):
    """tests/golden/ml100k_pointwise.npz (the reference's pointwise MF + its own generator's batches, 6 dense-Adam
    steps): the numpy restatement reproduces losses and tables."""
    g = golden["ml100k_pointwise"]
    for lf in ("ce", "mse"):
        U, V = g[f"{lf}_U0"].copy(), g[f"{lf}_V0"].copy()
        opt = O.DenseAdam([U.shape, V.shape], lr=1e-3)
        off = 0
        for s_, n in enumerate(g[f"{lf}_lens"]):
            u, i, r = g[f"{lf}_users"][off:off + n], g[f"{lf}_items"][off:off + n], g[f"{lf}_ratings"][off:off + n]
            off += n
            loss = O.pointwise_loss(U, V, u, i, r, lf)
            loss = loss[0] if isinstance(loss, tuple) else loss
            assert abs(float(loss) - float(g[f"{lf}_losses"][s_])) < 2e-5 * max(1.0, float(g[f"{lf}_losses"][s_]))
            dU, dV, _, _ = O.pointwise_grads(U, V, u, i, r, lf)
            U, V = opt.step([U, V], [dU, dV])
        for got, ref in ((U, g[f"{lf}_U"]), (V, g[f"{lf}_V"])):
            bad = ~np.isclose(got, ref, rtol=2e-4, atol=2e-5)
            assert bad.mean() < 0.01 and np.abs(got - ref).max() < 7 * 1e-3


def test_ngcf_oracle_matches_reference_forward_and_autograd(golden):
    """tests/golden/ngcf_ml100k.npz (reference NGCF, models/NGCF.py:182-221 + autograd): the numpy restatement of the
    propagation and its closed-form backward reproduce the propagated tables, the loss and every parameter gradient."""
    g, lg = golden["ngcf_ml100k"], golden["lightgcn_ml100k"]
    nu, ni = g["U0"].shape[0], g["V0"].shape[0]
    import scipy.sparse as sp
    A = sp.csr_matrix((lg["adj_vals"], (lg["adj_rows"], lg["adj_cols"])), shape=(nu + ni, nu + ni))   # same graph recipe
    E0 = np.concatenate([g["U0"], g["V0"]])
    Wg, bg, Wb, bb = ([g["%s_%d" % (nm, k)] for k in range(2)] for nm in ("W_gc", "b_gc", "W_bi", "b_bi"))
    out, cache = O.ngcf_forward(A, E0, Wg, bg, Wb, bb, keep=True)
    np.testing.assert_allclose(out[:nu], g["prop_U"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(out[nu:], g["prop_V"], rtol=2e-5, atol=2e-6)
    u, i, j = g["users"], g["pos"], g["neg"]
    loss, _ = O.bpr_loss(out[:nu], out[nu:], u, i, j)
    assert abs(float(loss) - float(g["loss"])) < 2e-6
    dU, dV, _, _ = O.bpr_grads(out[:nu], out[nu:], u, i, j)
    dE0, dWg, dbg, dWb, dbb = O.ngcf_backward(A, np.concatenate([dU, dV]), cache, Wg, Wb)
    np.testing.assert_allclose(dE0[:nu], g["dU0"], rtol=2e-4, atol=2e-7)
    np.testing.assert_allclose(dE0[nu:], g["dV0"], rtol=2e-4, atol=2e-7)
    for k in range(2):
        np.testing.assert_allclose(dWg[k], g["dW_gc_%d" % k], rtol=2e-4, atol=2e-7)
        np.testing.assert_allclose(dbg[k], g["db_gc_%d" % k], rtol=2e-4, atol=2e-7)
        np.testing.assert_allclose(dWb[k], g["dW_bi_%d" % k], rtol=2e-4, atol=2e-7)
        np.testing.assert_allclose(dbb[k], g["db_bi_%d" % k], rtol=2e-4, atol=2e-7)


def test_synthetic_interactions_recipe_on_cpu():
    """synthetic.make_interactions_raw (SURVEY 8(d) recipe; plain torch, here on CPU - the host arm of bench.py builds its
    sample of the dataset with it): sorted duplicate-free rows, every user keeps >= 1 train and >= 1 target item, train and
    target are disjoint, ~20 % held out, degrees inside [dmin, dmax], Zipf popularity, and shards drawn with different
    user seeds but one `item_seed` agree on which items are popular (what the multi-GPU layouts rely on)."""
    import torch
    from recsys_pytorch_b200 import synthetic
    nu, ni = 4000, 3000
    (tp, ti), (vp, vi) = synthetic.make_interactions_raw(nu, ni, seed=11, device="cpu")
    assert tp.dtype == torch.int64 and ti.dtype == torch.int32 and tp.numel() == nu + 1 and vp.numel() == nu + 1
    tp, ti, vp, vi = (x.numpy() for x in (tp, ti, vp, vi))
    assert tp[0] == 0 and vp[0] == 0 and tp[-1] == len(ti) and vp[-1] == len(vi)
    dt, dv = np.diff(tp), np.diff(vp)
    assert dt.min() >= 1 and dv.min() >= 1
    tot = dt + dv
    assert tot.max() <= 1000 and tot.min() >= 2 and np.median(tot) > 15           # clip(lognormal(3.5, .8), 10, 1000) minus in-row duplicates
    assert np.all(dv == np.minimum(np.maximum(np.ceil(tot * 0.2), 1), tot - 1))    # ceil(20 %) of every row, >= 1, train keeps >= 1
    for r in range(0, nu, 37):
        a, b = ti[tp[r]:tp[r + 1]], vi[vp[r]:vp[r + 1]]
        assert np.all(np.diff(a) > 0) and np.all(np.diff(b) > 0) and len(np.intersect1d(a, b)) == 0
        assert a.min() >= 0 and a.max() < ni
    pop = np.bincount(np.concatenate([ti, vi]), minlength=ni)
    assert pop.max() > 20 * np.median(pop)                                           # Zipf(1) head
    # one catalogue popularity for shards with different users
    (ap, ai), _ = synthetic.make_interactions_raw(nu, ni, seed=1, device="cpu", item_seed=99)
    (bp, bi), _ = synthetic.make_interactions_raw(nu, ni, seed=2, device="cpu", item_seed=99)
    assert not np.array_equal(ai.numpy()[:200], bi.numpy()[:200])
    ha, hb = np.bincount(ai.numpy(), minlength=ni), np.bincount(bi.numpy(), minlength=ni)
    top_a, top_b = set(np.argsort(-ha)[:30].tolist()), set(np.argsort(-hb)[:30].tolist())
    assert len(top_a & top_b) >= 24
    # deterministic in the seed
    (cp, ci), _ = synthetic.make_interactions_raw(nu, ni, seed=1, device="cpu", item_seed=99)
    assert torch.equal(ap, cp) and torch.equal(ai, ci)


def test_oracle_at_cfg2_row_width_and_batch_regime(golden):
    """SURVEY 8(c) golden (iii): d = 128, ONE 65,536-triple batch with a Zipf head (hottest item 7,164 times) through the
    reference's forward / process_one_batch / autograd / SGD swap (oracle/make_golden_cfg2shape.py).  The numpy oracle
    reproduces scores, loss, the sampled dense-gradient rows and the post-step rows; its fp64 accumulation differs from the
    reference's fp32 `embedding_dense_backward` by < 1e-5 of the gradient scale."""
    from oracle.make_golden_cfg2shape import LR, inputs
    g = golden["cfg2shape_bpr"]
    U0, V0, u, i, j, su, si = inputs(int(g["seed"]))
    assert np.array_equal(su, g["sample_users"]) and np.array_equal(si, g["sample_items"])     # same regenerated inputs
    dU, dV, _, x = O.bpr_grads(U0, V0, u, i, j)
    loss, _ = O.bpr_loss(U0, V0, u, i, j)
    np.testing.assert_allclose(x, g["x"], rtol=1e-5, atol=5e-7)
    assert abs(float(loss) - float(g["loss"])) < 1e-6
    for got, ref in ((dU[su], g["dU_rows"]), (dV[si], g["dV_rows"])):
        np.testing.assert_allclose(got, ref, rtol=2e-5, atol=1e-5 * np.abs(ref).max())
    assert abs(np.abs(dU.astype(np.float64)).sum() / float(g["dU_abs_sum"]) - 1) < 1e-6          # the unsampled rows too
    assert abs(np.abs(dV.astype(np.float64)).sum() / float(g["dV_abs_sum"]) - 1) < 1e-6
    Ur, Vr, _ = O.sgd_step(U0, V0, u, i, j, float(LR), 0.0)
    np.testing.assert_allclose(Ur[su], g["U_rows"], rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(Vr[si], g["V_rows"], rtol=2e-6, atol=1e-7)


def test_torch_port_lightgcn_matches_reference_golden(golden):
    """oracle/torch_port.py's LightGCN restatement (the timed CPU baseline of bench.py's cfg4 leg) against what the
    reference's getSparseGraph / _lightgcn_embedding produced on ml-100k."""
    import torch
    from oracle import torch_port as TP
    g, ml = golden["lightgcn_ml100k"], golden["ml100k"]
    nu, ni = int(ml["num_users"]), int(ml["num_items"])
    G = TP.lightgcn_graph(ml["train_indptr"], ml["train_indices"], nu, ni)
    assert G._nnz() == int(g["adj_nnz"])
    order = np.lexsort((g["adj_cols"], g["adj_rows"]))
    idx = G.indices().numpy()
    np.testing.assert_array_equal(idx[0], g["adj_rows"][order]); np.testing.assert_array_equal(idx[1], g["adj_cols"][order])
    np.testing.assert_allclose(G.values().numpy(), g["adj_vals"][order], rtol=5e-7, atol=0)
    out = TP.lightgcn_embedding(G, torch.from_numpy(np.concatenate([g["U0"], g["V0"]])), 3).numpy()
    np.testing.assert_allclose(out[:nu], g["prop_U"], rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(out[nu:], g["prop_V"], rtol=1e-5, atol=1e-8)
    fw, fb = TP.time_lightgcn(G, 16, 3, reps=1)
    assert fw > 0 and fb > 0


# ---- randomised differential pin: the C restatement vs the reference's OWN C++ (oracle/_ref, compiled from
# /root/reference/evaluation/backend/cython/include where it lies; the built library travels to the GPU box) ----
def _ref_native():
    path = os.path.join(ROOT, "oracle", "_ref", "libref_eval.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libref_eval.so not built (needs /root/reference at build time)")
    from oracle.ref_harness import RefNative
    return RefNative()


@pytest.mark.parametrize("seed", range(12))
def test_oracle_topk_differential_vs_reference_cpp(oracle_c, seed):
    """func.h:12-31 vs oracle/eval_oracle.c on random blocks: tie-free scores -> identical ids (every k up to the row
    length, masked -inf columns, negative and subnormal values); tied scores -> identical score at every rank."""
    ref = _ref_native()
    rng = np.random.default_rng(100 + seed)
    rows, cols = int(rng.integers(1, 40)), int(rng.integers(1, 3000))
    vals = (rng.permutation(rows * cols).astype(np.float64) - rows * cols / 2).reshape(rows, cols)
    S = (vals * rng.choice([1e-3, 1.0, 37.5, 1e-42])).astype(np.float32)                     # 1e-42: fp32 subnormals
    if len(np.unique(S)) < S.size:                                                            # rounding produced ties
        S = vals.astype(np.float32)
    n_mask = int(rng.integers(0, max(cols // 3, 1)))
    for r in range(rows):
        S[r, rng.choice(cols, n_mask, replace=False)] = -np.inf
    for k in sorted({1, min(5, cols), min(100, cols), max(cols - n_mask, 1)}):
        np.testing.assert_array_equal(oracle_c.topk(S, k), ref.topk(S, k))
    # ties: quantised scores
    T = np.round(rng.standard_normal((rows, cols)) * 3).astype(np.float32)
    k = min(20, cols)
    a, b = oracle_c.topk(T, k).astype(np.int64), ref.topk(T, k).astype(np.int64)
    np.testing.assert_array_equal(np.take_along_axis(T, a, 1), np.take_along_axis(T, b, 1))
    assert all(len(set(row)) == k for row in a.tolist())                                     # k distinct items per row
    if cols > 1:                                                                              # documented tie order: id ascending
        sc = np.take_along_axis(T, a, 1)
        same = sc[:, 1:] == sc[:, :-1]
        assert np.all(a[:, 1:][same] > a[:, :-1][same])


@pytest.mark.parametrize("seed", range(12))
def test_oracle_metrics_differential_vs_reference_cpp(oracle_c, seed):
    """holdout.h:20-103 / loo.h:20-85 vs oracle/eval_oracle.c and the numpy twin, bit for bit, on random rankings: truth
    sets of 1 .. 3K items, hits anywhere or nowhere, K lists with K = 1 and K = max_k."""
    ref = _ref_native()
    rng = np.random.default_rng(500 + seed)
    users, items = int(rng.integers(1, 300)), int(rng.integers(60, 5000))
    max_k = int(rng.integers(1, 51))
    ks = np.array(sorted(set([1, max_k] + rng.integers(1, max_k + 1, 3).tolist())), np.int32)
    topk = np.stack([rng.choice(items, max_k, replace=False) for _ in range(users)]).astype(np.int32)
    truths = []
    for u in range(users):
        n = int(rng.integers(1, 3 * max_k + 2))
        t = rng.choice(items, min(n, items), replace=False)
        if rng.random() < 0.5:                                                               # plant hits at random ranks
            m = min(len(t), 3, max_k)
            t[:m] = topk[u, rng.choice(max_k, m, replace=False)]
        truths.append(np.unique(t).astype(np.int32))
    want = ref.holdout(topk, truths, ks)
    np.testing.assert_array_equal(oracle_c.holdout(topk, truths, ks), want)
    np.testing.assert_array_equal(O.holdout_metrics(topk, truths, ks), want)
    assert np.isfinite(want).all() and want.min() >= 0 and want.max() <= 1.0 + 1e-6
    loo_truths = [t[:1] for t in truths]
    want = ref.loo(topk, loo_truths, ks)
    np.testing.assert_array_equal(oracle_c.loo(topk, loo_truths, ks), want)
    np.testing.assert_array_equal(O.loo_metrics(topk, loo_truths, ks), want)


# ---- the reference-side binding INTEGRATION.md shows (Option A) actually builds and binds ----
def test_integration_cython_shim_builds_and_binds(built_lib, tmp_path):
    """Extract the Cython shim from INTEGRATION.md section 2 (the maintainer's replacement of
    evaluation/backend/cython/func.pyx:8-25), build it against include/b200rec.h + libb200rec.so, import it and call it.
    On a box with a GPU it must return the top-k; without one the library's error must surface as RuntimeError - never a
    silent CPU answer."""
    import subprocess
    import sys
    import textwrap
    pytest.importorskip("Cython")
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    m = re.search(r"```cython\n(.*?)```", doc, re.S)
    assert m, "INTEGRATION.md lost its cython block"
    (tmp_path / "func_b200.pyx").write_text(m.group(1))
    libdir = os.path.join(ROOT, "recsys_pytorch_b200")
    (tmp_path / "setup.py").write_text(textwrap.dedent(f"""
        from setuptools import setup, Extension
        from Cython.Build import cythonize
        import numpy as np
        ext = Extension("func_b200", ["func_b200.pyx"], language="c++", include_dirs=[np.get_include(), r"{ROOT}/include"],
                        library_dirs=[r"{libdir}"], libraries=["b200rec"], runtime_library_dirs=[r"{libdir}"],
                        extra_compile_args=["-std=c++11", "-w"])
        setup(ext_modules=cythonize([ext], language_level=3, quiet=True), script_args=["build_ext", "--inplace", "-q"])
    """))
    out = subprocess.run([sys.executable, "setup.py"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    probe = textwrap.dedent("""
        import numpy as np, func_b200
        S = np.random.default_rng(0).standard_normal((7, 50)).astype(np.float32)
        try:
            top = func_b200.predict_topk_cy(S, 5)
        except RuntimeError as e:
            print("RAISED", str(e)[:200])
        else:
            ref = np.argsort(-S, axis=1, kind="stable")[:, :5]
            print("OK" if np.array_equal(top, ref) else "WRONG")
    """)
    out = subprocess.run([sys.executable, "-c", probe], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-3000:]
    import torch
    if torch.cuda.is_available():
        assert out.stdout.strip() == "OK", out.stdout
    else:
        assert out.stdout.startswith("RAISED") and "CUDA" in out.stdout.upper(), out.stdout


# ---- ABI: the ctypes mirrors of the argument structs against the C header, field by field ----
def test_ctypes_struct_layouts_match_the_header(tmp_path):
    """Every field of BprArgs / P2PRouteArgs / P2PStepArgs (recsys_pytorch_b200/_lib.py) sits at the offset gcc gives
    the same-named member of the struct in include/b200rec.h, and the sizes agree - a silent drift here would hand the
    kernels garbage pointers.  The struct INTEGRATION.md section 3b shows a maintainer is checked the same way."""
    import ctypes as C
    import subprocess
    from recsys_pytorch_b200 import _lib
    structs = {"b200rec_bpr_args": _lib.BprArgs, "b200rec_p2p_route_args": _lib.P2PRouteArgs,
               "b200rec_p2p_step_args": _lib.P2PStepArgs}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "b200rec.h"', 'int main(void) {']
    for cname, cls in structs.items():
        lines.append('printf("%s sizeof %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('printf("%s %s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines.append("return 0; }")
    (tmp_path / "abi.c").write_text("\n".join(lines))
    exe = str(tmp_path / "abi")
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(tmp_path / "abi.c"), "-o", exe], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split("\n")
    got = {tuple(l.split()[:2]): int(l.split()[2]) for l in out if l.strip()}
    for cname, cls in structs.items():
        assert got[(cname, "sizeof")] == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert got[(cname, fname)] == getattr(cls, fname).offset, (cname, fname)
    # the hand-written binding in INTEGRATION.md (Option C): same fields, same order, same types
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    m = re.search(r"class BprArgs\(C\.Structure\):.*?_fields_ = (\[.*?\])\n", doc, re.S)
    assert m, "INTEGRATION.md lost its BprArgs binding"
    fields = eval(m.group(1), {"C": C})
    assert [(n, t) for n, t in fields] == [(n, t) for n, t in _lib.BprArgs._fields_]


def test_ctypes_prototypes_match_the_header():
    """Every prototype in include/b200rec.h against the ctypes signature _lib.py installs: same number of parameters,
    pointer where the header has a pointer, the same scalar class (int32 / int64 / uint64 / float) elsewhere, and the
    same return class."""
    import ctypes as C
    from recsys_pytorch_b200 import _lib
    hdr = re.sub(r"/\*.*?\*/", " ", open(os.path.join(ROOT, "include", "b200rec.h")).read(), flags=re.S)
    protos = re.findall(r"(?:^|\n)\s*((?:const\s+)?[A-Za-z_0-9]+\s*\**)\s*(b200rec_\w+)\s*\(([^;{]*?)\)\s*;", hdr)
    seen = set()

    def klass_of_c(text):
        text = text.strip()
        if "*" in text:
            return "ptr"
        base = text.replace("const", " ").split()[0] if text else "void"
        return {"int": "i32", "int32_t": "i32", "int64_t": "i64", "uint64_t": "u64", "float": "f32", "void": "void"}[base]

    def klass_of_ctypes(t):
        if t is C.c_void_p or t is C.c_char_p or (isinstance(t, type) and issubclass(t, C._Pointer)):
            return "ptr"
        return {C.c_int: "i32", C.c_int32: "i32", C.c_int64: "i64", C.c_uint64: "u64", C.c_float: "f32"}[t]

    for ret, name, params in protos:
        res, args = _lib._PROTOS[name]
        seen.add(name)
        plist = [] if params.strip() in ("", "void") else [p for p in params.split(",")]
        assert len(plist) == len(args), (name, len(plist), len(args))
        for k, (p, a) in enumerate(zip(plist, args)):
            assert klass_of_c(p) == klass_of_ctypes(a), (name, k, p.strip(), a)
        assert klass_of_c(ret) == klass_of_ctypes(res), (name, ret)
    assert seen == set(_lib._PROTOS), set(_lib._PROTOS) ^ seen


def test_integration_python_snippets_are_valid_python():
    """Every ```python block of INTEGRATION.md parses; the Option C block also executes up to its definitions (loads the
    library, declares the struct, installs the argtypes) from the repo root."""
    import textwrap
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = [textwrap.dedent(b) for b in re.findall(r"```python\n(.*?)```", doc, re.S)]
    assert len(blocks) >= 4
    for b in blocks:
        compile(b, "INTEGRATION.md", "exec")
    opt_c = [b for b in blocks if "class BprArgs" in b]
    assert len(opt_c) == 1
    cwd = os.getcwd()
    os.chdir(ROOT)
    try:
        ns = {}
        exec(compile(opt_c[0], "INTEGRATION.md", "exec"), ns)
    finally:
        os.chdir(cwd)
    assert callable(ns["fused_bpr_step"]) and callable(ns["fused_predict_topk"])
