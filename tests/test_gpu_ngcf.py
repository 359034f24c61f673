"""`-m gpu`: NGCF (csrc/ngcf.cu + recsys_pytorch_b200/ngcf.py) against tests/golden/ngcf_ml100k.npz -
outputs of the reference's own models/NGCF.py (propagated tables, loss, autograd gradients of every parameter, one Adam
step) - and the numpy oracle that tests/test_oracle_cpu.py pins to the same golden."""
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _model(golden, dev, **hp):
    import scipy.sparse as sp
    from recsys_pytorch_b200.ngcf import NGCF
    g, m1 = golden["ngcf_ml100k"], golden["ml100k"]
    nu, ni = int(m1["num_users"]), int(m1["num_items"])
    tr = sp.csr_matrix((np.ones(len(m1["train_indices"])), m1["train_indices"], m1["train_indptr"]), shape=(nu, ni))
    va = sp.csr_matrix((np.ones(len(m1["valid_indices"])), m1["valid_indices"], m1["valid_indptr"]), shape=(nu, ni))
    ds = types.SimpleNamespace(num_users=nu, num_items=ni, train_data=tr, valid_input=tr, valid_target=va,
                               protocol="holdout", dataname="ml-100k")
    h = {"emb_dim": 16, "num_layers": 2, "node_dropout": 0.0, "mess_dropout": 0.0, "split": False, "num_folds": 100,
         "graph_dir": "graph", "reg": 1e-4}
    h.update(hp)
    m = NGCF(ds, h, dev)
    m.load_parameters(g["U0"], g["V0"], {k: g[k] for k in g.files})
    m.Graph = m.getSparseGraph(tr)
    return m, g, ds, nu, ni


def test_ngcf_propagation_matches_reference(golden, dev):
    m, g, ds, nu, ni = _model(golden, dev)
    m.update_ngcf_embedding(training=True)
    np.testing.assert_allclose(m.user_embeddings.cpu().numpy()[:, :16], g["prop_U"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(m.item_embeddings.cpu().numpy()[:, :16], g["prop_V"], rtol=2e-5, atol=2e-6)
    assert float(m.out[:, 16:].abs().sum()) == 0.0


def test_ngcf_loss_gradients_and_adam_step_match_reference(golden, dev):
    m, g, ds, nu, ni = _model(golden, dev)
    m.train()
    u, i, j = (torch.from_numpy(g[k]).to(dev) for k in ("users", "pos", "neg"))
    assert abs(float(m.process_one_batch(u, i, j)) - float(g["loss"])) < 2e-6
    slot = torch.zeros(1, dtype=torch.float64, device=dev)
    triples = m.train_batch(u, i, j, loss_slot=slot)
    assert abs(slot.item() / 256 - float(g["loss"])) < 2e-6
    got = {n: gr.cpu().numpy() for n, _, gr in triples}
    np.testing.assert_allclose(got["E0"][:nu, :16], g["dU0"], rtol=3e-4, atol=3e-7)
    np.testing.assert_allclose(got["E0"][nu:, :16], g["dV0"], rtol=3e-4, atol=3e-7)
    for k in range(2):
        np.testing.assert_allclose(got["W_gc_%d" % k], g["dW_gc_%d" % k], rtol=3e-4, atol=3e-7)
        np.testing.assert_allclose(got["W_bi_%d" % k], g["dW_bi_%d" % k], rtol=3e-4, atol=3e-7)
        np.testing.assert_allclose(got["b_gc_%d" % k], g["db_gc_%d" % k].reshape(-1), rtol=3e-4, atol=3e-7)
        np.testing.assert_allclose(got["b_bi_%d" % k], g["db_bi_%d" % k].reshape(-1), rtol=3e-4, atol=3e-7)
    # one Adam step (NGCF.py:46): the first step moves every element with a non-zero gradient by ~lr, sign(grad)
    after = {"U": m.E0[:nu, :16].cpu().numpy(), "V": m.E0[nu:, :16].cpu().numpy()}
    for nm in ("U", "V"):
        bad = ~np.isclose(after[nm], g["adam_" + nm], rtol=2e-4, atol=2e-5)
        assert bad.mean() < 0.01 and np.abs(after[nm] - g["adam_" + nm]).max() < 2.5e-3
    for k in range(2):
        for nm, t in (("W_gc", m.W_gc), ("W_bi", m.W_bi), ("b_gc", m.b_gc), ("b_bi", m.b_bi)):
            ref = g["adam_%s_%d" % (nm, k)].reshape(t[k].shape)
            bad = ~np.isclose(t[k].cpu().numpy(), ref, rtol=2e-4, atol=2e-5)
            assert bad.mean() < 0.02 and np.abs(t[k].cpu().numpy() - ref).max() < 2.5e-3


def test_ngcf_message_dropout_and_plugin_fit(golden, dev):
    """mess_dropout > 0 (conf/NGCF.yaml default 0.1): masks come from the counter RNG - right rate, reproducible per step
    key, absent in evaluation mode; fit() runs through the shared sampler / evaluator path."""
    from recsys_pytorch_b200.evaluation import Evaluator
    m, g, ds, nu, ni = _model(golden, dev, mess_dropout=0.3)
    m.train(); m._step_key = 5
    m.update_ngcf_embedding(training=True)
    e1 = m._ego[0][:, :16].clone()
    frac = float((e1 == 0).float().mean())
    assert 0.27 < frac < 0.33
    m.update_ngcf_embedding(training=True)
    assert torch.equal(m._ego[0][:, :16], e1)                          # same (seed, step, layer) -> same mask
    m._step_key = 6
    m.update_ngcf_embedding(training=True)
    assert not torch.equal(m._ego[0][:, :16] == 0, e1 == 0)
    m.update_ngcf_embedding(training=False)
    assert float((m._ego[0][:, :16] == 0).float().mean()) < 0.01
    ev = Evaluator(ds.valid_input, ds.valid_target, protocol="holdout", ks=[10])
    ret = m.fit(ds, types.SimpleNamespace(num_epochs=2, batch_size=256, verbose=0, test_from=2, test_step=2), evaluator=ev)
    assert 0.0 <= float(ret["scores"]["NDCG@10"]) <= 1.0
    assert all(bool(torch.isfinite(t).all()) for _, t in m.parameter_tensors())
