"""B200-native NGCF behind the reference's plugin surface.

Mirrors `models/NGCF.py:21-290`: constructor `NGCF(dataset, hparams, device)` reading `emb_dim, num_layers, node_dropout,
mess_dropout, split, num_folds, reg, graph_dir`; `fit`, `forward`, `process_one_batch`, `predict`.  What runs:

* `getSparseGraph` (:240-290)          -> `lightgcn.build_norm_adj` (the same symmetric-normalised adjacency, on the device)
* `_ngcf_embedding` (:182-221)         -> per layer one CSR SpMM (`b200rec_spmm_csr`) + one fused layer kernel
  (`b200rec_ngcf_layer_forward`: both [d,d] transforms, leaky-relu, message dropout, L2 normalisation, running mean)
* loss / backward (:126-132 + autograd) -> the fused BPR kernel in SINK_GRAD mode gives dL/d(out) as sparse rows; each
  layer is then walked backwards with `b200rec_ngcf_layer_backward` + one SpMM (A_hat is symmetric)
* optimiser (:46 `Adam(lr=1e-3)` over embeddings AND the W/b dictionary) -> the dense Adam sweep kernel on every tensor
* `predict` (:141-163)                  -> propagate in eval mode (no dropout), then the shared scoring / top-K kernels

`reg` is read but never used by the reference (:37 vs :131) - same here.  `node_dropout > 0` calls a private method with
the wrong arity in the reference (:191 vs :175) and crashes; refused here.  `mess_dropout > 0` draws its mask from the
library's counter RNG (same distribution as F.dropout, different stream): trajectory parity is defined at 0.
`split/num_folds` only chunk the reference's SpMM for memory - ignored.
"""
from __future__ import annotations

import torch

from . import _lib, engine
from ._lib import B200RecError, ECUDA, SCORE_EXACT, SCORE_TC, SINK_GRAD, SINK_NONE, check, current_stream, ptr
from .lightgcn import build_norm_adj
from .mf import MF, BaseModel, _hp


class NGCF(MF):
    def __init__(self, dataset, hparams, device):
        BaseModel.__init__(self)
        self.data_name = getattr(dataset, "dataname", "data")
        self.num_users, self.num_items = dataset.num_users, dataset.num_items
        self.emb_dim = self.hidden_dim = int(hparams["emb_dim"])
        self.num_layers = int(hparams["num_layers"])
        self.node_dropout = float(_hp(hparams, "node_dropout", 0.0))
        if self.node_dropout > 0:
            raise NotImplementedError("node_dropout > 0 crashes in the reference (models/NGCF.py:191 vs :175)")
        self.mess_dropout = float(_hp(hparams, "mess_dropout", 0.0))
        self.reg = float(_hp(hparams, "reg", 0.0))                  # read, never used (NGCF.py:37,131)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise B200RecError(ECUDA, "recsys_pytorch_b200.NGCF needs a CUDA device: there is no CPU path")
        if self.emb_dim > 64:
            raise NotImplementedError("NGCF layer kernels hold the [d,d] transforms in shared memory: emb_dim <= 64")
        self.optimizer_name = str(_hp(hparams, "optimizer", "adam")).lower()
        self.lr = float(_hp(hparams, "lr", 1e-3))
        self.sampler = str(_hp(hparams, "sampler", "device")).lower()
        self.gather = "ldg"
        self.score_algo = SCORE_TC if str(_hp(hparams, "score_algo", "exact")).lower() == "tc" else SCORE_EXACT
        self.seed = int(_hp(hparams, "seed", 2020))
        self.pointwise = False
        self.lr_per_triple = None
        N, d, L = self.num_users + self.num_items, self.emb_dim, self.num_layers
        g = torch.Generator(device=self.device); g.manual_seed(self.seed)
        self.E0 = engine.alloc_table(N, d, self.device, float(_hp(hparams, "init_std", 0.01)), g)   # NGCF.py:54-55
        mk = lambda *shape: torch.randn(*shape, device=self.device, generator=g, dtype=torch.float32)  # nn.init.normal_ :60-64
        self.W_gc = [mk(d, d) for _ in range(L)]; self.b_gc = [mk(d) for _ in range(L)]
        self.W_bi = [mk(d, d) for _ in range(L)]; self.b_bi = [mk(d) for _ in range(L)]
        z = lambda: torch.zeros_like(self.E0)
        self.out, self.gOut, self.gE0 = z(), z(), z()
        self._ego = [z() for _ in range(L)]                         # ego_1 .. ego_L (ego_0 = E0)
        self._side = [z() for _ in range(L)]
        self._nrm = [torch.zeros(N, dtype=torch.float32, device=self.device) for _ in range(L)]
        self._gz, self._gside, self._ga, self._gb = z(), z(), z(), z()
        self._graph = None
        self._adam = None
        self._csr_cache = {}
        self._fresh = None                                           # 'train' | 'eval' | None: what `out` currently holds
        self.global_step = 0
        self._step_key = 0

    # ---- parameters ------------------------------------------------------------------------------------
    def parameter_tensors(self):
        """Every trainable tensor, in the order the reference's ParameterDict is built (NGCF.py:52-66)."""
        ps = [("E0", self.E0)]
        for k in range(self.num_layers):
            ps += [("W_gc_%d" % k, self.W_gc[k]), ("b_gc_%d" % k, self.b_gc[k]), ("W_bi_%d" % k, self.W_bi[k]),
                   ("b_bi_%d" % k, self.b_bi[k])]
        return ps

    def load_parameters(self, U0, V0, weights):
        """Copy reference parameters in (`weights`: dict name -> array as in NGCF.weight_dict)."""
        d = self.emb_dim
        with torch.no_grad():
            self.E0.zero_()
            self.E0[:self.num_users, :d] = torch.as_tensor(U0, dtype=torch.float32)
            self.E0[self.num_users:, :d] = torch.as_tensor(V0, dtype=torch.float32)
            for k in range(self.num_layers):
                self.W_gc[k].copy_(torch.as_tensor(weights["W_gc_%d" % k])); self.b_gc[k].copy_(torch.as_tensor(weights["b_gc_%d" % k]).reshape(-1))
                self.W_bi[k].copy_(torch.as_tensor(weights["W_bi_%d" % k])); self.b_bi[k].copy_(torch.as_tensor(weights["b_bi_%d" % k]).reshape(-1))
        self._fresh = None

    @property
    def Graph(self):
        return self._graph

    @Graph.setter
    def Graph(self, g):
        self._graph, self._fresh = g, None

    def getSparseGraph(self, rating_matrix):                         # NGCF.py:240-290
        csr = rating_matrix if isinstance(rating_matrix, engine.DeviceCSR) else \
            engine.DeviceCSR.from_scipy(rating_matrix, self.device)
        return build_norm_adj(csr)

    # ---- propagation -----------------------------------------------------------------------------------
    def _spmm(self, X, Y=None, acc=None, acc_scale=1.0):
        if self._graph is None:
            raise RuntimeError("NGCF: no graph yet - call fit(), or set model.Graph = model.getSparseGraph(train)")
        indptr, cols, vals = self._graph
        if getattr(self, "_plan_for", None) is not indptr:
            self._plan, self._plan_for = engine.spmm_plan(indptr), indptr
        return engine.spmm_csr(indptr, cols, vals, X, self.emb_dim, Y=Y, acc=acc, acc_scale=acc_scale, plan=self._plan)

    def update_ngcf_embedding(self, training=None):                 # NGCF.py:68-69 -> _ngcf_embedding :182-221
        training = self.training if training is None else training
        L, d = self.num_layers, self.emb_dim
        N, ld = self.E0.shape
        s = 1.0 / (L + 1)
        self.out.zero_()
        engine.sgd_dense(self.out, self.E0, -s)                      # out = E0 / (L+1): the first entry of `embs`
        p = self.mess_dropout if training else 0.0
        ego = self.E0
        for k in range(L):
            self._spmm(ego, Y=self._side[k])
            check(_lib.lib().b200rec_ngcf_layer_forward(
                ptr(ego), ptr(self._side[k]), ptr(self.W_gc[k]), ptr(self.b_gc[k]), ptr(self.W_bi[k]), ptr(self.b_bi[k]),
                N, ld, d, k, float(p), self.seed, self._step_key, ptr(self._ego[k]), ptr(self._nrm[k]), ptr(self.out),
                float(s), current_stream()))
            ego = self._ego[k]
        self.user_embeddings = self.out[:self.num_users]
        self.item_embeddings = self.out[self.num_users:]
        self._fresh = "train" if (training and p > 0) else "eval"

    def backward_from_gout(self):
        """dL/d(out) in self.gOut -> gradients of every parameter: returns [(name, param, grad)]."""
        L, d = self.num_layers, self.emb_dim
        N, ld = self.E0.shape
        s = 1.0 / (L + 1)
        p = self.mess_dropout if self._fresh == "train" else 0.0
        grads = {}
        gnext = None
        for k in range(L - 1, -1, -1):
            dWg, dWb, db = torch.zeros_like(self.W_gc[k]), torch.zeros_like(self.W_bi[k]), torch.zeros_like(self.b_gc[k])
            ego_in = self.E0 if k == 0 else self._ego[k - 1]
            gego = self._ga if gnext is not self._ga else self._gb
            check(_lib.lib().b200rec_ngcf_layer_backward(
                ptr(self.gOut), ptr(gnext) if gnext is not None else None, ptr(ego_in), ptr(self._side[k]), ptr(self._ego[k]),
                ptr(self._nrm[k]), ptr(self.W_gc[k]), ptr(self.W_bi[k]), N, ld, d, k, float(p), self.seed, self._step_key,
                float(s), ptr(self._gz), ptr(self._gside), ptr(gego), ptr(dWg), ptr(dWb), ptr(db), current_stream()))
            self._spmm(self._gside, acc=gego, acc_scale=1.0)          # d(ego) += A_hat d(side)   (A_hat symmetric)
            gnext = gego
            grads["W_gc_%d" % k], grads["W_bi_%d" % k] = dWg, dWb
            grads["b_gc_%d" % k] = grads["b_bi_%d" % k] = db
        self.gE0.zero_()
        engine.sgd_dense(self.gE0, self.gOut, -s)                     # direct path of E0 into the mean
        if gnext is not None:
            engine.sgd_dense(self.gE0, gnext, -1.0)
        grads["E0"] = self.gE0
        return [(n, t, grads[n]) for n, t in self.parameter_tensors()]

    # ---- tables the shared MF paths use ----------------------------------------------------------------------
    def score_tables(self):
        return self.out[:self.num_users], self.out[self.num_users:], self.emb_dim

    @property
    def U(self):
        return self.out[:self.num_users]

    @property
    def V(self):
        return self.out[self.num_users:]

    # ---- reference API -----------------------------------------------------------------------------------
    def forward(self, user_ids, item_ids):                           # NGCF.py:71-76 (on the propagated tables)
        return engine.mf_forward(self.U, self.V, self.emb_dim, self._i32(user_ids, self.device),
                                 self._i32(item_ids, self.device))

    def process_one_batch(self, users, pos_items, neg_items):        # NGCF.py:126-132, forward only
        self.update_ngcf_embedding()
        users, pos, neg = (self._i32(t, self.device) for t in (users, pos_items, neg_items))
        loss = torch.zeros(1, dtype=torch.float64, device=self.device)
        engine.bpr_step(self.U, self.V, self.emb_dim, users, pos, neg, sink=SINK_NONE, loss_sum=loss)
        return (loss / users.numel()).to(torch.float32)[0]

    def train_batch(self, users, pos=None, neg=None, csr=None, step_key=0, users_unique=False, loss_slot=None):
        """zero_grad -> process_one_batch -> backward -> optimizer.step (NGCF.py:93-99)."""
        users = self._i32(users, self.device)
        pos = self._i32(pos, self.device) if pos is not None else None
        neg = self._i32(neg, self.device) if neg is not None else None
        nu, d = self.num_users, self.emb_dim
        self.global_step += 1
        self._step_key = self.global_step
        self.update_ngcf_embedding(training=True)
        self.gOut.zero_()
        engine.bpr_step(self.U, self.V, d, users, pos, neg, csr=csr, sink=SINK_GRAD, gU=self.gOut[:nu], gV=self.gOut[nu:],
                        seed=self.seed, step=step_key, loss_sum=loss_slot)
        triples = self.backward_from_gout()
        if self.optimizer_name == "adam":
            if self._adam is None:
                self._adam = {n: (torch.zeros_like(t), torch.zeros_like(t)) for n, t, _ in triples}
                self._adam_t = 0
            self._adam_t += 1
            for n, t, g in triples:
                engine.adam_dense(t, g, self._adam[n][0], self._adam[n][1], self._adam_t, lr=self.lr)
        else:
            for n, t, g in triples:
                engine.sgd_dense(t, g, self.lr)
        self._fresh = None
        return triples

    def fit(self, dataset, exp_config, evaluator=None, early_stop=None, loggers=None):   # NGCF.py:78-124
        self.Graph = self.getSparseGraph(dataset.train_data)
        self.step_mode = "n/a"
        return MF.fit(self, dataset, exp_config, evaluator, early_stop, loggers)

    def predict_batch_users(self, user_ids):                         # NGCF.py:134-137
        return engine.predict_dense(self.U, self.V, self.emb_dim, self._i32(user_ids, self.device), None)

    def predict(self, eval_users, eval_pos, test_batch_size):        # NGCF.py:139-163
        self.update_ngcf_embedding(training=False)
        return MF.predict(self, eval_users, eval_pos, test_batch_size)

    def predict_topk_device(self, eval_users, eval_pos, k, want_scores=False):
        if self._fresh != "eval":
            self.update_ngcf_embedding(training=False)               # re-propagate once per evaluation, no dropout (:140)
        return MF.predict_topk_device(self, eval_users, eval_pos, k, want_scores)
