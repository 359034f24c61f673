"""B200-native LightGCN behind the reference's plugin surface (BASELINE configs[3]).

Mirrors `models/LightGCN.py:19-267`: constructor `LightGCN(dataset, hparams,
device)` reading `emb_dim, num_layers, node_dropout, split, num_folds, reg,
graph_dir`; `fit`, `forward`, `process_one_batch`, `predict`.  What runs:

* `getSparseGraph` (:228-267)  ->  `build_norm_adj`: the symmetric-normalised
  bipartite adjacency A_hat = D^-1/2 [[0,R],[R^T,0]] D^-1/2 as ONE device CSR (int64
  indptr, int32 cols, fp32 values), built from the train CSR with torch sort /
  bincount plumbing (no scipy dok/lil, no ./graph npz cache).
* `_lightgcn_embedding` (:174-202)  ->  `propagate`: L launches of the sm_100a CSR
  SpMM kernel, the stack+mean over layers fused as a running accumulation.
* loss / backward (:117-123 + autograd)  ->  the fused BPR kernel in SINK_GRAD mode on
  the PROPAGATED tables gives dL/d(out) as sparse rows in a dense buffer; because
  A_hat is symmetric, dL/dE_0 = mean_l A_hat^l dL/d(out) = `propagate` applied to that
  buffer - the backward pass reuses the forward kernel.
* optimiser (:45 `torch.optim.Adam(lr=1e-3)`, dense)  ->  the dense Adam sweep kernel
  ('sgd' selectable for trajectory parity tests).
* `predict` (:130-150)  ->  propagate once, then the shared scoring / top-K kernels.

`reg` is read but never used by the reference (:33 vs :122) - same here.
`node_dropout > 0` crashes in the reference (:182 vs :165); refused here.
`split/num_folds` only chunk the reference's SpMM for memory - ignored.
"""
from __future__ import annotations

import numpy as np
import torch

from . import engine
from ._lib import B200RecError, ECUDA, SCORE_EXACT, SCORE_TC, SINK_GRAD, SINK_NONE
from .generators import PairwiseGenerator
from .mf import MF, BaseModel, _hp


def build_norm_adj(train: engine.DeviceCSR):
    """models/LightGCN.py:228-267 on the device.  Returns (indptr int64 [N+1], cols int32, vals fp32), N = U+I."""
    nu, ni = train.shape
    dev = train.indptr.device
    nnz = train.nnz
    deg_u = (train.indptr[1:] - train.indptr[:-1])
    rows_u = torch.repeat_interleave(torch.arange(nu, device=dev), deg_u)
    items = train.indices.long()
    deg_i = torch.bincount(items, minlength=ni)
    du = deg_u.double().pow(-0.5); du[torch.isinf(du)] = 0.0       # :249-250
    di = deg_i.double().pow(-0.5); di[torch.isinf(di)] = 0.0
    # reference: norm_adj = D.dot(A).dot(D) in float32 sparse arithmetic -> (d_u * 1) * d_i, each product rounded to fp32
    vals_ui = ((du[rows_u].float() * 1.0) * di[items].float()).float()
    # user rows (cols offset by nu), then item rows = transpose sorted by (item, user)
    order = torch.argsort(items * nu + rows_u)
    t_rows = items[order]
    t_cols = rows_u[order]
    t_vals = ((di[t_rows].float() * 1.0) * du[t_cols].float()).float()
    indptr = torch.zeros(nu + ni + 1, dtype=torch.int64, device=dev)
    indptr[1:nu + 1] = torch.cumsum(deg_u, 0)
    indptr[nu + 1:] = nnz + torch.cumsum(deg_i, 0)
    cols = torch.cat([(items + nu).to(torch.int32), t_cols.to(torch.int32)]).contiguous()
    vals = torch.cat([vals_ui, t_vals]).contiguous()
    return indptr.contiguous(), cols, vals


class LightGCN(MF):
    def __init__(self, dataset, hparams, device):
        BaseModel.__init__(self)
        self.data_name = getattr(dataset, "dataname", "data")
        self.num_users = dataset.num_users
        self.num_items = dataset.num_items
        self.emb_dim = self.hidden_dim = int(hparams["emb_dim"])
        self.num_layers = int(hparams["num_layers"])
        self.node_dropout = float(_hp(hparams, "node_dropout", 0.0))
        if self.node_dropout > 0:
            raise NotImplementedError("node_dropout > 0 crashes in the reference (models/LightGCN.py:182 vs :165)")
        self.reg = float(_hp(hparams, "reg", 0.0))                  # read, never used (LightGCN.py:33,122)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise B200RecError(ECUDA, "recsys_pytorch_b200.LightGCN needs a CUDA device: there is no CPU path")
        self.optimizer_name = str(_hp(hparams, "optimizer", "adam")).lower()
        self.lr = float(_hp(hparams, "lr", 1e-3))
        self.sampler = str(_hp(hparams, "sampler", "device")).lower()
        self.gather = str(_hp(hparams, "gather", "ldg")).lower()
        self.score_algo = SCORE_TC if str(_hp(hparams, "score_algo", "exact")).lower() == "tc" else SCORE_EXACT
        self.seed = int(_hp(hparams, "seed", 2020))
        self.pointwise, self.lr_per_triple = False, None             # read by the shared MF.fit loop
        std = float(_hp(hparams, "init_std", 0.01))                 # nn.init.normal_(w, 0, 0.01), LightGCN.py:50-51

        N, d = self.num_users + self.num_items, self.emb_dim
        g = torch.Generator(device=self.device); g.manual_seed(self.seed)
        self.E0 = engine.alloc_table(N, d, self.device, std, g)     # cat([user_w, item_w]) as ONE table (:177)
        self.out = torch.zeros_like(self.E0)                        # mean over layers (:198-200)
        self._tmp = [torch.zeros_like(self.E0), torch.zeros_like(self.E0)]
        self.gOut = torch.zeros_like(self.E0)
        self.gE0 = torch.zeros_like(self.E0)
        self.user_embedding = _View(self.E0, 0, self.num_users, d, self)
        self.item_embedding = _View(self.E0, self.num_users, N, d, self)
        self._graph = None
        self._adam = None
        self._csr_cache = {}
        self._prop_fresh = False
        self.global_step = 0

    # ---- graph / propagation ----------------------------------------- #
    @property
    def Graph(self):
        return self._graph

    @Graph.setter
    def Graph(self, g):                                               # a new graph invalidates the propagated tables
        self._graph = g
        self._prop_fresh = False

    def getSparseGraph(self, rating_matrix):                         # LightGCN.py:228-267
        csr = rating_matrix if isinstance(rating_matrix, engine.DeviceCSR) else \
            engine.DeviceCSR.from_scipy(rating_matrix, self.device)
        return build_norm_adj(csr)

    def propagate(self, src, dst):
        """dst = mean_{l=0..L} A_hat^l src  (LightGCN.py:174-202).  2L reads/writes of an [N, ld] table."""
        if self._graph is None:
            raise RuntimeError("LightGCN: no graph yet - call fit(), or set model.Graph = model.getSparseGraph(train) "
                               "before propagating / predicting (models/LightGCN.py:38-40 builds it in fit too)")
        indptr, cols, vals = self._graph
        if getattr(self, "_plan_for", None) is not indptr:          # long rows (popular items) are split: plan per graph
            self._plan, self._plan_for = engine.spmm_plan(indptr), indptr
        L, d = self.num_layers, self.emb_dim
        s = 1.0 / (L + 1)
        if L == 0:
            dst.copy_(src)
            return dst
        cur = src
        for layer in range(L):
            nxt = self._tmp[layer & 1]
            engine.spmm_csr(indptr, cols, vals, cur, d, Y=nxt if layer + 1 < L else None, acc=dst, acc_scale=s,
                            acc_init=(layer == 0), plan=self._plan)
            cur = nxt
        return dst

    def update_lightgcn_embedding(self):                             # LightGCN.py:58-59
        self.propagate(self.E0, self.out)
        self.user_embeddings = self.out[:self.num_users]
        self.item_embeddings = self.out[self.num_users:]
        self._prop_fresh = True

    def score_tables(self):
        return self.out[:self.num_users], self.out[self.num_users:], self.emb_dim

    @property
    def U(self):
        return self.out[:self.num_users]

    @property
    def V(self):
        return self.out[self.num_users:]

    # ---- reference API ------------------------------------------------- #
    def forward(self, user_ids, item_ids):                           # LightGCN.py:61-66 (on the propagated tables)
        return engine.mf_forward(self.U, self.V, self.emb_dim, self._i32(user_ids, self.device),
                                 self._i32(item_ids, self.device))

    def process_one_batch(self, users, pos_items, neg_items):        # LightGCN.py:117-123, forward only
        self.update_lightgcn_embedding()
        users, pos, neg = (self._i32(t, self.device) for t in (users, pos_items, neg_items))
        loss = torch.zeros(1, dtype=torch.float64, device=self.device)
        engine.bpr_step(self.U, self.V, self.emb_dim, users, pos, neg, sink=SINK_NONE, loss_sum=loss)
        return (loss / users.numel()).to(torch.float32)[0]

    def train_batch(self, users, pos=None, neg=None, csr=None, step_key=0, users_unique=False, loss_slot=None):
        """zero_grad -> process_one_batch -> backward -> optimizer.step (LightGCN.py:77-84)."""
        users = self._i32(users, self.device)
        pos = self._i32(pos, self.device) if pos is not None else None
        neg = self._i32(neg, self.device) if neg is not None else None
        nu, d = self.num_users, self.emb_dim
        self.global_step += 1
        self.update_lightgcn_embedding()                             # forward propagation, per batch (:118)
        self.gOut.zero_()
        engine.bpr_step(self.U, self.V, d, users, pos, neg, csr=csr, sink=SINK_GRAD, gU=self.gOut[:nu],
                        gV=self.gOut[nu:], seed=self.seed, step=step_key, loss_sum=loss_slot)
        self.propagate(self.gOut, self.gE0)                          # backward through the propagation (A_hat symmetric)
        if self.optimizer_name == "adam":
            if self._adam is None:
                self._adam = [torch.zeros_like(self.E0), torch.zeros_like(self.E0), 0]
            self._adam[2] += 1
            engine.adam_dense(self.E0, self.gE0, self._adam[0], self._adam[1], self._adam[2], lr=self.lr)
        else:
            engine.sgd_dense(self.E0, self.gE0, self.lr)
        self._prop_fresh = False

    def fit(self, dataset, exp_config, evaluator=None, early_stop=None, loggers=None):   # LightGCN.py:68-115
        self.Graph = self.getSparseGraph(dataset.train_data)
        self.step_mode = "n/a"
        return MF.fit(self, dataset, exp_config, evaluator, early_stop, loggers)

    def predict_batch_users(self, user_ids):                         # LightGCN.py:125-128
        return engine.predict_dense(self.U, self.V, self.emb_dim, self._i32(user_ids, self.device), None)

    def predict(self, eval_users, eval_pos, test_batch_size):        # LightGCN.py:130-150
        self.update_lightgcn_embedding()
        return MF.predict(self, eval_users, eval_pos, test_batch_size)

    def predict_topk_device(self, eval_users, eval_pos, k, want_scores=False):
        if not self._prop_fresh:
            self.update_lightgcn_embedding()                         # re-propagate once per evaluation (:131)
        return MF.predict_topk_device(self, eval_users, eval_pos, k, want_scores)


class _View:
    """`.weight` view of rows [lo,hi) of the joint table (user_embedding / item_embedding of LightGCN.py:48-49)."""

    def __init__(self, store, lo, hi, d, owner=None):
        self.store = store[lo:hi]
        self.embedding_dim = d
        self._owner = owner

    @property
    def weight(self):
        return self.store[:, :self.embedding_dim]

    def load_weight(self, w):
        with torch.no_grad():
            self.store.zero_()
            self.store[:, :self.embedding_dim].copy_(torch.as_tensor(w, dtype=torch.float32))
        if self._owner is not None:                                   # E_0 changed: the propagated tables are stale
            self._owner._prop_fresh = False
