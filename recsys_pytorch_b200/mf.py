"""B200-native BPR matrix factorisation behind the reference's model plugin surface.

Mirrors `models/BaseModel.py:3-14` and `models/MF.py:13-132` (constructor
`MF(dataset, hparams, device)`, `forward`, `fit`, `process_one_batch`,
`predict_batch_users`, `predict`) so it can be registered in the reference's
`models/__init__.py` and driven by its `main.py` unchanged - but every tensor op on
the path is a hand-written sm_100a kernel reached through the C ABI
(include/b200rec.h).  There is no PyTorch / CPU fallback: a CPU device raises.

Extra hparams (all optional).  With `conf/MF.yaml` as it is (hidden_dim, pointwise, loss_func only) the model trains
exactly like the reference: dense Adam(lr=1e-3) on N(0,1) tables - including the reference's slow start (NDCG@10 stays
near 0.014 on ml-100k for hundreds of epochs at batch 256, one triple per user and epoch; CPU simulation with
oracle/bpr_oracle.py).  The throughput path is opt-in:
    optimizer   'adam' (default = the reference's dense torch.optim.Adam(lr=1e-3), MF.py:30) |
                'sgd' (L2-regularised SGD, the fused one-kernel step of BASELINE north_star) |
                'lazy_adam' (row-wise Adam = torch.optim.SparseAdam semantics: only touched rows move)
    lr          step size.  For 'sgd' it multiplies the MEAN-reduced gradient (MF.py:105 `.mean()`), so one triple moves
                its rows by lr / batch_size; `lr_per_triple` sets that per-triple step directly (lr = lr_per_triple x
                batch_size at fit time; 0.05-0.1 is a sensible value, what bench.py uses)
    reg         per-occurrence L2 (the reference has none: Q1)
    step        'fused' (ONE kernel per batch, Hogwild inside a step) |
                'exact' (stage + apply: autograd's pre-step-weights semantics)
    sampler     'device' (default) | 'reference' (generators.py:168-224 call for call)
    gather      'ldg' (default: per-lane float4 loads, 2.2 G triples/s) | 'tma' (cp.async.bulk ring; the
                TMA engine issues ~1 bulk copy per 50 cycles, so 512-byte row gathers cap at 1.75 G triples/s)
    init_std    embedding init std (nn.Embedding default N(0,1), MF.py:23-24)
    score_algo  'exact' | 'tc'   scoring kernel used by predict_topk
    seed        sampler / permutation seed
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn

from . import engine
from ._lib import (B200RecError, ECUDA, F_TMA_GATHER, F_USERS_UNIQUE, SCORE_EXACT, SCORE_TC, SINK_GRAD, SINK_NONE, SINK_STAGE,
                   SINK_UPDATE)
from .generators import PairwiseGenerator


def _hp(hparams, key, default):
    try:
        return hparams[key]
    except Exception:
        return default


class _StepHandle:
    """Result of `MF.train_batch_async`: `.loss()` = mean -log sigmoid of that step (blocks until it is done)."""

    def __init__(self, event, host_slot, batch):
        self._event, self._host, self._batch = event, host_slot, batch

    def loss(self):
        self._event.synchronize()
        return float(self._host[0]) / self._batch


class BaseModel(nn.Module):
    """models/BaseModel.py:3-14 - the plugin base class (three no-op methods)."""

    def __init__(self):
        super().__init__()

    def forward(self, *input):
        pass

    def fit(self, *input):
        pass

    def predict(self, eval_users, eval_pos, test_batch_size):
        pass


class EmbeddingTable(nn.Module):
    """Stand-in for nn.Embedding (models/MF.py:23-24): `.weight` is an [n, d]
    Parameter viewing 16-byte-aligned padded storage `[n, ld]` that the kernels use."""

    def __init__(self, num_embeddings, embedding_dim, device, std=1.0, generator=None):
        super().__init__()
        self.num_embeddings, self.embedding_dim = int(num_embeddings), int(embedding_dim)
        self.store = engine.alloc_table(num_embeddings, embedding_dim, device, std, generator)
        self.weight = nn.Parameter(self.store[:, :embedding_dim], requires_grad=False)

    def _apply(self, fn, recurse=True):  # keep weight a view of the padded store across .to()/.cuda()
        new = fn(self.store)
        if new.data_ptr() != self.store.data_ptr():
            self.store = new.contiguous()
            self.weight = nn.Parameter(self.store[:, :self.embedding_dim], requires_grad=False)
        return self

    def load_weight(self, w):
        with torch.no_grad():
            self.store.zero_()
            self.store[:, :self.embedding_dim].copy_(torch.as_tensor(w, dtype=torch.float32))


class MF(BaseModel):
    def __init__(self, dataset, hparams, device):
        super().__init__()
        self.num_users = dataset.num_users
        self.num_items = dataset.num_items
        self.hidden_dim = int(hparams["hidden_dim"])
        self.pointwise = bool(_hp(hparams, "pointwise", False))          # MF.py:19
        self.loss_func_name = str(_hp(hparams, "loss_func", "ce")).lower()   # MF.py:21: 'ce' -> BCE-with-logits, else MSE
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise B200RecError(ECUDA, "recsys_pytorch_b200.MF needs a CUDA device: there is no CPU path")

        self.optimizer_name = str(_hp(hparams, "optimizer", "adam")).lower()
        self.lr = float(_hp(hparams, "lr", 1e-3 if "adam" in self.optimizer_name else 0.05))
        self.lr_per_triple = _hp(hparams, "lr_per_triple", None)
        self.reg = float(_hp(hparams, "reg", 0.0))
        self.step_mode = str(_hp(hparams, "step", "fused")).lower()
        self.sampler = str(_hp(hparams, "sampler", "device")).lower()
        self.gather = str(_hp(hparams, "gather", "ldg")).lower()
        self.score_algo = SCORE_TC if str(_hp(hparams, "score_algo", "exact")).lower() == "tc" else SCORE_EXACT
        self.seed = int(_hp(hparams, "seed", 2020))
        std = float(_hp(hparams, "init_std", 1.0))

        g = torch.Generator(device=self.device)
        g.manual_seed(self.seed)
        self.user_embedding = EmbeddingTable(self.num_users, self.hidden_dim, self.device, std, g)
        self.item_embedding = EmbeddingTable(self.num_items, self.hidden_dim, self.device, std, g)
        self._adam = None       # (m_u, v_u, m_i, v_i, g_u, g_i, t)
        self._stage = None
        self._csr_cache = {}
        self.global_step = 0

    # ---- tables -------------------------------------------------------- #
    @property
    def U(self):
        return self.user_embedding.store

    @property
    def V(self):
        return self.item_embedding.store

    def _flags(self, users_unique):
        from ._lib import GATHER_FLAGS
        f = GATHER_FLAGS.get(self.gather, 0)
        return f | (F_USERS_UNIQUE if users_unique else 0)

    @staticmethod
    def _i32(t, device):
        if isinstance(t, np.ndarray):
            t = torch.from_numpy(t)
        return t.to(device=device, dtype=torch.int32).contiguous()

    # ---- reference API -------------------------------------------------- #
    def embeddings(self, user_ids, item_ids):                      # MF.py:32-36
        return self.user_embedding.weight[user_ids.long()], self.item_embedding.weight[item_ids.long()]

    def forward(self, user_ids, item_ids):                         # MF.py:38-42
        return engine.mf_forward(self.U, self.V, self.hidden_dim, self._i32(user_ids, self.device),
                                 self._i32(item_ids, self.device))

    def process_one_batch(self, users, items, ratings):            # MF.py:99-107, forward only
        if self.pointwise:                                           # :101-102  loss_func(pos_ratings, ratings)
            users, items = self._i32(users, self.device), self._i32(items, self.device)
            r = torch.as_tensor(ratings, dtype=torch.float32).to(self.device).contiguous()
            loss = torch.zeros(1, dtype=torch.float64, device=self.device)
            engine.pointwise_step(self.U, self.V, self.hidden_dim, users, items, r, loss_func=self.loss_func_name,
                                  sink=SINK_NONE, loss_sum=loss)
            return (loss / users.numel()).to(torch.float32)[0]
        users, items, neg = (self._i32(t, self.device) for t in (users, items, ratings))
        loss = torch.zeros(1, dtype=torch.float64, device=self.device)
        engine.bpr_step(self.U, self.V, self.hidden_dim, users, items, neg, sink=SINK_NONE, loss_sum=loss)
        return (loss / users.numel()).to(torch.float32)[0]

    def _grad_buffers(self):
        if getattr(self, "_gbuf", None) is None:
            self._gbuf = (torch.zeros_like(self.U), torch.zeros_like(self.V))
        return self._gbuf

    def train_batch(self, users, pos=None, neg=None, csr=None, step_key=0, users_unique=False, loss_slot=None):
        """zero_grad -> process_one_batch -> backward -> optimizer.step (MF.py:64-68)."""
        users = self._i32(users, self.device)
        pos = self._i32(pos, self.device) if pos is not None else None
        neg = self._i32(neg, self.device) if neg is not None else None
        d, B = self.hidden_dim, users.numel()
        self.global_step += 1
        if self.optimizer_name == "adam":
            gU, gV = self._grad_buffers()
            gU.zero_(); gV.zero_()                                   # optimizer.zero_grad()
            engine.bpr_step(self.U, self.V, d, users, pos, neg, csr=csr, reg=self.reg, sink=SINK_GRAD, gU=gU, gV=gV,
                            seed=self.seed, step=step_key, loss_sum=loss_slot, flags=self._flags(False))
            if self._adam is None:
                self._adam = [torch.zeros_like(self.U), torch.zeros_like(self.U), torch.zeros_like(self.V),
                              torch.zeros_like(self.V), 0]
            self._adam[4] += 1
            t = self._adam[4]
            engine.adam_dense(self.U, gU, self._adam[0], self._adam[1], t, lr=self.lr)
            engine.adam_dense(self.V, gV, self._adam[2], self._adam[3], t, lr=self.lr)
        elif self.optimizer_name in ("lazy_adam", "sparse_adam"):
            # row-wise Adam (torch.optim.SparseAdam semantics): gradients into dense scratch rows,
            # then one claimed update per touched row; the scratch is zeroed row by row, never memset
            if getattr(self, "_lazy", None) is None:
                gU, gV = self._grad_buffers()
                gU.zero_(); gV.zero_()
                self._lazy = dict(mU=torch.zeros_like(self.U), vU=torch.zeros_like(self.U), mV=torch.zeros_like(self.V),
                                  vV=torch.zeros_like(self.V),
                                  sU=torch.zeros(self.U.shape[0], dtype=torch.int32, device=self.device),
                                  sV=torch.zeros(self.V.shape[0], dtype=torch.int32, device=self.device), t=0)
            lz = self._lazy
            gU, gV = self._grad_buffers()
            if pos is None or neg is None:
                pos, neg = engine.sample_triples(users, csr, self.seed, step_key)
            engine.bpr_step(self.U, self.V, d, users, pos, neg, reg=self.reg, sink=SINK_GRAD, gU=gU, gV=gV,
                            loss_sum=loss_slot, flags=self._flags(False))
            lz["t"] += 1
            engine.adam_rows(self.U, gU, lz["mU"], lz["vU"], lz["sU"], users, lz["t"], lr=self.lr)
            engine.adam_rows(self.V, gV, lz["mV"], lz["vV"], lz["sV"], pos, lz["t"], lr=self.lr)
            engine.adam_rows(self.V, gV, lz["mV"], lz["vV"], lz["sV"], neg, lz["t"], lr=self.lr)
        elif self.step_mode == "exact":
            if pos is None or neg is None:
                pos, neg = engine.sample_triples(users, csr, self.seed, step_key)
            if self._stage is None or self._stage.shape[0] < B:
                self._stage = torch.empty((B, 3, self.U.shape[1]), dtype=torch.float32, device=self.device)
            engine.bpr_step(self.U, self.V, d, users, pos, neg, lr=self.lr, reg=self.reg, sink=SINK_STAGE,
                            stage=self._stage, loss_sum=loss_slot, flags=self._flags(False))
            engine.bpr_apply(self.U, self.V, users, pos, neg, self._stage)
        else:
            engine.bpr_step(self.U, self.V, d, users, pos, neg, csr=csr, lr=self.lr, reg=self.reg, sink=SINK_UPDATE,
                            seed=self.seed, step=step_key, loss_sum=loss_slot, flags=self._flags(users_unique))

    def train_batch_pointwise(self, users, items, ratings, loss_slot=None):
        """Pointwise mode (MF.py:49-52,63-68,101-102): zero_grad -> loss_func(forward(u, i), ratings) -> backward ->
        optimizer.step as one fused kernel (+ the dense Adam sweep for the reference's optimiser)."""
        users, items = self._i32(users, self.device), self._i32(items, self.device)
        r = torch.as_tensor(ratings, dtype=torch.float32).to(self.device).contiguous()
        d = self.hidden_dim
        self.global_step += 1
        if self.optimizer_name == "adam":
            gU, gV = self._grad_buffers()
            gU.zero_(); gV.zero_()
            engine.pointwise_step(self.U, self.V, d, users, items, r, loss_func=self.loss_func_name, reg=self.reg,
                                  sink=SINK_GRAD, gU=gU, gV=gV, loss_sum=loss_slot)
            if self._adam is None:
                self._adam = [torch.zeros_like(self.U), torch.zeros_like(self.U), torch.zeros_like(self.V),
                              torch.zeros_like(self.V), 0]
            self._adam[4] += 1
            t = self._adam[4]
            engine.adam_dense(self.U, gU, self._adam[0], self._adam[1], t, lr=self.lr)
            engine.adam_dense(self.V, gV, self._adam[2], self._adam[3], t, lr=self.lr)
        elif self.optimizer_name == "sgd":
            engine.pointwise_step(self.U, self.V, d, users, items, r, loss_func=self.loss_func_name, lr=self.lr,
                                  reg=self.reg, sink=SINK_UPDATE, loss_sum=loss_slot)
        else:
            raise NotImplementedError("pointwise MF supports optimizer 'adam' (reference) and 'sgd'")

    def train_batch_async(self, users_host, csr=None, step_key=0, users_unique=False):
        """Pipelined form of `train_batch` for host-resident batches: `users_host` (pinned int32 CPU tensor) is
        copied on a side stream into one of two device buffers while the previous step still computes; the
        step's loss is copied back into pinned memory.  Returns a handle whose `.loss()` blocks until THAT step
        has finished - read it one step late to keep copy, compute and read-back overlapped."""
        st = getattr(self, "_pipe", None)
        if st is None or st["cap"] < users_host.numel():
            cap = int(users_host.numel())
            st = self._pipe = dict(
                cap=cap, n=0, copy=torch.cuda.Stream(device=self.device),
                users=[torch.empty(cap, dtype=torch.int32, device=self.device) for _ in range(2)],
                loss=[torch.zeros(1, dtype=torch.float64, device=self.device) for _ in range(2)],
                host=[torch.zeros(1, dtype=torch.float64).pin_memory() for _ in range(2)],
                h2d=[torch.cuda.Event() for _ in range(2)], done=[torch.cuda.Event() for _ in range(2)])
        k = st["n"] & 1
        st["n"] += 1
        B = int(users_host.numel())
        main = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(st["copy"]):
            st["copy"].wait_event(st["done"][k])                      # the step that last used this buffer is over
            st["users"][k][:B].copy_(users_host, non_blocking=True)
            st["h2d"][k].record(st["copy"])
        main.wait_event(st["h2d"][k])
        st["loss"][k].zero_()
        self.train_batch(st["users"][k][:B], csr=csr, step_key=step_key, users_unique=users_unique,
                         loss_slot=st["loss"][k])
        st["host"][k].copy_(st["loss"][k], non_blocking=True)
        st["done"][k].record(main)
        return _StepHandle(st["done"][k], st["host"][k], B)

    def fit(self, dataset, exp_config, evaluator=None, early_stop=None, loggers=None):   # MF.py:44-97
        train_matrix = dataset.train_data
        if getattr(self, "lr_per_triple", None) is not None and self.optimizer_name == "sgd":
            self.lr = float(self.lr_per_triple) * int(exp_config.batch_size)
        if getattr(self, "pointwise", False):
            return self._fit_pointwise(dataset, exp_config, evaluator, early_stop, loggers)
        gen = PairwiseGenerator(train_matrix, num_negatives=1, num_positives_per_user=1,
                                batch_size=exp_config.batch_size, shuffle=True, device=self.device,
                                sampler=self.sampler, seed=self.seed)
        num_batches = len(gen)
        scores = None
        for epoch in range(1, exp_config.num_epochs + 1):
            self.train()
            slots = torch.zeros(num_batches, dtype=torch.float64, device=self.device)
            sizes = []
            if gen.sampler == "device":
                for b, (users, key) in enumerate(gen.iter_device()):
                    self.train_batch(users, csr=gen.csr, step_key=key, users_unique=gen.users_unique,
                                     loss_slot=slots[b:b + 1])
                    sizes.append(users.numel())
            else:
                for b, (users, pos, neg) in enumerate(gen):
                    self.train_batch(users, pos, neg, users_unique=gen.users_unique, loss_slot=slots[b:b + 1])
                    sizes.append(users.numel())
            # epoch_loss = sum of per-batch mean losses (MF.py:70); one D2H per epoch
            epoch_loss = float((slots.cpu() / torch.tensor(sizes, dtype=torch.float64)).sum()) if sizes else 0.0
            if exp_config.verbose:
                print("epoch %3d loss = %.4f" % (epoch, epoch_loss))
            epoch_summary = {"loss": epoch_loss}
            if evaluator is not None and epoch >= exp_config.test_from and epoch % exp_config.test_step == 0:
                scores = evaluator.evaluate(self)
                epoch_summary.update(scores)
                if loggers is not None:
                    for logger in loggers:
                        logger.log_metrics(epoch_summary, epoch=epoch)
                if early_stop is not None:
                    is_update, should_stop = early_stop.step(scores, epoch)
                    if should_stop:
                        break
            elif loggers is not None:
                for logger in loggers:
                    logger.log_metrics(epoch_summary, epoch=epoch)
        best_score = early_stop.best_score if early_stop is not None else scores
        return {"scores": best_score}

    def _fit_pointwise(self, dataset, exp_config, evaluator, early_stop, loggers):       # MF.py:44-97, pointwise branch
        from .generators import PointwiseGenerator
        gen = PointwiseGenerator(dataset.train_data, return_rating=True, num_negatives=1,
                                 batch_size=exp_config.batch_size, shuffle=True, device=torch.device("cpu"))  # :49-52
        scores = None
        for epoch in range(1, exp_config.num_epochs + 1):
            self.train()
            slots = torch.zeros(len(gen), dtype=torch.float64, device=self.device)
            sizes = []
            for b, (users, items, ratings) in enumerate(gen):
                self.train_batch_pointwise(users, items, ratings, loss_slot=slots[b:b + 1])
                sizes.append(users.numel())
            epoch_loss = float((slots.cpu() / torch.tensor(sizes, dtype=torch.float64)).sum()) if sizes else 0.0
            if exp_config.verbose:
                print("epoch %3d loss = %.4f" % (epoch, epoch_loss))
            epoch_summary = {"loss": epoch_loss}
            if evaluator is not None and epoch >= exp_config.test_from and epoch % exp_config.test_step == 0:
                scores = evaluator.evaluate(self)
                epoch_summary.update(scores)
                if loggers is not None:
                    for logger in loggers:
                        logger.log_metrics(epoch_summary, epoch=epoch)
                if early_stop is not None:
                    is_update, should_stop = early_stop.step(scores, epoch)
                    if should_stop:
                        break
            elif loggers is not None:
                for logger in loggers:
                    logger.log_metrics(epoch_summary, epoch=epoch)
        best_score = early_stop.best_score if early_stop is not None else scores
        return {"scores": best_score}

    # ---- scoring -------------------------------------------------------- #
    def _device_csr(self, mat):
        if isinstance(mat, engine.DeviceCSR) or mat is None:
            return mat
        key = id(mat)
        hit = self._csr_cache.get(key)
        if hit is None or hit[0] is not mat:
            self._csr_cache = {key: (mat, engine.DeviceCSR.from_scipy(mat, self.device))}
            hit = self._csr_cache[key]
        return hit[1]

    def score_tables(self):
        """(U, V, d) scored by predict*; LightGCN overrides with the propagated tables."""
        return self.U, self.V, self.hidden_dim

    def predict_batch_users(self, user_ids):                       # MF.py:109-112
        U, V, d = self.score_tables()
        return engine.predict_dense(U, V, d, self._i32(user_ids, self.device), None)

    def predict(self, eval_users, eval_pos, test_batch_size):     # MF.py:114-132 (dense contract, small U only)
        U, V, d = self.score_tables()
        eval_users = np.asarray(eval_users)
        mask = self._device_csr(eval_pos)
        pred_matrix = np.zeros(eval_pos.shape)
        for st in range(0, len(eval_users), test_batch_size):
            bu = eval_users[st:st + test_batch_size]
            chunk = engine.predict_dense(U, V, d, self._i32(bu, self.device), mask)
            pred_matrix[bu] = chunk.cpu().numpy()
        if hasattr(eval_pos, "nonzero"):                           # MF.py:130 masks EVERY row of eval_pos, evaluated or not
            pred_matrix[eval_pos.nonzero()] = float("-inf")
        return pred_matrix

    def predict_topk_device(self, eval_users, eval_pos, k, want_scores=False):
        """Fused predict -> mask -> top-k (MF.py:109-132 + func.h:12-31) without the
        dense [U,I] matrix.  Returns CUDA (idx int32 [n,k], scores|None)."""
        U, V, d = self.score_tables()
        return engine.score_topk(U, V, d, self._i32(eval_users, self.device), self._device_csr(eval_pos), k,
                                 algo=self.score_algo, want_scores=want_scores)

    def predict_topk(self, eval_users, eval_pos, k):
        idx, _ = self.predict_topk_device(eval_users, eval_pos, k)
        return idx.cpu().numpy()
