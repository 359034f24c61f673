"""Multi-GPU BPR-MF with the exchange fused into the step kernel (csrc/p2p.cu).

Layout (BASELINE north_star "item table sharded by item-id range", made to scale): rank r holds the item rows
[item_bounds[r], item_bounds[r+1]) AND the user rows [user_bounds[r], user_bounds[r+1]) plus the CSR rows of those
users.  Nothing is replicated.  Per step every rank

    route    samples (i, j) for its own batch users (j from owner(i)'s range) and buckets the triples by owner(i)
             into an outbox in its own exported memory                                   [local kernel]
    barrier  one 4-byte all-reduce, stream-ordered                                        [the only NCCL call]
    compute  pulls the triples addressed to it from every outbox, loads each user row from its HOME rank through
             NVSwitch peer memory, updates its own item rows with vector atomics and stores the user row back
             to its home                                                                   [fused kernel, P2P ld/st]

Wire cost per triple: 4*ld bytes in + 4*ld bytes out on the owner's NVLink ports for the (W-1)/W remote triples -
point to point, never broadcast.  The replicated-user-table form of the north_star (dist.ItemShardedBPR: all-reduce of
the [B, ld] user-delta buffer) makes EVERY rank receive EVERY delta row, W/2 x more bytes per port.

The reference has no distributed code; the oracle is the single-device result on the triples the
ranks actually used (tests/test_p2p.py), and `pos`/`neg` may be given explicitly (fixed-triple parity mode, SURVEY
8(e) bullet 2: the negative may then live on another shard and is reached through the peer table as well).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, engine
from ._lib import F_USERS_UNIQUE, MAX_RANKS, PEER_HANDLE_BYTES, P2PRouteArgs, P2PStepArgs, check, current_stream, ptr


# ---------------------------------------------------------------------------------------------------------------
# peer memory
# ---------------------------------------------------------------------------------------------------------------
class _CudaArray:
    """Minimal __cuda_array_interface__ carrier: lets torch view a raw cudaMalloc'd block without copying."""

    def __init__(self, address, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(address), False),
                                         "version": 2, "strides": None}


class PeerArena:
    """One IPC-exportable device allocation (cudaMalloc through the C ABI - torch's caching allocator sub-allocates and
    cannot be exported block by block), carved into torch views."""

    ALIGN = 256

    def __init__(self, nbytes, device):
        self.device = torch.device(device)
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        with torch.cuda.device(self.device):
            check(_lib.lib().b200rec_peer_alloc(self.nbytes, C.byref(p)))
        self.ptr = int(p.value)
        self._carrier = _CudaArray(self.ptr, self.nbytes)
        self.buf = torch.as_tensor(self._carrier, device=self.device)
        assert self.buf.data_ptr() == self.ptr and self.buf.numel() == self.nbytes
        self._off = 0

    def carve(self, shape, dtype):
        """(tensor view, byte offset) of the next `shape` block."""
        n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        off = (self._off + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        assert off + n <= self.nbytes, "arena too small"
        self._off = off + n
        return self.buf[off:off + n].view(dtype).view(*shape), off

    def handle(self) -> bytes:
        h = (C.c_ubyte * PEER_HANDLE_BYTES)()
        check(_lib.lib().b200rec_peer_export(self.ptr, h))
        return bytes(h)

    def free(self):
        if self.ptr:
            self.buf = None
            check(_lib.lib().b200rec_peer_free(self.ptr))
            self.ptr = 0


def import_peer(handle: bytes) -> int:
    """Map another process's arena into this one (CUDA IPC, lazy peer access); returns its base address here."""
    h = (C.c_ubyte * PEER_HANDLE_BYTES).from_buffer_copy(handle)
    p = C.c_void_p()
    check(_lib.lib().b200rec_peer_import(h, C.byref(p)))
    return int(p.value)


# ---------------------------------------------------------------------------------------------------------------
# shard bounds
# ---------------------------------------------------------------------------------------------------------------
def uniform_bounds(n: int, world: int):
    return [(n * r) // world for r in range(world + 1)]


def balanced_item_bounds(item_counts, world: int):
    """Contiguous item-id ranges of (nearly) equal POSITIVE MASS: the owner of a triple is the shard of its positive
    item, and under a Zipf catalogue equal-width ranges leave the rank holding the head items with up to ~1.5x the
    work.  `item_counts` = global number of train interactions per item (1-D tensor / array).  Every shard keeps at
    least one item; bounds are identical on every rank because the histogram is."""
    c = torch.as_tensor(item_counts).double().cpu()
    n = int(c.numel())
    assert n >= world
    cum = torch.cumsum(c, 0)
    total = float(cum[-1]) if n else 0.0
    bounds = [0]
    for r in range(1, world):
        target = total * r / world
        b = int(torch.searchsorted(cum, torch.tensor([target], dtype=torch.float64)).item()) + 1 if total > 0 else (n * r) // world
        b = max(b, bounds[-1] + 1)
        b = min(b, n - (world - r))
        bounds.append(b)
    bounds.append(n)
    return bounds


def owner_from_bounds(ids, bounds):
    """Rank owning each id (numpy)."""
    return np.searchsorted(np.asarray(bounds[1:-1]), np.asarray(ids), side="right")


# ---------------------------------------------------------------------------------------------------------------
# the layout
# ---------------------------------------------------------------------------------------------------------------
class P2PShardedBPR:
    """One rank of the layout.  `train_local`: DeviceCSR whose rows are this rank's users (local ids) and whose
    columns are GLOBAL item ids.  Call `connect()` (all ranks, collectively) before the first step; single-process
    emulation of W ranks on one GPU: build W objects and call `P2PShardedBPR.connect_local(objs)`."""

    def __init__(self, num_users, num_items, d, train_local, rank, world, device, item_bounds, user_bounds=None,
                 lr=0.05, reg=0.0, init_std=0.01, seed=2020, max_batch=None):
        assert 1 <= world <= MAX_RANKS, "at most %d ranks" % MAX_RANKS
        self.num_users, self.num_items, self.d = int(num_users), int(num_items), int(d)
        self.rank, self.world, self.device = int(rank), int(world), torch.device(device)
        self.lr, self.reg, self.seed = float(lr), float(reg), int(seed)
        self.train = train_local
        self.item_bounds = [int(b) for b in item_bounds]
        self.user_bounds = [int(b) for b in (user_bounds if user_bounds is not None else uniform_bounds(num_users, world))]
        assert len(self.item_bounds) == world + 1 and self.item_bounds[0] == 0 and self.item_bounds[-1] == num_items
        assert len(self.user_bounds) == world + 1 and self.user_bounds[-1] == num_users
        self.ulo, self.uhi = self.user_bounds[rank], self.user_bounds[rank + 1]
        self.ilo, self.ihi = self.item_bounds[rank], self.item_bounds[rank + 1]
        nu, ni = self.uhi - self.ulo, self.ihi - self.ilo
        self.ld = engine.padded_dim(d)
        self.cap = int(max_batch if max_batch is not None else nu)
        W, cap, ld = world, self.cap, self.ld
        need = (6 * W * cap * 4 + 2 * 256 + 4 * (nu + ni) * ld + 16 * PeerArena.ALIGN)
        self.arena = PeerArena(need, self.device)
        # identical carve order on every rank: the outbox offsets depend on (world, cap) only
        self.ob, self.ob_off = [], []
        for _ in range(2):
            t, o = zip(*[self.arena.carve((W, cap), torch.int32) for _ in range(3)])
            c, oc = self.arena.carve((64,), torch.int32)
            c.zero_()
            self.ob.append(dict(u=t[0], i=t[1], j=t[2], cnt=c))
            self.ob_off.append(dict(u=o[0], i=o[1], j=o[2], cnt=oc))
        self.U, self.off_U = self.arena.carve((nu, ld), torch.float32)
        self.V, self.off_V = self.arena.carve((ni, ld), torch.float32)
        gu = torch.Generator(device=self.device); gu.manual_seed(seed * 7919 + 1 + rank)
        self.U.zero_(); self.V.zero_()
        if init_std > 0:
            self.U[:, :d].normal_(0.0, init_std, generator=gu)
            self.V[:, :d].normal_(0.0, init_std, generator=gu)
        self._bar = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.n_processed = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._step_args = None
        self._n = 0
        self._peer_base = None
        self._emulated = False

    # ---- wiring ------------------------------------------------------------------------------------
    def _meta(self):
        return dict(off_U=self.off_U, off_V=self.off_V, ob=self.ob_off, cap=self.cap)

    def _wire(self, bases, metas):
        """bases[s] = address of rank s's arena in THIS process; metas[s] = its offsets."""
        self._peer_base = list(bases)
        self._step_args = []
        for b in range(2):
            a = P2PStepArgs()
            a.world, a.rank, a.ld, a.d = self.world, self.rank, self.ld, self.d
            for k, v in enumerate(self.item_bounds):
                a.item_bounds[k] = v
            for s in range(self.world):
                m = metas[s]
                assert m["cap"] >= 1
                seg = self.rank * m["cap"] * 4                      # this rank's segment inside s's [W, cap] outbox
                a.U_peer[s] = bases[s] + m["off_U"]
                a.V_peer[s] = bases[s] + m["off_V"]
                a.in_u[s] = bases[s] + m["ob"][b]["u"] + seg
                a.in_i[s] = bases[s] + m["ob"][b]["i"] + seg
                a.in_j[s] = bases[s] + m["ob"][b]["j"] + seg
                a.in_cnt[s] = bases[s] + m["ob"][b]["cnt"] + 4 * self.rank
            self._step_args.append(a)
        self._peer_V = [bases[s] + metas[s]["off_V"] for s in range(self.world)]

    def connect(self):
        """Collective: exchange the IPC handles (host side, once) and map every peer arena."""
        if self.world == 1 or not dist.is_initialized():
            self._wire([self.arena.ptr], [self._meta()])
            return self
        mine = (self.arena.handle(), self._meta())
        allm = [None] * self.world
        dist.all_gather_object(allm, mine)
        bases = []
        with torch.cuda.device(self.device):
            for s in range(self.world):
                bases.append(self.arena.ptr if s == self.rank else import_peer(allm[s][0]))
        self._wire(bases, [m[1] for m in allm])
        self.barrier()
        return self

    @staticmethod
    def connect_local(ranks):
        """Single-process emulation: every 'peer' pointer is a plain local pointer of a sibling object."""
        bases = [r.arena.ptr for r in ranks]
        metas = [r._meta() for r in ranks]
        for r in ranks:
            r._wire(bases, metas)
            r._emulated = True
        return ranks

    # ---- one step ----------------------------------------------------------------------------------
    def route(self, users_local, step_key, pos=None, neg=None, dbg_pos=None, dbg_neg=None):
        """Sample (unless pos/neg are given: GLOBAL item ids) and bucket this rank's batch by owner(pos)."""
        users_local = _lib.require_cuda(users_local, "users_local", torch.int32)
        B = int(users_local.numel())
        assert B <= self.cap, "batch larger than max_batch"
        ob = self.ob[self._n & 1]
        a = P2PRouteArgs()
        a.users, a.B = ptr(users_local), B
        a.pos = ptr(_lib.require_cuda(pos, "pos", torch.int32)) if pos is not None else None
        a.neg = ptr(_lib.require_cuda(neg, "neg", torch.int32)) if neg is not None else None
        a.csr_indptr, a.csr_indices = ptr(self.train.indptr), ptr(self.train.indices)
        a.seed = self.seed & (2**64 - 1)
        a.step = (int(step_key) * self.world + self.rank) & (2**64 - 1)
        a.world, a.rank = self.world, self.rank
        for k, v in enumerate(self.item_bounds):
            a.item_bounds[k] = v
        a.out_u, a.out_i, a.out_j, a.out_cnt, a.cap = ptr(ob["u"]), ptr(ob["i"]), ptr(ob["j"]), ptr(ob["cnt"]), self.cap
        a.dbg_pos = ptr(dbg_pos) if dbg_pos is not None else None
        a.dbg_neg = ptr(dbg_neg) if dbg_neg is not None else None
        check(_lib.lib().b200rec_p2p_route(C.byref(a), current_stream()))

    def barrier(self):
        """Stream-ordered rendezvous of all ranks (a 4-byte all-reduce; nothing else crosses NCCL during training)."""
        if not self._emulated and self.world > 1 and dist.is_initialized():
            dist.all_reduce(self._bar)

    def compute(self, global_batch, loss_sum=None, users_unique=True):
        """Fused step over the triples every rank routed to this one (buffer of the current step)."""
        assert self._step_args is not None, "call connect() first"
        a = self._step_args[self._n & 1]
        a.lr, a.reg, a.inv_batch = self.lr, self.reg, 1.0 / float(global_batch)
        a.flags = F_USERS_UNIQUE if users_unique else 0
        a.loss_sum = ptr(_lib.require_cuda(loss_sum, "loss_sum", torch.float64)) if loss_sum is not None else None
        a.n_processed = ptr(self.n_processed)
        check(_lib.lib().b200rec_p2p_step(C.byref(a), current_stream()))
        self._n += 1

    def step(self, users_local, step_key, global_batch, loss_sum=None, users_unique=True, pos=None, neg=None):
        """route -> barrier -> compute.  `users_local`: int32 local row ids of this rank's batch (unique)."""
        self.route(users_local, step_key, pos, neg)
        self.barrier()
        self.compute(global_batch, loss_sum, users_unique)

    # ---- evaluation ---------------------------------------------------------------------------------
    def gather_items(self):
        """Full [num_items, ld] item table on this rank: W device-to-device copies out of the peers' shards (the
        one-time exchange of an evaluation, SURVEY 8(e) 'Scoring'); bracketed by barriers so no rank is training."""
        self.barrier()
        full = torch.empty((self.num_items, self.ld), dtype=torch.float32, device=self.device)
        row = self.ld * 4
        for s in range(self.world):
            lo, hi = self.item_bounds[s], self.item_bounds[s + 1]
            if hi > lo:
                check(_lib.lib().b200rec_peer_copy(full.data_ptr() + lo * row, self._peer_V[s], (hi - lo) * row,
                                                   current_stream()))
        self.barrier()
        return full

    def evaluate(self, eval_users_local, truth_local, ks, protocol="holdout", V_full=None):
        """Every rank scores its own users (local row ids; `truth_local` rows are local too) against the gathered item
        table; only the metric sums are all-reduced."""
        from .dist import evaluate_user_shard
        V_full = self.gather_items() if V_full is None else V_full
        return evaluate_user_shard(self.U, V_full, self.d, eval_users_local, self.train, truth_local, ks, protocol)

    def close(self):
        if self._peer_base is not None and not self._emulated:
            for s, b in enumerate(self._peer_base):
                if s != self.rank and b:
                    _lib.lib().b200rec_peer_close(b)
        self._peer_base = None
        self.arena.free()
