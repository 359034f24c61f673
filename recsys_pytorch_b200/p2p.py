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
    """One peer-mappable device allocation per rank, carved into torch views.

    symmetric=True (multi-process runs): the block comes from torch's symmetric-memory allocator (CUDA VMM: cuMemCreate
    + shareable handles, 2 MB granules); `rendezvous()` exchanges the handles and returns every rank's address in THIS
    process.  Measured reason (profiles/r02_p2p_notes.md): with legacy cudaIpc mappings a kernel that gathers from
    several peers at once runs 5-7x slower than the same pattern over same-process / VMM mappings.
    symmetric=False: plain cudaMalloc through the C ABI (single-process emulation; exportable with the legacy 64-byte
    cudaIpc handle via `handle()` / `import_peer`)."""

    ALIGN = 256

    def __init__(self, nbytes, device, symmetric=False):
        self.device = torch.device(device)
        self.nbytes = int(nbytes)
        self.symmetric = bool(symmetric)
        self._hdl = None
        if self.symmetric:
            import torch.distributed._symmetric_memory as symm_mem
            self.buf = symm_mem.empty(self.nbytes, dtype=torch.uint8, device=self.device)
            self.ptr = int(self.buf.data_ptr())
        else:
            p = C.c_void_p()
            with torch.cuda.device(self.device):
                check(_lib.lib().b200rec_peer_alloc(self.nbytes, C.byref(p)))
            self.ptr = int(p.value)
            self._carrier = _CudaArray(self.ptr, self.nbytes)
            self.buf = torch.as_tensor(self._carrier, device=self.device)
            assert self.buf.data_ptr() == self.ptr and self.buf.numel() == self.nbytes
        self._off = 0

    def rendezvous(self, group=None):
        """Collective (symmetric arenas): addresses of every rank's arena in this process."""
        import torch.distributed._symmetric_memory as symm_mem
        self._hdl = symm_mem.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        return [int(a) for a in self._hdl.buffer_ptrs]

    def carve(self, shape, dtype):
        """(tensor view, byte offset) of the next `shape` block."""
        n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        off = (self._off + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        assert off + n <= self.nbytes, "arena too small"
        self._off = off + n
        return self.buf[off:off + n].view(dtype).view(*shape), off

    def handle(self) -> bytes:
        assert not self.symmetric
        h = (C.c_ubyte * PEER_HANDLE_BYTES)()
        check(_lib.lib().b200rec_peer_export(self.ptr, h))
        return bytes(h)

    def free(self):
        if self.symmetric:
            self.buf, self._hdl, self.ptr = None, None, 0
        elif self.ptr:
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


def balanced_item_bounds(item_counts, world: int, head: int = 0):
    """Contiguous item-id ranges of (nearly) equal POSITIVE MASS: the owner of a triple is the shard of its positive
    item, and under a Zipf catalogue equal-width ranges leave the rank holding the head items with up to ~1.5x the
    work.  `item_counts` = global number of train interactions per item (1-D tensor / array).  Every shard keeps at
    least one item; bounds are identical on every rank because the histogram is.  `head` > 0: the first `head` ids are
    replicated everywhere and only [head, n) is split (bounds[0] == head)."""
    if head:
        return [head + b for b in balanced_item_bounds(torch.as_tensor(item_counts)[head:], world)]
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
    """Rank owning each (tail) id (numpy)."""
    return np.searchsorted(np.asarray(bounds[1:-1]), np.asarray(ids), side="right")


def relabel_columns(indptr, indices, shape, new_id):
    """Column ids of a CSR renumbered through `new_id` (old -> new) and re-sorted inside every row (the sampler's
    rejection test binary-searches the row).  Plain tensor code, any device."""
    dev = indices.device
    rows = torch.repeat_interleave(torch.arange(shape[0], device=dev), indptr[1:] - indptr[:-1])
    key, _ = torch.sort(rows * shape[1] + new_id[indices.long()])
    return (key - rows * shape[1]).to(torch.int32).contiguous()


def popularity_order(item_counts, head, seed=0):
    """(new_order, new_id): the `head` most popular items first, in RANDOM order inside the head; the tail keeps its
    original relative order.  Identical on every rank for identical (all-reduced) counts and seed."""
    counts = torch.as_tensor(item_counts)
    order = torch.argsort(counts, descending=True, stable=True)
    g = torch.Generator(device="cpu"); g.manual_seed(int(seed))
    head_items = order[:head][torch.randperm(int(head), generator=g).to(counts.device)] if head else order[:0]
    tail_items, _ = torch.sort(order[head:])
    new_order = torch.cat([head_items, tail_items])
    new_id = torch.empty_like(new_order)
    new_id[new_order] = torch.arange(new_order.numel(), device=counts.device)
    return new_order, new_id


def relabel_by_popularity(csrs, item_counts, head, seed=0):
    """Renumber the catalogue so that the `head` most popular items are the ids [0, head) - in RANDOM order inside the
    head, and the tail keeps its original relative order.  (Measured, profiles/r02_p2p_notes.md: laying the rows out in
    popularity ORDER makes the step kernel 35 % slower - the few dozen hottest rows then sit next to each other and their
    vector atomics pile onto the same L2 slices; scattered over the head they do not.)  Returns (relabelled DeviceCSRs
    with re-sorted rows, new_id_of_old int64 tensor, counts in the new order).  Pure data preparation, once per dataset;
    every rank computes the same permutation (same histogram, same seed)."""
    counts = torch.as_tensor(item_counts)
    new_order, new_id = popularity_order(counts, head, seed)
    out = []
    for csr in csrs:
        cols = relabel_columns(csr.indptr, csr.indices, csr.shape, new_id.to(csr.indices.device))
        out.append(engine.DeviceCSR(csr.indptr, cols, csr.shape))
    return out, new_id, counts[new_order]


# ---------------------------------------------------------------------------------------------------------------
# the layout
# ---------------------------------------------------------------------------------------------------------------
class P2PShardedBPR:
    """One rank of the layout.  `train_local`: DeviceCSR whose rows are this rank's users (local ids) and whose
    columns are GLOBAL item ids.  Call `connect()` (all ranks, collectively) before the first step; single-process
    emulation of W ranks on one GPU: build W objects and call `P2PShardedBPR.connect_local(objs)`."""

    def __init__(self, num_users, num_items, d, train_local, rank, world, device, item_bounds, user_bounds=None,
                 lr=0.05, reg=0.0, init_std=0.01, seed=2020, max_batch=None, head=0, head_reduce="sum"):
        """`head` > 0: the items [0, head) (the most popular ones after `relabel_by_popularity`) are REPLICATED on every
        rank; `item_bounds` then splits [head, num_items) only.  A triple whose positive is a head item is processed on
        its user's home rank without any NVLink traffic.  Every rank updates its replica IN PLACE during the step
        (Hogwild inside the rank, like the single-GPU kernel); between steps the per-rank differences to the pre-step
        snapshot are all-reduced (the same collective doubles as the step barrier) and every replica becomes
            snapshot + sum_r diff_r        head_reduce='sum'   exact data-parallel SGD (what the parity tests check)
            snapshot + mean_r diff_r       head_reduce='mean'  per-step parameter averaging (local SGD)
        'mean' is what large batches need: a hot row receives ~1e5 updates per rank and step, each rank's in-place
        trajectory saturates on its own, and ADDING W saturated displacements overshoots W-fold (observed: NaN at 8
        ranks, 8M-triple steps); averaging them does not.  Under a Zipf catalogue the head holds most positives, so
        most user rows never travel."""
        assert 1 <= world <= MAX_RANKS, "at most %d ranks" % MAX_RANKS
        self.head = int(head)
        self.num_users, self.num_items, self.d = int(num_users), int(num_items), int(d)
        self.rank, self.world, self.device = int(rank), int(world), torch.device(device)
        self.lr, self.reg, self.seed = float(lr), float(reg), int(seed)
        self.train = train_local
        self.item_bounds = [int(b) for b in item_bounds]
        self.user_bounds = [int(b) for b in (user_bounds if user_bounds is not None else uniform_bounds(num_users, world))]
        assert len(self.item_bounds) == world + 1 and self.item_bounds[0] == self.head and self.item_bounds[-1] == num_items
        assert len(self.user_bounds) == world + 1 and self.user_bounds[-1] == num_users
        self.ulo, self.uhi = self.user_bounds[rank], self.user_bounds[rank + 1]
        self.ilo, self.ihi = self.item_bounds[rank], self.item_bounds[rank + 1]
        nu, ni = self.uhi - self.ulo, self.ihi - self.ilo
        self.ld = engine.padded_dim(d)
        self.cap = int(max_batch if max_batch is not None else nu)
        W, cap, ld = world, self.cap, self.ld
        need = (6 * W * cap * 4 + 2 * 256 + 4 * (nu + ni) * ld + 16 * PeerArena.ALIGN)
        # multi-process: a symmetric (VMM) arena of the same size on every rank; B200REC_P2P_ARENA=ipc forces legacy IPC
        import os
        self._symm = (world > 1 and dist.is_initialized() and os.environ.get("B200REC_P2P_ARENA", "symm") != "ipc")
        self.arena = None
        if self._symm:
            t = torch.tensor([need], dtype=torch.int64, device=self.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)                # collective: every rank constructs its shard
            need = int(t.item())
            ok = torch.ones(1, dtype=torch.int32, device=self.device)
            try:
                self.arena = PeerArena(need, self.device, symmetric=True)
            except Exception:                                       # no VMM / symmetric-memory support here
                ok.zero_()
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)               # every rank takes the same path
            if int(ok.item()) == 0:
                self.arena, self._symm = None, False
        if self.arena is None:
            self.arena = PeerArena(need, self.device, symmetric=False)
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
        # replicated head rows: identical init on every rank (seed only), delta buffer zero
        self.Vh = self.dVh = None
        self.head_reduce = str(head_reduce)
        assert self.head_reduce in ("sum", "mean")
        if self.head:
            gh = torch.Generator(device=self.device); gh.manual_seed(seed * 104729 + 7)
            self.Vh = engine.alloc_table(self.head, d, self.device, init_std, gh)
            self._snap = torch.empty_like(self.Vh)
            self._diff = torch.empty_like(self.Vh)
            self._head_dirty = False
        self._bar = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.n_processed = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._step_args = None
        self._n = 0
        self._peer_base = None
        self._emulated = False
        self._side = torch.cuda.Stream(device=self.device)
        self._ev_sync, self._ev_route = torch.cuda.Event(), torch.cuda.Event()
        self._ev_sync.record(torch.cuda.current_stream(self.device))

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
            a.head = self.head
            a.Vh, a.dVh = (ptr(self.Vh), ptr(self.Vh)) if self.head else (None, None)     # replica updated in place
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
        mine = (None if self._symm else self.arena.handle(), self._meta())
        allm = [None] * self.world
        dist.all_gather_object(allm, mine)
        if self._symm:
            bases = self.arena.rendezvous()
            assert bases[self.rank] == self.arena.ptr
        else:
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
        a.head = self.head
        a.out_u, a.out_i, a.out_j, a.out_cnt, a.cap = ptr(ob["u"]), ptr(ob["i"]), ptr(ob["j"]), ptr(ob["cnt"]), self.cap
        a.dbg_pos = ptr(dbg_pos) if dbg_pos is not None else None
        a.dbg_neg = ptr(dbg_neg) if dbg_neg is not None else None
        check(_lib.lib().b200rec_p2p_route(C.byref(a), current_stream()))

    def barrier(self):
        """Stream-ordered rendezvous of all ranks (a 4-byte all-reduce)."""
        if not self._emulated and self.world > 1 and dist.is_initialized():
            dist.all_reduce(self._bar)

    def sync_head(self):
        """All-reduce the head delta of the previous step and apply it to this replica (Vh += sum_r dVh_r); doubles as
        the step barrier.  Without a head (or when no step is pending) it is the plain barrier."""
        if not self.head or not self._head_dirty:
            return self.barrier()
        engine.delta_diff(self.Vh, self._snap, self._diff, self._diff)          # what this rank's step did to its replica
        if not self._emulated and self.world > 1 and dist.is_initialized():
            dist.all_reduce(self._diff)
        engine.snap_apply(self.Vh, self._snap, self._diff, 1.0 if self.head_reduce == "sum" else 1.0 / self.world)
        self._head_dirty = False

    @staticmethod
    def sync_head_local(ranks):
        """Emulation of `sync_head` for sibling objects in one process."""
        if not ranks[0].head or not ranks[0]._head_dirty:
            return
        tot = torch.stack([r.Vh - r._snap for r in ranks]).sum(0)
        scale = 1.0 if ranks[0].head_reduce == "sum" else 1.0 / len(ranks)
        for r in ranks:
            engine.snap_apply(r.Vh, r._snap, tot, scale)
            r._head_dirty = False

    def compute(self, global_batch, loss_sum=None, users_unique=True):
        """Fused step over the triples every rank routed to this one (buffer of the current step)."""
        assert self._step_args is not None, "call connect() first"
        a = self._step_args[self._n & 1]
        a.lr, a.reg, a.inv_batch = self.lr, self.reg, 1.0 / float(global_batch)
        # chunk order: all homes round-robin over VMM (symmetric) mappings - the switch sees uniform all-to-all traffic;
        # one peer at a time over legacy IPC mappings, where round-robin collapses (profiles/r02_p2p_notes.md)
        order = _lib.F_P2P_ROUND_ROBIN if getattr(self, "_symm", False) else 0
        a.flags = (F_USERS_UNIQUE if users_unique else 0) | int(getattr(self, "extra_flags", order))
        a.loss_sum = ptr(_lib.require_cuda(loss_sum, "loss_sum", torch.float64)) if loss_sum is not None else None
        a.n_processed = ptr(self.n_processed)
        if self.head:
            self._snap.copy_(self.Vh)
        check(_lib.lib().b200rec_p2p_step(C.byref(a), current_stream()))
        self._n += 1
        if self.head:
            self._head_dirty = True

    def step(self, users_local, step_key, global_batch, loss_sum=None, users_unique=True, pos=None, neg=None):
        """route -> sync_head (all-reduce of the previous step's head delta = the barrier) -> compute.
        `users_local`: int32 local row ids of this rank's batch (unique).

        The route kernel runs on a side stream (`self.side`): sampling does not read the tables, so the route of step
        s overlaps the step kernel of step s-1 that is still executing on the main stream.  Its only dependency is the
        sync point of step s-1 (every peer has finished reading the outbox buffer it is about to overwrite).  The same
        triples, the same arithmetic - only the schedule differs from the serial order.  `users_local` (and pos/neg)
        must be ready for the side stream: produce them on `self.side` or before a synchronisation."""
        main = torch.cuda.current_stream(self.device)
        if self._emulated or self._side is None:
            self.route(users_local, step_key, pos, neg)
        else:
            self._side.wait_event(self._ev_sync)
            with torch.cuda.stream(self._side):
                self.route(users_local, step_key, pos, neg)
                self._ev_route.record(self._side)
            main.wait_event(self._ev_route)
        self.sync_head()
        if self._side is not None:
            self._ev_sync.record(main)
        self.compute(global_batch, loss_sum, users_unique)

    @property
    def side(self):
        """The stream the route kernel runs on (produce host-fed batches here)."""
        return self._side if self._side is not None else torch.cuda.current_stream(self.device)

    # ---- evaluation ---------------------------------------------------------------------------------
    def gather_items(self):
        """Full [num_items, ld] item table on this rank: W device-to-device copies out of the peers' shards (the
        one-time exchange of an evaluation, SURVEY 8(e) 'Scoring'); bracketed by barriers so no rank is training."""
        self.sync_head()
        self.barrier()
        full = torch.empty((self.num_items, self.ld), dtype=torch.float32, device=self.device)
        row = self.ld * 4
        if self.head:
            full[:self.head].copy_(self.Vh)
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
        if self._peer_base is not None and not self._emulated and not self._symm:
            for s, b in enumerate(self._peer_base):
                if s != self.rank and b:
                    _lib.lib().b200rec_peer_close(b)
        self._peer_base = None
        self.arena.free()


# ---------------------------------------------------------------------------------------------------------------
# bench.py --gpus N (N > 1), default layout
# ---------------------------------------------------------------------------------------------------------------
PER_GPU = dict(users=1_250_000, items=125_000)     # x8 = BASELINE configs[2]: 10M users x 1M items


def default_head(num_items, world):
    """Replicated head size used when none is given: 0 - pure range sharding.  A replicated head (e.g. 16,384 rows =
    8 MB at d=128, ~70 % of the positives of a Zipf(1) catalogue of 1M items) removes most NVLink traffic, but its
    synchronous exchange has no combiner that is right at every batch size: 'sum' overshoots once a hot row saturates
    inside one rank's step (NaN at 8 ranks x 1M triples), 'mean' divides the head's learning rate by the number of ranks
    when it does not (tests/test_p2p.py::test_p2p_same_data_ndcg_flat_across_world_sizes).  Opt-in: --head."""
    return 0


def build_rank(c, rank, world, dev, d=None, max_batch=None, lr=None, head=None, head_reduce="sum"):
    """This rank's shard of the weak-scaling dataset: `PER_GPU` users and items per GPU (N = 8 is cfg3), one global Zipf
    popularity order (item_seed), catalogue renumbered by popularity, balanced item bounds over the tail from the
    all-reduced item histogram."""
    from . import synthetic
    import os
    nu, ni = PER_GPU["users"] * world, int(os.environ.get("B200REC_PROBE_ITEMS", PER_GPU["items"])) * world
    if c.get("small"):
        nu, ni = 60_000 * world, 8_000 * world
    ub = uniform_bounds(nu, world)
    n_loc = ub[rank + 1] - ub[rank]
    train, target = synthetic.make_interactions(n_loc, ni, seed=c["seed"] + 1 + rank, device=dev, item_seed=c["seed"])
    hist = torch.bincount(train.indices.long(), minlength=ni).to(torch.int64)
    if world > 1 and dist.is_initialized():
        dist.all_reduce(hist)
    head = default_head(ni, world) if head is None else int(head)
    if head:
        (train, target), _, hist = relabel_by_popularity([train, target], hist, head, seed=c["seed"])
    ib = balanced_item_bounds(hist, world, head)
    m = P2PShardedBPR(nu, ni, d or c["d"], train, rank, world, dev, ib, ub, lr=lr if lr is not None else c["lr"],
                      reg=c["reg"], init_std=c["init_std"], seed=c["seed"], max_batch=max_batch or n_loc, head=head,
                      head_reduce=head_reduce)
    m.connect()
    return m, train, target, hist


def measure_p2p(c, rank, world, dev, timed_region, steps=8, warmup=3):
    """Compact measurement of the P2P layout at the weak-scaling shape (used by dist.bench_multi_gpu to report this
    layout beside the default one in the same run).  Returns a dict (rank 0) / None."""
    B_local = c["batch"]
    B_glob = B_local * world
    m, train, target, hist = build_rank(c, rank, world, dev, lr=c["lr_per_triple"] * B_glob, max_batch=B_local, head=0)
    n_loc = m.uhi - m.ulo
    g = torch.Generator(device=dev); g.manual_seed(c["seed"] + rank)
    perms = [torch.randperm(n_loc, device=dev, generator=g)[:B_local].to(torch.int32).contiguous() for _ in range(4)]
    loss = torch.zeros(1, dtype=torch.float64, device=dev)
    for s in range(warmup):
        m.step(perms[s % 4], s + 1, B_glob, loss_sum=loss)
    ms = timed_region(lambda s: m.step(perms[s % 4], warmup + s + 1, B_glob, loss_sum=loss), steps, world)
    n_ev = min(int(c["eval_users"]), n_loc)
    scores, n = m.evaluate(torch.arange(n_ev, dtype=torch.int32, device=dev), target, [c["eval_k"]])
    out = {"value": B_glob * steps / (ms * 1e-3), "unit": "triples/s", "steps": steps, "ms_per_step": ms / steps,
           "batch_triples": B_glob, "ndcg@%d" % c["eval_k"]: scores["NDCG@%d" % c["eval_k"]],
           "layout": "item table sharded by item-id range (north_star) + user table sharded by user-id range; user rows "
                     "through NVSwitch peer memory inside the fused step kernel; one 4-byte all-reduce per step (barrier)",
           "workload": "%dx%d d=%d" % (m.num_users, m.num_items, c["d"])}
    m.close()
    del train, target
    torch.cuda.empty_cache()
    return out if rank == 0 else None


def bench_p2p(args, c, rank, world, dev, timed_region, timed_under_load, hbm_gbs, peak_src):
    """Weak scaling, one fused P2P step per batch: every GPU brings 1.25M users, 125k items and 1M triples per step."""
    import json
    d, B_local = c["d"], c["batch"]
    B_glob = B_local * world
    lr = c["lr_per_triple"] * B_glob                      # the per-triple step does not depend on N (lr * 1/B_glob)
    m, train, target, hist = build_rank(c, rank, world, dev, lr=lr, max_batch=B_local, head=getattr(args, "head", None),
                                            head_reduce=getattr(args, "head_reduce", "sum"))
    n_loc = m.uhi - m.ulo
    head_mass = float(hist[:m.head].sum()) / float(hist.sum()) if m.head else 0.0
    g = torch.Generator(device=dev); g.manual_seed(c["seed"] + rank)
    perms = [torch.randperm(n_loc, device=dev, generator=g)[:B_local].to(torch.int32).contiguous() for _ in range(4)]
    loss = torch.zeros(1, dtype=torch.float64, device=dev)
    scratch = torch.zeros(1, dtype=torch.float64, device=dev)

    def step(s):
        m.step(perms[s % 4], s + 1, B_glob, loss_sum=loss)

    def step_idle(s):          # same kernels, same traffic, learning rate 0: the clock-sampling window only
        keep = m.lr
        m.lr = 0.0
        try:
            m.step(perms[s % 4], 100000 + s, B_glob, loss_sum=scratch)
        finally:
            m.lr = keep

    for s in range(args.warmup):
        step(s)
    counted = [0, 0]

    def step_counted(s):
        if s == 0:
            counted[0] = _lib.launch_count()
        step(args.warmup + s)
        if s == args.steps - 1:
            counted[1] = _lib.launch_count()
    reps = []
    ms, clk = timed_under_load(step_counted, step_idle, args.steps, world, dev.index or 0, rank=rank, pre_steps=400,
                               post_steps=120)
    reps.append(ms)
    for r in range(2):                                    # median of 3 timed repeats of the K steps
        reps.append(timed_region(lambda s: step(args.warmup + args.steps * (r + 1) + s), args.steps, world))
    ms = float(np.median(reps))
    launches = counted[1] - counted[0]
    # share of the triples each rank processed (load balance of the item bounds)
    share = torch.zeros(world, dtype=torch.float64, device=dev)
    share[rank] = float(m.n_processed.item())
    dist.all_reduce(share)
    share = (share / share.sum()).cpu().numpy()

    tot_loss = loss.clone()
    dist.all_reduce(tot_loss)
    # ---- e2e: user ids arrive from pinned host memory every step, the step's loss is read back every step ----
    host = [p.cpu().pin_memory() for p in perms]
    dbuf = [torch.empty_like(perms[0]) for _ in range(2)]
    hl = torch.zeros(1, dtype=torch.float64).pin_memory()

    def step_e2e(s):
        u = dbuf[s & 1]
        with torch.cuda.stream(m.side):                       # the batch arrives on the stream that routes it
            u.copy_(host[s % 4], non_blocking=True)
        loss.zero_()
        m.step(u, 5000 + s, B_glob, loss_sum=loss)
        hl.copy_(loss, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(hl[0])

    for s in range(2):
        step_e2e(s)
    ms_e2e = timed_region(step_e2e, args.steps, world)

    # ---- evaluation: one gather of the item shards, every rank scores its own users ----
    n_ev = min(int(c["eval_users"]), n_loc)
    ev_users = torch.arange(n_ev, dtype=torch.int32, device=dev)
    m.evaluate(ev_users, target, [c["eval_k"]])
    res = {}

    def ev_step(s):
        res["scores"], res["n"] = m.evaluate(ev_users, target, [c["eval_k"]])
    ms_ev = timed_region(ev_step, 3, world) / 3
    cfg5 = None
    if not getattr(args, "no_legs", False):
        from . import bench_legs
        m.close(); del train, target
        torch.cuda.empty_cache()
        cfg5 = bench_legs.cfg5_leg(dev, rank, world, float(getattr(args, "bf16_tf", 1590.0)), peak_src,
                                   small=bool(c.get("small")))
    if rank == 0:
        bpt = 24 * d + 8
        per_gpu = bpt * B_local / (ms / args.steps * 1e-3) / 1e9
        remote = (world - 1) / world * (1.0 - head_mass)                  # triples whose user row crosses NVLink
        # per GPU per direction per step: the rows this rank pulls + the rows its peers write back into it
        nvl_bytes = remote * B_local * (2 * 4 * m.ld + 12) + (2.0 * (world - 1) / world) * m.head * m.ld * 4
        nvl_gbs = nvl_bytes / (ms / args.steps * 1e-3) / 1e9
        out = {"metric": "BPR triples/sec (train)", "value": B_glob * args.steps / (ms * 1e-3), "unit": "triples/s",
               "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": {"workload": "BPRMF synthetic %dx%d d=%d, item-sharded + user-sharded over %d GPUs, user rows "
                                      "through NVSwitch peer memory inside the fused step%s"
                                      % (m.num_users, m.num_items, d, world,
                                         " (BASELINE configs[2])" if (m.num_users, m.num_items) == (10_000_000, 1_000_000) else ""),
                          "batch_triples": B_glob, "per_gpu_triples": B_local, "optimizer": "sgd+l2",
                          "lr_per_triple": c["lr_per_triple"], "reg": c["reg"], "parallelism": "p2p%d" % world,
                          "collective": ("user rows: P2P loads/stores inside the step kernel; replicated head of the "
                                         "catalogue (%d most popular items = %.1f %% of the positives), updated in place per "
                                         "rank: one all-reduce of its [%d, %d] fp32 difference per step (%d MiB, combined "
                                         "as the %s over ranks), which doubles as the step barrier"
                                         % (m.head, 100 * head_mass, m.head, m.ld, m.head * m.ld * 4 >> 20, m.head_reduce)) if m.head else
                                        "none on the data path: P2P loads/stores of user rows inside the step kernel; "
                                        "one 4-byte all-reduce per step as the barrier",
                          "head": m.head, "head_positive_mass": head_mass,
                          "item_bounds": "contiguous ranges of equal positive mass", "l2_policy": "inputs larger than L2",
                          "timing": "median of 3 repeats of the K steps", "repeats_ms": reps,
                          "triple_share_per_rank": [round(float(x), 4) for x in share]},
               "clocks": clk,
               "e2e": {"value": B_glob * args.steps / (ms_e2e * 1e-3), "unit": "triples/s",
                       "h2d_bytes_per_step": int(B_local * 4 * world), "d2h_bytes_per_step": 8 * world,
                       "ms_per_step": ms_e2e / args.steps,
                       "api": "recsys_pytorch_b200.p2p.P2PShardedBPR.step (user ids from pinned host memory, loss read "
                              "back every step, per rank)"},
               "gpu_launches": int(launches),
               "roofline": {"bound": "hbm", "achieved": per_gpu, "peak": hbm_gbs, "unit": "GB/s",
                            "frac": per_gpu / hbm_gbs, "traffic": None, "peak_source": peak_src,
                            "note": "per-GPU algorithmic bytes (24d+8 per triple) of the fused step over the WHOLE step "
                                    "time (route + barrier + step kernel)",
                            "nvlink": {"bytes_per_gpu_per_direction_per_step": nvl_bytes, "achieved_gbs": nvl_gbs,
                                       "model": "remote fraction x B_local x (row pulled + row written back by a peer + ids) "
                                                "+ ring all-reduce of the head delta",
                                       "peak_gbs": 770.0, "frac": nvl_gbs / 770.0,
                                       "peak_source": "B200_PROFILING.md measured peer copy per direction"}},
               "cpu_baseline": None,
               "eval": {"scored_pairs_per_sec": float(res["n"]) * m.num_items / (ms_ev * 1e-3), "ms": ms_ev,
                        "ndcg@%d" % c["eval_k"]: res["scores"]["NDCG@%d" % c["eval_k"]], "users": res["n"],
                        "k": c["eval_k"], "algo": "tc", "flops_per_pair": 2 * d,
                        "note": "whole evaluate() per rank: gather of the item shards through peer memory + scoring of "
                                "the rank's own users + metrics + metric all-reduce; the dataset grows with N, so NDCG is "
                                "not comparable across N (same-data parity across N: tests/test_p2p.py)"},
               "final_loss": float(tot_loss.item()) / (B_glob * (args.steps * 3 + args.warmup))}
        if cfg5 is not None:
            out["eval_cfg5"] = cfg5
        print(json.dumps(out))
    dist.barrier()
    if cfg5 is None:
        m.close()
    dist.destroy_process_group()
