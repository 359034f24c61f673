"""`-m "not gpu"`: the multi-GPU kernels of csrc/p2p.cu - `p2p_route_kernel` (sample + bucket by owner of the positive)
and the fused P2P step (`p2p_step_group_kernel<16>` = the default at d = 128, `<8>`, and the warp-per-row `p2p_step_kernel`
for every other width) - executed ON THE HOST by the SIMT emulator of tests/simt_host.py from their own source text.
W ranks live in one process exactly as in tests/test_p2p.py on the device: every "peer" pointer is a pointer into a
sibling rank's buffers, wired the way recsys_pytorch_b200/p2p.py::_wire does.  The oracle is the single-device numpy step
on the union of the triples the ranks used (oracle/bpr_oracle.py); the host mirror of the sampler predicts the draws."""
import ctypes as C

import numpy as np
import pytest

from oracle import bpr_oracle as O
from recsys_pytorch_b200 import _lib
from recsys_pytorch_b200._lib import F_P2P_ROUND_ROBIN, F_USERS_UNIQUE
from recsys_pytorch_b200.p2p import owner_from_bounds, uniform_bounds

pytestmark = pytest.mark.timeout(900)


@pytest.fixture(scope="module")
def simt(tmp_path_factory):
    from tests.simt_host import build_p2p
    return build_p2p(str(tmp_path_factory.mktemp("simt_p2p")))


class Rank:
    """Host twin of p2p.P2PShardedBPR: numpy buffers instead of the peer arena."""

    def __init__(self, rank, W, d, U_rows, V_rows, Vh_rows, csr, ib, ub, cap, head, lr, reg, seed):
        self.rank, self.W, self.d, self.head, self.cap = rank, W, d, head, cap
        self.lr, self.reg, self.seed = lr, reg, seed
        self.ib, self.ub = list(ib), list(ub)
        ld = (d + 3) // 4 * 4
        self.ld = ld
        pad = lambda a: np.ascontiguousarray(np.pad(a, ((0, 0), (0, ld - d))), np.float32)
        self.U, self.V = pad(U_rows), pad(V_rows)
        self.Vh = pad(Vh_rows) if head else np.zeros((1, ld), np.float32)
        self.indptr, self.indices = np.ascontiguousarray(csr[0], np.int64), np.ascontiguousarray(csr[1], np.int32)
        self.ob = {k: np.full((W, cap), -1, np.int32) for k in ("u", "i", "j")}
        self.cnt = np.zeros(64, np.int32)
        self.n_processed = np.zeros(1, np.int32)
        self.loss = np.zeros(1, np.float64)
        self.keep = []

    def route(self, simt, users_local, step_key, pos=None, neg=None):
        a = _lib.P2PRouteArgs()
        users_local = np.ascontiguousarray(users_local, np.int32)
        self.keep = [users_local]
        a.users, a.B = users_local.ctypes.data, len(users_local)
        for name, arr in (("pos", pos), ("neg", neg)):
            if arr is not None:
                self.keep.append(np.ascontiguousarray(arr, np.int32)); setattr(a, name, self.keep[-1].ctypes.data)
        a.csr_indptr, a.csr_indices = self.indptr.ctypes.data, self.indices.ctypes.data
        a.seed, a.step = self.seed, step_key * self.W + self.rank
        a.world, a.rank, a.head, a.cap = self.W, self.rank, self.head, self.cap
        for k, v in enumerate(self.ib):
            a.item_bounds[k] = v
        a.out_u, a.out_i, a.out_j = (self.ob[k].ctypes.data for k in "uij")
        a.out_cnt = self.cnt.ctypes.data
        self.dbg_pos, self.dbg_neg = np.full(len(users_local), -9, np.int32), np.full(len(users_local), -9, np.int32)
        a.dbg_pos, a.dbg_neg = self.dbg_pos.ctypes.data, self.dbg_neg.ctypes.data
        simt.emu_p2p_route(C.addressof(a), 2)

    def step(self, simt, ranks, global_batch, variant, flags=F_USERS_UNIQUE, want_loss=True, grid=2):
        a = _lib.P2PStepArgs()
        a.world, a.rank, a.ld, a.d, a.head = self.W, self.rank, self.ld, self.d, self.head
        for k, v in enumerate(self.ib):
            a.item_bounds[k] = v
        a.Vh = a.dVh = self.Vh.ctypes.data if self.head else None                       # replica updated in place
        for s, peer in enumerate(ranks):
            seg = self.rank * peer.cap * 4                                                # this rank's segment of s's outbox
            a.U_peer[s], a.V_peer[s] = peer.U.ctypes.data, peer.V.ctypes.data
            a.in_u[s], a.in_i[s], a.in_j[s] = (peer.ob[k].ctypes.data + seg for k in "uij")
            a.in_cnt[s] = peer.cnt.ctypes.data + 4 * self.rank
        a.lr, a.reg, a.inv_batch, a.flags = self.lr, self.reg, 1.0 / float(global_batch), flags
        a.loss_sum = self.loss.ctypes.data if want_loss else None
        a.n_processed = self.n_processed.ctypes.data
        assert simt.emu_p2p_step(C.addressof(a), variant, grid) == 0


def _make(W, nu, ni, d, seed, bounds=None, init=0.3, deg=12, head=0, lr=0.9, reg=0.01):
    rng = np.random.default_rng(seed)
    U0 = (rng.standard_normal((nu, d)) * init).astype(np.float32)
    V0 = (rng.standard_normal((ni, d)) * init).astype(np.float32)
    rows = [np.sort(rng.choice(ni, size=rng.integers(1, deg), replace=False)).astype(np.int32) for _ in range(nu)]
    indptr = np.zeros(nu + 1, np.int64); indptr[1:] = np.cumsum([len(r) for r in rows])
    indices = np.concatenate(rows)
    ib = bounds if bounds is not None else [head + b for b in uniform_bounds(ni - head, W)]
    ub = uniform_bounds(nu, W)
    ranks = []
    for r in range(W):
        lo, hi = ub[r], ub[r + 1]
        csr = (indptr[lo:hi + 1] - indptr[lo], indices[indptr[lo]:indptr[hi]])
        ranks.append(Rank(r, W, d, U0[lo:hi], V0[ib[r]:ib[r + 1]], V0[:head], csr, ib, ub, hi - lo, head, lr, reg, 11))
    return ranks, U0, V0, indptr, indices, ib, ub


def _sync_head(ranks, snap):
    """p2p.P2PShardedBPR.sync_head_local: every replica becomes snapshot + sum over ranks of what each did to its own."""
    if not ranks[0].head:
        return
    tot = sum(r.Vh - snap for r in ranks)
    for r in ranks:
        r.Vh[...] = snap + tot


def _tables(ranks, d):
    U = np.concatenate([r.U[:, :d] for r in ranks])
    V = np.concatenate(([ranks[0].Vh[:, :d]] if ranks[0].head else []) + [r.V[:, :d] for r in ranks])
    return U, V


@pytest.mark.parametrize("W", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("d,variant", [(128, 16), (128, 8), (128, 0), (50, 0), (200, 0)])
@pytest.mark.parametrize("head", [0, 80])
def test_given_triples_cross_shard_equals_single_device(simt, W, d, variant, head):
    """Fixed-triple parity mode (SURVEY 8(e) bullet 2; tests/test_p2p.py on the device): arbitrary (u, i, j) with i and j
    on DIFFERENT shards, no id shared between triples -> the fused P2P step is the exact step: tables == oracle."""
    if W == 3 and (d, variant) != (128, 16):
        pytest.skip("W = 3 is covered with the default kernel only (emulation time)")
    nu, ni = 320, 1100
    bounds = None if W == 1 else sorted({head, ni} | set(np.random.default_rng(W).choice(np.arange(head + 1, ni), W - 1, replace=False).tolist()))
    ranks, U0, V0, _, _, ib, ub = _make(W, nu, ni, d, seed=W * 100 + d, bounds=bounds, head=head)
    rng = np.random.default_rng(5)
    items = rng.permutation(ni)
    gu, gi, gj, k = [], [], [], 0
    for r in ranks:
        n_loc = ub[r.rank + 1] - ub[r.rank]
        B = n_loc * 3 // 4
        ul = rng.permutation(n_loc)[:B].astype(np.int32)
        pi, pj = items[k:k + B].astype(np.int32), items[k + B:k + 2 * B].astype(np.int32); k += 2 * B
        gu.append(ul + ub[r.rank]); gi.append(pi); gj.append(pj)
        r.route(simt, ul, 1, pos=pi, neg=pj)
    gu, gi, gj = np.concatenate(gu), np.concatenate(gi), np.concatenate(gj)
    Bg = len(gu)
    snap = ranks[0].Vh.copy()
    flags = F_USERS_UNIQUE | (F_P2P_ROUND_ROBIN if W % 2 == 0 else 0)                 # both visiting orders
    for r in ranks:
        r.step(simt, ranks, Bg, variant, flags=flags)
    assert sum(int(r.n_processed[0]) for r in ranks) == Bg
    _sync_head(ranks, snap)
    Ur, Vr, lref = O.sgd_step(U0, V0, gu, gi, gj, 0.9, 0.01)
    U, V = _tables(ranks, d)
    np.testing.assert_allclose(U, Ur, rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(V, Vr, rtol=2e-5, atol=2e-6)
    assert abs(sum(r.loss[0] for r in ranks) / Bg - float(lref)) < 2e-5 * max(1.0, float(lref))
    for r in ranks:
        assert not r.U[:, d:].any() and not r.V[:, d:].any()


@pytest.mark.parametrize("W", [1, 2, 4, 8])
@pytest.mark.parametrize("head", [0, 60])
def test_sampled_step_routing_and_oracle(simt, W, head):
    """On-device sampling + routing (tests/test_p2p.py::test_p2p_sampled_step_routing_and_oracle): the triples drawn are
    the host mirror's, every triple lands in the outbox segment of owner(pos) exactly once with its negative inside that
    owner's range (or the head), and two Hogwild steps stay within the second-order bound of the exact oracle step."""
    nu, ni, d = 512, 700, 128
    ranks, U0, V0, indptr, indices, ib, ub = _make(W, nu, ni, d, seed=W, head=head, lr=8.0, reg=0.0)
    rng = np.random.default_rng(1)
    Uc, Vc = U0, V0
    for step in (1, 2):
        gu, gi, gj = [], [], []
        for r in ranks:
            n_loc = ub[r.rank + 1] - ub[r.rank]
            B = n_loc - 7
            ul = rng.permutation(n_loc)[:B].astype(np.int32)
            r.route(simt, ul, step)
            dph, dnh = r.dbg_pos, r.dbg_neg
            cnt = r.cnt[:W]
            own = np.where(dph < head, r.rank, owner_from_bounds(dph, ib))           # head positives stay at home
            assert (dph >= 0).all() and (dnh >= 0).all()
            assert (cnt == np.bincount(own, minlength=W)).all()
            for dst in range(W):                                                      # segment content == triples owned by dst
                seg = np.stack([r.ob[k_][dst, :cnt[dst]] for k_ in "uij"], 1)
                exp = np.stack([ul[own == dst], dph[own == dst], dnh[own == dst]], 1)
                assert (seg[np.lexsort(seg.T[::-1])] == exp[np.lexsort(exp.T[::-1])]).all()
            assert ((dnh < head) | (owner_from_bounds(dnh, ib) == own)).all()
            for t in range(0, B, 5):                                                   # host mirror of the counter-RNG draws
                p_, n_ = O.sample_triple(11, step * W + r.rank, t, int(ul[t]) + ub[r.rank], indptr, indices, ni,
                                         item_bounds=ib, head=head, rank=r.rank)
                assert (p_, n_) == (int(dph[t]), int(dnh[t]))
            gu.append(ul + ub[r.rank]); gi.append(dph.copy()); gj.append(dnh.copy())
        gu, gi, gj = np.concatenate(gu), np.concatenate(gi), np.concatenate(gj)
        for t in range(0, len(gu), 3):                                                 # a negative is never a positive of the user
            assert gj[t] not in indices[indptr[gu[t]]:indptr[gu[t] + 1]]
        snap = ranks[0].Vh.copy()
        for r in ranks:
            r.step(simt, ranks, len(gu), 16, want_loss=False)
        _sync_head(ranks, snap)
        Ur, Vr, _ = O.sgd_step(Uc, Vc, gu, gi, gj, 8.0, 0.0)
        U, V = _tables(ranks, d)
        stepsz = max(np.abs(Ur - Uc).max(), np.abs(Vr - Vc).max())
        dev_ = max(np.abs(U - Ur).max(), np.abs(V - Vr).max())
        assert stepsz > 2e-3 and dev_ < 0.05 * stepsz, (stepsz, dev_)
        Uc, Vc = U, V


def test_hot_positive_prereduction_path(simt):
    """A chunk in which many triples share one positive (the Zipf head on its owner rank): the group kernel processes them
    first, sums their item update in registers and issues ONE reduction - same result as the oracle's accumulation."""
    W, nu, ni, d = 2, 256, 300, 128
    ranks, U0, V0, _, _, ib, ub = _make(W, nu, ni, d, seed=9)
    rng = np.random.default_rng(2)
    gu, gi, gj = [], [], []
    negs = rng.permutation(np.arange(10, ni))
    k = 0
    for r in ranks:
        B = 96
        ul = rng.permutation(ub[r.rank + 1] - ub[r.rank])[:B].astype(np.int32)
        pi = np.where(rng.random(B) < 0.6, 3, rng.integers(0, 10, B)).astype(np.int32)     # item 3: ~60 % of the positives
        pj = negs[k:k + B].astype(np.int32); k += B
        gu.append(ul + ub[r.rank]); gi.append(pi); gj.append(pj)
        r.route(simt, ul, 1, pos=pi, neg=pj)
    gu, gi, gj = np.concatenate(gu), np.concatenate(gi), np.concatenate(gj)
    for variant in (16, 8):
        rs, _, _, _, _, _, _ = _make(W, nu, ni, d, seed=9)
        for r, src in zip(rs, ranks):
            r.ob, r.cnt = {k_: v.copy() for k_, v in src.ob.items()}, src.cnt.copy()
        for r in rs:
            r.step(simt, rs, len(gu), variant, want_loss=False)
        # item rows are read before any update of the SAME chunk lands only for the hot row; the other positives collide
        # too (10 items), so compare at the Hogwild bound, and the hot row's update against the exact sum
        Ur, Vr, _ = O.sgd_step(U0, V0, gu, gi, gj, 0.9, 0.01)
        U, V = _tables(rs, d)
        stepsz = np.abs(Vr - V0).max()
        assert np.abs(V - Vr).max() < 0.05 * stepsz and np.abs(U - Ur).max() < 0.05 * stepsz
        assert np.abs(V[3] - Vr[3]).max() < 0.05 * np.abs(Vr[3] - V0[3]).max()
