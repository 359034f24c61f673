"""Randomised exactness campaigns on the host SIMT emulator (tests/simt_host.py): the kernels' own source text against the
oracle on random shapes.  Scoring: d, k, catalogue size, norm spread, ties, zero rows, heavy masks, anti-aligned users,
item-split counts, and (tensor-core path) an adversarial accumulator in half of the cases.  P2P step: world size up to 16,
every row width, replicated head, random shard bounds, empty batches, the three visiting orders, fixed triples across shards.

    python tools/fuzz_on_emulator.py tc    600 1000     # seconds, first seed  (round 2: 200 cases, 0 mismatches)
    python tools/fuzz_on_emulator.py exact 300 5000     #                      (round 2: 105 cases, 0 mismatches)
    python tools/fuzz_on_emulator.py p2p   300 100      #                      (round 2: 843 cases, 0 mismatches)
"""
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import simt_host  # noqa: E402
from tests.util import OracleC  # noqa: E402

P = lambda a: a.ctypes.data if a is not None else None


def fuzz_tc(seconds, seed):
    from tests.test_tc_path_on_simt_host import tc_score_topk
    d_ = tempfile.mkdtemp()
    simt = (simt_host.build_tc(d_), simt_host.build_score(d_))
    oc = OracleC(os.path.join(ROOT, "oracle", "liboracle.so"))
    t_end, n, bad = time.time() + seconds, 0, 0
    while time.time() < t_end:
        rng = np.random.default_rng(seed); seed += 1
        d = int(rng.choice([8, 16, 20, 32, 48, 64, 100, 128, 136, 200, 256])); k = int(rng.choice([1, 2, 5, 10, 20, 32, 33, 50, 100, 128]))
        ni = int(rng.integers(max(k + 5, 65), 5000)); nu = int(rng.integers(1, 300))
        ld = (d + 3) // 4 * 4
        sig, usig = float(rng.choice([0, 0.3, 0.8, 1.5])), float(rng.choice([0, 0.5]))
        U = np.zeros((nu, ld), np.float32); V = np.zeros((ni, ld), np.float32)
        U[:, :d] = rng.standard_normal((nu, d)) * 0.1 * np.exp(usig * rng.standard_normal((nu, 1)))
        V[:, :d] = rng.standard_normal((ni, d)) * 0.1 * np.exp(sig * rng.standard_normal((ni, 1)))
        if rng.random() < 0.3:                                               # ties and exact zeros
            V[:, :d] = np.round(V[:, :d] * 16) / 16; U[:, :d] = np.round(U[:, :d] * 16) / 16
        if rng.random() < 0.2:
            V[rng.integers(0, ni, ni // 10)] = 0                             # zero rows
        if rng.random() < 0.2:
            U[:, 0] -= 0.5; V[:, 0] += np.abs(V[:, 0])                       # users anti-aligned with the norm direction
        mask = None
        if rng.random() < 0.8:
            hi = int(rng.choice([5, 40, min(ni - 1, 400)]))
            rows = [np.sort(rng.choice(ni, int(rng.integers(0, hi)), replace=False)).astype(np.int32) for _ in range(nu)]
            if rng.random() < 0.3:
                rows[0] = np.sort(rng.choice(ni, ni - int(rng.integers(0, k + 3)), replace=False)).astype(np.int32)
            ip = np.zeros(nu + 1, np.int64); ip[1:] = np.cumsum([len(r) for r in rows])
            mask = (ip, np.concatenate(rows).astype(np.int32) if ip[-1] else np.zeros(0, np.int32))
        users = rng.permutation(nu).astype(np.int32)
        adv = np.random.default_rng(seed) if rng.random() < 0.5 else None
        idx, sc, _ = tc_score_topk(simt, U, V, ld, d, users, ni, mask, k, adversary=adv)
        ri, rs = oc.score_topk(U, V, d, users, ni, mask[0] if mask else None, mask[1] if mask else None, k)
        n += 1
        if not (np.array_equal(idx, ri) and np.array_equal(sc, rs)):
            bad += 1
            print("MISMATCH seed", seed - 1, dict(d=d, k=k, ni=ni, nu=nu, sig=sig), flush=True)
    print("tensor-core path: %d cases, %d mismatches, next seed %d" % (n, bad, seed))


def fuzz_exact(seconds, seed):
    simt = simt_host.build_score(tempfile.mkdtemp())
    oc = OracleC(os.path.join(ROOT, "oracle", "liboracle.so"))
    t_end, n, bad = time.time() + seconds, 0, 0
    while time.time() < t_end:
        rng = np.random.default_rng(seed); seed += 1
        d = int(rng.integers(1, 70)); ni = int(rng.integers(1, 500)); k = int(rng.integers(1, min(ni, 330) + 1)); nu = int(rng.integers(1, 80))
        ld = (d + 3) // 4 * 4 + int(rng.choice([0, 4]))
        U = np.zeros((nu, ld), np.float32); V = np.zeros((ni, ld), np.float32)
        U[:, :d] = rng.standard_normal((nu, d)); V[:, :d] = rng.standard_normal((ni, d))
        if rng.random() < 0.4:
            U[:, :d] = np.round(U[:, :d]); V[:, :d] = np.round(V[:, :d])
        mask = None
        if rng.random() < 0.7:
            rows = [np.sort(rng.choice(ni, int(rng.integers(0, ni + 1)) if rng.random() < 0.2 else int(rng.integers(0, min(ni, 20) + 1)),
                                       replace=False)).astype(np.int32) for _ in range(nu)]
            ip = np.zeros(nu + 1, np.int64); ip[1:] = np.cumsum([len(r) for r in rows])
            mask = (ip, np.concatenate(rows).astype(np.int32) if ip[-1] else np.zeros(1, np.int32))
        users = rng.integers(0, nu, int(rng.integers(1, nu + 1))).astype(np.int32)
        nr, splits = len(users), int(rng.choice([1, 1, 2, 3, 7]))
        idx = np.zeros((nr, k), np.int32); sc = np.zeros((nr, k), np.float32)
        simt.emu_score_topk_exact(P(U), P(V), ld, d, P(users), nr, ni, P(mask[0]) if mask else None, P(mask[1]) if mask else None, k,
                                  P(idx), P(sc), None, splits)
        ri, rs = oc.score_topk(U, V, d, users, ni, mask[0] if mask else None, mask[1] if mask else None, k)
        n += 1
        if not (np.array_equal(idx, ri) and np.array_equal(sc, rs)):
            bad += 1
            print("MISMATCH seed", seed - 1, dict(d=d, k=k, ni=ni, nu=nu, nr=nr, ld=ld, splits=splits), flush=True)
    print("exact kernel: %d cases, %d mismatches, next seed %d" % (n, bad, seed))


def fuzz_p2p(seconds, seed):
    from oracle import bpr_oracle as O
    from recsys_pytorch_b200._lib import F_P2P_PURE_SEQUENTIAL, F_P2P_ROUND_ROBIN, F_USERS_UNIQUE
    from tests.test_p2p_on_simt_host import _make, _sync_head, _tables
    simt = simt_host.build_p2p(tempfile.mkdtemp())
    t_end, n, bad = time.time() + seconds, 0, 0
    while time.time() < t_end:
        rng = np.random.default_rng(seed); seed += 1
        W = int(rng.choice([1, 2, 3, 4, 5, 8, 16])); d = int(rng.choice([128, 128, 128, 50, 64, 200, 256, 400, 8]))
        variant = int(rng.choice([16, 8, 0])) if d == 128 else 0
        head = int(rng.choice([0, 0, 17, 100]))
        nu = int(rng.integers(max(W, 20), 400)); ni = int(rng.integers(head + 2 * nu + W + 10, head + 2 * nu + 1500))
        bounds = None if W == 1 else sorted({head, ni} | set(rng.choice(np.arange(head + 1, ni), W - 1, replace=False).tolist()))
        ranks, U0, V0, _, _, ib, ub = _make(W, nu, ni, d, seed=seed, bounds=bounds, head=head)
        items = rng.permutation(ni); gu, gi, gj, k = [], [], [], 0
        for r in ranks:
            n_loc = ub[r.rank + 1] - ub[r.rank]
            B = int(rng.integers(0, n_loc + 1))                              # empty batches included
            ul = rng.permutation(n_loc)[:B].astype(np.int32)
            pi, pj = items[k:k + B].astype(np.int32), items[k + B:k + 2 * B].astype(np.int32); k += 2 * B
            gu.append(ul + ub[r.rank]); gi.append(pi); gj.append(pj)
            r.route(simt, ul, 1, pos=pi, neg=pj)
        gu, gi, gj = np.concatenate(gu), np.concatenate(gi), np.concatenate(gj)
        Bg = max(len(gu), 1)
        snap = ranks[0].Vh.copy()
        flags = F_USERS_UNIQUE | int(rng.choice([0, F_P2P_ROUND_ROBIN, F_P2P_PURE_SEQUENTIAL]))
        for r in ranks:
            r.step(simt, ranks, Bg, variant, flags=flags)
        _sync_head(ranks, snap)
        ok = sum(int(r.n_processed[0]) for r in ranks) == len(gu)
        if len(gu):
            Ur, Vr, lref = O.sgd_step(U0, V0, gu, gi, gj, 0.9, 0.01)
            U, V = _tables(ranks, d)
            ok = ok and np.allclose(U, Ur, rtol=2e-5, atol=2e-6) and np.allclose(V, Vr, rtol=2e-5, atol=2e-6) and \
                abs(sum(r.loss[0] for r in ranks) / Bg - float(lref)) < 2e-5 * max(1, float(lref))
        n += 1
        if not ok:
            bad += 1
            print("MISMATCH seed", seed - 1, dict(W=W, d=d, variant=variant, head=head, nu=nu, ni=ni, flags=flags), flush=True)
    print("P2P step: %d cases, %d mismatches, next seed %d" % (n, bad, seed))


if __name__ == "__main__":
    which, seconds, seed = sys.argv[1], float(sys.argv[2]), int(sys.argv[3])
    {"tc": fuzz_tc, "exact": fuzz_exact, "p2p": fuzz_p2p}[which](seconds, seed)
