"""`-m "not gpu"`: the pointwise-MF step kernel (csrc/pointwise_step.cu) and the NGCF layer kernels (csrc/ngcf.cu: fused
forward, row backward, weight gradients) executed ON THE HOST by the SIMT emulator of tests/simt_host.py from their own
source text - against the golden vectors the reference produced (tests/golden/tiny_pointwise.npz) and the numpy oracle
(oracle/bpr_oracle.py, itself pinned to the reference's NGCF autograd in tests/test_oracle_cpu.py)."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import bpr_oracle as O
from recsys_pytorch_b200._lib import SINK_GRAD, SINK_NONE, SINK_UPDATE

pytestmark = pytest.mark.timeout(1200)
P = lambda a: a.ctypes.data if a is not None else None


@pytest.fixture(scope="module")
def simt(tmp_path_factory):
    from tests.simt_host import build_pw_ngcf
    return build_pw_ngcf(str(tmp_path_factory.mktemp("simt_pw")))


def _pad(W):
    ld = (W.shape[1] + 3) // 4 * 4
    out = np.zeros((W.shape[0], ld), np.float32); out[:, :W.shape[1]] = W
    return out


def _pw(simt, U, V, d, u, i, y, lf, lr=0.0, reg=0.0, sink=SINK_UPDATE, grid=2):
    u, i, y = np.ascontiguousarray(u, np.int32), np.ascontiguousarray(i, np.int32), np.ascontiguousarray(y, np.float32)
    gU, gV, loss = np.zeros_like(U), np.zeros_like(V), np.zeros(1, np.float64)
    rc = simt.emu_pointwise_step(P(U), P(V), U.shape[1], d, P(u), P(i), P(y), len(u), 1 if lf == "mse" else 0, lr, reg, sink,
                                 P(gU), P(gV), P(loss), 0.0, grid)
    return gU, gV, loss[0], rc


@pytest.mark.parametrize("lf", ["ce", "mse"])
def test_pointwise_kernel_matches_the_reference_golden(simt, golden, lf):
    """models/MF.py:101-102 through the reference's own autograd (duplicate users AND items in the batch): loss and the
    dense gradient rows; SINK_GRAD / SINK_NONE leave the tables alone."""
    g, t = golden["tiny_pointwise"], golden["tiny_bpr"]
    sc = float(g[f"{lf}_scale"])
    U, V = _pad((t["U0"] * sc).astype(np.float32)), _pad((t["V0"] * sc).astype(np.float32))
    U0, V0, d = U.copy(), V.copy(), t["U0"].shape[1]
    u, i, y = g["users"][0], g["items"][0], g["ratings"][0]
    gU, gV, loss, _ = _pw(simt, U, V, d, u, i, y, lf, sink=SINK_GRAD)
    B = len(u)
    assert abs(loss / B - float(g[f"{lf}_loss"])) <= 2e-5 * max(1.0, abs(float(g[f"{lf}_loss"])))
    np.testing.assert_allclose(gU[:, :d], g[f"{lf}_dU"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(gV[:, :d], g[f"{lf}_dV"], rtol=2e-5, atol=2e-6)
    assert np.array_equal(U, U0) and np.array_equal(V, V0)
    _, _, loss2, _ = _pw(simt, U, V, d, u, i, y, lf, sink=SINK_NONE)
    assert abs(loss2 - loss) < 1e-9 and np.array_equal(U, U0)


@pytest.mark.parametrize("d", [8, 50, 64, 128, 200, 256])
@pytest.mark.parametrize("lf", ["ce", "mse"])
def test_pointwise_kernel_in_place_update_all_widths(simt, d, lf):
    """In-place SGD with per-occurrence L2 at every row-width class, no repeated rows -> exact: W -= lr (g other + reg/B W)."""
    rng = np.random.default_rng(d)
    nu, ni, B = 300, 400, 211
    U0 = (rng.standard_normal((nu, d)) * 0.3).astype(np.float32); V0 = (rng.standard_normal((ni, d)) * 0.3).astype(np.float32)
    u, i = rng.permutation(nu)[:B], rng.permutation(ni)[:B]
    y = (rng.random(B) < 0.4).astype(np.float32)
    U, V = _pad(U0), _pad(V0)
    _, _, loss, rc = _pw(simt, U, V, d, u, i, y, lf, lr=0.8, reg=0.02)
    d4 = U.shape[1] // 4
    G = 1
    while G < d4 and G < 32:
        G <<= 1
    assert rc == G * 100 + (d4 + G - 1) // G
    dU, dV, _, _ = O.pointwise_grads(U0, V0, u, i, y, lf)
    Ur = U0 - np.float32(0.8) * (dU + np.float32(0.02 / B) * np.where(np.isin(np.arange(nu), u)[:, None], U0, 0))
    Vr = V0 - np.float32(0.8) * (dV + np.float32(0.02 / B) * np.where(np.isin(np.arange(ni), i)[:, None], V0, 0))
    np.testing.assert_allclose(U[:, :d], Ur, rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(V[:, :d], Vr, rtol=2e-5, atol=2e-6)
    lref, _ = O.pointwise_loss(U0, V0, u, i, y, lf)
    assert abs(loss / B - float(lref)) < 2e-5 * max(1.0, float(lref))
    assert not U[:, d:].any() and not V[:, d:].any()


# ---- NGCF ----------------------------------------------------------------------------------------------------------------
def _ngcf_problem(seed, n, d, L):
    rng = np.random.default_rng(seed)
    rows = [np.sort(rng.choice(n, int(rng.integers(1, 12)), replace=False)) for _ in range(n)]
    indptr = np.zeros(n + 1, np.int64); indptr[1:] = np.cumsum([len(r) for r in rows])
    A = sp.csr_matrix((rng.random(int(indptr[-1])).astype(np.float32) * 0.3, np.concatenate(rows), indptr), shape=(n, n))
    A = ((A + A.T) * 0.5).tocsr().astype(np.float32)                          # symmetric, like the normalised adjacency
    E0 = (rng.standard_normal((n, d)) * 0.3).astype(np.float32)
    Wg = [(rng.standard_normal((d, d)) * 0.3).astype(np.float32) for _ in range(L)]
    Wb = [(rng.standard_normal((d, d)) * 0.3).astype(np.float32) for _ in range(L)]
    bg = [(rng.standard_normal((1, d)) * 0.1).astype(np.float32) for _ in range(L)]
    bb = [(rng.standard_normal((1, d)) * 0.1).astype(np.float32) for _ in range(L)]
    return A, E0, Wg, bg, Wb, bb


@pytest.mark.parametrize("d", [16, 40, 64])
def test_ngcf_layer_kernels_forward_and_backward_match_the_oracle(simt, d):
    """ngcf.py::update_ngcf_embedding / backward_from_gout with the SpMM done by scipy: per layer the fused forward kernel
    (two [d,d] transforms, leaky-relu, L2 normalise, running mean), then the backward kernels (normalise / leaky-relu
    backward, gz W^T products, weight and bias gradients) - propagated table, dE0 and every dW / db against the oracle."""
    n, L = 150, 2
    A, E0, Wg, bg, Wb, bb = _ngcf_problem(d, n, d, L)
    ld = (d + 3) // 4 * 4
    s = 1.0 / (L + 1)
    out = _pad(E0) * np.float32(s)                                            # out = E0 / (L+1): first entry of `embs`
    ego = [_pad(E0)] + [np.zeros((n, ld), np.float32) for _ in range(L)]
    side = [np.zeros((n, ld), np.float32) for _ in range(L)]
    nrm = [np.zeros(n, np.float32) for _ in range(L)]
    for k in range(L):
        side[k][:, :d] = (A @ ego[k][:, :d]).astype(np.float32)
        simt.emu_ngcf_forward(P(ego[k]), P(side[k]), P(Wg[k]), P(bg[k]), P(Wb[k]), P(bb[k]), n, ld, d, k, 0.0, 5, 1, P(ego[k + 1]),
                              P(nrm[k]), P(out), s, 3)
    ref_out, cache = O.ngcf_forward(A, E0, Wg, bg, Wb, bb, keep=True)
    np.testing.assert_allclose(out[:, :d], ref_out, rtol=2e-5, atol=2e-6)
    G = (np.random.default_rng(1).standard_normal((n, d)) * 0.1).astype(np.float32)
    gout = _pad(G)
    gnext = None
    got = {}
    for k in range(L - 1, -1, -1):
        gz, gside, gego = (np.zeros((n, ld), np.float32) for _ in range(3))
        dWg, dWb, db = np.zeros((d, d), np.float32), np.zeros((d, d), np.float32), np.zeros(d, np.float32)
        simt.emu_ngcf_backward(P(gout), P(gnext), P(ego[k]), P(side[k]), P(ego[k + 1]), P(nrm[k]), P(Wg[k]), P(Wb[k]), n, ld, d, k,
                               0.0, 5, 1, s, P(gz), P(gside), P(gego), P(dWg), P(dWb), P(db), 3)
        gego[:, :d] += (A @ gside[:, :d]).astype(np.float32)                  # d(ego) += A_hat d(side)
        gnext = gego
        got[k] = (dWg, dWb, db)
    dE0 = s * G + gnext[:, :d]
    rE0, rWg, rbg, rWb, rbb = O.ngcf_backward(A, G, cache, Wg, Wb)
    np.testing.assert_allclose(dE0, rE0, rtol=3e-4, atol=2e-7)
    for k in range(L):
        np.testing.assert_allclose(got[k][0], rWg[k], rtol=3e-4, atol=2e-6)
        np.testing.assert_allclose(got[k][1], rWb[k], rtol=3e-4, atol=2e-6)
        np.testing.assert_allclose(got[k][2], rbg[k].reshape(-1), rtol=3e-4, atol=2e-6)


def test_ngcf_message_dropout_mask_is_shared_by_forward_and_backward(simt):
    """mess_dropout > 0: the counter-RNG mask keeps ~ (1 - p) of the activations, scales the kept ones by 1 / (1 - p), is a
    pure function of (seed, step, layer, row, column) - the same launch twice gives the same bits, another step another
    mask - and the backward pass zeroes exactly the dropped positions."""
    n, d, p = 200, 32, 0.3
    A, E0, Wg, bg, Wb, bb = _ngcf_problem(7, n, d, 1)
    ego, side = _pad(E0), np.zeros((n, d), np.float32)
    side[:, :d] = (A @ E0).astype(np.float32)

    def fwd(step, p_):
        nxt, nrm, acc = np.zeros((n, d), np.float32), np.zeros(n, np.float32), np.zeros((n, d), np.float32)
        simt.emu_ngcf_forward(P(ego), P(side), P(Wg[0]), P(bg[0]), P(Wb[0]), P(bb[0]), n, d, d, 0, p_, 9, step, P(nxt), P(nrm), P(acc),
                              1.0, 2)
        return nxt, nrm
    full, _ = fwd(1, 0.0)
    a, nrm_a = fwd(1, p)
    b, _ = fwd(1, p)
    c, _ = fwd(2, p)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    kept = a != 0
    assert abs(kept.mean() - (1 - p)) < 0.03
    np.testing.assert_allclose(a[kept], full[kept] / np.float32(1 - p), rtol=1e-6)
    gz, gside, gego = (np.zeros((n, d), np.float32) for _ in range(3))
    dWg, dWb, db = np.zeros((d, d), np.float32), np.zeros((d, d), np.float32), np.zeros(d, np.float32)
    gout = np.ones((n, d), np.float32)
    simt.emu_ngcf_backward(P(gout), None, P(ego), P(side), P(a), P(nrm_a), P(Wg[0]), P(Wb[0]), n, d, d, 0, p, 9, 1, 1.0, P(gz),
                           P(gside), P(gego), P(dWg), P(dWb), P(db), 2)
    assert (gz[~kept] == 0).all() and (gz[kept] != 0).mean() > 0.99
