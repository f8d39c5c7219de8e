#!/usr/bin/env python
"""Algebra check (fp64, numpy) of the w-axis F(2,3) Winograd scheme used by csrc/conv_wg.cu:
   V-format input transform per pair t (voxels w=2t, 2t+1; dilation dl: pair (wa, wa+dl)), per-frequency weights U_f = G g,
   four per-frequency 'convolutions' over the remaining (kd, kh) taps and the output transform, against the direct 3x3x3 'same' conv."""
import numpy as np

def direct(x, w, dil=1):
    C, D, H, W = x.shape
    O = w.shape[0]
    xp = np.zeros((C, D + 2 * dil, H + 2 * dil, W + 2 * dil)); xp[:, dil:-dil, dil:-dil, dil:-dil] = x
    y = np.zeros((O, D, H, W))
    for kd in range(3):
        for kh in range(3):
            for kw in range(3):
                y += np.einsum("oc,cdhw->odhw", w[:, :, kd, kh, kw], xp[:, kd * dil:kd * dil + D, kh * dil:kh * dil + H, kw * dil:kw * dil + W])
    return y

def pairs_of_row(W, dil):
    """(wa, wb) per pair index t.  dil=1: (2t, 2t+1).  dil=2: t = 2j + p -> wa = p + 4j, wb = wa + 2 (sub-lattice of parity p)."""
    if dil == 1:
        return [(2 * t, 2 * t + 1) for t in range(W // 2)]
    return [((t % 2) + 4 * (t // 2), (t % 2) + 4 * (t // 2) + 2) for t in range(W // 2)]

def to_wino(x, dil=1):
    C, D, H, W = x.shape
    g = lambda w: x[..., w] if 0 <= w < W else np.zeros(x.shape[:-1])
    V = np.zeros((4, C, D, H, W // 2))
    for t, (wa, wb) in enumerate(pairs_of_row(W, dil)):
        d0, d1, d2, d3 = g(wa - dil), g(wa), g(wb), g(wb + dil)
        V[0, ..., t], V[1, ..., t], V[2, ..., t], V[3, ..., t] = d0 - d2, d1 + d2, d2 - d1, d1 - d3
    return V

def wino_conv(V, w, dil=1):
    G = np.array([[1, 0, 0], [.5, .5, .5], [.5, -.5, .5], [0, 0, 1]])
    U = np.einsum("fk,ocdhk->focdh", G, w)                       # (4, O, C, kd, kh)
    _, C, D, H, T = V.shape
    O = w.shape[0]
    M = np.zeros((4, O, D, H, T))
    Vp = np.zeros((4, C, D + 2 * dil, H + 2 * dil, T)); Vp[:, :, dil:-dil, dil:-dil] = V
    for f in range(4):
        for kd in range(3):
            for kh in range(3):
                M[f] += np.einsum("oc,cdht->odht", U[f, :, :, kd, kh], Vp[f, :, kd * dil:kd * dil + D, kh * dil:kh * dil + H])
    y0, y1 = M[0] + M[1] + M[2], M[1] - M[2] - M[3]
    y = np.zeros((O, D, H, 2 * T))
    for t, (wa, wb) in enumerate(pairs_of_row(2 * T, dil)):
        y[..., wa], y[..., wb] = y0[..., t], y1[..., t]
    return y

rs = np.random.RandomState(0)
for dil, S in [(1, 8), (2, 8), (2, 16)]:
    x, w = rs.randn(5, S, S, S), rs.randn(7, 5, 3, 3, 3)
    err = np.abs(wino_conv(to_wino(x, dil), w, dil) - direct(x, w, dil)).max()
    print("dil", dil, "S", S, "max err", err)
    assert err < 1e-10
print("OK")
