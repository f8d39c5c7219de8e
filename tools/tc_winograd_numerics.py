#!/usr/bin/env python
"""CPU experiment (not a test): would a 1-D Winograd F(2,3) along w keep the exact mode's accuracy?  Emulates, in float64 with
explicit roundings, a merge_conv2-like unit (100 -> 100 channels, 3x3x3) three ways against the float64 convolution:
  fp32      : float32 operands (what the reference computes with)
  split3    : fp16 hi+lo split of both operands, three cross products (the shipped exact mode; accumulation error ignored)
  winograd  : input transform of the 22-bit activations in fp32, re-split to fp16 hi+lo; weight transform in float64, split;
              three cross products per frequency; output transform in fp32
    python tests/tc_winograd_numerics.py"""
import numpy as np
import torch

f16 = lambda x: x.to(torch.float16).to(torch.float64)
f32 = lambda x: x.to(torch.float32).to(torch.float64)


def split(x):
    hi = f16(x)
    return hi, f16(x - hi)


def conv(a, w):
    return torch.nn.functional.conv3d(a, w, padding=1)


def prod3(ah, al, wh, wl, convf):
    return convf(ah, wh) + convf(ah, wl) + convf(al, wh)


def main():
    torch.manual_seed(0)
    C, S = 100, 12
    a = torch.relu(torch.randn(1, C, S, S, S, dtype=torch.float64))
    w = torch.randn(C, C, 3, 3, 3, dtype=torch.float64) * (2.0 / (27 * C)) ** 0.5
    scale = 2.0 ** np.floor(np.log2(1024.0 / w.abs().max().item()))             # the kernel's power-of-two pre-scaling
    ref = conv(a, w)
    sig = ref.std().item()
    rep = lambda name, y: print("%-9s max-abs error / sigma(out) = %.3g" % (name, (y - ref).abs().max().item() / sig))
    rep("fp32", conv(f32(a), f32(w)))
    rep("fp16", conv(f16(a), f16(w * scale)) / scale)
    a22 = sum(split(a))                                                        # what the previous unit's epilogue stores
    ah, al = split(a22)
    wh, wl = split(w * scale)
    rep("split3", prod3(ah, al, wh, wl, conv) / scale)
    # Winograd F(2,3) along the last axis: outputs (2j, 2j+1) from inputs 2j-1 .. 2j+2
    ap = torch.nn.functional.pad(a22, (1, 1))                                   # zero 'same' padding along w only
    d = [ap[..., k:k + S:2] for k in range(4)]                                  # d0..d3 for every output pair
    V = [f32(d[0] - d[2]), f32(d[1] + d[2]), f32(d[2] - d[1]), f32(d[1] - d[3])]
    g = [w[..., k] for k in range(3)]
    U = [g[0], (g[0] + g[1] + g[2]) / 2, (g[0] - g[1] + g[2]) / 2, g[2]]
    conv2 = lambda x, k: torch.nn.functional.conv3d(x, k.unsqueeze(-1), padding=(1, 1, 0))     # taps in d, h only
    M = []
    for v, u in zip(V, U):
        vh, vl = split(v)
        uh, ul = split(u * scale)
        M.append(f32(prod3(vh, vl, uh, ul, conv2) / scale))
    y = torch.empty_like(ref)
    y[..., 0::2] = f32(f32(M[0] + M[1]) + M[2])
    y[..., 1::2] = f32(f32(M[1] - M[2]) - M[3])
    rep("winograd", y)


if __name__ == "__main__":
    main()
