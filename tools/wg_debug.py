#!/usr/bin/env python
"""Debug aid (GPU box): one conv unit through sn_net_layer_conv in exact mode with and without the Winograd path, error maps per
output channel / position.  usage: python tools/wg_debug.py <unit> <S> [n]"""
import os, subprocess, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def run(name, S, n):
    import torch
    from oracle import surfacenet_oracle as so
    from surfacenet_b200 import SurfaceNet, _lib, weights
    params = weights.synthetic_params(0)
    net = SurfaceNet.Net(params)
    names = [u[0] for u in weights.UNITS]
    u = names.index(name)
    _, kind, cin, cout, k = weights.UNITS[u]
    rs = np.random.RandomState(100 + u)
    x = (rs.standard_normal((n, cin, S, S, S)) * 1.5).astype(np.float32)
    with torch.no_grad():
        ref = so.conv_bn(torch.from_numpy(x), params, weights.unit_index()[name], "relu", dilated=(kind == "dil")).numpy()
    xd = torch.from_numpy(x).cuda()
    out = torch.full((n, cout, S, S, S), float("nan"), dtype=torch.float32, device="cuda")
    _lib.check(_lib.lib.sn_net_layer_conv(net.handle, u, _lib.ptr(xd), n, S, _lib.ptr(out), _lib.MODES["exact"], _lib.stream_ptr()))
    torch.cuda.synchronize()
    o = out.cpu().numpy()
    e = np.abs(o - ref)
    print("SN_WG=%s %s S=%d n=%d: max-abs %.3g, ref max %.3g, nan %d" % (os.environ.get("SN_WG", "1"), name, S, n, np.nanmax(e), np.abs(ref).max(), int(np.isnan(o).sum())))
    if np.nanmax(e) > 1e-3 or np.isnan(o).any():
        e = np.nan_to_num(e, nan=9.0)
        print(" per sample  :", e.max(axis=(1, 2, 3, 4)))
        print(" per channel :", np.round(e.max(axis=(0, 2, 3, 4)), 3)[:40])
        print(" per d       :", np.round(e.max(axis=(0, 1, 3, 4)), 3))
        print(" per h       :", np.round(e.max(axis=(0, 1, 2, 4)), 3))
        print(" per w       :", np.round(e.max(axis=(0, 1, 2, 3)), 3))


if __name__ == "__main__":
    name, S = sys.argv[1], int(sys.argv[2])
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    run(name, S, n)
