"""Manual debugging helper (not a pytest): run single conv units through the tensor-core path and print
error statistics against the torch-CPU oracle.   python tests/tc_debug.py [mode] [unit S]..."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import surfacenet_oracle as so
from surfacenet_b200 import _lib, weights, SurfaceNet

mode = sys.argv[1] if len(sys.argv) > 1 else "exact"
cases = [("side_op1", 8), ("conv1_2", 8), ("conv1_2", 16), ("merge_conv", 8), ("conv4_2", 8), ("conv2_2", 17)]
if len(sys.argv) > 3:
    cases = [(sys.argv[i], int(sys.argv[i + 1])) for i in range(2, len(sys.argv) - 1, 2)]
params = weights.synthetic_params(0)
net = SurfaceNet.Net(params)
names = [u[0] for u in weights.UNITS]
for name, S in cases:
    u = names.index(name)
    _, kind, cin, cout, k = weights.UNITS[u]
    rs = np.random.RandomState(u)
    x = (rs.standard_normal((2, cin, S, S, S)) * 1.5).astype(np.float32)
    act = "sigmoid" if name in weights.SIGMOID_UNITS else "relu"
    with torch.no_grad():
        ref = so.conv_bn(torch.from_numpy(x), params, weights.unit_index()[name], act, dilated=(kind == "dil")).numpy()
    out = torch.full((2, cout, S, S, S), float("nan"), dtype=torch.float32, device="cuda")
    rc = _lib.lib.sn_net_layer_conv(net.handle, u, _lib.ptr(torch.from_numpy(x).cuda()), 2, S, _lib.ptr(out), _lib.MODES[mode], _lib.stream_ptr())
    if rc != 0:
        print(name, S, "rc", rc, _lib.last_error()); continue
    torch.cuda.synchronize()
    o = out.cpu().numpy()
    e = np.abs(o - ref)
    bad = ~np.isfinite(o)
    print("%-12s S=%-3d %s: max-abs %.3g  mean-abs %.3g  ref max %.3g  nan %d  frac(err>1e-3) %.4f" %
          (name, S, mode, np.nanmax(e), np.nanmean(e), np.abs(ref).max(), bad.sum(), np.nanmean(e > 1e-3)), flush=True)
    if np.nanmax(e) > 1e-3:
        idx = np.argwhere(np.nan_to_num(e, nan=9) > 1e-3)
        print("   first bad idx", idx[:5].tolist(), " by channel:", np.unique(idx[:, 1])[:20].tolist(), " by w:", np.unique(idx[:, 4]).tolist()[:20],
              "by h:", np.unique(idx[:, 3]).tolist()[:20], "by d:", np.unique(idx[:, 2]).tolist()[:20])
        print("   sample got/ref:", o[tuple(idx[0])], ref[tuple(idx[0])])
