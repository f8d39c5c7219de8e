#!/usr/bin/env python
"""Per-unit error of the exact / fp32 modes on REALISTIC activations (GPU box): the fp64 graph (tools/parity_survey.truth_forward with
taps) provides every unit's true input; each unit is run alone through sn_net_layer_conv on that input (cast to fp32) and compared with
the fp64 result of the same unit on the same fp32 input.  Reports max-abs and rms error relative to the rms of the unit's output."""
import os, sys, json
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
import torch.nn.functional as F
from oracle.surfacenet_oracle import LAYOUT as L
from surfacenet_b200 import SurfaceNet, _lib, weights
from tests import util
sys.path.insert(0, os.path.join(REPO, "tools"))
from parity_survey import make_inputs

D = int(sys.argv[1]) if len(sys.argv) > 1 else 64
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
params = weights.synthetic_params(seed)
net = SurfaceNet.Net(params)
X, _ = make_inputs(util.dtu_cameras(), D, 1, 1, seed)
dev, dt = "cuda", torch.float64
t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(device=dev, dtype=dt)


def cb(x, name, act, dil=False):
    i = L[name]
    W = t(params[i])
    if dil:
        W = W.permute(1, 0, 2, 3, 4).contiguous()
        y = F.conv3d(x, W, padding=2 * (W.shape[-1] // 2), dilation=2)
    else:
        y = F.conv3d(x, W, padding=W.shape[-1] // 2)
    beta, gamma, mean, inv_std = (t(params[i + k]) for k in (1, 2, 3, 4))
    sh = (1, -1, 1, 1, 1)
    y = (y - mean.view(sh)) * (gamma * inv_std).view(sh) + beta.view(sh)
    return torch.relu(y) if act == "relu" else torch.sigmoid(y)


def up(x, W, f):
    n, c, d, h, w = x.shape
    z = torch.zeros((n, c, d * f, h * f, w * f), dtype=dt, device=dev)
    z[:, :, ::f, ::f, ::f] = x
    W = t(W)
    return F.conv3d(z.reshape(n * c, 1, d * f, h * f, w * f), W, padding=W.shape[-1] // 2).reshape(n, c, d * f, h * f, w * f)


names = [u[0] for u in weights.UNITS]
rows = []


def unit(x, name, act, dil=False):
    """x: fp64 true input.  Returns the fp64 output on the TRUE input (for the chain); records the unit's own error."""
    x32 = x.float()
    ref = cb(x32.double(), name, act, dil)
    u = names.index(name)
    n, cin, S = x32.shape[0], x32.shape[1], x32.shape[2]
    cout = ref.shape[1]
    row = dict(unit=name, S=S, out_rms=float(ref.pow(2).mean().sqrt()))
    for mode in ("exact", "fp32"):
        out = torch.empty((n, cout, S, S, S), dtype=torch.float32, device="cuda")
        _lib.check(_lib.lib.sn_net_layer_conv(net.handle, u, _lib.ptr(x32.contiguous()), n, S, _lib.ptr(out), _lib.MODES[mode], _lib.stream_ptr()))
        torch.cuda.synchronize()
        e = (out.double() - ref)
        row[mode + "_max"] = float(e.abs().max()); row[mode + "_rms"] = float(e.pow(2).mean().sqrt())
    rows.append(row)
    print(json.dumps(row), flush=True)
    return cb(x, name, act, dil)


with torch.no_grad():
    x = t(X)
    c13 = unit(unit(unit(x, "conv1_1", "relu"), "conv1_2", "relu"), "conv1_3", "relu")
    s1 = unit(c13, "side_op1", "sigmoid")
    c23 = unit(unit(unit(F.max_pool3d(c13, 2, 2), "conv2_1", "relu"), "conv2_2", "relu"), "conv2_3", "relu")
    s2u = up(unit(c23, "side_op2", "sigmoid"), params[L["up2_W"]], 2)
    c33 = unit(unit(unit(F.max_pool3d(c23, 2, 2), "conv3_1", "relu"), "conv3_2", "relu"), "conv3_3", "relu")
    s3u = up(unit(c33, "side_op3", "sigmoid"), params[L["up3_W"]], 4)
    c43 = unit(unit(unit(c33, "conv4_1", "relu", True), "conv4_2", "relu", True), "conv4_3", "relu", True)
    s4u = up(unit(c43, "side_op4", "sigmoid", True), params[L["up4_W"]], 4)
    m2 = unit(unit(torch.cat([s1, s2u, s3u, s4u], dim=1), "merge_conv", "relu"), "merge_conv2", "relu")
    unit(m2, "merge_conv3", "sigmoid")
print("%-12s %4s %10s | %10s %10s | %10s %10s" % ("unit", "S", "out_rms", "exact_max", "exact_rms", "fp32_max", "fp32_rms"))
for r in rows:
    print("%-12s %4d %10.3g | %10.3g %10.3g | %10.3g %10.3g" % (r["unit"], r["S"], r["out_rms"], r["exact_max"], r["exact_rms"], r["fp32_max"], r["fp32_rms"]))
