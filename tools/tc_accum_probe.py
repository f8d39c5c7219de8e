"""Manual probe (not a pytest): how does tcgen05 kind::f16 accumulate in fp32?  Operands are chosen exactly
representable in fp16 (products and the true sum are exact in float64), so the only error is the accumulation."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.nn.functional as F
from surfacenet_b200 import _lib, weights, SurfaceNet

params = weights.synthetic_params(0, calibrated=False)
idx = weights.unit_index()
names = [u[0] for u in weights.UNITS]
rs = np.random.RandomState(0)
for name in ("conv4_2", "side_op4", "conv1_2", "merge_conv2"):
    i = idx[name]
    params[i] = (rs.randint(-512, 513, size=params[i].shape) / 1024.0).astype(np.float32)
    params[i].flat[0] = 0.5
net = SurfaceNet.Net(weights.validate(params))
for name, S, signed in (("conv4_2", 8, True), ("conv4_2", 8, False), ("conv1_2", 8, True), ("merge_conv2", 8, True), ("merge_conv2", 8, False)):
    u = names.index(name)
    _, kind, cin, cout, k = weights.UNITS[u]
    x = (rs.randint(-1024 if signed else 0, 1025, size=(1, cin, S, S, S)) / 256.0).astype(np.float32)
    W = torch.from_numpy(params[idx[name]]).double()
    if kind == "dil":
        W = W.permute(1, 0, 2, 3, 4).contiguous()
        ref = F.conv3d(torch.from_numpy(x).double(), W, padding=2 * (k // 2), dilation=2)
    else:
        ref = F.conv3d(torch.from_numpy(x).double(), W, padding=k // 2)
    ref = ref.numpy()
    for mode in ("exact", "fp32"):
        out = torch.empty((1, cout, S, S, S), dtype=torch.float32, device="cuda")
        _lib.check(_lib.lib.sn_net_layer_conv(net.handle, u, _lib.ptr(torch.from_numpy(x).cuda()), 1, S, _lib.ptr(out), _lib.MODES[mode], _lib.stream_ptr()))
        o = out.cpu().numpy().astype(np.float64)
        m = ref > 1.0                       # positive outputs survive the ReLU (identity BatchNorm)
        ulp = np.spacing(ref[m].astype(np.float32)).astype(np.float64)
        e = (o[m] - ref[m]) / ulp
        sig_ok = name not in weights.SIGMOID_UNITS
        print("%-12s signed=%d %-5s n=%d  |ref| mean %.1f  err/ulp: mean %+.2f  std %.2f  min %+.1f max %+.1f" %
              (name, signed, mode, m.sum(), ref[m].mean(), e.mean(), e.std(), e.min(), e.max()), flush=True)
