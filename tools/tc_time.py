"""Manual timing driver (not a pytest): SurfaceNet forward on random input, per-unit CUDA-event times.
    python tests/tc_time.py [mode] [n_pair_cubes] [D]"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from surfacenet_b200 import weights, SurfaceNet, _lib
mode = sys.argv[1] if len(sys.argv) > 1 else "exact"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
D = int(sys.argv[3]) if len(sys.argv) > 3 else 64
net = SurfaceNet.Net(weights.synthetic_params(0))
X = (torch.rand((n, 6, D, D, D), device="cuda") * 255 - 110)
for _ in range(2):
    net.forward(X, None, 1, mode)
torch.cuda.synchronize()
_lib.lib.sn_profile_enable(1)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
R = 3
for _ in range(R):
    net.forward(X, None, 1, mode)
b.record(); torch.cuda.synchronize()
nu = len(weights.UNITS)
ms = (C.c_double * nu)(); cnt = (C.c_int64 * nu)()
_lib.lib.sn_profile_collect(ms, cnt, nu)
tot = a.elapsed_time(b) / R
flop = 1358389.0 * n * D ** 3
print("env NB=%s ROT=%s mode=%s n=%d D=%d: %.2f ms/forward  %.1f TFLOP/s  %.3g pair-voxels/s" % (os.environ.get("SN_TC_NB"), os.environ.get("SN_TC_ROT"), mode, n, D, tot, flop / tot / 1e9, n * D ** 3 / tot * 1e3))
res = {"conv1": 1, "side_op1": 1, "merge": 1, "conv2": 8, "side_op2": 8, "conv3": 64, "side_op3": 64, "conv4": 64, "side_op4": 64}
line = []
for i, (name, kind, cin, cout, k) in enumerate(weights.UNITS):
    if not cnt[i]: continue
    key = [p for p in res if name.startswith(p)][0]
    f = 2.0 * cin * cout * k ** 3 / res[key] * n * D ** 3
    line.append("%s %.2fms/%.0fTF" % (name, ms[i] / R, f * R / ms[i] / 1e9))
print("   " + "  ".join(line))
