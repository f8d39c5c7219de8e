"""Manual profiling driver (not a pytest): one SurfaceNet forward on random input, for ncu.
    python tests/tc_profile.py [mode] [n_pair_cubes] [D]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from surfacenet_b200 import weights, SurfaceNet
mode = sys.argv[1] if len(sys.argv) > 1 else "exact"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10
D = int(sys.argv[3]) if len(sys.argv) > 3 else 64
net = SurfaceNet.Net(weights.synthetic_params(0))
X = (torch.rand((n, 6, D, D, D), device="cuda") * 255 - 110)
for _ in range(2):
    fused, _ = net.forward(X, None, 1, mode)
torch.cuda.synchronize()
print("ok", float(fused.mean()))
