#!/usr/bin/env python
"""One exact-mode forward of n pair-cubes (default 8 x 64^3) for ncu: python tools/wg_profile.py [n] [D]"""
import os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from surfacenet_b200 import SurfaceNet, weights

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
D = int(sys.argv[2]) if len(sys.argv) > 2 else 64
net = SurfaceNet.Net(weights.synthetic_params(0))
rs = np.random.RandomState(0)
X = torch.from_numpy((rs.randint(0, 256, size=(n, 6, D, D, D)).astype(np.float32) - 115.0)).cuda()
for _ in range(2):
    fused, _ = net.forward(X, None, 1, "exact")
torch.cuda.synchronize()
print("ok", float(fused.mean()))
