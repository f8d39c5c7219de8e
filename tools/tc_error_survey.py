"""Manual survey (not a pytest): exact-mode max-abs error vs the torch-CPU fp32 oracle over several random batches."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import surfacenet_oracle as so, cvc_oracle
from surfacenet_b200 import SurfaceNet, weights
from tests import util
cams = util.dtu_cameras()
params = weights.synthetic_params(0)
mode = sys.argv[1] if len(sys.argv) > 1 else "exact"
_, fn = SurfaceNet.SurfaceNet_inference(2, params, mode=mode)
used = list(range(0, 49, 3))
imgs = util.image_list(49, used)
worst = 0
for seed in range(int(sys.argv[2]) if len(sys.argv) > 2 else 6):
    rs = np.random.RandomState(100 + seed)
    D = 32
    pairs = rs.choice(used, size=(2, 2, 2))
    xyz = (util.SCAN9_BB[0] + rs.rand(2, 3) * (util.SCAN9_BB[1] - util.SCAN9_BB[0] - D * 0.4)).astype(np.float32)
    X = cvc_oracle.gen_coloredCubes(pairs, xyz, np.full(2, 0.4, np.float32), cams, imgs, D)
    _, X = cvc_oracle.preprocess_augmentation(None, X, util.MEAN6[None, :, None, None, None], False, False)
    w = (rs.rand(2, 2) + 0.1).astype(np.float32)
    fo, uo = so.nViewPair_SurfaceNet_fn(X, params, w, N_vp=2)
    f, u = fn(X, w)
    ef, eu = np.abs(f - fo).max(), np.abs(u - uo).max()
    worst = max(worst, ef, eu)
    print("seed %d: fused %.3g unfused %.3g  (nonzero input frac %.2f, prob mean %.3f)" % (seed, ef, eu, float((X + util.MEAN6[None, :, None, None, None] != 0).mean()), fo.mean()), flush=True)
print("worst", worst)
