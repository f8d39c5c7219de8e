#!/usr/bin/env python
"""Calibrate the BatchNorm statistics of the synthetic SurfaceNet weights
and write surfacenet_b200/data/synth_bn_seed<seed>.npz.  Run once in the build container:

    python tests/golden/make_synth_bn.py [seed]

Inputs: real DTU scan9 images + cameras from /root/reference/inputs (read only), 2 cubes x 2 view
pairs at s=32 pushed through the CPU oracle; at every conv+BN unit the stored mean / inv_std are set
from the batch so the unit's output is zero-mean / unit-variance per channel (eps = 1e-4 as in
Lasagne), ReLU units get gamma~U(0.8,1.2), beta~N(0,0.1^2); sigmoid units gamma=2, beta~N(0,0.3^2);
the final 1-channel unit gamma=2, beta=-0.5 (about 40 % of voxels end above min_prob=0.46).
"""
import os, sys
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
from oracle import cvc_oracle, surfacenet_oracle as so
from surfacenet_b200 import weights
from tests import util


def main(seed=0):
    from PIL import Image
    cams = util.dtu_cameras()
    used = [8, 9, 22, 23, 30, 33]
    imgs = [None] * 49
    for v in used:
        imgs[v] = np.asarray(Image.open("/root/reference/inputs/DTU_MVS/Rectified/scan9/rect_{:03}_3_r5000.jpg".format(v + 1)).convert("RGB"))
    pairs = np.array([[[8, 9], [22, 23]], [[30, 33], [9, 22]]])
    xyz = np.array([[10.0, -30.0, 620.0], [30.0, 0.0, 650.0]], dtype=np.float32)
    X = cvc_oracle.gen_coloredCubes(pairs, xyz, np.full(2, 0.4, np.float32), cams, imgs, 32)
    _, X = cvc_oracle.preprocess_augmentation(None, X, util.MEAN6[None, :, None, None, None], False, False)
    print("CVC input: nonzero voxels", float((X + util.MEAN6[None, :, None, None, None] != 0).mean()))

    p = weights.synthetic_params(seed, calibrated=False)
    rs = np.random.RandomState(seed + 1)
    idx = weights.unit_index()
    name_of = {idx[n]: n for n, k, *_ in weights.UNITS if k != "up"}
    out = {}

    def calib_bn_act(x, p, i, act):
        name = name_of[i]
        C = x.shape[1]
        xm = x.transpose(0, 1).reshape(C, -1).double()
        mean = xm.mean(1).float().numpy()
        inv_std = (1.0 / torch.sqrt(xm.var(1, unbiased=False) + 1e-4)).float().numpy()
        if name == "merge_conv3":
            gamma, beta = np.full(C, 2.0, np.float32), np.full(C, -0.5, np.float32)
        elif act == "sigmoid":
            gamma, beta = np.full(C, 2.0, np.float32), (0.3 * rs.standard_normal(C)).astype(np.float32)
        else:
            gamma, beta = rs.uniform(0.8, 1.2, C).astype(np.float32), (0.1 * rs.standard_normal(C)).astype(np.float32)
        p[i + 1], p[i + 2], p[i + 3], p[i + 4] = beta, gamma, mean, inv_std
        for j, key in enumerate(("beta", "gamma", "mean", "inv_std")):
            out[name + "." + key] = p[i + 1 + j]
        return orig(x, p, i, act)

    orig = so._bn_act
    so._bn_act = calib_bn_act
    with torch.no_grad():
        y = so.one_viewpair_forward(X, p)
    so._bn_act = orig
    y = y.numpy()
    print("output prob: mean %.3f  frac>0.46 %.3f  min %.3f max %.3f" % (y.mean(), (y > 0.46).mean(), y.min(), y.max()))
    path = os.path.join(REPO, "surfacenet_b200", "data", "synth_bn_seed{}.npz".format(seed))
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
