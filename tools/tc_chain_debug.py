"""Manual debugging helper (not a pytest): run the network unit by unit through sn_net_layer_conv (mode given),
feeding each unit the DEVICE's previous output, and print the error against the oracle's taps at every unit."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.nn.functional as F
from oracle import surfacenet_oracle as so, cvc_oracle
from surfacenet_b200 import _lib, weights, SurfaceNet
from tests import util

mode = sys.argv[1] if len(sys.argv) > 1 else "exact"
D = 32
cams = util.dtu_cameras()
rs = np.random.RandomState(0)
used = [8, 9, 22, 23, 30, 33]
imgs = util.image_list(49, used)
pairs = rs.choice(used, size=(2, 2, 2))
xyz = (np.array([10.0, -30.0, 620.0]) + rs.rand(2, 3) * 20).astype(np.float32)
X = cvc_oracle.gen_coloredCubes(pairs, xyz, np.full(2, 0.4, np.float32), cams, imgs, D)
_, X = cvc_oracle.preprocess_augmentation(None, X, util.MEAN6[None, :, None, None, None], False, False)
params = weights.synthetic_params(0)
with torch.no_grad():
    out_o, taps = so.one_viewpair_forward(X, params, return_taps=True)
net = SurfaceNet.Net(params)
names = [u[0] for u in weights.UNITS]
M = _lib.MODES[mode]

def conv(name, x):
    u = names.index(name)
    cout = weights.UNITS[u][3]
    n, _, S = x.shape[0], x.shape[1], x.shape[2]
    out = torch.empty((n, cout, S, S, S), dtype=torch.float32, device="cuda")
    _lib.check(_lib.lib.sn_net_layer_conv(net.handle, u, _lib.ptr(x.contiguous()), n, S, _lib.ptr(out), M, _lib.stream_ptr()))
    return out

def rep(name, t, key=None):
    ref = taps[key or name].numpy()
    e = np.abs(t.cpu().numpy() - ref)
    print("%-12s max-abs %.3g  mean-abs %.3g  ref max %.3g  rel %.3g" % (name, e.max(), e.mean(), np.abs(ref).max(), e.max() / np.abs(ref).max()), flush=True)

def up(t, key, f):
    return so.upsample(t.cpu(), params[so.LAYOUT[key]], f).cuda()

x = torch.from_numpy(X).cuda()
c11 = conv("conv1_1", x); rep("conv1_1", c11)
c12 = conv("conv1_2", c11); rep("conv1_2", c12)
c13 = conv("conv1_3", c12); rep("conv1_3", c13)
s1 = conv("side_op1", c13); rep("side_op1", s1)
p1 = F.max_pool3d(c13, 2, 2)
c21 = conv("conv2_1", p1); rep("conv2_1", c21)
c22 = conv("conv2_2", c21); rep("conv2_2", c22)
c23 = conv("conv2_3", c22); rep("conv2_3", c23)
s2 = conv("side_op2", c23); rep("side_op2", s2)
p2 = F.max_pool3d(c23, 2, 2)
c31 = conv("conv3_1", p2); rep("conv3_1", c31)
c32 = conv("conv3_2", c31); rep("conv3_2", c32)
c33 = conv("conv3_3", c32); rep("conv3_3", c33)
s3 = conv("side_op3", c33); rep("side_op3", s3)
c41 = conv("conv4_1", c33); rep("conv4_1", c41)
c42 = conv("conv4_2", c41); rep("conv4_2", c42)
c43 = conv("conv4_3", c42); rep("conv4_3", c43)
s4 = conv("side_op4", c43); rep("side_op4", s4)
cat = torch.cat([s1, up(s2, "up2_W", 2), up(s3, "up3_W", 4), up(s4, "up4_W", 4)], 1); rep("concat", cat)
m1 = conv("merge_conv", cat); rep("merge_conv", m1)
m2 = conv("merge_conv2", m1); rep("merge_conv2", m2)
o = conv("merge_conv3", m2); rep("out", o)
