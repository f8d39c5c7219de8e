"""Host-side description of the SurfaceNet parameter list (105 arrays).

The reference stores weights as the flat list returned by ``lasagne.layers.get_all_param_values``
on the layers named in params.py:105 and restores it with ``set_all_param_values``
(nets/SurfaceNet.py:397-400).  This module owns that layout: shapes, validation, the py2-pickle
loader, and a deterministic synthetic generator (the real ``.model`` file is not distributed with
the reference: inputs/SurfaceNet_models/README.txt:1-2).
"""
import os
import pickle
import numpy as np

# (name, kind, C_in, C_out, k).  kind "conv": W (C_out,C_in,k,k,k)  [Conv3DDNNLayer]
#                                kind "dil" : W (C_in,C_out,k,k,k)  [DilatedConv3DLayer, nets/layers.py:200-253]
#                                kind "up"  : fixed W (1,1,k,k,k)    [Bilinear_3DInterpolation, nets/layers.py:376-390]
UNITS = [
    ("conv1_1", "conv", 6, 32, 3), ("conv1_2", "conv", 32, 32, 3), ("conv1_3", "conv", 32, 32, 3),
    ("side_op1", "conv", 32, 16, 1),
    ("conv2_1", "conv", 32, 80, 3), ("conv2_2", "conv", 80, 80, 3), ("conv2_3", "conv", 80, 80, 3),
    ("side_op2", "conv", 80, 16, 1), ("up2", "up", 1, 1, 3),
    ("conv3_1", "conv", 80, 160, 3), ("conv3_2", "conv", 160, 160, 3), ("conv3_3", "conv", 160, 160, 3),
    ("side_op3", "conv", 160, 16, 1), ("up3", "up", 1, 1, 5),
    ("conv4_1", "dil", 160, 300, 3), ("conv4_2", "dil", 300, 300, 3), ("conv4_3", "dil", 300, 300, 3),
    ("side_op4", "dil", 300, 16, 1), ("up4", "up", 1, 1, 5),
    ("merge_conv", "conv", 64, 100, 3), ("merge_conv2", "conv", 100, 100, 3), ("merge_conv3", "conv", 100, 1, 1),
]
SIGMOID_UNITS = ("side_op1", "side_op2", "side_op3", "side_op4", "merge_conv3")
D_VIEWPAIR_FEATURE = 258        # params.py:94
SIMILNET_HIDDEN = 100           # params.py:95
N_ARRAYS = 105
MEAN_CVC_RGBRGB = np.asarray([123.68, 116.779, 103.939, 123.68, 116.779, 103.939], dtype=np.float32)  # params.py:129
MEAN_PATCHES_BGR = np.asarray([103.939, 116.779, 123.68], dtype=np.float32)                           # params.py:130


def unit_index():
    """name -> index of the unit's first array in the 105-list."""
    idx, i = {}, 0
    for name, kind, cin, cout, k in UNITS:
        idx[name] = i
        i += 1 if kind == "up" else 5
    idx["fc1_W"], idx["fc1_bn"], idx["linear1_W"], idx["linear1_b"] = i, i + 1, i + 5, i + 6
    return idx


def expected_shapes():
    shapes = []
    for name, kind, cin, cout, k in UNITS:
        if kind == "up":
            shapes.append((1, 1, k, k, k))
            continue
        shapes.append((cout, cin, k, k, k) if kind == "conv" else (cin, cout, k, k, k))
        shapes += [(cout,)] * 4                                  # beta, gamma, mean, inv_std
    shapes.append((D_VIEWPAIR_FEATURE, SIMILNET_HIDDEN))
    shapes += [(SIMILNET_HIDDEN,)] * 4
    shapes += [(SIMILNET_HIDDEN, 1), (1,)]
    assert len(shapes) == N_ARRAYS
    return shapes


def validate(params):
    """Raise ValueError unless ``params`` is the 105-array list with the App. B shapes."""
    if len(params) != N_ARRAYS:
        raise ValueError("SurfaceNet parameter list must hold {} arrays, got {}".format(N_ARRAYS, len(params)))
    for i, (a, s) in enumerate(zip(params, expected_shapes())):
        if tuple(np.shape(a)) != s:
            raise ValueError("parameter {} has shape {}, expected {}".format(i, tuple(np.shape(a)), s))
    return [np.ascontiguousarray(a, dtype=np.float32) for a in params]


def load_model_file(model_file):
    """Read a reference ``.model`` file: a python-2 pickle of the flat array list
    (nets/SurfaceNet.py:397-399).  ``.npz`` with keys '0'..'104' is accepted too."""
    if model_file.endswith(".npz"):
        z = np.load(model_file)
        return validate([z[str(i)] for i in range(N_ARRAYS)])
    with open(model_file, "rb") as f:
        data = pickle.load(f, encoding="latin1")
    return validate(data)


def upsample_W(k):
    """The fixed separable taps 1-|i-c|/((k+1)//2) (nets/layers.py:363-374)."""
    factor = (k + 1) // 2
    t = 1.0 - np.abs(np.arange(k) - (factor - 1)) / factor
    return (t[:, None, None] * t[None, :, None] * t[None, None, :])[None, None].astype(np.float32)


_BN_FIXTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "synth_bn_seed{}.npz")


def synthetic_params(seed=0, calibrated=True):
    """Deterministic synthetic weights: conv W ~ N(0,1)/sqrt(C_in*k^3) from the
    frozen ``np.random.RandomState(seed)`` stream; BatchNorm statistics either identity or, when
    ``calibrated``, the per-channel vectors committed in ``data/synth_bn_seed<seed>.npz`` (measured
    once so that every layer is unit-variance on DTU scan9 CVCs and the output probability is spread;
    generator: tests/golden/make_synth_bn.py)."""
    rs = np.random.RandomState(seed)
    params = []
    for name, kind, cin, cout, k in UNITS:
        if kind == "up":
            params.append(upsample_W(k))
            continue
        shape = (cout, cin, k, k, k) if kind == "conv" else (cin, cout, k, k, k)
        W = (rs.standard_normal(shape) / np.sqrt(cin * k ** 3)).astype(np.float32)
        params += [W, np.zeros(cout, np.float32), np.ones(cout, np.float32),
                   np.zeros(cout, np.float32), np.ones(cout, np.float32)]
    params.append((rs.standard_normal((D_VIEWPAIR_FEATURE, SIMILNET_HIDDEN)) / np.sqrt(D_VIEWPAIR_FEATURE)).astype(np.float32))
    params += [np.zeros(SIMILNET_HIDDEN, np.float32), np.ones(SIMILNET_HIDDEN, np.float32),
               np.zeros(SIMILNET_HIDDEN, np.float32), np.ones(SIMILNET_HIDDEN, np.float32)]
    params.append((rs.standard_normal((SIMILNET_HIDDEN, 1)) / np.sqrt(SIMILNET_HIDDEN)).astype(np.float32))
    params.append(np.zeros(1, np.float32))
    if calibrated:
        path = _BN_FIXTURE.format(seed)
        if not os.path.exists(path):
            raise FileNotFoundError("no calibrated BatchNorm fixture for seed {} ({})".format(seed, path))
        z = np.load(path)
        idx = unit_index()
        for name, kind, cin, cout, k in UNITS:
            if kind == "up":
                continue
            i = idx[name]
            for j, key in enumerate(("beta", "gamma", "mean", "inv_std")):
                params[i + 1 + j] = z[name + "." + key].astype(np.float32)
    return validate(params)
