"""Oracle restatement (torch-CPU, fp32) of the SurfaceNet inference graph.  Test infrastructure only.

PARITY UNPINNED for the convolutional network: the reference graph is Theano 0.9 + Lasagne@7992faa
on cuDNN-only layers (nets/SurfaceNet.py:2, nets/layers.py:8) and cannot execute in this image.
Restated from
  nets/SurfaceNet.py:18-76    __1viewPair_SurfaceNet__   (layer list)
  nets/SurfaceNet.py:84-100   __relativeWeight_net__
  nets/SurfaceNet.py:126,343-357  reshape + fusion, N_vp == 1 special case
  nets/layers.py:200-253      DilatedConv3DLayer (W stored (C_in, C_out, k,k,k), cross-correlation)
  nets/layers.py:321-339      ChannelPool_weightedAverage
  nets/layers.py:363-390      Bilinear_3DInterpolation (zero-stuff + fixed k^3 conv, 'same')
and Lasagne semantics: Conv3DDNNLayer(pad='same', flip_filters=False) = cross-correlation;
batch_norm() removes the bias and appends BatchNormLayer whose deterministic output is
(x - mean) * (gamma * inv_std) + beta, then the nonlinearity; Pool3DDNNLayer((2,2,2), stride=2) max;
Upscale3DLayer(mode='dilate') writes the input at [::f] of a zero tensor; DenseLayer = x.W + b.

``params`` is the flat list of 105 arrays in lasagne.layers.get_all_param_values order
(SURVEY.md App. B); index constants live in surfacenet_b200/weights.py (host-side layout table).
"""
import numpy as np
import torch
import torch.nn.functional as F

# (name, first index in the 105-array list, kind) -- App. B.  kind: c3 = 3x3x3 conv, c1 = 1x1x1 conv,
# d3 / d1 = DilatedConv3DLayer (W is (C_in, C_out, ...)), all followed by [beta, gamma, mean, inv_std].
LAYOUT = {
    "conv1_1": 0, "conv1_2": 5, "conv1_3": 10, "side_op1": 15,
    "conv2_1": 20, "conv2_2": 25, "conv2_3": 30, "side_op2": 35, "up2_W": 40,
    "conv3_1": 41, "conv3_2": 46, "conv3_3": 51, "side_op3": 56, "up3_W": 61,
    "conv4_1": 62, "conv4_2": 67, "conv4_3": 72, "side_op4": 77, "up4_W": 82,
    "merge_conv": 83, "merge_conv2": 88, "merge_conv3": 93,
    "fc1_W": 98, "fc1_bn": 99, "linear1_W": 103, "linear1_b": 104,
}


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))


def _bn_act(x, p, i, act):
    beta, gamma, mean, inv_std = (_t(p[i + k]) for k in (1, 2, 3, 4))
    sh = (1, -1) + (1,) * (x.dim() - 2)
    y = (x - mean.view(sh)) * (gamma * inv_std).view(sh) + beta.view(sh)    # lasagne BatchNormLayer, deterministic
    return torch.relu(y) if act == "relu" else torch.sigmoid(y)


def conv_bn(x, p, i, act, dilated=False, taps=None):
    """One `batch_norm(Conv3DDNNLayer(..., pad='same'))` or `batch_norm(DilatedConv3DLayer(PadLayer(2)))` unit."""
    W = _t(p[i])
    if dilated:
        W = W.permute(1, 0, 2, 3, 4).contiguous()           # (C_in,C_out,k,k,k) -> (C_out,C_in,k,k,k), layers.py:200-253
        k = W.shape[-1]
        y = F.conv3d(x, W, padding=2 * (k // 2), dilation=2)  # PadLayer(width=2) + dilation (2,2,2): SurfaceNet.py:60-67
    else:
        k = W.shape[-1]
        y = F.conv3d(x, W, padding=k // 2)                    # pad='same', cross-correlation
    if taps is not None:
        taps["pre"] = y
    return _bn_act(y, p, i, act)


def upsample(x, W, f):
    """Bilinear_3DInterpolation (layers.py:376-390): zero-stuff by f (Upscale3DLayer mode='dilate'),
    fold channels into batch, conv with the fixed (1,1,k,k,k) W, 'same' pad, no bias."""
    n, c, d, h, w = x.shape
    up = torch.zeros((n, c, d * f, h * f, w * f), dtype=x.dtype)
    up[:, :, ::f, ::f, ::f] = x
    W = _t(W)
    k = W.shape[-1]
    y = F.conv3d(up.reshape(n * c, 1, d * f, h * f, w * f), W, padding=k // 2)
    return y.reshape(n, c, d * f, h * f, w * f)


def one_viewpair_forward(X, p, return_taps=False):
    """nets/SurfaceNet.py:18-76.  X (N,6,s,s,s) f32 (mean already subtracted) -> (N,1,s,s,s)."""
    L = LAYOUT
    x = _t(X) if isinstance(X, np.ndarray) else X
    taps = {}
    c11 = conv_bn(x, p, L["conv1_1"], "relu")
    c12 = conv_bn(c11, p, L["conv1_2"], "relu")
    c13 = conv_bn(c12, p, L["conv1_3"], "relu")
    pool1 = F.max_pool3d(c13, 2, 2)
    s1 = conv_bn(c13, p, L["side_op1"], "sigmoid")
    c21 = conv_bn(pool1, p, L["conv2_1"], "relu")
    c22 = conv_bn(c21, p, L["conv2_2"], "relu")
    c23 = conv_bn(c22, p, L["conv2_3"], "relu")
    pool2 = F.max_pool3d(c23, 2, 2)
    s2 = conv_bn(c23, p, L["side_op2"], "sigmoid")
    s2u = upsample(s2, p[L["up2_W"]], 2)
    c31 = conv_bn(pool2, p, L["conv3_1"], "relu")
    c32 = conv_bn(c31, p, L["conv3_2"], "relu")
    c33 = conv_bn(c32, p, L["conv3_3"], "relu")
    s3 = conv_bn(c33, p, L["side_op3"], "sigmoid")
    s3u = upsample(s3, p[L["up3_W"]], 4)
    c41 = conv_bn(c33, p, L["conv4_1"], "relu", dilated=True)
    c42 = conv_bn(c41, p, L["conv4_2"], "relu", dilated=True)
    c43 = conv_bn(c42, p, L["conv4_3"], "relu", dilated=True)
    s4 = conv_bn(c43, p, L["side_op4"], "sigmoid", dilated=True)
    s4u = upsample(s4, p[L["up4_W"]], 4)
    cat = torch.cat([s1, s2u, s3u, s4u], dim=1)                              # SurfaceNet.py:71
    m1 = conv_bn(cat, p, L["merge_conv"], "relu")
    m2 = conv_bn(m1, p, L["merge_conv2"], "relu")
    out = conv_bn(m2, p, L["merge_conv3"], "sigmoid")
    if return_taps:
        taps.update(conv1_1=c11, conv1_2=c12, conv1_3=c13, pool1=pool1, side_op1=s1, conv2_1=c21, conv2_2=c22,
                    conv2_3=c23, pool2=pool2, side_op2=s2, side_op2_up=s2u, conv3_1=c31, conv3_2=c32, conv3_3=c33,
                    side_op3=s3, side_op3_up=s3u, conv4_1=c41, conv4_2=c42, conv4_3=c43, side_op4=s4,
                    side_op4_up=s4u, concat=cat, merge_conv=m1, merge_conv2=m2, out=out)
        return out, taps
    return out


def nViewPair_SurfaceNet_fn(X, p, w=None, N_vp=1, chunk=8):
    """What the compiled Theano callable returns (SurfaceNet.py:343-376): [fused (B,1,s,s,s), unfused (B,N_vp,s,s,s)].
    N_vp == 1: both are the raw network output (SurfaceNet.py:354-357)."""
    with torch.no_grad():
        outs = [one_viewpair_forward(X[i:i + chunk], p) for i in range(0, X.shape[0], chunk)]
        out = torch.cat(outs, 0)
        if N_vp == 1:
            o = out.numpy()
            return [o, o]
        s = out.shape[-1]
        unf = out.reshape(-1, N_vp, s, s, s)                                 # SurfaceNet.py:126
        wt = _t(w)
        cw = wt / wt.sum(dim=1, keepdim=True)                                # layers.py:330-331
        fused = (unf * cw[:, :, None, None, None]).sum(dim=1, keepdim=True)  # layers.py:334-335
        return [fused.numpy(), unf.numpy()]


def viewPair_relativeImpt_fn(features, p, n_samples_perGroup):
    """nets/SurfaceNet.py:84-100: softmax_over_group(Dense1(sigmoid(BN(Dense100(f)))))."""
    L = LAYOUT
    with torch.no_grad():
        f = _t(features)
        h = f @ _t(p[L["fc1_W"]])                                            # batch_norm() removed the bias
        beta, gamma, mean, inv_std = (_t(p[L["fc1_bn"] + k]) for k in range(4))
        h = torch.sigmoid((h - mean) * (gamma * inv_std) + beta)
        o = h @ _t(p[L["linear1_W"]]) + _t(p[L["linear1_b"]])
        o = o.reshape(-1, n_samples_perGroup)
        return torch.softmax(o, dim=1).numpy()


def W_5D(size):
    """nets/layers.py:363-374 (__W_5D__), py3: `np.ogrid[:size]` needs ints."""
    size = float(size)
    factor = (size + 1) // 2
    center = factor - 1 if size % 2 == 1 else factor - 0.5
    n = int(size)
    og = np.ogrid[:n, :n, :n]
    W = (1 - abs(og[0] - center) / factor) * (1 - abs(og[1] - center) / factor) * (1 - abs(og[2] - center) / factor)
    return W[None, None].astype(np.float32)
