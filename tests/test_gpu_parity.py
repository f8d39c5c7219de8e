"""GPU parity tests proper: the CUDA path (through the C ABI, libsurfacenet_b200.so) against the CPU
oracle on the same seeded inputs and against the golden fixtures produced by the reference code.

Bars: bit-exact for the voxel->pixel index map, the gathered colours and the ray-pool votes;
<= 1e-4 max-abs on probabilities (BASELINE.json north_star), tolerance written at each assert.
"""
import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu

PROB_TOL = 1e-4          # north_star: <= 1e-4 max-abs on the surface-probability volume


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


@pytest.fixture(scope="module")
def params():
    from surfacenet_b200 import weights
    return weights.synthetic_params(0)


@pytest.fixture(scope="module")
def net(torch_cuda, params):
    from surfacenet_b200 import SurfaceNet
    return SurfaceNet.Net(params)


# ---- perspectiveProj ---------------------------------------------------------------------------------
def test_perspectiveProj_reference_doctest(torch_cuda):
    from surfacenet_b200 import camera
    np.random.seed(201611)                                    # utils/camera.py:144-160
    Ms = np.random.rand(2, 3, 4)
    pts_3D = np.random.rand(2, 3)
    h, w = camera.perspectiveProj(Ms, pts_3D, return_int_hw=False)
    assert np.allclose(w, np.array([[1.35860185, 0.9878389], [0.64522543, 0.76079278]]))
    hi, wi = camera.perspectiveProj(Ms, pts_3D, return_int_hw=True)
    assert hi.dtype == np.int64 and np.array_equal(wi, np.array([[1, 1], [1, 1]]))
    h1, w1 = camera.perspectiveProj(Ms[1], pts_3D[0], return_int_hw=False)
    assert np.allclose(np.r_[h1, w1], np.stack((h, w))[:, 1, 0])
    with pytest.raises(ValueError):
        camera.perspectiveProj(np.zeros((4, 4)), np.zeros((2, 3)))
    with pytest.raises(ValueError):
        camera.perspectiveProj(np.zeros((3, 4)), np.zeros((2, 2)))


def test_perspectiveProj_golden_bit_exact(torch_cuda, golden, cams):
    from surfacenet_b200 import camera
    h, w, d = camera.perspectiveProj(cams[golden["pp_dtu_views"]], golden["pp_dtu_pts"], True, True)
    assert np.array_equal(h, golden["pp_dtu_h"]) and np.array_equal(w, golden["pp_dtu_w"])
    assert np.array_equal(d, golden["pp_dtu_depth"])           # fp64 depth, bit-exact


# ---- CVC -----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["basic", "ragged_sizes", "dup_views", "c1_s32", "outside"])
def test_cvc_matches_reference_outputs(torch_cuda, golden, cams, name):
    from surfacenet_b200 import CVC
    case = util.cvc_cases(cams)[name]
    X = CVC.gen_coloredCubes(case["pairs"], case["xyz"], case["resol"], case["cameraPOs"], case["images"], case["D"])
    assert X.dtype == np.float32 and X.shape == golden["cvc_" + name].shape
    assert np.array_equal(X, golden["cvc_" + name].astype(np.float32))            # exact colours
    _, X2 = CVC.preprocess_augmentation(None, X, util.MEAN6[None, :, None, None, None], False, False)
    assert np.array_equal(X2.reshape(-1)[::997], golden["cvc_" + name + "_pre_sample"])


@pytest.mark.parametrize("name", ["basic", "ragged_sizes", "outside"])
def test_cvc_index_map_bit_exact(torch_cuda, cams, name):
    from oracle import cvc_oracle
    from surfacenet_b200 import CVC
    from surfacenet_b200.device import DeviceScene
    case = util.cvc_cases(cams)[name]
    scene = DeviceScene(case["cameraPOs"], case["images"])
    X, iw, ih, ins = CVC.gen_coloredCubes_device(scene, case["pairs"], case["xyz"], case["resol"], case["D"], return_index=True)
    ow, oh, oin = cvc_oracle.gen_index_map(case["pairs"], case["xyz"], case["resol"], case["cameraPOs"], case["images"], case["D"])
    assert iw.dtype == torch_cuda.int32
    assert np.array_equal(iw.cpu().numpy(), ow), "w index map differs"
    assert np.array_equal(ih.cpu().numpy(), oh), "h index map differs"
    assert np.array_equal(ins.cpu().numpy().astype(bool), oin)


def test_cvc_full_size_s64_index_and_colour(torch_cuda, cams):
    """BASELINE full size (s=64): compare with the oracle on 2 cubes x 5 pairs (oracle ~2 s)."""
    from oracle import cvc_oracle
    from surfacenet_b200 import CVC
    rs = np.random.RandomState(7)
    used = list(range(0, 49, 4))
    imgs = util.image_list(49, used)
    pairs = rs.choice(used, size=(2, 5, 2))
    xyz = np.array([[10.0, -30.0, 620.0], [30.0, 0.0, 650.0]], np.float32)
    resol = np.full(2, 0.4, np.float32)
    X = CVC.gen_coloredCubes(pairs, xyz, resol, cams, imgs, 64)
    Xo = cvc_oracle.gen_coloredCubes(pairs, xyz, resol, cams, imgs, 64)
    assert np.array_equal(X, Xo)


def test_cvc_errors(torch_cuda, cams):
    from surfacenet_b200 import CVC
    imgs = util.image_list(49, [0, 1])
    with pytest.raises(ValueError):
        CVC.gen_coloredCubes(np.zeros((1, 2), np.int64), np.zeros((1, 3), np.float32), np.ones(1, np.float32), cams, imgs, 8)
    with pytest.raises(ValueError):     # view without an image
        CVC.gen_coloredCubes(np.array([[[0, 7]]]), np.zeros((1, 3), np.float32), np.ones(1, np.float32), cams, imgs, 8)
    out = CVC.gen_coloredCubes(np.zeros((0, 1, 2), np.int64), np.zeros((0, 3), np.float32), np.ones(0, np.float32), cams, imgs, 8)
    assert out.shape == (0, 6, 8, 8, 8)


# ---- ray pooling ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["sheet16", "sheet32_dup", "ties", "lowres_collide", "all_ones", "empty",
                                  "none_thresh_f32", "exact_thresh"])
def test_raypool_matches_reference_outputs(torch_cuda, golden, cams, name):
    from surfacenet_b200 import rayPooling
    case = util.raypool_cases(cams)[name]
    votes = rayPooling.rayPooling_1cube_numpy(case["cameraPOs"], None, case["pred"], case["pairs"], case["xyz"],
                                              case["resol"], prediction_thresh=case["thresh"])
    assert votes.dtype == np.int64
    assert np.array_equal(votes.astype(np.uint8), golden["rp_" + name])


def test_raypool_s64_batched_vs_oracle(torch_cuda, cams):
    """Full-size cubes, batched call, 5 pairs with a duplicated view: exact vote equality."""
    import torch
    from oracle import raypool_oracle
    from surfacenet_b200 import rayPooling
    preds = np.stack([util.sheet_prediction(64, 0.3 * i, 0.06) for i in range(3)])
    pairs = np.array([[[0, 5], [17, 22], [5, 30], [40, 41], [8, 9]],
                      [[1, 2], [3, 4], [5, 6], [7, 8], [9, 10]],
                      [[12, 12], [12, 13], [20, 44], [45, 46], [47, 48]]], np.int32)
    xyz = np.array([[20.0, -12.5, 630.0], [-10.3, 30.7, 655.1], [55.25, -60.0, 600.5]], np.float32)
    resol = np.full(3, 0.4, np.float32)
    votes = rayPooling.votes_device(torch.from_numpy(preds).cuda(), torch.from_numpy(pairs).cuda(), torch.from_numpy(xyz).cuda(),
                                    torch.from_numpy(resol).cuda(), torch.from_numpy(cams).cuda(), cams.shape[0], 0.46).cpu().numpy()
    for b in range(3):
        ref = raypool_oracle.rayPooling_1cube_numpy(cams, None, preds[b], pairs[b], xyz[b], resol[b], prediction_thresh=0.46)
        assert np.array_equal(votes[b], ref.astype(np.uint8)), "cube %d votes differ" % b
    assert votes.max() <= 10 and votes.max() >= 2


def test_raypool_errors(torch_cuda, cams):
    from surfacenet_b200 import rayPooling
    with pytest.raises(ValueError):
        rayPooling.rayPooling_1cube_numpy(cams, None, np.zeros((2, 2, 4, 4, 4)), np.array([[0, 1]]), np.zeros(3, np.float32), np.float32(1))
    with pytest.raises(ValueError):     # domain: selected predictions must be > 0
        rayPooling.rayPooling_1cube_numpy(cams, None, np.zeros((4, 4, 4), np.float32), np.array([[0, 1]]), np.zeros(3, np.float32),
                                          np.float32(1), prediction_thresh=None)


# ---- network: single layers ----------------------------------------------------------------------------
def _layer_in(rs, n, C, S):
    return (rs.standard_normal((n, C, S, S, S)) * 1.5).astype(np.float32)


CONV_CASES = [("conv1_1", 12), ("conv1_2", 9), ("side_op1", 8), ("conv2_1", 8), ("conv3_2", 6), ("conv4_1", 7), ("conv4_2", 8),
              ("side_op4", 5), ("merge_conv", 8), ("merge_conv2", 8), ("merge_conv3", 8), ("conv1_3", 20), ("conv4_3", 16), ("conv2_2", 17)]
# max-abs tolerance relative to max(1, |ref|max): fp32 = accumulation-order noise; exact = fp16 hi+lo split operands
# (22-bit products, fp32 accumulate); fast = single fp16 rounding of both operands (NOT a parity mode)
CONV_TOL = {"fp32": 2e-5, "exact": 3e-5, "fast": 2e-2}


@pytest.mark.parametrize("mode", ["fp32", "exact", "fast"])
@pytest.mark.parametrize("name,S", CONV_CASES)
def test_conv_units(torch_cuda, net, params, name, S, mode):
    """conv + BatchNorm + activation units, odd sizes included (tile edges), vs torch-CPU fp32."""
    import torch
    from oracle import surfacenet_oracle as so
    from surfacenet_b200 import _lib, weights
    names = [u[0] for u in weights.UNITS]
    u = names.index(name)
    _, kind, cin, cout, k = weights.UNITS[u]
    rs = np.random.RandomState(u)
    x = _layer_in(rs, 2, cin, S)
    act = "sigmoid" if name in weights.SIGMOID_UNITS else "relu"
    with torch.no_grad():
        ref = so.conv_bn(torch.from_numpy(x), params, weights.unit_index()[name], act, dilated=(kind == "dil")).numpy()
    xd = torch.from_numpy(x).cuda()
    out = torch.full((2, cout, S, S, S), float("nan"), dtype=torch.float32, device="cuda")
    _lib.check(_lib.lib.sn_net_layer_conv(net.handle, u, _lib.ptr(xd), 2, S, _lib.ptr(out), _lib.MODES[mode], _lib.stream_ptr()))
    torch.cuda.synchronize()
    o = out.cpu().numpy()
    assert np.isfinite(o).all(), "%s: unwritten / non-finite outputs" % name
    err = np.abs(o - ref).max()
    scale = max(1.0, np.abs(ref).max())
    print("%s S=%d %s: max-abs %.3g (ref max %.3g)" % (name, S, mode, err, scale))
    assert err <= CONV_TOL[mode] * scale, "%s: max-abs %g" % (name, err)


# sizes with a Winograd F(2,3) instance (csrc/conv_wg.cu: S in {8, 16, 32, 64}; rows of S/2 output pairs, N tiles 32 / 80 / 112)
CONV_CASES_WG = [("conv1_1", 16), ("conv1_2", 32), ("conv1_3", 16), ("conv2_1", 16), ("conv2_2", 32), ("conv3_1", 16), ("conv3_2", 16),
                 ("merge_conv", 16), ("merge_conv2", 16), ("merge_conv2", 32), ("conv1_2", 64), ("conv4_1", 16), ("conv4_2", 16),
                 ("conv4_3", 32), ("conv3_2", 8), ("conv4_1", 8), ("conv2_2", 8), ("merge_conv", 8)]     # S = 8: 4 planes x 8 rows x 4 pairs per tile


@pytest.mark.parametrize("name,S", CONV_CASES_WG)
def test_conv_units_winograd(torch_cuda, net, params, name, S):
    """The 3x3x3 units (dilated conv4_x included) at the sizes the exact mode runs as w-axis Winograd F(2,3) (4 frequency passes,
    output transform in the epilogue registers), vs torch-CPU fp32; same tolerance as the direct kernel.  n = 3 samples so that the persistent tile loop
    wraps around the accumulator double buffer with an odd tile count."""
    import torch
    from oracle import surfacenet_oracle as so
    from surfacenet_b200 import _lib, weights
    names = [u[0] for u in weights.UNITS]
    u = names.index(name)
    _, kind, cin, cout, k = weights.UNITS[u]
    n = 3 if S < 64 else 1
    rs = np.random.RandomState(100 + u)
    x = _layer_in(rs, n, cin, S)
    with torch.no_grad():
        ref = so.conv_bn(torch.from_numpy(x), params, weights.unit_index()[name], "relu", dilated=(kind == "dil")).numpy()
    xd = torch.from_numpy(x).cuda()
    out = torch.full((n, cout, S, S, S), float("nan"), dtype=torch.float32, device="cuda")
    _lib.check(_lib.lib.sn_net_layer_conv(net.handle, u, _lib.ptr(xd), n, S, _lib.ptr(out), _lib.MODES["exact"], _lib.stream_ptr()))
    torch.cuda.synchronize()
    o = out.cpu().numpy()
    assert np.isfinite(o).all(), "%s: unwritten / non-finite outputs" % name
    err = np.abs(o - ref).max()
    scale = max(1.0, np.abs(ref).max())
    print("%s S=%d winograd: max-abs %.3g (ref max %.3g)" % (name, S, err, scale))
    assert err <= CONV_TOL["exact"] * scale, "%s: max-abs %g" % (name, err)


@pytest.mark.parametrize("unit,f,S", [("up2", 2, 5), ("up3", 4, 3), ("up4", 4, 4)])
def test_upsample_units(torch_cuda, net, params, unit, f, S):
    import torch
    from oracle import surfacenet_oracle as so
    from surfacenet_b200 import _lib, weights
    names = [u[0] for u in weights.UNITS]
    u = names.index(unit)
    rs = np.random.RandomState(3)
    x = rs.rand(2, 16, S, S, S).astype(np.float32)
    ref = so.upsample(torch.from_numpy(x), params[weights.unit_index()[unit]], f).numpy()
    out = torch.zeros((2, 64, S * f, S * f, S * f), dtype=torch.float32, device="cuda")
    _lib.check(_lib.lib.sn_net_layer_upsample(net.handle, u, _lib.ptr(torch.from_numpy(x).cuda()), 2, 16, S, _lib.ptr(out), 64, 32,
                                              _lib.stream_ptr()))
    o = out.cpu().numpy()
    assert np.abs(o[:, 32:48] - ref).max() <= 1e-6
    assert o[:, :32].max() == 0 and o[:, 48:].max() == 0


def test_upsample_micro_cases_device(torch_cuda, net):
    """[1,2] -> x4 [1, 2/3, 1, 4/3, 2, 4/3, 2/3, 0];  x2 [1, 1.5, 2, 1]."""
    import torch
    from surfacenet_b200 import _lib, weights
    names = [u[0] for u in weights.UNITS]
    x = torch.zeros(1, 1, 2, 2, 2); x[0, 0, :, 0, 0] = torch.tensor([1.0, 2.0])
    for unit, f, want in (("up3", 4, [1, 2 / 3, 1.0, 4 / 3, 2, 4 / 3, 2 / 3, 0]), ("up2", 2, [1, 1.5, 2, 1.0])):
        out = torch.zeros((1, 1, 2 * f, 2 * f, 2 * f), dtype=torch.float32, device="cuda")
        xd = x.cuda()
        _lib.check(_lib.lib.sn_net_layer_upsample(net.handle, names.index(unit), _lib.ptr(xd), 1, 1, 2, _lib.ptr(out), 1, 0,
                                                  _lib.stream_ptr()))
        assert np.allclose(out[0, 0, :, 0, 0].cpu().numpy(), want, atol=1e-6)


def test_maxpool_and_fusion(torch_cuda):
    import torch
    import torch.nn.functional as F
    from surfacenet_b200 import _lib
    rs = np.random.RandomState(0)
    x = torch.from_numpy(rs.standard_normal((2, 5, 6, 6, 6)).astype(np.float32))
    out = torch.empty((2, 5, 3, 3, 3), dtype=torch.float32, device="cuda")
    xd = x.cuda()                                           # keep the device tensors alive across the raw-pointer calls
    _lib.check(_lib.lib.sn_maxpool2(_lib.ptr(xd), 2, 5, 6, _lib.ptr(out), _lib.stream_ptr()))
    assert torch.equal(out.cpu(), F.max_pool3d(x, 2, 2))
    p = torch.from_numpy(rs.rand(3, 4, 100).astype(np.float32))
    w = torch.from_numpy((rs.rand(3, 4) + 0.1).astype(np.float32))
    fo = torch.empty((3, 100), dtype=torch.float32, device="cuda")
    pd, wd = p.cuda(), w.cuda()
    _lib.check(_lib.lib.sn_fuse_weighted_average(_lib.ptr(pd), _lib.ptr(wd), 3, 4, 100, _lib.ptr(fo), _lib.stream_ptr()))
    ref = (p * (w / w.sum(1, keepdim=True))[:, :, None]).sum(1)
    assert (fo.cpu() - ref).abs().max() <= 1e-6


# ---- network: whole forward -----------------------------------------------------------------------------
def _real_like_X(cams, D, n_cubes=2, n_vp=2, seed=0):
    from oracle import cvc_oracle
    rs = np.random.RandomState(seed)
    used = [8, 9, 22, 23, 30, 33]
    imgs = util.image_list(49, used)
    pairs = rs.choice(used, size=(n_cubes, n_vp, 2))
    xyz = (np.array([10.0, -30.0, 620.0]) + rs.rand(n_cubes, 3) * 20).astype(np.float32)
    resol = np.full(n_cubes, 0.4, np.float32)
    X = cvc_oracle.gen_coloredCubes(pairs, xyz, resol, cams, imgs, D)
    _, X = cvc_oracle.preprocess_augmentation(None, X, util.MEAN6[None, :, None, None, None], False, False)
    return X, pairs, xyz, resol, imgs


@pytest.mark.parametrize("mode", ["fp32", "exact"])
def test_forward_s32_two_pairs(torch_cuda, params, cams, mode):
    """Whole network + weighted fusion on CVC input, s=32, 2 cubes x 2 pairs vs the torch-CPU fp32 oracle."""
    from oracle import surfacenet_oracle as so
    from surfacenet_b200 import SurfaceNet
    X, *_ = _real_like_X(cams, 32)
    w = np.array([[0.7, 0.2], [0.3, 0.9]], np.float32)
    fused_o, unf_o = so.nViewPair_SurfaceNet_fn(X, params, w, N_vp=2)
    _, fn = SurfaceNet.SurfaceNet_inference(2, params, ["output_SurfaceNet_reshape", "output_softmaxWeights"], mode=mode)
    fused, unf = fn(X, w)
    assert fused.shape == (2, 1, 32, 32, 32) and unf.shape == (2, 2, 32, 32, 32) and fused.dtype == np.float32
    e_f, e_u = np.abs(fused - fused_o).max(), np.abs(unf - unf_o).max()
    print("mode %s: max-abs fused %.3g unfused %.3g; prob mean %.3f frac>0.46 %.3f" % (mode, e_f, e_u, fused_o.mean(), (fused_o > 0.46).mean()))
    assert e_f <= PROB_TOL and e_u <= PROB_TOL
    assert 0.05 < (fused_o > 0.46).mean() < 0.95            # the synthetic weights give a spread output


@pytest.mark.parametrize("mode", ["fp32", "exact"])
def test_forward_single_pair_and_odd_batch(torch_cuda, params, cams, mode):
    """N_vp == 1: fused and unfused are the same tensor (SurfaceNet.py:354-357); s=16, 3 cubes."""
    from oracle import surfacenet_oracle as so
    from surfacenet_b200 import SurfaceNet
    X, *_ = _real_like_X(cams, 16, n_cubes=3, n_vp=1, seed=2)
    o, _ = so.nViewPair_SurfaceNet_fn(X, params, None, N_vp=1)
    _, fn = SurfaceNet.SurfaceNet_inference(1, params, mode=mode)
    fused, unf = fn(X)
    assert fused.shape == (3, 1, 16, 16, 16) and unf is fused
    assert np.abs(fused - o).max() <= PROB_TOL


def test_forward_s64_full_size_exact(torch_cuda, params, cams):
    """BASELINE full cube size (64^3), 1 cube x 2 pairs, tensor-core exact mode vs the torch-CPU fp32 oracle."""
    from oracle import surfacenet_oracle as so
    from surfacenet_b200 import SurfaceNet
    X, *_ = _real_like_X(cams, 64, n_cubes=1, n_vp=2, seed=11)
    w = np.array([[0.35, 0.8]], np.float32)
    fused_o, unf_o = so.nViewPair_SurfaceNet_fn(X, params, w, N_vp=2, chunk=1)
    _, fn = SurfaceNet.SurfaceNet_inference(2, params, mode="exact")
    fused, unf = fn(X, w)
    e_f, e_u = np.abs(fused - fused_o).max(), np.abs(unf - unf_o).max()
    print("s=64 exact: max-abs fused %.3g unfused %.3g" % (e_f, e_u))
    assert e_f <= PROB_TOL and e_u <= PROB_TOL


# ---- parity as a distribution: weight seeds x view-pair counts x cubes --------------------------------------
def _survey_case(cams, params, D, n_cubes, n_vp, seed):
    """-> (max-abs fused, max-abs unfused) of the exact mode vs the torch-CPU fp32 oracle on CVC input."""
    from oracle import surfacenet_oracle as so
    from surfacenet_b200 import SurfaceNet
    X, *_ = _real_like_X(cams, D, n_cubes=n_cubes, n_vp=n_vp, seed=1000 + 17 * seed + n_vp)
    w = (np.random.RandomState(seed * 7 + n_vp).rand(n_cubes, n_vp) + 0.1).astype(np.float32)
    fused_o, unf_o = so.nViewPair_SurfaceNet_fn(X, params, w if n_vp > 1 else None, N_vp=n_vp, chunk=4 if D <= 32 else 1)
    _, fn = SurfaceNet.SurfaceNet_inference(n_vp, params)                 # the documented drop-in call: default mode
    fused, unf = fn(X, w) if n_vp > 1 else fn(X)
    return float(np.abs(fused - fused_o).max()), float(np.abs(unf - unf_o).max())


SURVEY_64 = [(0, 1), (1, 1), (2, 1), (3, 1), (4, 1), (1, 5), (0, 8), (3, 8)]        # (weight seed, N_vp) at the BASELINE cube size, 1 cube each


@pytest.mark.parametrize("seed,n_vp", SURVEY_64)
def test_parity_distribution_s64(torch_cuda, cams, seed, n_vp):
    """Five synthetic weight sets (BatchNorm calibrated on real DTU CVCs per seed: tests/golden/make_synth_bn.py), N_vp in {1, 5, 8},
    64^3 cubes: fused AND unfused probabilities within 1e-4 of the oracle (nets/layers.py:321-339 fusion, SurfaceNet.py:126,343-357)."""
    from surfacenet_b200 import weights
    e_f, e_u = _survey_case(cams, weights.synthetic_params(seed), 64, 1, n_vp, seed)
    print("s=64 seed %d N_vp %d: max-abs fused %.3g unfused %.3g" % (seed, n_vp, e_f, e_u))
    assert e_f <= PROB_TOL and e_u <= PROB_TOL


@pytest.mark.parametrize("seed", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("n_vp", [5, 8])
def test_parity_distribution_s32_four_cubes(torch_cuda, cams, seed, n_vp):
    """Same survey at s=32 with 4 random cubes per case (mixed path: Winograd levels at 32 / 16, direct kernels at 8)."""
    from surfacenet_b200 import weights
    e_f, e_u = _survey_case(cams, weights.synthetic_params(seed), 32, 4, n_vp, seed)
    print("s=32 seed %d N_vp %d x 4 cubes: max-abs fused %.3g unfused %.3g" % (seed, n_vp, e_f, e_u))
    assert e_f <= PROB_TOL and e_u <= PROB_TOL


@pytest.mark.parametrize("name,S", [("merge_conv2", 32), ("conv4_2", 16), ("conv2_2", 17), ("conv4_1", 7)])
@pytest.mark.parametrize("kind", ["all_positive", "zero_mean", "one_sided_large"])
def test_accumulation_truncation_extremes(torch_cuda, params, name, S, kind):
    """tcgen05 accumulates with round-toward-zero; the expected loss is folded back into the BatchNorm scale (kRzLoss, conv_tc.cu /
    conv_wg.cu).  Bound that statistical compensation on adversarial operands: all-positive weights and activations (every partial sum
    grows monotonically: the largest possible truncation bias), zero-mean operands (no bias to compensate), and large one-sided
    activations; Winograd units (S in {16,32}) and direct units (odd S), against an fp64 evaluation."""
    import torch
    import torch.nn.functional as F
    from surfacenet_b200 import SurfaceNet, _lib, weights
    names = [u[0] for u in weights.UNITS]
    u = names.index(name)
    _, ukind, cin, cout, k = weights.UNITS[u]
    i0 = weights.unit_index()[name]
    rs = np.random.RandomState(31 + u)
    p2 = [a.copy() for a in params]
    W = p2[i0]
    x = rs.standard_normal((1, cin, S, S, S)).astype(np.float32) * 1.5
    if kind == "all_positive":
        p2[i0] = np.abs(W); x = np.abs(x)
    elif kind == "one_sided_large":
        x = (np.abs(x) * 40.0 + 10.0).astype(np.float32)
    # identity BatchNorm so that the pre-activation itself is compared: beta 0, gamma 1, mean 0, inv_std 1
    p2[i0 + 1][:] = 0; p2[i0 + 2][:] = 1; p2[i0 + 3][:] = 0; p2[i0 + 4][:] = 1
    net2 = SurfaceNet.Net(p2)
    Wt = torch.from_numpy(p2[i0]).double()
    dil = ukind == "dil"
    if dil:
        Wt = Wt.permute(1, 0, 2, 3, 4).contiguous()
    ref = torch.relu(F.conv3d(torch.from_numpy(x).double(), Wt, padding=(2 if dil else 1) * (k // 2), dilation=2 if dil else 1)).numpy()
    xd = torch.from_numpy(x).cuda()
    out = torch.full((1, cout, S, S, S), float("nan"), dtype=torch.float32, device="cuda")
    _lib.check(_lib.lib.sn_net_layer_conv(net2.handle, u, _lib.ptr(xd), 1, S, _lib.ptr(out), _lib.MODES["exact"], _lib.stream_ptr()))
    torch.cuda.synchronize()
    err = np.abs(out.cpu().numpy() - ref).max()
    scale = max(1.0, np.abs(ref).max())
    print("%s S=%d %s: max-abs %.3g / ref max %.3g = %.3g" % (name, S, kind, err, scale, err / scale))
    # measured on B200: zero-mean <= 2.9e-6, one-sided activations <= 3.6e-6 (what a BatchNorm + ReLU network feeds its units), all-positive
    # weights AND activations 4.5e-6 (Winograd, 63 MMAs per accumulator) .. 2.0e-5 (direct kernel, 270 MMAs): the worst case of the
    # statistical compensation stays inside the per-unit tolerance of the exact mode
    tol = CONV_TOL["exact"] if kind == "all_positive" else 5e-6
    assert err <= tol * scale, "%s %s: relative error %g" % (name, kind, err / scale)


def test_documented_import_swap_runs_the_tensor_core_kernels(torch_cuda, params, cams):
    """`from surfacenet_b200 import SurfaceNet` + SurfaceNet_inference(N, model) with NO mode argument must run
    the tcgen05 kernels (conv_wg / conv_tc), not the CUDA-core fp32 cross-check path."""
    from surfacenet_b200 import SurfaceNet, _lib
    X, *_ = _real_like_X(cams, 16, n_cubes=1, n_vp=1, seed=5)
    _, fn = SurfaceNet.SurfaceNet_inference(1, params)
    import ctypes as C
    _lib.lib.sn_launch_count_reset()
    fn(X)
    counts = (C.c_int64 * 3)()
    _lib.lib.sn_conv_path_counts(counts)
    # 22 conv units: side_op1 runs as a fused CUDA-core pass into the Winograd layout, merge_conv3 inside merge_conv2's epilogue, the rest on tcgen05
    assert counts[0] == 0 and counts[1] + counts[2] >= 16 and counts[2] >= 5, "default call ran fp32/direct/winograd = %s" % list(counts)
    _, fn32 = SurfaceNet.SurfaceNet_inference(1, params, mode="fp32")
    _lib.lib.sn_launch_count_reset()
    fn32(X)
    _lib.lib.sn_conv_path_counts(counts)
    assert counts[0] >= 18 and counts[1] == 0 and counts[2] == 0


def test_forward_fast_mode_reports_error(torch_cuda, params, cams):
    """Single-pass fp16 operands: explicitly NOT a parity mode; its error is measured and bounded loosely."""
    from oracle import surfacenet_oracle as so
    from surfacenet_b200 import SurfaceNet
    X, *_ = _real_like_X(cams, 32, n_cubes=1, n_vp=2, seed=4)
    w = np.array([[0.5, 0.5]], np.float32)
    fused_o, _ = so.nViewPair_SurfaceNet_fn(X, params, w, N_vp=2)
    _, fn = SurfaceNet.SurfaceNet_inference(2, params, mode="fast")
    fused, _ = fn(X, w)
    err = np.abs(fused - fused_o).max()
    print("fast mode max-abs %.3g" % err)
    assert err <= 5e-2


def test_forward_errors(torch_cuda, params):
    from surfacenet_b200 import SurfaceNet
    _, fn = SurfaceNet.SurfaceNet_inference(2, params)
    with pytest.raises(TypeError):
        fn(np.zeros((2, 6, 8, 8, 8), np.float32))
    with pytest.raises(ValueError):
        fn(np.zeros((3, 6, 8, 8, 8), np.float32), np.ones((1, 2), np.float32))
    with pytest.raises(ValueError):
        fn(np.zeros((2, 5, 8, 8, 8), np.float32), np.ones((1, 2), np.float32))
    with pytest.raises(ValueError):     # two 2^3 poolings need D % 4 == 0
        fn(np.zeros((2, 6, 6, 6, 6), np.float32), np.ones((1, 2), np.float32))
    with pytest.raises(ValueError):
        SurfaceNet.Net(params[:50])


def test_relative_importance(torch_cuda, params):
    from oracle import surfacenet_oracle as so
    from surfacenet_b200 import SurfaceNet
    rs = np.random.RandomState(9)
    f = rs.standard_normal((4 * 6, 258)).astype(np.float32)
    imp_fn, _ = SurfaceNet.SurfaceNet_inference(2, params)
    out = imp_fn(f, n_samples_perGroup=6)
    ref = so.viewPair_relativeImpt_fn(f, params, 6)
    assert out.shape == (4, 6) and np.abs(out - ref).max() <= 1e-5
    assert np.allclose(out.sum(1), 1, atol=1e-5)
    with pytest.raises(ValueError):
        imp_fn(f, n_samples_perGroup=5)


# ---- the fused hot loop body ----------------------------------------------------------------------------
@pytest.mark.parametrize("mode", ["fp32", "exact"])
def test_infer_batch_host_matches_oracle_pipeline(torch_cuda, params, cams, mode):
    """main_reconstruct.py:134-162 as one call with host buffers: CVC -> net -> fusion -> f16 -> votes."""
    from oracle import cvc_oracle, raypool_oracle, surfacenet_oracle as so
    from surfacenet_b200 import SurfaceNet, pipeline
    from surfacenet_b200.device import DeviceScene
    D = 32
    X, pairs, xyz, resol, imgs = _real_like_X(cams, D, n_cubes=2, n_vp=3, seed=5)
    w = (np.random.RandomState(3).rand(2, 3) + 0.1).astype(np.float32)
    fused_o, _ = so.nViewPair_SurfaceNet_fn(X, params, w, N_vp=3)
    net = SurfaceNet.Net(params)
    hp = pipeline.HotPath(net, DeviceScene(cams, imgs), mode=mode)
    out = hp.infer_batch_host(pairs, xyz, resol, w, D)
    assert np.abs(out["fused"] - fused_o).max() <= PROB_TOL
    # votes: bit-exact against the oracle run on the DEVICE's own float16 prediction (a 1e-5 difference in the
    # probability may legitimately flip a float16 rounding; the vote logic itself must be exact)
    p16 = out["pred16"]
    assert p16.dtype == np.float16 and np.array_equal(p16, out["fused"][:, 0].astype(np.float16))
    for b in range(2):
        ref = raypool_oracle.rayPooling_1cube_numpy(cams, None, p16[b], pairs[b], xyz[b], resol[b], prediction_thresh=0.46)
        assert np.array_equal(out["votes"][b], ref.astype(np.uint8))
    assert out["votes"].max() >= 1
    assert hp.h2d_bytes > 0 and hp.d2h_bytes == 2 * D ** 3 * 7


def test_infer_batch_device_equals_host_entry(torch_cuda, params, cams):
    import torch
    from surfacenet_b200 import SurfaceNet, pipeline
    from surfacenet_b200.device import DeviceScene
    D = 16
    _, pairs, xyz, resol, imgs = _real_like_X(cams, D, n_cubes=3, n_vp=2, seed=6)
    w = (np.random.RandomState(4).rand(3, 2) + 0.1).astype(np.float32)
    hp = pipeline.HotPath(SurfaceNet.Net(params), DeviceScene(cams, imgs))
    h = hp.infer_batch_host(pairs, xyz, resol, w, D)
    d = hp.infer_batch(torch.from_numpy(pairs.astype(np.int32)).cuda(), torch.from_numpy(xyz).cuda(), torch.from_numpy(resol).cuda(),
                       torch.from_numpy(w).cuda(), D, want_unfused=True)
    assert np.array_equal(d["fused"].cpu().numpy(), h["fused"]) and np.array_equal(d["votes"].cpu().numpy(), h["votes"])
    assert d["unfused"].shape == (3, 2, D, D, D)
    empty = hp.infer_batch_host(pairs[:0], xyz[:0], resol[:0], w[:0], D)
    assert empty["fused"].shape == (0, 1, D, D, D)


@pytest.mark.parametrize("D,n_cubes,n_vp", [(64, 1, 2), (32, 3, 3), (16, 2, 1)])
def test_fused_gather_is_bit_identical_to_cvc_then_forward(torch_cuda, params, cams, D, n_cubes, n_vp):
    """infer_batch colours conv1_1's Winograd operand straight from the images (conv_wg.cu:cvc_wino_kernel); the documented
    two-step form -- CVC.gen_coloredCubes + preprocess_augmentation, then nViewPair_SurfaceNet_fn(X, w) -- must give the SAME bits,
    also for voxels outside the image (one cube is placed so that part of it projects out of the frame)."""
    import torch
    from surfacenet_b200 import CVC, SurfaceNet, pipeline
    from surfacenet_b200.device import DeviceScene
    _, pairs, xyz, resol, imgs = _real_like_X(cams, 4, n_cubes=n_cubes, n_vp=n_vp, seed=21 + D)
    xyz[-1] = np.array([-160.0, 95.0, 640.0], np.float32)                 # partly outside the 1600 x 1200 frame of most views
    w = (np.random.RandomState(D).rand(n_cubes, n_vp) + 0.1).astype(np.float32)
    net = SurfaceNet.Net(params)
    hp = pipeline.HotPath(net, DeviceScene(cams, imgs))
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    got = hp.infer_batch(dev(pairs.astype(np.int32)), dev(xyz), dev(resol), dev(w) if n_vp > 1 else None, D, want_unfused=True, ray_pool=False)
    X = CVC.gen_coloredCubes(pairs, xyz, resol, cams, imgs, D)
    _, X = CVC.preprocess_augmentation(None, X, mean_rgb=util.MEAN6[None, :, None, None, None], augment_ON=False, crop_ON=False)
    scope = (X != -util.MEAN6[None, :, None, None, None]).mean()
    assert 0.05 < scope < 0.999, scope                                     # both in-image and out-of-image voxels are present
    fused, unf = net.forward(dev(X), dev(w) if n_vp > 1 else None, n_vp)
    assert np.array_equal(got["fused"].cpu().numpy(), fused.cpu().numpy())
    if n_vp > 1:
        assert np.array_equal(got["unfused"].cpu().numpy(), unf.cpu().numpy())


def test_raypool_dense_selection_uses_full_tables(torch_cuda, cams):
    """Ray-pool hash tables are sized by the cube's selected-voxel count (rp_dyn_cap): every voxel selected drives them to the
    allocated worst case, a handful keeps the 256-slot minimum; both against the oracle, in one batched call."""
    from oracle import raypool_oracle
    from surfacenet_b200 import rayPooling
    D = 32
    rs = np.random.RandomState(12)
    pred = np.stack([(rs.rand(D, D, D) * 0.5 + 0.47).astype(np.float16),                     # all selected
                     np.where(rs.rand(D, D, D) > 0.9995, 0.9, 0.1).astype(np.float16)])      # ~16 voxels selected
    pairs = np.array([[[3, 4], [4, 9]], [[22, 23], [30, 3]]], np.int32)
    xyz = np.array([[5.0, -20.0, 640.0], [-14.0, 12.0, 655.0]], np.float32)
    resol = np.full(2, 0.4, np.float32)
    import torch
    votes = rayPooling.votes_device(torch.from_numpy(pred).cuda(), torch.from_numpy(pairs).cuda(), torch.from_numpy(xyz).cuda(),
                                    torch.from_numpy(resol).cuda(), torch.from_numpy(cams).cuda(), cams.shape[0], 0.46).cpu().numpy()
    for b in range(2):
        ref = raypool_oracle.rayPooling_1cube_numpy(cams, None, pred[b], pairs[b], xyz[b], resol[b], prediction_thresh=0.46)
        assert np.array_equal(votes[b], ref.astype(np.uint8))
    assert votes[0].max() >= 1 and (pred[1] > 0.46).sum() < 128


# ---- "next" rows N1 / N2: colour fusion and dense -> sparse ------------------------------------------------
def test_color_fusion_matches_reference_outputs(torch_cuda, golden):
    from surfacenet_b200 import utils as sn_utils
    out = sn_utils.generate_voxelLevelWeighted_coloredCubes(golden["cf_cc"].astype(np.float32), golden["cf_pred"], golden["cf_w"])
    assert out.dtype == np.uint8 and np.array_equal(out, golden["cf_out"])          # bit-exact


def test_color_fusion_s32_vs_oracle(torch_cuda):
    from oracle import sparse_oracle
    from surfacenet_b200 import utils as sn_utils
    rs = np.random.RandomState(5)
    cc = rs.randint(0, 256, size=(3 * 5, 6, 32, 32, 32)).astype(np.float32)
    cc = (cc - util.MEAN6[None, :, None, None, None]) + util.MEAN6[None, :, None, None, None]     # main_reconstruct.py:143,150 round trip
    p = rs.rand(3, 5, 32, 32, 32).astype(np.float32) * 0.98 + 0.01
    w = (rs.rand(3, 5) + 0.1).astype(np.float32)
    assert np.array_equal(sn_utils.generate_voxelLevelWeighted_coloredCubes(cc, p, w),
                          sparse_oracle.generate_voxelLevelWeighted_coloredCubes(cc, p, w))
    with pytest.raises(ValueError):
        sn_utils.generate_voxelLevelWeighted_coloredCubes(cc, p, w[:, :3])


@pytest.mark.parametrize("name", ["d16", "d32", "all_empty"])
def test_dense2sparse_matches_reference_outputs(torch_cuda, golden, cams, name):
    from surfacenet_b200 import sparseCubes
    from tests.test_oracle_golden import _check_sparse
    case = util.sparse_cases(cams)[name]
    res = sparseCubes.append_dense_2sparseList(case["pred"], case["rgb"], case["param"], case["pairs"], min_prob=case["min_prob"],
                                               rayPool_thresh=0, enable_centerCrop=True, cube_Dcenter=case["Dcenter"],
                                               enable_rayPooling=True, cameraPOs=cams, cameraTs=None,
                                               prediction_list=[], rgb_list=[], vxl_ijk_list=[], rayPooling_votes_list=[])
    _check_sparse(res, golden, name)


def test_dense2sparse_s64_raypool_threshold_vs_oracle(torch_cuda, cams):
    """Full cube size (64^3, centre 52^3: params.py:107) and the votes >= rayPool_thresh branch (sparseCubes.py:64)."""
    from oracle import sparse_oracle
    from surfacenet_b200 import sparseCubes
    rs = np.random.RandomState(8)
    n, D, Dc = 2, 64, 52
    pred = np.stack([util.sheet_prediction(D, 0.4 * i, 0.07) for i in range(n)])
    rgb = rs.randint(0, 256, size=(n, D, D, D, 3)).astype(np.uint8)
    param = np.zeros(n, util.PARAM_DTYPE)
    param["xyz"] = np.array([[20.0, -12.5, 630.0], [-10.3, 30.7, 655.1]], np.float32); param["resol"] = np.float32(0.4)
    pairs = np.array([[[0, 5], [17, 22], [5, 30]], [[1, 2], [3, 4], [5, 6]]], np.uint16)
    for thresh in (0, 2):
        a = sparseCubes.dense2sparse(pred, rgb, param, pairs, 0.46, thresh, True, Dc, True, cams, None)
        b = sparse_oracle.dense2sparse(pred, rgb, param, pairs, 0.46, thresh, True, Dc, True, cams, None)
        assert a[0] == b[0] and len(a[0]) == n
        for k in (1, 2, 3, 4):
            assert all(np.array_equal(x, y) and x.dtype == y.dtype for x, y in zip(a[k], b[k])), "list %d differs (thresh %d)" % (k, thresh)
        assert np.array_equal(a[5]["xyz"], b[5]["xyz"])


def test_infer_batch_sparse_matches_dense_pipeline(torch_cuda, params, cams):
    """The fully fused call (CVC -> net -> fusion -> colours -> votes -> compaction) against the reference sequence of
    main_reconstruct.py:143-162 applied to the device's own dense outputs."""
    import torch
    from oracle import cvc_oracle, sparse_oracle
    from surfacenet_b200 import SurfaceNet, pipeline
    from surfacenet_b200.device import DeviceScene
    D, Dc = 32, 26
    X, pairs, xyz, resol, imgs = _real_like_X(cams, D, n_cubes=2, n_vp=3, seed=9)
    w = (np.random.RandomState(6).rand(2, 3) + 0.1).astype(np.float32)
    hp = pipeline.HotPath(SurfaceNet.Net(params), DeviceScene(cams, imgs), mode="exact")
    sp = hp.infer_batch_sparse(pairs, xyz, resol, w, D, Dc)
    dense = hp.infer_batch(torch.from_numpy(pairs.astype(np.int32)).cuda(), torch.from_numpy(xyz).cuda(), torch.from_numpy(resol).cuda(),
                           torch.from_numpy(w).cuda(), D, want_unfused=True)
    Xc = X + util.MEAN6[None, :, None, None, None]                                              # main_reconstruct.py:150
    rgb = sparse_oracle.generate_voxelLevelWeighted_coloredCubes(Xc, dense["unfused"].cpu().numpy(), w)
    param = np.zeros(2, util.PARAM_DTYPE); param["xyz"] = xyz; param["resol"] = resol
    ref = sparse_oracle.append_dense_2sparseList(dense["fused"].cpu().numpy(), rgb, param, pairs, min_prob=0.46, rayPool_thresh=0,
                                                 enable_centerCrop=True, cube_Dcenter=Dc, enable_rayPooling=True, cameraPOs=cams, cameraTs=None)
    pl, rl, il, vl = ref[:4]
    assert sp["counts"].sum() == sum(len(x) for x in pl) > 0
    assert np.array_equal(sp["pred"], np.concatenate(pl)) and np.array_equal(sp["ijk"], np.concatenate(il))
    assert np.array_equal(sp["rgb"], np.concatenate(rl)) and np.array_equal(sp["votes"], np.concatenate(vl))


def test_reconstruct_cubes_batching_and_npz(torch_cuda, params, cams, tmp_path):
    """The inference section of main_reconstruct.reconstruction over a small cube grid: batched == one call, ranks
    partition the cubes, NPZ round trip in the reference schema."""
    from surfacenet_b200 import SurfaceNet, pipeline, reconstruct, sparseCubes
    from surfacenet_b200.device import DeviceScene
    D, Dc = 16, 12
    cubes, side = reconstruct.initialize_cubes(np.float32(0.4), D, Dc, 0.5, np.array([[10, 14], [-30, -26], [620, 624]]))
    assert side == np.float32(0.4) * D and len(cubes) > 20 and np.array_equal(cubes["ijk"][1], [0, 0, 1])
    cubes = cubes[::4][:7]
    rs = np.random.RandomState(2)
    used = [8, 9, 22, 23]
    imgs = util.image_list(49, used)
    pairs = rs.choice(used, size=(len(cubes), 2, 2))
    w = (rs.rand(len(cubes), 2) + 0.1).astype(np.float32)
    hot = pipeline.HotPath(SurfaceNet.Net(params), DeviceScene(cams, imgs), mode="exact")
    res = reconstruct.reconstruct_cubes(hot, cubes, pairs, w, D, Dc, batch_size=3)
    one = hot.infer_batch_sparse(pairs, cubes["xyz"], cubes["resol"], w, D, Dc)
    ref = sparseCubes.lists_from_flat(one, cubes, pairs, Dc, D)
    assert len(res[0]) == len(ref[0]) > 0
    for a, b in zip(res, ref):
        if isinstance(a, list):
            assert all(np.array_equal(x, y) for x, y in zip(a, b))
        else:
            assert np.array_equal(a, b)
    # two "ranks" see disjoint cubes whose union is everything
    parts = [reconstruct.reconstruct_cubes(hot, cubes, pairs, w, D, Dc, batch_size=2, rank=r, world_size=2) for r in range(2)]
    got = sorted(tuple(x) for p in parts if p != "Empty!" for x in p[4].tolist())
    assert got == sorted(tuple(x) for x in ref[4].tolist())
    f = str(tmp_path / "model.npz")
    ply = str(tmp_path / "fixThresh.ply")
    masks, denoised = reconstruct.finish(res, f, ply_path=ply, tau=0.5, gamma=0.0, cube_D=D)
    back = sparseCubes.load_sparseCubes(f)
    assert all(np.array_equal(x, y) for x, y in zip(back[0], res[0])) and np.array_equal(back[5]["xyz"], res[5]["xyz"])
    assert len(masks) == len(res[0]) and masks[0].dtype == bool
    from oracle import postprocess_oracle as post                              # main_reconstruct.py:170-176 on the CPU
    m_o = post.filter_voxels([], res[0], 0.5, res[3], 0.0)
    d_o = post.denoise_crossCubes(res[4], res[2], m_o, D)
    assert all(np.array_equal(a, b) for a, b in zip(masks, m_o)) and all(np.array_equal(a, b) for a, b in zip(denoised, d_o))
    xyz_ply, _ = util.read_ply(ply)
    assert xyz_ply.shape[0] == int(sum(d.sum() for d in denoised))
    assert reconstruct.reconstruct_cubes(hot, cubes[:0], pairs[:0], w[:0], D, Dc) == "Empty!"


def test_forward_with_fused_side_epilogue_subprocess(torch_cuda):
    """SN_TC_FUSE_SIDE=1 (side_op1 / side_op2 evaluated in the epilogue of conv1_3 / conv2_3; read once per process, hence the
    subprocess): same <= 1e-4 bound against the oracle."""
    import os, subprocess, sys
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import numpy as np\n"
        "from tests import util\n"
        "from tests.test_gpu_parity import _real_like_X\n"
        "from oracle import surfacenet_oracle as so\n"
        "from surfacenet_b200 import SurfaceNet, weights\n"
        "params = weights.synthetic_params(0)\n"
        "X, *_ = _real_like_X(util.dtu_cameras(), 32, n_cubes=1, n_vp=2, seed=3)\n"
        "w = np.array([[0.4, 0.6]], np.float32)\n"
        "fo, uo = so.nViewPair_SurfaceNet_fn(X, params, w, N_vp=2)\n"
        "_, fn = SurfaceNet.SurfaceNet_inference(2, params, mode='exact')\n"
        "f, u = fn(X, w)\n"
        "print('ERR', float(np.abs(f - fo).max()), float(np.abs(u - uo).max()))\n" % util.REPO)
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, SN_TC_FUSE_SIDE="1"), stdout=subprocess.PIPE,
                         stderr=subprocess.PIPE, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    e = [float(x) for x in out.stdout.strip().splitlines()[-1].split()[1:]]
    assert max(e) <= PROB_TOL, e


def test_fused_passes_are_bit_identical_to_the_separate_ones(torch_cuda, params, cams, tmp_path):
    """The default chain (images -> conv1_1's operand in one pass; side_op1 + pool1 in one pass over conv1_3's output) against the separate
    passes it replaces (SN_CVC_FUSED=0 SN_SIDE_POOL=0: cvc_gather + pack_wino, side_wino + pool_blk), the latter also with the opt-in form of
    the Winograd units as CTA pairs sharing one multicast weight stream (SN_WG_CLUSTER=2), in a child process
    because the switches are read once: probabilities, float16 predictions and votes must be the SAME bits, 64^3 and 32^3."""
    import os, subprocess, sys
    body = ("import sys; sys.path.insert(0, %r)\n"
            "import numpy as np, torch\n"
            "from surfacenet_b200 import SurfaceNet, pipeline, weights\n"
            "from surfacenet_b200.device import DeviceScene\n"
            "from tests import util\n"
            "from tests.test_gpu_parity import _real_like_X\n"
            "cams = util.dtu_cameras(); hp = None\n"
            "for D, nc, nv in ((64, 2, 2), (32, 3, 3)):\n"
            "    _, pairs, xyz, resol, imgs = _real_like_X(cams, 4, n_cubes=nc, n_vp=nv, seed=40 + D)\n"
            "    w = (np.random.RandomState(D).rand(nc, nv) + 0.1).astype(np.float32)\n"
            "    hp = hp or pipeline.HotPath(SurfaceNet.Net(weights.synthetic_params(0)), DeviceScene(cams, imgs))\n"
            "    o = hp.infer_batch_host(pairs, xyz, resol, w, D)\n"
            "    np.savez(sys.argv[1] + '_%%d.npz' %% D, fused=o['fused'], pred16=o['pred16'], votes=o['votes'])\n" % util.REPO)
    for tag, env in (("fused", {}), ("separate", {"SN_CVC_FUSED": "0", "SN_SIDE_POOL": "0", "SN_WG_CLUSTER": "2"})):
        out = subprocess.run([sys.executable, "-c", body, str(tmp_path / tag)], env=dict(os.environ, **env), stdout=subprocess.PIPE,
                             stderr=subprocess.PIPE, text=True, timeout=900)
        assert out.returncode == 0, out.stderr[-2000:]
    for D in (64, 32):
        a, b = np.load(str(tmp_path / ("fused_%d.npz" % D))), np.load(str(tmp_path / ("separate_%d.npz" % D)))
        for k in ("fused", "pred16", "votes"):
            assert np.array_equal(a[k], b[k]), (D, k)
        assert a["votes"].max() >= 1 and 0.02 < (a["fused"] > 0.46).mean() < 0.98
