"""GPU parity tests of the "next" row N3: view-pair angles, feature assembly + relative-importance + top-N selection, early rejection
(patch cropping, similarityNet embedding, pair dissimilarity, cube selection) against the reference's doctest known answers, golden
vectors produced by the reference's numpy code, and the torch-CPU oracle of the similarityNet.
Bars: bit-exact for index / uint8 work and the crop+preprocess arithmetic; floating point tolerances written at each assert."""
import math
import os
import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(util.REPO, "tests", "golden", "select_golden.npz"))


@pytest.fixture(scope="module")
def case(cams):
    return util.select_case(cams)


def test_viewPairAngles_reference_doctest_and_golden(g, case):
    from surfacenet_b200 import camera
    pts = np.array([[0, 0, 0], [1, 1, 1]], dtype=np.float32)           # utils/camera.py:290-294
    cT = np.array([[0, 0, 1], [0, 1, 1], [1, 0, 1]], dtype=np.float32)
    a = camera.viewPairAngles_wrt_pts(cT, pts)
    assert a.dtype == np.float32 and np.allclose(a * 180 / math.pi, [[45., 45., 60.], [45., 45., 90.]], atol=1e-4)
    assert np.abs(a - g["ang_doc"]).max() <= 2e-7                      # float32 acos: <= 2 ulp of pi/2
    cTs = g["cameraTs"][case["views"]]
    a64 = camera.viewPairAngles_wrt_pts(cTs, case["centers"].astype(np.float32))
    assert a64.dtype == np.float64 and np.abs(a64 - g["ang_dtu64"]).max() <= 1e-14
    a32 = camera.viewPairAngles_wrt_pts(cTs.astype(np.float32), case["centers"].astype(np.float32))
    assert a32.dtype == np.float32 and np.abs(a32 - g["ang_dtu32"]).max() <= 2e-7
    with pytest.raises(ValueError):
        camera.viewPairAngles_wrt_pts(np.zeros((3, 2)), pts)


def test_argmaxN_viewPairs_reference_doctest_and_golden(g, case):
    from surfacenet_b200 import viewPairSelection as vps
    from surfacenet_b200.utils import k_combination_np
    vp = k_combination_np(range(3), k=2)                               # utils/viewPairSelection.py:17-31
    w = np.array([[3, 1, 2], [0, -1, 70]])
    a, b = vps.__argmaxN_viewPairs__(vp, w, 1)
    assert a.tolist() == [[[0, 1]], [[1, 2]]] and b.tolist() == [[3], [70]]
    a, b = vps.__argmaxN_viewPairs__(vp, w, 2)
    assert a.tolist() == [[[1, 2], [0, 1]], [[0, 1], [1, 2]]] and b.tolist() == [[2, 3], [0, 70]]
    a, b = vps.__argmaxN_viewPairs__(case["viewPairs"], case["w_rand"], 5)
    assert np.array_equal(a, g["argmax_rand_pairs"]) and np.array_equal(b, g["argmax_rand_w"])
    # 1176 pairs (49 views), N = 5: against numpy's argsort on a large random matrix
    rs = np.random.RandomState(5)
    big = rs.rand(37, 1176).astype(np.float32)
    vp49 = k_combination_np(range(49), k=2)
    a, b = vps.__argmaxN_viewPairs__(vp49, big, 5)
    idx = np.argsort(big, axis=1, kind="stable")[:, -5:]
    assert np.array_equal(a, vp49[idx]) and np.array_equal(b, np.take_along_axis(big, idx, axis=1))
    a, b = vps.__argmaxN_viewPairs__(vp, w, 4)                           # N > n: argsort()[:, -4:] returns all 3 columns
    assert a.shape == (2, 3, 2) and b.tolist() == [[1, 2, 3], [-1, 0, 70]]
    # NaN rows with n not a power of two: numpy sorts NaN last, so NaN entries are among the "largest" (index order)
    wn = np.array([[0.5, np.nan, 0.1, 0.9, np.nan, 0.3, 0.2], [np.nan] * 7, [1, 2, 3, 4, 5, 6, 7.]])
    vp7 = np.stack([np.arange(7), np.arange(7) + 1], axis=1)
    a, b = vps.__argmaxN_viewPairs__(vp7, wn, 3)
    idx = np.argsort(wn, axis=1, kind="stable")[:, -3:]
    assert np.array_equal(a, vp7[idx]) and np.array_equal(np.isnan(b), np.isnan(np.take_along_axis(wn, idx, axis=1)))
    assert idx.max() < 7 and a.max() <= 7


def test_crop_and_preprocess_bit_exact(g, case):
    from surfacenet_b200 import image
    img = util.synth_image(3, 300, 400)
    p = image.cropImgPatches(img, case["range_h"], case["range_w"], patchSize=64, pyramidRate=1, interp_order=2,
                             cubeCenter_hw=(case["crop_ch"], case["crop_cw"]))
    assert p.dtype == np.uint8 and np.array_equal(p, g["crop_patches"])
    assert np.array_equal(image.preprocess_patches(np.zeros((2, 2, 5, 3)), np.array([1, 2, 3])), g["pre_doc"])
    with pytest.raises(ValueError):
        image.cropImgPatches(img, case["range_h"], case["range_w"], pyramidRate=1.2)


def test_early_rejection_matches_reference_outputs(g, case):
    """The reference's earlyRejection functions were run with deterministic stand-in networks; the same stand-ins (numpy callables)
    are plugged into the GPU drop-ins, so cropping / preprocessing / scatter / pair enumeration / selection are compared exactly."""
    import torch
    from surfacenet_b200 import earlyRejection as er
    def emb_fn(p):                                                       # accepts numpy or cuda tensors like the real callable
        if torch.is_tensor(p):
            return torch.from_numpy(util.fake_patch2embedding_fn(p.cpu().numpy())).cuda()
        return util.fake_patch2embedding_fn(p)
    def pair_fn(p):
        return torch.from_numpy(util.fake_pair2simil_fn(p.cpu().numpy())).cuda()
    emb, inscope = er.patch2embedding(case["images"], case["h_corner"], case["w_corner"], emb_fn, util.MEAN_BGR, case["N_cubes"],
                                      len(case["views"]), 16, patchSize=64, batchSize=5, cubeCenter_hw=case["center_hw"])
    assert np.array_equal(inscope, g["er_inscope"]) and np.array_equal(emb, g["er_emb"])
    dis = er.embeddingPairs2simil(emb, len(case["views"]), inscope, pair_fn, 7, case["viewPairs"])
    assert np.array_equal(dis, g["er_dissim"])
    assert np.array_equal(er.selectFromSimilarity(dis, 3), g["er_select"])
    assert np.array_equal(er.selectFromSimilarity(g["er_dissim"], 0), np.ones(case["N_cubes"], bool))


def test_similarityNet_vs_oracle(case):
    from oracle import selection_oracle as so
    from surfacenet_b200 import similarityNet
    params = similarityNet.synthetic_params(0)
    p2e, pair = similarityNet.similarityNet_inference(params, (64, 64))
    rs = np.random.RandomState(1)
    img = util.synth_image(7, 400, 500)
    patches = so.preprocess_patches(so.cropImgPatches_rate1(img, 64, (rs.rand(9) * 400, rs.rand(9) * 500)).astype(np.float32), util.MEAN_BGR)
    patches = np.concatenate([patches, so.preprocess_patches(np.zeros((1, 64, 64, 3), np.float32), util.MEAN_BGR)])
    e = p2e(np.ascontiguousarray(patches, np.float32))
    e_o = so.patch2embedding_fn(patches, params)
    assert e.shape == (10, 128) and e.dtype == np.float32
    scale = np.abs(e_o).max()
    assert np.abs(e - e_o).max() <= 1e-4 * scale, (np.abs(e - e_o).max(), scale)      # fp32 FMA vs torch-CPU fp32, 13 conv layers deep
    pairs = np.ascontiguousarray(e_o[[0, 1, 2, 3, 4, 4, 9, 0]], np.float32)
    s = pair(pairs)
    s_o = so.embeddingPair2simil_fn(pairs, params)
    assert s.shape == (4, 1) and np.abs(s - s_o).max() <= 1e-6
    assert abs(float(s[2, 0]) - 1.0 / (1.0 + math.exp(-2.0))) <= 1e-6               # identical embeddings: distance 0 -> sigmoid(b)
    with pytest.raises(ValueError):
        p2e(np.zeros((2, 3, 32, 32), np.float32))
    with pytest.raises(ValueError):
        similarityNet.similarityNet_inference(params[:-1], (64, 64))


def test_viewPairSelection_vs_reference_output(g, case):
    """Reference viewPairSelection was run with the oracle's relative-importance MLP; here the GPU MLP (sn_net_relative_importance)
    produces the weights: same selected pairs, weights within 1e-5."""
    from surfacenet_b200 import SurfaceNet, viewPairSelection as vps, weights
    fn, _ = SurfaceNet.SurfaceNet_inference(4, weights.synthetic_params(0))
    selp, w = vps.viewPairSelection(g["cameraTs"][case["views"]], g["vps_e"], g["vps_d"], g["vps_valid"], case["centers"].astype(np.float32), fn,
                                    4 * case["viewPairs"].shape[0] + 3, 4, case["viewPairs"])
    assert w.shape == g["vps_w"].shape and np.abs(w - g["vps_w"]).max() <= 1e-5
    gaps = np.diff(np.sort(g["vps_w"], axis=1), axis=1).min()
    if gaps > 1e-5:                                                    # well separated weights: the discrete choice must agree
        assert np.array_equal(selp, g["vps_sel"])


def test_end_to_end_early_rejection_and_selection_vs_oracle(case, g):
    """similarityNet embeddings -> dissimilarity -> valid cubes -> view pairs + weights, GPU drop-ins vs the oracle chain."""
    from oracle import selection_oracle as so, surfacenet_oracle
    from surfacenet_b200 import SurfaceNet, earlyRejection as er, similarityNet, viewPairSelection as vps, weights
    sp = similarityNet.synthetic_params(0)
    p2e, pair = similarityNet.similarityNet_inference(sp, (64, 64))
    imgs, V = case["images"][:3], 3
    args = (imgs, case["h_corner"][:V], case["w_corner"][:V])
    emb, inscope = er.patch2embedding(*args, p2e, util.MEAN_BGR, case["N_cubes"], V, 128, patchSize=64, batchSize=16, cubeCenter_hw=case["center_hw"][:, :V])
    emb_o, inscope_o = so.patch2embedding(*args, lambda p: so.patch2embedding_fn(p, sp), util.MEAN_BGR, case["N_cubes"], V, 128, 64, 16, case["center_hw"][:, :V])
    assert np.array_equal(inscope, inscope_o)
    assert np.abs(emb - emb_o).max() <= 1e-4 * np.abs(emb_o).max()
    vp = so.k_combination_np(range(V), 2)
    dis = er.embeddingPairs2simil(emb, V, inscope, pair, 1000, vp)
    dis_o = so.embeddingPairs2simil(emb_o, V, lambda p: so.embeddingPair2simil_fn(p, sp), 1000)
    assert dis.shape == (case["N_cubes"], 3) and np.abs(dis - dis_o).max() <= 1e-4
    valid = np.ones(case["N_cubes"], bool)
    params = weights.synthetic_params(0)
    fn, _ = SurfaceNet.SurfaceNet_inference(2, params)
    cT = g["cameraTs"][case["views"][:V]]
    selp, w = vps.viewPairSelection(cT, emb, dis, valid, case["centers"].astype(np.float32), fn, 100, 2, vp)
    fo = lambda f, n_samples_perGroup: surfacenet_oracle.viewPair_relativeImpt_fn(f, params, n_samples_perGroup)
    selo, wo = so.viewPairSelection(cT, emb_o, dis_o, valid, case["centers"].astype(np.float32), fo, 100, 2, vp)
    assert np.abs(w - wo).max() <= 1e-3 and selp.shape == selo.shape == (case["N_cubes"], 2, 2)


def test_whole_scene_reconstruction_runs_end_to_end(tmp_path, cams):
    """main_reconstruct.reconstruction from arrays: cube grid -> early rejection -> view-pair selection -> SurfaceNet inference ->
    thresholding + denoising -> PLY / NPZ -> adapthresh, every stage on the GPU drop-ins (each stage has its own parity test;
    here the chain is checked for consistency)."""
    from surfacenet_b200 import adapthresh, reconstruct, similarityNet, sparseCubes, weights
    views = [3, 8, 15, 22, 30]
    images = [util.synth_image(200 + v, 1200, 1600) for v in views]
    P = cams[views]
    BB = np.array([[-10.0, 15.0], [-20.0, 5.0], [620.0, 645.0]])
    # permissive similarity head (sigmoid(-0.02 d - 0.4) in (0.1, 0.5) for most pairs) so that cubes survive early rejection
    sp = similarityNet.synthetic_params(0)
    sp[28] = np.array([[-0.02]], np.float32); sp[29] = np.array([-0.4], np.float32)
    out = reconstruct.reconstruction(images, P, BB, np.float32(0.4), 2, weights.synthetic_params(0), sp, outputFolder=str(tmp_path),
                                     cube_D=32, mode="exact", batch_size=8, tau=0.5, gamma=0.0, model="synth")
    assert out != "Empty!"
    n_valid = int(out["validCubes"].sum())
    assert n_valid > 0 and out["viewPairs4Reconstr"].shape == (n_valid, 2, 2) and out["w_viewPairs4Reconstr"].shape == (n_valid, 2)
    assert np.all(out["viewPairs4Reconstr"] < len(views)) and np.allclose(out["w_viewPairs4Reconstr"].sum(axis=1) <= 1.0 + 1e-5, True)
    pl, rl, il, vl, cube_ijk, param, vp = out["result"]
    assert len(pl) == len(il) == cube_ijk.shape[0] > 0 and il[0].max() < 26
    back = sparseCubes.load_sparseCubes(out["npz_path"])
    assert all(np.array_equal(a, b) for a, b in zip(back[0], pl))
    xyz, _ = util.read_ply(out["ply_path"])
    assert xyz.shape[0] == int(sum(d.sum() for d in out["vxl_maskDenoised_list"]))
    last = adapthresh.adapthresh(save_result_fld=str(tmp_path), N_refine_iter=2, D_cube=26, init_probThresh=0.5, min_probThresh=0.5,
                                 max_probThresh=0.9, rayPool_thresh=0, beta=6, gamma=0.8, npz_file=out["npz_path"], RGB_visual_ply=False)
    assert os.path.exists(last) and last.endswith("iter1.ply")


def test_perspectiveProj_cubesCorner_matches_oracle(cams):
    """utils/camera.py:186-250: (N_Ms, N_pts, 8) projections of the cube corners == the oracle's perspectiveProj of the same corners."""
    from oracle import camera_oracle
    from surfacenet_b200 import camera
    xyz = np.array([[10.0, -30.0, 620.0], [30.0, 0.0, 650.0], [-20.5, 12.25, 600.0]], np.float32)
    D_mm = np.float32(0.4) * 64
    h, w = camera.perspectiveProj_cubesCorner(cams[[3, 8]], xyz, D_mm, return_int_hw=False)
    assert h.shape == w.shape == (2, 3, 8)
    corners = (xyz[:, None, :] + np.indices((2, 2, 2)).reshape((3, -1)).T[None] * D_mm).reshape(-1, 3)
    ho, wo = camera_oracle.perspectiveProj(cams[[3, 8]], corners, return_int_hw=False)
    assert np.array_equal(h.reshape(2, -1), ho) and np.array_equal(w.reshape(2, -1), wo)
    hi, wi = camera.perspectiveProj_cubesCorner(cams[3], xyz[0], D_mm, return_int_hw=True)
    assert hi.shape == (1, 1, 8) and hi.dtype == np.int64
    with pytest.raises(ValueError):
        camera.perspectiveProj_cubesCorner(np.zeros((4, 4)), xyz, D_mm)



def test_main_reconstruct_dropin_from_files(tmp_path, cams):
    """surfacenet_b200.main_reconstruct.reconstruction with the reference's argument list: images / cameras read from files, NPZ path back."""
    from PIL import Image
    from surfacenet_b200 import main_reconstruct, similarityNet, sparseCubes, weights
    views = [3, 8, 15, 22]
    os.makedirs(tmp_path / "cal")
    for v in views:
        np.savetxt(str(tmp_path / "cal" / ("pos_%03d.txt" % v)), cams[v], delimiter=' ')
        Image.fromarray(util.synth_image(200 + v, 1200, 1600)).save(str(tmp_path / ("rect_%03d_3.png" % v)))
    sp = similarityNet.synthetic_params(0)
    sp[28] = np.array([[-0.02]], np.float32); sp[29] = np.array([-0.4], np.float32)
    npz = main_reconstruct.reconstruction(str(tmp_path), 9, "rect_#_3.png", "cal/pos_#.txt", None, str(tmp_path / "out"), 2, np.float32(0.4),
                                          np.array([[-10.0, 15.0], [-20.0, 5.0], [620.0, 645.0]]), views, surfacenet_model=weights.synthetic_params(0),
                                          similnet_model=sp, cube_D=32, tau=0.5, gamma=0.0)
    assert npz.endswith("model9-4views.npz") and os.path.exists(npz)
    back = sparseCubes.load_sparseCubes(npz)
    assert len(back[0]) > 0 and back[6].shape[1:] == (2, 2) and os.path.exists(str(tmp_path / "out" / "fixThresh_tau0.5_gamma0.0.ply"))
    grid_xyz, _ = util.read_ply(str(tmp_path / "out" / "initialCubes.ply"))                     # main_reconstruct.py:61
    # the same scene from an initial point cloud (main_reconstruct.py:57-60): the cubes cover the points only
    pts = grid_xyz[::7][:3].astype(np.float32)
    sparseCubes.save2ply(str(tmp_path / "init_pts.ply"), pts)
    npz2 = main_reconstruct.reconstruction(str(tmp_path), 9, "rect_#_3.png", "cal/pos_#.txt", "init_pts.ply", str(tmp_path / "out2"), 2, np.float32(0.4),
                                           np.array([[-10.0, 15.0], [-20.0, 5.0], [620.0, 645.0]]), views, surfacenet_model=weights.synthetic_params(0),
                                           similnet_model=sp, cube_D=32, tau=0.5, gamma=0.0)
    cubes2, _ = util.read_ply(str(tmp_path / "out2" / "initialCubes.ply"))
    assert 2 <= cubes2.shape[0] <= 6 and cubes2.shape[0] < grid_xyz.shape[0]                   # cells floor and floor + 1 per point (scene.py:92-98), shared cells once
    assert npz2 == "Empty!" or len(sparseCubes.load_sparseCubes(npz2)[0]) <= cubes2.shape[0]
