"""GPU parity on REAL inputs (real DTU scan9 JPEGs; BASELINE configs[4]: Middlebury dinoSparseRing views 7-12).

The images are the reference's own input files, committed as fixtures under tests/golden/real/ (six rectified scan9 views, the six
dinoSparseRing views params.py:174-182 selects, the Middlebury calibration file); cameras of scan9 come from the cal18 fixture.
Network weights are the calibrated synthetic sets (the pre-trained .model files are not distributed with the reference)."""
import os
import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu
REAL = os.path.join(util.REPO, "tests", "golden", "real")
SCAN9_VIEWS = [8, 9, 22, 23, 30, 33]                       # 0-based positions of rect_009/010/023/024/031/034
PROB_TOL = 1e-4


def _scan9_images():
    from PIL import Image
    imgs = [None] * 49
    for v in SCAN9_VIEWS:
        imgs[v] = np.array(Image.open(os.path.join(REAL, "scan9", "rect_{:03}_3_r5000.jpg".format(v + 1))).convert("RGB"))
    return imgs


@pytest.mark.parametrize("D", [32, 64])
def test_scan9_real_images_cvc_bit_exact_and_network_parity(cams, D):
    """Cubes placed on the scan9 surface (inside the reference's debug range x -40..40, y 80..160, z 630..680, params.py:171), real JPEG
    pixels: the CVC gather is bit-exact against the oracle (utils/CVC.py:6-104), the mean-subtracted CVCs go through the network within
    1e-4 of the torch-CPU fp32 oracle (the BatchNorm statistics of the synthetic weights were calibrated on exactly this kind of input)."""
    from oracle import cvc_oracle, surfacenet_oracle as so
    from surfacenet_b200 import CVC, SurfaceNet, weights
    imgs = _scan9_images()
    n_cubes, n_vp = (3, 2) if D == 32 else (1, 3)
    rs = np.random.RandomState(D)
    pairs = np.stack([np.stack([rs.choice(SCAN9_VIEWS, 2, replace=False) for _ in range(n_vp)]) for _ in range(n_cubes)])
    xyz = (np.array([-20.0, 95.0, 635.0]) + rs.rand(n_cubes, 3) * np.array([30.0, 40.0, 25.0])).astype(np.float32)
    resol = np.full(n_cubes, 0.4, np.float32)
    X_o = cvc_oracle.gen_coloredCubes(pairs, xyz, resol, cams, imgs, D)
    X = CVC.gen_coloredCubes(pairs, xyz, resol, cams, imgs, D)
    assert np.array_equal(X, X_o)                                           # exact colours of real pixels
    assert (X_o != 0).mean() > 0.5 and X_o.std() > 20                       # the cubes do see the object (black background pixels are 0)
    mean = util.MEAN6[None, :, None, None, None]
    _, Xm_o = cvc_oracle.preprocess_augmentation(None, X_o, mean, False, False)
    _, Xm = CVC.preprocess_augmentation(None, X, mean, False, False)
    assert np.array_equal(Xm, Xm_o)
    w = (rs.rand(n_cubes, n_vp) + 0.1).astype(np.float32)
    worst = 0.0
    for seed in (0, 2):
        params = weights.synthetic_params(seed)
        fused_o, unf_o = so.nViewPair_SurfaceNet_fn(Xm_o, params, w, N_vp=n_vp, chunk=1 if D == 64 else 4)
        _, fn = SurfaceNet.SurfaceNet_inference(n_vp, params)
        fused, unf = fn(Xm, w)
        e_f, e_u = float(np.abs(fused - fused_o).max()), float(np.abs(unf - unf_o).max())
        print("scan9 real D=%d seed %d: max-abs fused %.3g unfused %.3g, prob mean %.3f" % (D, seed, e_f, e_u, fused_o.mean()))
        worst = max(worst, e_f, e_u)
    assert worst <= PROB_TOL


def _dino_inputs():
    from surfacenet_b200 import camera, image
    views = list(range(7, 13))                                                           # params.py:182 viewList
    imgs = image.readImages(os.path.join(REAL), "dinoSparseRing/dinoSR0#.png", views)   # imgNamePattern, params.py:179
    P = camera.readCameraPOs_as_np(REAL, "Middlebury", "dinoSparseRing/dinoSR_par.txt", "dinoSparseRing", views)
    BB = np.array([(-0.061897, 0.010897), (-0.018874, 0.068227), (-0.057845, 0.015495)], dtype=np.float32)   # params.py:181
    return imgs, P, BB


def test_dinoSparseRing_scene_reconstruction_vs_oracle_chain(tmp_path):
    """BASELINE configs[4] on one GPU: Middlebury dinoSparseRing, views 7-12, s=64 cubes, N_viewPairs4inference = 3, resol 0.00025
    (params.py:174-182) through reconstruct.reconstruction (cube grid -> early rejection -> view-pair selection -> SurfaceNet -> sparse
    lists -> fixed threshold + cross-cube denoising -> PLY / NPZ) and adapthresh.  Three of the surviving cubes are then recomputed by the
    oracle chain (numpy CVC -> torch-CPU network + fusion with the SAME selected pairs / weights -> float16 -> numpy ray pooling -> dense2sparse)
    and compared voxel by voxel."""
    from oracle import cvc_oracle, raypool_oracle, surfacenet_oracle as so
    from surfacenet_b200 import adapthresh, reconstruct, similarityNet, sparseCubes, weights
    imgs, P, BB = _dino_inputs()
    assert len(imgs) == 6 and imgs[0].shape == (480, 640, 3) and P.shape == (6, 3, 4)
    params = weights.synthetic_params(1)
    sp = similarityNet.synthetic_params(0)
    sp[28] = np.array([[-0.02]], np.float32); sp[29] = np.array([-0.4], np.float32)     # permissive similarity head: cubes survive early rejection
    D, Dc, N = 64, 52, 3
    out = reconstruct.reconstruction(imgs, P, BB, np.float32(0.00025), N, params, sp, outputFolder=str(tmp_path), cube_D=D, batch_size=16,
                                     tau=0.7, gamma=0.8, model="dinoSparseRing")
    assert not isinstance(out, str)
    n_valid = int(out["validCubes"].sum())
    pl, rl, il, vl, cube_ijk, param, vp = out["result"]
    print("dinoSparseRing: %d cubes in the grid, %d after early rejection, %d non-empty, %d sparse voxels" %
          (out["validCubes"].size, n_valid, len(pl), sum(len(x) for x in pl)))
    assert n_valid >= 20 and len(pl) >= 3 and vp.shape[1] == N
    back = sparseCubes.load_sparseCubes(out["npz_path"])
    assert all(np.array_equal(a, b) for a, b in zip(back[0], pl)) and os.path.exists(out["ply_path"])
    # ---- oracle chain on three cubes (first, middle, last of the non-empty ones) ----
    mean = util.MEAN6[None, :, None, None, None]
    margin = (D - Dc) // 2
    sel_pairs, sel_w = out["viewPairs4Reconstr"], out["w_viewPairs4Reconstr"]
    valid_idx = np.flatnonzero(out["validCubes"])
    grid, _ = reconstruct.initialize_cubes(np.float32(0.00025), D, Dc, 0.5, BB)
    for k in sorted({0, len(pl) // 2, len(pl) - 1}):
        g = int(np.flatnonzero((grid["ijk"][valid_idx] == cube_ijk[k]).all(axis=1))[0])          # row of this cube among the valid cubes
        cube = grid[valid_idx[g]]
        pairs_k, w_k = sel_pairs[g][None].astype(np.int64), sel_w[g][None].astype(np.float32)
        assert np.array_equal(pairs_k[0].astype(np.uint16), vp[k])
        X = cvc_oracle.gen_coloredCubes(pairs_k, cube["xyz"][None], cube["resol"][None], P, imgs, D)
        _, Xm = cvc_oracle.preprocess_augmentation(None, X, mean, False, False)
        fused_o, _ = so.nViewPair_SurfaceNet_fn(Xm, params, w_k, N_vp=N, chunk=1)
        p16 = fused_o[0, 0].astype(np.float16)
        votes_o = raypool_oracle.rayPooling_1cube_numpy(P, None, p16, pairs_k[0], cube["xyz"], cube["resol"], 0.46)
        crop = (slice(margin, margin + Dc),) * 3
        p_c, v_c, f_c = p16[crop], votes_o[crop], fused_o[0, 0][crop]
        keep_o = p_c > np.float16(0.46)
        ijk = il[k].astype(np.int64)
        got = np.zeros((Dc, Dc, Dc), bool); got[ijk[:, 0], ijk[:, 1], ijk[:, 2]] = True
        near = np.abs(f_c - 0.46) < 5e-4                                     # a 1e-4 probability difference may flip the threshold / the f16 rounding here
        assert np.array_equal(got[~near], keep_o[~near]), "cube %d: kept-voxel set differs away from the threshold" % k
        both = got & keep_o
        gp = np.zeros((Dc, Dc, Dc), np.float32); gp[ijk[:, 0], ijk[:, 1], ijk[:, 2]] = pl[k].astype(np.float32)
        gv = np.zeros((Dc, Dc, Dc), np.int64); gv[ijk[:, 0], ijk[:, 1], ijk[:, 2]] = vl[k]
        assert np.abs(gp[both] - p_c[both].astype(np.float32)).max() <= 2 ** -10 + 1e-4    # one float16 ulp below 1.0 + the parity bound
        agree = (gv[both] == v_c[both]).mean()
        print("cube %d: %d kept voxels, votes agree on %.4f of them" % (k, int(both.sum()), agree))
        assert agree >= 0.999                                                # votes are exact GIVEN the float16 prediction; flips only where f16 values differ
    last = adapthresh.adapthresh(save_result_fld=str(tmp_path), N_refine_iter=2, D_cube=Dc, init_probThresh=0.5, min_probThresh=0.5,
                                 max_probThresh=0.9, rayPool_thresh=0, beta=6, gamma=0.8, npz_file=out["npz_path"], RGB_visual_ply=False)
    assert os.path.exists(last)
