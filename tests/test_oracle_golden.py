"""CPU: the oracle restatement against (a) the reference's own doctest known answers and
(b) golden outputs produced by executing the reference code (tests/golden/make_golden.py)."""
import numpy as np
import pytest
from oracle import camera_oracle, cvc_oracle, raypool_oracle, sparse_oracle, surfacenet_oracle
from tests import util


def test_perspectiveProj_reference_doctest_known_answers():
    # utils/camera.py:144-160 -- the only known answers the reference pins on this path
    np.random.seed(201611)
    Ms = np.random.rand(2, 3, 4)
    pts_3D = np.random.rand(2, 3)
    pts_2Dh, pts_2Dw = camera_oracle.perspectiveProj(Ms, pts_3D, return_int_hw=False)
    assert np.allclose(pts_2Dw, np.array([[1.35860185, 0.9878389], [0.64522543, 0.76079278]]))
    pts_2Dh_int, pts_2Dw_int = camera_oracle.perspectiveProj(Ms, pts_3D, return_int_hw=True)
    assert np.allclose(pts_2Dw_int, np.array([[1, 1], [1, 1]]))
    assert np.allclose(np.r_[camera_oracle.perspectiveProj(Ms[1], pts_3D[0], return_int_hw=False)],
                       np.stack((pts_2Dh, pts_2Dw))[:, 1, 0])


def test_perspectiveProj_errors():
    with pytest.raises(ValueError):
        camera_oracle.perspectiveProj(np.zeros((4, 4)), np.zeros((2, 3)))
    with pytest.raises(ValueError):
        camera_oracle.perspectiveProj(np.zeros((3, 4)), np.zeros((2, 2)))


def test_perspectiveProj_golden(golden, cams):
    h, w = camera_oracle.perspectiveProj(golden["pp_doctest_Ms"], golden["pp_doctest_pts"], return_int_hw=False)
    assert np.array_equal(h, golden["pp_doctest_h"]) and np.array_equal(w, golden["pp_doctest_w"])
    h, w, d = camera_oracle.perspectiveProj(cams[golden["pp_dtu_views"]], golden["pp_dtu_pts"], True, True)
    assert h.dtype == np.int64
    assert np.array_equal(h, golden["pp_dtu_h"]) and np.array_equal(w, golden["pp_dtu_w"])
    assert np.array_equal(d, golden["pp_dtu_depth"])
    hw = camera_oracle.perspectiveProj(cams[3], golden["pp_dtu_pts"][0], return_int_hw=False)
    assert np.array_equal(np.r_[hw], golden["pp_dtu_single_hw"])


@pytest.mark.parametrize("name", ["basic", "ragged_sizes", "dup_views", "c1_s32", "outside"])
def test_cvc_oracle_matches_reference_outputs(golden, cams, name):
    case = util.cvc_cases(cams)[name]
    X = cvc_oracle.gen_coloredCubes(case["pairs"], case["xyz"], case["resol"], case["cameraPOs"], case["images"], case["D"])
    assert X.dtype == np.float32 and X.shape == golden["cvc_" + name].shape
    assert np.array_equal(X, golden["cvc_" + name].astype(np.float32))
    _, X2 = cvc_oracle.preprocess_augmentation(None, X, util.MEAN6[None, :, None, None, None], False, False)
    assert X2.dtype == np.float32
    assert np.array_equal(X2.reshape(-1)[::997], golden["cvc_" + name + "_pre_sample"])
    assert X2.astype(np.float64).sum() == golden["cvc_" + name + "_pre_sum"][0]


def test_cvc_cases_cover_out_of_scope(golden):
    # the ragged case must really exercise the zero fill, the basic case must be fully in scope
    assert (golden["cvc_ragged_sizes"].reshape(4, 2, 3, -1).max(axis=2) == 0).mean() > 0.05
    assert golden["cvc_outside"].max() == 0


@pytest.mark.parametrize("name", ["sheet16", "sheet32_dup", "ties", "lowres_collide", "all_ones", "empty",
                                  "none_thresh_f32", "exact_thresh"])
def test_raypool_oracle_matches_reference_outputs(golden, cams, name):
    case = util.raypool_cases(cams)[name]
    votes = raypool_oracle.rayPooling_1cube_numpy(case["cameraPOs"], None, case["pred"], case["pairs"], case["xyz"],
                                                  case["resol"], prediction_thresh=case["thresh"])
    assert np.array_equal(votes.astype(np.uint8), golden["rp_" + name])
    assert votes.max() <= 2 * case["pairs"].shape[0]


def test_raypool_oracle_bad_dims():
    with pytest.raises(ValueError):
        raypool_oracle.rayPooling_1cube_numpy(np.zeros((2, 3, 4)), None, np.zeros((2, 2, 4, 4, 4)), np.array([[0, 1]]),
                                              np.zeros(3, np.float32), np.float32(1))


def test_upsample_kernels_match_reference(golden):
    from surfacenet_b200 import weights
    for k in (3, 5):
        assert np.array_equal(surfacenet_oracle.W_5D(k), golden["W5D_%d" % k])
        assert np.array_equal(weights.upsample_W(k), golden["W5D_%d" % k])


def test_upsample_micro_cases():
    # zero-stuff + fixed taps, 'same' zero border
    import torch
    x = torch.zeros(1, 1, 2, 2, 2); x[0, 0, :, 0, 0] = torch.tensor([1.0, 2.0])
    y4 = surfacenet_oracle.upsample(x, surfacenet_oracle.W_5D(5), 4)[0, 0, :, 0, 0].numpy()
    assert np.allclose(y4, [1, 2 / 3, 1.0, 4 / 3, 2, 4 / 3, 2 / 3, 0], atol=1e-6)
    y2 = surfacenet_oracle.upsample(x, surfacenet_oracle.W_5D(3), 2)[0, 0, :, 0, 0].numpy()
    assert np.allclose(y2, [1, 1.5, 2, 1.0], atol=1e-6)


def test_colorfusion_oracle_matches_reference_outputs(golden):
    out = sparse_oracle.generate_voxelLevelWeighted_coloredCubes(golden["cf_cc"].astype(np.float32), golden["cf_pred"], golden["cf_w"])
    assert out.dtype == np.uint8 and np.array_equal(out, golden["cf_out"])


def _check_sparse(res, golden, name):
    pl, rl, il, vl, cube_ijk, param_np, vp_np = res
    assert np.array_equal(np.array([len(x) for x in pl], np.int64), golden["sp_" + name + "_counts"])
    cat = lambda l, shape, dt: np.concatenate(l) if l else np.zeros(shape, dt)
    assert np.array_equal(cat(pl, 0, np.float16), golden["sp_" + name + "_pred"])
    assert np.array_equal(cat(rl, (0, 3), np.uint8), golden["sp_" + name + "_rgb"])
    assert np.array_equal(cat(il, (0, 3), np.uint8), golden["sp_" + name + "_ijk"])
    assert np.array_equal(cat(vl, 0, np.uint8), golden["sp_" + name + "_votes"])
    assert np.array_equal(np.asarray(cube_ijk), golden["sp_" + name + "_cube_ijk"])
    assert np.array_equal(np.asarray(param_np["xyz"]), golden["sp_" + name + "_xyz"])
    assert np.array_equal(np.asarray(vp_np), golden["sp_" + name + "_viewPair"])
    assert all(a.dtype == np.float16 for a in pl) and all(a.dtype == np.uint8 for a in il + rl + vl)


@pytest.mark.parametrize("name", ["d16", "d32", "all_empty"])
def test_dense2sparse_oracle_matches_reference_outputs(golden, cams, name):
    case = util.sparse_cases(cams)[name]
    res = sparse_oracle.append_dense_2sparseList(case["pred"], case["rgb"], case["param"], case["pairs"], min_prob=case["min_prob"],
                                                 rayPool_thresh=0, enable_centerCrop=True, cube_Dcenter=case["Dcenter"],
                                                 enable_rayPooling=True, cameraPOs=cams, cameraTs=None)
    _check_sparse(res, golden, name)


# ---- "next" row N4: utils/denoising.py, utils/adapthresh.py -------------------------------------------------------------
from oracle import postprocess_oracle as post


@pytest.fixture(scope="module")
def post_golden():
    import os
    return np.load(os.path.join(util.REPO, "tests", "golden", "postprocess_golden.npz"))


def _doc_cluster_inputs():
    ijk = [np.array([[1, 0, 0], [2, 2, 2], [3, 3, 3], [1, 0, 1], [2, 3, 3], [0, 3, 3], [1, 2, 2]]),
           np.array([[0, 2, 3], [0, 1, 0], [0, 0, 0], [0, 3, 3]]), np.array([[0, 2, 3], [0, 1, 0], [0, 2, 3]]),
           np.array([[0, 2, 3], [0, 1, 3], [0, 0, 0], [0, 3, 3], [3, 3, 3]], dtype=np.uint8)]
    mask = [np.array([1, 0, 1, 1, 1, 1, 1], dtype=bool), np.array([1, 1, 0, 1], dtype=bool), np.array([0, 0, 0], dtype=bool),
            np.array([1, 1, 1, 1, 1], dtype=bool)]
    return ijk, mask


def test_cluster_inCube_reference_doctest_known_answers(post_golden):
    # utils/denoising.py:26-38
    ijk, mask = _doc_cluster_inputs()
    lab, n = post.cluster_inCube(ijk, mask)
    assert [l.tolist() for l in lab] == [[2, 0, 4, 2, 4, 1, 3], [2, 1, 0, 2], [0, 0, 0], [2, 2, 1, 2, 3]] and n == [4, 2, 0, 3]
    lab, n = post.cluster_inCube(ijk, mask, neighbor_dist=3)
    assert [l.tolist() for l in lab] == [[2, 0, 1, 2, 1, 1, 1], [2, 1, 0, 2], [0, 0, 0], [2, 2, 1, 2, 3]] and n == [2, 2, 0, 3]
    for nd in (1, 2, 3):
        lab, n = post.cluster_inCube(ijk, mask, neighbor_dist=nd)
        assert np.array_equal(np.concatenate(lab).astype(np.int64), post_golden["doc_cluster_nd%d_labels" % nd])
        assert np.array_equal(np.asarray(n), post_golden["doc_cluster_nd%d_n" % nd])


def test_mark_overlappingLabels_reference_doctest_known_answers():
    # utils/denoising.py:82-94
    cube_ijk = np.array([[1, 6, 8], [2, 6, 8], [2, 7, 8], [2, 5, 8]], dtype=np.uint8)
    ijk = [np.array([[1, 0, 0], [2, 2, 2], [3, 2, 3], [3, 3, 3], [1, 0, 1], [2, 3, 3], [3, 0, 3]], dtype=np.uint8),
           np.array([[0, 2, 3], [0, 1, 3], [0, 0, 0], [0, 3, 3], [1, 0, 3], [3, 3, 0]], dtype=np.uint8),
           np.array([[0, 2, 3], [0, 1, 3], [0, 0, 0], [0, 3, 3]], dtype=np.uint8),
           np.array([[0, 2, 3], [0, 1, 3], [0, 0, 0], [0, 3, 3], [3, 3, 3]], dtype=np.uint8)]
    mask = [np.array([1, 0, 0, 1, 1, 1, 1], dtype=bool), np.array([1, 1, 0, 1, 1, 1], dtype=bool), np.array([0, 0, 0, 0], dtype=bool),
            np.array([1, 1, 1, 1, 1], dtype=bool)]
    ovl, lab = post.mark_overlappingLabels(cube_ijk, ijk, mask, D_cube=4)
    assert ovl == [[2, 3], [1, 2], [], [2]]
    assert [l.tolist() for l in lab] == [[1, 0, 0, 2, 1, 2, 3], [1, 1, 0, 1, 2, 3], [0, 0, 0, 0], [2, 2, 1, 2, 3]]


def test_denoise_crossCubes_reference_doctest_known_answers():
    # utils/denoising.py:160-172
    cube_ijk = np.array([[1, 6, 8], [2, 6, 8], [2, 7, 8], [2, 5, 8]], dtype=np.uint8)
    ijk = [np.array([[1, 0, 0], [2, 2, 2], [3, 3, 3], [1, 0, 1], [2, 3, 3]], dtype=np.uint8),
           np.array([[0, 2, 3], [0, 1, 3], [0, 0, 0], [0, 3, 3], [3, 3, 0]], dtype=np.uint8),
           np.array([[0, 2, 3], [0, 1, 3], [0, 0, 0], [0, 3, 3]], dtype=np.uint8),
           np.array([[0, 2, 3], [0, 1, 3], [0, 0, 0], [0, 3, 3], [3, 3, 3]], dtype=np.uint8)]
    mask = [np.array([1, 0, 1, 1, 1], dtype=bool), np.array([1, 1, 0, 1, 1], dtype=bool), np.array([0, 0, 0, 0], dtype=bool),
            np.array([1, 1, 1, 1, 1], dtype=bool)]
    keep = post.denoise_crossCubes(cube_ijk, ijk, mask, D_cube=4)
    assert [k.tolist() for k in keep] == [[False, False, True, False, True], [True, True, False, True, False],
                                          [False, False, False, False], [True, True, False, True, False]]


def test_adapthresh_helpers_reference_doctest_known_answers():
    # utils/adapthresh.py:33-41
    Occ = np.array([[1, 5, 2], [5, 2, 0], [0, 1, 5], [2, 1, 1], [4, 5, 5]])
    gt = Occ[2:3] - np.array([[0, 0, 3]])
    res = post.access_partial_Occupancy_ijk(Occ.copy(), (-1, 0, 1), D_cube=6)
    assert gt.shape == res.shape and np.array_equal(gt, res)
    assert np.array_equal(Occ[1:4, :2], post.access_partial_Occupancy_ijk(Occ[:, :2].copy(), (0, -1), D_cube=6))
    # utils/adapthresh.py:73-76
    ijk1 = np.array([[1, 0], [2, 3], [222, 666], [0, 0]])
    ijk2 = np.array([[11, 10], [2, 3], [22, 66], [0, 0], [7, 17]])
    assert post.sparseOccupancy_AND_XOR(ijk1, ijk2) == (2, 5)


@pytest.mark.parametrize("name", ["g322_d12", "g233_d16_shuffled", "g222_d13_odd_dup", "g141_d10_thin"])
def test_postprocess_oracle_matches_reference_outputs(post_golden, name):
    case = util.post_cases()[name]
    sc = case["scene"]
    g = lambda k: post_golden["post_%s_%s" % (name, k)]
    cat = lambda lst: np.concatenate([np.asarray(x) for x in lst])
    mask0 = post.filter_voxels([], sc["pred_list"], 0.7, sc["votes_list"], case["rp"])
    assert np.array_equal(cat(mask0).astype(np.uint8), g("mask_tau"))
    for nd in (1, 3):
        ovl, lab = post.mark_overlappingLabels(sc["cube_ijk"], sc["ijk_list"], mask0, sc["D"], neighbor_dist=nd)
        assert np.array_equal(cat(lab).astype(np.int64), g("labels_nd%d" % nd))
        assert np.array_equal(cat([np.isin(l, o) for l, o in zip(lab, ovl)]).astype(np.uint8), g("ovl_nd%d" % nd))
    keep = post.denoise_crossCubes(sc["cube_ijk"], sc["ijk_list"], mask0, sc["D"])
    assert np.array_equal(cat(keep).astype(np.uint8), g("denoised_tau"))
    res = post.adapthresh_core(sc["pred_list"], sc["ijk_list"], sc["votes_list"], sc["cube_ijk"], case["iters"], sc["D"], case["init"],
                               case["init"], case["maxp"], case["rp"], case["beta"])
    assert np.array_equal(cat(res["init_denoised"]).astype(np.uint8), g("ada_init_denoised"))
    for i, it in enumerate(res["iters"]):
        assert np.array_equal(it["probThresh"], g("ada_thresh")[i])
        assert np.array_equal(cat(it["denoised"]).astype(np.uint8), g("ada_denoised")[i])


# ---- "next" row N3: view-pair angles / selection / early rejection host code ------------------------------------------------
from oracle import selection_oracle as sel


@pytest.fixture(scope="module")
def select_golden():
    import os
    return np.load(os.path.join(util.REPO, "tests", "golden", "select_golden.npz"))


def test_viewPairAngles_reference_doctest_known_answers(select_golden, cams):
    import math                                                        # utils/camera.py:290-294
    pts = np.array([[0, 0, 0], [1, 1, 1]], dtype=np.float32)
    cT = np.array([[0, 0, 1], [0, 1, 1], [1, 0, 1]], dtype=np.float32)
    a = sel.viewPairAngles_wrt_pts(cT, pts)
    assert a.dtype == np.float32 and np.allclose(a * 180 / math.pi, [[45., 45., 60.], [45., 45., 90.]], atol=1e-4)
    assert np.array_equal(a, select_golden["ang_doc"])
    case = util.select_case(cams)
    cTs = select_golden["cameraTs"][case["views"]]
    assert np.array_equal(sel.viewPairAngles_wrt_pts(cTs, case["centers"].astype(np.float32)), select_golden["ang_dtu64"])
    assert np.array_equal(sel.viewPairAngles_wrt_pts(cTs.astype(np.float32), case["centers"].astype(np.float32)), select_golden["ang_dtu32"])


def test_argmaxN_viewPairs_reference_doctest_known_answers(select_golden, cams):
    vp = sel.k_combination_np(range(3), k=2)                           # utils/viewPairSelection.py:17-31
    w = np.array([[3, 1, 2], [0, -1, 70]])
    a, b = sel.argmaxN_viewPairs(vp, w, 1)
    assert a.tolist() == [[[0, 1]], [[1, 2]]] and b.tolist() == [[3], [70]]
    a, b = sel.argmaxN_viewPairs(vp, w, 2)
    assert a.tolist() == [[[1, 2], [0, 1]], [[0, 1], [1, 2]]] and b.tolist() == [[2, 3], [0, 70]]
    case = util.select_case(cams)
    a, b = sel.argmaxN_viewPairs(case["viewPairs"], case["w_rand"], 5)
    assert np.array_equal(a, select_golden["argmax_rand_pairs"]) and np.array_equal(b, select_golden["argmax_rand_w"])


def test_selection_oracle_matches_reference_outputs(select_golden, cams):
    case = util.select_case(cams)
    g = select_golden
    assert np.array_equal(sel.preprocess_patches(np.zeros((2, 2, 5, 3)), np.array([1, 2, 3])), g["pre_doc"])
    img = util.synth_image(3, 300, 400)
    assert np.array_equal(sel.cropImgPatches_rate1(img, 64, (case["crop_ch"], case["crop_cw"])), g["crop_patches"])
    emb, inscope = sel.patch2embedding(case["images"], case["h_corner"], case["w_corner"], util.fake_patch2embedding_fn, util.MEAN_BGR,
                                       case["N_cubes"], len(case["views"]), 16, 64, 5, case["center_hw"])
    assert np.array_equal(inscope, g["er_inscope"]) and np.array_equal(emb, g["er_emb"])
    dis = sel.embeddingPairs2simil(emb, len(case["views"]), util.fake_pair2simil_fn, 7)
    assert np.array_equal(dis, g["er_dissim"])
    assert np.array_equal(sel.selectFromSimilarity(dis, 3), g["er_select"]) and 0 < g["er_select"].sum() < case["N_cubes"]
    from oracle import surfacenet_oracle
    from surfacenet_b200 import weights
    params = weights.synthetic_params(0)
    fn = lambda f, n_samples_perGroup: surfacenet_oracle.viewPair_relativeImpt_fn(f, params, n_samples_perGroup)
    s, w = sel.viewPairSelection(g["cameraTs"][case["views"]], g["vps_e"], g["vps_d"], g["vps_valid"], case["centers"].astype(np.float32), fn,
                                 4 * case["viewPairs"].shape[0] + 3, 4, case["viewPairs"])
    assert np.array_equal(s, g["vps_sel"]) and np.array_equal(w, g["vps_w"])
