"""GPU parity tests of the "next" row N4 (utils/denoising.py, utils/adapthresh.py, sparseCubes.filter_voxels): the CUDA path
(csrc/postprocess.cu through the C ABI) against the reference's doctest known answers, the golden vectors produced by
executing the reference code, and the CPU oracle on larger scenes.  Bar: bit-exact (integer / byte / float16-compare work)."""
import os
import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def post_golden():
    return np.load(os.path.join(util.REPO, "tests", "golden", "postprocess_golden.npz"))


def cat(lst, dtype=None):
    a = np.concatenate([np.asarray(x) for x in lst])
    return a if dtype is None else a.astype(dtype)


# ---- the reference's own doctests, run against the GPU implementation ---------------------------------------------------
def test_cluster_inCube_reference_doctest():
    from surfacenet_b200 import denoising                      # utils/denoising.py:26-38
    ijk = [np.array([[1, 0, 0], [2, 2, 2], [3, 3, 3], [1, 0, 1], [2, 3, 3], [0, 3, 3], [1, 2, 2]]),
           np.array([[0, 2, 3], [0, 1, 0], [0, 0, 0], [0, 3, 3]]), np.array([[0, 2, 3], [0, 1, 0], [0, 2, 3]]),
           np.array([[0, 2, 3], [0, 1, 3], [0, 0, 0], [0, 3, 3], [3, 3, 3]], dtype=np.uint8)]
    mask = [np.array([1, 0, 1, 1, 1, 1, 1], dtype=bool), np.array([1, 1, 0, 1], dtype=bool), np.array([0, 0, 0], dtype=bool),
            np.array([1, 1, 1, 1, 1], dtype=bool)]
    lab, n = denoising.__cluster_inCube__(ijk, mask)
    assert [l.tolist() for l in lab] == [[2, 0, 4, 2, 4, 1, 3], [2, 1, 0, 2], [0, 0, 0], [2, 2, 1, 2, 3]] and n == [4, 2, 0, 3]
    assert lab[0].dtype == np.uint32 and lab[2].dtype == np.float64
    lab, n = denoising.__cluster_inCube__(ijk, mask, neighbor_dist=3)
    assert [l.tolist() for l in lab] == [[2, 0, 1, 2, 1, 1, 1], [2, 1, 0, 2], [0, 0, 0], [2, 2, 1, 2, 3]] and n == [2, 2, 0, 3]


def test_mark_overlappingLabels_reference_doctest():
    from surfacenet_b200 import denoising                      # utils/denoising.py:82-94
    cube_ijk = np.array([[1, 6, 8], [2, 6, 8], [2, 7, 8], [2, 5, 8]], dtype=np.uint8)
    ijk = [np.array([[1, 0, 0], [2, 2, 2], [3, 2, 3], [3, 3, 3], [1, 0, 1], [2, 3, 3], [3, 0, 3]], dtype=np.uint8),
           np.array([[0, 2, 3], [0, 1, 3], [0, 0, 0], [0, 3, 3], [1, 0, 3], [3, 3, 0]], dtype=np.uint8),
           np.array([[0, 2, 3], [0, 1, 3], [0, 0, 0], [0, 3, 3]], dtype=np.uint8),
           np.array([[0, 2, 3], [0, 1, 3], [0, 0, 0], [0, 3, 3], [3, 3, 3]], dtype=np.uint8)]
    mask = [np.array([1, 0, 0, 1, 1, 1, 1], dtype=bool), np.array([1, 1, 0, 1, 1, 1], dtype=bool), np.array([0, 0, 0, 0], dtype=bool),
            np.array([1, 1, 1, 1, 1], dtype=bool)]
    ovl, lab = denoising.__mark_overlappingLabels__(cube_ijk, ijk, mask, D_cube=4)
    assert ovl == [[2, 3], [1, 2], [], [2]]
    assert [l.tolist() for l in lab] == [[1, 0, 0, 2, 1, 2, 3], [1, 1, 0, 1, 2, 3], [0, 0, 0, 0], [2, 2, 1, 2, 3]]


def test_denoise_crossCubes_reference_doctest():
    from surfacenet_b200 import denoising                      # utils/denoising.py:160-172
    cube_ijk = np.array([[1, 6, 8], [2, 6, 8], [2, 7, 8], [2, 5, 8]], dtype=np.uint8)
    ijk = [np.array([[1, 0, 0], [2, 2, 2], [3, 3, 3], [1, 0, 1], [2, 3, 3]], dtype=np.uint8),
           np.array([[0, 2, 3], [0, 1, 3], [0, 0, 0], [0, 3, 3], [3, 3, 0]], dtype=np.uint8),
           np.array([[0, 2, 3], [0, 1, 3], [0, 0, 0], [0, 3, 3]], dtype=np.uint8),
           np.array([[0, 2, 3], [0, 1, 3], [0, 0, 0], [0, 3, 3], [3, 3, 3]], dtype=np.uint8)]
    mask = [np.array([1, 0, 1, 1, 1], dtype=bool), np.array([1, 1, 0, 1, 1], dtype=bool), np.array([0, 0, 0, 0], dtype=bool),
            np.array([1, 1, 1, 1, 1], dtype=bool)]
    keep = denoising.denoise_crossCubes(cube_ijk, ijk, mask, D_cube=4)
    assert [k.tolist() for k in keep] == [[False, False, True, False, True], [True, True, False, True, False],
                                          [False, False, False, False], [True, True, False, True, False]]
    assert all(k.dtype == bool for k in keep)


# ---- golden vectors produced by the reference code ----------------------------------------------------------------------
@pytest.mark.parametrize("name", ["g322_d12", "g233_d16_shuffled", "g222_d13_odd_dup", "g141_d10_thin"])
def test_post_matches_reference_outputs(post_golden, name):
    from surfacenet_b200 import adapthresh, denoising
    from surfacenet_b200.sparse_device import DeviceSparseCubes
    case = util.post_cases()[name]
    sc = case["scene"]
    g = lambda k: post_golden["post_%s_%s" % (name, k)]
    dsc = DeviceSparseCubes(sc["cube_ijk"], sc["ijk_list"], sc["pred_list"], sc["votes_list"])
    mask0 = dsc.filter_voxels(None, prob_thresh=0.7, rayPool_thresh=case["rp"])
    assert np.array_equal(mask0.cpu().numpy(), g("mask_tau"))
    mask0_l = dsc.split(mask0, bool)
    for nd in (1, 3):
        ovl, lab = denoising.__mark_overlappingLabels__(sc["cube_ijk"], sc["ijk_list"], mask0_l, sc["D"], neighbor_dist=nd)
        assert np.array_equal(cat(lab, np.int64), g("labels_nd%d" % nd))
        assert np.array_equal(cat([np.isin(l, o) for l, o in zip(lab, ovl)], np.uint8), g("ovl_nd%d" % nd))
    keep = denoising.denoise_crossCubes(sc["cube_ijk"], sc["ijk_list"], mask0_l, sc["D"])
    assert np.array_equal(cat(keep, np.uint8), g("denoised_tau"))
    seen = []
    res = adapthresh.adapthresh_lists(sc["pred_list"], sc["ijk_list"], sc["votes_list"], sc["cube_ijk"], case["iters"], sc["D"],
                                      case["init"], case["maxp"], case["rp"], case["beta"],
                                      on_iteration=lambda i, m, d, a: seen.append((i, cat(d, np.uint8))))
    assert np.array_equal(cat(res["init_denoised"], np.uint8), g("ada_init_denoised"))
    assert len(seen) == case["iters"]
    for i, d in seen:
        assert np.array_equal(d, g("ada_denoised")[i]), "iteration %d" % i
    assert np.array_equal(res["probThresh"], g("ada_thresh")[-1])
    # one call running all iterations on the device == the iteration-by-iteration loop
    import torch
    init_mask = dsc.filter_voxels(None, prob_thresh=case["init"], rayPool_thresh=case["rp"])
    mask = init_mask.clone()
    thresh = torch.full((dsc.C,), float(case["init"]), dtype=torch.float64, device="cuda")
    arg = dsc.adapthresh(init_mask, mask, thresh, sc["D"], case["maxp"], case["beta"], n_iter=case["iters"], want_argmin=True)
    assert np.array_equal(thresh.cpu().numpy(), g("ada_thresh")[-1])
    assert np.array_equal(arg.cpu().numpy(), res["argmin"])
    assert np.array_equal(cat(dsc.split(dsc.denoise(mask, sc["D"])["keep"], np.uint8)), g("ada_denoised")[-1])


# ---- the oracle on larger scenes (counts above 2048: the float16 cost accumulation rounds) ------------------------------
def test_post_full_size_d52_vs_oracle():
    from oracle import postprocess_oracle as post
    from surfacenet_b200 import adapthresh, denoising
    sc = util.sparse_scene((2, 2, 2), 52, seed=7, floaters=40, thick=0.05)          # cube_Dcenter of 64^3 cubes (params.py:107)
    n_vox = sum(x.shape[0] for x in sc["ijk_list"])
    assert n_vox > 100000
    mask_o = post.filter_voxels([], sc["pred_list"], 0.7, sc["votes_list"], 8)
    keep_o = post.denoise_crossCubes(sc["cube_ijk"], sc["ijk_list"], mask_o, 52)
    keep = denoising.denoise_crossCubes(sc["cube_ijk"], sc["ijk_list"], mask_o, 52)
    assert np.array_equal(cat(keep), cat(keep_o)) and 0 < cat(keep).sum() < cat(mask_o).sum()
    # main_reconstruct.py:172 passes D_cube = 64 for 52^3 centre cubes: mirror that call too
    assert np.array_equal(cat(denoising.denoise_crossCubes(sc["cube_ijk"], sc["ijk_list"], mask_o, 64)),
                          cat(post.denoise_crossCubes(sc["cube_ijk"], sc["ijk_list"], mask_o, 64)))
    lab, n = denoising.__cluster_inCube__(sc["ijk_list"], mask_o, neighbor_dist=2)
    lab_o, n_o = post.cluster_inCube(sc["ijk_list"], mask_o, neighbor_dist=2)
    assert n == n_o and np.array_equal(cat(lab), cat(lab_o))
    ref = post.adapthresh_core(sc["pred_list"], sc["ijk_list"], sc["votes_list"], sc["cube_ijk"], 2, 52, 0.5, 0.5, 0.9, 6, 6)
    res = adapthresh.adapthresh_lists(sc["pred_list"], sc["ijk_list"], sc["votes_list"], sc["cube_ijk"], 2, 52, 0.5, 0.9, 6, 6)
    assert np.array_equal(res["probThresh"], ref["iters"][-1]["probThresh"])
    assert np.array_equal(res["argmin"], np.stack([it["argmin"] for it in ref["iters"]]))
    assert np.array_equal(cat(res["mask"]), cat(ref["iters"][-1]["mask"]))
    assert np.array_equal(cat(res["denoised"]), cat(ref["iters"][-1]["denoised"]))
    # (the float16 costs of 52^3 cubes overflow to -inf exactly as in the reference: every cube takes the first perturbation)
    # cube_Dcenter of 32^3 cubes (params.py:107): counts in the thousands, no overflow, thresholds diverge
    sc = util.sparse_scene((3, 2, 2), 26, seed=7, floaters=20, thick=0.05)
    ref = post.adapthresh_core(sc["pred_list"], sc["ijk_list"], sc["votes_list"], sc["cube_ijk"], 3, 26, 0.5, 0.5, 0.9, 4, 6)
    res = adapthresh.adapthresh_lists(sc["pred_list"], sc["ijk_list"], sc["votes_list"], sc["cube_ijk"], 3, 26, 0.5, 0.9, 4, 6)
    assert np.array_equal(res["probThresh"], ref["iters"][-1]["probThresh"])
    assert np.array_equal(res["argmin"], np.stack([it["argmin"] for it in ref["iters"]]))
    assert np.array_equal(cat(res["denoised"]), cat(ref["iters"][-1]["denoised"]))
    assert len(set(res["probThresh"].tolist())) > 1, "the scene should move some thresholds"


def test_post_large_scene_properties():
    """6x6x3 grid of 52^3 cubes (~1.6 M voxels): size-independent properties -- denoising is idempotent and only removes
    voxels, the refinement only shrinks masks, thresholds stay within [init - 0.1*iters, max]."""
    import torch
    from surfacenet_b200.sparse_device import DeviceSparseCubes
    sc = util.sparse_scene((6, 6, 3), 52, seed=11, floaters=60, thick=0.05)
    dsc = DeviceSparseCubes(sc["cube_ijk"], sc["ijk_list"], sc["pred_list"], sc["votes_list"])
    assert dsc.N > 1000000
    m0 = dsc.filter_voxels(None, prob_thresh=0.5, rayPool_thresh=6)
    k1 = dsc.denoise(m0, 52)["keep"]
    k2 = dsc.denoise(k1, 52)["keep"]
    assert torch.equal(k1, k2) and bool((k1 <= m0).all()) and 0 < int(k1.sum()) < int(m0.sum())
    mask = m0.clone()
    thresh = torch.full((dsc.C,), 0.5, dtype=torch.float64, device="cuda")
    prev = mask.clone()
    for it in range(3):
        dsc.adapthresh(m0, mask, thresh, 52, 0.9, 6, n_iter=1)
        assert bool((mask <= prev).all())
        prev = mask.clone()
    t = thresh.cpu().numpy()
    assert t.max() <= 0.9 and t.min() >= 0.5 - 0.3 - 1e-12
    # filter_voxels with a per-cube threshold list == the mask the refinement ended with
    again = dsc.filter_voxels(m0.clone(), prob_thresh=thresh)
    assert bool((mask <= again).all())


def test_filter_voxels_device_matches_host_contract():
    from surfacenet_b200 import sparseCubes
    from surfacenet_b200.sparse_device import DeviceSparseCubes
    sc = util.sparse_scene((2, 2, 1), 12, seed=5)
    dsc = DeviceSparseCubes(sc["cube_ijk"], sc["ijk_list"], sc["pred_list"], sc["votes_list"])
    per_cube = [0.5, 0.61, 0.7001, 0.9]
    for kw in (dict(prob_thresh=0.7), dict(prob_thresh=per_cube), dict(rayPool_thresh=8.0), dict(prob_thresh=0.55, rayPool_thresh=3)):
        host = sparseCubes.filter_voxels([], sc["pred_list"] if "prob_thresh" in kw else None, kw.get("prob_thresh"),
                                         sc["votes_list"] if "rayPool_thresh" in kw else None, kw.get("rayPool_thresh"))
        dev = dsc.filter_voxels(None, **kw)
        assert np.array_equal(dev.cpu().numpy().astype(bool), cat(host)), kw


def test_adapthresh_file_contract(tmp_path):
    """utils/adapthresh.py:91 end to end: NPZ in, PLY files out; vertex counts == the denoised masks."""
    from surfacenet_b200 import adapthresh, sparseCubes
    sc = util.sparse_scene((2, 2, 2), 16, seed=9)
    npz = str(tmp_path / "model9-49views.npz")
    vp = np.zeros((len(sc["ijk_list"]), 5, 2), np.uint16)
    sparseCubes.save_sparseCubes(npz, sc["pred_list"], sc["rgb_list"], sc["ijk_list"], sc["votes_list"], sc["cube_ijk"], sc["param"], vp)
    last = adapthresh.adapthresh(save_result_fld=str(tmp_path), N_refine_iter=3, D_cube=16, init_probThresh=0.5, min_probThresh=0.5,
                                 max_probThresh=0.9, rayPool_thresh=4, beta=6, gamma=0.8, npz_file=npz, RGB_visual_ply=True)
    fld = os.path.join(str(tmp_path), "adapThresh_gamma0.8_beta6")
    assert last == os.path.join(fld, "iter2.ply")
    assert sorted(os.listdir(fld)) == sorted(["initialization.ply"] + ["iter%d.ply" % i for i in range(3)] +
                                             ["iter%d_tmprgb4debug.ply" % i for i in range(3)])
    res = adapthresh.adapthresh_lists(sc["pred_list"], sc["ijk_list"], sc["votes_list"], sc["cube_ijk"], 3, 16, 0.5, 0.9, 4, 6)
    xyz, rgb = util.read_ply(last)
    keep = cat(res["denoised"])
    assert xyz.shape[0] == int(keep.sum())
    exp = np.vstack([sc["ijk_list"][c][res["denoised"][c]] * sc["param"][c]["resol"] + sc["param"][c]["xyz"][None, :]
                     for c in range(len(sc["ijk_list"]))])
    assert np.array_equal(xyz, exp.astype(np.float32)) and np.array_equal(rgb, np.vstack(sc["rgb_list"])[keep])


def test_post_errors_and_empty():
    from surfacenet_b200 import denoising
    from surfacenet_b200.sparse_device import DeviceSparseCubes
    with pytest.raises(ValueError):
        DeviceSparseCubes(np.zeros((2, 3), np.int32), [np.zeros((1, 3), np.uint8)])
    with pytest.raises(ValueError):
        DeviceSparseCubes(np.zeros((1, 3), np.int32), [np.zeros((4, 2), np.uint8)])
    dsc = DeviceSparseCubes(np.zeros((1, 3), np.int32), [np.zeros((2, 3), np.uint8)])
    with pytest.raises(ValueError):
        dsc.filter_voxels(None, prob_thresh=0.5)                   # no predictions given
    with pytest.raises(ValueError):
        dsc.denoise(dsc.upload_mask([np.ones(2, bool)]), 4, neighbor_dist=4)
    with pytest.raises(Warning):
        dsc.upload_mask([np.ones(3, bool)])
    # no cubes / cubes without voxels / nothing masked
    assert denoising.denoise_crossCubes(np.zeros((0, 3), np.int32), [], [], 8) == []
    keep = denoising.denoise_crossCubes(np.array([[0, 0, 0], [1, 0, 0]]), [np.zeros((0, 3), np.uint8), np.array([[1, 2, 3]], np.uint8)],
                                        [np.zeros(0, bool), np.zeros(1, bool)], 8)
    assert [k.tolist() for k in keep] == [[], [False]]


def test_from_flat_equals_list_constructor():
    """DeviceSparseCubes.from_flat (the NPZ / infer_batch_sparse arrays, no python lists) == the list constructor."""
    import torch
    from surfacenet_b200.sparse_device import DeviceSparseCubes
    sc = util.sparse_scene((2, 2, 2), 16, seed=13)
    a = DeviceSparseCubes(sc["cube_ijk"], sc["ijk_list"], sc["pred_list"], sc["votes_list"])
    off = np.concatenate([[0], np.cumsum([x.shape[0] for x in sc["ijk_list"]])]).astype(np.uint32)     # cube_1st_vxlIndx_np dtype
    b = DeviceSparseCubes.from_flat(sc["cube_ijk"], off, np.vstack(sc["ijk_list"]), np.concatenate(sc["pred_list"]), np.concatenate(sc["votes_list"]))
    c = DeviceSparseCubes.from_flat(a.cube_ijk, a.offsets, a.ijk, a.pred, a.votes, grid_extent=16)       # cuda tensors in
    for x in (b, c):
        assert (x.C, x.N) == (a.C, a.N) and x.G in (a.G, 16)
        m = x.filter_voxels(None, prob_thresh=0.6, rayPool_thresh=4)
        assert torch.equal(m, a.filter_voxels(None, prob_thresh=0.6, rayPool_thresh=4))
        assert torch.equal(x.denoise(m, 16)["keep"], a.denoise(m, 16)["keep"])
    with pytest.raises(ValueError):
        DeviceSparseCubes.from_flat(sc["cube_ijk"], off[:-1], np.vstack(sc["ijk_list"]))
