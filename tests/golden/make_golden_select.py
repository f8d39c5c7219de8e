#!/usr/bin/env python
"""Golden fixtures for the "next" row N3 (selection side), produced by EXECUTING THE REFERENCE's numpy code (/root/reference,
read-only): utils/camera.py:viewPairAngles_wrt_pts, utils/viewPairSelection.py, utils/earlyRejection.py, utils/image.py.
Writes tests/golden/select_golden.npz (committed).   Usage:  python tests/golden/make_golden_select.py

Shims, in memory only: module-level doctest.testmod() dropped; utils/image.py `patchSize / 2` -> `//` (python-2 integer division)
and a numpy proxy restoring the removed aliases np.int / np.bool; the similarityNet callables (Theano) are replaced by the
deterministic stand-ins of tests/util.py (fake_patch2embedding_fn / fake_pair2simil_fn), which is what pins the cropping,
preprocessing, batching and index ordering of the reference's host code."""
import os, sys, types
import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REF, "utils"))


class _NpProxy(types.ModuleType):
    def __init__(self):
        super().__init__("numpy_proxy")
        self.bool = bool
        self.int = int

    def __getattr__(self, name):
        return getattr(np, name)


def load(name, repl=()):
    src = open(os.path.join(REF, "utils", name + ".py")).read().replace("import doctest\ndoctest.testmod()", "")
    for a, b in repl:
        assert src.count(a) >= 1, a
        src = src.replace(a, b)
    mod = types.ModuleType("ref_" + name)
    exec(compile(src, os.path.join(REF, "utils", name + ".py"), "exec"), mod.__dict__)
    mod.np = _NpProxy()
    return mod


def main():
    from tests import util
    import camera                                   # utils/camera.py imports as-is
    ref_utils = load("utils")
    sys.modules["utils"] = ref_utils
    image = load("image", [("patchSize_r = patchSize / 2", "patchSize_r = patchSize // 2")])
    sys.modules["image"] = image
    er = load("earlyRejection")
    vps = load("viewPairSelection")
    out = {}
    cams = util.dtu_cameras()
    cameraTs = camera.cameraPs2Ts(cams)
    out["cameraTs"] = cameraTs

    # viewPairAngles_wrt_pts: the doctest inputs (camera.py:290-294) and a DTU case in float64 / float32
    pts = np.array([[0, 0, 0], [1, 1, 1]], dtype=np.float32)
    cT = np.array([[0, 0, 1], [0, 1, 1], [1, 0, 1]], dtype=np.float32)
    out["ang_doc"] = camera.viewPairAngles_wrt_pts(cT, pts)
    case = util.select_case(cams)
    out["ang_dtu64"] = camera.viewPairAngles_wrt_pts(cameraTs[case["views"]], case["centers"].astype(np.float32))
    out["ang_dtu32"] = camera.viewPairAngles_wrt_pts(cameraTs[case["views"]].astype(np.float32), case["centers"].astype(np.float32))

    # __argmaxN_viewPairs__: doctest (viewPairSelection.py:17-31) + random float32 weights
    vp3 = ref_utils.k_combination_np(range(3), k=2)
    w = np.array([[3, 1, 2], [0, -1, 70]])
    for N in (1, 2):
        a, b = vps.__argmaxN_viewPairs__(vp3, w, N)
        out["argmax_doc%d_pairs" % N], out["argmax_doc%d_w" % N] = a, b
    a, b = vps.__argmaxN_viewPairs__(case["viewPairs"], case["w_rand"], 5)
    out["argmax_rand_pairs"], out["argmax_rand_w"] = a, b

    # image helpers
    out["pre_doc"] = image.preprocess_patches(np.zeros((2, 2, 5, 3)), mean_BGR=np.array([1, 2, 3]))
    img = util.synth_image(3, 300, 400)
    patches = image.cropImgPatches(img=img, range_h=case["range_h"], range_w=case["range_w"], patchSize=64, pyramidRate=1, interp_order=2,
                                   cubeCenter_hw=(case["crop_ch"], case["crop_cw"]))
    out["crop_patches"] = patches

    # earlyRejection.patch2embedding / embeddingPairs2simil / selectFromSimilarity with deterministic stand-in networks
    emb, inscope = er.patch2embedding(case["images"], case["h_corner"], case["w_corner"], util.fake_patch2embedding_fn, util.MEAN_BGR,
                                      case["N_cubes"], len(case["views"]), 16, patchSize=64, batchSize=5, cubeCenter_hw=case["center_hw"])
    out["er_emb"], out["er_inscope"] = emb, inscope
    dis = er.embeddingPairs2simil(embeddings=emb, embeddingPair2simil_fn=util.fake_pair2simil_fn, inScope_cubes_vs_views=inscope,
                                  viewPairs=case["viewPairs"], N_views=len(case["views"]), batchSize=7)
    out["er_dissim"] = dis
    out["er_select"] = er.selectFromSimilarity(dissimilarityProb=dis, N_viewPairs4inference=3)
    print("inScope per view", inscope.sum(axis=0), "valid cubes", out["er_select"].sum(), "of", case["N_cubes"])

    # viewPairSelection with the oracle's relative-importance network as the callable
    from oracle import surfacenet_oracle
    from surfacenet_b200 import weights
    params = weights.synthetic_params(0)
    rs = np.random.RandomState(77)
    e128 = rs.randn(case["N_cubes"], len(case["views"]), 128).astype(np.float32)
    d = rs.rand(case["N_cubes"], case["viewPairs"].shape[0]).astype(np.float32)
    valid = rs.rand(case["N_cubes"]) < 0.7
    fn = lambda f, n_samples_perGroup: surfacenet_oracle.viewPair_relativeImpt_fn(f, params, n_samples_perGroup)
    sel, sw = vps.viewPairSelection(cameraTs_np=cameraTs[case["views"]], e_viewPairs=e128, d_viewPairs=d, validCubes=valid,
                                    cubeCenters_xyz=case["centers"].astype(np.float32), viewPair_relativeImpt_fn=fn, batchSize=4 * case["viewPairs"].shape[0] + 3,
                                    N_viewPairs4inference=4, viewPairs=case["viewPairs"])
    out["vps_e"], out["vps_d"], out["vps_valid"], out["vps_sel"], out["vps_w"] = e128, d, valid, sel, sw
    np.savez_compressed(os.path.join(HERE, "select_golden.npz"), **out)
    print("wrote select_golden.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
