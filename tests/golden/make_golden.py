#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by EXECUTING THE REFERENCE CODE ITSELF
(/root/reference, read-only) in the build container.  Not run on the GPU box; the fixtures it
writes are committed.  Usage:  python tests/golden/make_golden.py

The reference is python-2 / numpy-1.13 code.  It is run unmodified except for these mechanical
shims, applied in memory:
  * utils/CVC.py: the py2 ``print '...'`` statement at line 74 (visualisation branch, never taken)
    is rewritten to a print() call so the module text compiles; nothing else is touched.
  * utils/rayPooling.py: executed with a numpy proxy whose ``unravel_index`` accepts the removed
    ``dims=`` keyword, ``bool`` alias restored, and ``unique(return_inverse=True)`` flattened to 1-D
    (numpy 2 returns the (N,1) shape of the structured view).
  * nets/layers.py ``__W_5D__``: the function text is extracted and executed alone (the module
    imports theano); ``np.ogrid[:3.0]`` float slices are given ints (numpy 2 rejects float slices).
"""
import os, re, sys, types
import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REF, "utils"))


def load_ref_camera():
    import camera            # utils/camera.py imports as-is under py3 (its doctests run at import)
    return camera


class _NpProxy(types.ModuleType):
    """numpy with the three removed/changed behaviours the 2017 code relies on."""
    def __init__(self):
        super().__init__("numpy_proxy")
        self.bool = bool

    def __getattr__(self, name):
        return getattr(np, name)

    @staticmethod
    def unravel_index(indices, shape=None, dims=None, order="C"):
        return np.unravel_index(indices, shape if shape is not None else dims, order=order)

    @staticmethod
    def unique(ar, return_inverse=False, **kw):
        out = np.unique(ar, return_inverse=return_inverse, **kw)
        if return_inverse:
            return out[0], out[1].ravel()
        return out


def load_ref_raypool():
    import rayPooling
    rayPooling.np = _NpProxy()
    return rayPooling


def load_ref_cvc():
    src = open(os.path.join(REF, "utils", "CVC.py")).read()
    src, n = re.subn(r"print ('error: \[func\]gen_coloredCubes[^\n]*')", r"print(\1)", src)
    assert n == 1
    src = src.replace("import doctest\ndoctest.testmod()", "")
    mod = types.ModuleType("ref_CVC")
    exec(compile(src, os.path.join(REF, "utils", "CVC.py"), "exec"), mod.__dict__)
    return mod


def load_ref_W5D():
    src = open(os.path.join(REF, "nets", "layers.py")).read()
    m = re.search(r"def __W_5D__\(size\):.*?return W\[None,None\]\.astype\(np\.float32\)", src, re.S)
    body = m.group(0).replace("og = np.ogrid[:size, :size, :size]", "og = np.ogrid[:int(size), :int(size), :int(size)]")
    ns = {"np": np}
    exec(body, ns)
    return ns["__W_5D__"]


def load_ref_colorfusion():
    import utils as ref_utils
    return ref_utils.generate_voxelLevelWeighted_coloredCubes


def load_ref_sparse(rp_module):
    """utils/sparseCubes.py imports cPickle / plyfile at module level; dense2sparse and append_dense_2sparseList are
    extracted as text and executed alone.  Mechanical fix: the py2 integer divisions `(D_orig-cube_Dcenter)/2`
    (sparseCubes.py:51) become `//`."""
    src = open(os.path.join(REF, "utils", "sparseCubes.py")).read()
    a = src.index("def dense2sparse(")
    b = src.index("\ndef ", src.index("def append_dense_2sparseList(") + 10)
    body = src[a:b]
    assert body.count("(D_orig-cube_Dcenter)/2") == 2
    body = body.replace("(D_orig-cube_Dcenter)/2", "(D_orig-cube_Dcenter)//2")
    ns = {"np": np, "rayPooling": rp_module}
    exec(compile(body, os.path.join(REF, "utils", "sparseCubes.py"), "exec"), ns)
    return ns["dense2sparse"], ns["append_dense_2sparseList"]


def dtu_cameras():
    cams = np.empty((49, 3, 4), dtype=np.float64)
    for v in range(1, 50):
        cams[v - 1] = np.loadtxt(os.path.join(REF, "inputs/DTU_MVS/SampleSet/MVS Data/Calibration/cal18",
                                              "pos_{:03}.txt".format(v)), dtype=np.float64, delimiter=" ")
    return cams


def synth_image(seed, H, W):
    """Deterministic smooth-ish uint8 RGB image (shared with tests/util.py:synth_image)."""
    from tests.util import synth_image as f
    return f(seed, H, W)


def main():
    from tests import util
    cam = load_ref_camera()
    cams = dtu_cameras()
    np.save(os.path.join(REPO, "surfacenet_b200", "data", "dtu_cal18_cameras.npy"), cams)

    out = {}
    # ---- 1. camera.perspectiveProj -------------------------------------------------------------
    np.random.seed(201611)                         # the doctest's own inputs (camera.py:144-147)
    Ms = np.random.rand(2, 3, 4); pts = np.random.rand(2, 3)
    h, w = cam.perspectiveProj(Ms, pts, return_int_hw=False)
    out["pp_doctest_Ms"], out["pp_doctest_pts"], out["pp_doctest_h"], out["pp_doctest_w"] = Ms, pts, h, w
    rs = np.random.RandomState(7)
    pts = rs.uniform([-73, -197, 472], [129, 183, 810], size=(257, 3))
    vsel = np.array([0, 5, 17, 48])
    h, w, d = cam.perspectiveProj(cams[vsel], pts, return_int_hw=True, return_depth=True)
    out["pp_dtu_pts"], out["pp_dtu_views"], out["pp_dtu_h"], out["pp_dtu_w"], out["pp_dtu_depth"] = pts, vsel, h, w, d
    hf, wf = cam.perspectiveProj(cams[3], pts[0], return_int_hw=False)
    out["pp_dtu_single_hw"] = np.r_[hf, wf]

    # ---- 2. CVC.gen_coloredCubes ----------------------------------------------------------------
    CVC = load_ref_cvc()
    for name, case in util.cvc_cases(cams).items():
        X = CVC.gen_coloredCubes(case["pairs"], case["xyz"], case["resol"], case["cameraPOs"], case["images"], case["D"])
        assert X.dtype == np.float32 and float(np.abs(X - np.round(X)).max()) == 0.0
        out["cvc_" + name] = X.astype(np.uint8)
        _, X2 = CVC.preprocess_augmentation(None, X, util.MEAN6[None, :, None, None, None], augment_ON=False, crop_ON=False)
        out["cvc_" + name + "_pre_sum"] = np.array([X2.astype(np.float64).sum()])
        out["cvc_" + name + "_pre_sample"] = X2.reshape(-1)[::997].copy()

    # ---- 3. rayPooling.rayPooling_1cube_numpy ---------------------------------------------------
    RP = load_ref_raypool()
    cameraTs = np.zeros((cams.shape[0], 3))        # only indexed, never used (rayPooling.py:214)
    for name, case in util.raypool_cases(cams).items():
        votes = RP.rayPooling_1cube_numpy(case["cameraPOs"], np.zeros((case["cameraPOs"].shape[0], 3)), case["pred"],
                                          case["pairs"], case["xyz"], case["resol"], prediction_thresh=case["thresh"])
        out["rp_" + name] = votes.astype(np.uint8)
        print("raypool", name, "votes hist", np.bincount(votes.ravel())[:12])

    # ---- 4. fixed up-sampling kernels -----------------------------------------------------------
    W5 = load_ref_W5D()
    out["W5D_3"], out["W5D_5"] = W5(3), W5(5)

    # ---- 5. colour fusion (utils/utils.py:8-42, "next" row N2) ----------------------------------
    fuse = load_ref_colorfusion()
    rs = np.random.RandomState(11)
    cc = rs.randint(0, 256, size=(2 * 3, 6, 8, 8, 8)).astype(np.float32)
    pr = rs.rand(2, 3, 8, 8, 8).astype(np.float32); ww = (rs.rand(2, 3) + 0.1).astype(np.float32)
    out["cf_cc"], out["cf_pred"], out["cf_w"], out["cf_out"] = cc.astype(np.uint8), pr, ww, fuse(cc, pr, ww)

    # ---- 6. sparseCubes.dense2sparse / append_dense_2sparseList ("next" row N1) -------------------
    d2s, append = load_ref_sparse(RP)
    for name, case in util.sparse_cases(cams).items():
        res = append(prediction_sub=case["pred"], rgb_sub=case["rgb"], param_sub=case["param"], viewPair_sub=case["pairs"],
                     min_prob=case["min_prob"], rayPool_thresh=0, enable_centerCrop=True, cube_Dcenter=case["Dcenter"],
                     enable_rayPooling=True, cameraPOs=cams, cameraTs=np.zeros((cams.shape[0], 3)),
                     prediction_list=[], rgb_list=[], vxl_ijk_list=[], rayPooling_votes_list=[], cube_ijk_np=None, param_np=None, viewPair_np=None)
        pl, rl, il, vl, cube_ijk, param_np, vp_np = res
        out["sp_" + name + "_counts"] = np.array([len(x) for x in pl], np.int64)
        out["sp_" + name + "_pred"] = np.concatenate(pl) if pl else np.zeros(0, np.float16)
        out["sp_" + name + "_rgb"] = np.concatenate(rl) if rl else np.zeros((0, 3), np.uint8)
        out["sp_" + name + "_ijk"] = np.concatenate(il) if il else np.zeros((0, 3), np.uint8)
        out["sp_" + name + "_votes"] = np.concatenate(vl) if vl else np.zeros(0, np.uint8)
        out["sp_" + name + "_cube_ijk"] = np.asarray(cube_ijk)
        out["sp_" + name + "_xyz"] = np.asarray(param_np["xyz"]); out["sp_" + name + "_resol"] = np.asarray(param_np["resol"])
        out["sp_" + name + "_viewPair"] = np.asarray(vp_np)
        assert not pl or (pl[0].dtype == np.float16 and il[0].dtype == np.uint8 and vl[0].dtype == np.uint8)
        print("sparse", name, "kept voxels per non-empty cube", out["sp_" + name + "_counts"])

    np.savez_compressed(os.path.join(HERE, "reference_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "reference_golden.npz"), {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
