#!/usr/bin/env python
"""Golden fixtures for the "next" row N4 (utils/denoising.py, utils/adapthresh.py), produced by EXECUTING THE REFERENCE CODE
(/root/reference, read-only) in the build container.  Writes tests/golden/postprocess_golden.npz (committed).
Usage:  python tests/golden/make_golden_post.py

The reference is python-2 / numpy-1.13 code; it is executed from its own source text with these mechanical shims,
applied in memory only:
  * `d.has_key(k)` -> `(k in d)`;  the python-2 integer divisions `3**3/2`, `(D_cube / 2)`, `D_mid = D_cube / 2` -> `//`
  * py2 `print '...'` statement in adapthresh.py:177 -> print();  the module-level doctest.testmod() calls are dropped
    (the doctest known answers are checked separately in tests/test_oracle_golden.py)
  * numpy proxy: `np.bool` alias restored, `np.in1d` -> isin(...).ravel();  `scipy.ndimage.measurements` -> scipy.ndimage
  * adapthresh.py imports `sparseCubes` (needs cPickle/plyfile): replaced by a stub holding the reference's own
    filter_voxels (extracted from the source text), load_sparseCubes (returns the test scene) and save_sparseCubes_2ply
    (records the masks it is asked to write instead of writing PLY files).
numpy-2 note: element_cost accumulates counts in float16 (adapthresh.py:141,162,165); numpy 1.13 and numpy 2 round
identically while every count stays <= 2048, which is asserted here for every case.
"""
import os, re, sys, tempfile, types
import numpy as np
import scipy.ndimage as ndim

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)


class _NpProxy(types.ModuleType):
    def __init__(self):
        super().__init__("numpy_proxy")
        self.bool = bool

    def __getattr__(self, name):
        return getattr(np, name)

    @staticmethod
    def in1d(a, b):
        return np.isin(a, b).ravel()


def _common_shims(src):
    src = re.sub(r"(\w+)\.has_key\((tuple\(\w+\))\)", r"(\2 in \1)", src)
    assert ".has_key(" not in src
    src = src.replace("import doctest\ndoctest.testmod()", "")
    return src


def load_ref_denoising():
    src = open(os.path.join(REF, "utils", "denoising.py")).read()
    src = _common_shims(src)
    assert src.count("3**3/2") == 1 and src.count("(D_cube / 2)") == 1
    src = src.replace("3**3/2", "3**3//2").replace("(D_cube / 2)", "(D_cube // 2)")
    src = src.replace("import scipy.ndimage.measurements as measure", "measure = ndim")
    mod = types.ModuleType("ref_denoising")
    exec(compile(src, os.path.join(REF, "utils", "denoising.py"), "exec"), mod.__dict__)
    mod.np = _NpProxy()
    return mod


def load_ref_filter_voxels():
    src = open(os.path.join(REF, "utils", "sparseCubes.py")).read()
    a = src.index("def filter_voxels(")
    b = src.index("\ndef ", a + 10)
    ns = {"np": np}
    exec(compile(src[a:b], os.path.join(REF, "utils", "sparseCubes.py"), "exec"), ns)
    return ns["filter_voxels"]


def load_ref_adapthresh(denoising_mod, sparse_stub):
    src = open(os.path.join(REF, "utils", "adapthresh.py")).read()
    src = _common_shims(src)
    src = src.replace("import sparseCubes\n", "").replace("import denoising\n", "")
    assert src.count("D_mid = D_cube / 2") == 1
    src = src.replace("D_mid = D_cube / 2", "D_mid = D_cube // 2")
    src, n = re.subn(r"print ('updated iteration[^\n]*)", r"print(\1)", src)
    assert n == 1
    mod = types.ModuleType("ref_adapthresh")
    mod.__dict__.update(sparseCubes=sparse_stub, denoising=denoising_mod)
    exec(compile(src, os.path.join(REF, "utils", "adapthresh.py"), "exec"), mod.__dict__)
    return mod


def cat(lst, dtype):
    return np.concatenate([np.asarray(x).astype(dtype) for x in lst]) if lst else np.zeros(0, dtype)


def main():
    from tests import util
    out = {}
    den = load_ref_denoising()
    ref_filter = load_ref_filter_voxels()

    # ---- the reference doctest inputs (denoising.py:26-38, 82-94, 160-172), outputs of the code itself -------------------
    ijk_l = [np.array([[1, 0, 0], [2, 2, 2], [3, 3, 3], [1, 0, 1], [2, 3, 3], [0, 3, 3], [1, 2, 2]]),
             np.array([[0, 2, 3], [0, 1, 0], [0, 0, 0], [0, 3, 3]]), np.array([[0, 2, 3], [0, 1, 0], [0, 2, 3]]),
             np.array([[0, 2, 3], [0, 1, 3], [0, 0, 0], [0, 3, 3], [3, 3, 3]], dtype=np.uint8)]
    mask_l = [np.array([1, 0, 1, 1, 1, 1, 1], dtype=bool), np.array([1, 1, 0, 1], dtype=bool), np.array([0, 0, 0], dtype=bool),
              np.array([1, 1, 1, 1, 1], dtype=bool)]
    for nd in (1, 2, 3):
        lab, nlab = den.__cluster_inCube__(ijk_l, mask_l, neighbor_dist=nd)
        out["doc_cluster_nd%d_labels" % nd] = cat(lab, np.int64)
        out["doc_cluster_nd%d_n" % nd] = np.asarray(nlab, np.int64)

    max_count = [0]
    for name, case in util.post_cases().items():
        sc = case["scene"]
        D = sc["D"]
        C = len(sc["ijk_list"])
        sizes = np.array([x.shape[0] for x in sc["ijk_list"]], np.int64)
        # -- denoise_crossCubes on a fixed-threshold mask (main_reconstruct.py:169-173), D_cube as passed there
        mask0 = ref_filter(vxl_mask_list=[], prediction_list=sc["pred_list"], prob_thresh=0.7,
                           rayPooling_votes_list=sc["votes_list"], rayPool_thresh=case["rp"])
        out["post_%s_mask_tau" % name] = cat(mask0, np.uint8)
        for nd in (1, 3):
            ovl, lab = den.__mark_overlappingLabels__(sc["cube_ijk"], sc["ijk_list"], mask0, D_cube=D, neighbor_dist=nd)
            out["post_%s_labels_nd%d" % (name, nd)] = cat(lab, np.int64)
            ovl_mask = [np.isin(l, o) for l, o in zip(lab, ovl)]
            out["post_%s_ovl_nd%d" % (name, nd)] = cat(ovl_mask, np.uint8)
        keep = den.denoise_crossCubes(sc["cube_ijk"], sc["ijk_list"], vxl_mask_list=mask0, D_cube=D)
        out["post_%s_denoised_tau" % name] = cat(keep, np.uint8)
        print(name, "cubes", C, "voxels", int(sizes.sum()), "masked", int(out["post_%s_mask_tau" % name].sum()),
              "kept after denoise", int(out["post_%s_denoised_tau" % name].sum()))

        # -- adapthresh(...) with the I/O stubbed
        rec = dict(ply={}, thresh=[])

        def filter_spy(vxl_mask_list=[], prediction_list=None, prob_thresh=None, rayPooling_votes_list=None, rayPool_thresh=None):
            if isinstance(prob_thresh, list) and len(prob_thresh) == C and rayPooling_votes_list is None:
                rec["thresh"].append(np.asarray(prob_thresh, np.float64))
            return ref_filter(vxl_mask_list=vxl_mask_list, prediction_list=prediction_list, prob_thresh=prob_thresh,
                              rayPooling_votes_list=rayPooling_votes_list, rayPool_thresh=rayPool_thresh)

        def ply_spy(vxl_mask_list, vxl_ijk_list, rgb_list, param, ply_filePath, normal_list=None):
            rec["ply"][os.path.basename(ply_filePath)] = cat(vxl_mask_list, np.uint8)
            return 1

        stub = types.SimpleNamespace(
            load_sparseCubes=lambda f: ([p.copy() for p in sc["pred_list"]], [r.copy() for r in sc["rgb_list"]],
                                        [i.copy() for i in sc["ijk_list"]], [v.copy() for v in sc["votes_list"]],
                                        sc["cube_ijk"].copy(), sc["param"].copy(), None),
            filter_voxels=filter_spy, save_sparseCubes_2ply=ply_spy)
        ada = load_ref_adapthresh(den, stub)
        real_andxor = ada.sparseOccupancy_AND_XOR

        def andxor_spy(a, b):
            r = real_andxor(a, b)
            max_count[0] = max(max_count[0], a.shape[0], b.shape[0], r[1])
            return r
        ada.sparseOccupancy_AND_XOR = andxor_spy
        with tempfile.TemporaryDirectory() as tmp:
            ada.adapthresh(save_result_fld=tmp, N_refine_iter=case["iters"], D_cube=D, init_probThresh=case["init"],
                           min_probThresh=case["init"], max_probThresh=case["maxp"], rayPool_thresh=case["rp"], beta=case["beta"],
                           gamma=0.8, npz_file="unused", RGB_visual_ply=False)
        assert len(rec["thresh"]) == case["iters"]
        out["post_%s_ada_init_denoised" % name] = rec["ply"]["initialization.ply"]
        out["post_%s_ada_thresh" % name] = np.stack(rec["thresh"])
        out["post_%s_ada_denoised" % name] = np.stack([rec["ply"]["iter%d.ply" % i] for i in range(case["iters"])])
        print("   adapthresh thresholds after the last iteration:", np.round(rec["thresh"][-1], 2),
              "kept", out["post_%s_ada_denoised" % name].sum(axis=1))
    assert max_count[0] <= 2048, max_count
    print("largest per-half-cube count:", max_count[0])
    np.savez_compressed(os.path.join(HERE, "postprocess_golden.npz"), **out)
    print("wrote postprocess_golden.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
