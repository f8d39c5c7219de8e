"""CPU-only tests of the host logic: the C ABI library loads and exports every symbol the header
declares (no compute calls), the parameter-list layout, cube sharding (gloo, world_size 2), and that
the product path refuses to run without a GPU instead of falling back."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from tests import util

REPO = util.REPO


def _declared_symbols():
    src = open(os.path.join(REPO, "include", "surfacenet_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sn_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import ctypes
    lib_path = os.path.join(REPO, "surfacenet_b200", "libsurfacenet_b200.so")
    if not os.path.exists(lib_path):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(lib_path)
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libsurfacenet_b200.so does not export " + n
    from surfacenet_b200 import _lib
    assert sorted(_lib.SIGNATURES) == names, "ctypes binding and header disagree"
    assert _lib.lib.sn_version() >= 100
    assert _lib.lib.sn_raypool_workspace_bytes(2, 5, 64) > 0         # pure host arithmetic, no GPU needed
    assert _lib.lib.sn_raypool_workspace_bytes(-1, 5, 64) == -1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from surfacenet_b200 import CVC, SurfaceNet, rayPooling, weights
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        SurfaceNet.SurfaceNet_inference(1, weights.synthetic_params(0))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        rayPooling.rayPooling_1cube_numpy(util.dtu_cameras(), None, np.ones((4, 4, 4), np.float16), np.array([[0, 1]]),
                                          np.zeros(3, np.float32), np.float32(0.4), 0.46)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        CVC.gen_coloredCubes(np.array([[[0, 1]]]), np.zeros((1, 3), np.float32), np.ones(1, np.float32), util.dtu_cameras(),
                             util.image_list(49, [0, 1]), 8)


def test_product_never_imports_oracle():
    for root, _, files in os.walk(os.path.join(REPO, "surfacenet_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f + " imports the oracle"


def test_parameter_layout():
    from surfacenet_b200 import weights
    shapes = weights.expected_shapes()
    assert len(shapes) == 105
    idx = weights.unit_index()
    assert idx["up2"] == 40 and idx["up3"] == 61 and idx["up4"] == 82 and idx["merge_conv"] == 83 and idx["fc1_W"] == 98   # App. B
    assert shapes[idx["conv4_1"]] == (160, 300, 3, 3, 3) and shapes[idx["side_op4"]] == (300, 16, 1, 1, 1)               # (C_in, C_out, ...)
    assert sum(int(np.prod(s)) for i, s in enumerate(shapes) if len(s) == 5 and s[0] != 1 or i == idx["merge_conv3"]) == 8811252
    p = weights.synthetic_params(0)
    assert np.array_equal(p[40], weights.upsample_W(3)) and np.allclose(weights.upsample_W(5)[0, 0, :, 2, 2], [1 / 3, 2 / 3, 1, 2 / 3, 1 / 3])
    with pytest.raises(ValueError):
        weights.validate(p[:-1])
    bad = list(p); bad[0] = bad[0][:, :5]
    with pytest.raises(ValueError):
        weights.validate(bad)
    p2 = weights.synthetic_params(0)
    assert all(np.array_equal(a, b) for a, b in zip(p, p2))


def test_model_file_roundtrip(tmp_path):
    import pickle
    from surfacenet_b200 import weights
    p = weights.synthetic_params(0, calibrated=False)
    f = tmp_path / "x.model"
    with open(f, "wb") as fh:
        pickle.dump(p, fh, protocol=2)                    # the reference writes py2 pickles (SurfaceNet.py:397-399)
    q = weights.load_model_file(str(f))
    assert all(np.array_equal(a, b) for a, b in zip(p, q))
    np.savez(tmp_path / "x.npz", **{str(i): a for i, a in enumerate(p)})
    q = weights.load_model_file(str(tmp_path / "x.npz"))
    assert all(np.array_equal(a, b) for a, b in zip(p, q))


def test_shard_bounds():
    from surfacenet_b200 import pipeline
    assert pipeline.shard_bounds(512, 8) == (64, [(64 * r, 64 * r + 64) for r in range(8)])
    per, b = pipeline.shard_bounds(10, 4)
    assert per == 3 and b == [(0, 3), (3, 6), (6, 9), (9, 10)]
    per, b = pipeline.shard_bounds(2, 4)
    assert per == 1 and b == [(0, 1), (1, 2), (2, 2), (2, 2)]
    assert pipeline.shard_bounds(0, 2) == (0, [(0, 0), (0, 0)])


_WORKER = r"""
import os, sys
sys.path.insert(0, {repo!r})
import numpy as np, torch, torch.distributed as dist
from surfacenet_b200 import pipeline
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=rank, world_size=world)
n = {n}
full_p = torch.arange(n * 8, dtype=torch.float32).reshape(n, 2, 2, 2)
full_v = (torch.arange(n * 8) % 251).to(torch.uint8).reshape(n, 2, 2, 2)
calls = []
def compute(lo, hi):
    calls.append((lo, hi))
    return full_p[lo:hi] * 2, full_v[lo:hi]
p, v = pipeline.infer_sharded(compute, n, [((2, 2, 2), torch.float32), ((2, 2, 2), torch.uint8)], rank, world, device="cpu")
assert torch.equal(p, full_p * 2) and torch.equal(v, full_v), "rank %d reassembly wrong" % rank
per, bounds = pipeline.shard_bounds(n, world)
assert calls == ([bounds[rank]] if bounds[rank][1] > bounds[rank][0] else [])
dist.barrier(); dist.destroy_process_group()
print("OK", rank)
"""


@pytest.mark.parametrize("n", [5, 1])
def test_infer_sharded_gloo_world2(tmp_path, n):
    """cube-sharded batch + one all-gather per output, world_size 2 over gloo (ragged and nearly empty shards)."""
    port = 29600 + os.getpid() % 200 + n
    script = tmp_path / "w.py"
    script.write_text(_WORKER.format(repo=REPO, port=port, n=n))
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and "OK %d" % r in o, o[-2000:]


def test_bench_reference_arm_json():
    """--impl reference prints one JSON line with the contract keys (tiny workload for speed)."""
    import json
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--workload", "c2"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["unit"] == "voxels/s"
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["gpu_launches"] == 0


def _ref_sparse_funcs():
    """filter_voxels / save_sparseCubes / load_sparseCubes executed from the reference source when it is present (build
    container only); mechanical py3 fix: np.load needs a binary file handle (`open(filePath)` -> `open(filePath, 'rb')`)."""
    ref = "/root/reference/utils/sparseCubes.py"
    if not os.path.exists(ref):
        pytest.skip("reference tree not present")
    src = open(ref).read()
    a, b, c = src.index("def filter_voxels("), src.index("def save2ply("), src.index("def save_sparseCubes(")
    d = src.index("def __debug():")
    body = (src[a:b] + src[c:d]).replace("with open(filePath) as f:", "with open(filePath, 'rb') as f:")
    ns = {"np": np}
    exec(compile(body, ref, "exec"), ns)
    return ns


def test_sparse_host_consumers_match_reference(tmp_path, capsys):
    from surfacenet_b200 import sparseCubes
    ns = _ref_sparse_funcs()
    rs = np.random.RandomState(3)
    counts = [5, 0, 17, 3]
    pl = [rs.rand(n).astype(np.float16) for n in counts]
    rl = [rs.randint(0, 256, (n, 3)).astype(np.uint8) for n in counts]
    il = [rs.randint(0, 52, (n, 3)).astype(np.uint8) for n in counts]
    vl = [rs.randint(0, 11, n).astype(np.uint8) for n in counts]
    cube_ijk = rs.randint(0, 40, (4, 3)).astype(np.uint32)
    param = np.zeros(4, util.PARAM_DTYPE); param["xyz"] = rs.rand(4, 3); param["resol"] = 0.4
    vp = rs.randint(0, 49, (4, 5, 2)).astype(np.uint16)
    # masks
    m_ref = ns["filter_voxels"]([], pl, 0.7, vl, 8.0)
    m_new = sparseCubes.filter_voxels([], pl, 0.7, vl, 8.0)
    assert all(np.array_equal(a, b) for a, b in zip(m_ref, m_new)) and len(m_new) == 4
    m_new2 = sparseCubes.filter_voxels([m.copy() for m in m_new], prediction_list=pl, prob_thresh=[0.9, 0.1, 0.5, 0.2])
    m_ref2 = ns["filter_voxels"]([m.copy() for m in m_ref], prediction_list=pl, prob_thresh=[0.9, 0.1, 0.5, 0.2])
    assert all(np.array_equal(a, b) for a, b in zip(m_ref2, m_new2))
    with pytest.raises(Warning):
        sparseCubes.filter_voxels([], pl, None)
    # NPZ schema: written by us -> read by the reference, and the other way round
    f1, f2 = str(tmp_path / "a.npz"), str(tmp_path / "b.npz")
    sparseCubes.save_sparseCubes(f1, pl, rl, il, vl, cube_ijk, param, vp)
    ns["save_sparseCubes"](f2, pl, rl, il, vl, cube_ijk, param, vp)
    z1, z2 = np.load(f1), np.load(f2)
    assert sorted(z1.files) == sorted(z2.files) and all(np.array_equal(z1[k], z2[k]) and z1[k].dtype == z2[k].dtype for k in z1.files)
    back_ref = ns["load_sparseCubes"](f1)
    back_new = sparseCubes.load_sparseCubes(f2)
    for a, b in zip(back_ref, back_new):
        if isinstance(a, list):
            assert all(np.array_equal(x, y) for x, y in zip(a, b))
        else:
            assert np.array_equal(a, b)
    assert all(np.array_equal(x, y) for x, y in zip(back_new[0], pl))


def test_lists_from_flat():
    from surfacenet_b200 import sparseCubes
    counts = np.array([2, 0, 3], np.int32)
    sp = dict(counts=counts, offsets=np.array([0, 2, 2, 5], np.int32), pred=np.arange(5).astype(np.float16), rgb=np.zeros((5, 3), np.uint8),
              ijk=np.arange(15).reshape(5, 3).astype(np.uint8), votes=np.arange(5).astype(np.uint8))
    param = np.zeros(3, util.PARAM_DTYPE); param["xyz"] = 10.0; param["resol"] = 0.5; param["ijk"] = np.arange(9).reshape(3, 3)
    vp = np.arange(12).reshape(3, 2, 2)
    pl, rl, il, vl, cube_ijk, param_np, vp_np = sparseCubes.lists_from_flat(sp, param, vp, 52, 64)
    assert [len(x) for x in pl] == [2, 3] and np.array_equal(il[1], sp["ijk"][2:5]) and np.array_equal(vl[1], [2, 3, 4])
    assert np.array_equal(cube_ijk, param["ijk"][[0, 2]]) and np.allclose(param_np["xyz"], 10.0 + 0.5 * 6) and vp_np.dtype == np.uint16


def test_initialize_cubes_matches_reference_log():
    """utils/scene.py:43-58 on DTU scan9 (params.py:161-172): the shipped job log reports 24,420 cubes at s=64
    (q.log/inference.0000022:116); s=32 gives 40 x 74 x 66."""
    from surfacenet_b200 import reconstruct
    BB = np.array([[-73, 129], [-197, 183], [472, 810]])
    c, side = reconstruct.initialize_cubes(np.float32(0.4), 64, 52, 0.5, BB)
    assert len(c) == 24420 and np.array_equal(c["ijk"].max(0) + 1, [20, 37, 33]) and abs(side - 25.6) < 1e-5
    assert np.allclose(c["xyz"][0], BB[:, 0] - (64 - 52) * 0.4 / 2, atol=1e-5) and np.array_equal(c["ijk"][1], [0, 0, 1])
    assert np.allclose(c["xyz"][1] - c["xyz"][0], [0, 0, 52 * 0.4 * 0.5], atol=1e-5)
    c, _ = reconstruct.initialize_cubes(np.float32(0.4), 32, 26, 0.5, BB)
    assert len(c) == 40 * 74 * 66


def test_save2ply_layout(tmp_path):
    """utils/sparseCubes.py:246-327: the PLY the N4 consumers write (binary little endian, x y z red green blue)."""
    from surfacenet_b200 import sparseCubes
    rs = np.random.RandomState(3)
    ijk = [rs.randint(0, 52, size=(5, 3)).astype(np.uint8), rs.randint(0, 52, size=(3, 3)).astype(np.uint8)]
    rgb = [rs.randint(0, 256, size=(5, 3)).astype(np.uint8), rs.randint(0, 256, size=(3, 3)).astype(np.uint8)]
    mask = [np.array([1, 0, 1, 1, 0], bool), np.array([0, 1, 0], bool)]
    param = np.zeros(2, util.PARAM_DTYPE)
    param["xyz"] = [[1.5, -2.0, 600.0], [10.0, 0.0, 610.25]]
    param["resol"] = np.float32(0.4)
    path = str(tmp_path / "sub" / "x.ply")
    assert sparseCubes.save_sparseCubes_2ply(mask, ijk, rgb, param, path) == 1
    xyz, col = util.read_ply(path)
    exp = np.vstack([ijk[c][mask[c]] * param[c]["resol"] + param[c]["xyz"][None, :] for c in range(2)]).astype(np.float32)
    assert xyz.dtype == np.float32 and np.array_equal(xyz, exp)
    assert np.array_equal(col, np.vstack(rgb)[np.concatenate(mask)])
    with pytest.raises(Warning):
        sparseCubes.save_sparseCubes_2ply(mask, ijk, rgb[:1], param, path)


def test_post_workspace_and_symbols():
    from surfacenet_b200 import _lib
    assert _lib.lib.sn_sparse_post_workspace_bytes(24420, 50000000, 52) > 24420 * (52 ** 3 // 32) * 8
    assert _lib.lib.sn_sparse_post_workspace_bytes(4, 100, 0) == -1 and _lib.lib.sn_sparse_post_workspace_bytes(4, 100, 257) == -1
    assert _lib.lib.sn_sparse_post_workspace_bytes(0, 0, 1) > 0


def test_bench_simnet_reference_arm_json():
    """bench.py --workload simnet --impl reference: one JSON line with the contract keys (torch-CPU VGG-16 embedding, no GPU)."""
    import json, subprocess, sys
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--workload", "simnet", "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "patches/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0 and line["gpu_launches"] == 0


def test_batching_helpers_reference_doctest_known_answers():
    """utils/utils.py:45-230: the reference's doctest answers for the host-side batching helpers."""
    from surfacenet_b200 import utils as U
    assert U.gen_batch_index(6, 3) == [[0, 1, 2], [3, 4, 5]] and U.gen_batch_index(7, 3) == [[0, 1, 2], [3, 4, 5], [6]]
    assert U.gen_batch_index(8, 3) == [[0, 1, 2], [3, 4, 5], [6, 7]]
    assert U.gen_batch_npBool(6, 3).tolist() == [[True] * 3 + [False] * 3, [False] * 3 + [True] * 3]
    assert U.gen_batch_npBool(6, 100).tolist() == [[True] * 6]
    sel = U.gen_batch_npBool(7, 3)
    assert sel.tolist() == [[True, True, True, False, False, False, False], [False, False, False, True, True, True, False],
                            [False, False, False, False, False, False, True]]
    assert np.arange(14).reshape((7, 2))[sel[2]].tolist() == [[12, 13]]
    ind = np.array([0, 1, 1, 1, 0, 0, 1, 0, 1, 1, 1], dtype=bool)
    s = U.gen_non0Batch_npBool(ind, 3)
    assert s.shape == (3, 11)
    assert s[0].tolist() == [bool(x) for x in [0, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0]] and s[1].tolist() == [bool(x) for x in [0, 0, 0, 0, 0, 0, 1, 0, 1, 1, 0]]
    assert s[2].tolist() == [bool(x) for x in [0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1]]
    feat = np.arange(3 * 6 * 8).reshape((3, 6, 8))
    b1 = [feat[i, j] for i, j in U.yield_batch_ij_npBool((range(3), range(6)), 5)]
    b2 = [feat.reshape((18, 8))[b] for b in U.gen_batch_npBool(18, 5)]
    assert len(b1) == len(b2) == 4 and all(np.array_equal(x, y) for x, y in zip(b1, b2))
    assert [b.sum() for b in U.yield_batch_npBool(7, 3)] == [3, 3, 1]
    assert U.k_combination_np([2, 5, 8]).tolist() == [[2, 5], [2, 8], [5, 8]] and U.k_combination_np([2, 5, 8]).dtype == np.int64
    assert np.allclose(U.k_combination_np([2.2, 5.5, 8.8, 9.9], k=3), [[2.2, 5.5, 8.8], [2.2, 5.5, 9.9], [2.2, 8.8, 9.9], [5.5, 8.8, 9.9]])


def test_scene_and_camera_host_reference_doctest_known_answers():
    """utils/scene.py:24-40 (initializeCubes) and utils/camera.py:88-96 (__cameraP2T__) known answers."""
    from surfacenet_b200 import camera, reconstruct
    cubes, side = reconstruct.initialize_cubes(resol=1, cube_D=22, cube_Dcenter=10, cube_overlapping_ratio=0.5,
                                               BB=np.array([[3, 88], [-11, 99], [-110, -11]]))
    assert side == 22
    # the doctest at scene.py:28-36 prints the first cube at BB_min; the CODE below it (scene.py:43-58, the version that produced the
    # shipped log's 24,420 cubes) starts at BB_min - (cube_D - cube_Dcenter) * resol / 2 = BB_min - 6: the code is followed
    assert cubes["xyz"][:3].tolist() == [[-3., -17., -116.], [-3., -17., -111.], [-3., -17., -106.]] and cubes["ijk"][:3].tolist() == [[0, 0, 0], [0, 0, 1], [0, 0, 2]]
    n_z = int(np.ceil(((-11 + 6) - (-110 - 6)) / 5.0))
    assert cubes["ijk"][n_z].tolist() == [0, 1, 0] and cubes["xyz"][n_z].tolist() == [-3., -12., -116.]
    assert np.all(cubes["resol"] == 1.0)
    P = np.array([[798.693916, -2438.153488, 1568.674338, -542599.034996], [-44.838945, 1433.912029, 2576.399630, -1176685.647358],
                  [-0.840873, -0.344537, 0.417405, 382.793511]])
    assert np.allclose(camera.__cameraP2T__(P), [555.64348632032, 191.10837560939, 360.02470478273])
    assert np.allclose(camera.cameraPs2Ts(np.stack([P, P]))[1], camera.__cameraP2T__(P)) and isinstance(camera.cameraPs2Ts([P]), list)


def test_new_dropins_have_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from surfacenet_b200 import adapthresh, camera, denoising, earlyRejection, similarityNet, viewPairSelection
    ijk, mask = [np.array([[1, 0, 0]], np.uint8)], [np.ones(1, bool)]
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        denoising.denoise_crossCubes(np.zeros((1, 3), np.int32), ijk, mask, 4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        adapthresh.adapthresh_lists([np.ones(1, np.float16)], ijk, [np.ones(1, np.uint8)], np.zeros((1, 3), np.int32), 1, 4, 0.5, 0.9, 0, 6)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        similarityNet.similarityNet_inference(similarityNet.synthetic_params(0), (64, 64))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        earlyRejection.selectFromSimilarity(np.zeros((2, 3), np.float32), 1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        viewPairSelection.__argmaxN_viewPairs__(np.array([[0, 1]]), np.ones((1, 1)), 1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        camera.viewPairAngles_wrt_pts(np.zeros((2, 3)), np.ones((1, 3)))


def test_similarityNet_model_file_roundtrip(tmp_path):
    """nets/similarityNet.py:241-244: the model file is a pickled list of the 30 parameter arrays (python-2 protocol)."""
    import pickle
    from surfacenet_b200 import similarityNet
    params = similarityNet.synthetic_params(1)
    assert len(params) == 30 and [p.shape for p in params] == [tuple(s) for s in similarityNet.PARAM_SHAPES]
    f = tmp_path / "epoch33.model"
    with open(f, "wb") as fh:
        pickle.dump([np.asarray(p, np.float64) for p in params[:2]] + params[2:], fh, protocol=2)
    back = similarityNet.load_model_file(str(f))
    assert all(b.dtype == np.float32 and np.array_equal(a, b) for a, b in zip(params, back))


def test_bench_post_reference_arm_json():
    """bench.py --workload post --impl reference: the numpy/scipy post-processing restatement, one JSON line."""
    import json, subprocess, sys
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--workload", "post", "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "voxels/s" and line["value"] > 0 and line["cpu_baseline"]["cores"] == 1


def test_cube_grid_builders_match_reference_outputs():
    """utils/scene.py:7-107 executed by tests/golden/make_golden_scene.py: the scan9 grid (24,420 cubes, q.log) and quantizePts2Cubes."""
    from surfacenet_b200 import reconstruct
    g = np.load(os.path.join(REPO, "tests", "golden", "scene_golden.npz"))
    cubes, side = reconstruct.initialize_cubes(np.float32(0.4), 64, 52, 0.5, np.array([[-73., 129.], [-197., 183.], [472., 810.]]))
    assert len(cubes) == int(g["init_n"][0]) == 24420
    assert np.array_equal(cubes["xyz"][::997], g["init_xyz_sample"]) and np.array_equal(cubes["ijk"][::997], g["init_ijk_sample"])
    q, D_mm = reconstruct.quantize_pts_to_cubes(g["q_pts"], np.float32(0.4), 32, 26, 0.5, BB=np.array([[0., 40.], [-5., 20.], [598., 612.]]))
    assert D_mm == g["q_D_mm"][0] and np.array_equal(q["ijk"], g["q_ijk"]) and np.array_equal(q["xyz"], g["q_xyz"]) and np.array_equal(q["resol"], g["q_resol"])


def test_camera_and_image_file_readers(tmp_path):
    """utils/camera.py:7-82, utils/image.py:50-89: the readers against matrices the reference readers produced from the bundled DTU / Middlebury
    calibration files (the file texts are stored in the fixture, generated in the build container)."""
    from PIL import Image
    from surfacenet_b200 import camera, image
    g = np.load(os.path.join(REPO, "tests", "golden", "camera_files_golden.npz"))
    os.makedirs(tmp_path / "cal18")
    for v, txt in zip((1, 10, 49), g["dtu_text"]):
        (tmp_path / "cal18" / ("pos_%03d.txt" % v)).write_text(str(txt))
    P = camera.readCameraPOs_as_np(str(tmp_path), "DTU", "cal18/pos_#.txt", 9, [1, 10, 49])
    assert P.dtype == np.float64 and np.array_equal(P, g["P_dtu"])
    (tmp_path / "dinoSR_par.txt").write_text(str(g["mid_par_text"]))
    Pm = camera.readCameraPOs_as_np(str(tmp_path), "Middlebury", "dinoSR_par.txt", "dinoSparseRing", list(range(7, 13)))
    assert np.array_equal(Pm, g["P_mid"])
    img = util.synth_image(1, 30, 40)
    Image.fromarray(img).save(str(tmp_path / "rect_007_3.png"))
    back = image.readImages(str(tmp_path), "rect_#_3.png", [7, 7], return_list=False)
    assert back.shape == (2, 30, 40, 3) and back.dtype == np.uint8 and np.array_equal(back[0], img)


def test_main_reconstruct_dropin_reads_inputs_then_needs_gpu(tmp_path):
    """surfacenet_b200.main_reconstruct.reconstruction: the reference's argument list; without a GPU it gets through the file readers and
    then refuses to run (no CPU fallback)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from PIL import Image
    from surfacenet_b200 import main_reconstruct, similarityNet, weights
    cams = util.dtu_cameras()
    os.makedirs(tmp_path / "cal")
    for v in (1, 2):
        np.savetxt(str(tmp_path / "cal" / ("pos_%03d.txt" % v)), cams[v - 1], delimiter=' ')
        Image.fromarray(util.synth_image(v, 120, 160)).save(str(tmp_path / ("rect_%03d.png" % v)))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        main_reconstruct.reconstruction(str(tmp_path), 9, "rect_#.png", "cal/pos_#.txt", None, str(tmp_path / "out"), 1, np.float32(0.4),
                                        np.array([[0., 10.], [0., 10.], [600., 610.]]), [1, 2], surfacenet_model=weights.synthetic_params(0),
                                        similnet_model=similarityNet.synthetic_params(0), cube_D=32)
    with pytest.raises(FileNotFoundError):                       # an initial point cloud (main_reconstruct.py:57) is read before anything else
        main_reconstruct.reconstruction(str(tmp_path), 9, "rect_#.png", "cal/pos_#.txt", "pts.ply", str(tmp_path / "out"), 1, np.float32(0.4),
                                        np.zeros((3, 2)), [1, 2])


_WORKER_RECON = """
import os, sys
sys.path.insert(0, {repo!r})
import numpy as np, torch.distributed as dist
from surfacenet_b200 import reconstruct
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=rank, world_size=world)

class FakeHot:                       # stands in for pipeline.HotPath.infer_batch_sparse (no GPU): deterministic per cube
    def infer_batch_sparse(self, pairs, xyz, resol, w, D, Dc, thresh):
        counts = np.array([int(abs(x[0])) % 4 for x in xyz], np.int32)          # some cubes are empty
        off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
        T = int(off[-1])
        seed = np.repeat(np.asarray(xyz)[:, 0], counts).astype(np.float32)
        return dict(counts=counts, offsets=off, ijk=(np.arange(T * 3).reshape(T, 3) % 7).astype(np.uint8), pred=(seed / 100).astype(np.float16),
                    rgb=np.full((T, 3), 9, np.uint8), votes=(seed.astype(np.int64) % 5).astype(np.uint8))

n = {n}
cubes = np.zeros(n, dtype=reconstruct.PARAM_DTYPE)
cubes["xyz"] = np.arange(n * 3, dtype=np.float32).reshape(n, 3); cubes["ijk"] = np.arange(n * 3).reshape(n, 3); cubes["resol"] = 0.4
pairs = np.arange(n * 4).reshape(n, 2, 2); w = np.ones((n, 2), np.float32)
whole = reconstruct.reconstruct_cubes(FakeHot(), cubes, pairs, w, 32, 26, batch_size=2)
got = reconstruct.reconstruct_cubes(FakeHot(), cubes, pairs, w, 32, 26, batch_size=2, rank=rank, world_size=world, gather=True)
part = reconstruct.reconstruct_cubes(FakeHot(), cubes, pairs, w, 32, 26, batch_size=2, rank=rank, world_size=world)
assert len(part[0]) < len(whole[0])                                        # a rank alone holds only its share
for a, b in zip(whole[:4], got[:4]):
    assert len(a) == len(b) and all(np.array_equal(x, y) for x, y in zip(a, b)), "rank %d: gathered lists differ" % rank
for a, b in zip(whole[4:], got[4:]):
    assert np.array_equal(a, b), "rank %d: gathered cube tables differ" % rank
dist.barrier(); dist.destroy_process_group()
print("OK", rank)
"""


def test_reconstruct_cubes_gather_gloo_world2(tmp_path):
    """world_size 2: every rank ends up with the WHOLE scene in single-rank batch order before the cross-cube stages
    (denoise_crossCubes / NPZ need every neighbour cube; main_reconstruct.py:168-183)."""
    port = 29850 + os.getpid() % 100
    script = tmp_path / "w.py"
    script.write_text(_WORKER_RECON.format(repo=REPO, port=port, n=9))
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and "OK %d" % r in o, o[-2000:]


def test_initialize_cubes_float32_resol_uses_float64_arithmetic():
    """params.py passes resol = np.float32(0.4); under the reference's numpy 1.x, float32_scalar * int is float64 arithmetic."""
    from surfacenet_b200 import reconstruct
    BB = [[-40, 40], [80, 160], [630, 680]]
    cubes32, side32 = reconstruct.initialize_cubes(np.float32(0.4), 64, 52, 0.5, BB)
    r = float(np.float32(0.4))
    stride, margin = r * 52 * 0.5, (r * 64 - r * 52) / 2
    n_axis = [int(np.ceil(((b[1] + margin) - (b[0] - margin)) / stride)) for b in BB]
    ijk = np.indices(tuple(n_axis)).reshape(3, -1).T
    want = (ijk * stride + (np.array(BB, np.float64)[:, 0][None] - margin)).astype(np.float32)
    assert cubes32.shape[0] == ijk.shape[0] and np.array_equal(cubes32["xyz"], want) and side32 == r * 64
    pts = np.random.default_rng(0).uniform(-30, 30, (50, 3))
    q32, _ = reconstruct.quantize_pts_to_cubes(pts, np.float32(0.4), 64, 52, 0.5)
    q64, _ = reconstruct.quantize_pts_to_cubes(pts, r, 64, 52, 0.5)
    assert np.array_equal(q32["xyz"], q64["xyz"]) and np.array_equal(q32["ijk"], q64["ijk"])


def test_readPointCloud_xyz_formats(tmp_path):
    """utils/scene.py:111-114 without plyfile: the x / y / z columns of the vertex element, binary (both byte orders, extra properties
    such as normals / colours in between) and ascii; the file's dtype is kept, as np.c_ of the plyfile columns does."""
    from surfacenet_b200 import sparseCubes
    rs = np.random.RandomState(2)
    xyz = (rs.rand(37, 3) * 100 - 50).astype(np.float32)
    rgb = rs.randint(0, 256, size=(37, 3)).astype(np.uint8)
    nrm = rs.randn(37, 3).astype(np.float32)
    p = str(tmp_path / "a.ply")
    sparseCubes.save2ply(p, xyz, rgb_np=rgb, normal_np=nrm)                               # x y z nx ny nz red green blue, little endian
    got = sparseCubes.readPointCloud_xyz(p)
    assert got.dtype == np.float32 and np.array_equal(got, xyz)
    # big endian, double coordinates, a face element after the vertices
    dt = np.dtype([("x", ">f8"), ("red", "u1"), ("y", ">f8"), ("z", ">f8")])
    v = np.zeros(5, dt); v["x"], v["y"], v["z"], v["red"] = xyz[:5, 0], xyz[:5, 1], xyz[:5, 2], 7
    hdr = ("ply\nformat binary_big_endian 1.0\ncomment made by hand\nelement vertex 5\nproperty double x\nproperty uchar red\n"
           "property double y\nproperty double z\nelement face 1\nproperty list uchar int vertex_indices\nend_header\n")
    with open(str(tmp_path / "b.ply"), "wb") as f:
        f.write(hdr.encode("ascii")); f.write(v.tobytes()); f.write(bytes([3]) + np.array([0, 1, 2], ">i4").tobytes())
    got = sparseCubes.readPointCloud_xyz(str(tmp_path / "b.ply"))
    assert got.dtype == np.float64 and np.array_equal(got, xyz[:5].astype(np.float64))
    with open(str(tmp_path / "c.ply"), "w") as f:
        f.write("ply\nformat ascii 1.0\nelement vertex 3\nproperty float x\nproperty float y\nproperty float z\nproperty uchar red\nend_header\n")
        for r in xyz[:3]:
            f.write("%r %r %r 9\n" % (float(r[0]), float(r[1]), float(r[2])))
    got = sparseCubes.readPointCloud_xyz(str(tmp_path / "c.ply"))
    assert got.dtype == np.float32 and np.array_equal(got, xyz[:3])
    with open(str(tmp_path / "d.ply"), "wb") as f:
        f.write(b"not a ply")
    with pytest.raises(ValueError):
        sparseCubes.readPointCloud_xyz(str(tmp_path / "d.ply"))
    # the cubes of a point cloud read back from a file equal the cubes of the array (main_reconstruct.py:57-60)
    from surfacenet_b200 import reconstruct
    pts = np.array([[10.0, -5.0, 620.0], [10.3, -5.2, 620.1], [25.0, 3.0, 640.0]], np.float32)
    sparseCubes.save2ply(str(tmp_path / "pts.ply"), pts)
    a, _ = reconstruct.quantize_pts_to_cubes(sparseCubes.readPointCloud_xyz(str(tmp_path / "pts.ply")), np.float32(0.4), 32, 26, 0.5)
    b, _ = reconstruct.quantize_pts_to_cubes(pts, np.float32(0.4), 32, 26, 0.5)
    assert np.array_equal(a["xyz"], b["xyz"]) and np.array_equal(a["ijk"], b["ijk"])


def test_fast_mode_warns_and_exact_is_the_default():
    """the single-pass fp16 mode misses the 1e-4 bound by ~100x -- selecting it must be loud; no environment variable may
    switch the default of the drop-ins away from the parity mode."""
    import warnings
    from surfacenet_b200 import _lib
    assert _lib.DEFAULT_MODE == "exact"
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter("always")
        assert _lib.resolve_mode("exact") == _lib.MODE_TC_EXACT and _lib.resolve_mode("fp32") == _lib.MODE_FP32
        assert not rec
        assert _lib.resolve_mode("fast") == _lib.MODE_TC_FAST
        assert len(rec) == 1 and issubclass(rec[0].category, RuntimeWarning) and "1e-4" in str(rec[0].message)
    with pytest.raises((KeyError, ValueError)):
        _lib.resolve_mode("bf16")
