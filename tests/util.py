"""Deterministic input builders shared by the golden generator (tests/golden/make_golden.py, run
against the reference in the build container) and by the parity tests (run anywhere)."""
import os
import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(REPO, "tests", "golden", "reference_golden.npz")
MEAN6 = np.asarray([123.68, 116.779, 103.939, 123.68, 116.779, 103.939], dtype=np.float32)   # params.py:129
SCAN9_BB = np.array([[-73, -197, 472], [129, 183, 810]], dtype=np.float64)


def dtu_cameras():
    return np.load(os.path.join(REPO, "surfacenet_b200", "data", "dtu_cal18_cameras.npy"))


def synth_image(seed, H, W):
    """Blocky random colours + per-pixel noise, uint8 (H,W,3).  Bit-stable (legacy RandomState)."""
    rs = np.random.RandomState(1000 + seed)
    coarse = rs.randint(0, 256, size=((H + 15) // 16, (W + 15) // 16, 3)).astype(np.int16)
    img = np.repeat(np.repeat(coarse, 16, axis=0), 16, axis=1)[:H, :W]
    img = img + rs.randint(-20, 21, size=(H, W, 3)).astype(np.int16)
    return np.clip(img, 0, 255).astype(np.uint8)


def image_list(n_views, used, sizes=None):
    """Reference-style ``models_img``: python list indexed by view position (image.py:80-89)."""
    imgs = [None] * n_views
    for v in used:
        H, W = (1200, 1600) if sizes is None or v not in sizes else sizes[v]
        imgs[v] = synth_image(v, H, W)
    return imgs


def cvc_cases(cams):
    c = {}
    f32 = np.float32
    pairs = np.array([[[0, 5], [17, 22]], [[5, 48], [22, 0]], [[17, 5], [48, 22]]], dtype=np.int64)
    xyz = np.array([[20.0, -12.5, 630.0], [-10.3, 30.7, 655.1], [55.25, -60.0, 600.5]], dtype=f32)
    c["basic"] = dict(pairs=pairs, xyz=xyz, resol=np.full(3, 0.4, f32), cameraPOs=cams, D=16,
                      images=image_list(49, [0, 5, 17, 22, 48]))
    # images of different (small) sizes: most voxels of some views fall outside -> zeros (CVC.py:42-46)
    c["ragged_sizes"] = dict(pairs=pairs[:2], xyz=xyz[:2], resol=np.array([0.4, 0.8], f32), cameraPOs=cams, D=16,
                             images=image_list(49, [0, 5, 17, 22, 48], sizes={5: (640, 800), 22: (700, 760), 0: (620, 1600)}))
    # duplicate views inside and across pairs (CVC.py:83)
    c["dup_views"] = dict(pairs=np.array([[[3, 3], [3, 7]]], dtype=np.int64), xyz=xyz[:1], resol=np.full(1, 0.4, f32),
                          cameraPOs=cams, D=16, images=image_list(49, [3, 7]))
    # BASELINE config 1 shape: one 32^3 cube, one view pair
    c["c1_s32"] = dict(pairs=np.array([[[10, 11]]], dtype=np.int64), xyz=np.array([[15.0, -20.0, 640.0]], f32),
                       resol=np.full(1, 0.4, f32), cameraPOs=cams, D=32, images=image_list(49, [10, 11]))
    # far outside the frustum / behind the camera: everything out of scope or wrapped (CVC.py:39,45)
    c["outside"] = dict(pairs=np.array([[[0, 1]]], dtype=np.int64), xyz=np.array([[900.0, -900.0, -300.0]], f32),
                        resol=np.full(1, 2.0, f32), cameraPOs=cams, D=8, images=image_list(49, [0, 1]))
    return c


def sheet_prediction(D, phase=0.0, sigma=0.08):
    """Smooth sheet exp(-(x+0.3 sin(6y+phase)-0.5)^2/2 sigma^2), float16."""
    g = (np.arange(D) + 0.5) / D
    x, y, z = np.meshgrid(g, g, g, indexing="ij")
    p = np.exp(-((x + 0.3 * np.sin(6 * y + phase) * (0.5 + 0.5 * z) - 0.5) ** 2) / (2 * sigma ** 2))
    return p.astype(np.float16)


def raypool_cases(cams):
    c = {}
    f32 = np.float32
    xyz = np.array([20.0, -12.5, 630.0], dtype=f32)
    c["sheet16"] = dict(cameraPOs=cams, pred=sheet_prediction(16), pairs=np.array([[0, 5], [17, 22]]), xyz=xyz,
                        resol=f32(0.4), thresh=0.46)
    c["sheet32_dup"] = dict(cameraPOs=cams, pred=sheet_prediction(32, 1.0, 0.15), pairs=np.array([[3, 3], [3, 7], [9, 40]]),
                            xyz=np.array([-10.3, 30.7, 655.1], dtype=f32), resol=f32(0.4), thresh=0.46)
    q = np.round(sheet_prediction(16, 0.5, 0.3).astype(np.float32) * 4) / 4          # values in {0,.25,.5,.75,1}: ties
    c["ties"] = dict(cameraPOs=cams, pred=q.astype(np.float16), pairs=np.array([[1, 30], [12, 44]]), xyz=xyz,
                     resol=f32(0.4), thresh=0.46)
    low = cams.copy(); low[:, :2, :] /= 8.0                                          # 8x coarser pixels: many voxels per
    c["lowres_collide"] = dict(cameraPOs=low, pred=sheet_prediction(16, 2.0, 0.4), pairs=np.array([[0, 5], [17, 22]]),
                               xyz=xyz, resol=f32(0.4), thresh=0.46)               # (pixel, depth-bin) cell: last write wins
    c["all_ones"] = dict(cameraPOs=low, pred=np.ones((8, 8, 8), np.float16), pairs=np.array([[2, 6]]), xyz=xyz,
                         resol=f32(0.8), thresh=0.46)
    c["empty"] = dict(cameraPOs=cams, pred=np.full((8, 8, 8), 0.3, np.float16), pairs=np.array([[2, 6]]), xyz=xyz,
                      resol=f32(0.4), thresh=0.46)
    rs = np.random.RandomState(5)
    c["none_thresh_f32"] = dict(cameraPOs=cams, pred=(rs.rand(8, 8, 8) * 0.9 + 0.05).astype(np.float32),
                                pairs=np.array([[20, 21], [21, 25]]), xyz=xyz, resol=f32(0.4), thresh=None)
    c["exact_thresh"] = dict(cameraPOs=cams, pred=np.where(rs.rand(8, 8, 8) < 0.5, np.float16(0.46), np.float16(0.4602)).astype(np.float16),
                             pairs=np.array([[20, 21]]), xyz=xyz, resol=f32(0.4), thresh=0.46)
    return c


PARAM_DTYPE = np.dtype([("xyz", np.float32, (3,)), ("ijk", np.uint32, (3,)), ("resol", np.float32)])   # utils/scene.py:55


def sparse_cases(cams):
    """Dense batches as main_reconstruct.py:154-162 hands them to sparseCubes.append_dense_2sparseList."""
    c = {}
    rs = np.random.RandomState(21)
    def batch(D, n, n_vp, empty=()):
        pred = np.stack([sheet_prediction(D, 0.7 * i, 0.12).astype(np.float32) for i in range(n)])[:, None]     # (N,1,D,D,D) f32
        for e in empty:
            pred[e] = 0.2
        rgb = rs.randint(0, 256, size=(n, 3, D, D, D)).astype(np.uint8)
        param = np.zeros(n, PARAM_DTYPE)
        param["xyz"] = (np.array([10.0, -30.0, 620.0]) + rs.rand(n, 3) * 30).astype(np.float32)
        param["ijk"] = rs.randint(0, 40, size=(n, 3))
        param["resol"] = np.float32(0.4)
        pairs = rs.randint(0, 49, size=(n, n_vp, 2))
        return dict(pred=pred, rgb=rgb, param=param, pairs=pairs, min_prob=0.46)
    c["d16"] = dict(batch(16, 3, 2, empty=(1,)), Dcenter=12)
    c["d32"] = dict(batch(32, 2, 3), Dcenter=26)                   # params.py:107 __cube_Dcenter[32] = 26
    c["all_empty"] = dict(batch(8, 2, 1, empty=(0, 1)), Dcenter=6)
    return c
