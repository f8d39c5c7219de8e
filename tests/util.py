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


# ---- "next" row N4: sparse scenes for denoising.denoise_crossCubes / adapthresh.adapthresh ------------------------------
def sparse_scene(grid, D, seed=0, noise=0.15, floaters=6, min_prob=0.46, shuffle=False, empty=(), dup_last=False, thick=0.09,
                 max_votes=10, gain_spread=0.35):
    """A wavy sheet cut by a grid of half-overlapping cubes (cube (i,j,k) starts at voxel (i,j,k)*D/2 of a global grid), as
    main_reconstruct.py leaves it after append_dense_2sparseList: per-cube voxel lists (np.where order unless `shuffle`),
    float16 predictions (each cube sees its own noise, as overlapping CNN windows do), uint8 ray-pool votes, plus a few
    isolated floaters per cube (the noise denoise_crossCubes removes).
    -> dict(cube_ijk (C,3) uint32, ijk_list, pred_list, votes_list, rgb_list, param (C,) PARAM_DTYPE)"""
    rs = np.random.RandomState(seed)
    h = D // 2
    gi, gj, gk = grid
    ext = np.array([(gi + 1) * h, (gj + 1) * h, (gk + 1) * h], np.float64)
    cube_ijk, ijk_l, pred_l, votes_l, rgb_l = [], [], [], [], []
    g = np.arange(D)
    for ci in range(gi):
        for cj in range(gj):
            for ck in range(gk):
                n = len(cube_ijk)
                X, Y, Z = np.meshgrid(ci * h + g, cj * h + g, ck * h + g, indexing="ij")
                x, y, z = X / ext[0], Y / ext[1], Z / ext[2]
                d = z - (0.5 + 0.22 * np.sin(5.0 * x + seed) * np.cos(4.0 * y))
                sig = thick * (1.0 + gain_spread * (2 * rs.rand() - 1))            # each cube sees a slightly thicker / thinner sheet
                p = np.exp(-d * d / (2 * sig * sig)) * (1.0 - noise * rs.rand(D, D, D))
                fl = rs.randint(0, D, size=(floaters, 3))
                p[fl[:, 0], fl[:, 1], fl[:, 2]] = 0.6 + 0.39 * rs.rand(floaters)
                if n in empty:
                    p[:] = 0.1
                p16 = p.astype(np.float16)
                sel = np.where(p16 > np.float16(min_prob))
                ijk = np.c_[sel].astype(np.uint8)
                pr = p16[sel]
                votes = np.clip(np.round(pr.astype(np.float64) * max_votes + rs.randint(-2, 3, size=pr.shape)), 0, max_votes).astype(np.uint8)
                rgb = rs.randint(0, 256, size=(ijk.shape[0], 3)).astype(np.uint8)
                if shuffle:
                    o = rs.permutation(ijk.shape[0])
                    ijk, pr, votes, rgb = ijk[o], pr[o], votes[o], rgb[o]
                cube_ijk.append((ci, cj, ck)); ijk_l.append(ijk); pred_l.append(pr); votes_l.append(votes); rgb_l.append(rgb)
    if dup_last:                                             # a repeated cube index: the reference's dict keeps the later one
        cube_ijk.append(cube_ijk[0]); ijk_l.append(ijk_l[1].copy()); pred_l.append(pred_l[1].copy())
        votes_l.append(votes_l[1].copy()); rgb_l.append(rgb_l[1].copy())
    C = len(cube_ijk)
    param = np.zeros(C, PARAM_DTYPE)
    param["ijk"] = np.asarray(cube_ijk, np.uint32)
    param["resol"] = np.float32(0.4)
    param["xyz"] = (np.asarray(cube_ijk, np.float64) * h * 0.4 + np.array([-30.0, 10.0, 600.0])).astype(np.float32)
    return dict(cube_ijk=np.asarray(cube_ijk, np.uint32), ijk_list=ijk_l, pred_list=pred_l, votes_list=votes_l, rgb_list=rgb_l,
                param=param, D=D)


def post_cases():
    """Small scenes (every per-half-cube count stays <= 2048, see oracle/postprocess_oracle.py) for the golden vectors."""
    c = {}
    c["g322_d12"] = dict(scene=sparse_scene((3, 2, 2), 12, seed=1), init=0.5, maxp=0.9, rp=3, beta=6, iters=4)
    c["g233_d16_shuffled"] = dict(scene=sparse_scene((2, 3, 3), 16, seed=2, shuffle=True, empty=(4,), thick=0.06), init=0.5, maxp=0.9, rp=2,
                                  beta=6, iters=5)
    c["g222_d13_odd_dup"] = dict(scene=sparse_scene((2, 2, 2), 13, seed=3, dup_last=True), init=0.6, maxp=0.8, rp=0, beta=3, iters=3)
    c["g141_d10_thin"] = dict(scene=sparse_scene((1, 4, 1), 10, seed=4, thick=0.04, floaters=12), init=0.5, maxp=0.9, rp=4, beta=6, iters=3)
    return c


def read_ply(path):
    """Minimal reader of the binary little-endian vertex PLY files sparseCubes.save2ply writes -> (xyz f32 (N,3), rgb u8 (N,3) | None)."""
    with open(path, "rb") as f:
        raw = f.read()
    end = raw.index(b"end_header\n") + len(b"end_header\n")
    lines = raw[:end].decode("ascii").split("\n")
    assert lines[0] == "ply" and lines[1] == "format binary_little_endian 1.0"
    n = int([l for l in lines if l.startswith("element vertex")][0].split()[-1])
    props = [l.split()[1:] for l in lines if l.startswith("property")]
    dt = np.dtype([(name, {"float": "<f4", "uchar": "u1"}[t]) for t, name in props])
    v = np.frombuffer(raw[end:], dtype=dt, count=n)
    xyz = np.stack([v["x"], v["y"], v["z"]], axis=1)
    rgb = np.stack([v["red"], v["green"], v["blue"]], axis=1) if "red" in dt.names else None
    return xyz, rgb


# ---- "next" row N3: early rejection / view-pair selection -----------------------------------------------------------------
MEAN_BGR = np.asarray([103.939, 116.779, 123.68], dtype=np.float32)                           # params.py:131


def fake_patch2embedding_fn(patches):
    """Deterministic stand-in for the similarityNet embedding: 16 fixed linear functionals of the (N,3,64,64) patch."""
    p = np.asarray(patches, np.float32)
    rs = np.random.RandomState(9)
    A = rs.randn(16, 3, 8, 8).astype(np.float32) / 64.0
    pooled = p.reshape(p.shape[0], 3, 8, 8, 8, 8).mean(axis=(3, 5), dtype=np.float64)
    return np.einsum("ncij,kcij->nk", pooled, A.astype(np.float64)).astype(np.float32)


def fake_pair2simil_fn(pairs):
    e = np.asarray(pairs, np.float64).reshape(-1, 2, pairs.shape[-1])
    d = np.sqrt(((e[:, 0] - e[:, 1]) ** 2).sum(axis=1, keepdims=True))
    return (1.0 / (1.0 + np.exp(-(0.02 * d - 1.2)))).astype(np.float32)


def select_case(cams):
    """6 views of small synthetic images, a 3x3x2 grid of cube centres around the scan9 volume, their projected corners."""
    rs = np.random.RandomState(31)
    views = np.array([3, 8, 15, 22, 30, 41])
    centers = np.array([[x, y, z] for x in (-20.0, 10.0, 40.0) for y in (-40.0, 0.0, 40.0) for z in (600.0, 660.0)], np.float64)
    N = centers.shape[0]
    half = 12.8
    corners = centers[:, None, :] + half * np.array([[a, b, c] for a in (-1, 1) for b in (-1, 1) for c in (-1, 1)], np.float64)[None]
    P = cams[views]
    def proj(pts):                                   # camera.perspectiveProj arithmetic (float64), h = row, w = column
        ph = np.concatenate([pts, np.ones((pts.shape[0], 1))], axis=1)
        q = np.matmul(P, ph.T)                        # (V,3,n)
        return q[:, 1] / q[:, 2], q[:, 0] / q[:, 2]
    h_c, w_c = proj(corners.reshape(-1, 3))
    h_corner, w_corner = h_c.reshape(len(views), N, 8), w_c.reshape(len(views), N, 8)
    ch, cw = proj(centers)
    images = [synth_image(100 + int(v), 1200 if i % 2 == 0 else 900, 1600 if i % 3 else 1100) for i, v in enumerate(views)]
    viewPairs = np.asarray([(a, b) for a in range(len(views)) for b in range(a + 1, len(views))])
    n_crop = 7
    return dict(views=views, centers=centers, N_cubes=N, h_corner=h_corner, w_corner=w_corner, center_hw=np.stack([ch, cw], axis=0),
                images=images, viewPairs=viewPairs, w_rand=rs.rand(N, viewPairs.shape[0]).astype(np.float32),
                range_h=np.stack([rs.rand(n_crop) * 200, 200 + rs.rand(n_crop) * 90], axis=1), range_w=np.stack([rs.rand(n_crop) * 300, 300 + rs.rand(n_crop) * 90], axis=1),
                crop_ch=np.array([5.2, 31.9, 150.5, 299.99, 280.0, 32.0, 100.7]), crop_cw=np.array([390.1, 10.0, 200.49, 399.0, 33.3, 32.0, 64.5]))
