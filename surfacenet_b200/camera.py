"""Drop-ins for utils/camera.py: perspectiveProj (123-184, sn_perspective_proj) and viewPairAngles_wrt_pts (275-309,
sn_viewpair_angles), evaluated on the GPU."""
import numpy as np
from . import _lib


def perspectiveProj(projection_M, xyz_3D, return_int_hw=True, return_depth=False):
    """projection_M (3,4) / (N_Ms,3,4); xyz_3D (3,) / (N_pts,3) -> img_h, img_w[, depth], each
    (N_pts,) / (N_Ms,N_pts); int64 when return_int_hw (round half to even), float64 otherwise."""
    projection_M = np.asarray(projection_M)
    xyz_3D = np.asarray(xyz_3D)
    if projection_M.shape[-2:] != (3, 4):
        raise ValueError("perspectiveProj needs projection_M with shape (3,4), however got {}".format(projection_M.shape))
    if xyz_3D.ndim == 1:
        xyz_3D = xyz_3D[None, :]
    if xyz_3D.ndim != 2 or xyz_3D.shape[1] != 3:
        raise ValueError("perspectiveProj needs xyz_3D with shape (3,) or (N_pts, 3), however got {}".format(xyz_3D.shape))
    torch = _lib.require_cuda()
    single = projection_M.ndim == 2
    P = torch.from_numpy(np.ascontiguousarray(projection_M.reshape(-1, 3, 4), dtype=np.float64)).cuda()
    pts = torch.from_numpy(np.ascontiguousarray(xyz_3D, dtype=np.float64)).cuda()
    nM, nP = P.shape[0], pts.shape[0]
    h = torch.empty((nM, nP), dtype=torch.float64, device="cuda")
    w = torch.empty_like(h)
    d = torch.empty_like(h) if return_depth else None
    _lib.check(_lib.lib.sn_perspective_proj(_lib.ptr(P), nM, _lib.ptr(pts), nP, 1 if return_int_hw else 0,
                                            _lib.ptr(h), _lib.ptr(w), _lib.ptr(d), _lib.stream_ptr()))
    h, w = h.cpu().numpy(), w.cpu().numpy()
    if return_int_hw:
        with np.errstate(invalid="ignore"):
            h, w = h.astype(np.int64), w.astype(np.int64)
    if single:
        h, w = h[0], w[0]
    if return_depth:
        d = d.cpu().numpy()
        return h, w, (d[0] if single else d)
    return h, w


def viewPairAngles_wrt_pts(cameraTs, pts_xyz, viewPairs=None, device_out=False):
    """utils/camera.py:275-309: angle <camera_i, point, camera_j> for every point and every 2-combination of cameras (or the given
    viewPairs) -> (N_pts, N_viewPairs), in the promoted dtype of the inputs (float32 only when both are float32)."""
    from .utils import k_combination_np
    torch = _lib.require_cuda()
    cameraTs, pts_xyz = np.asarray(cameraTs), np.asarray(pts_xyz)
    if cameraTs.ndim != 2 or cameraTs.shape[1] != 3 or pts_xyz.ndim != 2 or pts_xyz.shape[1] != 3:
        raise ValueError("need cameraTs (N_views,3) and pts_xyz (N_pts,3), got {} {}".format(cameraTs.shape, pts_xyz.shape))
    dt = np.float32 if (cameraTs.dtype == np.float32 and pts_xyz.dtype == np.float32) else np.float64
    pairs = k_combination_np(range(cameraTs.shape[0]), k=2) if viewPairs is None else np.asarray(viewPairs)
    cam = torch.from_numpy(np.ascontiguousarray(cameraTs, dtype=dt)).cuda()
    pts = torch.from_numpy(np.ascontiguousarray(pts_xyz, dtype=dt)).cuda()
    vp = torch.from_numpy(np.ascontiguousarray(pairs.reshape(-1, 2), dtype=np.int32)).cuda()
    out = torch.empty((pts.shape[0], vp.shape[0]), dtype=torch.float64 if dt == np.float64 else torch.float32, device="cuda")
    _lib.check(_lib.lib.sn_viewpair_angles(_lib.ptr(cam), _lib.ptr(pts), cam.shape[0], pts.shape[0], _lib.ptr(vp), vp.shape[0],
                                           1 if dt == np.float64 else 0, _lib.ptr(out), _lib.stream_ptr()))
    return out if device_out else out.cpu().numpy()


def __cameraP2T__(cameraPO):
    """utils/camera.py:88-100: camera centre in world coordinates from a (3,4) projection matrix (host; four 3x3 determinants)."""
    cameraPO = np.asarray(cameraPO)
    homo4D = np.array([np.linalg.det(cameraPO[:, [1, 2, 3]]), -1 * np.linalg.det(cameraPO[:, [0, 2, 3]]),
                       np.linalg.det(cameraPO[:, [0, 1, 3]]), -1 * np.linalg.det(cameraPO[:, [0, 1, 2]])])
    return homo4D[:3] / homo4D[3]


def cameraPs2Ts(cameraPOs):
    """utils/camera.py:103-120."""
    Ts = [__cameraP2T__(P) for P in cameraPOs]
    return Ts if type(cameraPOs) is list else np.stack(Ts)


def perspectiveProj_cubesCorner(projection_M, cube_xyz_min, cube_D_mm, return_int_hw=True, return_depth=False):
    """utils/camera.py:186-250: projections of the 8 corners of every cube -> img_h, img_w (N_Ms, N_pts, 8) (GPU projection)."""
    projection_M, cube_xyz_min = np.asarray(projection_M), np.asarray(cube_xyz_min)
    if projection_M.shape[-2:] != (3, 4):
        raise ValueError("perspectiveProj needs projection_M with shape (3,4), however got {}".format(projection_M.shape))
    if cube_xyz_min.ndim == 1:
        cube_xyz_min = cube_xyz_min[None, :]
    if cube_xyz_min.ndim != 2 or cube_xyz_min.shape[1] != 3:
        raise ValueError("perspectiveProj needs cube_xyz_min with shape (3,) or (N_pts, 3), however got {}".format(cube_xyz_min.shape))
    N_pts = cube_xyz_min.shape[0]
    cubeCorner_shift = np.indices((2, 2, 2)).reshape((3, -1)).T[None, :, :] * cube_D_mm
    cubeCorner = cube_xyz_min[:, None, :] + cubeCorner_shift
    img_h, img_w = perspectiveProj(projection_M=projection_M, xyz_3D=cubeCorner.reshape((N_pts * 8, 3)), return_int_hw=return_int_hw)
    return img_h.reshape((-1, N_pts, 8)), img_w.reshape((-1, N_pts, 8))


def __readCameraPO_as_np_DTU__(cameraPO_file):
    """utils/camera.py:7-24: one (3,4) float64 projection matrix per text file."""
    return np.loadtxt(cameraPO_file, dtype=np.float64, delimiter=' ')


def __readCameraPOs_as_np_Middlebury__(cameraPO_file, viewList):
    """utils/camera.py:26-53.  Middlebury `*_par.txt`: a count line, then one line per image `name k11..k33 r11..r33 t1 t2 t3`;
    P = K [R | t].  Row n of the table is view n (1-based: row 0, the count line, is never selected)."""
    table = np.genfromtxt(cameraPO_file, skip_header=1, usecols=range(1, 22), dtype=np.float64).reshape(-1, 21)
    K, R, t = table[:, :9].reshape(-1, 3, 3), table[:, 9:18].reshape(-1, 3, 3), table[:, 18:21].reshape(-1, 3, 1)
    P = np.matmul(K, np.concatenate([R, t], axis=2))                        # (N_images, 3, 4)
    return P[np.asarray(list(viewList), dtype=np.int64) - 1]                 # view n <-> data row n - 1


def readCameraPOs_as_np(datasetFolder, datasetName, poseNamePattern, model, viewList):
    """utils/camera.py:55-82 -> (N_views,3,4) float64; '#' -> zero-padded view index, '@' -> plain view index."""
    import os
    if 'Middlebury' in datasetName:
        return __readCameraPOs_as_np_Middlebury__(os.path.join(datasetFolder, poseNamePattern), viewList)
    cameraPOs = np.empty((len(viewList), 3, 4), dtype=np.float64)
    for _i, _view in enumerate(viewList):
        cameraPOs[_i] = __readCameraPO_as_np_DTU__(os.path.join(
            datasetFolder, poseNamePattern.replace('#', '{:03}'.format(_view)).replace('@', '{}'.format(_view))))
    return cameraPOs
