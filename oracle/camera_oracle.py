"""Oracle restatement of utils/camera.py:123-184 (perspectiveProj).  Test infrastructure only."""
import numpy as np


def perspectiveProj(projection_M, xyz_3D, return_int_hw=True, return_depth=False):
    """(N_Ms,3,4)|(3,4) f64 x (N_pts,3)|(3,) -> img_h, img_w[, depth]   (utils/camera.py:123-184).

    fp64 matmul (camera.py:172-173), rows 0..1 divided in place by row 2 (176), optional
    round-half-even -> int64 (177-178), w = row 0, h = row 1 (179), depth = row 2 (181).
    """
    if projection_M.shape[-2:] != (3, 4):                                   # camera.py:163-164
        raise ValueError("perspectiveProj needs projection_M with shape (3,4), however got {}".format(projection_M.shape))
    if xyz_3D.ndim == 1:                                                    # camera.py:166-167
        xyz_3D = xyz_3D[None, :]
    if xyz_3D.shape[1] != 3 or xyz_3D.ndim != 2:                            # camera.py:169-170
        raise ValueError("perspectiveProj needs xyz_3D with shape (3,) or (N_pts, 3), however got {}".format(xyz_3D.shape))
    N_pts = xyz_3D.shape[0]
    xyz1 = np.c_[xyz_3D, np.ones((N_pts, 1))].astype(np.float64)            # camera.py:173
    pts_3D = np.matmul(projection_M, xyz1.T)                                # camera.py:174
    pts_2D = pts_3D[..., :2, :]
    pts_2D /= pts_3D[..., 2:3, :]                                           # camera.py:177
    if return_int_hw:
        pts_2D = pts_2D.round().astype(np.int64)                            # camera.py:179
    img_w, img_h = pts_2D[..., 0, :], pts_2D[..., 1, :]                     # camera.py:180
    if return_depth:
        depth = pts_3D[..., 2, :]                                           # camera.py:182
        return img_h, img_w, depth
    return img_h, img_w
