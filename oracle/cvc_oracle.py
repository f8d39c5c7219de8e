"""Oracle restatement of utils/CVC.py (Colored Voxel Cube).  Test infrastructure only.

Only mechanical py3 changes w.r.t. the reference (the py2 ``print`` at CVC.py:74 sits in a
visualisation branch that is never taken on the hot path and is dropped).
"""
import numpy as np


def colorize_cube(view_set, cameraPOs_np, model_imgs_np, xyz, resol, colorize_cube_D, return_index=False):
    """utils/CVC.py:6-53 (__colorize_cube__).  -> (N_views,3,D,D,D) float64.

    With ``return_index`` also returns the int32 (pts_w, pts_h) pairs and the in-scope mask
    (CVC.py:39-45) per view: the "voxel -> pixel index map" that must be bit exact.
    """
    D = colorize_cube_D
    min_x, min_y, min_z = xyz
    indx_xyz = range(0, D)
    indx_x, indx_y, indx_z = np.meshgrid(indx_xyz, indx_xyz, indx_xyz, indexing='ij')   # CVC.py:15
    indx_x = indx_x * resol + min_x                                                       # CVC.py:16 (int64*f32+f32 -> f64)
    indx_y = indx_y * resol + min_y
    indx_z = indx_z * resol + min_z
    homogen_1s = np.ones(D ** 3, dtype=np.float64)
    pts_4D = np.vstack([indx_x.flatten(), indx_y.flatten(), indx_z.flatten(), homogen_1s])  # CVC.py:20
    N_views = len(view_set)
    colored_cubes = np.zeros((N_views, 3, D, D, D))                                       # CVC.py:23
    idx_w, idx_h, idx_in = [], [], []
    for _n, _view in enumerate(view_set):                                                 # CVC.py:34
        projection_M = cameraPOs_np[_view]
        pts_3D = np.dot(projection_M, pts_4D)                                             # CVC.py:37
        pts_3D[:-1] /= pts_3D[-1]                                                         # CVC.py:38
        with np.errstate(invalid='ignore'):
            pts_2D = pts_3D[:-1].round().astype(np.int32)                                 # CVC.py:39
        pts_w, pts_h = pts_2D[0], pts_2D[1]
        pts_RGB = np.zeros((D ** 3, 3))                                                   # CVC.py:42
        img = model_imgs_np[_view]
        max_h, max_w, _ = img.shape
        inScope = (pts_w < max_w) & (pts_h < max_h) & (pts_w >= 0) & (pts_h >= 0)         # CVC.py:45
        pts_RGB[inScope] = img[pts_h[inScope], pts_w[inScope]]                            # CVC.py:46
        colored_cubes[_n] = pts_RGB.T.reshape((3, D, D, D))                               # CVC.py:47
        if return_index:
            idx_w.append(pts_w); idx_h.append(pts_h); idx_in.append(inScope)
    if return_index:
        return colored_cubes, np.stack(idx_w), np.stack(idx_h), np.stack(idx_in)
    return colored_cubes


def gen_coloredCubes(selected_viewPairs, xyz, resol, cameraPOs, models_img, colorize_cube_D,
                     visualization_ON=False, occupiedCubes_01=None):
    """utils/CVC.py:56-104.  -> (N_cubes*N_vp, 6, D,D,D) float32."""
    N_cubes, N_vp = selected_viewPairs.shape[:2]
    D = colorize_cube_D
    coloredCubes = np.zeros((N_cubes, N_vp * 2, 3) + (D,) * 3, dtype=np.float32)          # CVC.py:67
    for _n_cube in range(N_cubes):                                                        # CVC.py:71
        selected_views = selected_viewPairs[_n_cube].flatten()                            # CVC.py:82
        coloredCubes[_n_cube] = colorize_cube(selected_views, cameraPOs, models_img,      # CVC.py:84-86,102
                                              xyz[_n_cube], resol[_n_cube], D)
    return coloredCubes.reshape((N_cubes * N_vp, 3 * 2) + (D,) * 3)                       # CVC.py:104


def gen_index_map(selected_viewPairs, xyz, resol, cameraPOs, models_img, colorize_cube_D):
    """(w, h, in) per (cube, view slot, voxel): int32 (B, 2*N_vp, D^3) x2 and bool mask (CVC.py:39-45)."""
    N_cubes = selected_viewPairs.shape[0]
    W, H, I = [], [], []
    for b in range(N_cubes):
        _, w, h, m = colorize_cube(selected_viewPairs[b].flatten(), cameraPOs, models_img, xyz[b], resol[b],
                                   colorize_cube_D, return_index=True)
        W.append(w); H.append(h); I.append(m)
    return np.stack(W), np.stack(H), np.stack(I)


def preprocess_augmentation(gt_sub, X_sub, mean_rgb, augment_ON=True, crop_ON=True):
    """utils/CVC.py:108-122 with augment_ON=False, crop_ON=False (the only live branch: the other
    two call functions that are not defined anywhere in the reference)."""
    X_sub = X_sub.astype(np.float32)                                                      # CVC.py:110
    X_sub -= mean_rgb                                                                     # CVC.py:111
    if augment_ON or crop_ON:
        raise NotImplementedError("dead branches in the reference (undefined helpers, CVC.py:113-121)")
    return gt_sub, X_sub
