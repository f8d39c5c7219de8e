"""Oracle restatement of utils/sparseCubes.py:9-141 (dense2sparse, append_dense_2sparseList) and
utils/utils.py:8-42 (generate_voxelLevelWeighted_coloredCubes).  Test infrastructure only.

Mechanical py3 fix only: the integer divisions `(D_orig-cube_Dcenter)/2` of sparseCubes.py:51 are `//`.
"""
import numpy as np
from . import raypool_oracle as rayPooling


def generate_voxelLevelWeighted_coloredCubes(viewPair_coloredCubes, viewPair_surf_predictions, weight4viewPair):
    """utils/utils.py:8-42 -> (N_cubes, 3, D,D,D) uint8."""
    N_cubes, N_viewPairs, _D = viewPair_surf_predictions.shape[:3]
    voxel_weight = weight4viewPair[..., None, None, None] * viewPair_surf_predictions              # utils.py:32
    voxel_weight /= np.sum(voxel_weight, axis=1, keepdims=True)                                    # utils.py:33
    mean_viewPair_coloredCubes = np.mean(viewPair_coloredCubes.astype(np.float32).reshape(
        (N_cubes, N_viewPairs, 2, 3, _D, _D, _D)), axis=2)                                         # utils.py:37
    new_coloredCubes = np.sum(voxel_weight[:, :, None, ...] * mean_viewPair_coloredCubes, axis=1)  # utils.py:40
    return new_coloredCubes.astype(np.uint8)                                                       # utils.py:42


def dense2sparse(prediction, rgb, param, viewPair, min_prob=0.5, rayPool_thresh=0, enable_centerCrop=False, cube_Dcenter=None,
                 enable_rayPooling=False, cameraPOs=None, cameraTs=None):
    """utils/sparseCubes.py:9-77."""
    N_cubes, D_orig, _, _ = prediction.shape
    nonempty_cube_indx, vxl_ijk_list, prediction_list, rgb_list, rayPooling_votes_list = [], [], [], [], []
    param_new = np.copy(param)
    slc = np.s_[:, :, :]
    if enable_centerCrop:
        _Cmin, _Cmax = (D_orig - cube_Dcenter) // 2, (D_orig - cube_Dcenter) // 2 + cube_Dcenter   # :51
        slc = (slice(_Cmin, _Cmax, 1),) * 3
        param_new['xyz'] += param_new['resol'][:, None] * _Cmin                                    # :55
    for _n in range(N_cubes):                                                                      # :57
        if enable_rayPooling:
            rayPool_votes = rayPooling.rayPooling_1cube_numpy(cameraPOs, cameraTs, viewPair_viewIndx=viewPair[_n],
                                                              xyz=param[_n]['xyz'], resol=param[_n]['resol'],
                                                              cube_prediction=prediction[_n], prediction_thresh=min_prob).astype(np.uint8)
            vxl_ijk_tuple = np.where(rayPool_votes[slc] >= rayPool_thresh)                         # :64
        if (not enable_rayPooling) or rayPool_thresh == 0:
            vxl_ijk_tuple = np.where(prediction[_n][slc] > min_prob)                               # :66
        if vxl_ijk_tuple[0].size == 0:
            continue
        nonempty_cube_indx.append(_n)
        vxl_ijk_list.append(np.c_[vxl_ijk_tuple].astype(np.uint8))                                 # :71
        prediction_list.append(prediction[_n][slc][vxl_ijk_tuple].astype(np.float16))
        rgb_list.append(rgb[_n][slc][vxl_ijk_tuple].astype(np.uint8))
        if enable_rayPooling:
            rayPooling_votes_list.append(rayPool_votes[slc][vxl_ijk_tuple].astype(np.uint8))       # :75
    return nonempty_cube_indx, vxl_ijk_list, prediction_list, rgb_list, rayPooling_votes_list, param_new


def append_dense_2sparseList(prediction_sub, rgb_sub, param_sub, viewPair_sub, min_prob=0.5, rayPool_thresh=0,
                             enable_centerCrop=False, cube_Dcenter=None, enable_rayPooling=False, cameraPOs=None, cameraTs=None,
                             prediction_list=None, rgb_list=None, vxl_ijk_list=None, rayPooling_votes_list=None,
                             cube_ijk_np=None, param_np=None, viewPair_np=None):
    """utils/sparseCubes.py:82-141 (the mutable default lists of the reference are passed explicitly by its caller)."""
    prediction_list = [] if prediction_list is None else prediction_list
    rgb_list = [] if rgb_list is None else rgb_list
    vxl_ijk_list = [] if vxl_ijk_list is None else vxl_ijk_list
    rayPooling_votes_list = [] if rayPooling_votes_list is None else rayPooling_votes_list
    if prediction_sub.ndim == 5:
        prediction_sub = prediction_sub.astype(np.float16)[:, 0]                                   # :115
    rgb_sub = np.transpose(rgb_sub.astype(np.uint8), axes=(0, 2, 3, 4, 1))                         # :116
    cube_ijk_sub = param_sub['ijk']
    viewPair_sub = viewPair_sub.astype(np.uint16)                                                  # :119
    out = dense2sparse(prediction=prediction_sub, rgb=rgb_sub, param=param_sub, viewPair=viewPair_sub, min_prob=min_prob,
                       rayPool_thresh=rayPool_thresh, enable_centerCrop=enable_centerCrop, cube_Dcenter=cube_Dcenter,
                       enable_rayPooling=enable_rayPooling, cameraPOs=cameraPOs, cameraTs=cameraTs)
    idx, ijk_l, pred_l, rgb_l, votes_l, param_new_sub = out
    param_sub = param_new_sub[idx]
    viewPair_sub = viewPair_sub[idx]
    cube_ijk_sub = cube_ijk_sub[idx]
    if not len(pred_l) == len(rgb_l) == len(ijk_l) == param_sub.shape[0] == viewPair_sub.shape[0] == cube_ijk_sub.shape[0]:
        raise Warning('load dense data, # of cubes is not consistent.')                           # :131
    prediction_list.extend(pred_l); rgb_list.extend(rgb_l); vxl_ijk_list.extend(ijk_l); rayPooling_votes_list.extend(votes_l)
    param_np = param_sub if param_np is None else np.concatenate([param_np, param_sub], axis=0)
    viewPair_np = viewPair_sub if viewPair_np is None else np.vstack([viewPair_np, viewPair_sub])
    cube_ijk_np = cube_ijk_sub if cube_ijk_np is None else np.vstack([cube_ijk_np, cube_ijk_sub])
    return prediction_list, rgb_list, vxl_ijk_list, rayPooling_votes_list, cube_ijk_np, param_np, viewPair_np
