"""Oracle restatement of utils/rayPooling.py:143-260 and the utils/sparseCubes.py:44-77,114-120
consumer slice.  Test infrastructure only.

Mechanical numpy-2 fixes only: ``np.unravel_index(dims=)`` -> positional shape,
``np.bool`` -> ``bool``, ``np.unique(..., return_inverse=True)`` result ``.ravel()``-ed.
"""
import numpy as np
from . import camera_oracle as camera


def rayPooling_1cube_numpy(cameraPOs, cameraTs, cube_prediction, viewPair_viewIndx, xyz, resol, prediction_thresh=None):
    cube_prediction = cube_prediction.squeeze()                                           # rayPooling.py:200
    if cube_prediction.ndim != 3:
        raise ValueError('rayPooling method argument cube_prediction has {} dims'.format(cube_prediction.ndim))
    cube_shape = cube_prediction.shape[-3:]
    N_channels = 2
    viewIndx_set, viewIndx_inverseIndx = np.unique(viewPair_viewIndx.flatten(), return_inverse=True)  # :210
    viewIndx_inverseIndx = viewIndx_inverseIndx.ravel()
    N_views_set = viewIndx_set.size
    view_POs = cameraPOs[viewIndx_set]                                                    # :213
    min_x, min_y, min_z = xyz
    pts_select = np.arange(cube_prediction.size) if prediction_thresh is None else \
        np.where(cube_prediction.flatten() > prediction_thresh)[0]                        # :218-219
    ijk_select = np.asarray(np.unravel_index(pts_select, cube_shape))                     # :222
    pts_xyz = ijk_select * resol + np.array([min_x, min_y, min_z])[:, None]               # :223
    img_h_abs, img_w_abs, depth = camera.perspectiveProj(projection_M=view_POs, xyz_3D=pts_xyz.T,
                                                         return_int_hw=True, return_depth=True)  # :228
    depth_resol = resol
    depth_int = (depth / depth_resol).round().astype(np.int32)                            # :233
    channels_infor = np.vstack([cube_prediction.flatten()[pts_select][None, ...], pts_select])  # :234
    cube_eachView_vote = np.zeros((N_views_set,) + cube_shape).astype(bool)               # :235
    for _view in range(N_views_set):                                                      # :237
        _depth_int = depth_int[_view]
        if _depth_int.size == 0:
            continue
        D_NDC2 = _depth_int.max() - _depth_int.min() + 1
        _img_w_abs, _img_h_abs = img_w_abs[_view], img_h_abs[_view]
        _img_wh_abs = np.c_[_img_w_abs, _img_h_abs]
        _dtype_wh = _img_w_abs.dtype.descr * 2
        _wh_tpl = _img_wh_abs.view(_dtype_wh)
        _wh_tpl_set, _wh_tpl_indx = np.unique(_wh_tpl, return_inverse=True)               # :246
        _wh_tpl_indx = _wh_tpl_indx.ravel()
        D_NDC1 = len(_wh_tpl_set)
        views_prediction_NDC = np.zeros((N_channels, D_NDC1, D_NDC2))                     # :249
        indx_NDC2 = _depth_int.flatten() - _depth_int.min()
        views_prediction_NDC[:, _wh_tpl_indx, indx_NDC2] = channels_infor                 # :251 (duplicates: last wins)
        argmax_NDC2 = np.argmax(views_prediction_NDC[0], axis=-1)                         # :253 (first max)
        rayPooling_indx = views_prediction_NDC[-1:, np.arange(D_NDC1), argmax_NDC2].astype(np.int32)
        cube_eachView_vote[_view][np.unravel_index(rayPooling_indx, cube_shape)] = True   # :256
    cube_N_votes = np.sum(cube_eachView_vote[viewIndx_inverseIndx], axis=0)               # :258
    return cube_N_votes


def votes_batch(prediction_f32, viewPairs, xyz, resol, cameraPOs, min_prob):
    """What utils/sparseCubes.py:114-120 + 57-62 do to a dense fused prediction batch:
    cast to float16 (115), ray-pool each cube with prediction_thresh=min_prob (60-62), cast uint8."""
    pred16 = prediction_f32.astype(np.float16)
    if pred16.ndim == 5:
        pred16 = pred16[:, 0]
    out = np.zeros(pred16.shape, dtype=np.uint8)
    vp = viewPairs.astype(np.uint16)                                                      # sparseCubes.py:119
    for n in range(pred16.shape[0]):
        out[n] = rayPooling_1cube_numpy(cameraPOs, None, pred16[n], vp[n], xyz[n], resol[n],
                                        prediction_thresh=min_prob).astype(np.uint8)
    return out


def dense2sparse_select(prediction_f16, cube_Dcenter, min_prob):
    """sparseCubes.py:49-55,65-66 with rayPool_thresh == 0 (main_reconstruct.py:156): the kept
    voxels are the centre crop's ``prediction > min_prob``; returns the per-cube ijk (uint8) lists."""
    N, D = prediction_f16.shape[:2]
    cmin = (D - cube_Dcenter) // 2
    sl = (slice(cmin, cmin + cube_Dcenter),) * 3
    return [np.c_[np.where(prediction_f16[n][sl] > min_prob)].astype(np.uint8) for n in range(N)]
