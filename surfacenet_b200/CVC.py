"""Drop-in for the reference's utils/CVC.py (Colored Voxel Cube), computed on the GPU.

Signatures, argument meaning, output layout and dtype follow utils/CVC.py:56-57,108; the work is
done by ``sn_cvc_gather`` (surfacenet_b200/csrc/cvc.cu).
"""
import numpy as np
from . import _lib
from .device import DeviceScene

_scene_cache = {}


def _scene_for(cameraPOs, models_img, views):
    """Per-scene constants are uploaded once and reused while the caller passes the same objects."""
    key = (id(models_img), id(cameraPOs))
    sc = _scene_cache.get(key)
    need = set(int(v) for v in np.unique(views))
    if sc is None or not need <= sc[0].loaded_views or sc[1] is not models_img:
        have = set() if sc is None else sc[0].loaded_views
        _scene_cache.clear()
        sc = (DeviceScene(cameraPOs, models_img, views=sorted(need | have)), models_img, cameraPOs)
        _scene_cache[key] = sc
    return sc[0]


def _cube_args(torch, selected_viewPairs, xyz, resol):
    pairs = np.asarray(selected_viewPairs)
    if pairs.ndim != 3 or pairs.shape[2] != 2:
        raise ValueError("selected_viewPairs must have shape (N_cubes, N_viewPairs, 2), got {}".format(pairs.shape))
    B, n_vp = pairs.shape[:2]
    xyz = np.ascontiguousarray(np.asarray(xyz, dtype=np.float32).reshape(B, 3))
    resol = np.ascontiguousarray(np.asarray(resol, dtype=np.float32).reshape(B))
    views = np.ascontiguousarray(pairs.reshape(B, 2 * n_vp).astype(np.int32))
    return B, n_vp, torch.from_numpy(xyz).cuda(), torch.from_numpy(resol).cuda(), torch.from_numpy(views).cuda(), views


def gen_coloredCubes_device(scene, selected_viewPairs, xyz, resol, colorize_cube_D, mean6=None, return_index=False):
    """GPU-resident variant: returns a torch.cuda float32 tensor (N_cubes*N_vp, 6, D,D,D); with
    ``return_index`` also the int32 (w, h) maps and the in-scope mask of utils/CVC.py:39-45,
    each (N_cubes, 2*N_vp, D^3)."""
    torch = _lib.require_cuda()
    B, n_vp, xyz_d, resol_d, views_d, views_h = _cube_args(torch, selected_viewPairs, xyz, resol)
    scene.check_views(views_h)
    D = int(colorize_cube_D)
    X = torch.empty((B * n_vp, 6, D, D, D), dtype=torch.float32, device="cuda")
    iw = ih = ins = None
    if return_index:
        iw = torch.empty((B, 2 * n_vp, D ** 3), dtype=torch.int32, device="cuda")
        ih = torch.empty_like(iw)
        ins = torch.empty((B, 2 * n_vp, D ** 3), dtype=torch.uint8, device="cuda")
    mean_d = None if mean6 is None else torch.as_tensor(np.asarray(mean6, np.float32).reshape(6)).cuda()
    for b0 in range(0, B, 2048):                    # grid.y limit: <= 65535 (cube, slot) rows per launch
        b1 = min(B, b0 + 2048)
        _lib.check(_lib.lib.sn_cvc_gather(
            _lib.ptr(scene.images), _lib.ptr(scene.img_offset), _lib.ptr(scene.img_hw), scene.n_views, _lib.ptr(scene.P),
            _lib.ptr(xyz_d[b0:]), _lib.ptr(resol_d[b0:]), _lib.ptr(views_d[b0:]), b1 - b0, n_vp, D, _lib.ptr(mean_d),
            _lib.ptr(X[b0 * n_vp:]), _lib.ptr(None if iw is None else iw[b0:]), _lib.ptr(None if ih is None else ih[b0:]),
            _lib.ptr(None if ins is None else ins[b0:]), _lib.stream_ptr()))
    if return_index:
        return X, iw, ih, ins
    return X


def gen_coloredCubes(selected_viewPairs, xyz, resol, cameraPOs, models_img, colorize_cube_D, visualization_ON=False,
                     occupiedCubes_01=None):
    """utils/CVC.py:56-104.
    inputs:  selected_viewPairs (N_cubes, N_select_viewPairs, 2); xyz (N_cubes,3), resol (N_cubes,)
             cameraPOs (N_views,3,4) float64; models_img list of (H,W,3) uint8
    return:  coloredCubes (N_cubes*N_select_viewPairs, 3*2, D,D,D) float32
    """
    if visualization_ON:
        raise NotImplementedError("visualization_ON is a debugging branch of the reference (CVC.py:49-50) and is not provided")
    scene = _scene_for(cameraPOs, models_img, np.asarray(selected_viewPairs))
    X = gen_coloredCubes_device(scene, selected_viewPairs, xyz, resol, colorize_cube_D)
    return X.cpu().numpy()


def preprocess_augmentation(gt_sub, X_sub, mean_rgb, augment_ON=True, crop_ON=True):
    """utils/CVC.py:108-122.  Only augment_ON=False, crop_ON=False is live in the reference (the other
    branches call helpers that are defined nowhere); anything else raises NameError there, here
    NotImplementedError.  X_sub: numpy or torch.cuda tensor (N, C, D,D,D); mean_rgb broadcastable (1,C,1,1,1)."""
    if augment_ON or crop_ON:
        raise NotImplementedError("augment_ON / crop_ON use helpers the reference never defines (CVC.py:113-121)")
    torch = _lib.require_cuda()
    is_np = isinstance(X_sub, np.ndarray)
    X = torch.from_numpy(np.ascontiguousarray(X_sub, dtype=np.float32)).cuda() if is_np else X_sub.to(torch.float32).contiguous().clone()
    C_ = X.shape[1]
    mean = np.broadcast_to(np.asarray(mean_rgb, dtype=np.float32).reshape(-1), (C_,)) if np.size(mean_rgb) in (1, C_) else None
    if mean is None:
        raise ValueError("mean_rgb must broadcast over the channel axis ({} channels)".format(C_))
    mean_d = torch.from_numpy(np.ascontiguousarray(mean)).cuda()
    spatial = int(np.prod(X.shape[2:]))
    _lib.check(_lib.lib.sn_sub_channel_mean(_lib.ptr(X), X.shape[0], C_, spatial, _lib.ptr(mean_d), _lib.stream_ptr()))
    return gt_sub, (X.cpu().numpy() if is_np else X)
