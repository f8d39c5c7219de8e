"""Drop-in for utils/utils.py:8-42 (generate_voxelLevelWeighted_coloredCubes), computed on the GPU (sn_color_fusion).
"Next" row N2 of the scope table."""
import numpy as np
from . import _lib


def color_fusion_device(cvc, unfused, w, mean6=None):
    """cvc torch.cuda (B*N_vp,6,D,D,D) f32; unfused (B,N_vp,D,D,D) f32; w (B,N_vp) f32 | None (N_vp == 1);
    mean6: None, or a cuda (6,) f32 tensor added to the colours first (main_reconstruct.py:150).  -> (B,3,D,D,D) uint8."""
    torch = _lib.require_cuda()
    B, n_vp = int(unfused.shape[0]), int(unfused.shape[1])
    if cvc.shape[0] != B * n_vp or cvc.shape[1] != 6:
        raise ValueError("viewPair_coloredCubes must have shape (N_cubes*N_viewPairs, 6, D,D,D), got {}".format(tuple(cvc.shape)))
    vol = int(np.prod(unfused.shape[2:]))
    out = torch.empty((B, 3) + tuple(unfused.shape[2:]), dtype=torch.uint8, device="cuda")
    _lib.check(_lib.lib.sn_color_fusion(_lib.ptr(cvc.contiguous()), _lib.ptr(mean6), _lib.ptr(unfused.contiguous()),
                                        _lib.ptr(None if w is None else w.contiguous()), B, n_vp, vol, _lib.ptr(out), _lib.stream_ptr()))
    return out


def generate_voxelLevelWeighted_coloredCubes(viewPair_coloredCubes, viewPair_surf_predictions, weight4viewPair):
    """utils/utils.py:8.
    weight4viewPair (N_cubes, N_viewPairs); viewPair_surf_predictions (N_cubes, N_viewPairs, D,D,D);
    viewPair_coloredCubes (N_cubes * N_viewPairs, 6, D,D,D)  ->  new_coloredCubes (N_cubes, 3, D,D,D) uint8"""
    torch = _lib.require_cuda()
    p = np.asarray(viewPair_surf_predictions)
    if p.ndim != 5:
        raise ValueError("viewPair_surf_predictions must have shape (N_cubes, N_viewPairs, D,D,D), got {}".format(p.shape))
    w = np.asarray(weight4viewPair)
    if w.shape != p.shape[:2]:
        raise ValueError("weight4viewPair must have shape {}, got {}".format(p.shape[:2], w.shape))
    f32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()
    return color_fusion_device(f32(viewPair_coloredCubes), f32(p), f32(w)).cpu().numpy()


def k_combination_np(iterable, k=2):
    """utils/utils.py:233-256: all k-combinations of `iterable` as the rows of an array (host index bookkeeping)."""
    import itertools
    return np.asarray(list(itertools.combinations(iterable, k)))
