"""Drop-in for utils/viewPairSelection.py: view-pair angles, feature assembly, relative-importance network and the top-N
selection on the GPU.  "Next" row N3."""
import math
import numpy as np
from . import _lib, camera


def __argmaxN_viewPairs__(viewPairs, w_viewPairs, N_argmax):
    """utils/viewPairSelection.py:8-41 -> argmaxN_viewPairs (N_validCubes, N_argmax, 2), argmaxN_w (N_validCubes, N_argmax),
    ascending by weight like argsort()[:, -N:] (ties in index order)."""
    torch = _lib.require_cuda()
    viewPairs, w_viewPairs = np.asarray(viewPairs), np.asarray(w_viewPairs)
    if w_viewPairs.ndim != 2 or viewPairs.shape != (w_viewPairs.shape[1], 2):
        raise ValueError("need viewPairs (N_viewPairs,2) and w_viewPairs (N_validCubes,N_viewPairs), got {} {}".format(viewPairs.shape, w_viewPairs.shape))
    rows, n = w_viewPairs.shape
    N_argmax = min(int(N_argmax), n)              # argsort()[:, -N:] returns all n columns when N > n (viewPairSelection.py:36)
    w = torch.from_numpy(np.ascontiguousarray(w_viewPairs, dtype=np.float64)).cuda()
    idx = torch.zeros((rows, N_argmax), dtype=torch.int32, device="cuda")
    _lib.check(_lib.lib.sn_topn_rows(_lib.ptr(w), rows, n, int(N_argmax), _lib.ptr(idx), _lib.stream_ptr()))
    idx = idx.cpu().numpy().astype(np.int64)
    rr = np.arange(rows)[:, None]
    return viewPairs[idx], w_viewPairs[rr, idx]


def viewPairSelection(cameraTs_np, e_viewPairs, d_viewPairs, validCubes, cubeCenters_xyz, viewPair_relativeImpt_fn, batchSize,
                      N_viewPairs4inference, viewPairs):
    """utils/viewPairSelection.py:44-82 -> selected_viewPairs (N_validCubes, N, 2), selected_similNet_weight (N_validCubes, N)."""
    torch = _lib.require_cuda()
    validCubes = np.asarray(validCubes).astype(bool)
    viewPairs = np.asarray(viewPairs)
    N_cubes, N_viewPairs = d_viewPairs.shape[:2]
    N_validCubes = int(validCubes.sum())
    D_embedding = e_viewPairs.shape[-1]
    N_views = e_viewPairs.shape[1]
    theta = camera.viewPairAngles_wrt_pts(cameraTs=cameraTs_np, pts_xyz=cubeCenters_xyz[validCubes], viewPairs=viewPairs, device_out=True)
    e = torch.from_numpy(np.ascontiguousarray(e_viewPairs[validCubes], dtype=np.float32)).cuda()
    d = torch.from_numpy(np.ascontiguousarray(d_viewPairs[validCubes], dtype=np.float32)).cuda()
    vp = torch.from_numpy(np.ascontiguousarray(viewPairs, dtype=np.int32)).cuda()
    w_viewPairs = torch.empty((N_validCubes, N_viewPairs), dtype=torch.float32, device="cuda")
    per_batch = max(1, int(math.floor(float(batchSize) / N_viewPairs)))
    F = 2 * D_embedding + 2
    for b0 in range(0, N_validCubes, per_batch):
        b1 = min(N_validCubes, b0 + per_batch)
        feats = torch.empty(((b1 - b0) * N_viewPairs, F), dtype=torch.float32, device="cuda")
        _lib.check(_lib.lib.sn_viewpair_features(_lib.ptr(e[b0:b1]), _lib.ptr(vp), _lib.ptr(d[b0:b1]), _lib.ptr(theta[b0:b1]),
                                                 1 if theta.dtype == torch.float64 else 0, b1 - b0, N_views, N_viewPairs, D_embedding,
                                                 _lib.ptr(feats), _lib.stream_ptr()))
        w_viewPairs[b0:b1] = viewPair_relativeImpt_fn(feats, n_samples_perGroup=N_viewPairs)
    return __argmaxN_viewPairs__(viewPairs, w_viewPairs.cpu().numpy(), N_viewPairs4inference)
