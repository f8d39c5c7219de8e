"""Drop-in for utils/rayPooling.py:143-260, computed on the GPU (sn_raypool_votes)."""
import numpy as np
from . import _lib


def votes_device(pred, viewPairs, xyz, resol, P_dev, n_views, prediction_thresh, workspace=None):
    """pred: torch.cuda (B,D,D,D) float16|float32; viewPairs (B,N_vp,2) int32 cuda; xyz (B,3) f32, resol (B) f32 cuda.
    -> votes torch.cuda uint8 (B,D,D,D)."""
    torch = _lib.require_cuda()
    B, D = pred.shape[0], pred.shape[-1]
    n_vp = viewPairs.shape[1]
    is16 = pred.dtype == torch.float16
    if prediction_thresh is None:
        has, th = 0, 0.0
    else:
        # numpy compares the float16 / float32 array with the python scalar in the array's dtype
        has, th = 1, float(np.float16(prediction_thresh) if is16 else np.float32(prediction_thresh))
    votes = torch.empty((B, D, D, D), dtype=torch.uint8, device="cuda")
    need = _lib.lib.sn_raypool_workspace_bytes(B, n_vp, D)
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(need, dtype=torch.uint8, device="cuda")
    _lib.check(_lib.lib.sn_raypool_votes(_lib.ptr(pred), int(is16), has, th, _lib.ptr(viewPairs), _lib.ptr(P_dev), n_views,
                                         _lib.ptr(xyz), _lib.ptr(resol), B, n_vp, D, _lib.ptr(votes), _lib.ptr(workspace),
                                         workspace.numel(), _lib.stream_ptr()))
    return votes


def rayPooling_1cube_numpy(cameraPOs, cameraTs, cube_prediction, viewPair_viewIndx, xyz, resol, prediction_thresh=None):
    """utils/rayPooling.py:143.
    cameraPOs (N_views,3,4) float; cameraTs (N_views,3) (only indexed by the reference, unused);
    cube_prediction (D,D,D) float; viewPair_viewIndx (N_viewPair,2) int; xyz (3,), resol scalar;
    prediction_thresh None / scalar.   return: cube_N_votes (D,D,D) int64, max = N_viewPair*2.
    Domain: selected predictions must be > 0 (always true on the hot path, thresh = 0.46)."""
    torch = _lib.require_cuda()
    pred = np.asarray(cube_prediction).squeeze()
    if pred.ndim != 3:
        raise ValueError('rayPooling method argument cube_prediction has {} dims'.format(pred.ndim))
    if pred.dtype != np.float16:
        pred = pred.astype(np.float32)
    cameraPOs = np.ascontiguousarray(cameraPOs, dtype=np.float64)
    vp = np.ascontiguousarray(np.asarray(viewPair_viewIndx).astype(np.int32).reshape(1, -1, 2))
    if vp.min() < 0 or vp.max() >= cameraPOs.shape[0]:
        raise IndexError("viewPair_viewIndx out of range for {} cameras".format(cameraPOs.shape[0]))
    votes = votes_device(torch.from_numpy(np.ascontiguousarray(pred))[None].cuda(), torch.from_numpy(vp).cuda(),
                         torch.from_numpy(np.asarray(xyz, np.float32).reshape(1, 3)).cuda(),
                         torch.from_numpy(np.asarray(resol, np.float32).reshape(1)).cuda(),
                         torch.from_numpy(cameraPOs).cuda(), cameraPOs.shape[0], prediction_thresh)
    return votes[0].cpu().numpy().astype(np.int64)
