"""Drop-in for utils/sparseCubes.py:9-141 (dense2sparse, append_dense_2sparseList): ray pooling, centre crop,
thresholding and the ordered compaction run on the GPU (sn_raypool_votes, sn_dense2sparse).  "Next" row N1."""
import numpy as np
from . import _lib, rayPooling


def dense2sparse_device(pred16, rgb, votes, D, Dcenter, min_prob, rayPool_thresh=0):
    """pred16 torch.cuda (B,D,D,D) f16; rgb (B,3,D,D,D) u8 | None; votes (B,D,D,D) u8 | None.
    -> dict(counts i32 (B,), offsets i32 (B+1,), ijk u8 (T,3), pred f16 (T,), rgb u8 (T,3)|None, votes u8 (T,)|None) on the
    host, T = total kept voxels; only the compacted entries cross PCIe."""
    torch = _lib.require_cuda()
    B = int(pred16.shape[0])
    cap = B * Dcenter ** 3
    counts = torch.zeros(B, dtype=torch.int32, device="cuda")
    offsets = torch.zeros(B + 1, dtype=torch.int32, device="cuda")
    ijk = torch.empty((cap, 3), dtype=torch.uint8, device="cuda")
    pred_o = torch.empty(cap, dtype=torch.float16, device="cuda")
    rgb_o = torch.empty((cap, 3), dtype=torch.uint8, device="cuda") if rgb is not None else None
    votes_o = torch.empty(cap, dtype=torch.uint8, device="cuda") if votes is not None else None
    need = _lib.lib.sn_dense2sparse_workspace_bytes(B, D, Dcenter)
    if need < 0:
        raise ValueError("dense2sparse: bad sizes (N_cubes={}, D={}, cube_Dcenter={})".format(B, D, Dcenter))
    ws = torch.empty(int(need), dtype=torch.uint8, device="cuda")
    mp = float(np.float16(min_prob))                 # float16 prediction compared with the python float in float16
    _lib.check(_lib.lib.sn_dense2sparse(_lib.ptr(pred16), _lib.ptr(rgb), _lib.ptr(votes), B, D, Dcenter, mp, int(rayPool_thresh),
                                        _lib.ptr(counts), _lib.ptr(offsets), _lib.ptr(ijk), _lib.ptr(pred_o), _lib.ptr(rgb_o),
                                        _lib.ptr(votes_o), cap, _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
    off = offsets.cpu().numpy()
    T = int(off[-1]) if B else 0
    return dict(counts=counts.cpu().numpy(), offsets=off, ijk=ijk[:T].cpu().numpy(), pred=pred_o[:T].cpu().numpy(),
                rgb=None if rgb_o is None else rgb_o[:T].cpu().numpy(), votes=None if votes_o is None else votes_o[:T].cpu().numpy())


def _split(flat, offsets, idx):
    return [flat[offsets[n]:offsets[n + 1]] for n in idx]


def dense2sparse(prediction, rgb, param, viewPair, min_prob=0.5, rayPool_thresh=0, enable_centerCrop=False, cube_Dcenter=None,
                 enable_rayPooling=False, cameraPOs=None, cameraTs=None):
    """utils/sparseCubes.py:9.
    prediction float16 (N_cubes,D,D,D); rgb uint8 (N_cubes,D,D,D,3); param structured 'ijk'/'xyz'/'resol'; viewPair (N_cubes,N_vp,2).
    -> nonempty_cube_indx, vxl_ijk_list, prediction_list, rgb_list, rayPooling_votes_list, param_new"""
    torch = _lib.require_cuda()
    prediction = np.asarray(prediction)
    if prediction.ndim != 4:
        raise ValueError("prediction must have shape (N_cubes, D, D, D), got {}".format(prediction.shape))
    N_cubes, D = prediction.shape[:2]
    Dc = int(cube_Dcenter) if enable_centerCrop else D
    param_new = np.copy(param)
    if enable_centerCrop:
        param_new['xyz'] += param_new['resol'][:, None] * ((D - Dc) // 2)              # sparseCubes.py:55
    if N_cubes == 0:
        return [], [], [], [], [], param_new
    pred16 = torch.from_numpy(np.ascontiguousarray(prediction.astype(np.float16))).cuda()
    rgb_d = torch.from_numpy(np.ascontiguousarray(np.transpose(np.asarray(rgb).astype(np.uint8), (0, 4, 1, 2, 3)))).cuda()
    votes_d = None
    if enable_rayPooling:
        P = torch.from_numpy(np.ascontiguousarray(cameraPOs, dtype=np.float64)).cuda()
        vp = torch.from_numpy(np.ascontiguousarray(np.asarray(viewPair).astype(np.int32))).cuda()
        xyz = torch.from_numpy(np.ascontiguousarray(param['xyz'], dtype=np.float32)).cuda()
        resol = torch.from_numpy(np.ascontiguousarray(param['resol'], dtype=np.float32)).cuda()
        votes_d = rayPooling.votes_device(pred16, vp, xyz, resol, P, P.shape[0], min_prob)   # sparseCubes.py:60-62
    out = dense2sparse_device(pred16, rgb_d, votes_d, D, Dc, min_prob, rayPool_thresh if enable_rayPooling else 0)
    idx = [n for n in range(N_cubes) if out["counts"][n] > 0]                               # sparseCubes.py:67-68
    off = out["offsets"]
    votes_l = _split(out["votes"], off, idx) if enable_rayPooling else []
    return idx, _split(out["ijk"], off, idx), _split(out["pred"], off, idx), _split(out["rgb"], off, idx), votes_l, param_new


def append_dense_2sparseList(prediction_sub, rgb_sub, param_sub, viewPair_sub, min_prob=0.5, rayPool_thresh=0,
                             enable_centerCrop=False, cube_Dcenter=None, enable_rayPooling=False, cameraPOs=None, cameraTs=None,
                             prediction_list=[], rgb_list=[], vxl_ijk_list=[], rayPooling_votes_list=[],
                             cube_ijk_np=None, param_np=None, viewPair_np=None):
    """utils/sparseCubes.py:82 (same mutable-default signature; main_reconstruct.py always passes the lists)."""
    prediction_sub = np.asarray(prediction_sub)
    if prediction_sub.ndim == 5:
        prediction_sub = prediction_sub.astype(np.float16)[:, 0]                           # sparseCubes.py:115
    rgb_sub = np.transpose(np.asarray(rgb_sub).astype(np.uint8), axes=(0, 2, 3, 4, 1))     # sparseCubes.py:116
    cube_ijk_sub = param_sub['ijk']
    viewPair_sub = np.asarray(viewPair_sub).astype(np.uint16)
    idx, ijk_l, pred_l, rgb_l, votes_l, param_new_sub = dense2sparse(
        prediction=prediction_sub, rgb=rgb_sub, param=param_sub, viewPair=viewPair_sub, min_prob=min_prob, rayPool_thresh=rayPool_thresh,
        enable_centerCrop=enable_centerCrop, cube_Dcenter=cube_Dcenter, enable_rayPooling=enable_rayPooling, cameraPOs=cameraPOs, cameraTs=cameraTs)
    param_sub = param_new_sub[idx]
    viewPair_sub = viewPair_sub[idx]
    cube_ijk_sub = cube_ijk_sub[idx]
    if not len(pred_l) == len(rgb_l) == len(ijk_l) == param_sub.shape[0] == viewPair_sub.shape[0] == cube_ijk_sub.shape[0]:
        raise Warning('load dense data, # of cubes is not consistent.')                   # sparseCubes.py:131
    prediction_list.extend(pred_l); rgb_list.extend(rgb_l); vxl_ijk_list.extend(ijk_l); rayPooling_votes_list.extend(votes_l)
    param_np = param_sub if param_np is None else np.concatenate([param_np, param_sub], axis=0)
    viewPair_np = viewPair_sub if viewPair_np is None else np.vstack([viewPair_np, viewPair_sub])
    cube_ijk_np = cube_ijk_sub if cube_ijk_np is None else np.vstack([cube_ijk_np, cube_ijk_sub])
    return prediction_list, rgb_list, vxl_ijk_list, rayPooling_votes_list, cube_ijk_np, param_np, viewPair_np
