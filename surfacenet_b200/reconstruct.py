"""The SurfaceNet-inference section of main_reconstruct.reconstruction (main_reconstruct.py:119-183) on top of the fused hot
path: batches of cubes go through HotPath.infer_batch_sparse (CVC -> SurfaceNet -> fusion -> colours -> ray-pool votes ->
centre crop / threshold / compaction, all on the GPU), the per-cube sparse lists are accumulated exactly as
sparseCubes.append_dense_2sparseList would, thresholded with (tau, gamma) and written in the reference's NPZ schema.

`finish` adds what follows the loop (main_reconstruct.py:168-183): fixed-threshold masks, cross-cube denoising on the GPU, PLY + NPZ.

What comes BEFORE this section in the reference (image / camera loading, early rejection with similarityNet, view-pair
selection) is out of scope: the caller passes the selected view pairs and their weights.
"""
import math
import numpy as np
from . import sparseCubes
from .pipeline import HotPath

PARAM_DTYPE = np.dtype([("xyz", np.float32, (3,)), ("ijk", np.uint32, (3,)), ("resol", np.float32)])      # utils/scene.py:55
CUBE_DCENTER = {32: 26, 64: 52}            # params.py:107
TAU, GAMMA, MIN_PROB = 0.7, 0.8, 0.46      # params.py:66-68


def initialize_cubes(resol, cube_D, cube_Dcenter, cube_overlapping_ratio, BB):
    """Overlapping cube grid over the bounding box BB = [[x_min,x_max],[y_min,y_max],[z_min,z_max]], same layout as
    utils/scene.py:7-61 (stride = Dcenter * resol * overlap, first cube at BB_min - (D - Dcenter) * resol / 2, ijk C-order)."""
    # the reference passes resol = np.float32(0.4) (params.py) under numpy 1.x, where float32_scalar * python_int promotes to
    # float64: stride / margin / origins are float64 arithmetic, cast to float32 only on assignment into the structured array
    BB, resol = np.asarray(BB, dtype=np.float64), float(resol)
    side, centre = resol * cube_D, resol * cube_Dcenter
    stride, margin = centre * cube_overlapping_ratio, (side - centre) / 2
    n_axis = [int(math.ceil(((BB[a][1] + margin) - (BB[a][0] - margin)) / stride)) for a in range(3)]
    ijk = np.indices(tuple(n_axis)).reshape(3, -1).T
    cubes = np.empty(ijk.shape[0], dtype=PARAM_DTYPE)
    cubes["ijk"] = ijk
    cubes["xyz"] = ijk * stride + (BB[:, 0][None, :] - margin)
    cubes["resol"] = resol
    return cubes, side


def quantize_pts_to_cubes(pts_xyz, resol, cube_D, cube_Dcenter, cube_overlapping_ratio, BB=None):
    """Overlapping cubes covering a point cloud, same layout as utils/scene.py:63-107 (quantizePts2Cubes; main_reconstruct.py:57-60 when
    an initial point cloud is given): every point selects the two neighbouring grid cells per axis (floor and floor + 1 of its
    stride coordinate), the distinct cells are the cubes.  -> (cubes_param (N,) PARAM_DTYPE, cube side in mm)"""
    pts_xyz, resol = np.asarray(pts_xyz), float(resol)             # float64 scalar arithmetic as under numpy 1.x (see initialize_cubes)
    side, centre = resol * cube_D, resol * cube_Dcenter
    stride = centre * cube_overlapping_ratio
    if BB is not None:
        BB = np.asarray(BB, dtype=np.float64)
        margin = side / 2
        inBB = np.array([np.logical_and(pts_xyz[:, a] >= (BB[a, 0] - margin), pts_xyz[:, a] <= (BB[a, 1] + margin)) for a in range(3)]).all(axis=0)
        pts_xyz = pts_xyz[inBB]
    shift = pts_xyz.min(axis=0)[None, ...]
    lo = (pts_xyz - shift) // stride
    cells = np.unique(np.vstack([lo, lo + 1]), axis=0)                  # rows sorted lexicographically, like the structured np.unique
    cubes = np.empty(cells.shape[0], dtype=PARAM_DTYPE)
    cubes["ijk"] = cells
    cubes["xyz"] = (cubes["ijk"] * stride + shift) - side / 2
    cubes["resol"] = resol
    return cubes, side


def _concat_batches(batches):
    """per-batch 7-tuples (in batch order) -> the seven items append_dense_2sparseList accumulates, or "Empty!"."""
    lists = ([], [], [], [])
    cube_ijk_np = param_np = viewPair_np = None
    for pl, rl, il, vl, ijk_b, param_b, vp_b in batches:
        for dst, src in zip(lists, (pl, rl, il, vl)):
            dst.extend(src)
        if len(pl):
            cube_ijk_np = ijk_b if cube_ijk_np is None else np.vstack([cube_ijk_np, ijk_b])
            param_np = param_b if param_np is None else np.concatenate([param_np, param_b], axis=0)
            viewPair_np = vp_b if viewPair_np is None else np.vstack([viewPair_np, vp_b])
    if cube_ijk_np is None:
        return "Empty!"
    return lists + (cube_ijk_np, param_np, viewPair_np)


def reconstruct_cubes(hot, cubes_param, viewPairs, w, cube_D, cube_Dcenter=None, batch_size=16, rayPool_thresh=0,
                      rank=0, world_size=1, progress=None, gather=False, group=None):
    """main_reconstruct.py:119-166.  hot: pipeline.HotPath; cubes_param: structured array ('xyz','ijk','resol') of the valid
    cubes; viewPairs (N, N_vp, 2) int; w (N, N_vp) float32 (ignored when N_vp == 1).  With world_size > 1 the batches are
    dealt round-robin to the ranks (cubes are independent).  gather=False: every rank returns ITS cubes only;
    gather=True: the per-batch sparse lists of all ranks are exchanged (torch.distributed.all_gather_object: the sparse
    lists are host objects, a few MB per scene) and EVERY rank returns the whole scene in single-rank batch order -- what the
    cross-cube stages after the loop (denoise_crossCubes, NPZ, adapthresh) need.
    -> (prediction_list, rgb_list, vxl_ijk_list, rayPooling_votes_list, cube_ijk_np, param_np, viewPair_np) or "Empty!"."""
    Dc = int(cube_Dcenter if cube_Dcenter is not None else CUBE_DCENTER.get(int(cube_D), cube_D))
    N = len(cubes_param)
    if N == 0:
        return "Empty!"                                                # main_reconstruct.py:128-129
    viewPairs = np.asarray(viewPairs)
    n_vp = viewPairs.shape[1]
    all_starts = list(range(0, N, batch_size))
    starts = all_starts[rank::world_size]
    mine = []
    for i, b0 in enumerate(starts):
        sel = slice(b0, min(N, b0 + batch_size))
        sp = hot.infer_batch_sparse(viewPairs[sel], cubes_param["xyz"][sel], cubes_param["resol"][sel],
                                    None if n_vp == 1 else np.asarray(w)[sel], cube_D, Dc, rayPool_thresh)
        mine.append(sparseCubes.lists_from_flat(sp, cubes_param[sel], viewPairs[sel], Dc, cube_D))
        if progress is not None:
            progress(i + 1, len(starts))
    if gather and world_size > 1:
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("reconstruct_cubes(gather=True, world_size={}): torch.distributed is not initialised".format(world_size))
        per_rank = [None] * world_size
        dist.all_gather_object(per_rank, mine, group=group)
        mine = [per_rank[i % world_size][i // world_size] for i in range(len(all_starts))]       # back to single-rank batch order
    return _concat_batches(mine)


def finish(result, npz_path=None, ply_path=None, tau=TAU, gamma=GAMMA, cube_D=64):
    """main_reconstruct.py:168-183: the fixed-threshold mask `prediction >= tau & votes >= gamma * N_vp * 2`, the cross-cube
    denoising (GPU, surfacenet_b200.denoising; D_cube = cube_D exactly as the reference passes it at :173), the PLY of the
    denoised cloud and the NPZ file adapthresh / load_sparseCubes read back.  -> (vxl_mask_list, vxl_maskDenoised_list)"""
    from . import denoising
    prediction_list, rgb_list, vxl_ijk_list, votes_list, cube_ijk_np, param_np, viewPair_np = result
    n_vp = viewPair_np.shape[1]
    masks = sparseCubes.filter_voxels(vxl_mask_list=[], prediction_list=prediction_list, prob_thresh=tau,
                                      rayPooling_votes_list=votes_list, rayPool_thresh=gamma * n_vp * 2)
    denoised = denoising.denoise_crossCubes(cube_ijk_np, vxl_ijk_list, vxl_mask_list=masks, D_cube=cube_D)
    if ply_path is not None:
        sparseCubes.save_sparseCubes_2ply(denoised, vxl_ijk_list, rgb_list, param_np, ply_filePath=ply_path, normal_list=None)
    if npz_path is not None:
        sparseCubes.save_sparseCubes(npz_path, *result)
    return masks, denoised


def reconstruction(images_list, cameraPOs_np, BB, resol, N_viewPairs4inference, surfacenet_params, similnet_params, outputFolder=None,
                   cube_D=64, mode="exact", weighted_fusion=True, batch_size=16, min_prob=MIN_PROB, tau=TAU, gamma=GAMMA,
                   cube_overlapping_ratio=0.5, patchSize=64, batchSize_patch2embedding=1024, batchSize_pair=1 << 20, model="model",
                   rank=0, world_size=1, group=None, initial_pts_xyz=None):
    """main_reconstruct.reconstruction (main_reconstruct.py:28-183) from in-memory inputs (the reference reads `images_list`,
    `cameraPOs_np` and BB from files at :49-50 and params.load_modelSpecific_params): cube grid -> early rejection (similarityNet patch
    embeddings, pair dissimilarity) -> view-pair selection -> SurfaceNet inference on the fused sparse path -> fixed-threshold mask,
    cross-cube denoising, PLY + NPZ.  Every stage runs on the GPU drop-ins of this package.
    initial_pts_xyz (N, 3): the cubes cover this point cloud (scene.quantizePts2Cubes, main_reconstruct.py:57-60) instead of the whole
    bounding box; with an outputFolder the cube centres are written to 'initialCubes.ply' as at main_reconstruct.py:61.
    world_size > 1 (one process per GPU, torch.distributed initialised): the cube batches are dealt to the ranks, the sparse lists are
    gathered so that every rank holds the whole scene before the cross-cube stages, and ONLY rank 0 writes the PLY / NPZ files.
    -> "Empty!" or dict(npz_path, ply_path, result=(the seven sparse-list items), vxl_mask_list, vxl_maskDenoised_list, validCubes,
                        viewPairs4Reconstr, w_viewPairs4Reconstr)"""
    import os
    from . import SurfaceNet, camera, earlyRejection, similarityNet, viewPairSelection
    from .device import DeviceScene
    from .utils import k_combination_np
    from .weights import MEAN_PATCHES_BGR
    N = int(N_viewPairs4inference)
    Dc = CUBE_DCENTER.get(int(cube_D), int(cube_D))
    cameraPOs_np = np.asarray(cameraPOs_np, dtype=np.float64)
    cameraTs_np = camera.cameraPs2Ts(cameraPOs_np)                                                        # :51
    if initial_pts_xyz is None:
        cubes_param_np, cube_D_mm = initialize_cubes(resol, cube_D, Dc, cube_overlapping_ratio, BB)        # :53-55
    else:
        cubes_param_np, cube_D_mm = quantize_pts_to_cubes(initial_pts_xyz, resol, cube_D, Dc, cube_overlapping_ratio, BB)   # :57-60
    if outputFolder is not None and rank == 0:
        sparseCubes.save2ply(os.path.join(outputFolder, 'initialCubes.ply'), xyz_np=cubes_param_np['xyz'] + cube_D_mm / 2)    # :61
    img_h_corner, img_w_corner = camera.perspectiveProj_cubesCorner(cameraPOs_np, cubes_param_np['xyz'], cube_D_mm, return_int_hw=False)
    centers = cubes_param_np['xyz'] + cube_D_mm / 2
    img_h_center, img_w_center = camera.perspectiveProj(cameraPOs_np, centers, return_int_hw=False)        # :63-65
    N_views, N_cubes = img_h_corner.shape[:2]
    patch2embedding_fn, embeddingPair2simil_fn = similarityNet.similarityNet_inference(similnet_params, (patchSize, patchSize))   # :70
    viewPair_relativeImpt_fn, nViewPair_SurfaceNet_fn = SurfaceNet.SurfaceNet_inference(N, surfacenet_params, mode=mode)       # :72
    viewPairs = k_combination_np(range(N_views), k=2)                                                       # :80
    emb, inScope = earlyRejection.patch2embedding(images_list, img_h_corner, img_w_corner, patch2embedding_fn, MEAN_PATCHES_BGR, N_cubes,
                                                  N_views, similarityNet.D_EMBEDDING, patchSize=patchSize, batchSize=batchSize_patch2embedding,
                                                  cubeCenter_hw=np.stack([img_h_center, img_w_center], axis=0))            # :83-88
    dissimilarity = earlyRejection.embeddingPairs2simil(emb, N_views, inScope, embeddingPair2simil_fn, batchSize_pair, viewPairs)   # :90-95
    validCubes = earlyRejection.selectFromSimilarity(dissimilarity, N)                                      # :96
    if int(validCubes.sum()) == 0:
        return "Empty!"
    pairs, w = viewPairSelection.viewPairSelection(cameraTs_np, emb, dissimilarity, validCubes, centers, viewPair_relativeImpt_fn,
                                                   batchSize_pair, N, viewPairs)                            # :104-113
    w = np.ascontiguousarray(w, dtype=np.float32)
    if not weighted_fusion:
        w[:] = 1.0 / N                                                                                      # :115-116
    hot = HotPath(nViewPair_SurfaceNet_fn.net, DeviceScene(cameraPOs_np, images_list), mode=mode, min_prob=min_prob)
    res = reconstruct_cubes(hot, cubes_param_np[validCubes], pairs, w, cube_D, Dc, batch_size=batch_size, rank=rank, world_size=world_size,
                            gather=True, group=group)
    if isinstance(res, str):
        return res
    npz = ply = None
    if outputFolder is not None:
        os.makedirs(outputFolder, exist_ok=True)
        ply = os.path.join(outputFolder, 'fixThresh_tau{:.3}_gamma{:.3}.ply'.format(tau, gamma))           # :169
        npz = os.path.join(outputFolder, 'model{}-{}views.npz'.format(model, N_views))                      # :180
    write = rank == 0                                                   # every rank holds the whole scene; one writer
    masks, denoised = finish(res, npz_path=npz if write else None, ply_path=ply if write else None, tau=tau, gamma=gamma, cube_D=cube_D)
    return dict(npz_path=npz, ply_path=ply, result=res, vxl_mask_list=masks, vxl_maskDenoised_list=denoised, validCubes=validCubes,
                viewPairs4Reconstr=pairs, w_viewPairs4Reconstr=w)
