"""Drop-in for utils/adapthresh.py:91-178: the adaptive per-cube threshold refinement on the GPU (sn_sparse_filter_voxels,
sn_sparse_adapthresh, sn_sparse_denoise in csrc/postprocess.cu); only the PLY files are written on the host.  "Next" row N4.

The reference's helpers access_partial_Occupancy_ijk / sparseOccupancy_AND_XOR (adapthresh.py:11-86) have no stand-alone
counterpart: the half-cube selection and the AND / XOR counts are fused into one counting kernel (pp_ada_count_kernel)."""
import copy
import os
import numpy as np
from . import sparseCubes
from .sparse_device import DeviceSparseCubes


def adapthresh_lists(prediction_list, vxl_ijk_list, rayPooling_votes_list, cube_ijk_np, N_refine_iter, D_cube, init_probThresh,
                     max_probThresh, rayPool_thresh, beta, on_iteration=None, denoise_each_iter=True, on_init=None):
    """The computation of adapthresh.py:101-176 without file I/O.
    -> dict(init_mask, init_denoised, probThresh (C,) f64, mask, denoised, argmin (n_iter, C)); lists of per-cube bool arrays.
    on_iteration(i, mask_list, denoised_list | None, argmin (C,)) is called after every iteration (the reference writes a PLY there)."""
    dsc = DeviceSparseCubes(cube_ijk_np, vxl_ijk_list, prediction_list, rayPooling_votes_list)
    torch = dsc.torch
    init_mask = dsc.filter_voxels(None, prob_thresh=init_probThresh, rayPool_thresh=rayPool_thresh)              # :103
    out = dict(init_mask=dsc.split(init_mask, bool))
    out["init_denoised"] = dsc.split(dsc.denoise(init_mask, D_cube)["keep"], bool)                                # :110
    if on_init is not None:
        on_init(out["init_denoised"])
    mask = init_mask.clone()
    thresh = torch.full((dsc.C,), float(init_probThresh), dtype=torch.float64, device="cuda")
    argmins, denoised = [], None
    for it in range(N_refine_iter):
        arg = dsc.adapthresh(init_mask, mask, thresh, D_cube, max_probThresh, beta, n_iter=1, want_argmin=True)   # :131-174
        argmins.append(arg[0].cpu().numpy())
        last = it == N_refine_iter - 1
        if denoise_each_iter or last:
            denoised = dsc.split(dsc.denoise(mask, D_cube)["keep"], bool)                                         # :176
        if on_iteration is not None:
            on_iteration(it, dsc.split(mask, bool), denoised if (denoise_each_iter or last) else None, argmins[-1])
    out.update(probThresh=thresh.cpu().numpy(), mask=dsc.split(mask, bool), denoised=denoised,
               argmin=np.stack(argmins) if argmins else np.zeros((0, dsc.C), np.int32))
    return out


def adapthresh(save_result_fld, N_refine_iter, D_cube, init_probThresh, min_probThresh, max_probThresh, rayPool_thresh, beta, gamma,
               npz_file, RGB_visual_ply=True):
    """Same contract as utils/adapthresh.py:91: reads the NPZ of sparse cubes, writes initialization.ply and iter{i}.ply
    (and iter{i}_tmprgb4debug.ply when RGB_visual_ply) under save_result_fld/adapThresh_gamma{gamma}_beta{beta}, returns the
    path of the last PLY.  (min_probThresh is accepted and unused, as in the reference.)"""
    prediction_list, rgb_list, vxl_ijk_list, rayPooling_votes_list, cube_ijk_np, param_np, viewPair_np = \
        sparseCubes.load_sparseCubes(npz_file)
    save_result_fld = os.path.join(save_result_fld, "adapThresh_gamma{:.3}_beta{}".format(gamma, beta))
    if not os.path.exists(save_result_fld):
        os.makedirs(save_result_fld)
    state = dict(path=None)

    def write_iter(i, mask_list, denoised_list, argmin):
        state["path"] = os.path.join(save_result_fld, 'iter{}.ply'.format(i))
        sparseCubes.save_sparseCubes_2ply(denoised_list, vxl_ijk_list, rgb_list, param_np, ply_filePath=state["path"], normal_list=None)
        if RGB_visual_ply:
            tmp_rgb_list = copy.deepcopy(rgb_list)
            for c, a in enumerate(argmin):
                if a >= 0:
                    tmp_rgb_list[c][:, a] = 255                                  # adapthresh.py:171: R/G/B <-> chosen perturbation
            sparseCubes.save_sparseCubes_2ply(mask_list, vxl_ijk_list, tmp_rgb_list, param_np, normal_list=None,
                                              ply_filePath=os.path.join(save_result_fld, 'iter{}_tmprgb4debug.ply'.format(i)))

    def write_init(denoised_list):                                               # adapthresh.py:110-112
        sparseCubes.save_sparseCubes_2ply(denoised_list, vxl_ijk_list, rgb_list, param_np,
                                          ply_filePath=os.path.join(save_result_fld, 'initialization.ply'), normal_list=None)

    adapthresh_lists(prediction_list, vxl_ijk_list, rayPooling_votes_list, cube_ijk_np, N_refine_iter, D_cube, init_probThresh,
                     max_probThresh, rayPool_thresh, beta, on_iteration=write_iter, on_init=write_init)
    return state["path"]
