"""Oracle restatement of utils/denoising.py:8-184 (per-cube clustering, cross-cube overlap marking, denoise_crossCubes)
and utils/adapthresh.py:11-178 (partial-cube access, sparse AND/XOR, the adaptive-threshold refinement loop).
Test infrastructure only ("next" row N4 of SURVEY.md 8(f)).

Pinned by (tests/test_oracle_golden.py):
  * the reference's own doctest known answers (denoising.py:26-38,82-94,160-172; adapthresh.py:33-41,73-76), and
  * golden vectors produced by executing the reference code itself (tests/golden/make_golden_post.py).

Mechanical py3 / numpy-2 restatements (the reference is python-2 / numpy-1.13 code):
  * dict.has_key(k) -> k in dict;  `3**3/2`, `D_cube / 2` are python-2 integer divisions -> `//`
  * adapthresh.py:136,154-157: `element_cost` is a float16 array and `element_cost[i] += python_int` is evaluated by
    numpy 1.13 as float64(element_cost[i]) + int, rounded back to float16 on assignment (scalar + scalar promotes
    float16 with int64 to float64).  numpy 2 (NEP 50) would first round the integer to float16; both agree while the
    counts stay <= 2048 (exactly representable), which the golden cases guarantee.  Restated here explicitly as
    float16(float64(cost) + count): PARITY UNPINNED for counts above 2048 (numpy 1.13 cannot run in this image).
  * the function returns what the reference only writes to PLY files: per-iteration thresholds, masks, denoised masks.
"""
import copy
import numpy as np
import scipy.ndimage as ndim

dtype_clusterLabel = np.uint32                                                                     # denoising.py:5


def cluster_inCube(vxl_ijk_list, vxl_mask_list=[], neighbor_dist=1):
    """denoising.py:8-62 __cluster_inCube__ -> (vxl_labeles_list, N_labels_list)."""
    vxl_labeles_list, N_labels_list = [], []
    for _cube, _select in enumerate(vxl_mask_list):
        N_pts = _select.sum()
        vxl_ijk = vxl_ijk_list[_cube]
        if N_pts == 0:
            vxl_labeles_list.append(np.zeros(vxl_ijk.shape[:1]))                                   # :45 (float zeros)
            N_labels_list.append(0)
            continue
        vxl_ijk_masked = vxl_ijk[_select]
        matrix3D = np.zeros(vxl_ijk_masked.max(axis=0).astype(np.int64) + 1, dtype=bool)           # :51
        matrix3D[vxl_ijk_masked[:, 0], vxl_ijk_masked[:, 1], vxl_ijk_masked[:, 2]] = 1
        binary_structure = ndim.generate_binary_structure(3, neighbor_dist)
        labeled_3Darray, N_labels = ndim.label(matrix3D, binary_structure)                         # :54
        _vxl_labeles = np.zeros(_select.shape, dtype=dtype_clusterLabel)
        _vxl_labeles[_select] = labeled_3Darray[vxl_ijk_masked[:, 0], vxl_ijk_masked[:, 1], vxl_ijk_masked[:, 2]].astype(dtype_clusterLabel)
        vxl_labeles_list.append(_vxl_labeles)
        N_labels_list.append(N_labels)
    return vxl_labeles_list, N_labels_list


def _view1d(a):
    a = np.ascontiguousarray(a)
    return a.view(dtype=a.dtype.descr * 3)


def mark_overlappingLabels(cube_ijk_np, vxl_ijk_list, vxl_mask_list, D_cube, neighbor_dist=1):
    """denoising.py:67-140 __mark_overlappingLabels__ -> (overlappingLabels_list, vxl_labeles_list).
    The label lists are returned sorted (the reference returns list(set(...)) in arbitrary order)."""
    vxl_labeles_list, N_labels_list = cluster_inCube(vxl_ijk_list, vxl_mask_list=vxl_mask_list, neighbor_dist=neighbor_dist)
    N_cubes = len(N_labels_list)
    overlappingLabels_list = [[] for _ in range(N_cubes)]
    cube_ijk2index = {}
    neigh_shifts = np.delete(np.array(np.indices((3, 3, 3))).reshape((3, -1)).T - 1, 3 ** 3 // 2, axis=0)   # :103
    for _n, _ijk in enumerate(cube_ijk_np):
        if vxl_mask_list[_n].sum() > 0:                                                            # :107
            cube_ijk2index.update({tuple(int(x) for x in _ijk): _n})
    for _n, _ijk in enumerate(cube_ijk_np):
        key = tuple(int(x) for x in _ijk)
        if key not in cube_ijk2index:
            continue
        i_current = cube_ijk2index[key]
        _vxl_mask_current = vxl_mask_list[i_current]
        vxl_ijk_current = vxl_ijk_list[i_current][_vxl_mask_current].astype(np.int64)
        view1d_current = _view1d(vxl_ijk_current)
        for _ijk_shift in neigh_shifts:
            ijk_neigh = tuple(int(x) for x in (np.asarray(_ijk).astype(np.int64) + _ijk_shift))
            if ijk_neigh not in cube_ijk2index:
                continue
            i_neigh = cube_ijk2index[ijk_neigh]
            _vxl_mask_neigh = vxl_mask_list[i_neigh]
            vxl_ijk_neigh = vxl_ijk_list[i_neigh][_vxl_mask_neigh].astype(np.int64)
            vxl_ijk_newCoords_neigh = (vxl_ijk_neigh + (D_cube // 2) * _ijk_shift).astype(np.int64)     # :127
            view1d_neigh = _view1d(vxl_ijk_newCoords_neigh)
            view1d_intersect = np.intersect1d(view1d_current, view1d_neigh)
            if view1d_intersect.size != 0:
                intersectBool_current = np.isin(view1d_current, view1d_intersect).ravel()
                intersectBool_neigh = np.isin(view1d_neigh, view1d_intersect).ravel()
                overlappingLabels_current = vxl_labeles_list[i_current][_vxl_mask_current][intersectBool_current]
                overlappingLabels_neigh = vxl_labeles_list[i_neigh][_vxl_mask_neigh][intersectBool_neigh]
                overlappingLabels_list[i_current] = sorted(set(overlappingLabels_list[i_current] + [int(x) for x in overlappingLabels_current]))
                overlappingLabels_list[i_neigh] = sorted(set(overlappingLabels_list[i_neigh] + [int(x) for x in overlappingLabels_neigh]))
    return overlappingLabels_list, vxl_labeles_list


def denoise_crossCubes(cube_ijk_np, vxl_ijk_list, vxl_mask_list, D_cube):
    """denoising.py:145-184 -> vxl_maskDenoise_list (keep only voxels whose 26-connected cluster overlaps a neighbouring cube)."""
    vxl_maskDenoise_list = []
    overlappingLabels_list, vxl_labeles_list = mark_overlappingLabels(cube_ijk_np, vxl_ijk_list, vxl_mask_list, D_cube, neighbor_dist=3)
    for _cube, _vxl_labels in enumerate(vxl_labeles_list):
        vxl_maskDenoise_list.append(np.isin(_vxl_labels, overlappingLabels_list[_cube]))          # :181
    return vxl_maskDenoise_list


def access_partial_Occupancy_ijk(Occ_ijk, shift, D_cube):
    """adapthresh.py:11-55.  Changes Occ_ijk in place like the reference."""
    D_mid = D_cube // 2                                                                            # :43 (py2 int division)
    N_voxel, n_dim = Occ_ijk.shape
    select_ijk = np.ones(shape=(N_voxel, n_dim))
    for _dim in range(n_dim):
        if shift[_dim] == -1:
            select_ijk[:, _dim] = (Occ_ijk[:, _dim] >= 0) & (Occ_ijk[:, _dim] < D_mid)
        elif shift[_dim] == 0:
            select_ijk[:, _dim] = (Occ_ijk[:, _dim] >= 0) & (Occ_ijk[:, _dim] < D_cube)
        elif shift[_dim] == 1:
            select_ijk[:, _dim] = (Occ_ijk[:, _dim] >= D_mid) & (Occ_ijk[:, _dim] < D_cube)
            Occ_ijk[:, _dim] -= np.asarray(D_mid).astype(Occ_ijk.dtype)                            # :51 (uint8 wrap-around kept)
        else:
            raise Warning("shift only support 3 values: -1/0/1, but got {}".format(shift))
    select_ijk = select_ijk.all(axis=1)
    return Occ_ijk[select_ijk]


def sparseOccupancy_AND_XOR(Occ1, Occ2):
    """adapthresh.py:57-86 -> (resul_AND, resul_XOR)."""
    n1, ndim1 = Occ1.shape
    n2, ndim2 = Occ2.shape
    if (n1 == 0) or (n2 == 0):
        resul_AND = 0
    else:
        Occ1_1D = np.ascontiguousarray(Occ1).view(dtype=Occ1.dtype.descr * ndim1)
        Occ2_1D = np.ascontiguousarray(Occ2).view(dtype=Occ2.dtype.descr * ndim2)
        resul_AND = np.intersect1d(Occ1_1D, Occ2_1D).size
    resul_XOR = n1 + n2 - resul_AND * 2
    return resul_AND, resul_XOR


def filter_voxels(vxl_mask_list=None, prediction_list=None, prob_thresh=None, rayPooling_votes_list=None, rayPool_thresh=None):
    """utils/sparseCubes.py:205-243 (float16 predictions are compared with the python-float threshold in float16)."""
    vxl_mask_list = [] if vxl_mask_list is None else vxl_mask_list
    empty = len(vxl_mask_list) == 0
    if prediction_list is not None:
        if prob_thresh is None:
            raise Warning('prob_thresh should not be None.')
        for _c, _prediction in enumerate(prediction_list):
            _prob_thresh = prob_thresh[_c] if isinstance(prob_thresh, list) else prob_thresh
            _surf = _prediction >= np.float16(_prob_thresh)
            if empty:
                vxl_mask_list.append(_surf)
            else:
                vxl_mask_list[_c] &= _surf
    empty = len(vxl_mask_list) == 0
    if rayPooling_votes_list is not None:
        if rayPool_thresh is None:
            raise Warning('rayPool_thresh should not be None.')
        for _cube, _votes in enumerate(rayPooling_votes_list):
            _surf = _votes >= rayPool_thresh
            if empty:
                vxl_mask_list.append(_surf)
            else:
                vxl_mask_list[_cube] &= _surf
    return vxl_mask_list


def adapthresh_core(prediction_list, vxl_ijk_list, rayPooling_votes_list, cube_ijk_np, N_refine_iter, D_cube,
                    init_probThresh, min_probThresh, max_probThresh, rayPool_thresh, beta, denoise_each_iter=True):
    """adapthresh.py:91-178 without the file I/O.
    -> dict(init_mask, init_denoised, iters=[dict(probThresh, argmin, mask, denoised)])"""
    vxl_mask_init_list = filter_voxels([], prediction_list, init_probThresh, rayPooling_votes_list, rayPool_thresh)     # :103
    out = dict(init_mask=[m.copy() for m in vxl_mask_init_list],
               init_denoised=denoise_crossCubes(cube_ijk_np, vxl_ijk_list, vxl_mask_init_list, D_cube), iters=[])       # :110
    neigh_shifts = np.asarray([[1, 0, 0], [0, 1, 0], [0, 0, 1], [-1, 0, 0], [0, -1, 0], [0, 0, -1]]).astype(np.int8)    # :114
    thresh_perturb_list = [0.1, 0, -0.1]                                                                                # :115
    cube_ijk2indx = {}
    for _n, _ijk in enumerate(cube_ijk_np):
        if vxl_mask_init_list[_n].sum() > 0:                                                                            # :123
            cube_ijk2indx.update({tuple(int(x) for x in _ijk): _n})
    vxl_mask_list = copy.deepcopy(vxl_mask_init_list)
    probThresh_list = [init_probThresh] * len(prediction_list)
    update_probThresh_list = copy.deepcopy(probThresh_list)

    def occupied_vxl(indx, thresh_shift):                                                                               # :128-129
        return vxl_mask_list[indx] & (prediction_list[indx] >= np.float16(probThresh_list[indx] + thresh_shift))

    for _iter in range(N_refine_iter):
        argmin_list = np.full(len(prediction_list), -1, np.int32)
        for _ijk in cube_ijk_np:
            key = tuple(int(x) for x in _ijk)
            if key not in cube_ijk2indx:
                continue
            i_current = cube_ijk2indx[key]
            element_cost = np.array([0, 0, 0]).astype(np.float16)                                                       # :141
            for _ijk_shift in neigh_shifts:
                ijk_ovlp = tuple(int(x) for x in (np.asarray(_ijk).astype(np.int64) + _ijk_shift))
                if ijk_ovlp in cube_ijk2indx:
                    i_ovlp = cube_ijk2indx[ijk_ovlp]
                    tmp_occupancy_ovlp = vxl_ijk_list[i_ovlp][occupied_vxl(i_ovlp, 0)]
                    partial_occ_ovlp = access_partial_Occupancy_ijk(tmp_occupancy_ovlp, shift=_ijk_shift * -1, D_cube=D_cube)
                else:
                    partial_occ_ovlp = np.empty((0, 3), dtype=np.uint8)
                for _n_thresh, _thresh_perturb in enumerate(thresh_perturb_list):
                    tmp_occupancy_current = vxl_ijk_list[i_current][occupied_vxl(i_current, _thresh_perturb)]
                    partial_occ_current = access_partial_Occupancy_ijk(tmp_occupancy_current, shift=_ijk_shift, D_cube=D_cube)
                    ovlp_AND, ovlp_XOR = sparseOccupancy_AND_XOR(partial_occ_current, partial_occ_ovlp)
                    element_cost[_n_thresh] = np.float16(np.float64(element_cost[_n_thresh]) + ovlp_XOR)                # :162
                    if partial_occ_current.shape[0] >= 6:
                        if partial_occ_ovlp.shape[0] >= 6:
                            element_cost[_n_thresh] = np.float16(np.float64(element_cost[_n_thresh]) - beta * ovlp_AND)  # :165
            update_probThresh_list[i_current] = probThresh_list[i_current] + thresh_perturb_list[int(np.argmin(element_cost))]
            update_probThresh_list[i_current] = min(update_probThresh_list[i_current], max_probThresh)                  # :168
            argmin_list[i_current] = int(np.argmin(element_cost))
        probThresh_list = copy.deepcopy(update_probThresh_list)
        vxl_mask_list = filter_voxels(vxl_mask_list, prediction_list, probThresh_list, None, None)                      # :174
        it = dict(probThresh=np.asarray(probThresh_list, np.float64), argmin=argmin_list, mask=[m.copy() for m in vxl_mask_list])
        if denoise_each_iter:
            it["denoised"] = denoise_crossCubes(cube_ijk_np, vxl_ijk_list, vxl_mask_list, D_cube)                       # :176
        out["iters"].append(it)
    return out
