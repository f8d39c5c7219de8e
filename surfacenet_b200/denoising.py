"""Drop-in for utils/denoising.py: per-cube clustering, cross-cube overlap marking and denoise_crossCubes on the GPU
(sn_sparse_denoise, csrc/postprocess.cu).  "Next" row N4."""
import numpy as np
from .sparse_device import DeviceSparseCubes

dtype_clusterLabel = np.uint32                                                   # denoising.py:5


def _labels_lists(dsc, out, vxl_mask_list):
    labels = dsc.split(out["labels"], dtype_clusterLabel)
    n_labels = [int(x) for x in out["n_labels"].cpu().numpy()]
    for c, m in enumerate(vxl_mask_list):
        if np.asarray(m).sum() == 0:
            labels[c] = np.zeros(labels[c].shape[:1])                            # denoising.py:45: float zeros for an empty cube
    return labels, n_labels


def __cluster_inCube__(vxl_ijk_list, vxl_mask_list=[], neighbor_dist=1):
    """utils/denoising.py:8-62 -> (vxl_labeles_list, N_labels_list); labels numbered like scipy.ndimage.label, 0 = masked out."""
    C = len(vxl_mask_list)
    dsc = DeviceSparseCubes(np.zeros((C, 3), np.int32), list(vxl_ijk_list)[:C])
    out = dsc.denoise(dsc.upload_mask(vxl_mask_list), D_cube=0, neighbor_dist=neighbor_dist, want_keep=False, want_labels=True)
    return _labels_lists(dsc, out, vxl_mask_list)


def __mark_overlappingLabels__(cube_ijk_np, vxl_ijk_list, vxl_mask_list, D_cube, neighbor_dist=1):
    """utils/denoising.py:67-140 -> (overlappingLabels_list, vxl_labeles_list).  The label lists come back sorted (the
    reference returns list(set(...)), whose order is arbitrary)."""
    dsc = DeviceSparseCubes(cube_ijk_np, vxl_ijk_list)
    out = dsc.denoise(dsc.upload_mask(vxl_mask_list), D_cube=D_cube, neighbor_dist=neighbor_dist, want_keep=True, want_labels=True)
    labels, _ = _labels_lists(dsc, out, vxl_mask_list)
    keep = dsc.split(out["keep"], bool)
    overlapping = [sorted(set(int(x) for x in l[k])) for l, k in zip(labels, keep)]
    return overlapping, labels


def denoise_crossCubes(cube_ijk_np, vxl_ijk_list, vxl_mask_list, D_cube):
    """utils/denoising.py:145-184 -> vxl_maskDenoise_list: only voxels whose 26-connected cluster overlaps a neighbouring cube."""
    dsc = DeviceSparseCubes(cube_ijk_np, vxl_ijk_list)
    out = dsc.denoise(dsc.upload_mask(vxl_mask_list), D_cube=D_cube, neighbor_dist=3)
    return dsc.split(out["keep"], bool)
