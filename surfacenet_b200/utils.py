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


# ---- host-side batching helpers of utils/utils.py:45-230 (index bookkeeping used by main_reconstruct.py:123 and the callers of
# the similarityNet functions); same names, same results, written with numpy ranges instead of python-2 list ranges -----------
def gen_batch_index(N_all, batch_size):
    """utils/utils.py:45-74: index lists of consecutive batches, the last one possibly shorter."""
    return [list(range(s, int(min(s + batch_size, N_all)))) for s in range(0, int(N_all), int(batch_size))]


def gen_batch_npBool(N_all, batch_size):
    """utils/utils.py:113-146 -> (N_batches, N_all) bool selectors."""
    starts = np.arange(0, int(N_all), int(batch_size))
    idx = np.arange(int(N_all))[None, :]
    return (idx >= starts[:, None]) & (idx < starts[:, None] + int(batch_size))


def yield_batch_npBool(N_all, batch_size):
    """utils/utils.py:149-178: the rows of gen_batch_npBool one at a time."""
    for row in gen_batch_npBool(N_all, batch_size):
        yield row


def gen_non0Batch_npBool(boolIndicators, batch_size):
    """utils/utils.py:77-110: selectors over ALL elements that pick, batch by batch, the elements whose indicator is True."""
    boolIndicators = np.asarray(boolIndicators).astype(bool)
    order = np.cumsum(boolIndicators)                                   # 1-based rank of every True element
    n_true = int(boolIndicators.sum())
    sel = [(order >= s + 1) & (order <= min(s + int(batch_size), n_true)) & boolIndicators for s in range(0, n_true, int(batch_size))]
    return np.array(sel) if sel else np.zeros((0, boolIndicators.size), dtype=bool)


def yield_batch_ij_npBool(ij_lists, batch_size):
    """utils/utils.py:181-230: batches of (i, j) index pairs enumerating ij_lists[0] x ij_lists[1] in row-major order."""
    a, b = np.asarray(list(ij_lists[0])), np.asarray(list(ij_lists[1]))
    i = np.repeat(a, b.size).astype(np.uint32)
    j = np.tile(b, a.size).astype(np.uint32)
    for s in range(0, i.size, int(batch_size)):
        yield i[s:s + int(batch_size)], j[s:s + int(batch_size)]
