"""Device-resident form of the scene's sparse cubes (the per-cube lists sparseCubes.append_dense_2sparseList accumulates /
the flat arrays of the NPZ file, utils/sparseCubes.py:330-366) and the N4 operators on it: filter_voxels, denoise_crossCubes,
the adapthresh refinement.  All arithmetic runs in libsurfacenet_b200.so (csrc/postprocess.cu); there is no CPU path."""
import numpy as np
from . import _lib


def _flat(lst, dtype, width=None):
    if len(lst) == 0:
        return np.zeros((0,) if width is None else (0, width), dtype)
    return np.ascontiguousarray(np.concatenate([np.asarray(x) for x in lst], axis=0).astype(dtype, copy=False))


class DeviceSparseCubes:
    """cube_ijk_np (C,3) int; vxl_ijk_list[i] (n_i,3) uint8; prediction_list[i] (n_i,) float16; rayPooling_votes_list[i] (n_i,) uint8."""

    def __init__(self, cube_ijk_np, vxl_ijk_list, prediction_list=None, rayPooling_votes_list=None):
        torch = _lib.require_cuda()
        self.torch = torch
        cube_ijk_np = np.asarray(cube_ijk_np)
        self.C = len(vxl_ijk_list)
        if cube_ijk_np.shape[0] != self.C or (self.C and cube_ijk_np.shape[1] != 3):
            raise ValueError("cube_ijk_np must have shape ({},3), got {}".format(self.C, cube_ijk_np.shape))
        for a in vxl_ijk_list:
            if np.asarray(a).ndim != 2 or np.asarray(a).shape[1] != 3:
                raise ValueError("every vxl_ijk_list entry must have shape (N,3)")
            if np.asarray(a).size and (np.asarray(a).min() < 0 or np.asarray(a).max() > 255):
                raise ValueError("voxel indices must fit uint8")
        self.sizes = np.array([np.asarray(a).shape[0] for a in vxl_ijk_list], np.int64)
        self.offsets_np = np.concatenate([[0], np.cumsum(self.sizes)]).astype(np.int64)
        self.N = int(self.offsets_np[-1])
        ijk = _flat(vxl_ijk_list, np.uint8, 3)
        self.G = int(ijk.max()) + 1 if ijk.size else 1
        dev = lambda a: torch.from_numpy(a).cuda()
        self.cube_ijk = dev(np.ascontiguousarray(cube_ijk_np.astype(np.int64).astype(np.int32).reshape(-1, 3)))
        self.offsets = dev(self.offsets_np)
        self.ijk = dev(ijk)
        self.pred = None
        self.votes = None
        if prediction_list is not None:
            p = _flat(prediction_list, np.float16)
            if p.shape[0] != self.N:
                raise Warning('make sure # of voxels in each cube are consistent.')
            self.pred = dev(p)
        if rayPooling_votes_list is not None and len(rayPooling_votes_list):
            v = _flat(rayPooling_votes_list, np.uint8)
            if v.shape[0] != self.N:
                raise Warning('make sure # of voxels in each cube are consistent.')
            self.votes = dev(v)
        need = _lib.lib.sn_sparse_post_workspace_bytes(self.C, self.N, self.G)
        if need < 0:
            raise ValueError("sparse cubes: bad sizes (C={}, N={}, G={})".format(self.C, self.N, self.G))
        self.ws = torch.empty(int(need), dtype=torch.uint8, device="cuda")

    @classmethod
    def from_flat(cls, cube_ijk_np, offsets, ijk, pred=None, votes=None, grid_extent=None):
        """The same object from the flat arrays of the NPZ file / of HotPath.infer_batch_sparse (numpy arrays or cuda tensors):
        cube_ijk_np (C,3) int, offsets (C+1,) = cube_1st_vxlIndx_np, ijk (N,3) uint8, pred (N,) float16, votes (N,) uint8.  No per-cube
        python lists are built; grid_extent defaults to max(ijk) + 1."""
        torch = _lib.require_cuda()
        self = cls.__new__(cls)
        self.torch = torch
        dev = lambda a, dt: (a.to(device="cuda", dtype=dt) if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a).astype(dt_np[dt], copy=False)).cuda()).contiguous()
        dt_np = {torch.int32: np.int32, torch.int64: np.int64, torch.uint8: np.uint8, torch.float16: np.float16}
        self.cube_ijk = dev(np.asarray(cube_ijk_np).astype(np.int64) if not torch.is_tensor(cube_ijk_np) else cube_ijk_np, torch.int32).reshape(-1, 3)
        self.offsets = dev(offsets, torch.int64)
        self.offsets_np = self.offsets.cpu().numpy()
        self.C, self.N = int(self.cube_ijk.shape[0]), int(self.offsets_np[-1]) if self.offsets_np.size else 0
        if self.offsets_np.size != self.C + 1 or (self.C and np.any(np.diff(self.offsets_np) < 0)):
            raise ValueError("offsets must be a non-decreasing array of {} entries".format(self.C + 1))
        self.sizes = np.diff(self.offsets_np)
        self.ijk = dev(ijk, torch.uint8).reshape(-1, 3)
        if int(self.ijk.shape[0]) != self.N:
            raise Warning('make sure # of voxels in each cube are consistent.')
        self.G = int(grid_extent) if grid_extent is not None else (int(self.ijk.max().item()) + 1 if self.N else 1)
        self.pred = None if pred is None else dev(pred, torch.float16).reshape(-1)
        self.votes = None if votes is None else dev(votes, torch.uint8).reshape(-1)
        for t in (self.pred, self.votes):
            if t is not None and int(t.shape[0]) != self.N:
                raise Warning('make sure # of voxels in each cube are consistent.')
        need = _lib.lib.sn_sparse_post_workspace_bytes(self.C, self.N, self.G)
        if need < 0:
            raise ValueError("sparse cubes: bad sizes (C={}, N={}, G={})".format(self.C, self.N, self.G))
        self.ws = torch.empty(int(need), dtype=torch.uint8, device="cuda")
        return self

    # ---- list <-> flat -------------------------------------------------------------------------------------------
    def upload_mask(self, vxl_mask_list):
        m = _flat([np.asarray(x).astype(bool) for x in vxl_mask_list], np.uint8)
        if len(vxl_mask_list) != self.C or m.shape[0] != self.N:
            raise Warning('make sure # of voxels in each cube are consistent.')
        return self.torch.from_numpy(m).cuda()

    def split(self, flat, dtype=None):
        flat = flat.cpu().numpy() if hasattr(flat, "cpu") else np.asarray(flat)
        if dtype is not None:
            flat = flat.astype(dtype)
        o = self.offsets_np
        return [flat[o[n]:o[n + 1]] for n in range(self.C)]

    # ---- operators -----------------------------------------------------------------------------------------------
    def filter_voxels(self, mask=None, prob_thresh=None, rayPool_thresh=None):
        """utils/sparseCubes.py:205-243 on the device: -> mask tensor (N,) uint8.  prob_thresh: None | scalar | per-cube
        sequence / device tensor of float64; rayPool_thresh: None | number (votes >= thresh)."""
        torch = self.torch
        out = torch.ones(self.N, dtype=torch.uint8, device="cuda") if mask is None else mask
        per_cube, scalar, has_prob = None, 0.0, 0
        if prob_thresh is not None:
            if self.pred is None:
                raise ValueError("filter_voxels: no predictions were given")
            has_prob = 1
            if torch.is_tensor(prob_thresh):
                per_cube = prob_thresh
            elif isinstance(prob_thresh, (list, tuple, np.ndarray)):
                per_cube = torch.from_numpy(np.ascontiguousarray(prob_thresh, dtype=np.float64)).cuda()
                if per_cube.numel() != self.C:
                    raise ValueError("prob_thresh list must have one entry per cube")
            else:
                scalar = float(prob_thresh)
        rp = -1
        if rayPool_thresh is not None:
            if self.votes is None:
                raise ValueError("filter_voxels: no ray-pooling votes were given")
            rp = max(0, int(np.ceil(float(rayPool_thresh))))          # uint8 votes >= t  <=>  votes >= ceil(t)
        _lib.check(_lib.lib.sn_sparse_filter_voxels(_lib.ptr(self.pred), _lib.ptr(self.votes) if rp >= 0 else None, _lib.ptr(self.offsets),
                                                    self.C, self.N, _lib.ptr(per_cube), scalar, has_prob, rp, 0 if mask is None else 1,
                                                    _lib.ptr(out), _lib.ptr(self.ws), self.ws.numel(), _lib.stream_ptr()))
        return out

    def denoise(self, mask, D_cube, neighbor_dist=3, want_keep=True, want_labels=False):
        """utils/denoising.py:67-184 on the device -> dict(keep (N,) u8 | None, labels (N,) i32 | None, n_labels (C,) i32 | None)."""
        torch = self.torch
        keep = torch.zeros(self.N, dtype=torch.uint8, device="cuda") if want_keep else None
        labels = torch.zeros(self.N, dtype=torch.int32, device="cuda") if want_labels else None
        n_labels = torch.zeros(self.C, dtype=torch.int32, device="cuda") if want_labels else None
        _lib.check(_lib.lib.sn_sparse_denoise(_lib.ptr(self.cube_ijk), _lib.ptr(self.offsets), _lib.ptr(self.ijk), _lib.ptr(mask), self.C,
                                              self.N, self.G, int(D_cube), int(neighbor_dist), _lib.ptr(keep), _lib.ptr(labels),
                                              _lib.ptr(n_labels), _lib.ptr(self.ws), self.ws.numel(), _lib.stream_ptr()))
        return dict(keep=keep, labels=labels, n_labels=n_labels)

    def adapthresh(self, init_mask, mask, thresh, D_cube, max_probThresh, beta, n_iter=1, want_argmin=False):
        """utils/adapthresh.py:126-174: n_iter refinement iterations in place on `mask` (N,) u8 and `thresh` (C,) f64 tensors."""
        torch = self.torch
        if self.pred is None:
            raise ValueError("adapthresh: no predictions were given")
        arg = torch.zeros((n_iter, self.C), dtype=torch.int32, device="cuda") if want_argmin else None
        _lib.check(_lib.lib.sn_sparse_adapthresh(_lib.ptr(self.cube_ijk), _lib.ptr(self.offsets), _lib.ptr(self.ijk), _lib.ptr(self.pred),
                                                 _lib.ptr(init_mask), self.C, self.N, self.G, int(D_cube), float(max_probThresh), float(beta),
                                                 int(n_iter), _lib.ptr(thresh), _lib.ptr(mask), _lib.ptr(arg), _lib.ptr(self.ws),
                                                 self.ws.numel(), _lib.stream_ptr()))
        return arg
