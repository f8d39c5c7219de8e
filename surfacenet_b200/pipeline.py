"""The per-batch hot loop body of main_reconstruct.py:132-162 as one GPU call, and its cube-sharded
multi-GPU form.

    HotPath.infer_batch(...)        device tensors in / out, nothing leaves the GPU   (sn_infer_batch)
    HotPath.infer_batch_host(...)   numpy in / numpy out, the call a reference user makes; H2D of the
                                    per-batch arguments and D2H of the results inside (sn_infer_batch_host)
    shard_bounds / infer_sharded    cubes are independent: contiguous split of the cube
                                    axis over ranks, no data-path collective, ONE all-gather of the
                                    per-cube probability (+ votes) volumes at the end.
"""
import numpy as np
from . import _lib
from .device import DeviceScene

MIN_PROB = 0.46          # params.py:66  __min_prob


class HotPath:
    def __init__(self, net, scene, mode="exact", min_prob=MIN_PROB):
        """net: SurfaceNet.Net; scene: DeviceScene (images + cameras resident on the device)."""
        self.torch = _lib.require_cuda()
        self.net, self.scene = net, scene
        self.mode = _lib.resolve_mode(mode)
        self.min_prob = float(min_prob)
        self._ws = None
        self._pinned = {}

    # ---- helpers ---------------------------------------------------------------------------------
    def _workspace(self, need):
        if need < 0:
            raise ValueError(_lib.last_error())
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = self.torch.empty(int(need), dtype=self.torch.uint8, device="cuda")
        return self._ws

    def _min_prob_f16(self):
        # utils/sparseCubes.py:115 + rayPooling.py:218: the float16 prediction is compared with the python float in float16
        return float(np.float16(self.min_prob))

    @staticmethod
    def _check_batch(pairs, xyz, resol, w):
        if pairs.ndim != 3 or pairs.shape[2] != 2:
            raise ValueError("viewPairs must have shape (N_cubes, N_viewPairs, 2), got {}".format(tuple(pairs.shape)))
        B, n_vp = int(pairs.shape[0]), int(pairs.shape[1])
        if tuple(xyz.shape) != (B, 3) or tuple(resol.shape) != (B,):
            raise ValueError("xyz must be (N_cubes,3) and resol (N_cubes,), got {} and {}".format(tuple(xyz.shape), tuple(resol.shape)))
        if w is None and n_vp > 1:
            raise ValueError("w (N_cubes, N_viewPairs) is required when N_viewPairs4inference >= 2")
        if w is not None and tuple(w.shape) != (B, n_vp):
            raise ValueError("w must have shape ({}, {}), got {}".format(B, n_vp, tuple(w.shape)))
        return B, n_vp

    # ---- device-resident call ----------------------------------------------------------------------
    def infer_batch(self, viewPairs, xyz, resol, w, D, want_unfused=False, ray_pool=True):
        """viewPairs (B,N_vp,2) i32, xyz (B,3) f32, resol (B,) f32, w (B,N_vp) f32|None: torch.cuda.
        -> dict(fused f32 (B,1,D,D,D), unfused f32 (B,N_vp,D,D,D)|None, pred16 f16 (B,D,D,D), votes u8 (B,D,D,D)|None)."""
        t = self.torch
        B, n_vp = self._check_batch(viewPairs, xyz, resol, w)
        D = int(D)
        fused = t.empty((B, 1, D, D, D), dtype=t.float32, device="cuda")
        unf = t.empty((B, n_vp, D, D, D), dtype=t.float32, device="cuda") if want_unfused else None
        p16 = t.empty((B, D, D, D), dtype=t.float16, device="cuda")
        votes = t.empty((B, D, D, D), dtype=t.uint8, device="cuda") if ray_pool else None
        ws = self._workspace(_lib.lib.sn_infer_batch_workspace_bytes(self.net.handle, B, n_vp, D, self.mode))
        sc = self.scene
        _lib.check(_lib.lib.sn_infer_batch(
            self.net.handle, _lib.ptr(sc.images), _lib.ptr(sc.img_offset), _lib.ptr(sc.img_hw), sc.n_views, _lib.ptr(sc.P),
            _lib.ptr(xyz), _lib.ptr(resol), _lib.ptr(viewPairs), _lib.ptr(w if n_vp > 1 else None), B, n_vp, D, self._min_prob_f16(),
            _lib.ptr(fused), _lib.ptr(unf), _lib.ptr(p16), _lib.ptr(votes), _lib.ptr(ws), ws.numel(), self.mode, _lib.stream_ptr()))
        return dict(fused=fused, unfused=unf, pred16=p16, votes=votes)

    # ---- host-buffer call (the drop-in entry) ------------------------------------------------------
    def _pin(self, name, shape, dtype):
        buf = self._pinned.get(name)
        n = int(np.prod(shape))
        if buf is None or buf.dtype != dtype or buf.numel() < n:
            buf = self.torch.empty(max(n, 1), dtype=dtype).pin_memory()
            self._pinned[name] = buf
        return buf[:n].view(*shape)

    def infer_batch_host(self, viewPairs, xyz, resol, w, D, want_fused=True, ray_pool=True):
        """numpy arguments as main_reconstruct.py holds them: viewPairs (B,N_vp,2) int, xyz (B,3) f32,
        resol (B,) f32, w (B,N_vp) f32|None.  Returns numpy views of pinned result buffers:
        dict(fused f32 (B,1,D,D,D)|None, pred16 f16 (B,D,D,D), votes u8 (B,D,D,D)|None); they are
        overwritten by the next call."""
        t = self.torch
        pairs = np.asarray(viewPairs)
        B, n_vp = self._check_batch(pairs, np.asarray(xyz), np.asarray(resol), None if w is None else np.asarray(w))
        self.scene.check_views(pairs)
        D = int(D)
        if B == 0:                                  # main_reconstruct.py:128-129: an empty selection is not an error
            self.h2d_bytes = self.d2h_bytes = 0
            return dict(fused=np.zeros((0, 1, D, D, D), np.float32) if want_fused else None, pred16=np.zeros((0, D, D, D), np.float16),
                        votes=np.zeros((0, D, D, D), np.uint8) if ray_pool else None)
        h_pairs = self._pin("pairs", (B, n_vp, 2), t.int32); h_pairs.numpy()[...] = pairs
        h_xyz = self._pin("xyz", (B, 3), t.float32); h_xyz.numpy()[...] = xyz
        h_resol = self._pin("resol", (B,), t.float32); h_resol.numpy()[...] = resol
        h_w = None
        if n_vp > 1:
            h_w = self._pin("w", (B, n_vp), t.float32); h_w.numpy()[...] = w
        h_fused = self._pin("fused", (B, 1, D, D, D), t.float32) if want_fused else None
        h_p16 = self._pin("pred16", (B, D, D, D), t.float16)
        h_votes = self._pin("votes", (B, D, D, D), t.uint8) if ray_pool else None
        inner = _lib.lib.sn_infer_batch_workspace_bytes(self.net.handle, B, n_vp, D, self.mode)
        if inner < 0:
            raise ValueError(_lib.last_error())
        V = D ** 3
        ws = self._workspace(inner + B * (12 + 4 + n_vp * 12 + V * 7) + 16 * 256)
        sc = self.scene
        _lib.check(_lib.lib.sn_infer_batch_host(
            self.net.handle, _lib.ptr(sc.images), _lib.ptr(sc.img_offset), _lib.ptr(sc.img_hw), sc.n_views, _lib.ptr(sc.P),
            _lib.ptr(h_xyz), _lib.ptr(h_resol), _lib.ptr(h_pairs), _lib.ptr(h_w), B, n_vp, D, self._min_prob_f16(),
            _lib.ptr(h_fused), _lib.ptr(h_p16), _lib.ptr(h_votes), _lib.ptr(ws), ws.numel(), self.mode, _lib.stream_ptr()))
        self.h2d_bytes = B * (12 + 4 + n_vp * 8 + (n_vp * 4 if n_vp > 1 else 0))
        self.d2h_bytes = B * V * ((4 if want_fused else 0) + 2 + (1 if ray_pool else 0))
        return dict(fused=None if h_fused is None else h_fused.numpy(), pred16=h_p16.numpy(),
                    votes=None if h_votes is None else h_votes.numpy())


    # ---- fused call with colour fusion + sparsification (main_reconstruct.py:134-162 in full) ---------
    def infer_batch_sparse(self, viewPairs, xyz, resol, w, D, cube_Dcenter, rayPool_thresh=0):
        """numpy arguments as infer_batch_host.  Everything up to sparseCubes.dense2sparse stays on the GPU; only the kept
        voxels cross PCIe.  -> dict(counts (B,), offsets (B+1,), ijk u8 (T,3) crop coordinates, pred f16 (T,), rgb u8 (T,3),
        votes u8 (T,)), ordered per cube exactly like the lists append_dense_2sparseList builds."""
        t = self.torch
        pairs = np.asarray(viewPairs)
        B, n_vp = self._check_batch(pairs, np.asarray(xyz), np.asarray(resol), None if w is None else np.asarray(w))
        self.scene.check_views(pairs)
        D, Dc = int(D), int(cube_Dcenter)
        if B == 0:
            return dict(counts=np.zeros(0, np.int32), offsets=np.zeros(1, np.int32), ijk=np.zeros((0, 3), np.uint8),
                        pred=np.zeros(0, np.float16), rgb=np.zeros((0, 3), np.uint8), votes=np.zeros(0, np.uint8))
        dev = lambda a, dt: t.from_numpy(np.ascontiguousarray(a, dtype=dt)).cuda()
        d_pairs, d_xyz, d_resol = dev(pairs, np.int32), dev(xyz, np.float32), dev(resol, np.float32)
        d_w = dev(w, np.float32) if n_vp > 1 else None
        cap = B * Dc ** 3
        counts = t.zeros(B, dtype=t.int32, device="cuda"); offsets = t.zeros(B + 1, dtype=t.int32, device="cuda")
        ijk = t.empty((cap, 3), dtype=t.uint8, device="cuda"); pred = t.empty(cap, dtype=t.float16, device="cuda")
        rgb = t.empty((cap, 3), dtype=t.uint8, device="cuda"); votes = t.empty(cap, dtype=t.uint8, device="cuda")
        ws = self._workspace(_lib.lib.sn_infer_batch_sparse_workspace_bytes(self.net.handle, B, n_vp, D, Dc, self.mode))
        sc = self.scene
        _lib.check(_lib.lib.sn_infer_batch_sparse(
            self.net.handle, _lib.ptr(sc.images), _lib.ptr(sc.img_offset), _lib.ptr(sc.img_hw), sc.n_views, _lib.ptr(sc.P),
            _lib.ptr(d_xyz), _lib.ptr(d_resol), _lib.ptr(d_pairs), _lib.ptr(d_w), B, n_vp, D, Dc, self._min_prob_f16(), int(rayPool_thresh),
            _lib.ptr(counts), _lib.ptr(offsets), _lib.ptr(ijk), _lib.ptr(pred), _lib.ptr(rgb), _lib.ptr(votes), cap,
            _lib.ptr(ws), ws.numel(), self.mode, _lib.stream_ptr()))
        off = offsets.cpu().numpy()
        T = int(off[-1])
        self.d2h_bytes = (B * 2 + 1) * 4 + T * 8
        return dict(counts=counts.cpu().numpy(), offsets=off, ijk=ijk[:T].cpu().numpy(), pred=pred[:T].cpu().numpy(),
                    rgb=rgb[:T].cpu().numpy(), votes=votes[:T].cpu().numpy())


# ---- cube sharding ---------------------------------------------------------------
def shard_bounds(n_cubes, world_size):
    """Contiguous, equal-count split of the cube axis: every rank gets ceil(n/world) slots (the last
    ranks' tails are padding so that a single equal-count all-gather reassembles the batch)."""
    per = -(-int(n_cubes) // int(world_size)) if n_cubes else 0
    return per, [(min(r * per, n_cubes), min((r + 1) * per, n_cubes)) for r in range(world_size)]


def infer_sharded(compute_fn, n_cubes, out_specs, rank, world_size, group=None, device="cuda"):
    """Run ``compute_fn(lo, hi) -> tuple of tensors (hi-lo, ...)`` on this rank's contiguous slice of
    the cube axis and reassemble the full batch on every rank with ONE all-gather per output.
    out_specs: list of (trailing_shape, torch dtype) describing compute_fn's outputs.
    world_size == 1 does no communication."""
    import torch
    import torch.distributed as dist
    per, bounds = shard_bounds(n_cubes, world_size)
    lo, hi = bounds[rank]
    local = compute_fn(lo, hi) if hi > lo else tuple(torch.empty((0,) + tuple(s), dtype=dt, device=device) for s, dt in out_specs)
    if world_size == 1:
        return tuple(local)
    outs = []
    for x, (shape, dt) in zip(local, out_specs):
        send = torch.zeros((per,) + tuple(shape), dtype=dt, device=device)
        send[:hi - lo] = x
        full = torch.empty((world_size * per,) + tuple(shape), dtype=dt, device=device)
        dist.all_gather_into_tensor(full, send, group=group) if device != "cpu" else \
            dist.all_gather(list(full.view((world_size, per) + tuple(shape)).unbind(0)), send, group=group)
        outs.append(full[:n_cubes])
    return tuple(outs)
