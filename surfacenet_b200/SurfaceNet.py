"""Drop-in for the inference part of nets/SurfaceNet.py, computed on the GPU.

    SurfaceNet_inference(N_viewPairs4inference, model_file, layerNameList_2_load)
        -> (viewPair_relativeImpt_fn, nViewPair_SurfaceNet_fn)                 nets/SurfaceNet.py:385-402

    nViewPair_SurfaceNet_fn(X)            if N_viewPairs4inference == 1         main_reconstruct.py:145-146
    nViewPair_SurfaceNet_fn(X, w)         otherwise
        -> [fused (B,1,D,D,D) float32, unfused (B,N_vp,D,D,D) float32]         nets/SurfaceNet.py:352-357,374-376
    viewPair_relativeImpt_fn(features, n_samples_perGroup) -> (G, n) float32   nets/SurfaceNet.py:337, viewPairSelection.py:77

numpy in -> numpy out (as the compiled Theano callables behave); torch.cuda tensors in -> torch.cuda
tensors out (no host round trip).  The arithmetic is libsurfacenet_b200.so (csrc/net*.cu, conv_tc.cu).
"""
import ctypes as C
import numpy as np
from . import _lib, weights


class Net:
    """Owns the device copy of the 105-array parameter list (sn_net_create / sn_net_destroy)."""

    def __init__(self, params):
        _lib.require_cuda()
        self.params = weights.validate(params)
        arr = (C.c_void_p * len(self.params))(*[a.ctypes.data for a in self.params])
        sizes = (C.c_int64 * len(self.params))(*[a.size for a in self.params])
        h = C.c_void_p()
        _lib.check(_lib.lib.sn_net_create(arr, sizes, len(self.params), C.byref(h)))
        self.handle = h
        self._ws = None

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h and _lib is not None and getattr(_lib, "lib", None) is not None:      # interpreter shutdown clears module globals
            _lib.lib.sn_net_destroy(h)

    def workspace(self, nbytes):
        torch = _lib.require_cuda()
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(int(nbytes), dtype=torch.uint8, device="cuda")
        return self._ws

    def forward(self, X, w=None, N_vp=1, mode="exact", want_unfused=True):
        """X torch.cuda (B*N_vp,6,D,D,D) f32 (mean subtracted); w torch.cuda (B,N_vp) f32 or None."""
        torch = _lib.require_cuda()
        if X.dim() != 5 or X.shape[1] != 6 or X.shape[2] != X.shape[3] or X.shape[3] != X.shape[4]:
            raise ValueError("X must have shape (N_cubes*N_viewPairs, 6, D, D, D), got {}".format(tuple(X.shape)))
        if X.shape[0] % N_vp:
            raise ValueError("X holds {} samples, not a multiple of N_viewPairs4inference={}".format(X.shape[0], N_vp))
        B, D = X.shape[0] // N_vp, X.shape[-1]
        if N_vp > 1:
            if w is None:
                raise ValueError("w (N_cubes, N_viewPairs) is required when N_viewPairs4inference >= 2")
            if tuple(w.shape) != (B, N_vp):
                raise ValueError("w must have shape ({}, {}), got {}".format(B, N_vp, tuple(w.shape)))
        X = X.contiguous()
        m = _lib.resolve_mode(mode)
        fused = torch.empty((B, 1, D, D, D), dtype=torch.float32, device="cuda")
        unf = torch.empty((B, N_vp, D, D, D), dtype=torch.float32, device="cuda") if (want_unfused and N_vp > 1) else None
        need = _lib.lib.sn_net_workspace_bytes(self.handle, B * N_vp, D, m)
        if need < 0:
            raise ValueError(_lib.last_error() or "unsupported cube size {}".format(D))
        ws = self.workspace(need)
        _lib.check(_lib.lib.sn_net_forward(self.handle, _lib.ptr(X), B, N_vp, D, _lib.ptr(w if N_vp > 1 else None), _lib.ptr(fused),
                                           _lib.ptr(unf), _lib.ptr(ws), ws.numel(), m, _lib.stream_ptr()))
        return fused, (fused if N_vp == 1 else unf)

    def relative_importance(self, features, n_samples_perGroup):
        torch = _lib.require_cuda()
        if features.dim() != 2 or features.shape[1] != weights.D_VIEWPAIR_FEATURE:
            raise ValueError("features must have shape (N, {}), got {}".format(weights.D_VIEWPAIR_FEATURE, tuple(features.shape)))
        n = int(n_samples_perGroup)
        if n < 1 or features.shape[0] % n:
            raise ValueError("{} feature rows is not a multiple of n_samples_perGroup={}".format(features.shape[0], n))
        out = torch.empty((features.shape[0] // n, n), dtype=torch.float32, device="cuda")
        _lib.check(_lib.lib.sn_net_relative_importance(self.handle, _lib.ptr(features.contiguous()), features.shape[0], n,
                                                       _lib.ptr(out), _lib.stream_ptr()))
        return out


def _to_dev(torch, a):
    if isinstance(a, np.ndarray):
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda(), True
    return a.to(device="cuda", dtype=torch.float32), False


def make_inference_fns(net, N_viewPairs4inference, mode="exact"):
    """The two callables nets/SurfaceNet.py:382 compiles, bound to ``net``."""
    N_vp = int(N_viewPairs4inference)

    def nViewPair_SurfaceNet_fn(X, w=None):
        torch = _lib.require_cuda()
        if N_vp > 1 and w is None:
            raise TypeError("nViewPair_SurfaceNet_fn(X, w): w is required when N_viewPairs4inference >= 2")
        Xd, is_np = _to_dev(torch, X)
        wd = None if (w is None or N_vp == 1) else _to_dev(torch, w)[0]
        fused, unf = net.forward(Xd, wd, N_vp, mode)
        if is_np:
            f = fused.cpu().numpy()
            return [f, f if N_vp == 1 else unf.cpu().numpy()]
        return [fused, unf]

    def viewPair_relativeImpt_fn(features, n_samples_perGroup):
        torch = _lib.require_cuda()
        fd, is_np = _to_dev(torch, features)
        out = net.relative_importance(fd, n_samples_perGroup)
        return out.cpu().numpy() if is_np else out

    return viewPair_relativeImpt_fn, nViewPair_SurfaceNet_fn


def SurfaceNet_inference(N_viewPairs4inference, model_file, layerNameList_2_load=None, mode="exact"):
    """nets/SurfaceNet.py:385-402.  ``model_file``: path of the reference's pickled parameter list
    (or .npz), or an in-memory list of the 105 arrays.  ``layerNameList_2_load`` is accepted for call
    compatibility: the reference always loads ["output_SurfaceNet_reshape","output_softmaxWeights"]
    (params.py:105), i.e. the whole 105-array list.
    ``mode``: "exact" (default; tcgen05 tensor cores, fp16 hi/lo split operands, <= 1e-4 on the probability),
    "fp32" (CUDA-core cross-check, ~26x slower), "fast" (single-pass fp16 operands: MISSES the 1e-4 bound, warns)."""
    if layerNameList_2_load is not None and list(layerNameList_2_load) != ["output_SurfaceNet_reshape", "output_softmaxWeights"]:
        raise ValueError("only the full parameter list of params.py:105 can be loaded, got {}".format(layerNameList_2_load))
    params = weights.load_model_file(model_file) if isinstance(model_file, str) else model_file
    net = Net(params)
    fns = make_inference_fns(net, N_viewPairs4inference, mode)
    fns[1].net = net
    return fns
