"""Drop-in for nets/similarityNet.py:234-246 similarityNet_inference -> (patch2embedding_fn, embeddingPair2simil_fn), evaluated on
the GPU (csrc/simnet.cu: sn_simnet_patch2embedding, sn_simnet_embeddingpair2simil).  "Next" row N3.

Parameter list = lasagne.layers.get_all_param_values([embedding layer, similarity layer]) (similarityNet.py:241-244): 13 x
(conv W (C_out,C_in,3,3), b), dense W (5888,128), b (128), similarity W (1,1), b (1,) -- 30 float32 arrays."""
import ctypes as C
import pickle
import numpy as np
from . import _lib

CONV_CH = [(3, 64), (64, 64), (64, 128), (128, 128), (128, 256), (256, 256), (256, 256), (256, 512), (512, 512), (512, 512),
           (512, 512), (512, 512), (512, 512)]                                  # similarityNet.py:28-45 (VGG-16)
D_CONCAT, D_EMBEDDING, N_ARRAYS = 5888, 128, 30                                 # similarityNet.py:46-56, params.py:88
PARAM_SHAPES = [s for cin, cout in CONV_CH for s in ((cout, cin, 3, 3), (cout,))] + [(D_CONCAT, D_EMBEDDING), (D_EMBEDDING,), (1, 1), (1,)]


def synthetic_params(seed=0):
    """Deterministic He-initialised parameter list (the trained .model file is not distributed with the reference)."""
    rs = np.random.RandomState(4000 + seed)
    out = []
    for shp in PARAM_SHAPES[:-2]:
        if len(shp) > 1:
            fan_in = int(np.prod(shp[1:])) if len(shp) == 4 else shp[0]
            out.append((rs.randn(*shp) * np.sqrt(2.0 / fan_in)).astype(np.float32))
        else:
            out.append((rs.randn(*shp) * 0.05).astype(np.float32))
    out[-2] = (out[-2] * 4).astype(np.float32)                                  # embeddings of unit-ish scale
    out += [np.array([[-3.0]], np.float32), np.array([2.0], np.float32)]         # similarity = sigmoid(-3 * dist + 2)
    return out


def load_model_file(path):
    with open(path, "rb") as f:
        arrays = pickle.load(f, encoding="latin1")                               # python-2 pickle of numpy arrays
    return [np.asarray(a, dtype=np.float32) for a in arrays]


class SimNet:
    def __init__(self, params, patch=64):
        _lib.require_cuda()
        if len(params) != N_ARRAYS:
            raise ValueError("similarityNet needs {} parameter arrays, got {}".format(N_ARRAYS, len(params)))
        arrs = []
        for a, shp in zip(params, PARAM_SHAPES):
            a = np.ascontiguousarray(a, dtype=np.float32)
            if a.shape != tuple(shp):
                raise ValueError("similarityNet parameter has shape {}, expected {}".format(a.shape, shp))
            arrs.append(a)
        ptrs = (C.c_void_p * N_ARRAYS)(*[a.ctypes.data for a in arrs])
        sizes = (C.c_int64 * N_ARRAYS)(*[a.size for a in arrs])
        h = C.c_void_p()
        _lib.check(_lib.lib.sn_simnet_create(ptrs, sizes, N_ARRAYS, int(patch), C.byref(h)))
        self.handle, self.patch, self._ws = h, int(patch), None

    def __del__(self):
        if getattr(self, "handle", None):
            _lib.lib.sn_simnet_destroy(self.handle)
            self.handle = None

    def patch2embedding(self, patches):
        """patches cuda (n,3,64,64) f32 -> cuda (n,128) f32"""
        torch = _lib.require_cuda()
        if patches.dim() != 4 or tuple(patches.shape[1:]) != (3, self.patch, self.patch):
            raise ValueError("patches must have shape (N,3,{0},{0}), got {1}".format(self.patch, tuple(patches.shape)))
        n = int(patches.shape[0])
        out = torch.empty((n, D_EMBEDDING), dtype=torch.float32, device="cuda")
        need = int(_lib.lib.sn_simnet_workspace_bytes(self.handle, n))
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device="cuda")
        _lib.check(_lib.lib.sn_simnet_patch2embedding(self.handle, _lib.ptr(patches.contiguous()), n, _lib.ptr(out), _lib.ptr(self._ws),
                                                      self._ws.numel(), _lib.stream_ptr()))
        return out

    def embeddingpair2simil(self, pairs):
        """pairs cuda (2M, E) f32, rows (2m, 2m+1) form pair m -> cuda (M,1) f32"""
        torch = _lib.require_cuda()
        if pairs.dim() != 2 or pairs.shape[0] % 2:
            raise ValueError("embedding pairs must have shape (2M, D_embedding), got {}".format(tuple(pairs.shape)))
        M = int(pairs.shape[0]) // 2
        out = torch.empty((M, 1), dtype=torch.float32, device="cuda")
        _lib.check(_lib.lib.sn_simnet_embeddingpair2simil(self.handle, _lib.ptr(pairs.contiguous()), M, int(pairs.shape[1]), _lib.ptr(out),
                                                          _lib.stream_ptr()))
        return out


def _to_dev(torch, a):
    if isinstance(a, np.ndarray):
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda(), True
    return a.to(device="cuda", dtype=torch.float32), False


def similarityNet_inference(model_file, imgPatch_hw_size):
    """nets/similarityNet.py:234-246.  model_file: path of the pickled parameter list or the list itself.
    Both callables take numpy (-> numpy) or cuda tensors (-> cuda tensors)."""
    hw = tuple(imgPatch_hw_size)
    if len(hw) != 2 or hw[0] != hw[1]:
        raise ValueError("imgPatch_hw_size must be a square (h, w), got {}".format(imgPatch_hw_size))
    net = SimNet(load_model_file(model_file) if isinstance(model_file, str) else model_file, patch=hw[0])

    def patch2embedding_fn(patches):
        torch = _lib.require_cuda()
        p, is_np = _to_dev(torch, patches)
        out = net.patch2embedding(p)
        return out.cpu().numpy() if is_np else out

    def embeddingPair2simil_fn(embeddingPairs):
        torch = _lib.require_cuda()
        p, is_np = _to_dev(torch, embeddingPairs)
        out = net.embeddingpair2simil(p)
        return out.cpu().numpy() if is_np else out

    patch2embedding_fn.net = net
    return patch2embedding_fn, embeddingPair2simil_fn
