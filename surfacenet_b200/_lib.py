"""ctypes binding of libsurfacenet_b200.so (C ABI declared in include/surfacenet_b200.h).

There is NO fallback: if the shared library has not been built (``python -c "import
__graft_entry__ as g; g.build()"`` or ``make -C surfacenet_b200/csrc``) importing this module raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SN_LIB_PATH") or os.path.join(_HERE, "libsurfacenet_b200.so")     # SN_LIB_PATH: A/B runs of two builds (tools/)

SN_OK, SN_ERR_INVALID, SN_ERR_CUDA, SN_ERR_DOMAIN, SN_ERR_NOMEM = 0, -1, -2, -3, -4
MODE_FP32, MODE_TC_EXACT, MODE_TC_FAST = 0, 1, 2
DEFAULT_MODE = "exact"      # the mode every drop-in, bench.py and smoke() run unless told otherwise (no environment override)
MODES = {"fp32": MODE_FP32, "exact": MODE_TC_EXACT, "tc_exact": MODE_TC_EXACT, "fast": MODE_TC_FAST, "tc_fast": MODE_TC_FAST}


def resolve_mode(mode):
    """str / int -> SN_MODE_*.  "fast" (single-pass fp16 operands) misses the <= 1e-4 probability bound by ~100x
    (profiles/r01_bench_fast_c3.json: 9.8e-3), enough to flip tau / min_prob thresholds and ray-pool votes: it must be
    asked for explicitly and warns every time it is selected."""
    m = MODES[mode] if isinstance(mode, str) else int(mode)
    if m not in (MODE_FP32, MODE_TC_EXACT, MODE_TC_FAST):
        raise ValueError("unknown mode {!r}".format(mode))
    if m == MODE_TC_FAST:
        import warnings
        warnings.warn("surfacenet_b200: mode='fast' does NOT meet the 1e-4 parity bound of the reference (max-abs ~1e-2 on the "
                      "surface probability); use mode='exact' for reconstruction", RuntimeWarning, stacklevel=3)
    return m

if not os.path.exists(LIB_PATH):
    raise ImportError("surfacenet_b200: {} is missing -- build the CUDA library first (make -C surfacenet_b200/csrc); "
                      "there is no CPU fallback".format(LIB_PATH))
lib = C.CDLL(LIB_PATH)

_p, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float
# name -> (restype, argtypes); must list every symbol include/surfacenet_b200.h declares
SIGNATURES = {
    "sn_last_error": (C.c_char_p, []),
    "sn_version": (_i, []),
    "sn_launch_count": (_i64, []),
    "sn_launch_count_reset": (None, []),
    "sn_conv_path_counts": (None, [C.POINTER(C.c_int64)]),
    "sn_profile_enable": (None, [_i]),
    "sn_profile_collect": (_i, [C.POINTER(C.c_double), C.POINTER(C.c_int64), _i]),
    "sn_perspective_proj": (_i, [_p, _i, _p, _i64, _i, _p, _p, _p, _p]),
    "sn_cvc_gather": (_i, [_p, _p, _p, _i, _p, _p, _p, _p, _i, _i, _i, _p, _p, _p, _p, _p, _p]),
    "sn_sub_channel_mean": (_i, [_p, _i64, _i, _i64, _p, _p]),
    "sn_net_create": (_i, [C.POINTER(C.c_void_p), C.POINTER(C.c_int64), _i, C.POINTER(C.c_void_p)]),
    "sn_net_destroy": (None, [_p]),
    "sn_net_workspace_bytes": (_i64, [_p, _i, _i, _i]),
    "sn_net_forward": (_i, [_p, _p, _i, _i, _i, _p, _p, _p, _p, _i64, _i, _p]),
    "sn_net_relative_importance": (_i, [_p, _p, _i64, _i, _p, _p]),
    "sn_net_layer_conv": (_i, [_p, _i, _p, _i, _i, _p, _i, _p]),
    "sn_maxpool2": (_i, [_p, _i, _i, _i, _p, _p]),
    "sn_net_layer_upsample": (_i, [_p, _i, _p, _i, _i, _i, _p, _i, _i, _p]),
    "sn_fuse_weighted_average": (_i, [_p, _p, _i, _i, _i64, _p, _p]),
    "sn_raypool_workspace_bytes": (_i64, [_i, _i, _i]),
    "sn_raypool_votes": (_i, [_p, _i, _i, _f, _p, _p, _i, _p, _p, _i, _i, _i, _p, _p, _i64, _p]),
    "sn_cast_f32_to_f16": (_i, [_p, _i64, _p, _p]),
    "sn_infer_batch_workspace_bytes": (_i64, [_p, _i, _i, _i, _i]),
    "sn_infer_batch": (_i, [_p, _p, _p, _p, _i, _p, _p, _p, _p, _p, _i, _i, _i, _f, _p, _p, _p, _p, _p, _i64, _i, _p]),
    "sn_infer_batch_host": (_i, [_p, _p, _p, _p, _i, _p, _p, _p, _p, _p, _i, _i, _i, _f, _p, _p, _p, _p, _i64, _i, _p]),
    "sn_color_fusion": (_i, [_p, _p, _p, _p, _i, _i, _i64, _p, _p]),
    "sn_dense2sparse_workspace_bytes": (_i64, [_i, _i, _i]),
    "sn_dense2sparse": (_i, [_p, _p, _p, _i, _i, _i, _f, _i, _p, _p, _p, _p, _p, _p, _i64, _p, _i64, _p]),
    "sn_infer_batch_sparse_workspace_bytes": (_i64, [_p, _i, _i, _i, _i, _i]),
    "sn_sparse_post_workspace_bytes": (_i64, [_i, _i64, _i]),
    "sn_sparse_filter_voxels": (_i, [_p, _p, _p, _i, _i64, _p, C.c_double, _i, _i, _i, _p, _p, _i64, _p]),
    "sn_sparse_denoise": (_i, [_p, _p, _p, _p, _i, _i64, _i, _i, _i, _p, _p, _p, _p, _i64, _p]),
    "sn_sparse_adapthresh": (_i, [_p, _p, _p, _p, _p, _i, _i64, _i, _i, C.c_double, C.c_double, _i, _p, _p, _p, _p, _i64, _p]),
    "sn_viewpair_angles": (_i, [_p, _p, _i, _i64, _p, _i, _i, _p, _p]),
    "sn_viewpair_features": (_i, [_p, _p, _p, _p, _i, _i64, _i, _i, _i, _p, _p]),
    "sn_topn_rows": (_i, [_p, _i64, _i, _i, _p, _p]),
    "sn_select_from_similarity": (_i, [_p, _i64, _i, _i, _p, _p]),
    "sn_crop_patches": (_i, [_p, _i, _i, _p, _p, _i64, _i, _p, _p, _p]),
    "sn_simnet_create": (_i, [C.POINTER(C.c_void_p), C.POINTER(C.c_int64), _i, _i, C.POINTER(C.c_void_p)]),
    "sn_simnet_destroy": (None, [_p]),
    "sn_simnet_workspace_bytes": (_i64, [_p, _i64]),
    "sn_simnet_patch2embedding": (_i, [_p, _p, _i64, _p, _p, _i64, _p]),
    "sn_simnet_embeddingpair2simil": (_i, [_p, _p, _i64, _i, _p, _p]),
    "sn_infer_batch_sparse": (_i, [_p, _p, _p, _p, _i, _p, _p, _p, _p, _p, _i, _i, _i, _i, _f, _i, _p, _p, _p, _p, _p, _p, _i64, _p, _i64, _i, _p]),
}
for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)          # AttributeError here = the library does not export a declared symbol
    _fn.restype, _fn.argtypes = _res, _args


def last_error():
    return lib.sn_last_error().decode("utf-8", "replace")


def check(rc):
    """Map a status code to the exception the reference raises for the same condition:
    ValueError for bad shapes / arguments (utils/rayPooling.py:201-202, utils/camera.py:163-170),
    RuntimeError for everything the GPU runtime reports."""
    if rc == SN_OK:
        return
    msg = last_error()
    if rc in (SN_ERR_INVALID, SN_ERR_DOMAIN):
        raise ValueError(msg)
    raise RuntimeError("surfacenet_b200 [{}]: {}".format(rc, msg))


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("surfacenet_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """device/host pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())
