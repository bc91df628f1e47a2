"""ctypes binding of libtracknet_b200.so (C ABI declared in include/tracknet_b200.h).

The shared library is the product: there is no Python/PyTorch fallback for any op. If the library is
missing the import fails loudly with build instructions (``python -c "import __graft_entry__ as g; g.build()"``).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# TNB_LIBRARY: another build of the same ABI (experiments: A/B of a compile-time constant on one box)
LIB_PATH = os.environ.get("TNB_LIBRARY") or os.path.join(_HERE, "libtracknet_b200.so")


class TnbError(RuntimeError):
    pass


class Src(C.Structure):  # tnb_src_t
    _fields_ = [("ptr", C.c_void_p), ("scale", C.c_void_p), ("shift", C.c_void_p),
                ("C", C.c_int), ("Hs", C.c_int), ("Ws", C.c_int), ("mode", C.c_int)]


class View(C.Structure):  # tnb_view_t
    _fields_ = [("s", Src * 2), ("C0", C.c_int), ("C", C.c_int), ("N", C.c_int), ("H", C.c_int), ("W", C.c_int)]


class GradSrc(C.Structure):  # tnb_gradsrc_t
    _fields_ = [("ptr", C.c_void_p), ("C", C.c_int), ("coff", C.c_int), ("mode", C.c_int),
                ("Hs", C.c_int), ("Ws", C.c_int)]


class BnBwd(C.Structure):  # tnb_bnbwd_t
    _fields_ = [("g", GradSrc * 2), ("ng", C.c_int), ("z", C.c_void_p),
                ("scale", C.c_void_p), ("shift", C.c_void_p), ("mean", C.c_void_p), ("invstd", C.c_void_p),
                ("N", C.c_int), ("H", C.c_int), ("W", C.c_int), ("C", C.c_int),
                ("part", C.c_void_p), ("sums", C.c_void_p), ("dz", C.c_void_p), ("inv_count", C.c_float),
                ("amax", C.c_void_p), ("dz_format", C.c_int), ("act_presplit", C.c_void_p),
                ("gmax", C.c_void_p), ("dz_mul", C.c_void_p), ("act_pool", C.c_int), ("act_full", C.c_void_p)]


class TrackNetCfg(C.Structure):  # tnb_tracknet_cfg_t
    _fields_ = [("n", C.c_int), ("h", C.c_int), ("w", C.c_int), ("in_dim", C.c_int), ("out_dim", C.c_int),
                ("training", C.c_int), ("fwd_terms", C.c_int), ("bwd_terms", C.c_int), ("variant", C.c_int),
                ("bn_eps", C.c_float), ("bn_momentum", C.c_float)]


SRC_IDENTITY, SRC_AFFINE_RELU, SRC_AFFINE_RELU_POOL, SRC_AFFINE_RELU_UP, SRC_PRESPLIT, SRC_PRESPLIT_UP, SRC_PLANAR16 = 0, 1, 2, 3, 4, 5, 6
GRAD_SAME, GRAD_POOL, GRAD_UP = 0, 1, 2

vp, i32, i64, f32, f64, sz = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_double, C.c_size_t

# name -> (restype, argtypes); MUST list every symbol include/tracknet_b200.h declares (tests check this)
SIGNATURES = {
    "tnb_last_error": (C.c_char_p, []),
    "tnb_abi_version": (i32, []),
    "tnb_pack_nchw_to_nhwc": (i32, [vp, vp, i32, i32, i32, i32, i32, vp]),
    "tnb_conv3x3_wpack_elems": (sz, [i32, i32]),
    "tnb_conv3x3_pack_weights": (i32, [vp, vp, i32, i32, i32, i32, vp]),
    "tnb_conv3x3_stat_rows": (i32, [i32, i32, i32, i32, i32, i32]),
    "tnb_conv3x3_plan_query": (i32, [i32, i32, i32, i32, i32, i32, vp]),
    "tnb_conv3x3_fwd": (i32, [C.POINTER(View), vp, vp, vp, i32, i32, i32, i32, vp]),
    "tnb_conv3x3_dgrad_bnreduce_rows": (i32, [i32, i32, i32, i32, i32, i32]),
    "tnb_conv3x3_dgrad_bnreduce": (i32, [C.POINTER(View), vp, vp, vp, i32, i32, vp, vp, vp, vp, vp, vp]),
    "tnb_conv3x3_wgrad": (i32, [C.POINTER(View), vp, vp, i32, i32, i32, i32, i32, vp, vp]),
    "tnb_conv3x3_wgrad_ws_elems": (sz, [C.POINTER(View), i32]),
    "tnb_conv3x3_wgrad_ws": (i32, [C.POINTER(View), vp, vp, i32, i32, i32, i32, vp, i32, vp, vp]),
    "tnb_pack_nchw_to_planar16": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, vp]),
    "tnb_presplit_bf16": (i32, [vp, vp, i64, i32, vp]),
    "tnb_presplit_fp16": (i32, [vp, vp, i64, i32, f32, vp]),
    "tnb_view_presplit": (i32, [C.POINTER(View), vp, i32, vp]),
    "tnb_bn_finalize": (i32, [vp, i32, f64, vp, vp, vp, vp, f32, f32, i32, vp, vp, vp, vp, i32, vp]),
    "tnb_bn_bwd_blocks": (i32, [i32, i32, i32, i32]),
    "tnb_bn_relu_bwd_reduce": (i32, [C.POINTER(BnBwd), vp]),
    "tnb_bn_relu_bwd_finalize": (i32, [vp, i32, i32, vp, vp, vp, vp]),
    "tnb_bn_relu_bwd_apply": (i32, [C.POINTER(BnBwd), vp]),
    "tnb_conv1x1_bias_sigmoid_fwd": (i32, [C.POINTER(Src), i32, i32, i32, vp, vp, i32, vp, vp]),
    "tnb_conv1x1_bias_sigmoid_bwd_workspace_bytes": (sz, [i32, i32, i32, i32]),
    "tnb_conv1x1_bias_sigmoid_bwd": (i32, [C.POINTER(Src), i32, i32, i32, vp, i32, vp, vp, vp, vp, vp, vp, vp]),
    "tnb_wbce_workspace_bytes": (sz, [i32]),
    "tnb_wbce_fwd": (i32, [vp, vp, i32, i64, i32, vp, vp, vp]),
    "tnb_wbce_bwd": (i32, [vp, vp, vp, i32, i64, i32, vp, vp]),
    "tnb_mixup": (i32, [vp, vp, vp, vp, i32, i64, vp]),
    "tnb_adam_multi": (i32, [vp, i32, i64, f32, f32, f32, f32, f32, i32, vp]),
    "tnb_heatmap_decode_workspace_bytes": (sz, [i32, i32, i32]),
    "tnb_heatmap_decode": (i32, [vp, i32, f32, i32, i32, i32, vp, vp, vp]),
    "tnb_inpaintnet_fwd": (i32, [vp, vp, C.POINTER(vp), i32, i32, vp, vp]),
    "tnb_inpaintnet_rectify": (i32, [vp, vp, C.POINTER(vp), i32, i32, f32, vp, vp]),
    "tnb_median_u8": (i32, [vp, i32, i64, vp, vp, vp]),
    "tnb_label_discs": (i32, [vp, i32, i32, i32, f32, vp, vp]),
    "tnb_inpaintnet_bwd": (i32, [vp, vp, C.POINTER(vp), vp, C.POINTER(vp), i32, i32, vp, vp]),
    "tnb_resize_frames": (i32, [vp, i32, i32, i32, i32, vp, vp, i32, vp, vp, i32, i32, i32, vp, vp, i32, C.c_longlong, i32, i32, vp]),
    "tnb_bg_subtract_u8": (i32, [vp, vp, C.c_longlong, i32, i32, vp, vp]),
    "tnb_eval_stats": (i32, [vp, vp, vp, i32, i32, i32, vp, vp, vp]),
    "tnb_temporal_ensemble": (i32, [vp, vp, vp, C.POINTER(C.c_float), i32, C.c_longlong, i32, i32, i32, i32, vp]),
    "tnb_tracknet_workspace_bytes": (sz, [C.POINTER(TrackNetCfg)]),
    "tnb_tracknet_forward": (i32, [C.POINTER(TrackNetCfg), vp, C.POINTER(vp), vp, vp, sz, vp]),
    "tnb_tracknet_backward": (i32, [C.POINTER(TrackNetCfg), vp, vp, C.POINTER(vp), C.POINTER(vp), vp, sz, vp]),
    "tnb_tracknet_backward_range": (i32, [C.POINTER(TrackNetCfg), vp, vp, C.POINTER(vp), C.POINTER(vp), vp, sz, i32, i32, vp]),
    "tnb_tracknet_grad_split_layer": (i32, []),
    "tnb_set_graph_replay": (i32, [i32]),
    "tnb_graph_stats": (i32, [C.POINTER(C.c_longlong)]),
    "tnb_tracknet_num_launches": (i32, [C.POINTER(TrackNetCfg), i32]),
    "tnb_tracknet_debug_layer": (i32, [C.POINTER(TrackNetCfg), vp, i32, C.POINTER(vp), C.POINTER(i32)]),
    "tnb_profile_enable": (i32, [i32]),
    "tnb_profile_collect": (i32, [i32, vp, vp]),
}

_lib = None


def load():
    """Load the shared library once; raise with build instructions if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TnbError(
            f"{LIB_PATH} not found: the sm_100a extension is not built. Run "
            "`python -c \"import __graft_entry__ as g; g.build()\"` (or `make -C tracknetv3_b200/csrc`). "
            "There is no CPU / PyTorch fallback for this path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.tnb_abi_version() != 2:
        raise TnbError("libtracknet_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().tnb_last_error().decode("utf-8", "replace")
        raise TnbError(f"tracknet_b200 call failed (rc={rc}): {msg}")


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors):
    """The reference runs anywhere torch does; this path is CUDA-only by design (no CPU fallback)."""
    for t in tensors:
        if not t.is_cuda:
            raise TnbError("tracknet_b200: CPU tensor passed to a CUDA-only op (there is no CPU fallback)")


def ptr_array(tensors):
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr
