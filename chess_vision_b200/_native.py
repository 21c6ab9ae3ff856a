"""ctypes binding of libchessvision_b200.so (the C-ABI in include/chessvision_b200.h).

The library is the product: if it cannot be loaded this module raises -- there is no Python/PyTorch
fallback for any compute entry point.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# CV_B200_LIB: an experiment build of the same library (tools/, -DCV_FE_PROFILE ...); never a fallback
LIB_PATH = os.environ.get("CV_B200_LIB") or os.path.join(_HERE, "libchessvision_b200.so")

PRECISION_FP32, PRECISION_BF16, PRECISION_FP16, PRECISION_FP32_SPLIT = 0, 1, 2, 3
LAYOUT_HWC, LAYOUT_CHW = 0, 1
FEN_STRIDE = 80
PRECISIONS = {"fp32": PRECISION_FP32, "float32": PRECISION_FP32, "bf16": PRECISION_BF16, "bfloat16": PRECISION_BF16,
              "fp16": PRECISION_FP16, "float16": PRECISION_FP16, "fp32_split": PRECISION_FP32_SPLIT, "split": PRECISION_FP32_SPLIT}
# cv_square_set_impl bits (include/chessvision_b200.h; kept in step by tests/test_boundary.py)
IMPL_POINTWISE_UMMA, IMPL_DENSE_UMMA, IMPL_DEPTHWISE_VEC, IMPL_SPLIT_WEIGHTS = 1, 2, 4, 8
IMPL_FRONTEND, IMPL_TAIL, IMPL_MID, IMPL_EARLY, IMPL_FRONTEND3 = 16, 32, 64, 128, 512
IMPL_DEFAULT, IMPL_ALL = 1023, 2047


class NativeError(RuntimeError):
    pass


class LayerInfo(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("kind", "cin", "cout", "k", "stride", "relu", "hin", "hout", "skip")] + \
               [("w_offset", C.c_int64), ("b_offset", C.c_int64)]


class NamedTensor(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("numel", C.c_int64)]


_lib = None


def _sig(fn, res, *args):
    fn.restype = res
    fn.argtypes = list(args)


def lib():
    """Load and return the ctypes library; the default in-tree library is (re)built first when it is missing or older than a
    source under csrc/ and nvcc is available (a prebuilt library on a box without nvcc is used as it is)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.environ.get("CV_B200_LIB"):
        from . import _build
        try:
            _build.build()                       # no-op unless stale
        except RuntimeError:
            if not os.path.exists(LIB_PATH):
                raise
    try:
        L = C.CDLL(LIB_PATH)
    except OSError as e:  # pragma: no cover
        raise NativeError(f"cannot load {LIB_PATH}: {e} (chess_vision_b200 has no CPU fallback)") from e
    vp, i32, i64, sz, u32 = C.c_void_p, C.c_int, C.c_int64, C.c_size_t, C.c_uint32
    _sig(L.cv_last_error, C.c_char_p)
    _sig(L.cv_abi_version, i32)
    _sig(L.cv_num_layers, i32)
    _sig(L.cv_layer_info_get, i32, i32, C.POINTER(LayerInfo))
    _sig(L.cv_weight_blob_floats, sz)
    _sig(L.cv_square_create, i32, i32, C.POINTER(vp))
    _sig(L.cv_square_destroy, i32, vp)
    _sig(L.cv_square_load_weights, i32, vp, vp, sz, vp)
    _sig(L.cv_square_set_norm_lut, i32, vp, vp)
    _sig(L.cv_square_set_wave, i32, vp, i32)
    _sig(L.cv_square_set_impl, i32, vp, i32)
    _sig(L.cv_square_workspace_bytes, sz, vp, i32, i32, i32)
    _sig(L.cv_square_forward_f32, i32, vp, vp, i32, i32, i32, vp, vp, vp, vp, vp, sz, vp)
    _sig(L.cv_square_forward_u8, i32, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, sz, vp)
    _sig(L.cv_square_fen, i32, vp, vp, vp, vp, i32, vp, vp, vp)
    _sig(L.cv_square_predict_u8, i32, vp, vp, i32, vp, i32, i32, i32, vp, vp, vp, sz, vp)
    _sig(L.cv_square_predict_host_u8, i32, vp, vp, i32, vp, i32, i32, i32, vp, vp)
    _sig(L.cv_combine_type_color, i32, vp, vp, i64, vp, vp)
    _sig(L.cv_crop_squares_f32, i32, vp, i32, i32, vp, vp)
    _sig(L.cv_crop_squares_u8, i32, vp, i32, i32, i32, vp, vp)
    _sig(L.cv_crop_index_table, i32, i32, vp, vp, vp)
    _sig(L.cv_square_set_tap, i32, vp, i32, vp, sz)
    _sig(L.cv_synth_boards, i32, vp, i32, i64, i32, i32, u32, i32, vp, vp)
    _sig(L.cv_synth_boards_host, i32, vp, i32, i64, i32, i32, u32, i32, vp)
    _sig(L.cv_fen_from_classes_host, i32, vp, C.c_float, vp, vp)
    _sig(L.cv_square_launch_count, i64, vp)
    _sig(L.cv_jpeg_info, i32, vp, sz, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int))
    _sig(L.cv_jpeg_decode_batch, i32, vp, vp, i32, i32, i32, vp, i32, vp)
    _sig(L.cv_jpeg_decode_coefficients_host, i32, vp, sz, vp, sz, vp)
    _sig(L.cv_jpeg_decode_coefficients_host_chunked, i32, vp, sz, vp, sz, i32, i32, vp)
    _sig(L.cv_square_pack_weights, i32, vp, i32, vp, sz)
    _sig(L.cv_square_fp16_status, i32, vp, C.POINTER(C.c_int), C.POINTER(C.c_int))
    _sig(L.cv_square_profile, i32, vp, i32)
    _sig(L.cv_square_profile_read, i32, vp, vp, vp)
    _sig(L.cv_eval_accumulate, i32, vp, vp, vp, vp, vp, vp, vp, i32, vp, vp, vp, vp)
    _sig(L.cv_resize_bilinear_u8, i32, vp, i32, i32, i32, vp, i32, i32, vp)
    _sig(L.cv_resize_coeffs_host, i32, i32, i32, C.POINTER(C.c_int), vp, vp, i32)
    if L.cv_abi_version() != 1:
        raise NativeError("libchessvision_b200.so ABI version mismatch")
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise NativeError(f"chessvision_b200 native call failed ({rc}): {lib().cv_last_error().decode()}")


def layer_table():
    L = lib()
    out = []
    for i in range(L.cv_num_layers()):
        info = LayerInfo()
        check(L.cv_layer_info_get(i, C.byref(info)))
        out.append(info)
    return out


def ptr(t):
    """Device/host pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(device):
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
