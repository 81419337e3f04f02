"""ctypes binding of the pesr_b200 C-ABI (include/pesr_b200.h).

The shared library is built in-tree by ``__graft_entry__.build()`` (plain nvcc, sm_100a).  There is
no CPU fallback: if the library is missing the import fails loudly, and every entry point raises
``PesrError`` carrying ``pesr_last_error()`` on a non-zero return code.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PESR_B200_LIB: A/B a differently built library of the same ABI (tools/ab_*.sh); default = the in-tree build
LIB_PATH = os.environ.get("PESR_B200_LIB") or os.path.join(_HERE, "libpesr_b200.so")

MAX_TAPS = 9
MAX_SRC = 4

DT_F16, DT_BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_LRELU = 0, 1, 2
OUT_NORMAL, OUT_SHUFFLE2, OUT_UNSHUFFLE2 = 0, 1, 2
WMAP_OIHW, WMAP_OIHW_PS, WMAP_COL_IN, WMAP_COL_OUT = 0, 1, 2, 3


class PesrError(RuntimeError):
    pass


class ConvDesc(C.Structure):
    """Mirror of ``pesr_conv_desc``."""
    _fields_ = [
        ("dtype", C.c_int32), ("nb", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
        ("cin", C.c_int32), ("cout", C.c_int32), ("block_n", C.c_int32),
        ("tile_h", C.c_int32), ("tile_w", C.c_int32), ("ntaps", C.c_int32),
        ("tap_dh", C.c_int8 * MAX_TAPS), ("tap_dw", C.c_int8 * MAX_TAPS),
        ("tap_src", C.c_int8 * MAX_TAPS), ("tap_widx", C.c_int8 * MAX_TAPS),
        ("src", C.c_void_p * MAX_SRC),
        ("src_h", C.c_int32 * MAX_SRC), ("src_w", C.c_int32 * MAX_SRC),
        ("src_sn", C.c_int64 * MAX_SRC), ("src_sh", C.c_int64 * MAX_SRC), ("src_sw", C.c_int64 * MAX_SRC),
        ("nsrc", C.c_int32),
        ("wpacked", C.c_void_p), ("w_rows", C.c_int32),
        ("bias", C.c_void_p), ("alpha", C.c_float), ("alpha_dev", C.c_void_p),
        ("res32", C.c_void_p), ("ld_res32", C.c_int32),
        ("res16", C.c_void_p), ("ld_res16", C.c_int32),
        ("act", C.c_int32),
        ("mask16", C.c_void_p), ("ld_mask16", C.c_int32), ("mask_mode", C.c_int32),
        ("out32", C.c_void_p), ("ld_out32", C.c_int32),
        ("out16", C.c_void_p), ("ld_out16", C.c_int32),
        ("out_mode", C.c_int32), ("out_h", C.c_int32), ("out_w", C.c_int32),
        ("out_sy", C.c_int32), ("out_sx", C.c_int32), ("out_oy", C.c_int32), ("out_ox", C.c_int32),
        ("out_coff", C.c_int32), ("ps_c", C.c_int32), ("aux_mode", C.c_int32),
        ("ksplit", C.c_int32), ("b_mn_major", C.c_int32), ("split_stride32", C.c_int64),
        ("tile_n", C.c_int32), ("reserved0", C.c_int32),
        ("bn_sums", C.c_void_p),
        ("ncls", C.c_int32), ("cls_ntaps", C.c_int8 * 4), ("cls_oy", C.c_int8 * 4), ("cls_ox", C.c_int8 * 4),
    ]


class WgradDesc(C.Structure):
    """Mirror of ``pesr_wgrad_desc``."""
    _fields_ = [
        ("dtype", C.c_int32), ("nb", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
        ("m_total", C.c_int32), ("n_total", C.c_int32), ("block_m", C.c_int32), ("block_n", C.c_int32),
        ("ntaps", C.c_int32),
        ("tap_dh", C.c_int8 * MAX_TAPS), ("tap_dw", C.c_int8 * MAX_TAPS), ("tap_src", C.c_int8 * MAX_TAPS),
        ("a", C.c_void_p), ("a_c", C.c_int32),
        ("b", C.c_void_p * MAX_SRC),
        ("b_h", C.c_int32 * MAX_SRC), ("b_w", C.c_int32 * MAX_SRC),
        ("b_sn", C.c_int64 * MAX_SRC), ("b_sh", C.c_int64 * MAX_SRC), ("b_sw", C.c_int64 * MAX_SRC),
        ("nsrc", C.c_int32), ("splits", C.c_int32),
        ("partials", C.c_void_p), ("partials_elems", C.c_int64),
        ("out_mul", C.c_float), ("out_div_dev", C.c_void_p),
    ]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"pesr_b200: {LIB_PATH} is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU or PyTorch fallback for the hot path.")
    return C.CDLL(LIB_PATH)


lib = _load()

_vp, _i32, _i64, _f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float

# name -> (restype, argtypes). This table is also what tests/test_abi.py checks against the header.
SIGNATURES = {
    "pesr_last_error": (C.c_char_p, []),
    "pesr_version": (C.c_int, []),
    "pesr_launch_count": (C.c_longlong, [C.c_int]),
    "pesr_sizeof": (C.c_int, [C.c_int]),
    "pesr_profile_enable": (None, [C.c_int]),
    "pesr_profile_read": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_longlong), C.POINTER(C.c_double)]),
    "pesr_conv_igemm": (C.c_int, [C.POINTER(ConvDesc), _vp]),
    "pesr_set_option": (C.c_int, [C.c_int, C.c_int]),
    "pesr_conv_wgrad": (C.c_int, [C.POINTER(WgradDesc), C.POINTER(_i32), _vp]),
    "pesr_wgrad_reduce": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _f32, _vp, _i32, _vp, _vp]),
    "pesr_wgrad_reduce_bias": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _f32, _vp, _i32, _vp,
                                         _vp, _i64, _i32, _i32, _f32, _i32, _vp, _vp, _i32, _vp]),
    "pesr_pack_weights": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "pesr_pack_weights_multi": (C.c_int, [_vp, _i32, _vp, _i32, _i32, _vp]),
    "pesr_im2col3": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp]),
    "pesr_col2im3": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _f32, _vp, _i32, _vp, _vp, _vp]),
    "pesr_nchw32_to_nhwc16": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _vp, _vp]),
    "pesr_nhwc16_to_nchw32": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _f32, _vp, _i32, _vp, _vp]),
    "pesr_colsum16": (C.c_int, [_vp, _i64, _i32, _i32, _f32, _vp, _i32, _i32, _vp, _vp]),
    "pesr_amax_scale": (C.c_int, [_vp, _i64, _f32, _vp, _vp]),
    "pesr_moments3": (C.c_int, [_vp, _vp, _i32, _i64, _vp, _vp]),
    "pesr_blend_x8_to_u8": (C.c_int, [_vp, _vp, _i32, _i32, _f32, _i32, _vp, _vp, _vp]),
    "pesr_u8hwc_to_f32nchw": (C.c_int, [_vp, _i32, _i32, _vp, _vp]),
    "pesr_u8hwc_to_f32nchw_batch": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _vp]),
    "pesr_col2im3_tiled": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _f32, _vp, _i32, _vp, _vp, _vp, _vp]),
    "pesr_mean_shift": (C.c_int, [_vp, _i32, _i64, _vp, _vp, _vp, _vp]),
    "pesr_allreduce_p2p": (C.c_int, [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), _i32, _i32, C.c_uint64, _i64, _i64, _f32,
                                     _i32, C.c_uint32, _vp]),
    "pesr_split16": (C.c_int, [_vp, _i64, _i32, _vp, _vp, _i32, _f32, _vp, _i32, _vp, _vp, _vp]),
    "pesr_colmoments32": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _vp]),
    "pesr_affine_split": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp]),
    "pesr_maxpool2_f32_fwd": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    "pesr_maxpool2_f32_bwd": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "pesr_psnr_y_sse": (C.c_int, [_vp, _vp, _i32, _i64, _vp, _vp]),
    "pesr_gather_patches": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _vp, _vp]),
    "pesr_loss_l1": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _vp]),
    "pesr_loss_mse": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _vp]),
    "pesr_loss_tv": (C.c_int, [_vp, _i64, _i32, _i32, _vp, _vp, _vp]),
    "pesr_loss_gan": (C.c_int, [_vp, _vp, _i32, _f32, _f32, _f32, _i32, _f32, _vp, _vp, _vp, _vp]),
    "pesr_bn_reduce": (C.c_int, [_vp, _i64, _i32, _i32, _vp, _i32, _i32, _vp]),
    "pesr_bn_lrelu_fwd": (C.c_int, [_vp, _i64, _i32, _i32, _vp, _f32, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f32,
                                    _i32, _vp, _vp]),
    "pesr_bn_lrelu_bwd": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _i32, _f32, _vp, _i32, _vp, _vp, _vp,
                                    _i32, _vp]),
    "pesr_maxpool2_fwd": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "pesr_maxpool2_bwd": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "pesr_linear_workspace_floats": (C.c_int64, [_i32, _i32, _i32]),
    "pesr_linear_skinny_fwd": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "pesr_linear_finalize": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _i32, _i32, _vp, _vp, _vp]),
    "pesr_linear_skinny_dgrad": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    "pesr_linear_skinny_wgrad": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _f32, _vp, _i32, _i32, _vp, _vp]),
    "pesr_cast16": (C.c_int, [_vp, _i64, _i32, _vp, _vp]),
    "pesr_flatten_nchw16": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _vp]),
    "pesr_unflatten_nchw16": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _f32, _vp, _f32, _i32, _vp, _vp]),
    "pesr_adam_multi": (C.c_int, [_vp, _i32, _f32, _f32, _f32, _f32, _i32, _f32, _vp]),
    "pesr_adam_multi_dev": (C.c_int, [_vp, _i32, _vp, _f32, _f32, _f32, _vp, _f32, _vp]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here == the .so does not export what the header declares
    _fn.restype = _res
    _fn.argtypes = _args

if lib.pesr_sizeof(0) != C.sizeof(ConvDesc) or lib.pesr_sizeof(1) != C.sizeof(WgradDesc):
    raise ImportError("pesr_b200: ctypes struct layout does not match include/pesr_b200.h "
                      f"(conv {C.sizeof(ConvDesc)} vs {lib.pesr_sizeof(0)}, wgrad {C.sizeof(WgradDesc)} vs {lib.pesr_sizeof(1)})")


def check(rc, what=""):
    if rc != 0:
        msg = lib.pesr_last_error().decode("utf-8", "replace")
        raise PesrError(f"{what or 'pesr_b200'} failed (rc={rc}): {msg}")


def profile_enable(on=True):
    lib.pesr_profile_enable(1 if on else 0)


def profile_read(kind):
    """-> (total kernel ms, launches, algorithmic FLOPs) of the tensor-core launches recorded since the last read."""
    ms, n, fl = C.c_double(0), C.c_longlong(0), C.c_double(0)
    check(lib.pesr_profile_read(kind, C.byref(ms), C.byref(n), C.byref(fl)), "pesr_profile_read")
    return ms.value, n.value, fl.value


OPT_PAIR_MODE, OPT_SUB_STAGES, OPT_PDL, OPT_STAGED_EPILOGUE, OPT_SPECIALISED_EPILOGUE = 0, 1, 2, 3, 4
OPT_RESERVE_SMS, OPT_RESIDENT_WEIGHTS = 5, 6


def set_option(option, value):
    """Kernel-selection options of pesr_conv_igemm (include/pesr_b200.h PESR_OPT_*)."""
    check(lib.pesr_set_option(option, value), "pesr_set_option")


def launch_count(reset=False):
    return int(lib.pesr_launch_count(1 if reset else 0))
