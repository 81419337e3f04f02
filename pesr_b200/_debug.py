"""ctypes binding of the bring-up hooks (include/pesr_b200_debug.h).  They exist only in
pesr_b200/libpesr_b200_debug.so (tools/build_debug.sh); run the tools that use them with
PESR_B200_LIB=pesr_b200/libpesr_b200_debug.so.  The product library exports none of these symbols."""
import ctypes as C

from ._lib import LIB_PATH, lib

_vp, _i32, _i64 = C.c_void_p, C.c_int32, C.c_int64

DEBUG_SIGNATURES = {
    "pesr_debug_timeline": (None, [_vp]),
    "pesr_debug_wgrad_timeline": (None, [_vp]),
    "pesr_debug_wgrad_desc": (None, [C.c_int, C.c_int]),
    "pesr_debug_mma_rate": (C.c_int, [_i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "pesr_debug_sm_hog": (C.c_int, [_i32, _i64, _vp]),
}


def bind():
    """Sets the ctypes signatures of the debug hooks on the loaded library; raises if it is the product build."""
    for name, (res, args) in DEBUG_SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            raise ImportError(f"{LIB_PATH} has no {name}: build the debug library with tools/build_debug.sh and set "
                              "PESR_B200_LIB=pesr_b200/libpesr_b200_debug.so") from None
        fn.restype, fn.argtypes = res, args
    return lib
