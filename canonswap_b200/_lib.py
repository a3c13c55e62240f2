"""ctypes binding of libcanonswap_b200.so -- the C ABI declared in include/canonswap_b200.h.

There is deliberately no fallback: if the shared library is missing the import of the product
path fails with an explicit error (build it with `python -m canonswap_b200._build`).
"""
from __future__ import annotations

import ctypes as C
import os

from ._build import LIB

CS_F32, CS_I64, CS_U8 = 0, 1, 2
CS_FRAME_IN_U8_HWC = 1
CS_FRAME_DEBUG_DECODES = 2
CS_FRAME_V2I = 4
CS_OPT_CONV_IMPL = 1
CS_OPT_USE_GRAPH = 2
CS_OPT_TC_PASSES = 3
CS_OPT_TC_SETS = 4
CS_OPT_TC_COMP = 5
CS_OPT_TC_PAIR = 6
CS_OPT_TC_STACKED3 = 7
CS_OPT_TC_DOUBLE_BUFFER = 8
CS_OPT_TC_BN_MAX = 9
CS_OPT_LANES = 10
CS_OPT_WINOGRAD = 13
CS_OPT_TC_POSCOMP = 14
CS_OPT_TEST_AMUL = 15
CS_FRAME_MOTION = 8
CS_FRAME_V2I_FEATURE = 16
MOTION_HEADS = 328
PASTE_MAX_BATCH = 16
CS_OPT_TC_CHAIN_MAX = 12
CS_OPT_TC_SINGLE_CHAIN = 11

# every symbol include/canonswap_b200.h declares
SYMBOLS = [
    "cs_create", "cs_destroy", "cs_last_error", "cs_set_option", "cs_launch_count", "cs_workspace_bytes",
    "cs_load_weights", "cs_set_identity", "cs_appearance", "cs_warp", "cs_warp_out", "cs_warp_forward",
    "cs_swap", "cs_refine", "cs_spade", "cs_frame", "cs_profile", "cs_profile_read", "cs_profile_dump", "cs_motion", "cs_keypoints", "cs_paste_back", "cs_soft_erosion", "cs_parse_mask", "cs_calibrate", "cs_test_conv", "cs_test_grid_sample3d",
    "cs_test_instance_stats",
]


class TensorDesc(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("dtype", C.c_int32), ("ndim", C.c_int32),
                ("shape", C.c_int64 * 6)]


_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    # CANONSWAP_B200_LIB: another build of the SAME library (same-box A/B measurements); never a different implementation
    path = os.environ.get("CANONSWAP_B200_LIB", LIB)
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is missing. canonswap_b200 has no CPU / PyTorch fallback: build the CUDA library first "
            "(python -m canonswap_b200._build, or __graft_entry__.build()).")
    lib = C.CDLL(path)
    p, i, f, vp = C.c_void_p, C.c_int, C.c_float, C.c_void_p
    lib.cs_create.argtypes = [C.POINTER(vp), i, i, i, i]
    lib.cs_destroy.argtypes = [vp]
    lib.cs_destroy.restype = None
    lib.cs_last_error.argtypes = [vp]
    lib.cs_last_error.restype = C.c_char_p
    lib.cs_set_option.argtypes = [vp, i, i]
    lib.cs_launch_count.argtypes = [vp]
    lib.cs_launch_count.restype = C.c_int64
    lib.cs_workspace_bytes.argtypes = [vp]
    lib.cs_workspace_bytes.restype = C.c_size_t
    lib.cs_load_weights.argtypes = [vp, C.POINTER(TensorDesc), i]
    lib.cs_set_identity.argtypes = [vp, p, vp]
    lib.cs_appearance.argtypes = [vp, p, p, i, vp]
    lib.cs_warp.argtypes = [vp, p, p, p, p, p, p, i, vp]
    lib.cs_warp_out.argtypes = [vp, p, p, p, i, vp]
    lib.cs_warp_forward.argtypes = [vp, p, p, p, p, p, p, i, vp]
    lib.cs_swap.argtypes = [vp, p, p, p, i, vp]
    lib.cs_refine.argtypes = [vp, p, p, i, vp]
    lib.cs_spade.argtypes = [vp, p, p, p, i, vp]
    lib.cs_frame.argtypes = [vp, p, p, p, p, p, i, i, vp]
    lib.cs_motion.argtypes = [vp, p, p, i, vp]
    lib.cs_keypoints.argtypes = [vp, p, p, p, p, p, i, vp]
    lib.cs_paste_back.argtypes = [vp, p, p, C.POINTER(C.c_double), p, p, i, i, i, i, i, vp]
    lib.cs_soft_erosion.argtypes = [vp, p, p, p, p, i, i, i, i, f, i, vp]
    lib.cs_parse_mask.argtypes = [vp, p, i, i, i, i, i, i, C.c_uint64, p, p, vp]
    lib.cs_calibrate.argtypes = [vp, i, p, i]
    lib.cs_profile.argtypes = [vp, i]
    lib.cs_profile_read.argtypes = [vp, C.POINTER(C.c_double)]
    lib.cs_profile_dump.argtypes = [vp, C.c_char_p, i]
    lib.cs_test_conv.argtypes = [vp, p, p, p, p] + [i] * 13 + [f, i, vp]
    lib.cs_test_grid_sample3d.argtypes = [vp, p, p, p, i, i, i, i, i, vp]
    lib.cs_test_instance_stats.argtypes = [vp, p, p, p, i, i, i, f, vp]
    for name in SYMBOLS:
        fn = getattr(lib, name)          # AttributeError here == header / library mismatch
        if name not in ("cs_destroy", "cs_last_error", "cs_launch_count", "cs_workspace_bytes"):
            fn.restype = C.c_int
    _lib = lib
    return lib
