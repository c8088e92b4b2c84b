"""ctypes binding of libdmvs_b200.so (declared in include/dmvs_b200.h).

The library is the product; there is no Python/torch fallback for any entry point.
If it is missing or fails to load, every op raises ``NativeLibraryError``.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_longlong, c_size_t, c_ulonglong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdmvs_b200.so")

ABI_VERSION = 20
FMT_F32, FMT_CH16, FMT_CH16P = 0, 1, 2
FMT_NHWC2 = 4
FMT_NHWC2_F16 = 5
ENGINE_FP32, ENGINE_TENSOR = 0, 1
MAX_SRC = 16
REGNET_LAYERS = 11
LAYER_NAMES = ("conv0", "conv1", "conv2", "conv3", "conv4", "conv5", "conv6", "conv7", "conv9", "conv11", "prob")


class NativeLibraryError(RuntimeError):
    pass


class ConvLayer(ctypes.Structure):
    _fields_ = [("w", c_void_p), ("scale", c_void_p), ("shift", c_void_p), ("w_tc", c_void_p), ("w_tc_kd", c_void_p), ("w_tc_kw", c_void_p)]


class RegnetBranch(ctypes.Structure):
    _fields_ = [("layer", ConvLayer * REGNET_LAYERS), ("conv0_pair", ConvLayer)]


# name -> (restype, argtypes); mirrors include/dmvs_b200.h one to one
SIGNATURES = {
    "dmvs_abi_version": (c_int, []),
    "dmvs_last_error": (c_char_p, []),
    "dmvs_launch_count": (c_ulonglong, []),
    "dmvs_debug_set": (c_int, [c_char_p, c_int]),
    "dmvs_debug_set_ptr": (c_int, [c_char_p, c_void_p]),
    "dmvs_warp_corr_f32": (c_int, [c_void_p, c_longlong, POINTER(c_void_p), c_longlong, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "dmvs_warp_corr_nhwc_f32": (c_int, [c_void_p, c_longlong, c_int, POINTER(c_void_p), c_longlong, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "dmvs_warp_corr_staged_f32": (c_int, [c_void_p, c_longlong, c_int, POINTER(c_void_p), c_longlong, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                          c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "dmvs_warp_corr_flag_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "dmvs_warp_corr_h16_f32": (c_int, [c_void_p, c_longlong, c_int, POINTER(c_void_p), c_longlong, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "dmvs_features_nhwc_f16": (c_int, [c_void_p, c_longlong, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "dmvs_warp_corr_backward_f32": (c_int, [c_void_p, c_longlong, c_int, POINTER(c_void_p), c_longlong, c_int, c_int, c_void_p, c_void_p,
                                            c_void_p, c_void_p, POINTER(c_void_p), c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "dmvs_features_nhwc_f32": (c_int, [c_void_p, c_longlong, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "dmvs_features_s2d_cells_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "dmvs_conv2d_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "dmvs_regnet_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "dmvs_regnet_forward_f32": (c_int, [POINTER(RegnetBranch), c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                        c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "dmvs_regnet_forward_branches_f32": (c_int, [POINTER(RegnetBranch), c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                                 c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "dmvs_conv3d_f32": (c_int, [c_void_p, POINTER(ConvLayer), c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "dmvs_convert_layout": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "dmvs_conv3d_ch16": (c_int, [c_void_p, c_int, POINTER(ConvLayer), c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                 c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "dmvs_geo_consistency_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_float, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_void_p, c_void_p]),
    "dmvs_geo_consistency_dynamic_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_double, c_double, c_void_p, c_void_p,
                                                 c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dmvs_backproject_world_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "dmvs_depth_head_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_int, c_int, c_int, c_int, c_void_p]),
    "dmvs_refine_head_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p,
                                     c_int, c_int, c_int, c_void_p]),
    "dmvs_hypotheses_first_f32": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "dmvs_hypotheses_next_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                         c_int, c_void_p]),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load the shared library once; raise loudly if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryError(
            "libdmvs_b200.so is not built (%s). Run `python -m dmvsnet_b200.build` (needs nvcc); "
            "there is no CPU/PyTorch fallback for the cost-volume path." % LIB_PATH)
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as e:  # pragma: no cover - depends on the machine
        raise NativeLibraryError("cannot load %s: %s" % (LIB_PATH, e)) from e
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise NativeLibraryError("%s does not export %s (stale build?)" % (LIB_PATH, name)) from e
        fn.restype = res
        fn.argtypes = args
    if lib.dmvs_abi_version() != ABI_VERSION:
        raise NativeLibraryError("ABI version mismatch: library %d, binding %d" % (lib.dmvs_abi_version(), ABI_VERSION))
    # experiments on the GPU box: DMVS_DEBUG_SET="key=value,key=value" applies dmvs_debug_set knobs at load time
    for kv in filter(None, os.environ.get("DMVS_DEBUG_SET", "").split(",")):
        key, _, val = kv.partition("=")
        if lib.dmvs_debug_set(key.strip().encode(), int(val)) != 0:
            raise NativeLibraryError("DMVS_DEBUG_SET: %r rejected (%s)" % (kv, lib.dmvs_last_error().decode()))
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().dmvs_last_error()
        raise RuntimeError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))


def launch_count() -> int:
    return int(load().dmvs_launch_count())
