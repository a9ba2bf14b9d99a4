"""ctypes binding of libhowl_b200.so (the C ABI in include/howl_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import re
from typing import Dict, List

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libhowl_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "howl_b200.h")
DEBUG_HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "howl_b200_debug.h")   # tuning aids / test hooks

HOWL_OK = 0
FE_TIME_MAJOR, FE_MELS_ONLY, FE_STACKED, FE_ZMUV = 0x1, 0x2, 0x4, 0x10


class HowlB200Error(RuntimeError):
    pass


class FrontendCfg(C.Structure):
    _fields_ = [("sample_rate", C.c_int32), ("n_fft", C.c_int32), ("hop", C.c_int32), ("n_mels", C.c_int32)]


_vp, _i32, _i64, _u32, _f32, _sz = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_float, C.c_size_t

# name -> (restype, argtypes); mirrors include/howl_b200.h one to one
SIGNATURES: Dict[str, tuple] = {
    "howl_b200_abi_version": (C.c_int, []),
    "howl_b200_create": (C.c_int, [C.c_int, C.POINTER(FrontendCfg), C.POINTER(_vp)]),
    "howl_b200_destroy": (None, [_vp]),
    "howl_b200_last_error": (C.c_char_p, [_vp]),
    "howl_b200_sm_count": (C.c_int, [_vp]),
    "howl_b200_launch_count": (_i64, [_vp]),
    "howl_b200_set_option": (C.c_int, [_vp, C.c_char_p, _i64]),
    "howl_b200_selftest_umma": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32]),
    "howl_b200_debug_stream_profile": (C.c_int, [_vp, _vp, _i32]),
    "howl_b200_debug_umma_bench": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp]),
    "howl_b200_debug_mbn_workspace_bytes": (_i64, [_i64, C.c_int, C.c_int]),
    "howl_b200_debug_mbn_gemm": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _vp, _sz]),
    "howl_b200_debug_mbn_wgrad": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _vp, _sz]),
    "howl_b200_mobilenet_debug_mask_bytes": (_i64, [_i64, _i32, _i32]),
    "howl_b200_mobilenet_debug_masks": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _sz, _vp]),
    "howl_b200_res8_debug_masks": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _sz, _vp, _vp]),
    "howl_b200_profile_begin": (C.c_int, [_vp, _vp]),
    "howl_b200_profile_end": (C.c_int, [_vp, _vp, _sz, _vp, _i32]),
    "howl_b200_num_frames": (_i64, [_i64, _i32]),
    "howl_b200_compute_lengths": (C.c_int, [_vp, _i64, _i32, _i32, _vp]),
    "howl_b200_frontend_fwd": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _vp, _f32, _f32, _vp, _u32, _vp]),
    "howl_b200_deltas_fwd": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _vp]),
    "howl_b200_sum_sumsq": (C.c_int, [_vp, _vp, _vp, _i64, _vp]),
    "howl_b200_zmuv_fwd": (C.c_int, [_vp, _vp, _vp, _i64, _f32, _f32, _vp]),
    "howl_b200_spec_mask": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp]),
    "howl_b200_to_time_major": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp]),
    "howl_b200_batch_gather": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp]),
    "howl_b200_batch_gather_aug": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, C.c_uint64, _vp]),
    "howl_b200_res8_param_count": (_i64, [_i32]),
    "howl_b200_res8_workspace_bytes": (_i64, [_i64, _i32, _i32, _i32, C.c_int]),
    "howl_b200_res8_fwd": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp, C.c_int, _vp, _vp, _sz]),
    "howl_b200_res8_bwd": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i64, _vp, _vp, _vp, _vp, _sz]),
    "howl_b200_res8_bwd_dlogits": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _sz]),
    "howl_b200_lstm_param_count": (_i64, [_i32, _i32]),
    "howl_b200_lstm_workspace_bytes": (_i64, [_i64, _i32, _i32, _i32, C.c_int, C.c_int]),
    "howl_b200_lstm_fwd": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp, _vp, _vp, C.c_int, C.c_int, _vp, _vp, _sz]),
    "howl_b200_lstm_bwd": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _i64, _vp, _vp, _vp, _vp, _sz]),
    "howl_b200_lstm_bwd_dlogits": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, _i64, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _sz]),
    "howl_b200_lstm_ctc_bwd": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i64, _i32, _i32, _i32, _i32, _i64, _vp, _vp, _vp,
                                         _vp, _sz]),
    "howl_b200_seq_lstm_ctc_train_step": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _vp, _i64, _i64, _vp, _f32, _f32, _i32,
                                                    _i32, _vp, _vp, _vp, _vp, _vp, _i64, _f32, _f32, _vp, _vp, _vp, _sz]),
    "howl_b200_lstm_train_step": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _f32, _f32, _i32, _i32, _vp, _vp, _vp,
                                            _vp, _i64, _f32, _f32, _vp, _vp, _vp, _sz]),
    "howl_b200_mobilenet_param_count": (_i64, [_i32]),
    "howl_b200_mobilenet_bn_channels": (_i64, []),
    "howl_b200_mobilenet_bn_layers": (_i64, []),
    "howl_b200_mobilenet_workspace_bytes": (_i64, [_i64, _i32, _i32, _i32]),
    "howl_b200_mobilenet_fwd": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp, C.c_int, _f32, C.c_uint64, _vp, _vp, _sz]),
    "howl_b200_mobilenet_bwd": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i64, _vp, _vp, _f32, C.c_uint64, _vp, _vp, _sz]),
    "howl_b200_mobilenet_bwd_dlogits": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _f32, C.c_uint64, _vp, _sz]),
    "howl_b200_mobilenet_train_step": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i64, _vp, _f32, _f32, _i32, _vp, _vp, _vp, _vp, _vp, _vp,
                                                 _i64, _f32, _f32, _f32, C.c_uint64, _vp, _vp, _vp, _sz]),
    "howl_b200_las_param_count": (_i64, [_i32, _i32]),
    "howl_b200_las_workspace_bytes": (_i64, [_i64, _i32, _i32, _i32, C.c_int]),
    "howl_b200_las_lengths": (C.c_int, [_vp, _i64, _vp]),
    "howl_b200_las_fwd": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp, C.c_int, C.c_float, C.c_uint64, _vp, _vp, _sz]),
    "howl_b200_las_bwd": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i64, _vp, _vp, C.c_float, _vp, _vp, _sz]),
    "howl_b200_las_bwd_dlogits": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp, C.c_float, _vp, _sz]),
    "howl_b200_adamw": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _f32, _f32, _f32, _f32, _f32]),
    "howl_b200_res8_train_step": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i64, _vp, _f32, _f32, _vp, _i32, _vp, _vp, _vp,
                                            _vp, _vp, _vp, _i64, _f32, _f32, _vp, _vp, _vp, _sz]),
}

_lib = None


def header_symbols(debug: bool = True) -> List[str]:
    """Every function name declared in include/howl_b200.h (+ include/howl_b200_debug.h when `debug`)."""
    names = set()
    for path in [HEADER_PATH] + ([DEBUG_HEADER_PATH] if debug else []):
        text = re.sub(r"/\*.*?\*/", "", open(path).read(), flags=re.S)
        names |= set(re.findall(r"\b(howl_b200_[a-z0-9_]+)\s*\(", text))
    return sorted(names)


def load():
    """Load the shared library (built in-tree by `make` / __graft_entry__.build()); raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HowlB200Error(
            f"{LIB_PATH} not found: build it with `make` (or __graft_entry__.build()). howl_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.howl_b200_abi_version() != 1:
        raise HowlB200Error("libhowl_b200.so ABI version mismatch")
    _lib = lib
    return lib
