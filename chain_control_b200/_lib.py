"""ctypes binding of the C ABI in include/cc_b200.h (libccb200.so).

There is NO fallback: if the CUDA library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

from .arch import CcModule

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libccb200.so")

_vp, _i32, _i64, _sz = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t
_M = C.POINTER(CcModule)

_SIGS = {
    "cc_abi_version": (_i32, []),
    "cc_last_error": (C.c_char_p, []),
    "cc_launch_count": (_i64, []),
    "cc_param_count": (_i64, [_M]),
    "cc_module_supported": (_i32, [_M, _i32]),
    "cc_kernel_family": (_i32, [_M, _M, _i32]),
    "cc_kernel_family_at": (_i32, [_M, _M, _i32, _i64]),
    "cc_rollout_tape_bytes": (_sz, [_M, _i64, _i32, _i32]),
    "cc_rollout_workspace_bytes": (_sz, [_M, _i64, _i32, _i32]),
    "cc_rollout_fwd_workspace_bytes": (_sz, [_M, _i64, _i32]),
    "cc_rollout_fwd": (_i32, [_M, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i64, _i32, _vp, _sz, _vp]),
    "cc_rollout_bwd": (_i32, [_M, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _i64, _i32, _vp, _sz, _vp]),
    "cc_closedloop_tape_bytes": (_sz, [_M, _M, _i64, _i32, _i32]),
    "cc_closedloop_workspace_bytes": (_sz, [_M, _M, _i64, _i32, _i32]),
    "cc_closedloop_fwd_workspace_bytes": (_sz, [_M, _M, _i64, _i32]),
    "cc_closedloop_fwd": (_i32, [_M, _M, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i64, _i32, _vp, _sz, _vp]),
    "cc_closedloop_bwd": (_i32, [_M, _M, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _i64, _i32, _vp, _sz, _vp]),
    "cc_mse_value_and_cotangent": (_i32, [_vp, _vp, _vp, _vp, _i64, C.c_float, _vp]),
    "cc_opt_step": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64] + [C.c_float] * 9 + [_vp]),
    "cc_env_two_segments_rollout": (_i32, [_vp, _i64, _vp, _vp, _vp, _vp, _i64, _i32, _i32, C.c_double, _vp]),
    "cc_fp32_peak_probe": (_i32, [_vp, _i32, _vp]),
    "cc_fp32_tile_probe": (_i32, [_vp, _vp, _i32, _vp]),
}


class CcError(RuntimeError):
    pass


class CcLib:
    """Thin, checked view of the shared library.  Pointers are plain integers (device addresses)."""

    def __init__(self, path: str = LIB_PATH):
        if not os.path.exists(path):
            raise CcError(
                f"{path} not found: the CUDA extension is not built (run `python -c 'import __graft_entry__ as g; "
                "g.build()'`).  chain_control_b200 has no CPU fallback."
            )
        self.path = path
        self._dll = C.CDLL(path)
        for name, (res, args) in _SIGS.items():
            if not hasattr(self._dll, name):
                continue  # the export check lives in tests/test_abi.py
            fn = getattr(self._dll, name)
            fn.restype, fn.argtypes = res, args

    def last_error(self) -> str:
        return self._dll.cc_last_error().decode()

    def check(self, rc: int, what: str):
        if rc != 0:
            raise CcError(f"{what} failed ({rc}): {self.last_error()}")

    def __getattr__(self, name):
        return getattr(self._dll, name)


_lib = None


def lib() -> CcLib:
    global _lib
    if _lib is None:
        _lib = CcLib()
    return _lib
