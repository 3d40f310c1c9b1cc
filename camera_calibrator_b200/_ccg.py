"""ctypes binding of the ccg C-ABI (``include/ccg.h``) — the thin Python view of
``camera_calibrator_b200/libccg.so``.  There is no CPU fallback: if the library
is missing or no CUDA device is usable, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
# CCG_LIB selects an experimental build of the library (tools/build_variant.sh); default: the in-tree libccg.so
LIB_PATH = os.environ.get("CCG_LIB") or os.path.join(_PKG, "libccg.so")
CSRC = os.path.join(_PKG, "csrc")


def build(force: bool = False) -> str:
    """Compile libccg.so for sm_100a with nvcc (in-tree, see csrc/Makefile)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    srcs.append(os.path.join(_PKG, "..", "include", "ccg.h"))
    stale = force or not os.path.exists(LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if stale:
        subprocess.run(["make", "-C", CSRC, "-s"], check=True)
    return LIB_PATH


class Opts(C.Structure):
    _fields_ = [
        ("max_iterations", C.c_int32),
        ("profile", C.c_int32),
        ("initial_radius", C.c_double),
        ("max_radius", C.c_double),
        ("min_radius", C.c_double),
        ("min_relative_decrease", C.c_double),
        ("min_diagonal", C.c_double),
        ("max_diagonal", C.c_double),
        ("function_tolerance", C.c_double),
        ("gradient_tolerance", C.c_double),
        ("parameter_tolerance", C.c_double),
        ("huber_a", C.c_double),
        ("noise_factor", C.c_double),
        ("hessian_fp64", C.c_int32),
        ("reserved_", C.c_int32),
    ]


class Report(C.Structure):
    _fields_ = [
        ("iterations", C.c_int32),
        ("accepted", C.c_int32),
        ("rejected", C.c_int32),
        ("sweeps", C.c_int32),
        ("termination", C.c_int32),
        ("pad_", C.c_int32),
        ("initial_cost", C.c_double),
        ("final_cost", C.c_double),
        ("gradient_norm", C.c_double),
        ("rms", C.c_double),
    ]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "pad_"}


class Timing(C.Structure):
    _fields_ = [
        ("total_ms", C.c_double),
        ("sweep_ms", C.c_double),
        ("schur_ms", C.c_double),
        ("small_ms", C.c_double),
        ("sweep_launches", C.c_int32),
        ("schur_launches", C.c_int32),
        ("small_launches", C.c_int32),
        ("allreduces", C.c_int32),
        ("sweep_obs", C.c_int64),
        ("sweep_views", C.c_int64),
    ]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class CcgError(RuntimeError):
    def __init__(self, status, where, detail=""):
        self.status = status
        super().__init__(f"{where}: ccg status {status} ({status_string(status)}) {detail}".strip())


_lib = None


def lib():
    """Load libccg.so (must have been built; raises if absent)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OSError(
                f"{LIB_PATH} is missing: build it with camera_calibrator_b200.build() "
                "(nvcc, sm_100a). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        L.ccg_status_string.restype = C.c_char_p
        L.ccg_last_error.restype = C.c_char_p
        L.ccg_last_error.argtypes = [C.c_void_p]
        L.ccg_kernel_launches.restype = C.c_int64
        L.ccg_kernel_launches.argtypes = [C.c_void_p]
        L.ccg_intrinsics_device_bytes.restype = C.c_int64
        L.ccg_intrinsics_device_bytes.argtypes = [C.c_void_p]
        L.ccg_destroy.argtypes = [C.c_void_p]
        L.ccg_trim.argtypes = [C.c_void_p]
        L.ccg_intrinsics_destroy.argtypes = [C.c_void_p]
        if hasattr(L, "ccg_rig_destroy"):
            L.ccg_rig_destroy.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def status_string(status: int) -> str:
    return lib().ccg_status_string(C.c_int(status)).decode()


def default_opts(**kw) -> Opts:
    o = Opts()
    lib().ccg_default_opts(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def ptr(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)
