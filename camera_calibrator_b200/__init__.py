"""camera_calibrator_b200 — B200-native (sm_100a) implementation of the
Levenberg-Marquardt reprojection-error hot path of buq2/camera_calibrator,
behind the reference's Calibrator / ExtrinsicsCalibrator API.

Layout:
  csrc/      hand-written CUDA kernels + the ccg C-ABI (include/ccg.h) -> libccg.so
  host/      C++ host classes mirroring reference calibrator.hh /
             extrinsics_calibrator.hh and the pybind module `pycalibrator`
  api.py     array-level Python API over the C-ABI (ctypes)
  synth.py   synthetic planar-target inputs for bench.py / smoke()

There is no CPU fallback: every compute entry point needs libccg.so and a CUDA
device and raises otherwise.
"""
from ._ccg import CcgError, Opts, Report, Timing, build, default_opts, lib  # noqa: F401
from .api import Context, IntrinsicsProblem  # noqa: F401
