"""Host C++ layer: Calibrator / ExtrinsicsCalibrator (reference calibrator.hh /
extrinsics_calibrator.hh mirrors) and the pybind11 module ``pycalibrator``."""
import importlib
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.dirname(_HERE)


def build() -> None:
    subprocess.run(["make", "-C", _HERE, "-s", f"PY={sys.executable}"], check=True)


def load():
    """Import and return the ``pycalibrator`` module built in-tree."""
    if _PKG not in sys.path:
        sys.path.insert(0, _PKG)
    return importlib.import_module("pycalibrator")
