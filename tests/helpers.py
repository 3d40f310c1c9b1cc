"""Shared helpers of the parity tests (tests may use the oracle; the product may not)."""
import numpy as np


def quat_to_rot(q):
    q = np.asarray(q, dtype=np.float64)
    q = q / np.linalg.norm(q, axis=-1, keepdims=True)
    w, a, b, c = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0] = 1 - 2 * (b * b + c * c)
    R[..., 0, 1] = 2 * (a * b - w * c)
    R[..., 0, 2] = 2 * (a * c + w * b)
    R[..., 1, 0] = 2 * (a * b + w * c)
    R[..., 1, 1] = 1 - 2 * (a * a + c * c)
    R[..., 1, 2] = 2 * (b * c - w * a)
    R[..., 2, 0] = 2 * (a * c - w * b)
    R[..., 2, 1] = 2 * (b * c + w * a)
    R[..., 2, 2] = 1 - 2 * (a * a + b * b)
    return R


def rel_err(a, b, floor=0.0):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)) if a.size else 0.0


def block_err(a, b):
    """max |a-b| / max|b| — for blocks whose entries span many magnitudes."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    s = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / s) if s > 0 else float(np.max(np.abs(a - b)))


# BASELINE.json tolerances
TOL_PARAM = 1e-6   # relative, recovered intrinsics / extrinsics
TOL_RMS = 1e-9     # final RMS reprojection error


def start_from_estimate(oracle, d):
    """Initial point exactly as Calibrator::Estimate builds it (reference calibrator.cpp:47-67)."""
    _, K, q, t = oracle.estimate_init(d.u, d.v, d.X, d.Y, d.Z, d.view_offsets)
    intr = np.zeros(9)
    intr[0], intr[1], intr[2], intr[3] = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    return intr, q.astype(np.float64), t.astype(np.float64)
