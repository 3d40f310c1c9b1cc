"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes front-end of ``oracle/liboracle.so``, the CPU restatement of the
reference hot path (see ``oracle/oracle.h`` for what is restated and how the
oracle is pinned).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this
package; nothing under ``camera_calibrator_b200/`` does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (g++, no dependencies)."""
    if force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
        for f in os.listdir(_HERE)
        if f.endswith((".cpp", ".hh", ".h"))
    ):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _LIB_PATH


class Opts(C.Structure):
    _fields_ = [
        ("max_iterations", C.c_int32),
        ("verbose", C.c_int32),
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


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.oracle_num_threads.restype = C.c_int
    return _lib


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def default_opts(**kw) -> Opts:
    o = Opts()
    lib().oracle_default_opts(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def set_num_threads(n: int) -> None:
    lib().oracle_set_num_threads(C.c_int(n))


# --------------------------------------------------------------------------
# single camera
# --------------------------------------------------------------------------
def intrinsics_solve(u, v, X, Y, Z, view_offsets, intr, q, t, const_mask=0, opts=None):
    """Returns (intr[9], q[V,4], t[V,3], report dict).  Inputs are not modified."""
    u, v, X, Y, Z = map(_f32, (u, v, X, Y, Z))
    off = _i64(view_offsets)
    nv = len(off) - 1
    intr = _f64(intr).copy()
    q = _f64(q).reshape(nv, 4).copy()
    t = _f64(t).reshape(nv, 3).copy()
    o = opts if opts is not None else default_opts()
    rep = Report()
    rc = lib().oracle_intrinsics_solve(
        _p(u, C.c_float), _p(v, C.c_float), _p(X, C.c_float), _p(Y, C.c_float), _p(Z, C.c_float),
        _p(off, C.c_int64), C.c_int64(nv), _p(intr, C.c_double), _p(q, C.c_double),
        _p(t, C.c_double), C.c_uint32(const_mask), C.byref(o), C.byref(rep))
    if rc != 0:
        raise RuntimeError(f"oracle_intrinsics_solve failed: {rc}")
    return intr, q, t, rep.asdict()


def intrinsics_solve_batch(u, v, X, Y, Z, view_offsets, problem_offsets, intr, q, t, const_mask=0,
                           opts=None):
    u, v, X, Y, Z = map(_f32, (u, v, X, Y, Z))
    off = _i64(view_offsets)
    poff = _i64(problem_offsets)
    nv = len(off) - 1
    nb = len(poff) - 1
    intr = _f64(intr).reshape(nb, 9).copy()
    q = _f64(q).reshape(nv, 4).copy()
    t = _f64(t).reshape(nv, 3).copy()
    o = opts if opts is not None else default_opts()
    reps = (Report * nb)()
    rc = lib().oracle_intrinsics_solve_batch(
        _p(u, C.c_float), _p(v, C.c_float), _p(X, C.c_float), _p(Y, C.c_float), _p(Z, C.c_float),
        _p(off, C.c_int64), _p(poff, C.c_int64), C.c_int64(nb), _p(intr, C.c_double),
        _p(q, C.c_double), _p(t, C.c_double), C.c_uint32(const_mask), C.byref(o), reps)
    if rc != 0:
        raise RuntimeError(f"oracle_intrinsics_solve_batch failed: {rc}")
    return intr, q, t, [r.asdict() for r in reps]


def intrinsics_eval(u, v, X, Y, Z, view_offsets, intr, q, t):
    """One residual+Jacobian+normal-equation evaluation; returns a dict of blocks."""
    u, v, X, Y, Z = map(_f32, (u, v, X, Y, Z))
    off = _i64(view_offsets)
    nv = len(off) - 1
    n = int(off[-1] - off[0])
    intr = _f64(intr)
    q = _f64(q).reshape(nv, 4)
    t = _f64(t).reshape(nv, 3)
    res = np.zeros(2 * n)
    Vb = np.zeros((nv, 6, 6))
    Wb = np.zeros((nv, 6, 9))
    gv = np.zeros((nv, 6))
    U = np.zeros((9, 9))
    gu = np.zeros(9)
    cost = np.zeros(1)
    lib().oracle_intrinsics_eval(
        _p(u, C.c_float), _p(v, C.c_float), _p(X, C.c_float), _p(Y, C.c_float), _p(Z, C.c_float),
        _p(off, C.c_int64), C.c_int64(nv), _p(intr, C.c_double), _p(q, C.c_double),
        _p(t, C.c_double), _p(res, C.c_double), _p(Vb, C.c_double), _p(Wb, C.c_double),
        _p(gv, C.c_double), _p(U, C.c_double), _p(gu, C.c_double), _p(cost, C.c_double))
    return dict(res=res.reshape(n, 2), V=Vb, W=Wb, gv=gv, U=U, gu=gu, cost=float(cost[0]))


# --------------------------------------------------------------------------
# rig
# --------------------------------------------------------------------------
def rig_solve(u, v, obs_cam, obs_point, frame_offsets, Xw, q_c, t_c, cam_frozen, q_f, t_f,
              opts=None):
    u, v = _f32(u), _f32(v)
    cam = np.ascontiguousarray(obs_cam, dtype=np.int32)
    pt = _i64(obs_point)
    foff = _i64(frame_offsets)
    Xw = _f32(Xw).reshape(-1, 3)
    nf = len(foff) - 1
    q_c = _f64(q_c).reshape(-1, 4).copy()
    nc = q_c.shape[0]
    t_c = _f64(t_c).reshape(nc, 3).copy()
    q_f = _f64(q_f).reshape(nf, 4).copy()
    t_f = _f64(t_f).reshape(nf, 3).copy()
    frozen = np.ascontiguousarray(cam_frozen, dtype=np.uint8)
    n = int(foff[-1] - foff[0])
    cost = np.zeros(n)
    o = opts if opts is not None else default_opts(max_iterations=1000)
    rep = Report()
    rc = lib().oracle_rig_solve(
        _p(u, C.c_float), _p(v, C.c_float), _p(cam, C.c_int32), _p(pt, C.c_int64),
        _p(foff, C.c_int64), C.c_int64(nf), _p(Xw, C.c_float), C.c_int64(Xw.shape[0]),
        C.c_int32(nc), _p(q_c, C.c_double), _p(t_c, C.c_double), _p(frozen, C.c_uint8),
        _p(q_f, C.c_double), _p(t_f, C.c_double), _p(cost, C.c_double), C.byref(o), C.byref(rep))
    if rc != 0:
        raise RuntimeError(f"oracle_rig_solve failed: {rc}")
    return q_c, t_c, q_f, t_f, cost, rep.asdict()


def rig_eval(u, v, obs_cam, obs_point, frame_offsets, Xw, q_c, t_c, q_f, t_f, huber_a):
    u, v = _f32(u), _f32(v)
    cam = np.ascontiguousarray(obs_cam, dtype=np.int32)
    pt = _i64(obs_point)
    foff = _i64(frame_offsets)
    Xw = _f32(Xw).reshape(-1, 3)
    nf = len(foff) - 1
    q_c = _f64(q_c).reshape(-1, 4)
    nc = q_c.shape[0]
    t_c = _f64(t_c).reshape(nc, 3)
    q_f = _f64(q_f).reshape(nf, 4)
    t_f = _f64(t_f).reshape(nf, 3)
    n = int(foff[-1] - foff[0])
    n6 = 6 * nc
    res = np.zeros(2 * n)
    Vb = np.zeros((nf, 6, 6))
    Wb = np.zeros((nf, 6, n6))
    gv = np.zeros((nf, 6))
    U = np.zeros((n6, n6))
    gu = np.zeros(n6)
    cost = np.zeros(1)
    lib().oracle_rig_eval(
        _p(u, C.c_float), _p(v, C.c_float), _p(cam, C.c_int32), _p(pt, C.c_int64),
        _p(foff, C.c_int64), C.c_int64(nf), _p(Xw, C.c_float), C.c_int64(Xw.shape[0]),
        C.c_int32(nc), _p(q_c, C.c_double), _p(t_c, C.c_double), _p(q_f, C.c_double),
        _p(t_f, C.c_double), C.c_double(huber_a), _p(res, C.c_double), _p(Vb, C.c_double),
        _p(Wb, C.c_double), _p(gv, C.c_double), _p(U, C.c_double), _p(gu, C.c_double),
        _p(cost, C.c_double))
    return dict(res=res.reshape(n, 2), V=Vb, W=Wb, gv=gv, U=U, gu=gu, cost=float(cost[0]))


# --------------------------------------------------------------------------
# inputs / initialisation
# --------------------------------------------------------------------------
@dataclass
class Planar:
    u: np.ndarray
    v: np.ndarray
    X: np.ndarray
    Y: np.ndarray
    Z: np.ndarray
    view_offsets: np.ndarray
    planes: np.ndarray
    n_rejected: int


FIXTURE_K = np.array([[1000.0, 0, 800.0], [0, 1000.0, 500.0], [0, 0, 1]], dtype=np.float32)
FIXTURE_DIST = np.array([-4.0e-2, 5e-4, 1.0e-3, 2.0e-5, -3e-4], dtype=np.float32)  # test_calibrator.cpp:18


def datagen_planar(n_views, n_points, seed=0, width=1600, height=1000, K=FIXTURE_K, dist=FIXTURE_DIST,
                   noise_px=0.5) -> Planar:
    """DataGenerator::GetDistortedPointsPlanar(n_points) called n_views times (fixture of
    reference src/test_calibrator.cpp:9-28 by default)."""
    K = _f32(K).reshape(9)
    dist = _f32(dist)
    n = n_views * n_points
    u, v, X, Y, Z = (np.zeros(n, np.float32) for _ in range(5))
    planes = np.zeros((n_views, 4), np.float32)
    nrej = C.c_int64(0)
    lib().oracle_datagen_planar(
        C.c_uint32(seed), C.c_int32(width), C.c_int32(height), _p(K, C.c_float), _p(dist, C.c_float),
        C.c_float(noise_px), C.c_int32(n_points), C.c_int32(n_views), _p(u, C.c_float),
        _p(v, C.c_float), _p(X, C.c_float), _p(Y, C.c_float), _p(Z, C.c_float),
        _p(planes, C.c_float), C.byref(nrej))
    off = np.arange(n_views + 1, dtype=np.int64) * n_points
    return Planar(u, v, X, Y, Z, off, planes, int(nrej.value))


def estimate_init(u, v, X, Y, Z, view_offsets):
    """Initialisation of Calibrator::Estimate: returns (H[V,3,3], K[3,3], q[V,4], t[V,3]) float32."""
    u, v, X, Y, Z = map(_f32, (u, v, X, Y, Z))
    off = _i64(view_offsets)
    nv = len(off) - 1
    H = np.zeros((nv, 3, 3), np.float32)
    K = np.zeros((3, 3), np.float32)
    q = np.zeros((nv, 4), np.float32)
    t = np.zeros((nv, 3), np.float32)
    lib().oracle_estimate_init(
        _p(u, C.c_float), _p(v, C.c_float), _p(X, C.c_float), _p(Y, C.c_float), _p(Z, C.c_float),
        _p(off, C.c_int64), C.c_int64(nv), _p(H, C.c_float), _p(K, C.c_float), _p(q, C.c_float),
        _p(t, C.c_float))
    return H, K, q, t


def homography(x1, y1, x2, y2):
    x1, y1, x2, y2 = map(_f32, (x1, y1, x2, y2))
    H = np.zeros((3, 3), np.float32)
    lib().oracle_homography(_p(x1, C.c_float), _p(y1, C.c_float), _p(x2, C.c_float),
                            _p(y2, C.c_float), C.c_int64(len(x1)), _p(H, C.c_float))
    return H


def estimate(u, v, X, Y, Z, view_offsets, const_mask=0, dist0=None, opts=None):
    """Calibrator::Estimate = init + Optimize (reference src/calibrator.cpp:47-68)."""
    _, K, q, t = estimate_init(u, v, X, Y, Z, view_offsets)
    intr = np.zeros(9)
    intr[0], intr[1], intr[2], intr[3] = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    if dist0 is not None:
        intr[4:9] = np.asarray(dist0, dtype=np.float32)
    return intrinsics_solve(u, v, X, Y, Z, view_offsets, intr, q.astype(np.float64),
                            t.astype(np.float64), const_mask, opts)
