"""Array-level Python API over the ccg C-ABI: what bench.py, the tests and the
``pycalibrator`` host classes call.  All compute happens in libccg.so on the GPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ccg
from ._ccg import CcgError, Opts, Report, Timing, default_opts, f32, f64, i64, ptr


class Context:
    """One ccg context = one GPU = one host thread (ccg_init / ccg_destroy)."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        rc = _ccg.lib().ccg_init(C.c_int(device), C.byref(self._h))
        if rc != 0:
            raise CcgError(rc, "ccg_init", f"(device {device}; a CUDA device is required, there is no CPU path)")
        self.device = device

    # -- plumbing -----------------------------------------------------------
    @property
    def handle(self):
        return self._h

    def close(self):
        if self._h:
            _ccg.lib().ccg_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, where):
        if rc != 0:
            raise CcgError(rc, where, _ccg.lib().ccg_last_error(self._h).decode())

    def device_info(self):
        sm, ma, mi, mem = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int64()
        self._check(_ccg.lib().ccg_device_info(self._h, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(mem)),
                    "ccg_device_info")
        return dict(sm_count=sm.value, cc=(ma.value, mi.value), global_mem=mem.value)

    def trim(self):
        """Return the device memory cached from destroyed problems to the driver (ccg_trim)."""
        self._check(_ccg.lib().ccg_trim(self._h), "ccg_trim")

    def kernel_launches(self) -> int:
        return int(_ccg.lib().ccg_kernel_launches(self._h))

    def measure_fp64_peak(self) -> float:
        tf = C.c_double()
        self._check(_ccg.lib().ccg_measure_fp64_peak(self._h, C.byref(tf)), "ccg_measure_fp64_peak")
        return tf.value

    def measure_copy_bandwidth(self, nbytes: int = 1 << 30) -> float:
        g = C.c_double()
        self._check(_ccg.lib().ccg_measure_copy_bandwidth(self._h, C.c_int64(nbytes), C.byref(g)),
                    "ccg_measure_copy_bandwidth")
        return g.value

    # -- multi-GPU ------------------------------------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        rc = _ccg.lib().ccg_comm_unique_id(buf)
        if rc != 0:
            raise CcgError(rc, "ccg_comm_unique_id")
        return bytes(buf)

    def comm_init(self, rank: int, world: int, uid: bytes):
        buf = (C.c_uint8 * 128).from_buffer_copy(uid)
        self._check(_ccg.lib().ccg_comm_init(self._h, C.c_int(rank), C.c_int(world), buf), "ccg_comm_init")

    def barrier(self):
        self._check(_ccg.lib().ccg_comm_barrier(self._h), "ccg_comm_barrier")

    # -- single camera --------------------------------------------------------
    def intrinsics_solve(self, u, v, X, Y, Z, view_offsets, intr, q, t, const_mask=0, opts=None,
                         problem_offsets=None):
        """Host arrays in, host arrays out (H2D + LM solve + D2H).

        Returns (intr, q[V,4], t[V,3], report) — `intr` is [9] and `report` a dict
        for one problem, [B,9] and a list when `problem_offsets` is given."""
        u, v, X, Y, Z = map(f32, (u, v, X, Y, Z))
        off = i64(view_offsets)
        nv = len(off) - 1
        q = f64(q).reshape(nv, 4).copy()
        t = f64(t).reshape(nv, 3).copy()
        o = opts if opts is not None else default_opts()
        L = _ccg.lib()
        if problem_offsets is None:
            intr = f64(intr).reshape(9).copy()
            rep = Report()
            rc = L.ccg_intrinsics_solve(
                self._h, ptr(u, C.c_float), ptr(v, C.c_float), ptr(X, C.c_float), ptr(Y, C.c_float),
                ptr(Z, C.c_float), ptr(off, C.c_int64), C.c_int64(nv), ptr(intr, C.c_double),
                ptr(q, C.c_double), ptr(t, C.c_double), C.c_uint32(const_mask), C.byref(o), C.byref(rep))
            self._check(rc, "ccg_intrinsics_solve")
            return intr, q, t, rep.asdict()
        poff = i64(problem_offsets)
        nb = len(poff) - 1
        intr = f64(intr).reshape(nb, 9).copy()
        reps = (Report * nb)()
        rc = L.ccg_intrinsics_solve_batch(
            self._h, ptr(u, C.c_float), ptr(v, C.c_float), ptr(X, C.c_float), ptr(Y, C.c_float),
            ptr(Z, C.c_float), ptr(off, C.c_int64), C.c_int64(nv), ptr(poff, C.c_int64), C.c_int64(nb),
            ptr(intr, C.c_double), ptr(q, C.c_double), ptr(t, C.c_double), C.c_uint32(const_mask),
            C.byref(o), reps)
        self._check(rc, "ccg_intrinsics_solve_batch")
        return intr, q, t, [r.asdict() for r in reps]

    def undistort(self, px, py, K, dist):
        """Calibrator::Undistort on arrays: pixels -> undistorted normalised coordinates (float32)."""
        px, py = f32(px), f32(py)
        K32 = np.ascontiguousarray(K, dtype=np.float32).reshape(3, 3)
        Kinv = np.ascontiguousarray(np.linalg.inv(K32.astype(np.float64)).astype(np.float32))
        d5 = np.ascontiguousarray(dist, dtype=np.float32).reshape(5)
        xn, yn = np.zeros_like(px), np.zeros_like(py)
        self._check(_ccg.lib().ccg_undistort(self._h, ptr(px, C.c_float), ptr(py, C.c_float), C.c_int64(px.size),
                                             ptr(Kinv, C.c_float), ptr(d5, C.c_float), ptr(xn, C.c_float),
                                             ptr(yn, C.c_float)), "ccg_undistort")
        return xn, yn

    def distort(self, xn, yn, K, dist):
        """Calibrator::Distort on arrays: normalised coordinates -> distorted pixels (float32)."""
        xn, yn = f32(xn), f32(yn)
        K32 = np.ascontiguousarray(K, dtype=np.float32).reshape(3, 3)
        d5 = np.ascontiguousarray(dist, dtype=np.float32).reshape(5)
        px, py = np.zeros_like(xn), np.zeros_like(yn)
        self._check(_ccg.lib().ccg_distort(self._h, ptr(xn, C.c_float), ptr(yn, C.c_float), C.c_int64(xn.size),
                                           ptr(K32, C.c_float), ptr(d5, C.c_float), ptr(px, C.c_float),
                                           ptr(py, C.c_float)), "ccg_distort")
        return px, py

    def homography_batch(self, x1, y1, x2, y2, set_offsets):
        """EstimateHomography (reference geometry.cpp:70-107) for many point sets on the device: (n_sets, 3, 3),
        unit Frobenius norm; NaN in H[i, 0, 0] marks a degenerate set."""
        x1, y1, x2, y2 = f32(x1), f32(y1), f32(x2), f32(y2)
        off = np.ascontiguousarray(set_offsets, dtype=np.int64)
        n = off.size - 1
        H = np.zeros((n, 3, 3))
        self._check(_ccg.lib().ccg_homography_batch(self._h, ptr(x1, C.c_float), ptr(y1, C.c_float), ptr(x2, C.c_float),
                                                    ptr(y2, C.c_float), ptr(off, C.c_int64), C.c_int64(n),
                                                    ptr(H, C.c_double)), "ccg_homography_batch")
        return H

    def intrinsics_estimate(self, u, v, X, Y, Z, view_offsets, const_mask=0, opts=None, problem_offsets=None):
        """Calibrator::Estimate: closed-form initialisation on the device + LM solve.
        Returns (intr[9] or [B,9], q[V,4], t[V,3], report(s), intr_init)."""
        u, v, X, Y, Z = map(f32, (u, v, X, Y, Z))
        off = i64(view_offsets)
        nv = len(off) - 1
        o = opts if opts is not None else default_opts()
        q = np.zeros((nv, 4))
        t = np.zeros((nv, 3))
        if problem_offsets is None:
            intr = np.zeros(9)
            init = np.zeros(9)
            rep = Report()
            rc = _ccg.lib().ccg_intrinsics_estimate(
                self._h, ptr(u, C.c_float), ptr(v, C.c_float), ptr(X, C.c_float), ptr(Y, C.c_float),
                ptr(Z, C.c_float), ptr(off, C.c_int64), C.c_int64(nv), ptr(intr, C.c_double),
                ptr(q, C.c_double), ptr(t, C.c_double), ptr(init, C.c_double), C.c_uint32(const_mask),
                C.byref(o), C.byref(rep))
            self._check(rc, "ccg_intrinsics_estimate")
            return intr, q, t, rep.asdict(), init
        poff = i64(problem_offsets)
        nb = len(poff) - 1
        intr = np.zeros((nb, 9))
        init = np.zeros((nb, 9))
        reps = (Report * nb)()
        rc = _ccg.lib().ccg_intrinsics_estimate_batch(
            self._h, ptr(u, C.c_float), ptr(v, C.c_float), ptr(X, C.c_float), ptr(Y, C.c_float),
            ptr(Z, C.c_float), ptr(off, C.c_int64), C.c_int64(nv), ptr(poff, C.c_int64), C.c_int64(nb),
            ptr(intr, C.c_double), ptr(q, C.c_double), ptr(t, C.c_double), ptr(init, C.c_double),
            C.c_uint32(const_mask), C.byref(o), reps)
        self._check(rc, "ccg_intrinsics_estimate_batch")
        return intr, q, t, [r.asdict() for r in reps], init

    def intrinsics_eval(self, u, v, X, Y, Z, view_offsets, intr, q, t, hessian_fp64=0):
        """One sweep at the given point.  hessian_fp64: 0 tensor-core TF32 J'J (default sweep), 1 all FP64
        (ccg_opts.hessian_fp64)."""
        u, v, X, Y, Z = map(f32, (u, v, X, Y, Z))
        off = i64(view_offsets)
        nv = len(off) - 1
        n = int(off[-1] - off[0])
        intr = f64(intr).reshape(9)
        q = f64(q).reshape(nv, 4)
        t = f64(t).reshape(nv, 3)
        res = np.zeros(2 * n)
        Vb = np.zeros((nv, 6, 6))
        Wb = np.zeros((nv, 6, 9))
        gv = np.zeros((nv, 6))
        U = np.zeros((9, 9))
        gu = np.zeros(9)
        cost = np.zeros(1)
        rc = _ccg.lib().ccg_intrinsics_eval(
            self._h, ptr(u, C.c_float), ptr(v, C.c_float), ptr(X, C.c_float), ptr(Y, C.c_float),
            ptr(Z, C.c_float), ptr(off, C.c_int64), C.c_int64(nv), ptr(intr, C.c_double),
            ptr(q, C.c_double), ptr(t, C.c_double), ptr(res, C.c_double), ptr(Vb, C.c_double),
            ptr(Wb, C.c_double), ptr(gv, C.c_double), ptr(U, C.c_double), ptr(gu, C.c_double),
            ptr(cost, C.c_double), C.c_int(int(hessian_fp64)))
        self._check(rc, "ccg_intrinsics_eval")
        return dict(res=res.reshape(n, 2), V=Vb, W=Wb, gv=gv, U=U, gu=gu, cost=float(cost[0]))

    def intrinsics_problem(self, u, v, X, Y, Z, view_offsets, intr, q, t, const_mask=0,
                           problem_offsets=None, shared=False):
        return IntrinsicsProblem(self, u, v, X, Y, Z, view_offsets, intr, q, t, const_mask,
                                 problem_offsets, shared)


class IntrinsicsProblem:
    """Device-resident single-camera problem (batch): upload once, run many times."""

    def __init__(self, ctx: Context, u, v, X, Y, Z, view_offsets, intr, q, t, const_mask=0,
                 problem_offsets=None, shared=False):
        self.ctx = ctx
        u, v, X, Y, Z = map(f32, (u, v, X, Y, Z))
        off = i64(view_offsets)
        self.n_views = len(off) - 1
        self.n_obs = int(off[-1] - off[0])
        if problem_offsets is None:
            problem_offsets = np.array([0, self.n_views], dtype=np.int64)
        poff = i64(problem_offsets)
        self.n_problems = len(poff) - 1
        intr = f64(intr).reshape(self.n_problems, 9)
        q = f64(q).reshape(self.n_views, 4)
        t = f64(t).reshape(self.n_views, 3)
        self._h = C.c_void_p()
        rc = _ccg.lib().ccg_intrinsics_create(
            ctx.handle, ptr(u, C.c_float), ptr(v, C.c_float), ptr(X, C.c_float), ptr(Y, C.c_float),
            ptr(Z, C.c_float), ptr(off, C.c_int64), C.c_int64(self.n_views), ptr(poff, C.c_int64),
            C.c_int64(self.n_problems), ptr(intr, C.c_double), ptr(q, C.c_double), ptr(t, C.c_double),
            C.c_uint32(const_mask), C.c_int(1 if shared else 0), C.byref(self._h))
        ctx._check(rc, "ccg_intrinsics_create")

    def run(self, opts=None) -> dict:
        o = opts if opts is not None else default_opts()
        tm = Timing()
        self.ctx._check(_ccg.lib().ccg_intrinsics_run(self._h, C.byref(o), C.byref(tm)), "ccg_intrinsics_run")
        return tm.asdict()

    def time_sweep(self, n: int = 10) -> float:
        ms = C.c_float()
        self.ctx._check(_ccg.lib().ccg_intrinsics_time_sweep(self._h, C.c_int(n), C.byref(ms)),
                        "ccg_intrinsics_time_sweep")
        return ms.value

    def download(self):
        intr = np.zeros((self.n_problems, 9))
        q = np.zeros((self.n_views, 4))
        t = np.zeros((self.n_views, 3))
        reps = (Report * self.n_problems)()
        self.ctx._check(_ccg.lib().ccg_intrinsics_download(self._h, ptr(intr, C.c_double), ptr(q, C.c_double),
                                                           ptr(t, C.c_double), reps), "ccg_intrinsics_download")
        return intr, q, t, [r.asdict() for r in reps]

    def history(self, b: int = 0):
        """[n,6] per-sweep trace: cost, candidate cost, gain ratio, radius, rel. step norm, accepted."""
        rows = np.zeros((48, 6))
        n = C.c_int()
        self.ctx._check(_ccg.lib().ccg_intrinsics_history(self._h, C.c_int64(b), ptr(rows, C.c_double),
                                                          C.c_int(48), C.byref(n)), "ccg_intrinsics_history")
        return rows[:n.value].copy()

    def device_bytes(self) -> int:
        return int(_ccg.lib().ccg_intrinsics_device_bytes(self._h))

    def close(self):
        if self._h:
            _ccg.lib().ccg_intrinsics_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
