"""Array-level Python API of the rig path (ccg_rig_*), see include/ccg.h."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ccg
from ._ccg import Report, Timing, default_opts, f32, f64, i64, ptr
from .api import Context


def _prep(u, v, obs_cam, obs_point, frame_offsets, Xw, q_c, t_c, q_f, t_f):
    u, v = f32(u), f32(v)
    cam = np.ascontiguousarray(obs_cam, dtype=np.int32)
    pt = i64(obs_point)
    foff = i64(frame_offsets)
    Xw = f32(Xw).reshape(-1, 3)
    nf = len(foff) - 1
    q_c = f64(q_c).reshape(-1, 4).copy()
    nc = q_c.shape[0]
    t_c = f64(t_c).reshape(nc, 3).copy()
    q_f = f64(q_f).reshape(nf, 4).copy()
    t_f = f64(t_f).reshape(nf, 3).copy()
    return u, v, cam, pt, foff, Xw, nf, nc, q_c, t_c, q_f, t_f


def rig_solve(ctx: Context, u, v, obs_cam, obs_point, frame_offsets, Xw, q_c, t_c, cam_frozen, q_f, t_f,
              opts=None):
    """ExtrinsicsCalibrator::Optimize on host arrays (H2D + LM solve + D2H).
    Returns (q_c, t_c, q_f, t_f, obs_cost, report)."""
    u, v, cam, pt, foff, Xw, nf, nc, q_c, t_c, q_f, t_f = _prep(u, v, obs_cam, obs_point, frame_offsets, Xw,
                                                                q_c, t_c, q_f, t_f)
    frozen = np.ascontiguousarray(cam_frozen, dtype=np.uint8)
    n = int(foff[-1] - foff[0])
    cost = np.zeros(n)
    o = opts if opts is not None else default_opts(max_iterations=1000)
    rep = Report()
    rc = _ccg.lib().ccg_rig_solve(
        ctx.handle, ptr(u, C.c_float), ptr(v, C.c_float), ptr(cam, C.c_int32), ptr(pt, C.c_int64),
        ptr(foff, C.c_int64), C.c_int64(nf), ptr(Xw, C.c_float), C.c_int64(Xw.shape[0]), C.c_int32(nc),
        ptr(q_c, C.c_double), ptr(t_c, C.c_double), ptr(frozen, C.c_uint8), ptr(q_f, C.c_double),
        ptr(t_f, C.c_double), ptr(cost, C.c_double), C.byref(o), C.byref(rep))
    ctx._check(rc, "ccg_rig_solve")
    return q_c, t_c, q_f, t_f, cost, rep.asdict()


def rig_eval(ctx: Context, u, v, obs_cam, obs_point, frame_offsets, Xw, q_c, t_c, q_f, t_f, huber_a):
    u, v, cam, pt, foff, Xw, nf, nc, q_c, t_c, q_f, t_f = _prep(u, v, obs_cam, obs_point, frame_offsets, Xw,
                                                                q_c, t_c, q_f, t_f)
    n = int(foff[-1] - foff[0])
    n6 = 6 * nc
    res = np.zeros(2 * n)
    Vb = np.zeros((nf, 6, 6))
    Wb = np.zeros((nf, 6, n6))
    gv = np.zeros((nf, 6))
    U = np.zeros((n6, n6))
    gu = np.zeros(n6)
    cost = np.zeros(1)
    rc = _ccg.lib().ccg_rig_eval(
        ctx.handle, ptr(u, C.c_float), ptr(v, C.c_float), ptr(cam, C.c_int32), ptr(pt, C.c_int64),
        ptr(foff, C.c_int64), C.c_int64(nf), ptr(Xw, C.c_float), C.c_int64(Xw.shape[0]), C.c_int32(nc),
        ptr(q_c, C.c_double), ptr(t_c, C.c_double), ptr(q_f, C.c_double), ptr(t_f, C.c_double),
        C.c_double(huber_a), ptr(res, C.c_double), ptr(Vb, C.c_double), ptr(Wb, C.c_double),
        ptr(gv, C.c_double), ptr(U, C.c_double), ptr(gu, C.c_double), ptr(cost, C.c_double))
    ctx._check(rc, "ccg_rig_eval")
    return dict(res=res.reshape(n, 2), V=Vb, W=Wb, gv=gv, U=U, gu=gu, cost=float(cost[0]))


class RigProblem:
    """Device-resident rig problem: upload once, run many times."""

    def __init__(self, ctx: Context, u, v, obs_cam, obs_point, frame_offsets, Xw, q_c, t_c, cam_frozen, q_f,
                 t_f, shared=False):
        self.ctx = ctx
        u, v, cam, pt, foff, Xw, nf, nc, q_c, t_c, q_f, t_f = _prep(u, v, obs_cam, obs_point, frame_offsets,
                                                                    Xw, q_c, t_c, q_f, t_f)
        frozen = np.ascontiguousarray(cam_frozen, dtype=np.uint8)
        self.n_frames, self.n_cams = nf, nc
        self.n_obs = int(foff[-1] - foff[0])
        self._h = C.c_void_p()
        rc = _ccg.lib().ccg_rig_create(
            ctx.handle, ptr(u, C.c_float), ptr(v, C.c_float), ptr(cam, C.c_int32), ptr(pt, C.c_int64),
            ptr(foff, C.c_int64), C.c_int64(nf), ptr(Xw, C.c_float), C.c_int64(Xw.shape[0]), C.c_int32(nc),
            ptr(q_c, C.c_double), ptr(t_c, C.c_double), ptr(frozen, C.c_uint8), ptr(q_f, C.c_double),
            ptr(t_f, C.c_double), C.c_int(1 if shared else 0), C.byref(self._h))
        ctx._check(rc, "ccg_rig_create")

    def run(self, opts=None) -> dict:
        o = opts if opts is not None else default_opts(max_iterations=1000)
        tm = Timing()
        self.ctx._check(_ccg.lib().ccg_rig_run(self._h, C.byref(o), C.byref(tm)), "ccg_rig_run")
        return tm.asdict()

    def download(self, with_cost=True):
        q_c = np.zeros((self.n_cams, 4))
        t_c = np.zeros((self.n_cams, 3))
        q_f = np.zeros((self.n_frames, 4))
        t_f = np.zeros((self.n_frames, 3))
        cost = np.zeros(self.n_obs) if with_cost else None
        rep = Report()
        self.ctx._check(_ccg.lib().ccg_rig_download(self._h, ptr(q_c, C.c_double), ptr(t_c, C.c_double),
                                                    ptr(q_f, C.c_double), ptr(t_f, C.c_double),
                                                    ptr(cost, C.c_double), C.byref(rep)), "ccg_rig_download")
        return q_c, t_c, q_f, t_f, cost, rep.asdict()

    def close(self):
        if self._h:
            _ccg.lib().ccg_rig_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
