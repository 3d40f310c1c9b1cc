"""Synthetic calibration inputs for bench.py and smoke(): planar-target views with
the reference fixture's camera (reference src/test_calibrator.cpp:9-28) and the
same construction as DataGenerator::GetDistortedPointsPlanar (reference
src/data_generator.cpp:52-124: random plane through three frustum-corner rays at
U(0.2,1) distances, random pixels back-projected onto it, OpenCV 5-coefficient
lens model, U(+-noise) pixel noise, reject outside the image), vectorised over
views with numpy so that 4e7 observations are generated in seconds.

This generator draws from numpy's PCG64, so it is distribution-identical but NOT
stream-identical to the reference's std::mt19937 generator; the bit-faithful
restatement lives in oracle/datagen.cpp and is what the parity tests use.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

FIXTURE_K = np.array([[1000.0, 0, 800.0], [0, 1000.0, 500.0], [0, 0, 1]], dtype=np.float64)
FIXTURE_DIST = np.array([-4.0e-2, 5e-4, 1.0e-3, 2.0e-5, -3e-4], dtype=np.float64)  # k1 k2 p1 p2 k3
FIXTURE_SIZE = (1600, 1000)


@dataclass
class PlanarViews:
    u: np.ndarray
    v: np.ndarray
    X: np.ndarray
    Y: np.ndarray
    Z: np.ndarray
    view_offsets: np.ndarray
    q_true: np.ndarray  # [V,4] (w,x,y,z)
    t_true: np.ndarray  # [V,3]

    @property
    def n_views(self):
        return len(self.view_offsets) - 1

    @property
    def n_obs(self):
        return int(self.view_offsets[-1])


def _project(p, K, dist):
    x = p[..., 0] / p[..., 2]
    y = p[..., 1] / p[..., 2]
    k1, k2, p1, p2, k3 = dist
    r2 = x * x + y * y
    m = 1.0 + k1 * r2 + k2 * r2 * r2 + k3 * r2 * r2 * r2
    xd = x * m + 2.0 * p1 * x * y + p2 * (r2 + 2.0 * x * x)
    yd = y * m + 2.0 * p2 * x * y + p1 * (r2 + 2.0 * y * y)
    return K[0, 0] * xd + K[0, 2], K[1, 1] * yd + K[1, 2]


def rot_to_quat(R):
    """Batch rotation matrices [N,3,3] -> unit quaternions [N,4] (w,x,y,z), w >= 0."""
    R = np.asarray(R, dtype=np.float64)
    q = np.empty(R.shape[:-2] + (4,))
    tr = R[..., 0, 0] + R[..., 1, 1] + R[..., 2, 2]
    q[..., 0] = np.sqrt(np.maximum(0.0, 1.0 + tr)) / 2.0
    q[..., 1] = np.sqrt(np.maximum(0.0, 1.0 + R[..., 0, 0] - R[..., 1, 1] - R[..., 2, 2])) / 2.0
    q[..., 2] = np.sqrt(np.maximum(0.0, 1.0 - R[..., 0, 0] + R[..., 1, 1] - R[..., 2, 2])) / 2.0
    q[..., 3] = np.sqrt(np.maximum(0.0, 1.0 - R[..., 0, 0] - R[..., 1, 1] + R[..., 2, 2])) / 2.0
    q[..., 1] = np.copysign(q[..., 1], R[..., 2, 1] - R[..., 1, 2])
    q[..., 2] = np.copysign(q[..., 2], R[..., 0, 2] - R[..., 2, 0])
    q[..., 3] = np.copysign(q[..., 3], R[..., 1, 0] - R[..., 0, 1])
    return q / np.linalg.norm(q, axis=-1, keepdims=True)


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


CHUNK_VIEWS = 2048  # views are generated in chunks, chunk c from the stream seeded (seed, c)


def planar_views(n_views, n_points, seed=0, K=FIXTURE_K, dist=FIXTURE_DIST, size=FIXTURE_SIZE,
                 noise_px=0.5, view_range=None) -> PlanarViews:
    """Views [view_range[0], view_range[1]) of the `n_views`-view data set `seed` (default: all).
    Every chunk of CHUNK_VIEWS views has its own random stream, so a rank of a
    view-partitioned run generates exactly its share and all ranks agree on the whole."""
    a, b = (0, n_views) if view_range is None else view_range
    assert 0 <= a <= b <= n_views
    parts = []
    for c in range(a // CHUNK_VIEWS, (b - 1) // CHUNK_VIEWS + 1 if b > a else a // CHUNK_VIEWS):
        c0 = c * CHUNK_VIEWS
        nvc = min(CHUNK_VIEWS, n_views - c0)
        part = _planar_chunk(nvc, n_points, np.random.default_rng([seed, c]), K, dist, size, noise_px)
        lo, hi = max(a, c0) - c0, min(b, c0 + nvc) - c0
        parts.append(_slice_views(part, lo, hi, n_points))
    if not parts:
        z = np.zeros(0, np.float32)
        return PlanarViews(z, z, z, z, z, np.zeros(1, np.int64), np.zeros((0, 4)), np.zeros((0, 3)))
    cat = lambda k: np.concatenate([getattr(p, k) for p in parts])
    off = np.arange(b - a + 1, dtype=np.int64) * n_points
    return PlanarViews(cat("u"), cat("v"), cat("X"), cat("Y"), cat("Z"), off, cat("q_true"), cat("t_true"))


def _slice_views(p, lo, hi, n_points):
    s = slice(lo * n_points, hi * n_points)
    off = np.arange(hi - lo + 1, dtype=np.int64) * n_points
    return PlanarViews(p.u[s], p.v[s], p.X[s], p.Y[s], p.Z[s], off, p.q_true[lo:hi], p.t_true[lo:hi])


def _planar_chunk(n_views, n_points, rng, K, dist, size, noise_px, chunk_views=CHUNK_VIEWS) -> PlanarViews:
    w, h = size
    Kinv = np.linalg.inv(K)
    n = n_views * n_points
    u = np.empty(n, np.float32)
    v = np.empty(n, np.float32)
    X = np.empty(n, np.float32)
    Y = np.empty(n, np.float32)
    q_true = np.empty((n_views, 4))
    t_true = np.empty((n_views, 3))
    corners = np.array([[0.0, 0.0, 1.0], [w - 1.0, 0.0, 1.0], [w - 1.0, h - 1.0, 1.0]])
    rays_c = corners @ Kinv.T
    rays_c /= np.linalg.norm(rays_c, axis=1, keepdims=True)
    for v0 in range(0, n_views, chunk_views):
        nv = min(chunk_views, n_views - v0)
        # plane n.p = 1 through three points on the corner rays
        dists = rng.uniform(0.2, 1.0, size=(nv, 3))
        A = rays_c[None, :, :] * dists[:, :, None]  # [nv,3(points),3]
        pn = np.linalg.solve(A, np.ones((nv, 3, 1)))[..., 0]
        nrm = pn / np.linalg.norm(pn, axis=1, keepdims=True)
        v1 = np.cross(nrm, np.array([0.0, 0.0, 1.0]))
        v1 /= np.linalg.norm(v1, axis=1, keepdims=True)
        v2 = np.cross(nrm, v1)
        v2 /= np.linalg.norm(v2, axis=1, keepdims=True)
        R = np.stack([v1, v2, nrm], axis=1)  # rows v1, v2, normal  (camera -> board)
        uu = np.empty((nv, n_points))
        vv = np.empty((nv, n_points))
        pp = np.empty((nv, n_points, 3))
        need = np.ones((nv, n_points), dtype=bool)
        while need.any():
            idx = np.nonzero(need)
            m = idx[0].shape[0]
            px = np.stack([rng.uniform(0.0, w - 1.0, m), rng.uniform(0.0, h - 1.0, m), np.ones(m)], axis=1)
            ray = px @ Kinv.T
            s = 1.0 / np.einsum("ij,ij->i", ray, pn[idx[0]])
            p = ray * s[:, None]
            iu, iv = _project(p, K, dist)
            iu = (iu + rng.uniform(-noise_px, noise_px, m)).astype(np.float32)
            iv = (iv + rng.uniform(-noise_px, noise_px, m)).astype(np.float32)
            ok = (iu >= 0.0) & (iu < w - 1) & (iv >= 0.0) & (iv < h - 1) & (s > 0)
            sel = (idx[0][ok], idx[1][ok])
            uu[sel] = iu[ok]
            vv[sel] = iv[ok]
            pp[sel] = p[ok]
            need[sel] = False
        board = np.einsum("vij,vpj->vpi", R, pp)
        sl = slice(v0 * n_points, (v0 + nv) * n_points)
        u[sl] = uu.reshape(-1)
        v[sl] = vv.reshape(-1)
        X[sl] = board[..., 0].reshape(-1)
        Y[sl] = board[..., 1].reshape(-1)
        # camera-frame point = R^T (X, Y, d), d = n.p = 1/|pn|
        Rt = np.transpose(R, (0, 2, 1))
        q_true[v0:v0 + nv] = rot_to_quat(Rt)
        t_true[v0:v0 + nv] = nrm / np.linalg.norm(pn, axis=1, keepdims=True)
    off = np.arange(n_views + 1, dtype=np.int64) * n_points
    return PlanarViews(u, v, X, Y, np.zeros(n, np.float32), off, q_true, t_true)


def perturbed_start(views: PlanarViews, seed=1, K=FIXTURE_K, trans=0.02, rot_deg=1.0, focal_rel=0.02,
                    centre_px=10.0, n_views_total=None, view_start=0):
    """Solver-only starting point: ground-truth poses perturbed by
    U(+-trans) / U(+-rot_deg) per axis, focal lengths off by +-focal_rel, principal
    point off by +-centre_px, zero distortion (reference default, calibrator.hh:50).
    For a partition, pass the total view count and this partition's first view so
    that every rank draws the same perturbations for the same views."""
    rng = np.random.default_rng(seed)
    nv = views.n_views
    tot = nv if n_views_total is None else n_views_total
    pert = rng.uniform(-1.0, 1.0, size=(tot, 6))[view_start:view_start + nv]
    ang = np.deg2rad(rot_deg * pert[:, :3]) / 2.0
    dq = np.concatenate([np.ones((nv, 1)), ang], axis=1)
    dq /= np.linalg.norm(dq, axis=1, keepdims=True)
    q0 = views.q_true
    w0, x0, y0, z0 = dq.T
    w1, x1, y1, z1 = q0.T
    q = np.stack([w0 * w1 - x0 * x1 - y0 * y1 - z0 * z1, w0 * x1 + x0 * w1 + y0 * z1 - z0 * y1,
                  w0 * y1 - x0 * z1 + y0 * w1 + z0 * x1, w0 * z1 + x0 * y1 - y0 * x1 + z0 * w1], axis=1)
    t = views.t_true + trans * pert[:, 3:]
    intr = np.zeros(9)
    intr[0] = K[0, 0] * (1.0 + rng.uniform(-focal_rel, focal_rel))
    intr[1] = K[1, 1] * (1.0 + rng.uniform(-focal_rel, focal_rel))
    intr[2] = K[0, 2] + rng.uniform(-centre_px, centre_px)
    intr[3] = K[1, 2] + rng.uniform(-centre_px, centre_px)
    return intr, q, t


# --------------------------------------------------------------------------
# rig scenes (ExtrinsicsCalibrator): the scenario of reference
# src/test_extrinsics_calibrator.cpp:40-134, scaled per SURVEY.md App. B (C3)
# --------------------------------------------------------------------------
@dataclass
class RigScene:
    u: np.ndarray            # normalised image coordinates per observation (float32)
    v: np.ndarray
    obs_cam: np.ndarray      # int32
    obs_point: np.ndarray    # int64 global world_point_id
    frame_offsets: np.ndarray
    Xw: np.ndarray           # [n_points,3] float32 (perturbed world points handed to the solver)
    q_c: np.ndarray          # initial (perturbed) camera_T_rig, [C,4]
    t_c: np.ndarray
    cam_frozen: np.ndarray   # uint8
    q_f: np.ndarray          # initial (perturbed) rig_T_world, [F,4]
    t_f: np.ndarray
    q_c_true: np.ndarray
    t_c_true: np.ndarray
    q_f_true: np.ndarray
    t_f_true: np.ndarray

    @property
    def n_obs(self):
        return int(self.frame_offsets[-1])


def _quat_mul(a, b):
    w0, x0, y0, z0 = np.moveaxis(a, -1, 0)
    w1, x1, y1, z1 = np.moveaxis(b, -1, 0)
    return np.stack([w0 * w1 - x0 * x1 - y0 * y1 - z0 * z1, w0 * x1 + x0 * w1 + y0 * z1 - z0 * y1,
                     w0 * y1 - x0 * z1 + y0 * w1 + z0 * x1, w0 * z1 + x0 * y1 - y0 * x1 + z0 * w1], axis=-1)


def _perturb_pose(rng, q, t, trans, rot_deg):
    n = q.shape[0]
    ang = np.deg2rad(rng.uniform(-rot_deg, rot_deg, size=(n, 3))) / 2.0
    dq = np.concatenate([np.ones((n, 1)), ang], axis=1)
    dq /= np.linalg.norm(dq, axis=1, keepdims=True)
    return _quat_mul(q, dq), t + rng.uniform(-trans, trans, size=(n, 3))


def rig_scene(n_cams=4, n_frames=1000, grid=(17, 11), pitch=0.02, seed=0, noise_2d=2.0 / 500.0,
              noise_3d=0.001, cam_trans=0.005, cam_rot_deg=0.1, frame_trans=0.02, frame_rot_deg=1.0,
              outlier_fraction=0.0, order="frame-major") -> RigScene:
    """Camera 0 is the identity and frozen; the others sit at U(+-0.03) in x,y.  Each frame looks
    at a planar lattice of world points from U(0.3,1)^3; every camera sees every point."""
    rng = np.random.default_rng(seed)
    C, F = n_cams, n_frames
    q_c_true = np.tile(np.array([1.0, 0, 0, 0]), (C, 1))
    t_c_true = np.zeros((C, 3))
    t_c_true[1:, :2] = rng.uniform(-0.03, 0.03, size=(C - 1, 2))
    q_c, t_c = _perturb_pose(rng, q_c_true, t_c_true, cam_trans, cam_rot_deg)
    q_c[0], t_c[0] = q_c_true[0], t_c_true[0]
    # frames: translation U(0.3,1)^3, rotation rows (forward, right, up) as in the reference test
    t_f_true = rng.uniform(0.3, 1.0, size=(F, 3))
    fwd = t_f_true / np.linalg.norm(t_f_true, axis=1, keepdims=True)
    right = np.cross(np.array([0.0, 1.0, 0.0]), fwd)
    right /= np.linalg.norm(right, axis=1, keepdims=True)
    up = np.cross(fwd, right)
    Rf = np.stack([fwd, right, up], axis=1)
    q_f_true = rot_to_quat(Rf)
    q_f, t_f = _perturb_pose(rng, q_f_true, t_f_true, frame_trans, frame_rot_deg)
    gx, gy = grid
    xs = (np.arange(gx) - (gx - 1) / 2.0) * pitch
    ys = (np.arange(gy) - (gy - 1) / 2.0) * pitch
    # lattice in the plane facing the rig: world axes chosen so that it is seen at depth ~|t|
    lat = np.stack(np.meshgrid(xs, ys, indexing="ij"), axis=-1).reshape(-1, 2)
    P = lat.shape[0]
    pts = np.zeros((F, P, 3))
    pts[:, :, 1] = lat[:, 0]
    pts[:, :, 2] = lat[:, 1]
    # world -> rig -> camera (all cameras)
    Rc = quat_to_rot(q_c_true)
    p_rig = np.einsum("fij,fpj->fpi", quat_to_rot(q_f_true), pts) + t_f_true[:, None, :]
    p_cam = np.einsum("cij,fpj->fcpi", Rc, p_rig) + t_c_true[None, :, None, :]
    uu = p_cam[..., 0] / p_cam[..., 2] + rng.uniform(-noise_2d, noise_2d, size=p_cam.shape[:-1])
    vv = p_cam[..., 1] / p_cam[..., 2] + rng.uniform(-noise_2d, noise_2d, size=p_cam.shape[:-1])
    if outlier_fraction > 0:
        bad = rng.uniform(size=uu.shape) < outlier_fraction
        uu = np.where(bad, uu + rng.uniform(-0.05, 0.05, size=uu.shape), uu)
        vv = np.where(bad, vv + rng.uniform(-0.05, 0.05, size=vv.shape), vv)
    Xw = (pts + rng.uniform(-noise_3d, noise_3d, size=pts.shape)).reshape(-1, 3).astype(np.float32)
    pid = np.arange(F * P, dtype=np.int64).reshape(F, P)
    if order == "frame-major":  # the reference test: per point, every camera (point-major inside a frame)
        obs_cam = np.tile(np.arange(C, dtype=np.int32), (F, P, 1)).reshape(-1)
        obs_point = np.repeat(pid[:, :, None], C, axis=2).reshape(-1)
        u = np.transpose(uu, (0, 2, 1)).reshape(-1)
        v = np.transpose(vv, (0, 2, 1)).reshape(-1)
    else:  # camera-major inside a frame
        obs_cam = np.tile(np.repeat(np.arange(C, dtype=np.int32), P), F)
        obs_point = np.repeat(pid[:, None, :], C, axis=1).reshape(-1)
        u, v = uu.reshape(-1), vv.reshape(-1)
    frame_offsets = np.arange(F + 1, dtype=np.int64) * (C * P)
    frozen = np.zeros(C, np.uint8)
    frozen[0] = 1
    return RigScene(u.astype(np.float32), v.astype(np.float32), obs_cam, obs_point, frame_offsets, Xw,
                    q_c, t_c, frozen, q_f, t_f, q_c_true, t_c_true, q_f_true, t_f_true)
