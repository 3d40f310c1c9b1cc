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
