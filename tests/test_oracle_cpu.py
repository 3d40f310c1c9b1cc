"""CPU tests that PIN the oracle (test infrastructure) before it is trusted as the parity
reference: golden fixtures, cv2.calibrateCamera (the function the reference's EstimateOpenCv
calls, calibrator.cpp:38), scipy.optimize.least_squares, torch.float64 autograd Jacobians, and
the reference's own test assertions."""
import json
import os

import numpy as np
import pytest

import oracle
from helpers import quat_to_rot, start_from_estimate

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "intrinsics_golden.json")))


@pytest.mark.parametrize("name", sorted(GOLD))
def test_generator_matches_golden_stream(name):
    """std::mt19937 + uniform_real_distribution<float> + float32 geometry: bit-exact regression."""
    g = GOLD[name]
    d = oracle.datagen_planar(g["n_views"], g["n_points"])
    assert d.u.size == d.X.size == g["n_views"] * g["n_points"]  # counts (test_calibrator.cpp:30-43)
    assert d.n_rejected == g["n_rejected"]
    for k in "uvXY":
        assert [float(x) for x in getattr(d, k)[:8]] == g["first_obs"][k]
        assert float(np.sum(getattr(d, k).astype(np.float64))) == g["checksum"][k]
    assert np.all(d.Z == 0.0)
    assert np.all((d.u >= 0) & (d.u < 1599) & (d.v >= 0) & (d.v < 999))  # in-image rule, data_generator.cpp:25-29


def test_mt19937_stream_is_the_reference_one():
    """mt19937{0} + uniform_real_distribution<float>(0.2,1) -> 0.639050841, ..."""
    K = oracle.FIXTURE_K
    d = oracle.datagen_planar(1, 1, noise_px=0.0)
    # the first three draws are the plane distances; recover them from the plane through the rays
    import numpy.random as npr

    raw = npr.RandomState(0).randint(0, 2**32, size=3, dtype=np.uint64).astype(np.float32)
    want = np.float32(0.8) * (raw / np.float32(2.0**32)) + np.float32(0.2)
    assert np.allclose(want, [0.639050841, 0.674275696, 0.77215147], atol=1e-7)
    Kinv = np.linalg.inv(K.astype(np.float64))
    corners = np.array([[0, 0, 1.0], [1599.0, 0, 1.0], [1599.0, 999.0, 1.0]]) @ Kinv.T
    corners /= np.linalg.norm(corners, axis=1, keepdims=True)
    pts = corners * want[:, None].astype(np.float64)
    n = np.linalg.solve(pts, np.ones(3))
    assert np.allclose(d.planes[0, :3], n, rtol=2e-5)


@pytest.mark.parametrize("name", sorted(GOLD))
def test_oracle_matches_cv2_golden(name):
    """Converged oracle == converged cv2.calibrateCamera (golden) on the same inputs."""
    g = GOLD[name]
    d = oracle.datagen_planar(g["n_views"], g["n_points"])
    intr, q, t, rep = oracle.estimate(d.u, d.v, d.X, d.Y, d.Z, d.view_offsets)
    assert rep["termination"] in (1, 2, 3)
    cv = np.array(g["cv2"]["K"] + g["cv2"]["dist"])
    # cv2 stops on its own (looser) criterion: 1e-6 relative on K, 2e-4 on distortion, 1e-9 on RMS
    assert np.max(np.abs(intr[:4] - cv[:4]) / cv[:4]) <= 1e-6
    assert np.max(np.abs(intr[4:] - cv[4:]) / np.abs(cv[4:])) <= 2e-4
    assert abs(rep["rms"] - g["cv2"]["rms"]) <= 1e-9
    # regression against the oracle's own committed solution
    assert np.max(np.abs(intr - g["oracle"]["intr"]) / np.abs(g["oracle"]["intr"])) <= 1e-7
    assert abs(rep["rms"] - g["oracle"]["rms"]) <= 1e-11


def test_reference_test_assertion_k_within_one_percent():
    """test_calibrator.cpp:62-71: every non-zero K entry within 1 %, zero entries exactly zero."""
    d = oracle.datagen_planar(5, 100)
    intr, _, _, _ = oracle.estimate(d.u, d.v, d.X, d.Y, d.Z, d.view_offsets)
    K = oracle.FIXTURE_K
    for got, want in ((intr[0], K[0, 0]), (intr[1], K[1, 1]), (intr[2], K[0, 2]), (intr[3], K[1, 2])):
        assert (got - want) / want < 0.01
        assert abs(got - want) / want < 0.01


def test_oracle_matches_live_cv2():
    import cv2

    d = oracle.datagen_planar(8, 60, seed=11)
    obj = [np.stack([d.X[a:b], d.Y[a:b], d.Z[a:b]], 1) for a, b in zip(d.view_offsets[:-1], d.view_offsets[1:])]
    img = [np.stack([d.u[a:b], d.v[a:b]], 1) for a, b in zip(d.view_offsets[:-1], d.view_offsets[1:])]
    crit = (cv2.TERM_CRITERIA_EPS + cv2.TERM_CRITERIA_MAX_ITER, 2000, 1e-16)
    rms, K, dist, _, _ = cv2.calibrateCamera(obj, img, (1600, 1000), None, None, criteria=crit)
    intr, _, _, rep = oracle.estimate(d.u, d.v, d.X, d.Y, d.Z, d.view_offsets)
    assert abs(rep["rms"] - rms) <= 1e-9
    assert abs(intr[0] - K[0, 0]) / K[0, 0] <= 1e-6 and abs(intr[3] - K[1, 2]) / K[1, 2] <= 1e-6


def _torch_residuals(intr, q, t, X, meas):
    import torch

    qn = q / q.norm(dim=-1, keepdim=True)
    w, v = qn[..., :1], qn[..., 1:]
    c = torch.cross(v, X, dim=-1)
    p = X + 2 * w * c + 2 * torch.cross(v, c, dim=-1) + t
    xn, yn = p[..., 0] / p[..., 2], p[..., 1] / p[..., 2]
    fx, fy, px, py, k1, k2, p1, p2, k3 = intr
    r2 = xn * xn + yn * yn
    m = 1 + k1 * r2 + k2 * r2**2 + k3 * r2**3
    xd = xn * m + 2 * p1 * xn * yn + p2 * (r2 + 2 * xn * xn)
    yd = yn * m + 2 * p2 * xn * yn + p1 * (r2 + 2 * yn * yn)
    return torch.stack([fx * xd + px - meas[..., 0], fy * yd + py - meas[..., 1]], -1)


def test_jacobian_blocks_match_torch_autograd():
    """U, g_u, V, W, g_v of the oracle (Jets + Ceres plus-Jacobian) vs torch.float64 autograd with
    the manifold perturbation q <- exp_q(delta) * q applied explicitly."""
    import torch

    d = oracle.datagen_planar(3, 12, seed=4)
    intr, q, t = start_from_estimate(oracle, d)
    intr[4:] = [-0.03, 0.002, 0.001, -0.0005, 0.0007]
    e = oracle.intrinsics_eval(d.u, d.v, d.X, d.Y, d.Z, d.view_offsets, intr, q, t)
    X = torch.tensor(np.stack([d.X, d.Y, d.Z], 1).astype(np.float64)).reshape(3, 12, 3)
    meas = torch.tensor(np.stack([d.u, d.v], 1).astype(np.float64)).reshape(3, 12, 2)
    q0 = torch.tensor(q)
    t0 = torch.tensor(t)

    def f(params):
        i9, dl, dt = params[:9], params[9:18].reshape(3, 3), params[18:27].reshape(3, 3)
        n = dl.norm(dim=-1, keepdim=True)
        qd = torch.cat([torch.cos(n), torch.sinc(n / np.pi) * dl], -1)
        w0, x0, y0, z0 = qd.unbind(-1)
        w1, x1, y1, z1 = q0.unbind(-1)
        qq = torch.stack([w0 * w1 - x0 * x1 - y0 * y1 - z0 * z1, w0 * x1 + x0 * w1 + y0 * z1 - z0 * y1,
                          w0 * y1 - x0 * z1 + y0 * w1 + z0 * x1, w0 * z1 + x0 * y1 - y0 * x1 + z0 * w1], -1)
        return _torch_residuals(i9, qq[:, None, :], (t0 + dt)[:, None, :], X, meas).reshape(-1)

    p0 = torch.cat([torch.tensor(intr), torch.full((9,), 1e-300, dtype=torch.float64), torch.zeros(9, dtype=torch.float64)])
    J = torch.autograd.functional.jacobian(f, p0).numpy()
    r = f(p0).detach().numpy()
    assert np.max(np.abs(r.reshape(-1, 2) - e["res"])) < 1e-9
    Ji = J[:, :9]
    assert np.allclose(Ji.T @ Ji, e["U"], rtol=1e-9, atol=1e-6)
    assert np.allclose(Ji.T @ r, e["gu"], rtol=1e-9, atol=1e-6)
    for v in range(3):
        rows = slice(v * 24, (v + 1) * 24)
        Jp = np.concatenate([J[rows, 9 + 3 * v:12 + 3 * v], J[rows, 18 + 3 * v:21 + 3 * v]], 1)
        assert np.allclose(Jp.T @ Jp, e["V"][v], rtol=1e-8, atol=1e-5)
        assert np.allclose(Jp.T @ Ji[rows], e["W"][v], rtol=1e-8, atol=1e-5)
        assert np.allclose(Jp.T @ r[rows], e["gv"][v], rtol=1e-8, atol=1e-5)


def test_oracle_matches_scipy_least_squares():
    """Independent solver on the same residuals (fixed poses' rotation via rotation vectors)."""
    from scipy.optimize import least_squares
    from scipy.spatial.transform import Rotation

    d = oracle.datagen_planar(6, 40, seed=9)
    intr0, q0, t0 = start_from_estimate(oracle, d)
    ref_i, ref_q, ref_t, rep = oracle.intrinsics_solve(d.u, d.v, d.X, d.Y, d.Z, d.view_offsets, intr0, q0, t0)
    Xw = np.stack([d.X, d.Y, d.Z], 1).astype(np.float64).reshape(6, 40, 3)
    meas = np.stack([d.u, d.v], 1).astype(np.float64).reshape(6, 40, 2)

    def res(p):
        i9, rv, tt = p[:9], p[9:27].reshape(6, 3), p[27:45].reshape(6, 3)
        R = Rotation.from_rotvec(rv).as_matrix()
        pc = np.einsum("vij,vpj->vpi", R, Xw) + tt[:, None, :]
        xn, yn = pc[..., 0] / pc[..., 2], pc[..., 1] / pc[..., 2]
        fx, fy, px, py, k1, k2, p1, p2, k3 = i9
        r2 = xn * xn + yn * yn
        m = 1 + k1 * r2 + k2 * r2**2 + k3 * r2**3
        xd = xn * m + 2 * p1 * xn * yn + p2 * (r2 + 2 * xn * xn)
        yd = yn * m + 2 * p2 * xn * yn + p1 * (r2 + 2 * yn * yn)
        return np.stack([fx * xd + px - meas[..., 0], fy * yd + py - meas[..., 1]], -1).ravel()

    rv0 = Rotation.from_matrix(quat_to_rot(q0)).as_rotvec()
    sol = least_squares(res, np.concatenate([intr0, rv0.ravel(), t0.ravel()]), method="lm", x_scale="jac",
                        xtol=1e-15, ftol=1e-15, gtol=1e-15, max_nfev=4000)
    rms = np.sqrt(np.sum(sol.fun**2) / (sol.fun.size / 2))
    assert abs(rms - rep["rms"]) <= 1e-9
    assert np.max(np.abs(sol.x[:4] - ref_i[:4]) / ref_i[:4]) <= 1e-6


def test_two_starts_reach_the_same_minimum():
    """Parity is only meaningful at the converged minimum: two different
    starting points must agree at BASELINE.json's 1e-6 / 1e-9."""
    d = oracle.datagen_planar(20, 54)
    intr0, q0, t0 = start_from_estimate(oracle, d)
    a = oracle.intrinsics_solve(d.u, d.v, d.X, d.Y, d.Z, d.view_offsets, intr0, q0, t0)
    rng = np.random.default_rng(0)
    intr1 = intr0 * (1 + 0.01 * rng.uniform(-1, 1, 9))
    t1 = t0 + 0.005 * rng.uniform(-1, 1, t0.shape)
    b = oracle.intrinsics_solve(d.u, d.v, d.X, d.Y, d.Z, d.view_offsets, intr1, q0, t1)
    assert np.max(np.abs(a[0] - b[0]) / np.abs(a[0])) <= 1e-6
    assert abs(a[3]["rms"] - b[3]["rms"]) <= 1e-9


def test_homography_known_answer():
    """test_geometry.cpp:172-194: EstimateHomography recovers a random H to 1e-4 (H(2,2) = 1)."""
    rng = np.random.default_rng(0)
    for _ in range(10):
        H = rng.uniform(-1, 1, (3, 3)).astype(np.float32)
        H /= H[2, 2]
        p1 = np.concatenate([rng.uniform(-1, 1, (10, 2)), np.ones((10, 1))], 1).astype(np.float32)
        p2 = (H @ p1.T).T
        p2 /= p2[:, 2:]
        H2 = oracle.homography(p1[:, 0], p1[:, 1], p2[:, 0], p2[:, 1])
        H2 = H2 / H2[2, 2]
        assert np.max(np.abs(H - H2)) <= 1e-4


def test_rig_oracle_huber_and_costs():
    """Rig oracle: robust cost = sum of per-observation costs; outliers land in the linear zone."""
    from camera_calibrator_b200 import synth

    s = synth.rig_scene(n_cams=3, n_frames=25, grid=(4, 3), seed=7, outlier_fraction=0.1)
    qc, tc, qf, tf, cost, rep = oracle.rig_solve(s.u, s.v, s.obs_cam, s.obs_point, s.frame_offsets, s.Xw, s.q_c,
                                                 s.t_c, s.cam_frozen, s.q_f, s.t_f)
    assert rep["termination"] in (1, 2, 3)
    assert abs(cost.sum() - rep["final_cost"]) <= 1e-12 * rep["final_cost"]
    a = float(np.float32(3.0) / np.float32(500.0))
    assert np.any(cost > 0.5 * a * a)
    assert np.array_equal(qc[0], s.q_c[0]) and np.array_equal(tc[0], s.t_c[0])  # frozen camera
    assert np.max(np.abs(tc - s.t_c_true)) < 5e-3


def test_rig_oracle_matches_scipy_block_huber():
    """Independent pin of the rig oracle (the reference has no assertion on this path): scipy
    least_squares on the robustified residuals r * sqrt(rho(s) / s), s = ||r||^2 per 2-vector block
    - Huber per OBSERVATION as Ceres applies it (extrinsics_calibrator.cpp:175-176), not per scalar
    like scipy's own loss='huber' - must reach the oracle's cost and camera extrinsics."""
    from scipy.optimize import least_squares
    from scipy.spatial.transform import Rotation

    from camera_calibrator_b200 import synth

    C, F = 2, 12
    s = synth.rig_scene(n_cams=C, n_frames=F, grid=(4, 3), seed=11, outlier_fraction=0.08)
    qc, tc, qf, tf, cost, rep = oracle.rig_solve(s.u, s.v, s.obs_cam, s.obs_point, s.frame_offsets, s.Xw, s.q_c,
                                                 s.t_c, s.cam_frozen, s.q_f, s.t_f)
    assert rep["termination"] in (1, 2, 3)
    a = float(np.float32(3.0) / np.float32(500.0))
    frame_of_obs = np.repeat(np.arange(F), np.diff(s.frame_offsets))
    Xw = s.Xw.astype(np.float64)[s.obs_point]
    uv = np.stack([s.u, s.v], 1).astype(np.float64)
    cam = s.obs_cam.astype(np.int64)
    Rc0, Rf0 = quat_to_rot(s.q_c), quat_to_rot(s.q_f)

    def unpack(p):
        # camera 0 is frozen; cameras 1.. and all frames: rotation-vector increment (left) + translation
        dc = p[:6 * (C - 1)].reshape(C - 1, 6)
        df = p[6 * (C - 1):].reshape(F, 6)
        Rc = np.concatenate([Rc0[:1], np.einsum("cij,cjk->cik", Rotation.from_rotvec(dc[:, :3]).as_matrix(), Rc0[1:])])
        tcam = np.concatenate([s.t_c[:1], s.t_c[1:] + dc[:, 3:]])
        Rf = np.einsum("fij,fjk->fik", Rotation.from_rotvec(df[:, :3]).as_matrix(), Rf0)
        return Rc, tcam, Rf, s.t_f + df[:, 3:]

    def robust(p):
        Rc, tcam, Rf, tfr = unpack(p)
        x_rig = np.einsum("nij,nj->ni", Rf[frame_of_obs], Xw) + tfr[frame_of_obs]
        x = np.einsum("nij,nj->ni", Rc[cam], x_rig) + tcam[cam]
        r = np.stack([x[:, 0] / x[:, 2], x[:, 1] / x[:, 2]], 1) - uv
        sq = np.maximum(np.sum(r * r, 1), 1e-300)
        rho = np.where(sq <= a * a, sq, 2 * a * np.sqrt(sq) - a * a)
        return (r * np.sqrt(rho / sq)[:, None]).ravel()

    sol = least_squares(robust, np.zeros(6 * (C - 1) + 6 * F), method="trf", x_scale="jac", xtol=1e-15, ftol=1e-15,
                        gtol=1e-14, max_nfev=400)
    assert abs(0.5 * np.sum(sol.fun**2) - rep["final_cost"]) <= 1e-7 * rep["final_cost"]
    Rc, tcam, _, _ = unpack(sol.x)
    assert np.max(np.abs(tcam - tc)) <= 2e-6
    assert np.max(np.abs(Rc - quat_to_rot(qc))) <= 2e-6


def _rig_golden():
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "rig_golden.json")))
    i = g["inputs"]
    args = (np.array(i["u"], np.float32), np.array(i["v"], np.float32), np.array(i["obs_cam"], np.int32),
            np.array(i["obs_point"], np.int64), np.array(i["frame_offsets"], np.int64), np.array(i["Xw"], np.float32),
            np.array(i["q_c"]), np.array(i["t_c"]), np.array(i["cam_frozen"], np.uint8), np.array(i["q_f"]),
            np.array(i["t_f"]))
    return g, args


def test_rig_oracle_matches_golden_fixture():
    """Committed rig fixture (inputs + an independent scipy block-Huber solution, tests/golden/
    make_golden.py): the oracle must reproduce cost, extrinsics and per-observation costs."""
    g, args = _rig_golden()
    qc, tc, qf, tf, cost, rep = oracle.rig_solve(*args)
    sp = g["scipy"]
    assert abs(rep["final_cost"] - sp["cost"]) <= 1e-7 * sp["cost"]
    assert np.max(np.abs(tc - np.array(sp["t_c"]))) <= 2e-6
    assert np.max(np.abs(quat_to_rot(qc) - np.array(sp["R_c"]))) <= 2e-6
    assert np.max(np.abs(tf - np.array(sp["t_f"]))) <= 2e-5
    assert np.allclose(cost, np.array(sp["obs_cost"]), rtol=1e-3, atol=1e-9)
    assert abs(rep["final_cost"] - g["oracle"]["final_cost"]) <= 1e-12 * g["oracle"]["final_cost"]  # regression pin


def test_rig_jacobian_blocks_match_torch_autograd():
    """Rig oracle (two chained transforms, Ceres quaternion manifold, Huber corrector with
    rho'' <= 0 = residual and Jacobian scaled by sqrt(rho')) vs torch.float64 autograd with the manifold perturbation
    q <- exp_q(delta) * q applied explicitly to frame and camera poses: res, V_f, W_fc, g_f, U, g_c, cost."""
    import torch

    from camera_calibrator_b200 import synth

    C_, F_, G = 3, 4, (3, 2)
    P = G[0] * G[1]
    s = synth.rig_scene(n_cams=C_, n_frames=F_, grid=G, seed=11, outlier_fraction=0.3, order="frame-major")
    a = float(np.float32(3.0) / np.float32(500.0))
    # evaluated at the TRUE poses: inliers sit below the Huber threshold, the 30 % gross outliers above it
    e = oracle.rig_eval(s.u, s.v, s.obs_cam, s.obs_point, s.frame_offsets, s.Xw, s.q_c_true, s.t_c_true, s.q_f_true,
                        s.t_f_true, a)
    n = s.u.size
    frame_of = np.repeat(np.arange(F_), np.diff(s.frame_offsets))
    Xw = torch.tensor(s.Xw.astype(np.float64))[torch.tensor(s.obs_point)]
    meas = torch.tensor(np.stack([s.u, s.v], 1).astype(np.float64))
    qc0, tc0 = torch.tensor(s.q_c_true), torch.tensor(s.t_c_true)
    qf0, tf0 = torch.tensor(s.q_f_true), torch.tensor(s.t_f_true)
    cam = torch.tensor(s.obs_cam.astype(np.int64))
    frm = torch.tensor(frame_of)

    def plus(q, dl):
        nrm = dl.norm(dim=-1, keepdim=True)
        qd = torch.cat([torch.cos(nrm), torch.sinc(nrm / np.pi) * dl], -1)
        w0, x0, y0, z0 = qd.unbind(-1)
        w1, x1, y1, z1 = q.unbind(-1)
        return torch.stack([w0 * w1 - x0 * x1 - y0 * y1 - z0 * z1, w0 * x1 + x0 * w1 + y0 * z1 - z0 * y1,
                            w0 * y1 - x0 * z1 + y0 * w1 + z0 * x1, w0 * z1 + x0 * y1 - y0 * x1 + z0 * w1], -1)

    def rotate(q, x):  # ceres::QuaternionRotatePoint: normalises q
        q = q / q.norm(dim=-1, keepdim=True)
        w, v = q[..., :1], q[..., 1:]
        t2 = 2.0 * torch.cross(v, x, dim=-1)
        return x + w * t2 + torch.cross(v, t2, dim=-1)

    def f(params):
        df = params[:6 * F_].reshape(F_, 6)
        dc = params[6 * F_:].reshape(C_, 6)
        qf, tf = plus(qf0, df[:, :3]), tf0 + df[:, 3:]
        qc, tc = plus(qc0, dc[:, :3]), tc0 + dc[:, 3:]
        x_rig = rotate(qf[frm], Xw) + tf[frm]
        x = rotate(qc[cam], x_rig) + tc[cam]
        return (x[:, :2] / x[:, 2:3] - meas).reshape(-1)

    p0 = torch.zeros(6 * (F_ + C_), dtype=torch.float64)
    p0[:] = 1e-300  # (the norm in plus() is not differentiable at exactly 0)
    J = torch.autograd.functional.jacobian(f, p0).numpy()
    r = f(p0).detach().numpy().reshape(n, 2)
    sq = np.sum(r * r, axis=1)
    rho1 = np.where(sq <= a * a, 1.0, a / np.sqrt(np.maximum(sq, 1e-300)))
    rho = np.where(sq <= a * a, sq, 2 * a * np.sqrt(sq) - a * a)
    assert np.any(rho1 < 1.0) and np.any(rho1 == 1.0)  # both Huber zones are exercised
    w = np.repeat(np.sqrt(rho1), 2)
    Jw, rw = J * w[:, None], r.reshape(-1) * w
    assert abs(0.5 * rho.sum() - e["cost"]) <= 1e-12 * e["cost"]
    assert np.max(np.abs(rw.reshape(n, 2) - e["res"])) < 1e-12
    Jc = Jw[:, 6 * F_:]
    assert np.allclose(Jc.T @ Jc, e["U"], rtol=1e-9, atol=1e-9)
    assert np.allclose(Jc.T @ rw, e["gu"], rtol=1e-9, atol=1e-10)
    for fi in range(F_):
        rows = np.repeat(frame_of == fi, 2)
        Jf = Jw[rows, 6 * fi:6 * fi + 6]
        assert np.allclose(Jf.T @ Jf, e["V"][fi], rtol=1e-9, atol=1e-9)
        assert np.allclose(Jf.T @ Jc[rows], e["W"][fi], rtol=1e-9, atol=1e-9)
        assert np.allclose(Jf.T @ rw[rows], e["gv"][fi], rtol=1e-9, atol=1e-10)
