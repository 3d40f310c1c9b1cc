"""GPU tests through the reference-facing API: the pycalibrator module (pybind11 over the host
C++ classes over the C-ABI) against the oracle, on identical inputs."""
import numpy as np
import pytest

import oracle
from camera_calibrator_b200 import host, synth
from helpers import TOL_PARAM, quat_to_rot

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pc():
    return host.load()


def _views(d):
    img = [np.stack([d.u[a:b], d.v[a:b]], 1) for a, b in zip(d.view_offsets[:-1], d.view_offsets[1:])]
    world = [np.stack([d.X[a:b], d.Y[a:b], d.Z[a:b]], 1) for a, b in zip(d.view_offsets[:-1], d.view_offsets[1:])]
    return img, world


def test_calibrator_estimate_matches_oracle(pc):
    """Calibrator(1600,1000).Estimate(img_points, world_points) as cam_calibration.py:307-313 calls it."""
    d = oracle.datagen_planar(20, 54)
    img, world = _views(d)
    c = pc.Calibrator(1600, 1000)
    c.Estimate(img_points=img, world_points=world)
    K, dist = c.GetK(), c.GetDistortion()
    ref, _, _, rrep = oracle.estimate(d.u, d.v, d.X, d.Y, d.Z, d.view_offsets)
    got = np.array([K[0, 0], K[1, 1], K[0, 2], K[1, 2], *dist], dtype=np.float64)
    # API precision is float32 (calibrator.cpp:327-335): equality after rounding, 2 ulp slack
    assert np.max(np.abs(got - ref) / np.abs(ref)) <= 2.5e-7
    assert K[0, 1] == 0 and K[1, 0] == 0 and K[2, 2] == 1  # zero entries stay exactly zero (test_calibrator.cpp:69)
    rep = c.GetLastReport()
    assert rep["termination"] in (1, 2, 3) and abs(rep["rms"] - rrep["rms"]) <= 1e-9
    q, t = c.GetOptimizedPoses()
    assert q.shape == (20, 4) and t.shape == (20, 3)


def test_calibrator_frozen_k3_like_the_python_caller(pc):
    d = oracle.datagen_planar(10, 54, seed=8)
    img, world = _views(d)
    c = pc.Calibrator(1600, 1000)
    c.ForceDistortionToConstant(4)
    c.Estimate(img, world)
    ref, _, _, _ = oracle.estimate(d.u, d.v, d.X, d.Y, d.Z, d.view_offsets, const_mask=1 << 8)
    got = np.array([c.GetK()[0, 0], c.GetK()[1, 1], c.GetK()[0, 2], c.GetK()[1, 2], *c.GetDistortion()], np.float64)
    assert got[8] == 0.0
    assert np.max(np.abs(got[:8] - ref[:8]) / np.abs(ref[:8])) <= 2.5e-7


def test_calibrator_optimize_from_given_poses(pc):
    d = synth.planar_views(40, 80, seed=3)
    intr, q, t = synth.perturbed_start(d)
    img, world = _views(d)
    c = pc.Calibrator(1600, 1000)
    K0 = np.array([[intr[0], 0, intr[2]], [0, intr[1], intr[3]], [0, 0, 1]], np.float32)
    c.SetK(K0)
    qf, tf = q.astype(np.float32), t.astype(np.float32)
    c.Optimize(img, world, qf, tf)
    intr32 = np.array([K0[0, 0], K0[1, 1], K0[0, 2], K0[1, 2], 0, 0, 0, 0, 0], np.float64)
    ref, rq, rt, _ = oracle.intrinsics_solve(d.u, d.v, d.X, d.Y, d.Z, d.view_offsets, intr32,
                                             qf.astype(np.float64), tf.astype(np.float64))
    got = np.array([c.GetK()[0, 0], c.GetK()[1, 1], c.GetK()[0, 2], c.GetK()[1, 2], *c.GetDistortion()], np.float64)
    assert np.max(np.abs(got - ref) / np.abs(ref)) <= 2.5e-7
    gq, gt = c.GetOptimizedPoses()
    assert np.max(np.abs(quat_to_rot(gq) - quat_to_rot(rq))) <= TOL_PARAM


def _affine(q, t):
    T = np.eye(4, dtype=np.float32)
    T[:3, :3] = quat_to_rot(q)
    T[:3, 3] = t
    return T


def _build(pc, s, bulk=False):
    e = pc.ExtrinsicsCalibrator()
    C = s.q_c.shape[0]
    for c in range(C):
        e.AddCameraTRig(_affine(s.q_c[c], s.t_c[c]), freeze=bool(s.cam_frozen[c]))
    F = len(s.frame_offsets) - 1
    for f in range(F):
        e.AddObservationFrame(_affine(s.q_f[f], s.t_f[f]))
        a, b = s.frame_offsets[f], s.frame_offsets[f + 1]
        pts = np.unique(s.obs_point[a:b])
        if bulk:
            img = np.stack([s.u[a:b], s.v[a:b]], 1).reshape(len(pts), C, 2)
            e.AddObservations(f, s.Xw[pts], list(range(C)), img)
        else:
            ids = {int(p): e.AddWorldPoint(f, s.Xw[p]) for p in pts}
            for o in range(a, b):
                e.AddObservation(int(s.obs_cam[o]), ids[int(s.obs_point[o])], [s.u[o], s.v[o]])
    return e


def _oracle_rig(s):
    # the host class packs float32 transforms -> quaternion in double (extrinsics_calibrator.cpp:116-130)
    qc = synth.rot_to_quat(quat_to_rot(s.q_c).astype(np.float32).astype(np.float64))
    qf = synth.rot_to_quat(quat_to_rot(s.q_f).astype(np.float32).astype(np.float64))
    return oracle.rig_solve(s.u, s.v, s.obs_cam, s.obs_point, s.frame_offsets, s.Xw, qc,
                            s.t_c.astype(np.float32), s.cam_frozen, qf, s.t_f.astype(np.float32))


def test_extrinsics_calibrator_optimize_matches_oracle(pc, tmp_path):
    s = synth.rig_scene(n_cams=3, n_frames=60, grid=(5, 4), seed=4)
    e = _build(pc, s)
    e.Optimize()
    rqc, rtc, rqf, rtf, rcost, rrep = _oracle_rig(s)
    for c in range(3):
        T = e.GetCameraTRig(c)
        assert np.max(np.abs(T[:3, :3] - quat_to_rot(rqc[c]))) <= 1e-6
        assert np.max(np.abs(T[:3, 3] - rtc[c])) <= 1e-6
        assert np.array_equal(T[3], [0, 0, 0, 1])
    for f in (0, 17, 59):
        T = e.GetObservationFrame(f)
        assert np.max(np.abs(T[:3, :3] - quat_to_rot(rqf[f]))) <= 1e-6
        assert np.max(np.abs(T[:3, 3] - rtf[f])) <= 2e-6
    # per-observation cost written back in insertion order (extrinsics_calibrator.cpp:219-225)
    got_cost = np.array([o["cost"] for f in range(60) for o in e.GetObservations(f)])
    assert np.max(np.abs(got_cost - rcost)) <= 1e-8 * np.max(rcost)
    assert abs(e.GetLastReport()["rms"] - rrep["rms"]) <= 1e-9
    # Serialize -> Parse into a FRESH object -> Optimize again: same answer (wire format carries everything)
    e2 = _build(pc, s, bulk=True)
    fn = str(tmp_path / "rig.json")
    e2.Serialize(fn)
    e3 = pc.ExtrinsicsCalibrator()
    e3.Parse(fn)
    e3.Optimize()
    for c in range(3):
        assert np.max(np.abs(e3.GetCameraTRig(c) - e.GetCameraTRig(c))) <= 1e-6


def test_calibrator_lens_round_trip(pc):
    """Distort then Undistort (calibrator.cpp:97-166) through the reference-facing class."""
    c = pc.Calibrator(1600, 1000)
    K = np.array([[1000, 0, 800], [0, 1000, 500], [0, 0, 1]], np.float32)
    dist = np.array([-4e-2, 5e-4, 1e-3, 2e-5, -3e-4], np.float32)
    c.SetK(K=K)
    c.SetDistortion(distortion=dist)
    rng = np.random.default_rng(0)
    norm = rng.uniform(-0.5, 0.5, (50, 2)).astype(np.float32)
    px = c.Distort(normalized_points=norm)
    back = c.Undistort(img_points=px)
    assert px.shape == (50, 2) and back.shape == (50, 2)
    assert np.max(np.abs(back - norm)) < 2e-6  # float32 pixels in, float32 out
    # lists of rows work like the reference's stl+eigen casters
    assert np.allclose(c.Distort([[0.1, 0.2], [0.0, 0.0]])[1], [800, 500])


def test_estimate_opencv_matches_cv2_calibrate_camera(pc):
    """EstimateOpenCv (calibrator.cpp:16-45) against the function the reference calls there."""
    import cv2

    d = oracle.datagen_planar(12, 80, seed=4)
    img, world = _views(d)
    crit = (cv2.TERM_CRITERIA_EPS + cv2.TERM_CRITERIA_MAX_ITER, 2000, 1e-16)
    rms, K, dist, _, _ = cv2.calibrateCamera([w.astype(np.float32) for w in world], [i.astype(np.float32) for i in img],
                                             (1600, 1000), None, None, criteria=crit)
    c = pc.Calibrator(1600, 1000)
    c.ForceDistortionToConstant(4)  # ignored by this entry point, as cv::calibrateCamera ignores it
    c.EstimateOpenCv(img_points=img, world_points=world)
    Kg, dg = c.GetK(), c.GetDistortion()
    assert abs(c.GetLastReport()["rms"] - rms) <= 1e-8
    for got, want in ((Kg[0, 0], K[0, 0]), (Kg[1, 1], K[1, 1]), (Kg[0, 2], K[0, 2]), (Kg[1, 2], K[1, 2])):
        assert abs(got - want) / want <= 2e-6
    assert np.max(np.abs(dg - dist.ravel()[:5])) <= 1e-5 and dg[4] != 0.0
