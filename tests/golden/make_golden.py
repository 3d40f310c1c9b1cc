"""Generates tests/golden/*.json.  Run from the repo root:  python tests/golden/make_golden.py

The reference (C++/Ceres/Eigen/OpenCV) cannot be built or imported in this image, so the golden
vectors are (a) the restated generator's first outputs for the reference fixture (regression pin of
the std::mt19937 stream + float arithmetic), (b) cv2.calibrateCamera — the function the reference's
own Calibrator::EstimateOpenCv calls (reference src/calibrator.cpp:38) and which minimises the same
cost — run to convergence on those inputs, and (c) the oracle's converged solution (regression pin).
"""
import json
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

out = {}
for name, (nv, npts) in {"reference_test_5x100": (5, 100), "c1_20x54": (20, 54)}.items():
    d = oracle.datagen_planar(nv, npts)
    obj = [np.stack([d.X[a:b], d.Y[a:b], d.Z[a:b]], 1) for a, b in zip(d.view_offsets[:-1], d.view_offsets[1:])]
    img = [np.stack([d.u[a:b], d.v[a:b]], 1) for a, b in zip(d.view_offsets[:-1], d.view_offsets[1:])]
    crit = (cv2.TERM_CRITERIA_EPS + cv2.TERM_CRITERIA_MAX_ITER, 2000, 1e-16)
    rms, K, dist, _, _ = cv2.calibrateCamera(obj, img, (1600, 1000), None, None, criteria=crit)
    intr, q, t, rep = oracle.estimate(d.u, d.v, d.X, d.Y, d.Z, d.view_offsets)
    out[name] = {
        "n_views": nv, "n_points": npts, "n_rejected": d.n_rejected,
        "first_obs": {k: [float(x) for x in getattr(d, k)[:8]] for k in "uvXY"},
        "checksum": {k: float(np.sum(getattr(d, k).astype(np.float64))) for k in "uvXY"},
        "cv2": {"rms": float(rms), "K": [float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2])],
                "dist": [float(x) for x in dist.ravel()]},
        "oracle": {"intr": [float(x) for x in intr], "rms": rep["rms"], "final_cost": rep["final_cost"]},
    }
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "intrinsics_golden.json"), "w"), indent=1)
print(json.dumps(out, indent=1)[:1500])


# ---- rig (ExtrinsicsCalibrator::Optimize): a small scene with outliers, its inputs, the oracle's
# solution and an INDEPENDENT solution (scipy least_squares on the block-Huber robustified residuals,
# the construction of tests/test_oracle_cpu.py::test_rig_oracle_matches_scipy_block_huber)
from scipy.optimize import least_squares  # noqa: E402
from scipy.spatial.transform import Rotation  # noqa: E402

from camera_calibrator_b200 import synth  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import quat_to_rot  # noqa: E402

C, F = 2, 12
s = synth.rig_scene(n_cams=C, n_frames=F, grid=(4, 3), seed=11, outlier_fraction=0.08)
qc, tc, qf, tf, cost, rep = oracle.rig_solve(s.u, s.v, s.obs_cam, s.obs_point, s.frame_offsets, s.Xw, s.q_c, s.t_c,
                                             s.cam_frozen, s.q_f, s.t_f)
a = float(np.float32(3.0) / np.float32(500.0))
frame_of_obs = np.repeat(np.arange(F), np.diff(s.frame_offsets))
Xw = s.Xw.astype(np.float64)[s.obs_point]
uv = np.stack([s.u, s.v], 1).astype(np.float64)
Rc0, Rf0 = quat_to_rot(s.q_c), quat_to_rot(s.q_f)


def unpack(p):
    dc = p[:6 * (C - 1)].reshape(C - 1, 6)
    df = p[6 * (C - 1):].reshape(F, 6)
    Rc = np.concatenate([Rc0[:1], np.einsum("cij,cjk->cik", Rotation.from_rotvec(dc[:, :3]).as_matrix(), Rc0[1:])])
    tcam = np.concatenate([s.t_c[:1], s.t_c[1:] + dc[:, 3:]])
    Rf = np.einsum("fij,fjk->fik", Rotation.from_rotvec(df[:, :3]).as_matrix(), Rf0)
    return Rc, tcam, Rf, s.t_f + df[:, 3:]


def robust(p):
    Rc, tcam, Rf, tfr = unpack(p)
    x_rig = np.einsum("nij,nj->ni", Rf[frame_of_obs], Xw) + tfr[frame_of_obs]
    x = np.einsum("nij,nj->ni", Rc[s.obs_cam], x_rig) + tcam[s.obs_cam]
    r = np.stack([x[:, 0] / x[:, 2], x[:, 1] / x[:, 2]], 1) - uv
    sq = np.maximum(np.sum(r * r, 1), 1e-300)
    rho = np.where(sq <= a * a, sq, 2 * a * np.sqrt(sq) - a * a)
    return (r * np.sqrt(rho / sq)[:, None]).ravel()


sol = least_squares(robust, np.zeros(6 * (C - 1) + 6 * F), method="trf", x_scale="jac", xtol=1e-15, ftol=1e-15,
                    gtol=1e-14, max_nfev=400)
Rc, tcam, Rf, tfr = unpack(sol.x)
per_obs = 0.5 * np.sum(sol.fun.reshape(-1, 2) ** 2, 1)
rig = {
    "inputs": {"u": s.u.tolist(), "v": s.v.tolist(), "obs_cam": s.obs_cam.tolist(), "obs_point": s.obs_point.tolist(),
               "frame_offsets": s.frame_offsets.tolist(), "Xw": s.Xw.tolist(), "q_c": s.q_c.tolist(),
               "t_c": s.t_c.tolist(), "cam_frozen": s.cam_frozen.tolist(), "q_f": s.q_f.tolist(), "t_f": s.t_f.tolist()},
    "scipy": {"cost": float(0.5 * np.sum(sol.fun ** 2)), "R_c": Rc.tolist(), "t_c": tcam.tolist(), "R_f": Rf.tolist(),
              "t_f": tfr.tolist(), "obs_cost": per_obs.tolist()},
    "oracle": {"final_cost": rep["final_cost"], "q_c": qc.tolist(), "t_c": tc.tolist(), "iterations": rep["iterations"]},
}
json.dump(rig, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "rig_golden.json"), "w"))
print("rig golden: scipy cost", rig["scipy"]["cost"], "oracle cost", rep["final_cost"])
