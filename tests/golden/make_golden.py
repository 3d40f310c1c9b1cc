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
