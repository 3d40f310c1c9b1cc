"""Target for compute-sanitizer (memcheck / racecheck): a small single-camera solve through the
multi-launch path and the persistent kernel, a small batch, a lens call and a small rig solve."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import camera_calibrator_b200 as cc
from camera_calibrator_b200 import rig_api, synth

ctx = cc.Context(0)
for nv, npts in ((20, 54), (150, 400)):
    d = synth.planar_views(nv, npts, seed=1)
    intr, q, t = synth.perturbed_start(d)
    for no_persist in ("0", "1"):
        os.environ["CCG_NO_PERSIST"] = no_persist
        got = ctx.intrinsics_solve(d.u, d.v, d.X, d.Y, d.Z, d.view_offsets, intr, q, t)
        print("intrinsics", nv, npts, "no_persist", no_persist, got[3])
xn = np.linspace(-0.4, 0.4, 4096)
K = np.array([[1000.0, 0, 640], [0, 1000.0, 360], [0, 0, 1]])
dist = np.array([0.1, -0.05, 0.001, 0.001, 0.01])
px, py = ctx.distort(xn, xn[::-1].copy(), K, dist)
ctx.undistort(px, py, K, dist)
for n_cams in (3, 7):
    s = synth.rig_scene(n_cams=n_cams, n_frames=24, grid=(5, 4), seed=2, outlier_fraction=0.05)
    print("rig", n_cams, rig_api.rig_solve(ctx, s.u, s.v, s.obs_cam, s.obs_point, s.frame_offsets, s.Xw, s.q_c, s.t_c,
                                           s.cam_frozen, s.q_f, s.t_f)[-1])
print("launches", ctx.kernel_launches())
ctx.close()
