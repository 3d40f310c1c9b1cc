import numpy as np, sys, os
sys.path.insert(0, os.getcwd())
import oracle, camera_calibrator_b200 as cc
from camera_calibrator_b200 import rig_api, synth
ctx = cc.Context(0)
def args(s): return (s.u, s.v, s.obs_cam, s.obs_point, s.frame_offsets, s.Xw)
for name, kw in (("2cam", dict(n_cams=2, n_frames=200, grid=(2, 2), pitch=0.2, seed=0)),
                 ("4cam", dict(n_cams=4, n_frames=120, grid=(7, 5), seed=2, outlier_fraction=0.05)),
                 ("12cam", dict(n_cams=12, n_frames=60, grid=(5, 4), seed=7, outlier_fraction=0.03))):
    s = synth.rig_scene(**kw)
    ref = oracle.rig_solve(*args(s), s.q_c, s.t_c, s.cam_frozen, s.q_f, s.t_f)
    for h in (0, 1):
        o = cc.default_opts(max_iterations=1000, hessian_fp64=h)
        got = rig_api.rig_solve(ctx, *args(s), s.q_c.copy(), s.t_c.copy(), s.cam_frozen, s.q_f.copy(), s.t_f.copy(), opts=o)
        dc = np.max(np.abs(got[4] - ref[4])) / np.max(ref[4])
        dt = np.max(np.abs(got[3] - ref[3]))
        print(name, "hessian_fp64", h, "cost rel diff %.3e" % dc, "t_f abs diff %.3e" % dt, "rms diff %.3e" % abs(got[5]["rms"] - ref[5]["rms"]),
              "it", got[5]["iterations"], ref[5]["iterations"], "term", got[5]["termination"], ref[5]["termination"], "gnorm %.2e %.2e" % (got[5]["gradient_norm"], ref[5]["gradient_norm"]))
