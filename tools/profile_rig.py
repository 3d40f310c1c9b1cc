"""Small driver for ncu: the C3 rig problem (4 cameras x 1000 frames x 187 points), one full solve.
Run plain first, then under ncu (B200_PROFILING.md)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import camera_calibrator_b200 as cc
from camera_calibrator_b200 import rig_api, synth

ctx = cc.Context(0)
s = synth.rig_scene(n_cams=4, n_frames=1000, grid=(17, 11), seed=0)
p = rig_api.RigProblem(ctx, s.u, s.v, s.obs_cam, s.obs_point, s.frame_offsets, s.Xw, s.q_c, s.t_c, s.cam_frozen,
                       s.q_f, s.t_f)
tm = p.run(cc.default_opts(max_iterations=1000, profile=1))
print(tm, p.download(with_cost=False)[5])
