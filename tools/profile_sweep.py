"""Small driver for ncu: a quarter-size C5 problem (25,000 views x 400 points), one full
solve.  Run plain first, then under ncu (B200_PROFILING.md)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import camera_calibrator_b200 as cc
from camera_calibrator_b200 import synth

nv = int(sys.argv[1]) if len(sys.argv) > 1 else 25000
npts = int(sys.argv[2]) if len(sys.argv) > 2 else 400
ctx = cc.Context(0)
d = synth.planar_views(nv, npts, seed=0)
intr, q, t = synth.perturbed_start(d)
p = ctx.intrinsics_problem(d.u, d.v, d.X, d.Y, d.Z, d.view_offsets, intr, q, t)
tm = p.run(cc.default_opts(profile=1))
print(tm, p.download()[3])
print("sweep ms/launch", p.time_sweep(5))
