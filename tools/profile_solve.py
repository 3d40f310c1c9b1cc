"""Driver for ncu / timing of whole device-resident solves (default options: the persistent LM kernel for small
problems, the multi-launch loop otherwise).  Usage: python tools/profile_solve.py [views] [points] [repeats]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import camera_calibrator_b200 as cc
from camera_calibrator_b200 import synth

nv = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
npts = int(sys.argv[2]) if len(sys.argv) > 2 else 187
rep = int(sys.argv[3]) if len(sys.argv) > 3 else 3
ctx = cc.Context(0)
d = synth.planar_views(nv, npts, seed=0)
intr, q, t = synth.perturbed_start(d)
p = ctx.intrinsics_problem(d.u, d.v, d.X, d.Y, d.Z, d.view_offsets, intr, q, t)
for _ in range(rep):
    tm = p.run()
print(f"{nv} x {npts}: solve {tm['total_ms']:.4f} ms", p.download()[3][0])
