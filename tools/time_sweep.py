"""Time the sweep kernel alone (CUDA events inside libccg.so) and one full device-resident solve on a
C5-geometry problem.  Usage: python tools/time_sweep.py [views] [points]   (CCG_LIB selects a variant build)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import camera_calibrator_b200 as cc
from camera_calibrator_b200 import synth

nv = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
npts = int(sys.argv[2]) if len(sys.argv) > 2 else 400
ctx = cc.Context(0)
d = synth.planar_views(nv, npts, seed=0)
intr, q, t = synth.perturbed_start(d)
p = ctx.intrinsics_problem(d.u, d.v, d.X, d.Y, d.Z, d.view_offsets, intr, q, t)
ms = p.time_sweep(10)
gb = (20.0 * d.n_obs + 704.0 * d.n_views) / 1e9
lib = os.path.basename(os.environ.get("CCG_LIB", "libccg.so"))
print(f"{lib} views {nv} x {npts}: sweep {ms:.4f} ms/launch, "
      f"{gb / (ms * 1e-3):.1f} GB/s algorithmic, {d.n_obs / (ms * 1e-3) / 1e9:.2f} Gobs/s", flush=True)
p.run()
tm = p.run()
got = p.download()
print(f"   solve {tm['total_ms']:.3f} ms  intr {got[0][0][:4]} k {got[0][0][4:]} report {got[3][0]}", flush=True)
