"""Time the sweep kernel alone (CUDA events inside libccg.so) on a C5-geometry problem.
Usage: python tools/time_sweep.py [views] [points]   (CCG_TC_WARPS selects the tc variant)"""
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
print(f"CCG_TC_WARPS={os.environ.get('CCG_TC_WARPS', 'default')} views {nv} x {npts}: sweep {ms:.4f} ms/launch, "
      f"{gb / (ms * 1e-3):.1f} GB/s algorithmic, {d.n_obs / (ms * 1e-3) / 1e9:.2f} Gobs/s")
