"""Single-GPU proxy of the strong-scaling curve of config C5: the device-resident solve (graph replay) of the
per-rank share of 100,000 views at N = 1, 2, 4, 8 ranks (the partitioned solve adds two 64-double allreduces per
iteration on top).  Usage: python tools/scaling_proxy.py [points]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import camera_calibrator_b200 as cc
from camera_calibrator_b200 import synth

npts = int(sys.argv[1]) if len(sys.argv) > 1 else 400
ctx = cc.Context(0)
base = None
for n in (1, 2, 4, 8):
    nv = 100000 // n
    d = synth.planar_views(nv, npts, seed=0)
    intr, q, t = synth.perturbed_start(d)
    p = ctx.intrinsics_problem(d.u, d.v, d.X, d.Y, d.Z, d.view_offsets, intr, q, t)
    p.run()
    ms = min(p.run()["total_ms"] for _ in range(5))
    rep = p.download()[3][0]
    base = base or ms
    print(f"ranks {n}: {nv} views, solve {ms:.3f} ms ({rep['sweeps']} sweeps), proxy speed-up {base / ms:.2f}x", flush=True)
    p.close()
