"""Where does the end-to-end time of ccg_intrinsics_solve go (C5 geometry, 1 GPU)?"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import camera_calibrator_b200 as cc
from camera_calibrator_b200 import synth

nv = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
ctx = cc.Context(0)
d = synth.planar_views(nv, 400, seed=0)
intr, q, t = synth.perturbed_start(d)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
hu, hv, hX, hY, hZ = (pin(a) for a in (d.u, d.v, d.X, d.Y, d.Z))
for name, arrs in (("pageable", (d.u, d.v, d.X, d.Y, d.Z)), ("pinned", (hu, hv, hX, hY, hZ))):
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        p = ctx.intrinsics_problem(*arrs, d.view_offsets, intr, q, t)
        t1 = time.perf_counter()
        p.run()
        t2 = time.perf_counter()
        out = p.download()
        t3 = time.perf_counter()
        p.close()
        t4 = time.perf_counter()
        print(f"{name} rep {rep}: create {1e3*(t1-t0):.1f} ms  run {1e3*(t2-t1):.1f}  download {1e3*(t3-t2):.1f}  destroy {1e3*(t4-t3):.1f}  total {1e3*(t4-t0):.1f}")
    t0 = time.perf_counter()
    r = ctx.intrinsics_solve(*arrs, d.view_offsets, intr, q, t)
    print(f"{name} one-call solve: {1e3*(time.perf_counter()-t0):.1f} ms")

# the PCIe floor of this box: one 800 MB pinned host -> device copy, and the same as five 160 MB copies
x = torch.empty(200_000_000, dtype=torch.float32).pin_memory()
y = torch.empty(200_000_000, dtype=torch.float32, device="cuda")
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    y.copy_(x, non_blocking=True)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    for i in range(5):
        y[i * 40_000_000:(i + 1) * 40_000_000].copy_(x[i * 40_000_000:(i + 1) * 40_000_000], non_blocking=True)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"H2D 800 MB pinned: {1e3*(t1-t0):.2f} ms ({0.8/(t1-t0):.1f} GB/s); as 5 copies: {1e3*(t2-t1):.2f} ms")
