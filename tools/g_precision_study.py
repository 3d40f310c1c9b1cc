"""Where does the exact-Schur LM land if the GRADIENT g = J'r is taken from a low-precision Jacobian (oracle
experiment switch g_g_float, oracle/lm.cpp eval_single) while the cost stays FP64?  The next step of the sweep kernel
 would halve its FP64 work that way; this is the precision evidence for or against it.
Columns: iterations, sweeps, RMS, |RMS - RMS of the all-FP64 solve|, max relative difference of the intrinsics."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C

import numpy as np

import oracle
from camera_calibrator_b200 import synth

oracle.build()
L = oracle.lib()
L.oracle_set_h_float.argtypes = [C.c_int]
L.oracle_set_g_float.argtypes = [C.c_int]


def cases():
    d = oracle.datagen_planar(20, 54)
    _, K, q, t = oracle.estimate_init(d.u, d.v, d.X, d.Y, d.Z, d.view_offsets)
    intr = np.zeros(9)
    intr[:4] = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    yield "C1 (Estimate init)", d, intr, q, t
    d = oracle.datagen_planar(5, 100)
    _, K, q, t = oracle.estimate_init(d.u, d.v, d.X, d.Y, d.Z, d.view_offsets)
    intr = np.zeros(9)
    intr[:4] = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    yield "test_calibrator 5x100 (Estimate init)", d, intr, q, t
    sizes = ((50, 54, 3), (2000, 187, 0)) + (((6000, 400, 0),) if "--big" in sys.argv else ())
    for nv, npts, seed in sizes:
        d = synth.planar_views(nv, npts, seed=seed)
        intr, q, t = synth.perturbed_start(d)
        yield f"{nv}x{npts} (perturbed start)", d, intr, q, t


MODES = ((0, 0, "all FP64"),
         (311, 0, "J'J 11 bit, g FP64 (the shipped sweep)"),
         (311, 1, "J'J 11 bit, g = J32' r64, FP64 sums"),
         (311, 2, "J'J 11 bit, g = J32' r32, FP32 sums"),
         (311, 3, "J'J 11 bit, g = J11' r11, FP32 sums"))
for name, d, intr, q, t in cases():
    ref = None
    for h, g, label in MODES:
        L.oracle_set_h_float(h)
        L.oracle_set_g_float(g)
        try:
            got = oracle.intrinsics_solve(d.u, d.v, d.X, d.Y, d.Z, d.view_offsets, intr, q, t)
        except RuntimeError as e:
            print(f"{name:38s} {label:40s} FAILED {e}")
            continue
        if ref is None:
            ref = got
        rel = np.max(np.abs(got[0] - ref[0]) / np.abs(ref[0]))
        r = got[3]
        print(f"{name:38s} {label:40s} it {r['iterations']:3d} sweeps {r['sweeps']:3d} rej {r['rejected']:2d} "
              f"term {r['termination']} rms {r['rms']:.12f} d_rms {abs(r['rms'] - ref[3]['rms']):.1e} max rel d_intr {rel:.1e}",
              flush=True)
    L.oracle_set_h_float(0)
    L.oracle_set_g_float(0)
