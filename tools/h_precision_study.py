"""Convergence of the exact-Schur LM when the J'J blocks come from a split low-precision contraction
(oracle experiment modes, oracle/lm.cpp eval_single): iterations, sweeps and the distance of the
result from the all-FP64 solve.  Evidence for the tensor-core sweep design."""
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
    for nv, npts, seed in ((50, 54, 3), (2000, 187, 0), (6000, 400, 0)):
        d = synth.planar_views(nv, npts, seed=seed)
        intr, q, t = synth.perturbed_start(d)
        yield f"{nv}x{npts} (perturbed start)", d, intr, q, t


for name, d, intr, q, t in cases():
    ref = None
    for mode, label in ((0, "fp64"), (1, "fp32"), (111, "2-pass 11 bit"), (108, "2-pass 8 bit (bf16x2)"),
                        (211, "1-pass 11 bit"), (311, "1-pass 11 bit truncated"),
                        (208, "1-pass 8 bit")):
        L.oracle_set_h_float(mode)
        try:
            got = oracle.intrinsics_solve(d.u, d.v, d.X, d.Y, d.Z, d.view_offsets, intr, q, t)
        except RuntimeError as e:
            print(f"{name:40s} {label:24s} FAILED {e}")
            continue
        if ref is None:
            ref = got
        rel = np.max(np.abs(got[0] - ref[0]) / np.abs(ref[0]))
        r = got[3]
        print(f"{name:40s} {label:24s} it {r['iterations']:3d} sweeps {r['sweeps']:3d} rej {r['rejected']:2d} "
              f"term {r['termination']} rms {r['rms']:.12f} d_rms {abs(r['rms'] - ref[3]['rms']):.1e} max rel d_intr {rel:.1e}")
    L.oracle_set_h_float(0)
