"""Debug aid: GPU vs oracle on the C4-shaped problems, printing the per-sweep trace
of any problem whose intrinsics disagree by more than 1e-7."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import camera_calibrator_b200 as cc
import oracle
from helpers import start_from_estimate

np.set_printoptions(linewidth=200, precision=17)
ctx = cc.Context(0)
for b in range(int(sys.argv[1]) if len(sys.argv) > 1 else 24):
    d = oracle.datagen_planar(50, 54, seed=b)
    intr, q, t = start_from_estimate(oracle, d)
    p = ctx.intrinsics_problem(d.u, d.v, d.X, d.Y, d.Z, d.view_offsets, intr, q, t)
    p.run()
    gi, gq, gt, rep = p.download()
    ri, rq, rt, rrep = oracle.intrinsics_solve(d.u, d.v, d.X, d.Y, d.Z, d.view_offsets, intr, q, t)
    err = np.max(np.abs(gi[0] - ri) / np.abs(ri))
    print(b, "err %.2e" % err, rep[0]["iterations"], rep[0]["rejected"], rep[0]["termination"], "|",
          rrep["iterations"], rrep["rejected"], rrep["termination"],
          "gn %.2e %.2e" % (rep[0]["gradient_norm"], rrep["gradient_norm"]))
    if err > 1e-7:
        print(p.history(0))
        o = oracle.default_opts(verbose=1)
        oracle.intrinsics_solve(d.u, d.v, d.X, d.Y, d.Z, d.view_offsets, intr, q, t, opts=o)
    p.close()
