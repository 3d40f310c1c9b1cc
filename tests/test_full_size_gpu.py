"""BASELINE.json's full sizes on the GPU: oracle parity at the two sizes `north_star` targets (C5:
100,000 views x 400 points against the OpenMP oracle, C3: 4 cameras x 1,000 frames x 187 points with
Huber active), plus size-independent properties: recovery of the generator's ground truth, the
noise-level RMS, a vanishing gradient, additivity of the normal equations over a partition of the
views, run-to-run bit reproducibility, and oracle parity on a random sample of the C4 batch."""
import numpy as np
import pytest

import oracle
from camera_calibrator_b200 import rig_api, synth
from helpers import TOL_PARAM, TOL_RMS, quat_to_rot, rel_err

pytestmark = pytest.mark.gpu

TRUE_INTR = np.array([1000.0, 1000.0, 800.0, 500.0, -4e-2, 5e-4, 1e-3, 2e-5, -3e-4])
NOISE_RMS = np.sqrt(2.0 / 12.0)  # two coordinates, each U(-0.5, 0.5) px: sqrt(2 * 1/12)


def test_c5_full_size_properties(ctx):
    """configs[4]: one calibration of 100,000 views x 400 points (4e7 observations)."""
    d = synth.planar_views(100000, 400, seed=0)
    intr0, q0, t0 = synth.perturbed_start(d)
    p = ctx.intrinsics_problem(d.u, d.v, d.X, d.Y, d.Z, d.view_offsets, intr0, q0, t0)
    p.run()
    gi, gq, gt, reps = p.download()
    rep = reps[0]
    assert d.n_obs == 40_000_000 and gq.shape == (100000, 4) and gt.shape == (100000, 3)
    assert rep["termination"] in (1, 2, 3) and rep["iterations"] <= 12 and rep["rejected"] <= 2
    # 4e7 observations pin the intrinsics far below the test tolerances
    assert np.max(np.abs(gi[0][:4] - TRUE_INTR[:4]) / TRUE_INTR[:4]) <= 1e-5
    assert np.max(np.abs(gi[0][4:] - TRUE_INTR[4:])) <= 1e-4
    assert abs(rep["rms"] - NOISE_RMS) <= 3e-3
    assert rep["gradient_norm"] <= 1e-9
    # poses: rotation and translation of every view close to the generator's
    assert np.max(np.abs(gt - d.t_true)) <= 5e-3
    # a second run of the resident problem is bit-identical (fixed-order reductions)
    p.run()
    hi, hq, ht, hreps = p.download()
    assert np.array_equal(gi, hi) and np.array_equal(gq, hq) and np.array_equal(gt, ht)
    assert hreps[0]["final_cost"] == rep["final_cost"]
    p.close()


def test_normal_equations_are_additive_over_a_view_partition(ctx):
    """U, g_u and the cost of 20,000 views equal the sums over two halves (what the multi-GPU
    partition relies on); per-view blocks do not depend on which partition a view is in."""
    d = synth.planar_views(20000, 400, seed=4)
    intr, q, t = synth.perturbed_start(d)
    whole = ctx.intrinsics_eval(d.u, d.v, d.X, d.Y, d.Z, d.view_offsets, intr, q, t)
    parts = []
    for lo, hi in ((0, 9000), (9000, 20000)):
        o0, o1 = int(d.view_offsets[lo]), int(d.view_offsets[hi])
        parts.append(ctx.intrinsics_eval(d.u[o0:o1], d.v[o0:o1], d.X[o0:o1], d.Y[o0:o1], d.Z[o0:o1],
                                         d.view_offsets[lo:hi + 1] - o0, intr, q[lo:hi], t[lo:hi]))
    for key in ("U", "gu"):
        tot = parts[0][key] + parts[1][key]
        assert np.max(np.abs(tot - whole[key])) <= 1e-10 * np.max(np.abs(whole[key])), key
    assert abs(parts[0]["cost"] + parts[1]["cost"] - whole["cost"]) <= 1e-12 * whole["cost"]
    assert np.array_equal(np.concatenate([parts[0]["V"], parts[1]["V"]]), whole["V"])
    assert np.array_equal(np.concatenate([parts[0]["W"], parts[1]["W"]]), whole["W"])
    assert np.max(np.abs(np.concatenate([parts[0]["gv"], parts[1]["gv"]]) - whole["gv"])) <= 1e-9 * np.max(np.abs(whole["gv"]))


def test_c4_full_batch_properties(ctx):
    """configs[3]: 4,096 independent calibrations of 50 views x 54 points in one call."""
    nb, V, P = 4096, 50, 54
    ds = [synth.planar_views(V, P, seed=b) for b in range(nb)]
    st = [synth.perturbed_start(x, seed=1000 + b) for b, x in enumerate(ds)]
    u, v, X, Y, Z = (np.concatenate([getattr(x, k) for x in ds]) for k in "uvXYZ")
    off = np.arange(nb * V + 1, dtype=np.int64) * P
    poff = np.arange(nb + 1, dtype=np.int64) * V
    intr = np.stack([s[0] for s in st])
    q = np.concatenate([s[1] for s in st])
    t = np.concatenate([s[2] for s in st])
    gi, gq, gt, reps = ctx.intrinsics_solve(u, v, X, Y, Z, off, intr, q, t, problem_offsets=poff)
    assert gi.shape == (nb, 9) and len(reps) == nb
    assert all(r["termination"] in (1, 2, 3) for r in reps)
    rms = np.array([r["rms"] for r in reps])
    assert np.all(np.abs(rms - NOISE_RMS) <= 0.05)  # 2,700 observations each
    assert np.max(np.abs(gi[:, :2] - 1000.0)) <= 2.0 and np.max(np.abs(gi[:, 2] - 800.0)) <= 2.0  # sigma ~0.11 px
    assert np.max(np.abs(gi[:, 3] - 500.0)) <= 2.0
    # oracle parity on a random sample of the batch, each problem solved on its own
    for b in np.random.default_rng(0).choice(nb, 6, replace=False):
        x = ds[b]
        ref = oracle.intrinsics_solve(x.u, x.v, x.X, x.Y, x.Z, x.view_offsets, st[b][0], st[b][1], st[b][2])
        assert rel_err(gi[b], ref[0], floor=1e-300) <= TOL_PARAM, b
        assert abs(reps[b]["rms"] - ref[3]["rms"]) <= TOL_RMS, b


def test_c5_full_size_matches_oracle(ctx):
    """configs[4] at full size against the oracle (reference src/calibrator.cpp:221-336 restated; all host
    threads, about half a minute): intrinsics to 1e-6 relative, final RMS to 1e-9, every pose to 1e-6."""
    d = synth.planar_views(100000, 400, seed=0)
    intr0, q0, t0 = synth.perturbed_start(d)
    got = ctx.intrinsics_solve(d.u, d.v, d.X, d.Y, d.Z, d.view_offsets, intr0, q0, t0)
    ref = oracle.intrinsics_solve(d.u, d.v, d.X, d.Y, d.Z, d.view_offsets, intr0, q0, t0)
    gi, gq, gt, grep = got
    ri, rq, rt, rrep = ref
    assert grep["termination"] in (1, 2, 3) and rrep["termination"] in (1, 2, 3)
    assert rel_err(gi, ri, floor=1e-300) <= TOL_PARAM, (gi, ri)
    assert abs(grep["rms"] - rrep["rms"]) <= TOL_RMS
    assert np.max(np.abs(quat_to_rot(gq) - quat_to_rot(rq))) <= TOL_PARAM
    assert rel_err(gt, rt, floor=1e-3) <= TOL_PARAM
    assert grep["sweeps"] == rrep["sweeps"]  # the same LM path, not only the same minimum


def test_c3_full_size_matches_oracle(ctx):
    """configs[2] at full size: 4 cameras x 1,000 frames x 187 points = 748,000 observations, 5 % gross
    outliers so the Huber loss is active (reference src/extrinsics_calibrator.cpp:86-257 restated)."""
    s = synth.rig_scene(n_cams=4, n_frames=1000, grid=(17, 11), seed=0, outlier_fraction=0.05)
    assert s.u.size == 4 * 1000 * 187
    args = (s.u, s.v, s.obs_cam, s.obs_point, s.frame_offsets, s.Xw)
    got = rig_api.rig_solve(ctx, *args, s.q_c, s.t_c, s.cam_frozen, s.q_f, s.t_f)
    ref = oracle.rig_solve(*args, s.q_c, s.t_c, s.cam_frozen, s.q_f, s.t_f)
    gqc, gtc, gqf, gtf, gcost, grep = got
    rqc, rtc, rqf, rtf, rcost, rrep = ref
    assert grep["termination"] in (1, 2, 3)
    assert np.max(np.abs(quat_to_rot(gqc) - quat_to_rot(rqc))) <= TOL_PARAM
    assert np.max(np.abs(quat_to_rot(gqf) - quat_to_rot(rqf))) <= TOL_PARAM
    assert rel_err(gtc[1:], rtc[1:], floor=1e-3) <= TOL_PARAM
    assert rel_err(gtf, rtf, floor=1e-3) <= TOL_PARAM
    assert abs(grep["rms"] - rrep["rms"]) <= TOL_RMS
    assert np.array_equal(gqc[0], s.q_c[0]) and np.array_equal(gtc[0], s.t_c[0])  # frozen camera, bit-exact
    assert gcost.shape == rcost.shape == (748000,)
    assert np.max(np.abs(gcost - rcost)) <= 1e-9 * np.max(rcost)
    huber = float(np.float32(3.0) / np.float32(500.0))
    assert np.count_nonzero(rcost > 0.5 * huber * huber) > 10000  # the linear zone of the loss is populated
