"""GPU parity tests of the single-camera path: libccg.so (through the C-ABI) vs the
CPU oracle on identical inputs.  Tolerances are BASELINE.json's: relative 1e-6 on
recovered parameters, 1e-9 on the final RMS reprojection error."""
import numpy as np
import pytest

import oracle
from camera_calibrator_b200 import synth
from helpers import TOL_PARAM, TOL_RMS, block_err, quat_to_rot, rel_err, start_from_estimate

pytestmark = pytest.mark.gpu


def _check_solution(got, ref, poses=True):
    gi, gq, gt, grep = got
    ri, rq, rt, rrep = ref
    assert rel_err(gi, ri, floor=1e-300) <= TOL_PARAM, (gi, ri)
    assert abs(grep["rms"] - rrep["rms"]) <= TOL_RMS
    if poses:
        # rotations compared as matrices (q and -q are the same pose), entries are O(1)
        assert np.max(np.abs(quat_to_rot(gq) - quat_to_rot(rq))) <= TOL_PARAM
        assert rel_err(gt, rt, floor=1e-3) <= TOL_PARAM


@pytest.mark.parametrize("mode,npts", [(1, 54), (0, 54), (0, 150), (0, 400), (0, 33), (1, 150)])
def test_eval_blocks_match_oracle(ctx, mode, npts):
    """r, V, W, g_v, U, g_u, cost of one sweep at the Estimate() starting point.  Residuals, cost
    and gradients are FP64 in every sweep variant (1e-11).  The J'J blocks are FP64 (mode 1: 1e-11),
    or TF32 tensor-core products of float32 Jacobian rows in the default sweep (mode 0: operands
    truncated to 11 significant bits by design, 3e-3 of the block's largest entry).  The view sizes
    exercise partial 32-observation blocks."""
    d = oracle.datagen_planar(20, npts)
    intr, q, t = start_from_estimate(oracle, d)
    intr[4:] = [-0.03, 0.002, 0.001, -0.0005, 0.0007]  # exercise every Jacobian column
    g = ctx.intrinsics_eval(d.u, d.v, d.X, d.Y, d.Z, d.view_offsets, intr, q, t, hessian_fp64=mode)
    r = oracle.intrinsics_eval(d.u, d.v, d.X, d.Y, d.Z, d.view_offsets, intr, q, t)
    tol_h = {0: 3e-3, 1: 1e-11}[mode]
    assert np.max(np.abs(g["res"] - r["res"])) <= 1e-9 * max(1.0, np.max(np.abs(r["res"])))
    assert abs(g["cost"] - r["cost"]) <= 1e-12 * r["cost"]
    assert block_err(g["gu"], r["gu"]) <= 1e-11
    assert block_err(g["U"], r["U"]) <= tol_h
    for i in range(d.view_offsets.size - 1):
        assert block_err(g["gv"][i], r["gv"][i]) <= 1e-11, i
        for key in ("V", "W"):
            assert block_err(g[key][i], r[key][i]) <= tol_h, (key, i)


@pytest.mark.parametrize("mode", [0, 1])
def test_solve_c1_from_estimate(ctx, mode):
    """Config C1: 20 views x 54 points, Estimate() initialisation, all nine intrinsics free; both
    sweep variants (tensor-core = default, all-FP64) must land on the oracle's minimum."""
    import camera_calibrator_b200 as cc

    d = oracle.datagen_planar(20, 54)
    intr, q, t = start_from_estimate(oracle, d)
    got = ctx.intrinsics_solve(d.u, d.v, d.X, d.Y, d.Z, d.view_offsets, intr, q, t,
                               opts=cc.default_opts(hessian_fp64=mode))
    ref = oracle.intrinsics_solve(d.u, d.v, d.X, d.Y, d.Z, d.view_offsets, intr, q, t)
    assert got[3]["termination"] in (1, 2, 3)
    _check_solution(got, ref)
    # inputs untouched, counts preserved
    assert got[1].shape == (20, 4) and got[2].shape == (20, 3)


def test_solve_reference_test_scenario(ctx):
    """The reference's own test (test_calibrator.cpp:45-75): 5 views x 100 points, K within 1 %."""
    d = oracle.datagen_planar(5, 100)
    intr, q, t = start_from_estimate(oracle, d)
    gi, _, _, rep = ctx.intrinsics_solve(d.u, d.v, d.X, d.Y, d.Z, d.view_offsets, intr, q, t)
    K = oracle.FIXTURE_K
    for got, want in ((gi[0], K[0, 0]), (gi[1], K[1, 1]), (gi[2], K[0, 2]), (gi[3], K[1, 2])):
        assert (got - want) / want < 0.01
    ref = oracle.intrinsics_solve(d.u, d.v, d.X, d.Y, d.Z, d.view_offsets, intr, q, t)
    _check_solution((gi, _, _, rep), ref, poses=False)


def test_solve_frozen_k3(ctx):
    """Python caller's variant: ForceDistortionToConstant(4) freezes k3 (cam_calibration.py:308)."""
    d = oracle.datagen_planar(12, 54, seed=5)
    intr, q, t = start_from_estimate(oracle, d)
    mask = 1 << 8
    got = ctx.intrinsics_solve(d.u, d.v, d.X, d.Y, d.Z, d.view_offsets, intr, q, t, const_mask=mask)
    ref = oracle.intrinsics_solve(d.u, d.v, d.X, d.Y, d.Z, d.view_offsets, intr, q, t, const_mask=mask)
    assert got[0][8] == 0.0  # frozen coordinate untouched, bit-exact
    _check_solution(got, ref)


def test_solve_ragged_and_empty_views(ctx):
    """Views of different sizes, including an empty one and one smaller than a warp."""
    d = oracle.datagen_planar(10, 64, seed=2)
    intr, q, t = start_from_estimate(oracle, d)
    keep = [64, 33, 7, 64, 0, 50, 64, 1, 40, 64]
    idx = np.concatenate([np.arange(i * 64, i * 64 + k) for i, k in enumerate(keep)])
    off = np.concatenate([[0], np.cumsum(keep)]).astype(np.int64)
    u, v, X, Y, Z = (a[idx] for a in (d.u, d.v, d.X, d.Y, d.Z))
    g = ctx.intrinsics_eval(u, v, X, Y, Z, off, intr, q, t)
    r = oracle.intrinsics_eval(u, v, X, Y, Z, off, intr, q, t)
    assert g["res"].shape == r["res"].shape == (int(off[-1]), 2)
    assert abs(g["cost"] - r["cost"]) <= 1e-12 * r["cost"]
    assert block_err(g["gu"], r["gu"]) <= 1e-11
    assert block_err(g["U"], r["U"]) <= 3e-3  # default sweep: TF32 tensor-core products
    assert np.all(g["V"][4] == 0.0) and np.all(g["W"][4] == 0.0)  # empty view contributes nothing
    got = ctx.intrinsics_solve(u, v, X, Y, Z, off, intr, q, t)
    ref = oracle.intrinsics_solve(u, v, X, Y, Z, off, intr, q, t)
    # the empty / 1-point views have free gauge directions: compare intrinsics and RMS only
    _check_solution(got, ref, poses=False)


def test_solve_c2_perturbed_start(ctx):
    """Config C2 shape (2000 views x 187 points) from the perturbed ground-truth start."""
    d = synth.planar_views(2000, 187, seed=0)
    intr, q, t = synth.perturbed_start(d)
    got = ctx.intrinsics_solve(d.u, d.v, d.X, d.Y, d.Z, d.view_offsets, intr, q, t)
    ref = oracle.intrinsics_solve(d.u, d.v, d.X, d.Y, d.Z, d.view_offsets, intr, q, t)
    _check_solution(got, ref)


def test_solve_batch_matches_individual(ctx):
    """Config C4 shape, reduced: 24 independent problems (50 views x 54 points, seed = index)."""
    nb = 24
    ds = [oracle.datagen_planar(50, 54, seed=b) for b in range(nb)]
    starts = [start_from_estimate(oracle, d) for d in ds]
    u, v, X, Y, Z = (np.concatenate([getattr(d, k) for d in ds]) for k in "uvXYZ")
    off = np.arange(nb * 50 + 1, dtype=np.int64) * 54
    poff = np.arange(nb + 1, dtype=np.int64) * 50
    intr = np.stack([s[0] for s in starts])
    q = np.concatenate([s[1] for s in starts])
    t = np.concatenate([s[2] for s in starts])
    gi, gq, gt, greps = ctx.intrinsics_solve(u, v, X, Y, Z, off, intr, q, t, problem_offsets=poff)
    ri, rq, rt, rreps = oracle.intrinsics_solve_batch(u, v, X, Y, Z, off, poff, intr, q, t)
    for b in range(nb):
        sl = slice(b * 50, (b + 1) * 50)
        _check_solution((gi[b], gq[sl], gt[sl], greps[b]), (ri[b], rq[sl], rt[sl], rreps[b]))


def test_device_resident_run_is_repeatable(ctx):
    """Upload once, run twice: bit-identical results (fixed-order reductions)."""
    d = synth.planar_views(300, 100, seed=7)
    intr, q, t = synth.perturbed_start(d)
    p = ctx.intrinsics_problem(d.u, d.v, d.X, d.Y, d.Z, d.view_offsets, intr, q, t)
    p.run()
    a = p.download()
    p.run()
    b = p.download()
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    assert a[3][0]["final_cost"] == b[3][0]["final_cost"]
    p.close()


def test_bad_arguments_are_rejected(ctx):
    import camera_calibrator_b200 as cc

    d = oracle.datagen_planar(3, 10)
    intr, q, t = start_from_estimate(oracle, d)
    off = d.view_offsets.copy()
    off[1] = off[2] + 1  # not monotone
    with pytest.raises(cc.CcgError) as e:
        ctx.intrinsics_solve(d.u, d.v, d.X, d.Y, d.Z, off, intr, q, t)
    assert e.value.status == -1


@pytest.mark.parametrize("nv,npts,seed", [(20, 54, 0), (5, 100, 0), (200, 187, 3)])
def test_estimate_device_init_then_solve(ctx, nv, npts, seed):
    """ccg_intrinsics_estimate = Calibrator::Estimate (calibrator.cpp:47-68) with the closed-form
    initialisation on the device.  The device init works on Hartley-normalised normal equations
    (DESIGN.md section 4), so its K differs from the reference's raw-SVD K at the noise level
    (< 1 %); the LM solve that follows must land on the oracle's minimum (1e-6 / 1e-9), and the
    mirror ambiguity is resolved with the board in front of the camera."""
    d = oracle.datagen_planar(nv, npts, seed=seed)
    _, K_ref, _, _ = oracle.estimate_init(d.u, d.v, d.X, d.Y, d.Z, d.view_offsets)
    gi, gq, gt, grep, init = ctx.intrinsics_estimate(d.u, d.v, d.X, d.Y, d.Z, d.view_offsets)
    k_ref = np.array([K_ref[0, 0], K_ref[1, 1], K_ref[0, 2], K_ref[1, 2]], dtype=np.float64)
    assert np.max(np.abs(init[:4] - k_ref) / k_ref) <= 1e-2
    assert np.all(init[4:] == 0.0)
    ref = oracle.estimate(d.u, d.v, d.X, d.Y, d.Z, d.view_offsets)
    assert rel_err(gi, ref[0], floor=1e-300) <= TOL_PARAM
    assert abs(grep["rms"] - ref[3]["rms"]) <= TOL_RMS
    assert np.all(gt[:, 2] > 0.0)
    # the same views as a batch of two problems: identical intrinsics per problem
    if nv >= 20:
        half = nv // 2
        poff = np.array([0, half, nv], dtype=np.int64)
        bi, _, _, breps, _ = ctx.intrinsics_estimate(d.u, d.v, d.X, d.Y, d.Z, d.view_offsets, problem_offsets=poff)
        for b, (lo, hi) in enumerate(((0, half), (half, nv))):
            o0, o1 = d.view_offsets[lo], d.view_offsets[hi]
            sub = oracle.estimate(d.u[o0:o1], d.v[o0:o1], d.X[o0:o1], d.Y[o0:o1], d.Z[o0:o1],
                                  d.view_offsets[lo:hi + 1] - o0)
            assert rel_err(bi[b], sub[0], floor=1e-300) <= TOL_PARAM
            assert breps[b]["termination"] in (1, 2, 3)


def test_device_homography_known_answer(ctx):
    """The reference's own known-answer test for EstimateHomography (src/test_geometry.cpp:172-194: a random H with
    H(2,2) = 1 is recovered from 10 exact correspondences to 1e-4), run on the DEVICE kernel of the initialisation
    (k_init_homography behind ccg_homography_batch), 200 point sets in one call, against the oracle's SVD as well."""
    rng = np.random.default_rng(0)
    n_sets, n_pts = 200, 10
    Hs = rng.uniform(-1, 1, (n_sets, 3, 3)).astype(np.float32)
    Hs /= Hs[:, 2:3, 2:3]
    p1 = np.concatenate([rng.uniform(-1, 1, (n_sets, n_pts, 2)), np.ones((n_sets, n_pts, 1))], 2).astype(np.float32)
    p2 = np.einsum("sij,spj->spi", Hs.astype(np.float64), p1.astype(np.float64))
    keep = np.min(np.abs(p2[:, :, 2]), axis=1) > 0.2   # (sets whose points stay away from the line at infinity)
    p2 = (p2 / p2[:, :, 2:]).astype(np.float32)
    off = np.arange(n_sets + 1, dtype=np.int64) * n_pts
    got = ctx.homography_batch(p1[:, :, 0].ravel(), p1[:, :, 1].ravel(), p2[:, :, 0].ravel(), p2[:, :, 1].ravel(), off)
    assert keep.sum() >= 50
    for s in np.nonzero(keep)[0]:
        H2 = got[s] / got[s, 2, 2]
        assert np.max(np.abs(Hs[s] - H2)) <= 1e-4, s
        Ho = oracle.homography(p1[s, :, 0], p1[s, :, 1], p2[s, :, 0], p2[s, :, 1])
        assert np.max(np.abs(Ho / Ho[2, 2] - H2)) <= 1e-4
    # ragged and degenerate sets: 3 points -> NaN marker, the others unaffected
    off2 = np.array([0, 3, 13, 23], dtype=np.int64)
    x1 = np.concatenate([p1[0, :3, 0], p1[1, :, 0], p1[2, :, 0]])
    y1 = np.concatenate([p1[0, :3, 1], p1[1, :, 1], p1[2, :, 1]])
    x2 = np.concatenate([p2[0, :3, 0], p2[1, :, 0], p2[2, :, 0]])
    y2 = np.concatenate([p2[0, :3, 1], p2[1, :, 1], p2[2, :, 1]])
    g2 = ctx.homography_batch(x1, y1, x2, y2, off2)
    assert np.isnan(g2[0, 0, 0])
    assert np.array_equal(g2[1], got[1]) and np.array_equal(g2[2], got[2])


def test_estimate_needs_three_views(ctx):
    import camera_calibrator_b200 as cc

    d = oracle.datagen_planar(2, 54)
    with pytest.raises(cc.CcgError):
        ctx.intrinsics_estimate(d.u, d.v, d.X, d.Y, d.Z, d.view_offsets)


def test_lens_kernels_match_float64_model(ctx):
    """ccg_distort against the float32 model evaluated in numpy (bit-exact: same operations, no
    contraction) and ccg_undistort against the float64 root of the lens equations, on 1e6 points."""
    rng = np.random.default_rng(3)
    n = 1_000_000
    K = np.array([[1000, 0, 800], [0, 990, 500], [0, 0, 1]], np.float32)
    dist = np.array([-4e-2, 5e-4, 1e-3, 2e-5, -3e-4], np.float32)
    xn = rng.uniform(-0.7, 0.7, n).astype(np.float32)
    yn = rng.uniform(-0.45, 0.45, n).astype(np.float32)
    k1, k2, p1, p2, k3 = dist
    two = np.float32(2.0)
    r2 = xn * xn + yn * yn
    r4 = r2 * r2
    r6 = r4 * r2
    m = np.float32(1.0) + k1 * r2 + k2 * r4 + k3 * r6
    xd = xn * m + two * p1 * xn * yn + p2 * (r2 + two * xn * xn)
    yd = yn * m + two * p2 * xn * yn + p1 * (r2 + two * yn * yn)
    pu, pv = ctx.distort(xn, yn, K, dist)
    assert np.array_equal(pu, K[0, 0] * xd + K[0, 2]) and np.array_equal(pv, K[1, 1] * yd + K[1, 2])
    bx, by = ctx.undistort(pu, pv, K, dist)
    # pixels are float32 (ulp 6e-5 px at 1e3 px = 6e-8 normalised): the round trip is exact to a few float32 ulp
    assert np.max(np.abs(bx - xn)) < 1e-6 and np.max(np.abs(by - yn)) < 1e-6
    # and the returned point IS a root of the float64 model for the given float32 pixels
    x, y = bx.astype(np.float64), by.astype(np.float64)
    d = dist.astype(np.float64)
    r2 = x * x + y * y
    mm = 1 + d[0] * r2 + d[1] * r2 ** 2 + d[4] * r2 ** 3
    fx = (x * mm + 2 * d[2] * x * y + d[3] * (r2 + 2 * x * x)) * float(K[0, 0]) + float(K[0, 2])
    fy = (y * mm + 2 * d[3] * x * y + d[2] * (r2 + 2 * y * y)) * float(K[1, 1]) + float(K[1, 2])
    assert np.max(np.abs(fx - pu)) < 3e-4 and np.max(np.abs(fy - pv)) < 3e-4  # float32 rounding of the output (6e-8 x 1e3 px) and of the pixels


@pytest.mark.parametrize("name", ["reference_test_5x100", "c1_20x54"])
def test_estimate_matches_cv2_golden(ctx, name):
    """The committed golden vectors (tests/golden/intrinsics_golden.json): cv2.calibrateCamera - the
    function the reference's EstimateOpenCv calls (calibrator.cpp:38), same cost - run to convergence
    on the reference fixture's data.  The CUDA path, without the oracle's solver: 1e-6 on K, 2e-4 on
    the distortion (cv2 stops on its own criterion), 1e-9 on the RMS."""
    import json
    import os

    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "intrinsics_golden.json")))[name]
    d = oracle.datagen_planar(g["n_views"], g["n_points"])  # the generator only (pinned by the same file)
    for k in "uvXY":
        assert float(np.sum(getattr(d, k).astype(np.float64))) == g["checksum"][k]
    gi, _, _, rep, _ = ctx.intrinsics_estimate(d.u, d.v, d.X, d.Y, d.Z, d.view_offsets)
    cv = np.array(g["cv2"]["K"] + g["cv2"]["dist"])
    assert rep["termination"] in (1, 2, 3)
    assert np.max(np.abs(gi[:4] - cv[:4]) / cv[:4]) <= 1e-6
    assert np.max(np.abs(gi[4:] - cv[4:]) / np.abs(cv[4:])) <= 2e-4
    assert abs(rep["rms"] - g["cv2"]["rms"]) <= 1e-9
