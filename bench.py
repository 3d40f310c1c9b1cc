#!/usr/bin/env python
"""Benchmark of the LM reprojection-error hot path (BASELINE.json metric:
residual+Jacobian observations/s; LM iterations/s; full calibration wall time).

  python bench.py --gpus N --steps K --warmup W [--workload c5|c2|c4|c1] [--impl reference]

One STEP = one full device-resident Levenberg-Marquardt solve (reset to the
initial parameters, then residual+Jacobian+normal-equation sweeps, pose
elimination, reduced solves and accept/reject until convergence) of the
workload.  `value` = observations x sweeps / device time, inputs resident in
HBM; `e2e` = the same through the C-ABI call with host buffers (H2D + solve +
D2H inside the timed region).  Default workload: BASELINE.json configs[4]
geometry on ONE GPU — a single calibration of 100,000 views x 400 points
(4e7 observations, 800 MB of observations: larger than the 126 MB L2, so every
sweep streams from HBM; configs[1] is 7.5 MB and L2-resident).
With --gpus N > 1 the same 100,000 views are partitioned by view over the N
GPUs of the box (strong scaling) and every LM iteration all-reduces the packed
reduced system (two 64-double sums over NVLink peer memory, fused into the
accept/reject and reduced-solve kernels) inside libccg.so; rank 0 also solves the
un-partitioned problem and the line carries `partition_check`.  --workload c3 is
the rig (frames partitioned over the ranks), c4 the batch (problems sharded,
no communication).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (views, points per view, problems)
    "c1": (20, 54, 1),
    "c2": (2000, 187, 1),
    "c4": (50, 54, 4096),
    "c5": (100000, 400, 1),
}
# per observation per sweep of the default (warp-specialised tensor-core) sweep kernel, from the SASS of the
# full-chunk loops of its two roles (profiles/r02_sass_opcodes.md, tools/sass_loops.py): FP64 role 54 DFMA + 19 DMUL
# + 13 DADD, Jacobian role 32 FFMA + 43 FMUL + 15 FADD, and 8 HMMA.16816.F32 (m16n8k16, FP16 operands, FP32
# accumulation) per 32 observations
FP64_FLOP_PER_OBS = 140.0
FP32_FLOP_PER_OBS = 122.0
MMA_FLOP_PER_OBS = 8 * 2.0 * 16 * 8 * 16 / 32.0


def algorithmic_bytes(n_obs, n_views):
    """20 B/observation (u,v,X,Y,Z float32) + per view 56 B pose read and
    648 B of pose blocks (V 21, W 54, g 6 doubles) written."""
    return 20.0 * n_obs + 704.0 * n_views


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(np.max(smax)) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def pinned(a):
    """Copy a numpy array into page-locked host memory (torch is plumbing only)."""
    import torch

    t = torch.from_numpy(np.ascontiguousarray(a))
    try:
        return t.pin_memory().numpy()
    except RuntimeError:
        return t.numpy()


def make_workload(name, rank, world, scaling):
    from camera_calibrator_b200 import synth

    V, P, B = WORKLOADS[name]
    if B == 1:
        if scaling == "weak":
            total = V * world
        else:
            total = V
        a, b = total * rank // world, total * (rank + 1) // world
        d = synth.planar_views(total, P, seed=0, view_range=(a, b))
        intr, q, t = synth.perturbed_start(d, n_views_total=total, view_start=a)
        return dict(d=d, intr=intr, q=q, t=t, poff=None, shared=world > 1, total_views=total,
                    total_obs=total * P, n_problems=1)
    # batch of independent problems: shard the problems over the ranks, no communication
    b0, b1 = B * rank // world, B * (rank + 1) // world
    ds, intrs, qs, ts = [], [], [], []
    for b in range(b0, b1):
        d = synth.planar_views(V, P, seed=b)
        i0, q0, t0 = synth.perturbed_start(d, seed=1000 + b)
        ds.append(d), intrs.append(i0), qs.append(q0), ts.append(t0)
    cat = lambda k: np.concatenate([getattr(x, k) for x in ds])
    nb = b1 - b0
    d = synth.PlanarViews(cat("u"), cat("v"), cat("X"), cat("Y"), cat("Z"),
                          np.arange(nb * V + 1, dtype=np.int64) * P, cat("q_true"), cat("t_true"))
    return dict(d=d, intr=np.stack(intrs), q=np.concatenate(qs), t=np.concatenate(ts),
                poff=np.arange(nb + 1, dtype=np.int64) * V, shared=False, total_views=B * V,
                total_obs=B * V * P, n_problems=nb)


# ------------------------------------------------------------------ CPU arms
def cpu_sample(name):
    """Bounded sample of the workload for the CPU legs (about 10-30 s of host work)."""
    from camera_calibrator_b200 import synth

    V, P, B = WORKLOADS[name]
    if B == 1:
        nv = min(V, 6000)
        d = synth.planar_views(V, P, seed=0, view_range=(0, nv))
        intr, q, t = synth.perturbed_start(d, n_views_total=V, view_start=0)
        return d, intr, q, t, None, f"first {nv} of {V} views x {P} points of the same data set ({nv * P} observations), full LM solve"
    nb = 24
    ds = [synth.planar_views(V, P, seed=b) for b in range(nb)]
    st = [synth.perturbed_start(x, seed=1000 + b) for b, x in enumerate(ds)]
    cat = lambda k: np.concatenate([getattr(x, k) for x in ds])
    d = synth.PlanarViews(cat("u"), cat("v"), cat("X"), cat("Y"), cat("Z"),
                          np.arange(nb * V + 1, dtype=np.int64) * P, cat("q_true"), cat("t_true"))
    return (d, np.stack([s[0] for s in st]), np.concatenate([s[1] for s in st]),
            np.concatenate([s[2] for s in st]), np.arange(nb + 1, dtype=np.int64) * V,
            f"first {nb} of {B} problems ({nb * V * P} observations), full LM solves")


def cpu_solve_once(sample):
    import oracle

    d, intr, q, t, poff, _ = sample
    t0 = time.perf_counter()
    if poff is None:
        _, _, _, rep = oracle.intrinsics_solve(d.u, d.v, d.X, d.Y, d.Z, d.view_offsets, intr, q, t)
        sweeps_obs = d.n_obs * rep["sweeps"]
        iters = rep["iterations"]
    else:
        _, _, _, reps = oracle.intrinsics_solve_batch(d.u, d.v, d.X, d.Y, d.Z, d.view_offsets, poff, intr, q, t)
        per = d.n_obs // len(reps)
        sweeps_obs = sum(per * r["sweeps"] for r in reps)
        iters = sum(r["iterations"] for r in reps)
    dt = time.perf_counter() - t0
    return sweeps_obs / dt, dt, iters


def run_reference(args):
    """--impl reference: the restated reference algorithm (oracle port; the reference itself
    cannot be compiled here - Ceres/Eigen/OpenCV absent) on the host cores.  Every timed step is a
    full LM solve of a BOUNDED SAMPLE of the workload (config.sample says which); the thread count is
    set explicitly (torchrun exports OMP_NUM_THREADS=1), and the line also carries the one-thread value
    (what the reference does: Ceres num_threads is left at 1, calibrator.cpp:314-321) and one
    full-size pass."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle

    oracle.build()
    cores = os.cpu_count() or 1
    oracle.set_num_threads(cores)
    sample = cpu_sample(args.workload)
    for _ in range(args.warmup):
        cpu_solve_once(sample)
    vals, times = [], []
    for _ in range(args.steps):
        v, dt, _ = cpu_solve_once(sample)
        vals.append(v), times.append(dt)
    value = float(np.sum([v * t for v, t in zip(vals, times)]) / np.sum(times))
    oracle.set_num_threads(1)
    v1, dt1, _ = cpu_solve_once(sample)
    oracle.set_num_threads(cores)
    full = None
    V, P, B = WORKLOADS[args.workload]
    if args.workload == "c5":
        # one pass over the whole workload (about half a minute on 16 threads): the sample is representative
        from camera_calibrator_b200 import synth

        d = synth.planar_views(V, P, seed=0)
        intr, q, t = synth.perturbed_start(d)
        vf, dtf, _ = cpu_solve_once((d, intr, q, t, None, ""))
        full = {"value": vf, "unit": "obs/s", "seconds": dtf, "views": V}
    cfg = workload_config(args.workload, args.gpus, args.scaling)
    cfg["sample"] = sample[5]
    line = {
        "impl": "reference", "metric": "residual+Jacobian observations/s (full LM solve)", "value": value,
        "unit": "obs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": float(np.mean(times) * 1e3), "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": "obs/s", "cores": cores, "kind": "port", "sample": sample[5]},
        "cpu_baseline_1thread": {"value": v1, "unit": "obs/s", "cores": 1, "kind": "port", "sample": sample[5],
                                 "seconds": dt1},
        "full_size_pass": full,
        "e2e": {"value": value, "unit": "obs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(name, gpus, scaling):
    V, P, B = WORKLOADS[name]
    desc = {
        "c1": "BASELINE configs[0]: single-camera calibration, 20 views x 54 points",
        "c2": "BASELINE configs[1]: single-camera intrinsics+distortion, 2000 views x 187 points (L2-resident)",
        "c4": "BASELINE configs[3]: 4096 independent calibrations (50 views x 54 points), sharded by problem",
        "c5": "BASELINE configs[4] geometry: single calibration, 100000 views x 400 points (4e7 observations), "
              "view-partitioned over the GPUs with an NCCL allreduce of the reduced system per LM iteration",
    }[name]
    return {"workload": desc, "views": V, "points_per_view": P, "problems": B,
            "parallelism": f"view-partition x{gpus}" if B == 1 else f"problem-shard x{gpus}",
            "start": "ground-truth poses perturbed +-0.02 / +-1 deg, K off by +-2 % / +-10 px, zero distortion",
            "cache": "inputs larger than L2 (no flush needed)" if name in ("c5", "c4") else
                     "inputs smaller than L2: L2-resident after the first sweep (roofline not meaningful)"}



# ------------------------------------------------------------------ rig (C3) arm
def run_rig(args):
    """--workload c3: BASELINE configs[2], ExtrinsicsCalibrator::Optimize on 4 cameras x 1000 frames x
    187 shared board points (748,000 observations, L2-resident).  With --gpus N > 1 the frames are
    partitioned over the ranks (camera extrinsics replicated, reduced system all-reduced per iteration)."""
    import torch
    import torch.distributed as dist

    import camera_calibrator_b200 as cc
    from camera_calibrator_b200 import rig_api, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    ctx = cc.Context(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.tensor(list(cc.Context.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        ctx.comm_init(rank, world, bytes(uid.cpu().tolist()))

    def barrier():
        if world > 1:
            dist.barrier()
            ctx.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if world == 1:
            return x
        tns = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(tns, op=dist.ReduceOp.MAX)
        return float(tns.item())

    F, C, P = 1000, 4, 187
    s = synth.rig_scene(n_cams=C, n_frames=F, grid=(17, 11), seed=0)
    fa, fb = F * rank // world, F * (rank + 1) // world
    o0, o1 = int(s.frame_offsets[fa]), int(s.frame_offsets[fb])
    a = (s.u[o0:o1], s.v[o0:o1], s.obs_cam[o0:o1], s.obs_point[o0:o1], s.frame_offsets[fa:fb + 1] - o0, s.Xw)
    prob = rig_api.RigProblem(ctx, *a, s.q_c, s.t_c, s.cam_frozen, s.q_f[fa:fb], s.t_f[fa:fb], shared=world > 1)
    opts = cc.default_opts(max_iterations=1000)
    for _ in range(max(args.warmup, 3)):
        prob.run(opts)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    l0 = ctx.kernel_launches()
    tot = 0.0
    nar = 0
    for _ in range(args.steps):
        tm = prob.run(opts)
        tot += tm["total_ms"]
        nar += tm["allreduces"]
    barrier()
    launches = ctx.kernel_launches() - l0
    clocks = sampler.stop() if rank == 0 else None
    gqc, gtc, _, _, _, rep = prob.download(with_cost=False)
    ms = allmax(tot) / args.steps
    obs_sweeps = s.n_obs * rep["sweeps"]
    e2e_t = None
    partition_check = None
    if world == 1:
        # end to end: host arrays in, poses + per-observation costs out
        pa = tuple(pinned(x) for x in a)
        times = []
        for i in range(1 + min(args.steps, 3)):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            rig_api.rig_solve(ctx, *pa, s.q_c, s.t_c, s.cam_frozen, s.q_f, s.t_f, opts=opts)
            if i:
                times.append(time.perf_counter() - t0)
        e2e_t = float(np.mean(times))
    else:
        mine = torch.tensor(np.concatenate([gqc.ravel(), gtc.ravel(), [rep["rms"]]]), device="cuda")
        allv = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allv, mine)
        if rank == 0:
            ctx1 = cc.Context(local)
            full = (s.u, s.v, s.obs_cam, s.obs_point, s.frame_offsets, s.Xw)
            rqc, rtc, _, _, _, rrep = rig_api.rig_solve(ctx1, *full, s.q_c, s.t_c, s.cam_frozen, s.q_f, s.t_f, opts=opts)
            ctx1.close()
            partition_check = {"t_c_abs_err": float(np.max(np.abs(gtc - rtc))), "q_c_abs_err": float(np.max(np.abs(gqc - rqc))),
                               "rms_abs_err": float(abs(rep["rms"] - rrep["rms"])),
                               "replicated_bit_identical": bool(all(torch.equal(allv[0], x) for x in allv)),
                               "sweeps": [int(rep["sweeps"]), int(rrep["sweeps"])],
                               "reference": "the same rig solved un-partitioned on rank 0's GPU"}
        barrier()
    line = None
    if rank == 0:
        import oracle

        oracle.build()
        oracle.set_num_threads(os.cpu_count() or 1)
        full = (s.u, s.v, s.obs_cam, s.obs_point, s.frame_offsets, s.Xw)
        t0 = time.perf_counter()
        ref = oracle.rig_solve(*full, s.q_c, s.t_c, s.cam_frozen, s.q_f, s.t_f)
        cpu_dt = time.perf_counter() - t0
        alg = 12.0 * s.n_obs + 12.0 * F * P + (56 + 8 * (21 + 36 * (C - 1) + 6)) * F
        line = {
            "metric": "residual+Jacobian observations/s (full LM solve)", "value": obs_sweeps / (ms * 1e-3), "unit": "obs/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "BASELINE configs[2]: rig extrinsics (ExtrinsicsCalibrator::Optimize), 4 cameras x 1000 frames "
                                   "x 187 shared board points, Huber loss", "cameras": C, "frames": F, "points_per_frame": P,
                       "parallelism": f"frame-partition x{world}",
                       "cache": "inputs smaller than L2: L2-resident after the first sweep (roofline not meaningful)"},
            "lm_iterations_per_s": rep["iterations"] / (ms * 1e-3),
            "solve": {k: rep[k] for k in ("iterations", "sweeps", "rejected", "termination", "rms", "gradient_norm")},
            "allreduces_per_step": nar / args.steps,
            "roofline": {"kernel": "k_rig_sweep", "bound": "hbm", "achieved": None, "peak": measured_peaks()[0], "unit": "GB/s",
                         "frac": None, "traffic": None, "algorithmic_bytes_per_launch": alg,
                         "note": "9 MB working set is L2-resident: no HBM fraction quoted (SURVEY 8d)"},
            "cpu_baseline": {"value": s.n_obs * ref[5]["sweeps"] / cpu_dt, "unit": "obs/s", "cores": oracle.num_threads(),
                             "kind": "port", "sample": "the full C3 problem, full LM solve", "seconds": cpu_dt},
            "e2e": None if e2e_t is None else {
                "value": obs_sweeps / e2e_t, "unit": "obs/s", "h2d_bytes_per_step": int(sum(x.nbytes for x in a)),
                "d2h_bytes_per_step": int(8 * s.n_obs + 56 * (F + C)), "ms_per_step": e2e_t * 1e3},
            "partition_check": partition_check,
            "gpu_launches": int(launches), "clocks": clocks,
        }
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    if line is not None:
        print(json.dumps(line), flush=True)
    prob.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    import camera_calibrator_b200 as cc

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # libraries (NCCL's version banner) must not write into the one-JSON-line stdout contract
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local)
    ctx = cc.Context(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.tensor(list(cc.Context.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        ctx.comm_init(rank, world, bytes(uid.cpu().tolist()))

    def barrier():
        if world > 1:
            dist.barrier()
            ctx.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if world == 1:
            return x
        tns = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(tns, op=dist.ReduceOp.MAX)
        return float(tns.item())

    def allsum(x):
        if world == 1:
            return x
        tns = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(tns, op=dist.ReduceOp.SUM)
        return float(tns.item())

    wl = make_workload(args.workload, rank, world, args.scaling)
    d = wl["d"]
    # ---- what a first Calibrator::Optimize call sees: pageable host arrays, cold memory pool (no warm-up, timed once)
    e2e_cold_ms = None
    if world == 1:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.intrinsics_solve(d.u, d.v, d.X, d.Y, d.Z, d.view_offsets, wl["intr"], wl["q"], wl["t"],
                             problem_offsets=wl["poff"])
        e2e_cold_ms = (time.perf_counter() - t0) * 1e3
    opts = cc.default_opts()            # timed steps: no per-kernel events, LM iterations replayed as a CUDA graph
    opts_prof = cc.default_opts(profile=1)  # breakdown pass: CUDA events around every kernel class
    prob = ctx.intrinsics_problem(d.u, d.v, d.X, d.Y, d.Z, d.view_offsets, wl["intr"], wl["q"], wl["t"],
                                  problem_offsets=wl["poff"], shared=wl["shared"])
    fp64_peak = ctx.measure_fp64_peak() if rank == 0 else 0.0

    # ---- device-resident solves: W warm-up, K timed
    for _ in range(args.warmup):
        prob.run(opts)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # before the barrier: spawning the sampler must not skew rank 0 against the others
    barrier()
    launches0 = ctx.kernel_launches()
    tot_ms = 0.0
    n_allreduce = 0
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        tm = prob.run(opts)
        tot_ms += tm["total_ms"]
        n_allreduce += tm["allreduces"]
    barrier()
    wall_ms = (time.perf_counter() - t_wall0) * 1e3
    launches = ctx.kernel_launches() - launches0
    clocks = sampler.stop() if rank == 0 else None
    # ---- breakdown pass (not part of `value`): per-kernel-class device times of the same solve
    n_prof = 2
    sweep_ms = schur_ms = small_ms = 0.0
    n_sweeps = 0
    prob.run(opts_prof)  # untimed: small problems ran the persistent kernel so far, the multi-launch kernels load here
    for _ in range(n_prof):
        tm = prob.run(opts_prof)
        sweep_ms += tm["sweep_ms"]
        schur_ms += tm["schur_ms"]
        small_ms += tm["small_ms"]
        n_sweeps += tm["sweep_launches"]
    barrier()
    intr, q, t, reps = prob.download()
    # observations swept per step over the whole job (every problem: its obs x its sweeps)
    if wl["poff"] is None:
        obs_sweeps_local = d.n_obs * reps[0]["sweeps"]
    else:
        per = d.n_obs // max(len(reps), 1)
        obs_sweeps_local = sum(per * r["sweeps"] for r in reps)
    obs_sweeps = allsum(float(obs_sweeps_local))
    tot_ms = allmax(tot_ms)
    ms_per_step = tot_ms / args.steps
    value = obs_sweeps / (ms_per_step * 1e-3)
    iters = reps[0]["iterations"] if wl["poff"] is None else allsum(float(sum(r["iterations"] for r in reps)))

    # ---- the all-FP64 sweep variant (ccg_opts.hessian_fp64 = 1: J'J in FP64 with warp shuffles, no tensor cores -
    # north_star's literal formulation) on the same problem, printed beside the default
    fp64_variant = None
    if world == 1:
        o64 = cc.default_opts(hessian_fp64=1)
        prob.run(o64)
        t64 = [prob.run(o64)["total_ms"] for _ in range(2)]
        tm64 = prob.run(cc.default_opts(hessian_fp64=1, profile=1))
        r64 = prob.download()[3][0]
        sw64 = tm64["sweep_ms"] / max(tm64["sweep_launches"], 1)
        fp64_variant = {"ms_per_step": float(np.mean(t64)), "sweeps": r64["sweeps"], "rms_px": r64["rms"],
                        "value": d.n_obs * r64["sweeps"] / (float(np.mean(t64)) * 1e-3), "unit": "obs/s",
                        "sweep_launch_ms": sw64,
                        "roofline_frac": algorithmic_bytes(d.n_obs, d.n_views) / (sw64 * 1e-3) / 1e9 / measured_peaks()[0]}
        prob.run(opts)  # leave the default variant's result in the problem
        intr, q, t, reps = prob.download()

    # ---- partitioned solve against the un-partitioned one (rank 0 solves the whole problem on its GPU)
    partition_check = None
    if world > 1 and wl["poff"] is None:
        mine = torch.tensor(np.concatenate([intr[0], [reps[0]["rms"], reps[0]["final_cost"]]]), device="cuda")
        allv = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allv, mine)
        identical = all(torch.equal(allv[0], x) for x in allv)
        if rank == 0:
            from camera_calibrator_b200 import synth

            V, P, _ = WORKLOADS[args.workload]
            total = wl["total_views"]
            fd = synth.planar_views(total, P, seed=0)
            fi, fq, ft = synth.perturbed_start(fd)
            ctx1 = cc.Context(local)
            ri, _, _, rrep = ctx1.intrinsics_solve(fd.u, fd.v, fd.X, fd.Y, fd.Z, fd.view_offsets, fi, fq, ft)
            ctx1.close()
            partition_check = {"intr_rel_err": float(np.max(np.abs(intr[0] - ri) / np.abs(ri))),
                               "rms_abs_err": float(abs(reps[0]["rms"] - rrep["rms"])),
                               "replicated_bit_identical": bool(identical),
                               "sweeps": [int(reps[0]["sweeps"]), int(rrep["sweeps"])],
                               "reference": "the same 100,000 views solved un-partitioned on rank 0's GPU"}
            del fd
        barrier()

    # ---- dominant kernel against the roofline (this rank's share, rank 0 reports)
    sweep_ms_avg = sweep_ms / max(n_sweeps, 1)
    alg_bytes = algorithmic_bytes(d.n_obs, d.n_views)
    hbm_peak, peak_src = measured_peaks()
    achieved = alg_bytes / (sweep_ms_avg * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        traffic = tj.get(args.workload, {}).get("sweep_dram_bytes_per_launch")

    # ---- end to end through the C-ABI with host buffers (H2D + solve + D2H timed)
    hu, hv, hX, hY, hZ = (pinned(a) for a in (d.u, d.v, d.X, d.Y, d.Z))
    e2e_opts = cc.default_opts()
    h2d = 20 * d.n_obs + 8 * (d.n_views + 1) + 8 * 7 * d.n_views + 8 * 9 * wl["n_problems"]
    d2h = 8 * 7 * d.n_views + (8 * 9 + 56) * wl["n_problems"]
    e2e_times = []
    n_e2e = max(1, min(args.steps, 3))
    for i in range(1 + n_e2e):
        barrier()
        t0 = time.perf_counter()
        r = ctx.intrinsics_solve(hu, hv, hX, hY, hZ, d.view_offsets, wl["intr"], wl["q"], wl["t"],
                                 opts=e2e_opts, problem_offsets=wl["poff"]) if not wl["shared"] else None
        if wl["shared"]:
            p2 = ctx.intrinsics_problem(hu, hv, hX, hY, hZ, d.view_offsets, wl["intr"], wl["q"], wl["t"], shared=True)
            p2.run(e2e_opts)
            r = p2.download()
            p2.close()
        dt = time.perf_counter() - t0
        if i > 0:
            e2e_times.append(dt)
    e2e_t = allmax(float(np.mean(e2e_times)))
    e2e_value = obs_sweeps / e2e_t

    # ---- BASELINE metric (iii), full calibration wall time: Calibrator::Estimate through the C-ABI,
    # host buffers in, no starting point given (H2D + closed-form init on the device + LM solve + D2H)
    full_calib = None
    if world == 1:
        est_times = []
        for i in range(1 + n_e2e):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            est = ctx.intrinsics_estimate(hu, hv, hX, hY, hZ, d.view_offsets, problem_offsets=wl["poff"])
            if i > 0:
                est_times.append(time.perf_counter() - t0)
        erep = est[3] if wl["poff"] is None else est[3][0]
        full_calib = {"call": "ccg_intrinsics_estimate (Calibrator::Estimate)", "wall_ms": float(np.mean(est_times)) * 1e3,
                      "iterations": erep["iterations"], "sweeps": erep["sweeps"], "rms_px": erep["rms"],
                      "termination": erep["termination"]}

    line = None
    if rank == 0:
        # CPU baseline: oracle port on a bounded sample of the same workload, all host threads
        import oracle

        oracle.build()
        sample = cpu_sample(args.workload)
        cpu_v, cpu_dt, _ = cpu_solve_once(sample)
        if full_calib is not None and wl["poff"] is None:
            # the same call on the CPU port (init: one SVD per view, single thread as in the reference)
            nv_s = min(d.n_views, 1000)
            o1 = int(d.view_offsets[nv_s])
            t0 = time.perf_counter()
            oracle.estimate(d.u[:o1], d.v[:o1], d.X[:o1], d.Y[:o1], d.Z[:o1], d.view_offsets[:nv_s + 1])
            full_calib["cpu_port"] = {"sample": f"first {nv_s} of {d.n_views} views", "wall_ms": (time.perf_counter() - t0) * 1e3,
                                      "cores": oracle.num_threads()}
        line = {
            "metric": "residual+Jacobian observations/s (full LM solve)", "value": value, "unit": "obs/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": args.scaling if wl["poff"] is None else "weak",
            "vs_baseline": None,
            "dtype": "f64 residual/cost/gradient + reduced system; J'J blocks: fp16-operand tensor-core products (fp32 accumulate) of float32 rows",
            "data": "synthetic",
            "config": workload_config(args.workload, world, args.scaling),
            "lm_iterations_per_s": iters / (ms_per_step * 1e-3),
            "sweep_launches_per_step": n_sweeps / n_prof,
            "value_counts": "observations x accepted+rejected candidate evaluations (report.sweeps); the initial "
                            "evaluation is one more sweep launch per step and is not counted",
            "full_solve_wall_ms": wall_ms / args.steps,
            "solve": {"iterations": reps[0]["iterations"], "sweeps": reps[0]["sweeps"],
                      "rejected": reps[0]["rejected"], "termination": reps[0]["termination"],
                      "rms_px": reps[0]["rms"], "gradient_norm": reps[0]["gradient_norm"],
                      "intrinsics": [float(x) for x in intr[0]]},
            "breakdown_ms_per_step": {"sweep": sweep_ms / n_prof, "eliminate": schur_ms / n_prof,
                                      "reduce+decide+solve(+allreduce)": small_ms / n_prof,
                                      "allreduces_per_step": n_allreduce / args.steps,
                                      "note": "separate pass with CUDA events around every kernel class"},
            "roofline": {"kernel": "k_intr_sweep_ws (candidate poses + both roles, one launch)", "bound": "hbm", "achieved": achieved, "peak": hbm_peak,
                         "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic,
                         "peak_source": peak_src, "launch_ms": sweep_ms_avg, "launches": n_sweeps,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "fp64": {"achieved_tflops": FP64_FLOP_PER_OBS * d.n_obs / (sweep_ms_avg * 1e-3) / 1e12,
                                  "peak_tflops": fp64_peak, "flop_per_obs": FP64_FLOP_PER_OBS,
                                  "peak_source": "DFMA micro-benchmark run in this process"},
                         "fp32": {"achieved_tflops": FP32_FLOP_PER_OBS * d.n_obs / (sweep_ms_avg * 1e-3) / 1e12,
                                  "flop_per_obs": FP32_FLOP_PER_OBS},
                         "f16_mma": {"achieved_tflops": MMA_FLOP_PER_OBS * d.n_obs / (sweep_ms_avg * 1e-3) / 1e12,
                                      "flop_per_obs": MMA_FLOP_PER_OBS}},
            "cpu_baseline": {"value": cpu_v, "unit": "obs/s", "cores": oracle.num_threads(), "kind": "port",
                             "sample": sample[5], "seconds": cpu_dt},
            "e2e": {"value": e2e_value, "unit": "obs/s", "h2d_bytes_per_step": int(allsum(float(h2d))) if world > 1 else int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_t * 1e3,
                    "buffers": "pinned host memory, warm device memory pool (first call discarded)",
                    "pcie_floor_ms": h2d / 55e9 * 1e3},
            "e2e_cold_pageable": None if e2e_cold_ms is None else {
                "ms": e2e_cold_ms, "value": obs_sweeps / (e2e_cold_ms * 1e-3), "unit": "obs/s",
                "note": "first call of the process: pageable numpy arrays, device memory pool still empty"},
            "fp64_hessian": fp64_variant,
            "partition_check": partition_check,
            "full_calibration": full_calib,
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
    else:
        allsum(float(h2d))
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    if line is not None:
        print(json.dumps(line), flush=True)
    prob.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c5", choices=sorted(WORKLOADS) + ["c3"])
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.workload == "c3":
        if args.impl == "reference":
            raise SystemExit("--workload c3 (rig): the CPU port is timed beside the GPU arm (cpu_baseline)")
        run_rig(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
