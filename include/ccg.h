/* ccg — camera-calibrator GPU hot path, C-ABI (sm_100a / B200).
 *
 * This is the drop-in boundary: the entry points a reference maintainer binds
 * in place of the Ceres problem build + ceres::Solve inside
 *   Calibrator::Optimize            reference src/calibrator.cpp:221-336
 *   ExtrinsicsCalibrator::Optimize  reference src/extrinsics_calibrator.cpp:86-257
 * The host classes keep doing the float->double packing the reference does
 * (calibrator.cpp:245-303, extrinsics_calibrator.cpp:105-137) and call these.
 * Plain pointers and sizes only; no exceptions cross this boundary; every
 * function returns 0 or a negative ccg_status.  The caller owns all host
 * buffers, the library owns all device memory inside ccg_ctx / handles.
 * One context per host thread.  There is NO CPU fallback: without a CUDA
 * device every compute entry point returns CCG_ERR_CUDA.
 */
#ifndef CCG_H_
#define CCG_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  CCG_OK = 0,
  CCG_ERR_BAD_ARG = -1,
  CCG_ERR_CUDA = -2,
  CCG_ERR_NOT_FINITE = -3, /* initial cost is NaN/Inf */
  CCG_ERR_NCCL = -4,
  CCG_ERR_UNSUPPORTED = -5
} ccg_status;

typedef struct ccg_ctx ccg_ctx;
typedef struct ccg_intr_problem ccg_intr_problem; /* device-resident single-camera problem (batch) */
typedef struct ccg_rig_problem ccg_rig_problem;   /* device-resident rig problem */

/* Solver options.  The reference hard-codes its Ceres options
 * (calibrator.cpp:314-321, extrinsics_calibrator.cpp:206-213); the fields below
 * are the trust-region constants those options select [ext Ceres defaults] plus
 * the convergence tolerances of this solver (DESIGN.md "LM loop"). */
typedef struct {
  int32_t max_iterations;       /* 100 (calibrator.cpp:319) / 1000 (extrinsics_calibrator.cpp:211) */
  int32_t profile;              /* 1: record CUDA events around every kernel class (ccg_timing) */
  double initial_radius;        /* 1e4 */
  double max_radius;            /* 1e16 */
  double min_radius;            /* 1e-32 */
  double min_relative_decrease; /* 1e-3 */
  double min_diagonal;          /* 1e-6 */
  double max_diagonal;          /* 1e32 */
  double function_tolerance;    /* accepted step with cost decrease <= tol * cost */
  double gradient_tolerance;    /* sqrt(g' D^-1 g) <= tol * sqrt(2 cost), D = diag(J'J) */
  double parameter_tolerance;   /* sqrt(d' D d) <= tol * sqrt(2 cost): take the step if the cost does not rise, stop */
  double huber_a;               /* rig only: (double)(3.0f/500.0f), extrinsics_calibrator.cpp:176; <= 0: no loss */
  double noise_factor;          /* 4: a step whose predicted decrease is <= noise_factor * (eps/2) * max|measurement| *
                                   sqrt(2 cost) is below the FP64 resolution of the cost: take it and stop */
  int32_t hessian_fp64;         /* sweep variant (single camera and rig).  Residual, cost and gradient are FP64 in all of them (same
                                   fixed point); they differ in how the J'J blocks are formed:
                                   0 (default) tensor-core products (mma.sync m16n8k16, FP32 accumulation) of the
                                   float32 Jacobian rows rounded to FP16 (11 significant bits, as TF32; pose columns
                                   scaled by a power of two into range), 1 everything in FP64 */
  int32_t reserved_;
} ccg_opts;

typedef struct {
  int32_t iterations; /* linear solves attempted */
  int32_t accepted;
  int32_t rejected;
  int32_t sweeps;     /* residual+Jacobian+normal-equation passes over the observations */
  int32_t termination; /* 1 gradient, 2 function, 3 parameter, 4 max iterations, 5 radius underflow, <0 ccg_status */
  int32_t pad_;
  double initial_cost;
  double final_cost;
  double gradient_norm; /* sqrt(g' D^-1 g) / sqrt(2 cost) at the returned point */
  double rms;           /* sqrt(sum ||r||^2 / n_obs), raw (un-robustified) residuals */
} ccg_report;

/* Device timings of the last ccg_*_run with opts.profile = 1 (CUDA events on
 * the library's stream). */
typedef struct {
  double total_ms;      /* whole device-resident LM loop */
  double sweep_ms;      /* sum over residual+Jacobian+normal-equation sweep launches */
  double schur_ms;      /* sum over pose-elimination launches */
  double small_ms;      /* sum over reduce/decide/solve launches (+ allreduce) */
  int32_t sweep_launches;
  int32_t schur_launches;
  int32_t small_launches;
  int32_t allreduces;
  int64_t sweep_obs;    /* observations processed per sweep launch (this rank) */
  int64_t sweep_views;  /* views (frames) processed per sweep launch (this rank) */
} ccg_timing;

void ccg_default_opts(ccg_opts* o);
const char* ccg_status_string(int status);
const char* ccg_last_error(const ccg_ctx* ctx);

/* device = CUDA ordinal this context binds to (one context per GPU / rank). */
int ccg_init(int device, ccg_ctx** out);
void ccg_destroy(ccg_ctx* ctx);
/* Device buffers of destroyed problems stay cached in the context's memory pool (a second
 * solve of the same size allocates nothing); ccg_trim returns the cached memory to the driver. */
int ccg_trim(ccg_ctx* ctx);
int ccg_device_info(ccg_ctx* ctx, int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor,
                    int64_t* global_mem_bytes);

/* ---- multi-GPU: one process per GPU, one NCCL communicator per context.
 * The 128-byte id comes from rank 0 (ccg_comm_unique_id) and is distributed by
 * the launcher (torch.distributed / file / env).  After ccg_comm_init, problems
 * created with shared=1 are treated as ONE calibration whose views (frames) are
 * partitioned across ranks: each LM iteration all-reduces the packed reduced
 * system - over peer memory (CUDA IPC mailboxes, NVLink) inside
 * the accept/reject and reduced-solve kernels, for the single-camera and the
 * rig solver; over ncclAllReduce when the mailboxes cannot be mapped or
 * CCG_NO_PEER=1.  ccg_comm_init also exchanges the mailbox handles. */
int ccg_comm_unique_id(uint8_t id_out[128]);
int ccg_comm_init(ccg_ctx* ctx, int rank, int world_size, const uint8_t id[128]);
int ccg_comm_barrier(ccg_ctx* ctx);

/* ================= single camera (Calibrator::Optimize) =================
 * Observations: SoA float32 (reference API precision, types.hh:10-11), view i
 * owns [view_offsets[i], view_offsets[i+1]).  intr in/out: fx,fy,px,py,k1,k2,p1,
 * p2,k3 (calibrator.cpp:168-179).  q in/out: 4 per view (w,x,y,z), t in/out: 3 per
 * view (calibrator.cpp:273-281).  const_intr_mask bit k freezes intr[k]
 * (ForceDistortionToConstant(i) sets bit i+4, calibrator.cpp:338-340).
 * Host buffers in, host buffers out: H2D + solve + D2H. */
int ccg_intrinsics_solve(ccg_ctx* ctx, const float* u, const float* v, const float* X,
                         const float* Y, const float* Z, const int64_t* view_offsets,
                         int64_t n_views, double* intr, double* q, double* t,
                         uint32_t const_intr_mask, const ccg_opts* opts, ccg_report* report);

/* B independent calibrations in one call: problem b owns views
 * [problem_offsets[b], problem_offsets[b+1]); intr is 9 per problem. */
int ccg_intrinsics_solve_batch(ccg_ctx* ctx, const float* u, const float* v, const float* X,
                               const float* Y, const float* Z, const int64_t* view_offsets,
                               int64_t n_views, const int64_t* problem_offsets,
                               int64_t n_problems, double* intr, double* q, double* t,
                               uint32_t const_intr_mask, const ccg_opts* opts,
                               ccg_report* reports);

/* Calibrator::Estimate (reference src/calibrator.cpp:47-68) in one call: the closed-form
 * initialisation (per-view DLT homography, Zhang's K with zero skew, per-view RecoverExtrinsics +
 * FixRotationMatrix; reference src/geometry.cpp:70-203) runs on the device on the uploaded
 * observations (world points are read as planar: X, Y), then the LM solve of ccg_intrinsics_solve.
 * intr, q, t are OUTPUTS only (q, t may be NULL); intr_init (may be NULL) receives the closed-form
 * starting point fx fy px py 0 0 0 0 0.  Needs at least 3 views with 4 or more points per problem
 * (CCG_ERR_BAD_ARG otherwise); CCG_ERR_NOT_FINITE when the closed form has no real solution. */
int ccg_intrinsics_estimate(ccg_ctx* ctx, const float* u, const float* v, const float* X,
                            const float* Y, const float* Z, const int64_t* view_offsets,
                            int64_t n_views, double* intr, double* q, double* t, double* intr_init,
                            uint32_t const_intr_mask, const ccg_opts* opts, ccg_report* report);
int ccg_intrinsics_estimate_batch(ccg_ctx* ctx, const float* u, const float* v, const float* X,
                                  const float* Y, const float* Z, const int64_t* view_offsets,
                                  int64_t n_views, const int64_t* problem_offsets,
                                  int64_t n_problems, double* intr, double* q, double* t,
                                  double* intr_init, uint32_t const_intr_mask, const ccg_opts* opts,
                                  ccg_report* reports);

/* EstimateHomography (reference src/geometry.cpp:70-107) for n_sets point sets at once: set i maps the points
 * (x1, y1)[set_offsets[i] .. set_offsets[i+1]) to (x2, y2) of the same range.  H receives 9 doubles per set (row-major,
 * unit Frobenius norm; the caller divides by H[8] as the reference's users do); H[9 i] is NaN for a set with fewer than
 * 4 points or degenerate geometry.  This is the first kernel of ccg_intrinsics_estimate's initialisation. */
int ccg_homography_batch(ccg_ctx* ctx, const float* x1, const float* y1, const float* x2, const float* y2,
                         const int64_t* set_offsets, int64_t n_sets, double* H);

/* Batched lens model of calibrator::Calibrator, host buffers in and out.
 * ccg_undistort: pixels -> undistorted normalised coordinates (Calibrator::Undistort, reference
 *   src/calibrator.cpp:97-155); K_inv = the float32 inverse of K, row-major, as the class forms it;
 *   dist = k1 k2 p1 p2 k3.
 * ccg_distort: normalised coordinates -> distorted pixels (Calibrator::Distort, calibrator.cpp:157-166);
 *   K row-major float32 (zero skew). */
int ccg_undistort(ccg_ctx* ctx, const float* px, const float* py, int64_t n, const float K_inv[9],
                  const float dist[5], float* xn, float* yn);
int ccg_distort(ccg_ctx* ctx, const float* xn, const float* yn, int64_t n, const float K[9],
                const float dist[5], float* px, float* py);

/* One residual+Jacobian+normal-equation evaluation at the given parameters
 * (the blocks the LM loop consumes), for parity tests.  Any output may be NULL.
 *   res 2/obs | Vb 36/view (6x6) | Wb 54/view (6x9) | gvb 6/view | U 81 | gu 9 | cost 1
 * Tangent order per view: rotation (half-angle vector, Ceres QuaternionManifold
 * convention) then translation.  hessian_fp64 selects the sweep variant (see ccg_opts): with 0
 * V, W, U are TF32 tensor-core products (~1e-3 relative), with 1 FP64; res, gvb, gu and cost are
 * FP64 in both. */
int ccg_intrinsics_eval(ccg_ctx* ctx, const float* u, const float* v, const float* X,
                        const float* Y, const float* Z, const int64_t* view_offsets,
                        int64_t n_views, const double* intr, const double* q, const double* t,
                        double* res, double* Vb, double* Wb, double* gvb, double* U, double* gu,
                        double* cost, int hessian_fp64);

/* Device-resident form of the same solve: upload once, run many times.
 * shared != 0 (requires ccg_comm_init): the views given here are this rank's
 * partition of one calibration shared by all ranks (n_problems must be 1). */
int ccg_intrinsics_create(ccg_ctx* ctx, const float* u, const float* v, const float* X,
                          const float* Y, const float* Z, const int64_t* view_offsets,
                          int64_t n_views, const int64_t* problem_offsets, int64_t n_problems,
                          const double* intr, const double* q, const double* t,
                          uint32_t const_intr_mask, int shared, ccg_intr_problem** out);
int ccg_intrinsics_reset(ccg_intr_problem* p); /* restore the uploaded initial parameters */
int ccg_intrinsics_run(ccg_intr_problem* p, const ccg_opts* opts, ccg_timing* timing);
/* n back-to-back sweep launches at the current parameters, CUDA-event timed. */
int ccg_intrinsics_time_sweep(ccg_intr_problem* p, int n, float* ms_per_launch);
int ccg_intrinsics_download(ccg_intr_problem* p, double* intr, double* q, double* t,
                            ccg_report* reports);
/* Per-sweep trace of problem b of the last run: up to `cap` rows of 6 doubles
 * (cost, candidate cost, gain ratio, trust-region radius, relative step norm,
 * accepted flag), oldest first; *n_rows receives the row count (<= 48). */
int ccg_intrinsics_history(ccg_intr_problem* p, int64_t b, double* rows, int cap, int* n_rows);
int64_t ccg_intrinsics_device_bytes(const ccg_intr_problem* p);
void ccg_intrinsics_destroy(ccg_intr_problem* p);

/* ================= rig (ExtrinsicsCalibrator::Optimize) =================
 * Frame f owns observations [frame_offsets[f], frame_offsets[f+1]) in the
 * reference's order (frame by frame, insertion order inside a frame,
 * extrinsics_calibrator.cpp:155-160).  Observation o: camera obs_cam[o], world
 * point obs_point[o] (global world_point_id, index into Xw), normalised image
 * point (u[o], v[o]).  q_c/t_c per camera, q_f/t_f per frame, in/out.
 * cam_frozen[c] != 0 keeps camera c constant (extrinsics_calibrator.cpp:189-193).
 * obs_cost out (may be NULL): 0.5*rho(||r||^2) per observation at the solution
 * (extrinsics_calibrator.cpp:219-225), in the caller's observation order. */
int ccg_rig_solve(ccg_ctx* ctx, const float* u, const float* v, const int32_t* obs_cam,
                  const int64_t* obs_point, const int64_t* frame_offsets, int64_t n_frames,
                  const float* Xw, int64_t n_points, int32_t n_cams, double* q_c, double* t_c,
                  const uint8_t* cam_frozen, double* q_f, double* t_f, double* obs_cost,
                  const ccg_opts* opts, ccg_report* report);

/*   res 2/obs (scaled by sqrt(rho')) | Vb 36/frame | Wb 6 x 6C per frame | gvb 6/frame |
 *   U (6C)^2 | gu 6C | cost = 0.5 sum rho */
int ccg_rig_eval(ccg_ctx* ctx, const float* u, const float* v, const int32_t* obs_cam,
                 const int64_t* obs_point, const int64_t* frame_offsets, int64_t n_frames,
                 const float* Xw, int64_t n_points, int32_t n_cams, const double* q_c,
                 const double* t_c, const double* q_f, const double* t_f, double huber_a,
                 double* res, double* Vb, double* Wb, double* gvb, double* U, double* gu,
                 double* cost);

int ccg_rig_create(ccg_ctx* ctx, const float* u, const float* v, const int32_t* obs_cam,
                   const int64_t* obs_point, const int64_t* frame_offsets, int64_t n_frames,
                   const float* Xw, int64_t n_points, int32_t n_cams, const double* q_c,
                   const double* t_c, const uint8_t* cam_frozen, const double* q_f,
                   const double* t_f, int shared, ccg_rig_problem** out);
int ccg_rig_reset(ccg_rig_problem* p);
int ccg_rig_run(ccg_rig_problem* p, const ccg_opts* opts, ccg_timing* timing);
int ccg_rig_download(ccg_rig_problem* p, double* q_c, double* t_c, double* q_f, double* t_f,
                     double* obs_cost, ccg_report* report);
void ccg_rig_destroy(ccg_rig_problem* p);

/* ================= measurement helpers =================
 * FP64 FMA peak of this GPU (dependent-chain-free DFMA loop on every SM),
 * the second roofline denominator of this path. */
int ccg_measure_fp64_peak(ccg_ctx* ctx, double* tflops);
/* Device-to-device copy bandwidth (read+write bytes / s) over `bytes` bytes. */
int ccg_measure_copy_bandwidth(ccg_ctx* ctx, int64_t bytes, double* gbps);
/* Number of kernels this library has launched on this context so far. */
int64_t ccg_kernel_launches(const ccg_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* CCG_H_ */
