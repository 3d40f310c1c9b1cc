/* ORACLE — TEST INFRASTRUCTURE ONLY.
 *
 * C interface of the CPU restatement of the reference hot path (the
 * Levenberg-Marquardt reprojection-error minimisation behind
 * Calibrator::Optimize, reference src/calibrator.cpp:221-336, and
 * ExtrinsicsCalibrator::Optimize, reference src/extrinsics_calibrator.cpp:86-257).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product path
 * (camera_calibrator_b200/) never links, imports or calls it.
 *
 * PARITY PINNING: the reference cannot be compiled here (Ceres, Eigen, OpenCV
 * C++ and Catch2 are absent) and it stores no golden
 * vectors.  This oracle is pinned against (1) the reference's own test
 * assertion — recovered K within 1 % of the generator's K on the
 * test_calibrator.cpp:45-75 scenario — (2) cv2.calibrateCamera, the function the
 * reference's EstimateOpenCv calls (calibrator.cpp:38) and which minimises the
 * same cost, (3) scipy.optimize.least_squares on the same residuals, and
 * (4) torch.float64 autograd for every Jacobian block.  The Ceres ITERATE
 * PATH is unpinned (Ceres is not available); only the converged minimum is
 * compared.  See DESIGN.md "Oracle".
 */
#ifndef ORACLE_H_
#define ORACLE_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  int32_t max_iterations;       /* reference: 100 (calibrator.cpp:319) / 1000 (extrinsics_calibrator.cpp:211) */
  int32_t verbose;
  double initial_radius;        /* [ext] Ceres default 1e4 */
  double max_radius;            /* [ext] 1e16 */
  double min_radius;            /* [ext] 1e-32 */
  double min_relative_decrease; /* [ext] 1e-3 */
  double min_diagonal;          /* [ext] min_lm_diagonal 1e-6 */
  double max_diagonal;          /* [ext] max_lm_diagonal 1e32 */
  double function_tolerance;    /* accepted step with cost decrease <= tol * cost */
  double gradient_tolerance;    /* sqrt(g' D^-1 g) <= tol * sqrt(2 cost), D = diag(J'J) */
  double parameter_tolerance;   /* sqrt(d' D d) <= tol * sqrt(2 cost): take the step if the cost does not rise, then stop */
  double huber_a;               /* rig only; <= 0 disables the loss */
  double noise_factor;          /* predicted decrease <= noise_factor*eps/2*max|meas|*sqrt(2 cost): below cost resolution, take the step and stop */
} oracle_opts;

typedef struct {
  int32_t iterations;   /* LM iterations = linear solves attempted */
  int32_t accepted;
  int32_t rejected;
  int32_t sweeps;       /* residual+Jacobian evaluations over all observations */
  int32_t termination;  /* 1 gradient, 2 function, 3 parameter, 4 max iterations, 5 radius underflow, <0 failure */
  int32_t pad_;
  double initial_cost;
  double final_cost;
  double gradient_norm; /* sqrt(g' D^-1 g) / sqrt(2 cost) at the returned point */
  double rms;           /* sqrt(sum ||r||^2 / n_obs) at the returned point (robustified residuals NOT used) */
} oracle_report;

void oracle_default_opts(oracle_opts* o);
int oracle_num_threads(void);
void oracle_set_num_threads(int n);

/* Single-camera path.  Observations are SoA float32 (the precision of the
 * reference API, types.hh:10-11), views are [view_offsets[i], view_offsets[i+1]).
 * intr: fx,fy,px,py,k1,k2,p1,p2,k3.  q: 4 per view (w,x,y,z).  t: 3 per view. */
int oracle_intrinsics_solve(const float* u, const float* v, const float* X, const float* Y,
                            const float* Z, const int64_t* view_offsets, int64_t n_views,
                            double* intr, double* q, double* t, uint32_t const_intr_mask,
                            const oracle_opts* opts, oracle_report* report);

/* One evaluation at the given parameters.  Any output pointer may be NULL.
 *   res   : 2 per observation (u then v)
 *   Vb    : 36 per view (6x6 row-major, tangent order rot(3) then trans(3))
 *   Wb    : 54 per view (6x9 row-major: pose tangent x intrinsic)
 *   gvb   : 6 per view
 *   U     : 81 (9x9 row-major), gu : 9, cost : 1 (= 0.5 sum r^2) */
int oracle_intrinsics_eval(const float* u, const float* v, const float* X, const float* Y,
                           const float* Z, const int64_t* view_offsets, int64_t n_views,
                           const double* intr, const double* q, const double* t, double* res,
                           double* Vb, double* Wb, double* gvb, double* U, double* gu,
                           double* cost);

/* Batch of independent single-camera problems: problem b owns views
 * [problem_offsets[b], problem_offsets[b+1]); intr is 9 per problem. */
int oracle_intrinsics_solve_batch(const float* u, const float* v, const float* X, const float* Y,
                                  const float* Z, const int64_t* view_offsets,
                                  const int64_t* problem_offsets, int64_t n_problems, double* intr,
                                  double* q, double* t, uint32_t const_intr_mask,
                                  const oracle_opts* opts, oracle_report* reports);

/* Rig path.  Frame f owns observations [frame_offsets[f], frame_offsets[f+1]);
 * observation o sees world point obs_point[o] (index into Xw, 3 doubles-as-float
 * each) from camera obs_cam[o] at normalised image point (u[o], v[o]).
 * q_c/t_c: per camera, q_f/t_f: per frame.  cam_frozen: 1 byte per camera.
 * obs_cost (may be NULL): 0.5 * rho(||r||^2) per observation at the solution
 * (extrinsics_calibrator.cpp:219-225). */
int oracle_rig_solve(const float* u, const float* v, const int32_t* obs_cam,
                     const int64_t* obs_point, const int64_t* frame_offsets, int64_t n_frames,
                     const float* Xw, int64_t n_points, int32_t n_cams, double* q_c, double* t_c,
                     const uint8_t* cam_frozen, double* q_f, double* t_f, double* obs_cost,
                     const oracle_opts* opts, oracle_report* report);

/*   res : 2 per observation, already scaled by sqrt(rho')
 *   Vb  : 36 per frame, Wb : 6 x (6*n_cams) per frame row-major, gvb : 6 per frame
 *   U   : (6 n_cams)^2 row-major, gu : 6 n_cams, cost : 0.5 sum rho */
int oracle_rig_eval(const float* u, const float* v, const int32_t* obs_cam,
                    const int64_t* obs_point, const int64_t* frame_offsets, int64_t n_frames,
                    const float* Xw, int64_t n_points, int32_t n_cams, const double* q_c,
                    const double* t_c, const double* q_f, const double* t_f, double huber_a,
                    double* res, double* Vb, double* Wb, double* gvb, double* U, double* gu,
                    double* cost);

/* Restatement of DataGenerator::GetDistortedPointsPlanar (reference
 * src/data_generator.cpp:77-124) called n_views times on one generator object
 * seeded with `seed` (reference: fixed 0, data_generator.hh:43).  K is 3x3
 * row-major float, dist is k1,k2,p1,p2,k3.  Outputs: n_views*n_points each.
 * If plane_out != NULL it receives 4 floats per view (the random plane). */
int oracle_datagen_planar(uint32_t seed, int32_t width, int32_t height, const float* K,
                          const float* dist, float noise_px, int32_t n_points, int32_t n_views,
                          float* u, float* v, float* X, float* Y, float* Z, float* plane_out,
                          int64_t* n_rejected);

/* Restatement of the initialisation inside Calibrator::Estimate (reference
 * src/calibrator.cpp:47-65; geometry.cpp:70-105, 123-177, 179-203).
 * H_out: 9 floats per view (row-major), K_out: 9 floats, q_out: 4 floats per
 * view (w,x,y,z), t_out: 3 floats per view. */
int oracle_estimate_init(const float* u, const float* v, const float* X, const float* Y,
                         const float* Z, const int64_t* view_offsets, int64_t n_views,
                         float* H_out, float* K_out, float* q_out, float* t_out);

/* EstimateHomography (geometry.cpp:70-105) on n point pairs: p2 ~ H p1. */
int oracle_homography(const float* x1, const float* y1, const float* x2, const float* y2,
                      int64_t n, float* H_out);

#ifdef __cplusplus
}
#endif
#endif
