// Host side of calibrator::Calibrator (reference src/calibrator.cpp): float->double
// packing exactly where the reference does it (calibrator.cpp:245-303), the solve
// itself behind ccg_intrinsics_solve, float write-back of K and distortion only
// (calibrator.cpp:326-335).
#include "calibrator.hh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <stdexcept>
#include <string>

#include "../../include/ccg.h"
#include "geometry.hh"

namespace calibrator {

namespace {
struct CtxHolder {
  ccg_ctx* ctx = nullptr;
  ~CtxHolder() {
    if (ctx) ccg_destroy(ctx);
  }
};
}  // namespace

ccg_ctx* HostContext() {
  thread_local CtxHolder holder;
  if (!holder.ctx) {
    int device = 0;
    if (const char* e = std::getenv("CCG_DEVICE")) device = std::atoi(e);
    const int rc = ccg_init(device, &holder.ctx);
    if (rc != CCG_OK)
      throw std::runtime_error(std::string("camera_calibrator_b200: no usable CUDA device (") +
                               ccg_status_string(rc) + "); there is no CPU fallback");
  }
  return holder.ctx;
}

// intrinsic block order, reference calibrator.cpp:168-179
enum { OFF_FX = 0, OFF_FY, OFF_PX, OFF_PY, OFF_K1, OFF_K2, OFF_P1, OFF_P2, OFF_K3, NUM_INTR };

Calibrator::Calibrator(const int img_width, const int img_height) : width_(img_width), height_(img_height) {}

namespace {
SolveReport ToReport(const ccg_report& rep) {
  SolveReport r;
  r.iterations = rep.iterations;
  r.accepted = rep.accepted;
  r.rejected = rep.rejected;
  r.sweeps = rep.sweeps;
  r.termination = rep.termination;
  r.initial_cost = rep.initial_cost;
  r.final_cost = rep.final_cost;
  r.gradient_norm = rep.gradient_norm;
  r.rms = rep.rms;
  return r;
}
}  // namespace

namespace {
// float32 API arrays -> the SoA the C-ABI takes (the packing of calibrator.cpp:283-303)
void PackObservations(const std::vector<Points2D>& img, const std::vector<Points3D>& world, const char* who,
                      std::vector<int64_t>* off, std::vector<float>* u, std::vector<float>* v,
                      std::vector<float>* X, std::vector<float>* Y, std::vector<float>* Z) {
  const size_t n_img = img.size();
  off->assign(n_img + 1, 0);
  for (size_t i = 0; i < n_img; ++i) {
    if (img[i].size() != world[i].size())
      throw std::invalid_argument(std::string(who) + ": point counts differ in a view");
    (*off)[i + 1] = (*off)[i] + (int64_t)img[i].size();
  }
  const size_t n = (size_t)(*off)[n_img];
  u->resize(n); v->resize(n); X->resize(n); Y->resize(n); Z->resize(n);
  for (size_t i = 0; i < n_img; ++i) {
    size_t o = (size_t)(*off)[i];
    for (size_t j = 0; j < img[i].size(); ++j, ++o) {
      (*u)[o] = img[i][j].x();
      (*v)[o] = img[i][j].y();
      (*X)[o] = world[i][j].x();
      (*Y)[o] = world[i][j].y();
      (*Z)[o] = world[i][j].z();
    }
  }
}
}  // namespace

void Calibrator::Estimate(const std::vector<Points2D>& in_img_points,
                          const std::vector<Points3D>& in_world_points) {
  if (in_img_points.size() != in_world_points.size())
    throw std::invalid_argument("Calibrator::Estimate: image/world view counts differ");
  bool zero_dist = true;
  for (int k = 0; k < 5; ++k) zero_dist = zero_dist && dist_(k) == 0.0f;
  if (!zero_dist) {
    // a caller-set distortion is the starting point of the solve (calibrator.cpp:262-266): keep the
    // reference's literal sequence, closed forms on the host, then Optimize
    std::vector<Matrix3> Hs;
    Hs.reserve(in_img_points.size());
    for (size_t i = 0; i < in_img_points.size(); ++i)
      Hs.push_back(EstimateHomography(in_world_points[i], in_img_points[i]));
    K_ = EstimateKFromHomographies(Hs);
    std::vector<Quaternion> qs;
    std::vector<Point3D> ts;
    const Matrix3 K_inv = K_.inverse();
    for (const auto& H : Hs) {
      const auto [R, t] = RecoverExtrinsics(K_inv, H);
      qs.emplace_back(R);
      ts.push_back(t);
    }
    Optimize(in_img_points, in_world_points, qs, ts);
    return;
  }
  // default path: initialisation and solve on the device behind one call (the observations cross
  // PCIe once; the per-view homographies are one warp each instead of a serial SVD loop)
  const size_t n_img = in_img_points.size();
  std::vector<int64_t> off;
  std::vector<float> u, v, X, Y, Z;
  PackObservations(in_img_points, in_world_points, "Calibrator::Estimate", &off, &u, &v, &X, &Y, &Z);
  std::vector<double> q(4 * std::max<size_t>(n_img, 1)), t(3 * std::max<size_t>(n_img, 1));
  double intr[NUM_INTR], init[NUM_INTR];
  uint32_t mask = 0;
  for (int idx : constant_intrinsics_)
    if (idx >= 0 && idx < NUM_INTR) mask |= 1u << idx;
  ccg_opts opts;
  ccg_default_opts(&opts);
  opts.max_iterations = 100;  // calibrator.cpp:319
  ccg_report rep;
  ccg_ctx* ctx = HostContext();
  const int rc = ccg_intrinsics_estimate(ctx, u.data(), v.data(), X.data(), Y.data(), Z.data(), off.data(),
                                         (int64_t)n_img, intr, q.data(), t.data(), init, mask, &opts, &rep);
  if (rc != CCG_OK)
    throw std::runtime_error(std::string("Calibrator::Estimate: ") + ccg_status_string(rc) + " " +
                             ccg_last_error(ctx));
  // the reference overwrites the whole matrix here (K_ = EstimateKFromHomographies(Hs): zero skew, last row
  // 0 0 1, calibrator.cpp:56) - an earlier SetK with other entries must not survive
  K_ = Matrix3::Identity();
  StoreResult(ToReport(rep), intr, q.data(), t.data(), n_img);
}

void Calibrator::EstimateOpenCv(const std::vector<Points2D>& in_img_points,
                                const std::vector<Points3D>& in_world_points) {
  // The reference hands the points to cv::calibrateCamera with flags = 0 (calibrator.cpp:16-45): OpenCV
  // ignores the current K_ / dist_ and the frozen set, initialises from the homographies itself and
  // minimises the same reprojection error over fx fy cx cy k1 k2 p1 p2 k3 and one pose per view.  The
  // minimum of that cost is what ccg_intrinsics_estimate converges to (the oracle is pinned against
  // cv2.calibrateCamera: RMS to 1e-9, K to 1e-6), so the call runs the device path with nothing frozen.
  if (in_img_points.size() != in_world_points.size())
    throw std::invalid_argument("Calibrator::EstimateOpenCv: image/world view counts differ");
  const size_t n_img = in_img_points.size();
  std::vector<int64_t> off;
  std::vector<float> u, v, X, Y, Z;
  PackObservations(in_img_points, in_world_points, "Calibrator::EstimateOpenCv", &off, &u, &v, &X, &Y, &Z);
  std::vector<double> q(4 * std::max<size_t>(n_img, 1)), t(3 * std::max<size_t>(n_img, 1));
  double intr[NUM_INTR], init[NUM_INTR];
  ccg_opts opts;
  ccg_default_opts(&opts);
  ccg_report rep;
  ccg_ctx* ctx = HostContext();
  const int rc = ccg_intrinsics_estimate(ctx, u.data(), v.data(), X.data(), Y.data(), Z.data(), off.data(),
                                         (int64_t)n_img, intr, q.data(), t.data(), init, 0u, &opts, &rep);
  if (rc != CCG_OK)
    throw std::runtime_error(std::string("Calibrator::EstimateOpenCv: ") + ccg_status_string(rc) + " " +
                             ccg_last_error(ctx));
  K_ = Matrix3::Identity();  // the reference assigns OpenCV's whole camera matrix (calibrator.cpp:41)
  StoreResult(ToReport(rep), intr, q.data(), t.data(), n_img);
}

void Calibrator::Optimize(const std::vector<Points2D>& in_img_points,
                          const std::vector<Points3D>& in_world_points, std::vector<Quaternion>& qs,
                          std::vector<Point3D>& ts) {
  const size_t n_img = in_img_points.size();
  if (n_img != in_world_points.size() || n_img != qs.size() || n_img != ts.size())
    throw std::invalid_argument("Calibrator::Optimize: per-view argument sizes differ");
  std::vector<int64_t> off;
  std::vector<float> u, v, X, Y, Z;
  PackObservations(in_img_points, in_world_points, "Calibrator::Optimize", &off, &u, &v, &X, &Y, &Z);
  std::vector<double> q(4 * n_img), t(3 * n_img);
  for (size_t i = 0; i < n_img; ++i) {
    q[4 * i + 0] = qs[i].w();
    q[4 * i + 1] = qs[i].x();
    q[4 * i + 2] = qs[i].y();
    q[4 * i + 3] = qs[i].z();
    t[3 * i + 0] = ts[i].x();
    t[3 * i + 1] = ts[i].y();
    t[3 * i + 2] = ts[i].z();
  }
  double intr[NUM_INTR];
  intr[OFF_FX] = K_(0, 0);
  intr[OFF_FY] = K_(1, 1);
  intr[OFF_PX] = K_(0, 2);
  intr[OFF_PY] = K_(1, 2);
  intr[OFF_K1] = dist_(0);
  intr[OFF_K2] = dist_(1);
  intr[OFF_P1] = dist_(2);
  intr[OFF_P2] = dist_(3);
  intr[OFF_K3] = dist_(4);
  uint32_t mask = 0;
  for (int idx : constant_intrinsics_)
    if (idx >= 0 && idx < NUM_INTR) mask |= 1u << idx;

  ccg_opts opts;
  ccg_default_opts(&opts);
  opts.max_iterations = 100;  // calibrator.cpp:319
  ccg_report rep;
  ccg_ctx* ctx = HostContext();
  const int rc = ccg_intrinsics_solve(ctx, u.data(), v.data(), X.data(), Y.data(), Z.data(), off.data(),
                                      (int64_t)n_img, intr, q.data(), t.data(), mask, &opts, &rep);
  if (rc != CCG_OK)
    throw std::runtime_error(std::string("Calibrator::Optimize: ") + ccg_status_string(rc) + " " +
                             ccg_last_error(ctx));
  StoreResult(ToReport(rep), intr, q.data(), t.data(), n_img);
}

void Calibrator::StoreResult(const SolveReport& rep, const double* intr, const double* q, const double* t,
                             size_t n_img) {
  report_ = rep;
  opt_q_.resize(n_img);
  opt_t_.resize(n_img);
  for (size_t i = 0; i < n_img; ++i) {
    opt_q_[i] = {q[4 * i], q[4 * i + 1], q[4 * i + 2], q[4 * i + 3]};
    opt_t_[i] = {t[3 * i], t[3 * i + 1], t[3 * i + 2]};
  }
  // Put params back (float32, K and distortion only; calibrator.cpp:326-335)
  K_(0, 0) = static_cast<float>(intr[OFF_FX]);
  K_(1, 1) = static_cast<float>(intr[OFF_FY]);
  K_(0, 2) = static_cast<float>(intr[OFF_PX]);
  K_(1, 2) = static_cast<float>(intr[OFF_PY]);
  dist_(0) = static_cast<float>(intr[OFF_K1]);
  dist_(1) = static_cast<float>(intr[OFF_K2]);
  dist_(2) = static_cast<float>(intr[OFF_P1]);
  dist_(3) = static_cast<float>(intr[OFF_P2]);
  dist_(4) = static_cast<float>(intr[OFF_K3]);
}

void Calibrator::GetOptimizedPoses(std::vector<std::array<double, 4>>* qs,
                                   std::vector<std::array<double, 3>>* ts) const {
  if (qs) *qs = opt_q_;
  if (ts) *ts = opt_t_;
}

void Calibrator::ForceDistortionToConstant(const int distortion_idx) {
  constant_intrinsics_.insert(distortion_idx + static_cast<int>(OFF_K1));
}

Points2D Calibrator::Undistort(const Points2D& img_points) {
  // The reference solves, per point, a 2-unknown least-squares problem with Ceres (DENSE_QR,
  // <= 1000 iterations) starting from the distorted coordinates (calibrator.cpp:97-155).  The same
  // fixed point is found on the device, one thread per point, with Newton steps on the 2x2 system
  // in double (ccg_undistort).
  const Matrix3 K_inv = K_.inverse();
  const size_t n = img_points.size();
  std::vector<float> px(n), py(n), xn(n), yn(n);
  for (size_t i = 0; i < n; ++i) {
    px[i] = img_points[i].x();
    py[i] = img_points[i].y();
  }
  const float dist[5] = {dist_(0), dist_(1), dist_(2), dist_(3), dist_(4)};
  ccg_ctx* ctx = HostContext();
  const int rc = ccg_undistort(ctx, px.data(), py.data(), (int64_t)n, K_inv.m, dist, xn.data(), yn.data());
  if (rc != CCG_OK)
    throw std::runtime_error(std::string("Calibrator::Undistort: ") + ccg_status_string(rc) + " " + ccg_last_error(ctx));
  Points2D out;
  out.reserve(n);
  for (size_t i = 0; i < n; ++i) out.emplace_back(xn[i], yn[i]);
  return out;
}

Points2D Calibrator::Distort(const Points2D& normalized_points) {
  // float32 lens model + K (calibrator.cpp:157-166), one thread per point (ccg_distort)
  const size_t n = normalized_points.size();
  std::vector<float> xn(n), yn(n), px(n), py(n);
  for (size_t i = 0; i < n; ++i) {
    xn[i] = normalized_points[i].x();
    yn[i] = normalized_points[i].y();
  }
  const float dist[5] = {dist_(0), dist_(1), dist_(2), dist_(3), dist_(4)};
  ccg_ctx* ctx = HostContext();
  const int rc = ccg_distort(ctx, xn.data(), yn.data(), (int64_t)n, K_.m, dist, px.data(), py.data());
  if (rc != CCG_OK)
    throw std::runtime_error(std::string("Calibrator::Distort: ") + ccg_status_string(rc) + " " + ccg_last_error(ctx));
  Points2D out;
  out.reserve(n);
  for (size_t i = 0; i < n; ++i) out.emplace_back(px[i], py[i]);
  return out;
}

}  // namespace calibrator
