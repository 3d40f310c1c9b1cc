// ORACLE — TEST INFRASTRUCTURE ONLY (see jet.hh header).
//
// CPU restatement of the reference's two residual models, templated on the
// scalar type so the same code runs on double (cost only) and on Jet<N>
// (cost + Jacobian), exactly as the reference's functors do under Ceres
// autodiff.
//
//   lens model ............ reference src/calibrator.cpp:70-83  (DistortNormalized)
//   pixel mapping ......... reference src/calibrator.cpp:85-95  (DistortPixels)
//   intrinsic block order . reference src/calibrator.cpp:168-179
//   single-camera residual  reference src/calibrator.cpp:183-219
//   rig residual .......... reference src/extrinsics_calibrator.cpp:51-84
//   quaternion rotation ... [ext] ceres/rotation.h QuaternionRotatePoint:
//                           normalises q, then rotates; storage (w,x,y,z)
#pragma once
#include "jet.hh"

namespace oracle {

// Index of each entry of the 9-vector of intrinsics (calibrator.cpp:168-179).
enum IntrSlot { I_FX = 0, I_FY, I_PX, I_PY, I_K1, I_K2, I_P1, I_P2, I_K3, I_COUNT };

// p + 2w (v x p) + 2 v x (v x p) with (w, v) = q/|q|   [ext ceres/rotation.h]
template <typename T>
inline void rotate_by_quaternion(const T q[4], const T p[3], T out[3]) {
  using std::sqrt;
  const T n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const T w = q[0] / n, a = q[1] / n, b = q[2] / n, c = q[3] / n;
  // c1 = v x p, doubled
  T c1[3] = {b * p[2] - c * p[1], c * p[0] - a * p[2], a * p[1] - b * p[0]};
  c1[0] = c1[0] + c1[0];
  c1[1] = c1[1] + c1[1];
  c1[2] = c1[2] + c1[2];
  out[0] = p[0] + w * c1[0] + (b * c1[2] - c * c1[1]);
  out[1] = p[1] + w * c1[1] + (c * c1[0] - a * c1[2]);
  out[2] = p[2] + w * c1[2] + (a * c1[1] - b * c1[0]);
}

// Brown-Conrady radial (k1,k2,k3) + tangential (p1,p2) on normalised coords.
// calibrator.cpp:70-83
template <typename TP, typename T>
inline void lens_distort(const TP& k1, const TP& k2, const TP& k3, const TP& p1, const TP& p2,
                         const T& x, const T& y, T* xd, T* yd) {
  const T r2 = x * x + y * y;
  const T r4 = r2 * r2;
  const T r6 = r4 * r2;
  const T radial = 1.0 + k1 * r2 + k2 * r4 + k3 * r6;
  *xd = x * radial + 2.0 * p1 * x * y + p2 * (r2 + 2.0 * x * x);
  *yd = y * radial + 2.0 * p2 * x * y + p1 * (r2 + 2.0 * y * y);
}

// Single-camera reprojection residual (calibrator.cpp:183-219): world point ->
// camera frame by (q, t), pinhole division, lens model, pixels (no skew),
// minus the measured pixel.
template <typename T>
inline void residual_single(const T* intr, const T* q, const T* t, const T* X, double mx, double my,
                            T res[2]) {
  T pc[3];
  rotate_by_quaternion(q, X, pc);
  pc[0] = pc[0] + t[0];
  pc[1] = pc[1] + t[1];
  pc[2] = pc[2] + t[2];
  const T xn = pc[0] / pc[2];
  const T yn = pc[1] / pc[2];
  T xd, yd;
  lens_distort(intr[I_K1], intr[I_K2], intr[I_K3], intr[I_P1], intr[I_P2], xn, yn, &xd, &yd);
  res[0] = intr[I_FX] * xd + intr[I_PX] - mx;
  res[1] = intr[I_FY] * yd + intr[I_PY] - my;
}

// Rig residual (extrinsics_calibrator.cpp:51-84): world -> rig by the frame
// pose, rig -> camera by the camera pose, pinhole division; the measurement
// is already in normalised (undistorted) image coordinates.
template <typename T>
inline void residual_rig(const T* qf, const T* tf, const T* qc, const T* tc, const T* X, double mu,
                         double mv, T res[2]) {
  T pr[3], pc[3];
  rotate_by_quaternion(qf, X, pr);
  pr[0] = pr[0] + tf[0];
  pr[1] = pr[1] + tf[1];
  pr[2] = pr[2] + tf[2];
  rotate_by_quaternion(qc, pr, pc);
  pc[0] = pc[0] + tc[0];
  pc[1] = pc[1] + tc[1];
  pc[2] = pc[2] + tc[2];
  res[0] = pc[0] / pc[2] - mu;
  res[1] = pc[1] / pc[2] - mv;
}

// [ext] ceres::QuaternionManifold::PlusJacobian at x (4x3, row-major):
// d(exp_q(delta) * x)/d(delta) at delta = 0, delta being the HALF rotation vector.
inline void quaternion_plus_jacobian(const double x[4], double J[12]) {
  J[0] = -x[1]; J[1]  = -x[2]; J[2]  = -x[3];
  J[3] =  x[0]; J[4]  =  x[3]; J[5]  = -x[2];
  J[6] = -x[3]; J[7]  =  x[0]; J[8]  =  x[1];
  J[9] =  x[2]; J[10] = -x[1]; J[11] =  x[0];
}

// [ext] ceres::QuaternionManifold::Plus: x_plus = q_delta * x (left product),
// q_delta = (cos|d|, sin|d|/|d| d).
inline void quaternion_plus(const double x[4], const double d[3], double out[4]) {
  const double n2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
  double qd[4];
  if (n2 > 0.0) {
    const double n = std::sqrt(n2);
    const double s = std::sin(n) / n;
    qd[0] = std::cos(n); qd[1] = s * d[0]; qd[2] = s * d[1]; qd[3] = s * d[2];
  } else {
    qd[0] = 1.0; qd[1] = d[0]; qd[2] = d[1]; qd[3] = d[2];
  }
  out[0] = qd[0] * x[0] - qd[1] * x[1] - qd[2] * x[2] - qd[3] * x[3];
  out[1] = qd[0] * x[1] + qd[1] * x[0] + qd[2] * x[3] - qd[3] * x[2];
  out[2] = qd[0] * x[2] - qd[1] * x[3] + qd[2] * x[0] + qd[3] * x[1];
  out[3] = qd[0] * x[3] + qd[1] * x[2] - qd[2] * x[1] + qd[3] * x[0];
}

// [ext] ceres::HuberLoss(a): rho(s) = s for s <= a^2, else 2 a sqrt(s) - a^2.
// Returns rho and rho' (rho'' <= 0 always, so Ceres' corrector scales the
// residual and the Jacobian by sqrt(rho') only).
inline void huber(double a, double s, double* rho, double* rho1) {
  const double b = a * a;
  if (s > b) {
    const double r = std::sqrt(s);
    *rho = 2.0 * a * r - b;
    *rho1 = a / r;
  } else {
    *rho = s;
    *rho1 = 1.0;
  }
}

}  // namespace oracle
