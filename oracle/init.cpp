// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.h header).
//
// Restatement of the closed-form initialisation that Calibrator::Estimate runs
// before the LM solve (reference src/calibrator.cpp:47-65):
//   EstimateHomography (DLT, 2n x 9 double SVD) ... reference src/geometry.cpp:70-105
//   EstimateKFromHomographies (Zhang, zero skew) .. reference src/geometry.cpp:123-177
//   RecoverExtrinsics ............................. reference src/geometry.cpp:179-195
//   FixRotationMatrix (U V^T) ..................... reference src/geometry.cpp:197-203
//   Eigen::Quaternionf(R) ......................... [ext] Eigen rotation-matrix -> quaternion
// The sign of an SVD null vector is implementation defined (in Eigen too), so H
// and therefore (R, t) are defined up to the planar mirror ambiguity the
// reference leaves unresolved; K is sign-invariant.
#include <cmath>
#include <vector>

#include "oracle.h"
#include "svd.hh"

namespace {

void homography(const float* x1, const float* y1, const float* x2, const float* y2, int64_t n,
                float* H) {
  std::vector<double> A((size_t)2 * n * 9, 0.0);
  for (int64_t i = 0; i < n; ++i) {
    double* r0 = &A[(size_t)(2 * i) * 9];
    double* r1 = r0 + 9;
    // products are formed in float32 before widening, as in the reference
    r0[3] = -x1[i];
    r0[4] = -y1[i];
    r0[5] = -1.0;
    r0[6] = (double)(x1[i] * y2[i]);
    r0[7] = (double)(y1[i] * y2[i]);
    r0[8] = y2[i];
    r1[0] = x1[i];
    r1[1] = y1[i];
    r1[2] = 1.0;
    r1[6] = (double)(-x1[i] * x2[i]);
    r1[7] = (double)(-y1[i] * x2[i]);
    r1[8] = -x2[i];
  }
  const std::vector<double> h = oracle::null_vector(A, (int)(2 * n), 9);
  for (int k = 0; k < 9; ++k) H[k] = (float)h[k];
}

void inv3f(const float* A, float* B) {
  const float c00 = A[4] * A[8] - A[5] * A[7];
  const float c01 = A[5] * A[6] - A[3] * A[8];
  const float c02 = A[3] * A[7] - A[4] * A[6];
  const float id = 1.0f / (A[0] * c00 + A[1] * c01 + A[2] * c02);
  B[0] = c00 * id;
  B[1] = (A[2] * A[7] - A[1] * A[8]) * id;
  B[2] = (A[1] * A[5] - A[2] * A[4]) * id;
  B[3] = c01 * id;
  B[4] = (A[0] * A[8] - A[2] * A[6]) * id;
  B[5] = (A[2] * A[3] - A[0] * A[5]) * id;
  B[6] = c02 * id;
  B[7] = (A[1] * A[6] - A[0] * A[7]) * id;
  B[8] = (A[0] * A[4] - A[1] * A[3]) * id;
}

void k_from_homographies(const float* Hs, int64_t n, float* K) {
  std::vector<double> A((size_t)(2 * n + 1) * 6, 0.0);
  auto vij = [](const double* H, int i, int j, double* out) {
    // H row-major: column i entries are H[0*3+i], H[1*3+i], H[2*3+i]
    const double a0 = H[i], a1 = H[3 + i], a2 = H[6 + i];
    const double b0 = H[j], b1 = H[3 + j], b2 = H[6 + j];
    out[0] = a0 * b0;
    out[1] = a0 * b1 + a1 * b0;
    out[2] = a1 * b1;
    out[3] = a2 * b0 + a0 * b2;
    out[4] = a2 * b1 + a1 * b2;
    out[5] = a2 * b2;
  };
  for (int64_t i = 0; i < n; ++i) {
    double H[9];
    for (int k = 0; k < 9; ++k) H[k] = Hs[(size_t)i * 9 + k];
    double v01[6], v00[6], v11[6];
    vij(H, 0, 1, v01);
    vij(H, 0, 0, v00);
    vij(H, 1, 1, v11);
    for (int k = 0; k < 6; ++k) {
      A[(size_t)(2 * i) * 6 + k] = v01[k];
      A[(size_t)(2 * i + 1) * 6 + k] = v00[k] - v11[k];
    }
  }
  A[(size_t)(2 * n) * 6 + 1] = (double)n;  // soft zero-skew row, weighted by the view count
  const std::vector<double> b = oracle::null_vector(A, (int)(2 * n + 1), 6);
  const double B11 = b[0], B12 = b[1], B22 = b[2], B13 = b[3], B23 = b[4], B33 = b[5];
  const double v0 = (B12 * B13 - B11 * B23) / (B11 * B22 - B12 * B12);
  const double l = B33 - (B13 * B13 + v0 * (B12 * B13 - B11 * B23)) / B11;
  const double alpha = std::sqrt(l / B11);
  const double beta = std::sqrt(l * B11 / (B11 * B22 - B12 * B12));
  const double skew = 0.0;
  const double u0 = skew * v0 / beta - B13 * alpha * alpha / l;
  K[0] = (float)alpha; K[1] = (float)skew; K[2] = (float)u0;
  K[3] = 0.0f;         K[4] = (float)beta; K[5] = (float)v0;
  K[6] = 0.0f;         K[7] = 0.0f;        K[8] = 1.0f;
}

// nearest rotation U V^T of a 3x3 (float in, float out)
void fix_rotation(const float* R, float* out) {
  std::vector<double> A(9), V, s, U;
  for (int k = 0; k < 9; ++k) A[k] = R[k];
  oracle::jacobi_svd(A, 3, 3, &V, &s, &U);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double acc = 0.0;
      for (int k = 0; k < 3; ++k) acc += U[(size_t)i * 3 + k] * V[(size_t)j * 3 + k];
      out[i * 3 + j] = (float)acc;
    }
}

void recover_extrinsics(const float* Kinv, const float* H, float* R, float* t) {
  auto col = [&](int j, float* o) {
    for (int i = 0; i < 3; ++i) o[i] = Kinv[i * 3] * H[j] + Kinv[i * 3 + 1] * H[3 + j] + Kinv[i * 3 + 2] * H[6 + j];
  };
  float h0[3], h1[3], h2[3];
  col(0, h0);
  col(1, h1);
  col(2, h2);
  const float l = 1.0f / std::sqrt(h0[0] * h0[0] + h0[1] * h0[1] + h0[2] * h0[2]);
  const float r0[3] = {l * h0[0], l * h0[1], l * h0[2]};
  const float r1[3] = {l * h1[0], l * h1[1], l * h1[2]};
  const float r2[3] = {r0[1] * r1[2] - r0[2] * r1[1], r0[2] * r1[0] - r0[0] * r1[2], r0[0] * r1[1] - r0[1] * r1[0]};
  for (int i = 0; i < 3; ++i) t[i] = l * h2[i];
  const float Rraw[9] = {r0[0], r1[0], r2[0], r0[1], r1[1], r2[1], r0[2], r1[2], r2[2]};
  fix_rotation(Rraw, R);
}

// [ext] Eigen::Quaternionf(Matrix3f): trace-based branch selection
void quat_from_rotation(const float* m, float* q /* w,x,y,z */) {
  float tr = m[0] + m[4] + m[8];
  if (tr > 0.0f) {
    float s = std::sqrt(tr + 1.0f);
    q[0] = 0.5f * s;
    s = 0.5f / s;
    q[1] = (m[7] - m[5]) * s;
    q[2] = (m[2] - m[6]) * s;
    q[3] = (m[3] - m[1]) * s;
  } else {
    int i = 0;
    if (m[4] > m[0]) i = 1;
    if (m[8] > m[i * 3 + i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    float s = std::sqrt(m[i * 3 + i] - m[j * 3 + j] - m[k * 3 + k] + 1.0f);
    float xyz[3];
    xyz[i] = 0.5f * s;
    s = 0.5f / s;
    q[0] = (m[k * 3 + j] - m[j * 3 + k]) * s;
    xyz[j] = (m[j * 3 + i] + m[i * 3 + j]) * s;
    xyz[k] = (m[k * 3 + i] + m[i * 3 + k]) * s;
    q[1] = xyz[0];
    q[2] = xyz[1];
    q[3] = xyz[2];
  }
}

}  // namespace

extern "C" int oracle_homography(const float* x1, const float* y1, const float* x2, const float* y2,
                                 int64_t n, float* H_out) {
  homography(x1, y1, x2, y2, n, H_out);
  return 0;
}

extern "C" int oracle_estimate_init(const float* u, const float* v, const float* X, const float* Y,
                                    const float* Z, const int64_t* view_offsets, int64_t n_views,
                                    float* H_out, float* K_out, float* q_out, float* t_out) {
  (void)Z;  // planar target: the DLT reads x and y of the world point only (geometry.cpp:83-96)
  std::vector<float> Hs((size_t)n_views * 9);
  for (int64_t i = 0; i < n_views; ++i) {
    const int64_t o = view_offsets[i], n = view_offsets[i + 1] - o;
    homography(X + o, Y + o, u + o, v + o, n, &Hs[(size_t)i * 9]);
  }
  float K[9], Kinv[9];
  k_from_homographies(Hs.data(), n_views, K);
  inv3f(K, Kinv);
  for (int64_t i = 0; i < n_views; ++i) {
    float R[9], t[3], q[4];
    recover_extrinsics(Kinv, &Hs[(size_t)i * 9], R, t);
    quat_from_rotation(R, q);
    if (q_out) for (int k = 0; k < 4; ++k) q_out[(size_t)i * 4 + k] = q[k];
    if (t_out) for (int k = 0; k < 3; ++k) t_out[(size_t)i * 3 + k] = t[k];
  }
  if (H_out) for (size_t k = 0; k < Hs.size(); ++k) H_out[k] = Hs[k];
  if (K_out) for (int k = 0; k < 9; ++k) K_out[k] = K[k];
  return 0;
}
