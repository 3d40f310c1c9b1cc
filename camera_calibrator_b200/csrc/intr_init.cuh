// Closed-form initialisation of Calibrator::Estimate on the device (included by intr.cu).
//
// Replaces the serial host loop of reference src/calibrator.cpp:52-65:
//   per view   EstimateHomography (DLT)            reference src/geometry.cpp:70-105
//   once       EstimateKFromHomographies (Zhang)   reference src/geometry.cpp:123-177
//   per view   RecoverExtrinsics + FixRotationMatrix + Quaternionf(R)   src/geometry.cpp:179-203
// The reference takes a (2P x 9) and a ((2V+1) x 6) double SVD; at 100,000 views x 400 points that
// is 1e5 serial SVDs of 800 x 9.  Here every view is one warp: it forms the 9 x 9 normal matrix of
// the HARTLEY-NORMALISED DLT system from four 3 x 3 moment matrices (24 accumulators per lane) and
// extracts the null vector by inverse iteration; Zhang's 6 x 6 normal matrix is accumulated in
// pixel coordinates scaled to O(1) and treated the same way.  Normalising is what makes the normal
// equations usable (the raw systems have condition numbers whose squares exceed 1/eps); it also
// weights the algebraic errors differently from the reference's raw SVD, so the STARTING POINT
// differs from the reference's at the noise level (K within a fraction of a percent).  The LM
// solve that follows lands on the same minimum - that is what Estimate returns and what the
// parity tests compare.  The planar mirror ambiguity (sign of the DLT null vector, left open by
// the reference) is resolved towards t_z > 0.
#pragma once

namespace ccg {

// Cholesky-based inverse iteration for the eigenvector of the smallest eigenvalue of an N x N SPD
// matrix (full storage, destroyed).  Returns false when the factorisation breaks down.
template <int N>
__device__ inline bool smallest_eigenvector(double (&M)[N][N], double (&x)[N]) {
  double tr = 0.0;
  for (int i = 0; i < N; ++i) tr += M[i][i];
  if (!(tr > 0.0) || !isfinite(tr)) return false;
  const double shift = 1e-13 * tr;  // keeps exact (noise-free) data factorisable
  for (int i = 0; i < N; ++i) M[i][i] += shift;
  for (int j = 0; j < N; ++j) {
    double dgn = M[j][j];
    for (int k = 0; k < j; ++k) dgn -= M[j][k] * M[j][k];
    if (!(dgn > 0.0)) return false;
    dgn = sqrt(dgn);
    M[j][j] = dgn;
    for (int i = j + 1; i < N; ++i) {
      double s = M[i][j];
      for (int k = 0; k < j; ++k) s -= M[i][k] * M[j][k];
      M[i][j] = s / dgn;
    }
  }
  for (int i = 0; i < N; ++i) x[i] = 1.0 + 0.37 * i;  // generic start (not orthogonal to anything special)
  for (int it = 0; it < 8; ++it) {
    for (int i = 0; i < N; ++i) {
      double s = x[i];
      for (int k = 0; k < i; ++k) s -= M[i][k] * x[k];
      x[i] = s / M[i][i];
    }
    for (int i = N - 1; i >= 0; --i) {
      double s = x[i];
      for (int k = i + 1; k < N; ++k) s -= M[k][i] * x[k];
      x[i] = s / M[i][i];
    }
    double nn = 0.0;
    for (int i = 0; i < N; ++i) nn += x[i] * x[i];
    if (!(nn > 0.0) || !isfinite(nn)) return false;
    const double inv = rsqrt(nn);
    for (int i = 0; i < N; ++i) x[i] *= inv;
  }
  return true;
}

__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) x += __shfl_xor_sync(0xffffffffu, x, s);
  return x;
}

// One warp per view: homography world (X, Y) -> image (u, v), row-major, unit Frobenius norm;
// H[0] = NaN marks a view without a usable homography (fewer than 4 points, degenerate geometry).
constexpr int kInitWarps = 4;
__global__ void __launch_bounds__(kInitWarps * 32) k_init_homography(IntrDev d, int64_t n_views, double* __restrict__ H_out) {
  const int lane = threadIdx.x & 31;
  const int64_t v = (int64_t)blockIdx.x * kInitWarps + (threadIdx.x >> 5);
  if (v >= n_views) return;
  const int64_t o0 = d.view_off[v], o1 = d.view_off[v + 1];
  const double n = (double)(o1 - o0);
  double* H = H_out + v * 9;
  if (o1 - o0 < 4) {
    if (lane == 0) H[0] = nan("");
    return;
  }
  // ---- Hartley normalisation of both point sets: centroid to the origin, mean distance sqrt(2)
  double sx = 0.0, sy = 0.0, su = 0.0, sv = 0.0;
  for (int64_t o = o0 + lane; o < o1; o += 32) {
    sx += (double)d.X[o];
    sy += (double)d.Y[o];
    su += (double)d.u[o];
    sv += (double)d.v[o];
  }
  const double cx = warp_sum(sx) / n, cy = warp_sum(sy) / n, cu = warp_sum(su) / n, cv = warp_sum(sv) / n;
  double d1 = 0.0, d2 = 0.0;
  for (int64_t o = o0 + lane; o < o1; o += 32) {
    const double a = (double)d.X[o] - cx, b = (double)d.Y[o] - cy, p = (double)d.u[o] - cu, q = (double)d.v[o] - cv;
    d1 += sqrt(a * a + b * b);
    d2 += sqrt(p * p + q * q);
  }
  d1 = warp_sum(d1) / n;
  d2 = warp_sum(d2) / n;
  if (!(d1 > 0.0) || !(d2 > 0.0)) {
    if (lane == 0) H[0] = nan("");
    return;
  }
  const double s1 = 1.4142135623730951 / d1, s2 = 1.4142135623730951 / d2;
  // ---- moments: with p = (a, b, 1) the normal matrix of the DLT rows
  //   [0 0 0 | -p | y p]  and  [p | 0 0 0 | -x p]   is
  //   [ M0   0    -Mx ]
  //   [ 0    M0   -My ]      M0 = sum p p', Mx = sum x p p', My = sum y p p', Mr = sum (x^2 + y^2) p p'
  //   [-Mx  -My    Mr ]
  double m[24];
#pragma unroll
  for (int i = 0; i < 24; ++i) m[i] = 0.0;
  for (int64_t o = o0 + lane; o < o1; o += 32) {
    const double a = ((double)d.X[o] - cx) * s1, b = ((double)d.Y[o] - cy) * s1;
    const double x = ((double)d.u[o] - cu) * s2, y = ((double)d.v[o] - cv) * s2;
    const double pp[6] = {a * a, a * b, a, b * b, b, 1.0};
    const double r = x * x + y * y;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      m[i] += pp[i];
      m[6 + i] = fma(x, pp[i], m[6 + i]);
      m[12 + i] = fma(y, pp[i], m[12 + i]);
      m[18 + i] = fma(r, pp[i], m[18 + i]);
    }
  }
#pragma unroll
  for (int i = 0; i < 24; ++i) m[i] = warp_sum(m[i]);
  if (lane != 0) return;
  auto sym = [&](int base, int i, int j) {  // packed upper triangle of a 3 x 3: 00 01 02 11 12 22
    const int a = i < j ? i : j, b = i < j ? j : i;
    return m[base + (a == 0 ? b : (a == 1 ? 2 + b : 5))];
  };
  double M[9][9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      M[i][j] = sym(0, i, j);
      M[3 + i][3 + j] = sym(0, i, j);
      M[i][3 + j] = M[3 + i][j] = 0.0;
      M[i][6 + j] = M[6 + j][i] = -sym(6, i, j);
      M[3 + i][6 + j] = M[6 + j][3 + i] = -sym(12, i, j);
      M[6 + i][6 + j] = sym(18, i, j);
    }
  double h[9];
  if (!smallest_eigenvector<9>(M, h)) {
    H[0] = nan("");
    return;
  }
  // the null vector is ordered like the reference's unknown (geometry.cpp:83-96): rows of the
  // homography that maps (a, b, 1) to (x, y, 1).  Denormalise: H = T2^-1 Hn T1.
  const double Hn[3][3] = {{h[0], h[1], h[2]}, {h[3], h[4], h[5]}, {h[6], h[7], h[8]}};
  double A[3][3];  // Hn T1, T1 = [s1 0 -s1 cx; 0 s1 -s1 cy; 0 0 1]
  for (int i = 0; i < 3; ++i) {
    A[i][0] = Hn[i][0] * s1;
    A[i][1] = Hn[i][1] * s1;
    A[i][2] = Hn[i][2] - s1 * (Hn[i][0] * cx + Hn[i][1] * cy);
  }
  double B[3][3];  // T2^-1 A, T2^-1 = [1/s2 0 cu; 0 1/s2 cv; 0 0 1]
  for (int j = 0; j < 3; ++j) {
    B[0][j] = A[0][j] / s2 + cu * A[2][j];
    B[1][j] = A[1][j] / s2 + cv * A[2][j];
    B[2][j] = A[2][j];
  }
  double nn = 0.0;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) nn += B[i][j] * B[i][j];
  const double inv = rsqrt(nn);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) H[i * 3 + j] = B[i][j] * inv;
}

// One CTA per problem: Zhang's closed form with zero skew.  Image coordinates are scaled by
// 1 / meas_scale so that the entries of K (and of B = K^-T K^-1) are O(1) while the 6 x 6 normal
// matrix is accumulated; each homography enters with unit norm.  Writes intr0 = (fx fy px py 0 0 0 0 0)
// and status (0 ok, CCG_ERR_NOT_FINITE when the closed form has no real solution).
__global__ void __launch_bounds__(256) k_init_zhang(const double* __restrict__ H_all, const int64_t* __restrict__ problem_off,
                                                   const double* __restrict__ meas_scale, double* __restrict__ intr0,
                                                   int* __restrict__ status) {
  __shared__ double s_acc[8][22];
  const int b = blockIdx.x;
  const double sc = meas_scale[b] > 0.0 ? 1.0 / meas_scale[b] : 1.0;
  double acc[22];  // 21 upper-triangle entries + number of usable views
#pragma unroll
  for (int i = 0; i < 22; ++i) acc[i] = 0.0;
  for (int64_t v = problem_off[b] + threadIdx.x; v < problem_off[b + 1]; v += blockDim.x) {
    const double* Hp = H_all + v * 9;
    if (!isfinite(Hp[0])) continue;
    double H[9];
    double nn = 0.0;
    for (int k = 0; k < 9; ++k) {
      H[k] = Hp[k] * (k < 6 ? sc : 1.0);  // diag(sc, sc, 1) H
      nn += H[k] * H[k];
    }
    const double inv = rsqrt(nn);
    for (int k = 0; k < 9; ++k) H[k] *= inv;
    auto vij = [&](int i, int j, double* out) {
      const double a0 = H[i], a1 = H[3 + i], a2 = H[6 + i], b0 = H[j], b1 = H[3 + j], b2 = H[6 + j];
      out[0] = a0 * b0;
      out[1] = a0 * b1 + a1 * b0;
      out[2] = a1 * b1;
      out[3] = a2 * b0 + a0 * b2;
      out[4] = a2 * b1 + a1 * b2;
      out[5] = a2 * b2;
    };
    double r0[6], v00[6], v11[6], r1[6];
    vij(0, 1, r0);
    vij(0, 0, v00);
    vij(1, 1, v11);
    for (int k = 0; k < 6; ++k) r1[k] = v00[k] - v11[k];
    int idx = 0;
    for (int i = 0; i < 6; ++i)
      for (int j = i; j < 6; ++j) acc[idx++] += r0[i] * r0[j] + r1[i] * r1[j];
    acc[21] += 1.0;
  }
#pragma unroll
  for (int i = 0; i < 22; ++i) acc[i] = warp_sum(acc[i]);
  if ((threadIdx.x & 31) == 0)
    for (int i = 0; i < 22; ++i) s_acc[threadIdx.x >> 5][i] = acc[i];
  __syncthreads();
  if (threadIdx.x != 0) return;
  double tot[22];
  for (int i = 0; i < 22; ++i) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += s_acc[w][i];
    tot[i] = s;
  }
  double* out = intr0 + (size_t)b * 9;
  for (int k = 0; k < 9; ++k) out[k] = 0.0;
  if (tot[21] < 3.0) {
    status[b] = CCG_ERR_BAD_ARG;  // Zhang needs at least three homographies (geometry.cpp:124)
    return;
  }
  double M[6][6];
  int idx = 0;
  for (int i = 0; i < 6; ++i)
    for (int j = i; j < 6; ++j) {
      M[i][j] = M[j][i] = tot[idx];
      ++idx;
    }
  M[1][1] += tot[21] * tot[21];  // the soft zero-skew row (0, n, 0, 0, 0, 0) of geometry.cpp:147
  double bv[6];
  if (!smallest_eigenvector<6>(M, bv)) {
    status[b] = CCG_ERR_NOT_FINITE;
    return;
  }
  const double B11 = bv[0], B12 = bv[1], B22 = bv[2], B13 = bv[3], B23 = bv[4], B33 = bv[5];
  const double den = B11 * B22 - B12 * B12;
  const double v0 = (B12 * B13 - B11 * B23) / den;
  const double lam = B33 - (B13 * B13 + v0 * (B12 * B13 - B11 * B23)) / B11;
  const double alpha = sqrt(lam / B11);
  const double beta = sqrt(lam * B11 / den);
  const double u0 = -B13 * alpha * alpha / lam;  // zero skew forced, as in the reference
  if (!isfinite(alpha) || !isfinite(beta) || !isfinite(u0) || !isfinite(v0) || !(alpha > 0.0) || !(beta > 0.0)) {
    status[b] = CCG_ERR_NOT_FINITE;
    return;
  }
  // the reference keeps K as float32 (types.hh:14) before it enters the solver (calibrator.cpp:245)
  out[0] = (double)(float)(alpha / sc);
  out[1] = (double)(float)(beta / sc);
  out[2] = (double)(float)(u0 / sc);
  out[3] = (double)(float)(v0 / sc);
  status[b] = 0;
}

// One thread per view: RecoverExtrinsics with the problem's K, nearest rotation by Newton's polar
// iteration (the U V' of geometry.cpp:197-203), quaternion, pose0.
__global__ void __launch_bounds__(128) k_init_extrinsics(const double* __restrict__ H_all, const int64_t* __restrict__ problem_off,
                                                        int n_problems, const double* __restrict__ intr0,
                                                        double* __restrict__ pose0) {
  const int b = blockIdx.y;
  const int64_t v = problem_off[b] + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= problem_off[b + 1]) return;
  double* pose = pose0 + v * kPose;
  const double* H = H_all + v * 9;
  const double fx = intr0[(size_t)b * 9], fy = intr0[(size_t)b * 9 + 1], px = intr0[(size_t)b * 9 + 2],
               py = intr0[(size_t)b * 9 + 3];
  pose[7] = 0.0;
  if (!isfinite(H[0]) || !(fx > 0.0) || !(fy > 0.0)) {  // no homography: a neutral pose in front of the camera
    pose[0] = 1.0; pose[1] = pose[2] = pose[3] = 0.0;
    pose[4] = pose[5] = 0.0; pose[6] = 1.0;
    return;
  }
  // K^-1 h for the three columns (zero skew)
  double c[3][3];
  for (int j = 0; j < 3; ++j) {
    c[j][0] = (H[j] - px * H[6 + j]) / fx;
    c[j][1] = (H[3 + j] - py * H[6 + j]) / fy;
    c[j][2] = H[6 + j];
  }
  double l = rsqrt(c[0][0] * c[0][0] + c[0][1] * c[0][1] + c[0][2] * c[0][2]);
  if (c[2][2] < 0.0) l = -l;  // mirror ambiguity of the null vector: keep the board in front of the camera
  double R[3][3], t[3];
  for (int i = 0; i < 3; ++i) {
    R[i][0] = l * c[0][i];
    R[i][1] = l * c[1][i];
    t[i] = l * c[2][i];
  }
  R[0][2] = R[1][0] * R[2][1] - R[2][0] * R[1][1];
  R[1][2] = R[2][0] * R[0][1] - R[0][0] * R[2][1];
  R[2][2] = R[0][0] * R[1][1] - R[1][0] * R[0][1];
  // polar decomposition: R <- (R + R^-T) / 2 converges quadratically to U V'
  for (int it = 0; it < 12; ++it) {
    const double c00 = R[1][1] * R[2][2] - R[1][2] * R[2][1], c01 = R[1][2] * R[2][0] - R[1][0] * R[2][2],
                 c02 = R[1][0] * R[2][1] - R[1][1] * R[2][0];
    const double det = R[0][0] * c00 + R[0][1] * c01 + R[0][2] * c02;
    if (!(fabs(det) > 1e-300)) break;
    const double id = 1.0 / det;
    double T[3][3];  // R^-T = cofactor(R) / det
    T[0][0] = c00 * id; T[0][1] = c01 * id; T[0][2] = c02 * id;
    T[1][0] = (R[0][2] * R[2][1] - R[0][1] * R[2][2]) * id;
    T[1][1] = (R[0][0] * R[2][2] - R[0][2] * R[2][0]) * id;
    T[1][2] = (R[0][1] * R[2][0] - R[0][0] * R[2][1]) * id;
    T[2][0] = (R[0][1] * R[1][2] - R[0][2] * R[1][1]) * id;
    T[2][1] = (R[0][2] * R[1][0] - R[0][0] * R[1][2]) * id;
    T[2][2] = (R[0][0] * R[1][1] - R[0][1] * R[1][0]) * id;
    double change = 0.0;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        const double nv = 0.5 * (R[i][j] + T[i][j]);
        change = fmax(change, fabs(nv - R[i][j]));
        R[i][j] = nv;
      }
    if (change < 1e-15) break;
  }
  // [ext] Eigen::Quaternionf(Matrix3f): trace test, else the largest diagonal entry
  double q[4];
  const double tr = R[0][0] + R[1][1] + R[2][2];
  if (tr > 0.0) {
    double s = sqrt(tr + 1.0);
    q[0] = 0.5 * s;
    s = 0.5 / s;
    q[1] = (R[2][1] - R[1][2]) * s;
    q[2] = (R[0][2] - R[2][0]) * s;
    q[3] = (R[1][0] - R[0][1]) * s;
  } else {
    int i = 0;
    if (R[1][1] > R[0][0]) i = 1;
    if (R[2][2] > R[i][i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    double s = sqrt(R[i][i] - R[j][j] - R[k][k] + 1.0);
    q[1 + i] = 0.5 * s;
    s = 0.5 / s;
    q[0] = (R[k][j] - R[j][k]) * s;
    q[1 + j] = (R[j][i] + R[i][j]) * s;
    q[1 + k] = (R[k][i] + R[i][k]) * s;
  }
  // the reference hands Quaternionf / Vector3f (float32) to Optimize (calibrator.cpp:58-63)
  for (int k = 0; k < 4; ++k) pose[k] = (double)(float)q[k];
  for (int k = 0; k < 3; ++k) pose[4 + k] = (double)(float)t[k];
}

}  // namespace ccg
