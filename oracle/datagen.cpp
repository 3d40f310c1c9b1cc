// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.h header).
//
// Restatement of the reference's synthetic-input generator, the definition of
// "identical data_generator synthetic inputs" for parity:
//   DataGenerator ctor / distributions ...... reference src/data_generator.cpp:34-50, data_generator.hh:36-47
//   GetRandomPlane .......................... reference src/data_generator.cpp:52-75
//   GetDistortedPointsPlanar ................ reference src/data_generator.cpp:77-124
//   GetRandom3DPointVisibleToCamera ......... reference src/data_generator.cpp:126-146
//   ProjectToCameraAndDistort ............... reference src/data_generator.cpp:10-32
//   EstimatePlaneFinite / RotationMatrixFromPlane / ProjectToPlane
//                                             reference src/geometry.cpp:6-68
// The random stream is std::mt19937 + std::uniform_real_distribution<float>
// from this image's libstdc++, i.e. the very implementation the reference would
// link on Linux.  cv::projectPoints (zero rvec/tvec) is restated in double with
// the 5-coefficient OpenCV model and rounded to float32 like OpenCV's output
// [ext: OpenCV is not vendored; formula checked against cv2 4.13 in tests].
// Eigen's pseudoInverse of K is replaced by the plain 3x3 inverse (K is
// invertible), evaluated in float32.
#include <cmath>
#include <random>
#include <vector>

#include "oracle.h"

namespace {

struct V3 {
  float x, y, z;
};
inline V3 mulM(const float* M, const V3& p) {
  return {M[0] * p.x + M[1] * p.y + M[2] * p.z, M[3] * p.x + M[4] * p.y + M[5] * p.z,
          M[6] * p.x + M[7] * p.y + M[8] * p.z};
}
inline V3 normalized(const V3& p) {
  const float n = std::sqrt(p.x * p.x + p.y * p.y + p.z * p.z);
  return {p.x / n, p.y / n, p.z / n};
}
inline V3 cross(const V3& a, const V3& b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline V3 scale(const V3& p, float s) { return {p.x * s, p.y * s, p.z * s}; }

// cofactor inverse, float32
void inv3(const float* A, float* B) {
  const float c00 = A[4] * A[8] - A[5] * A[7];
  const float c01 = A[5] * A[6] - A[3] * A[8];
  const float c02 = A[3] * A[7] - A[4] * A[6];
  const float det = A[0] * c00 + A[1] * c01 + A[2] * c02;
  const float id = 1.0f / det;
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

}  // namespace

extern "C" int oracle_datagen_planar(uint32_t seed, int32_t width, int32_t height, const float* K,
                                     const float* dist, float noise_px, int32_t n_points,
                                     int32_t n_views, float* u, float* v, float* X, float* Y,
                                     float* Z, float* plane_out, int64_t* n_rejected) {
  std::mt19937 gen{seed};
  std::uniform_real_distribution<float> rand_w(0.0f, width - 1.0f);
  std::uniform_real_distribution<float> rand_h(0.0f, height - 1.0f);
  std::uniform_real_distribution<float> rand_dist(0.2f, 1.0f);
  std::uniform_real_distribution<float> rand_pixel(-noise_px, noise_px);
  float Kinv[9];
  inv3(K, Kinv);
  const double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
  const double k1 = dist[0], k2 = dist[1], p1 = dist[2], p2 = dist[3], k3 = dist[4];
  int64_t rejected = 0;
  for (int32_t iv = 0; iv < n_views; ++iv) {
    // random plane through three frustum-corner rays
    V3 c = {0.0f, 0.0f, 1.0f};
    V3 a1 = scale(normalized(mulM(Kinv, c)), rand_dist(gen));
    c.x = (float)(width - 1);
    V3 a2 = scale(normalized(mulM(Kinv, c)), rand_dist(gen));
    c.y = (float)(height - 1);
    V3 a3 = scale(normalized(mulM(Kinv, c)), rand_dist(gen));
    const float A[9] = {a1.x, a1.y, a1.z, a2.x, a2.y, a2.z, a3.x, a3.y, a3.z};
    float Ai[9];
    inv3(A, Ai);
    const V3 pn = mulM(Ai, V3{1.0f, 1.0f, 1.0f});
    const float plane[4] = {pn.x, pn.y, pn.z, -1.0f};
    if (plane_out) {
      for (int k = 0; k < 4; ++k) plane_out[(size_t)iv * 4 + k] = plane[k];
    }
    // rotation taking the plane normal to +z (rows v1, v2, normal)
    const V3 nrm = normalized(pn);
    const V3 v1 = normalized(cross(nrm, V3{0.0f, 0.0f, 1.0f}));
    const V3 v2 = normalized(cross(nrm, v1));
    const float R[9] = {v1.x, v1.y, v1.z, v2.x, v2.y, v2.z, nrm.x, nrm.y, nrm.z};
    int32_t got = 0;
    while (got < n_points) {
      V3 px;
      px.x = rand_w(gen);
      px.y = rand_h(gen);
      px.z = 1.0f;
      V3 p3 = normalized(mulM(Kinv, px));
      const float depth = rand_dist(gen);
      p3 = scale(p3, depth);
      // project onto the plane along the point's own ray, in double
      const double pd[3] = {p3.x, p3.y, p3.z};
      const double nd[3] = {plane[0], plane[1], plane[2]};
      const double tt = (pd[0] * nd[0] + pd[1] * nd[1] + pd[2] * nd[2] + (double)plane[3]) /
                        (pd[0] * nd[0] + pd[1] * nd[1] + pd[2] * nd[2]);
      const V3 pp = {(float)(pd[0] - pd[0] * tt), (float)(pd[1] - pd[1] * tt), (float)(pd[2] - pd[2] * tt)};
      const V3 pr = mulM(R, pp);
      // OpenCV projectPoints, identity pose, double inside, float out
      const double x = (double)pp.x / (double)pp.z, y = (double)pp.y / (double)pp.z;
      const double r2 = x * x + y * y, r4 = r2 * r2, r6 = r4 * r2;
      const double cd = 1.0 + k1 * r2 + k2 * r4 + k3 * r6;
      const double xd = x * cd + p1 * (2.0 * x * y) + p2 * (r2 + 2.0 * x * x);
      const double yd = y * cd + p1 * (r2 + 2.0 * y * y) + p2 * (2.0 * x * y);
      const float iu = (float)(xd * fx + cx), ivv = (float)(yd * fy + cy);
      const float ud = iu + rand_pixel(gen);
      const float vd = ivv + rand_pixel(gen);
      if (ud < 0.0f || ud >= (float)(width - 1) || vd < 0.0f || vd >= (float)(height - 1)) {
        ++rejected;
        continue;
      }
      const size_t o = (size_t)iv * n_points + got;
      u[o] = ud;
      v[o] = vd;
      X[o] = pr.x;
      Y[o] = pr.y;
      Z[o] = 0.0f;
      ++got;
    }
  }
  if (n_rejected) *n_rejected = rejected;
  return 0;
}
