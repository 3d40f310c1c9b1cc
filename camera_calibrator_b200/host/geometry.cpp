// Host-side initialisation geometry (reference src/geometry.cpp:6-203): DLT
// homography, Zhang's closed-form K, pose from homography, nearest rotation.
// The reference leans on Eigen::JacobiSVD; here the SVD is a small one-sided
// Jacobi (Hestenes) routine in double.  Singular vectors carry Eigen's sign
// ambiguity, i.e. the planar mirror ambiguity the reference leaves unresolved.
#include "geometry.hh"

#include <algorithm>
#include <cmath>
#include <stdexcept>

namespace calibrator {
namespace {

// Column-major dense matrix with just enough behaviour for the routines below.
struct Mat {
  int rows, cols;
  std::vector<double> a;
  Mat(int r, int c) : rows(r), cols(c), a((size_t)r * c, 0.0) {}
  double& at(int r, int c) { return a[(size_t)c * rows + r]; }
  double at(int r, int c) const { return a[(size_t)c * rows + r]; }
  double* colp(int c) { return &a[(size_t)c * rows]; }
};

struct Svd {
  Mat U, V;
  std::vector<double> s;
  Svd(int m, int n) : U(m, n), V(n, n), s(n, 0.0) {}
  int smallest() const { return (int)(std::min_element(s.begin(), s.end()) - s.begin()); }
};

// One-sided Jacobi: rotate column pairs of A until they are mutually orthogonal.
Svd jacobi_svd(Mat A) {
  const int m = A.rows, n = A.cols;
  Svd out(m, n);
  for (int i = 0; i < n; ++i) out.V.at(i, i) = 1.0;
  for (int sweep = 0; sweep < 64; ++sweep) {
    bool rotated = false;
    for (int p = 0; p + 1 < n; ++p) {
      for (int q = p + 1; q < n; ++q) {
        double* cp = A.colp(p);
        double* cq = A.colp(q);
        double alpha = 0.0, beta = 0.0, gamma = 0.0;
        for (int i = 0; i < m; ++i) {
          alpha += cp[i] * cp[i];
          beta += cq[i] * cq[i];
          gamma += cp[i] * cq[i];
        }
        if (gamma == 0.0 || std::fabs(gamma) <= 1e-16 * std::sqrt(alpha * beta)) continue;
        rotated = true;
        const double zeta = (beta - alpha) / (2.0 * gamma);
        const double t = std::copysign(1.0, zeta) / (std::fabs(zeta) + std::hypot(1.0, zeta));
        const double c = 1.0 / std::hypot(1.0, t), s = c * t;
        for (int i = 0; i < m; ++i) {
          const double x = cp[i], y = cq[i];
          cp[i] = c * x - s * y;
          cq[i] = s * x + c * y;
        }
        double* vp = out.V.colp(p);
        double* vq = out.V.colp(q);
        for (int i = 0; i < n; ++i) {
          const double x = vp[i], y = vq[i];
          vp[i] = c * x - s * y;
          vq[i] = s * x + c * y;
        }
      }
    }
    if (!rotated) break;
  }
  for (int j = 0; j < n; ++j) {
    double nn = 0.0;
    for (int i = 0; i < m; ++i) nn += A.at(i, j) * A.at(i, j);
    out.s[j] = std::sqrt(nn);
    for (int i = 0; i < m; ++i) out.U.at(i, j) = out.s[j] > 0.0 ? A.at(i, j) / out.s[j] : 0.0;
  }
  return out;
}

template <typename PA, typename PB>
Matrix3 homography_dlt(const PA& p1, const PB& p2) {
  if (p1.size() != p2.size()) throw std::invalid_argument("EstimateHomography: size mismatch");
  const int n = (int)p1.size();
  Mat A(2 * n, 9);
  for (int i = 0; i < n; ++i) {
    const float x1 = p1[i].x(), y1 = p1[i].y(), x2 = p2[i].x(), y2 = p2[i].y();
    const int r = 2 * i;
    A.at(r, 3) = -x1;
    A.at(r, 4) = -y1;
    A.at(r, 5) = -1.0;
    A.at(r, 6) = (double)(x1 * y2);  // float product, as the reference forms it
    A.at(r, 7) = (double)(y1 * y2);
    A.at(r, 8) = y2;
    A.at(r + 1, 0) = x1;
    A.at(r + 1, 1) = y1;
    A.at(r + 1, 2) = 1.0;
    A.at(r + 1, 6) = (double)(-x1 * x2);
    A.at(r + 1, 7) = (double)(-y1 * x2);
    A.at(r + 1, 8) = -x2;
  }
  const Svd svd = jacobi_svd(A);
  const int k = svd.smallest();
  Matrix3 H;
  for (int i = 0; i < 9; ++i) H.m[i] = (float)svd.V.at(i, k);
  return H;
}

}  // namespace

// ---------------------------------------------------------------- planes ----
Plane EstimatePlaneFinite(const Point3D& p1, const Point3D& p2, const Point3D& p3) {
  Matrix3 A;
  for (int c = 0; c < 3; ++c) {
    A(0, c) = p1(c);
    A(1, c) = p2(c);
    A(2, c) = p3(c);
  }
  const Point3D n = A.inverse() * Point3D(1.f, 1.f, 1.f);
  Plane out;
  out(0) = n(0);
  out(1) = n(1);
  out(2) = n(2);
  out(3) = -1.0f;
  return out;
}

Point3D PlaneNormal(const Plane& plane) {
  Point3D n(plane(0), plane(1), plane(2));
  const float l = n.norm();
  return {n(0) / l, n(1) / l, n(2) / l};
}

Matrix3 RotationMatrixFromPlane(const Plane& plane, const Point3D& new_normal) {
  auto unit = [](Point3D p) {
    const float l = p.norm();
    return Point3D(p(0) / l, p(1) / l, p(2) / l);
  };
  const Point3D normal = PlaneNormal(plane);
  const Point3D v1 = unit(normal.cross(unit(new_normal)));
  const Point3D v2 = unit(normal.cross(v1));
  Matrix3 R;
  for (int c = 0; c < 3; ++c) {
    R(0, c) = v1(c);
    R(1, c) = v2(c);
    R(2, c) = normal(c);
  }
  return R;
}

Point3D ProjectToPlane(const Plane& plane, const Point3D& p, const std::optional<Point3D>& direction) {
  if (!direction) {
    const float pn[3] = {plane(0), plane(1), plane(2)};
    const float t = (p(0) * pn[0] + p(1) * pn[1] + p(2) * pn[2] + plane(3)) /
                    (pn[0] * pn[0] + pn[1] * pn[1] + pn[2] * pn[2]);
    return {p(0) - pn[0] * t, p(1) - pn[1] * t, p(2) - pn[2] * t};
  }
  const double pn[3] = {plane(0), plane(1), plane(2)};
  const double dd[3] = {(*direction)(0), (*direction)(1), (*direction)(2)};
  const double pd[3] = {p(0), p(1), p(2)};
  const double t = (pd[0] * pn[0] + pd[1] * pn[1] + pd[2] * pn[2] + (double)plane(3)) /
                   (dd[0] * pn[0] + dd[1] * pn[1] + dd[2] * pn[2]);
  return {(float)(pd[0] - dd[0] * t), (float)(pd[1] - dd[1] * t), (float)(pd[2] - dd[2] * t)};
}

// ------------------------------------------------------------ homography ----
Matrix3 EstimateHomography(const Points2D& p1, const Points2D& p2) { return homography_dlt(p1, p2); }
Matrix3 EstimateHomography(const Points2D& p1, const Points3D& p2) { return homography_dlt(p1, p2); }
Matrix3 EstimateHomography(const Points3D& p1, const Points2D& p2) { return homography_dlt(p1, p2); }
Matrix3 EstimateHomography(const Points3D& p1, const Points3D& p2) { return homography_dlt(p1, p2); }

// ------------------------------------------------------ Zhang closed form ----
Matrix3 EstimateKFromHomographies(const std::vector<Matrix3>& Hs) {
  if (Hs.size() < 3) throw std::invalid_argument("EstimateKFromHomographies needs at least 3 homographies");
  const int n = (int)Hs.size();
  Mat A(2 * n + 1, 6);
  auto v = [](const Matrix3& H, int i, int j, double* out) {
    const double a0 = H(0, i), a1 = H(1, i), a2 = H(2, i), b0 = H(0, j), b1 = H(1, j), b2 = H(2, j);
    out[0] = a0 * b0;
    out[1] = a0 * b1 + a1 * b0;
    out[2] = a1 * b1;
    out[3] = a2 * b0 + a0 * b2;
    out[4] = a2 * b1 + a1 * b2;
    out[5] = a2 * b2;
  };
  for (int i = 0; i < n; ++i) {
    double v01[6], v00[6], v11[6];
    v(Hs[i], 0, 1, v01);
    v(Hs[i], 0, 0, v00);
    v(Hs[i], 1, 1, v11);
    for (int k = 0; k < 6; ++k) {
      A.at(2 * i, k) = v01[k];
      A.at(2 * i + 1, k) = v00[k] - v11[k];
    }
  }
  A.at(2 * n, 1) = (double)n;  // zero-skew row, weighted with the number of images
  const Svd svd = jacobi_svd(A);
  const int k = svd.smallest();
  const double B11 = svd.V.at(0, k), B12 = svd.V.at(1, k), B22 = svd.V.at(2, k), B13 = svd.V.at(3, k),
               B23 = svd.V.at(4, k), B33 = svd.V.at(5, k);
  const double v0 = (B12 * B13 - B11 * B23) / (B11 * B22 - B12 * B12);
  const double lam = B33 - (B13 * B13 + v0 * (B12 * B13 - B11 * B23)) / B11;
  const double alpha = std::sqrt(lam / B11);
  const double beta = std::sqrt(lam * B11 / (B11 * B22 - B12 * B12));
  const double skew = 0.0;  // forced, as in the reference
  const double u0 = skew * v0 / beta - B13 * alpha * alpha / lam;
  Matrix3 K;
  K(0, 0) = (float)alpha;
  K(0, 1) = (float)skew;
  K(0, 2) = (float)u0;
  K(1, 1) = (float)beta;
  K(1, 2) = (float)v0;
  K(2, 2) = 1.0f;
  return K;
}

Matrix3 FixRotationMatrix(const Matrix3& R) {
  Mat A(3, 3);
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) A.at(r, c) = R(r, c);
  const Svd svd = jacobi_svd(A);
  Matrix3 out;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      double acc = 0.0;
      for (int k = 0; k < 3; ++k) acc += svd.U.at(r, k) * svd.V.at(c, k);
      out(r, c) = (float)acc;
    }
  return out;
}

std::tuple<Matrix3, Point3D> RecoverExtrinsics(const Matrix3& K_inv, const Matrix3& H) {
  const Point3D h0 = K_inv * H.col(0);
  const float l = 1.0f / h0.norm();
  const Point3D h1 = K_inv * H.col(1), h2 = K_inv * H.col(2);
  const Point3D r0(l * h0(0), l * h0(1), l * h0(2));
  const Point3D r1(l * h1(0), l * h1(1), l * h1(2));
  const Point3D r2 = r0.cross(r1);
  const Point3D t(l * h2(0), l * h2(1), l * h2(2));
  Matrix3 R;
  for (int r = 0; r < 3; ++r) {
    R(r, 0) = r0(r);
    R(r, 1) = r1(r);
    R(r, 2) = r2(r);
  }
  return {FixRotationMatrix(R), t};
}

// [ext] Eigen::Quaternionf(Matrix3f): trace test, else the largest diagonal entry
Quaternion::Quaternion(const Matrix3& R) {
  const float tr = R(0, 0) + R(1, 1) + R(2, 2);
  if (tr > 0.0f) {
    float s = std::sqrt(tr + 1.0f);
    q[0] = 0.5f * s;
    s = 0.5f / s;
    q[1] = (R(2, 1) - R(1, 2)) * s;
    q[2] = (R(0, 2) - R(2, 0)) * s;
    q[3] = (R(1, 0) - R(0, 1)) * s;
    return;
  }
  int i = 0;
  if (R(1, 1) > R(0, 0)) i = 1;
  if (R(2, 2) > R(i, i)) i = 2;
  const int j = (i + 1) % 3, k = (j + 1) % 3;
  float s = std::sqrt(R(i, i) - R(j, j) - R(k, k) + 1.0f);
  q[1 + i] = 0.5f * s;
  s = 0.5f / s;
  q[0] = (R(k, j) - R(j, k)) * s;
  q[1 + j] = (R(j, i) + R(i, j)) * s;
  q[1 + k] = (R(k, i) + R(i, k)) * s;
}

}  // namespace calibrator
