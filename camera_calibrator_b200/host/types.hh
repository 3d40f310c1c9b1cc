// Boundary types of the host classes.
//
// The reference's API is written against float32 Eigen aliases (reference
// src/types.hh:10-20).  Eigen is not available in this image, so this header
// provides a minimal Eigen-free mirror with the members the reference's callers
// use (x(), y(), z(), operator(), Identity(), translation(), linear(), ...).
// Layout notes: Matrix3 / Affine3f are stored row-major here; the JSON wire
// format and the numpy casters convert explicitly, so this is not observable.
#pragma once
#include <array>
#include <cmath>
#include <cstddef>
#include <vector>

namespace calibrator {

struct Point2D {
  float v[2] = {0.f, 0.f};
  Point2D() = default;
  Point2D(float x, float y) : v{x, y} {}
  float x() const { return v[0]; }
  float y() const { return v[1]; }
  float& x() { return v[0]; }
  float& y() { return v[1]; }
  float operator()(int i) const { return v[i]; }
  float& operator()(int i) { return v[i]; }
};

struct Point3D {
  float v[3] = {0.f, 0.f, 0.f};
  Point3D() = default;
  Point3D(float x, float y, float z) : v{x, y, z} {}
  float x() const { return v[0]; }
  float y() const { return v[1]; }
  float z() const { return v[2]; }
  float& x() { return v[0]; }
  float& y() { return v[1]; }
  float& z() { return v[2]; }
  float operator()(int i) const { return v[i]; }
  float& operator()(int i) { return v[i]; }
  float norm() const { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
  Point3D cross(const Point3D& o) const {
    return {v[1] * o.v[2] - v[2] * o.v[1], v[2] * o.v[0] - v[0] * o.v[2], v[0] * o.v[1] - v[1] * o.v[0]};
  }
};
using Vector3 = Point3D;
using Points2D = std::vector<Point2D>;
using Points3D = std::vector<Point3D>;

struct Matrix3 {
  float m[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // row-major
  static Matrix3 Identity() {
    Matrix3 r;
    r.m[0] = r.m[4] = r.m[8] = 1.f;
    return r;
  }
  float operator()(int r, int c) const { return m[r * 3 + c]; }
  float& operator()(int r, int c) { return m[r * 3 + c]; }
  Point3D col(int c) const { return {m[c], m[3 + c], m[6 + c]}; }
  Point3D operator*(const Point3D& p) const {
    return {m[0] * p.v[0] + m[1] * p.v[1] + m[2] * p.v[2], m[3] * p.v[0] + m[4] * p.v[1] + m[5] * p.v[2],
            m[6] * p.v[0] + m[7] * p.v[1] + m[8] * p.v[2]};
  }
  Matrix3 inverse() const {
    const float c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
    const float id = 1.0f / (m[0] * c00 + m[1] * c01 + m[2] * c02);
    Matrix3 r;
    r.m[0] = c00 * id;
    r.m[1] = (m[2] * m[7] - m[1] * m[8]) * id;
    r.m[2] = (m[1] * m[5] - m[2] * m[4]) * id;
    r.m[3] = c01 * id;
    r.m[4] = (m[0] * m[8] - m[2] * m[6]) * id;
    r.m[5] = (m[2] * m[3] - m[0] * m[5]) * id;
    r.m[6] = c02 * id;
    r.m[7] = (m[1] * m[6] - m[0] * m[7]) * id;
    r.m[8] = (m[0] * m[4] - m[1] * m[3]) * id;
    return r;
  }
};

// VectorXf stand-in (distortion coefficients)
struct DynamicVector {
  std::vector<float> d;
  DynamicVector() = default;
  explicit DynamicVector(size_t n) : d(n, 0.f) {}
  static DynamicVector Zero(size_t n) { return DynamicVector(n); }
  size_t size() const { return d.size(); }
  float operator()(size_t i) const { return d[i]; }
  float& operator()(size_t i) { return d[i]; }
};

// Quaternionf stand-in, storage (w,x,y,z)
struct Quaternion {
  float q[4] = {1.f, 0.f, 0.f, 0.f};
  Quaternion() = default;
  Quaternion(float w, float x, float y, float z) : q{w, x, y, z} {}
  explicit Quaternion(const Matrix3& R);  // [ext] Eigen's rotation-matrix constructor (geometry.cpp)
  float w() const { return q[0]; }
  float x() const { return q[1]; }
  float y() const { return q[2]; }
  float z() const { return q[3]; }
};

// Affine3f stand-in: rotation (or general linear) part + translation
struct Affine3f {
  Matrix3 R = Matrix3::Identity();
  Point3D t;
  static Affine3f Identity() { return Affine3f(); }
  const Matrix3& linear() const { return R; }
  Matrix3& linear() { return R; }
  const Matrix3& rotation() const { return R; }
  const Point3D& translation() const { return t; }
  Point3D& translation() { return t; }
  // 4x4 homogeneous matrix, column-major like Eigen's .matrix().reshaped() (JSON wire format)
  std::array<float, 16> matrix_colmajor() const {
    std::array<float, 16> o{};
    for (int c = 0; c < 3; ++c)
      for (int r = 0; r < 3; ++r) o[c * 4 + r] = R(r, c);
    for (int r = 0; r < 3; ++r) o[12 + r] = t(r);
    o[15] = 1.f;
    return o;
  }
  static Affine3f from_colmajor(const float* a) {
    Affine3f T;
    for (int c = 0; c < 3; ++c)
      for (int r = 0; r < 3; ++r) T.R(r, c) = a[c * 4 + r];
    for (int r = 0; r < 3; ++r) T.t(r) = a[12 + r];
    return T;
  }
};

}  // namespace calibrator
