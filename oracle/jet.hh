// ORACLE — TEST INFRASTRUCTURE ONLY. Nothing under oracle/ is part of the
// shipped product path; only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may build, load or call it.
//
// Forward-mode dual number ("jet"), the same differentiation scheme the
// reference relies on: its residual functors are templated on T and handed to
// ceres::AutoDiffCostFunction (reference src/calibrator.cpp:294-297,
// src/extrinsics_calibrator.cpp:168-171), which evaluates them on
// ceres::Jet<double, N> [ext: Ceres Solver >=2.1, conanfile.txt:3 — not
// vendored, restated from its published behaviour].
#pragma once
#include <cmath>

namespace oracle {

template <int N>
struct Jet {
  double a;
  double v[N];

  Jet() : a(0.0) {
    for (int i = 0; i < N; ++i) v[i] = 0.0;
  }
  Jet(double s) : a(s) {  // NOLINT: implicit on purpose (constants in functors)
    for (int i = 0; i < N; ++i) v[i] = 0.0;
  }
  static Jet variable(double s, int k) {
    Jet j(s);
    j.v[k] = 1.0;
    return j;
  }
};

template <int N>
inline Jet<N> operator+(const Jet<N>& x, const Jet<N>& y) {
  Jet<N> r;
  r.a = x.a + y.a;
  for (int i = 0; i < N; ++i) r.v[i] = x.v[i] + y.v[i];
  return r;
}
template <int N>
inline Jet<N> operator-(const Jet<N>& x, const Jet<N>& y) {
  Jet<N> r;
  r.a = x.a - y.a;
  for (int i = 0; i < N; ++i) r.v[i] = x.v[i] - y.v[i];
  return r;
}
template <int N>
inline Jet<N> operator-(const Jet<N>& x) {
  Jet<N> r;
  r.a = -x.a;
  for (int i = 0; i < N; ++i) r.v[i] = -x.v[i];
  return r;
}
template <int N>
inline Jet<N> operator*(const Jet<N>& x, const Jet<N>& y) {
  Jet<N> r;
  r.a = x.a * y.a;
  for (int i = 0; i < N; ++i) r.v[i] = x.a * y.v[i] + x.v[i] * y.a;
  return r;
}
template <int N>
inline Jet<N> operator/(const Jet<N>& x, const Jet<N>& y) {
  Jet<N> r;
  // value part rounded exactly like a plain double division, so that a
  // cost-only (double) evaluation and a Jet evaluation agree bit for bit
  r.a = x.a / y.a;
  const double inv = 1.0 / y.a;
  for (int i = 0; i < N; ++i) r.v[i] = (x.v[i] - r.a * y.v[i]) * inv;
  return r;
}
template <int N>
inline Jet<N> operator+(const Jet<N>& x, double s) { Jet<N> r = x; r.a += s; return r; }
template <int N>
inline Jet<N> operator+(double s, const Jet<N>& x) { return x + s; }
template <int N>
inline Jet<N> operator-(const Jet<N>& x, double s) { Jet<N> r = x; r.a -= s; return r; }
template <int N>
inline Jet<N> operator-(double s, const Jet<N>& x) { return (-x) + s; }
template <int N>
inline Jet<N> operator*(const Jet<N>& x, double s) {
  Jet<N> r;
  r.a = x.a * s;
  for (int i = 0; i < N; ++i) r.v[i] = x.v[i] * s;
  return r;
}
template <int N>
inline Jet<N> operator*(double s, const Jet<N>& x) { return x * s; }
template <int N>
inline Jet<N> operator/(const Jet<N>& x, double s) { return x * (1.0 / s); }
template <int N>
inline Jet<N>& operator+=(Jet<N>& x, const Jet<N>& y) { x = x + y; return x; }
template <int N>
inline Jet<N> sqrt(const Jet<N>& x) {
  Jet<N> r;
  r.a = std::sqrt(x.a);
  const double d = 0.5 / r.a;
  for (int i = 0; i < N; ++i) r.v[i] = x.v[i] * d;
  return r;
}

inline double value_of(double x) { return x; }
template <int N>
inline double value_of(const Jet<N>& x) { return x.a; }

}  // namespace oracle
