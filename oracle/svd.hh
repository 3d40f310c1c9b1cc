// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.h header).
//
// One-sided (Hestenes) Jacobi SVD in double: stands in for the
// Eigen::JacobiSVD<MatrixXd> calls of reference src/geometry.cpp:99,154,200
// (Eigen is not available in this image).  Singular vectors are defined up to
// sign, exactly as with Eigen; callers only use sign-invariant quantities or
// document the ambiguity.
#pragma once
#include <cmath>
#include <vector>

namespace oracle {

// A: m x n row-major (m >= 1).  On return `V` (n x n row-major) holds the right
// singular vectors as columns and `sigma` the singular values (unsorted).
inline void jacobi_svd(std::vector<double> A, int m, int n, std::vector<double>* V,
                       std::vector<double>* sigma, std::vector<double>* Uout = nullptr) {
  V->assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) (*V)[(size_t)i * n + i] = 1.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        double app = 0.0, aqq = 0.0, apq = 0.0;
        for (int i = 0; i < m; ++i) {
          const double x = A[(size_t)i * n + p], y = A[(size_t)i * n + q];
          app += x * x;
          aqq += y * y;
          apq += x * y;
        }
        if (apq == 0.0) continue;
        const double denom = std::sqrt(app * aqq);
        if (denom == 0.0 || std::fabs(apq) <= 1e-17 * denom) continue;
        off = std::fmax(off, std::fabs(apq) / denom);
        const double zeta = (aqq - app) / (2.0 * apq);
        const double tt = (zeta >= 0.0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / std::sqrt(1.0 + tt * tt), s = c * tt;
        for (int i = 0; i < m; ++i) {
          const double x = A[(size_t)i * n + p], y = A[(size_t)i * n + q];
          A[(size_t)i * n + p] = c * x - s * y;
          A[(size_t)i * n + q] = s * x + c * y;
        }
        for (int i = 0; i < n; ++i) {
          const double x = (*V)[(size_t)i * n + p], y = (*V)[(size_t)i * n + q];
          (*V)[(size_t)i * n + p] = c * x - s * y;
          (*V)[(size_t)i * n + q] = s * x + c * y;
        }
      }
    if (off < 1e-15) break;
  }
  sigma->assign(n, 0.0);
  for (int j = 0; j < n; ++j) {
    double s = 0.0;
    for (int i = 0; i < m; ++i) s += A[(size_t)i * n + j] * A[(size_t)i * n + j];
    (*sigma)[j] = std::sqrt(s);
  }
  if (Uout) {
    Uout->assign((size_t)m * n, 0.0);
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < m; ++i)
        (*Uout)[(size_t)i * n + j] = (*sigma)[j] > 0.0 ? A[(size_t)i * n + j] / (*sigma)[j] : 0.0;
  }
}

// Right singular vector of the smallest singular value ("rightCols(1)" of a
// full-V JacobiSVD whose singular values are sorted in decreasing order).
inline std::vector<double> null_vector(const std::vector<double>& A, int m, int n) {
  std::vector<double> V, s;
  jacobi_svd(A, m, n, &V, &s);
  int best = 0;
  for (int j = 1; j < n; ++j)
    if (s[j] < s[best]) best = j;
  std::vector<double> out(n);
  for (int i = 0; i < n; ++i) out[i] = V[(size_t)i * n + best];
  return out;
}

}  // namespace oracle
