// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.h header).
//
// Levenberg-Marquardt with exact Schur elimination of the per-view (per-frame)
// 6-DoF pose blocks and a dense Cholesky solve of the reduced system, FP64,
// Jacobians by forward-mode Jets like Ceres autodiff.
//
// What this restates:
//   problem structure, parameter packing, free/frozen sets:
//       reference src/calibrator.cpp:245-312 and src/extrinsics_calibrator.cpp:105-204
//   trust-region LM constants (initial radius 1e4, accept if rho > 1e-3, radius
//   update 1/max(1/3, 1-(2rho-1)^3), halving with doubling factor on reject,
//   diagonal clamp [1e-6, 1e32]): [ext] Ceres Solver defaults, selected by
//       reference src/calibrator.cpp:314-321 / src/extrinsics_calibrator.cpp:206-213
//   Huber corrector (sqrt(rho') scaling):  [ext] Ceres, reference
//       src/extrinsics_calibrator.cpp:175-176
//   per-observation final cost: reference src/extrinsics_calibrator.cpp:219-225
//
// What deliberately differs from Ceres: the linear solve
// is exact (Schur + Cholesky) instead of CG on the implicit Schur complement,
// there are no inner iterations and no non-monotonic steps, and termination is
// tightened so the iteration runs to gradient convergence — only the converged
// minimum is comparable between solvers.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "models.hh"
#include "oracle.h"

namespace oracle {

// ---- small dense helpers -------------------------------------------------
// In-place lower Cholesky of a row-major n x n SPD matrix (upper part ignored).
static bool chol(double* A, int n) {
  for (int j = 0; j < n; ++j) {
    double d = A[j * n + j];
    for (int k = 0; k < j; ++k) d -= A[j * n + k] * A[j * n + k];
    if (!(d > 0.0) || !std::isfinite(d)) return false;
    d = std::sqrt(d);
    A[j * n + j] = d;
    for (int i = j + 1; i < n; ++i) {
      double s = A[i * n + j];
      for (int k = 0; k < j; ++k) s -= A[i * n + k] * A[j * n + k];
      A[i * n + j] = s / d;
    }
  }
  return true;
}
static void chol_solve(const double* L, int n, double* b) {
  for (int i = 0; i < n; ++i) {
    double s = b[i];
    for (int k = 0; k < i; ++k) s -= L[i * n + k] * b[k];
    b[i] = s / L[i * n + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = b[i];
    for (int k = i + 1; k < n; ++k) s -= L[k * n + i] * b[k];
    b[i] = s / L[i * n + i];
  }
}
static inline double clampd(double x, double lo, double hi) { return std::min(std::max(x, lo), hi); }

static int g_h_float = 0;  // experiment switch, see eval_single
// EXPERIMENT switch for the gradient g = J'r of eval_single (tools/g_precision_study.py): 0 = double (the solver);
// 1 = Jacobian entries rounded to float32, residual and sums in double; 2 = float32 Jacobian, float32 residual,
// float32 sums flushed every 13 observations (a lane's share of a view); 3 = Jacobian and residual rounded to 11
// significant bits (FP16 operands), float32 sums.  Answers: how far does the minimum move if a sweep kernel took the
// gradient from its float32 Jacobian rows instead of an FP64 reverse pass?
static int g_g_float = 0;

// relative cost noise allowed when taking a final sub-resolution step
static const double kNoise = 1e-12;
static const double kHalfEps = 1.1102230246251565e-16;
static const int kBlindSteps = 2;

// Rounding noise of the FP64 cost: every residual is a difference of numbers of
// size `meas_scale` (pixels ~1e3), so it carries ~eps*meas_scale absolute error
// and the cost ~eps*meas_scale*||r||.  A step whose PREDICTED decrease is below
// `noise_factor` times that cannot be validated by the gain ratio.
static inline double cost_noise(const oracle_opts& O, double meas_scale, double rnorm) {
  return O.noise_factor * kHalfEps * meas_scale * rnorm;
}
static double max_abs(const float* a, int64_t n) {
  float m = 0.0f;
  for (int64_t i = 0; i < n; ++i) m = std::max(m, std::fabs(a[i]));
  return (double)m;
}

// =========================================================================
//                       single-camera problem
// =========================================================================
struct SingleData {
  const float *u, *v, *X, *Y, *Z;
  const int64_t* voff;
  int64_t nv;
  int64_t nobs() const { return voff[nv] - voff[0]; }
};

struct SingleEval {
  double cost = 0.0;
  double U[81];
  double gu[9];
  std::vector<double> V, W, gv;  // 36, 54, 6 per view
};

// Full evaluation: residuals + Jacobians (Jets) + normal-equation blocks.
static void eval_single(const SingleData& D, const double* intr, const double* q, const double* t,
                        SingleEval* E, double* res_out) {
  const int64_t nv = D.nv;
  E->V.assign((size_t)nv * 36, 0.0);
  E->W.assign((size_t)nv * 54, 0.0);
  E->gv.assign((size_t)nv * 6, 0.0);
  std::vector<double> part((size_t)nv * 91, 0.0);  // per view: U 81, gu 9, cost 1
  typedef Jet<16> J;
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t iv = 0; iv < nv; ++iv) {
    J ji[9], jq[4], jt[3];
    for (int k = 0; k < 9; ++k) ji[k] = J::variable(intr[k], k);
    for (int k = 0; k < 4; ++k) jq[k] = J::variable(q[iv * 4 + k], 9 + k);
    for (int k = 0; k < 3; ++k) jt[k] = J::variable(t[iv * 3 + k], 13 + k);
    double PJ[12];
    quaternion_plus_jacobian(q + iv * 4, PJ);
    double* Vv = &E->V[(size_t)iv * 36];
    double* Wv = &E->W[(size_t)iv * 54];
    double* gvv = &E->gv[(size_t)iv * 6];
    double* Uv = &part[(size_t)iv * 91];
    double* guv = Uv + 81;
    double cost = 0.0;
    float Uf[81] = {0}, Vf[36] = {0}, Wf[54] = {0};
    float guf[9] = {0}, gvf[6] = {0};
    auto rnd11 = [](float x) {
      if (x == 0.0f) return 0.0f;
      int e;
      const float m = std::ldexp(std::frexp(x, &e), 11);
      return std::ldexp(std::nearbyint(m), e - 11);
    };
    // the gradient contribution of one residual row under the experiment switch g_g_float (guv / gvv += J' r)
    auto add_gradient = [&](const double* Ji, const double* Jp, double rr) {
      if (g_g_float == 0) {
        for (int a = 0; a < 9; ++a) guv[a] += Ji[a] * rr;
        for (int a = 0; a < 6; ++a) gvv[a] += Jp[a] * rr;
      } else if (g_g_float == 1) {
        for (int a = 0; a < 9; ++a) guv[a] += (double)(float)Ji[a] * rr;
        for (int a = 0; a < 6; ++a) gvv[a] += (double)(float)Jp[a] * rr;
      } else {
        const float rf = g_g_float == 3 ? rnd11((float)rr) : (float)rr;
        for (int a = 0; a < 9; ++a) guf[a] += (g_g_float == 3 ? rnd11((float)Ji[a]) : (float)Ji[a]) * rf;
        for (int a = 0; a < 6; ++a) gvf[a] += (g_g_float == 3 ? rnd11((float)Jp[a]) : (float)Jp[a]) * rf;
      }
    };
    auto flush_gradient = [&]() {
      for (int a = 0; a < 9; ++a) { guv[a] += guf[a]; guf[a] = 0; }
      for (int a = 0; a < 6; ++a) { gvv[a] += gvf[a]; gvf[a] = 0; }
    };
    for (int64_t o = D.voff[iv]; o < D.voff[iv + 1]; ++o) {
      if (g_g_float >= 2 && ((o - D.voff[iv]) % 13) == 12) flush_gradient();
      if (g_h_float && ((o - D.voff[iv]) % 13) == 12) {  // flush like a lane's ~13 observations
        for (int k = 0; k < 81; ++k) { Uv[k] += Uf[k]; Uf[k] = 0; }
        for (int k = 0; k < 36; ++k) { Vv[k] += Vf[k]; Vf[k] = 0; }
        for (int k = 0; k < 54; ++k) { Wv[k] += Wf[k]; Wf[k] = 0; }
      }
      J Xw[3] = {J((double)D.X[o]), J((double)D.Y[o]), J((double)D.Z[o])};
      J r[2];
      residual_single<J>(ji, jq, jt, Xw, (double)D.u[o], (double)D.v[o], r);
      if (res_out) {
        res_out[2 * (o - D.voff[0])] = r[0].a;
        res_out[2 * (o - D.voff[0]) + 1] = r[1].a;
      }
      for (int row = 0; row < 2; ++row) {
        double Ji[9], Jp[6];
        for (int k = 0; k < 9; ++k) Ji[k] = r[row].v[k];
        for (int k = 0; k < 3; ++k) {
          double s = 0.0;
          for (int a = 0; a < 4; ++a) s += r[row].v[9 + a] * PJ[a * 3 + k];
          Jp[k] = s;
          Jp[3 + k] = r[row].v[13 + k];
        }
        const double rr = r[row].a;
        cost += 0.5 * rr * rr;
        if (g_h_float) {
          // EXPERIMENT (ORACLE_H_FLOAT=1): J'J products and per-view sums in float32, g and r in
          // double — emulates a mixed-precision sweep kernel; only the fixed point g = 0 is exact
          float Jif[9], Jpf[6];
          for (int a = 0; a < 9; ++a) Jif[a] = (float)Ji[a] * (g_h_float == 2 ? (1.0f + 1e-6f * ((a * 7 + row) % 3 - 1)) : 1.0f);
          for (int a = 0; a < 6; ++a) Jpf[a] = (float)Jp[a] * (g_h_float == 2 ? (1.0f + 1e-6f * ((a * 5 + row) % 3 - 1)) : 1.0f);
          if (g_h_float >= 100) {
            // EXPERIMENT (modes 1BB / 2BB): the J'J products as a split low-precision tensor-core
            // contraction would form them.  Every Jacobian entry x = h + l with h, l rounded to BB
            // significant bits (8 = bfloat16, 11 = tf32 / half); mode 1BB accumulates hh + hl + lh
            // (two-pass split, ll dropped), modes 2BB / 3BB only hh (single pass); float32 accumulation.
            const int bits = g_h_float % 100;
            const bool two = g_h_float < 200;
            const bool trunc = g_h_float >= 300;  // mode 3BB: single pass, operands truncated (what
                                                  // mma .tf32 does to raw float32 bits)
            auto rnd = [bits, trunc](float x) {
              if (x == 0.0f) return 0.0f;
              int e;
              const float m = std::ldexp(std::frexp(x, &e), bits);
              return std::ldexp(trunc ? std::trunc(m) : std::nearbyint(m), e - bits);
            };
            float hi[15], lo[15];
            for (int a = 0; a < 6; ++a) { hi[a] = rnd(Jpf[a]); lo[a] = two ? rnd(Jpf[a] - hi[a]) : 0.0f; }
            for (int a = 0; a < 9; ++a) { hi[6 + a] = rnd(Jif[a]); lo[6 + a] = two ? rnd(Jif[a] - hi[6 + a]) : 0.0f; }
            auto prod = [&](int a, int b) { return hi[a] * hi[b] + (hi[a] * lo[b] + lo[a] * hi[b]); };
            add_gradient(Ji, Jp, rr);
            for (int a = 0; a < 9; ++a) {
              for (int b = 0; b < 9; ++b) Uf[a * 9 + b] += prod(6 + a, 6 + b);
            }
            for (int a = 0; a < 6; ++a) {
              for (int b = 0; b < 6; ++b) Vf[a * 6 + b] += prod(a, b);
              for (int b = 0; b < 9; ++b) Wf[a * 9 + b] += prod(a, 6 + b);
            }
            continue;
          }
          add_gradient(Ji, Jp, rr);
          for (int a = 0; a < 9; ++a) {
            for (int b = 0; b < 9; ++b) Uf[a * 9 + b] += Jif[a] * Jif[b];
          }
          for (int a = 0; a < 6; ++a) {
            for (int b = 0; b < 6; ++b) Vf[a * 6 + b] += Jpf[a] * Jpf[b];
            for (int b = 0; b < 9; ++b) Wf[a * 9 + b] += Jpf[a] * Jif[b];
          }
          continue;
        }
        add_gradient(Ji, Jp, rr);
        for (int a = 0; a < 9; ++a) {
          for (int b = 0; b < 9; ++b) Uv[a * 9 + b] += Ji[a] * Ji[b];
        }
        for (int a = 0; a < 6; ++a) {
          for (int b = 0; b < 6; ++b) Vv[a * 6 + b] += Jp[a] * Jp[b];
          for (int b = 0; b < 9; ++b) Wv[a * 9 + b] += Jp[a] * Ji[b];
        }
      }
    }
    if (g_g_float >= 2) flush_gradient();
    if (g_h_float) {
      for (int k = 0; k < 81; ++k) Uv[k] += Uf[k];
      for (int k = 0; k < 36; ++k) Vv[k] = (double)(float)(Vv[k] + Vf[k]);
      for (int k = 0; k < 54; ++k) Wv[k] = (double)(float)(Wv[k] + Wf[k]);
    }
    Uv[90] = cost;
  }
  // fixed-order (view-order) sum: independent of the thread count
  std::memset(E->U, 0, sizeof(E->U));
  std::memset(E->gu, 0, sizeof(E->gu));
  E->cost = 0.0;
  for (int64_t iv = 0; iv < nv; ++iv) {
    const double* p = &part[(size_t)iv * 91];
    for (int k = 0; k < 81; ++k) E->U[k] += p[k];
    for (int k = 0; k < 9; ++k) E->gu[k] += p[81 + k];
    E->cost += p[90];
  }
}

static double cost_single(const SingleData& D, const double* intr, const double* q, const double* t) {
  const int64_t nv = D.nv;
  std::vector<double> part((size_t)nv, 0.0);
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t iv = 0; iv < nv; ++iv) {
    double c = 0.0;
    for (int64_t o = D.voff[iv]; o < D.voff[iv + 1]; ++o) {
      const double Xw[3] = {(double)D.X[o], (double)D.Y[o], (double)D.Z[o]};
      double r[2];
      residual_single<double>(intr, q + iv * 4, t + iv * 3, Xw, (double)D.u[o], (double)D.v[o], r);
      c += 0.5 * r[0] * r[0];
      c += 0.5 * r[1] * r[1];
    }
    part[(size_t)iv] = c;
  }
  double c = 0.0;
  for (int64_t iv = 0; iv < nv; ++iv) c += part[(size_t)iv];
  return c;
}

static double grad_norm_single(const SingleEval& E, int64_t nv, uint32_t mask) {
  double s = 0.0;
  for (int k = 0; k < 9; ++k)
    if (!((mask >> k) & 1u) && E.U[k * 9 + k] > 0.0) s += E.gu[k] * E.gu[k] / E.U[k * 9 + k];
  for (int64_t iv = 0; iv < nv; ++iv)
    for (int k = 0; k < 6; ++k) {
      const double d = E.V[(size_t)iv * 36 + k * 6 + k];
      const double g = E.gv[(size_t)iv * 6 + k];
      if (d > 0.0) s += g * g / d;
    }
  return std::sqrt(s);
}

static int solve_single(const SingleData& D, double* intr, double* q, double* t, uint32_t mask,
                        const oracle_opts& O, oracle_report* R) {
  const int64_t nv = D.nv;
  const int64_t nobs = D.nobs();
  const double meas_scale = std::max(max_abs(D.u + D.voff[0], nobs), max_abs(D.v + D.voff[0], nobs));
  SingleEval E;
  eval_single(D, intr, q, t, &E, nullptr);
  oracle_report rep;
  std::memset(&rep, 0, sizeof(rep));
  rep.sweeps = 1;
  rep.initial_cost = E.cost;
  double radius = O.initial_radius, nu = 2.0;
  int blind = 0;  // sub-resolution steps taken so far
  std::vector<double> Y((size_t)nv * 54), z((size_t)nv * 6), dv((size_t)nv * 6), Dv((size_t)nv * 6);
  std::vector<double> qc((size_t)nv * 4), tc((size_t)nv * 3);
  rep.termination = 4;
  if (!std::isfinite(E.cost)) {
    rep.termination = -2;
    if (R) *R = rep;
    return -2;
  }
  for (int it = 0; it < O.max_iterations; ++it) {
    const double rnorm = std::sqrt(2.0 * E.cost);
    const double gn = grad_norm_single(E, nv, mask);
    if (E.cost == 0.0 || gn <= O.gradient_tolerance * rnorm) {
      rep.termination = 1;
      break;
    }
    rep.iterations++;
    const double lambda = 1.0 / radius;
    // --- Schur complement of the pose blocks onto the intrinsics
    double S[81], rhs[9], Du[9];
    for (int a = 0; a < 9; ++a) {
      Du[a] = clampd(E.U[a * 9 + a], O.min_diagonal, O.max_diagonal);
      rhs[a] = -E.gu[a];
      for (int b = 0; b < 9; ++b) S[a * 9 + b] = E.U[a * 9 + b];
      S[a * 9 + a] += lambda * Du[a];
    }
    bool ok = true;
    for (int64_t iv = 0; iv < nv && ok; ++iv) {
      double L[36];
      std::memcpy(L, &E.V[(size_t)iv * 36], sizeof(L));
      for (int k = 0; k < 6; ++k) {
        Dv[(size_t)iv * 6 + k] = clampd(L[k * 6 + k], O.min_diagonal, O.max_diagonal);
        L[k * 6 + k] += lambda * Dv[(size_t)iv * 6 + k];
      }
      if (!chol(L, 6)) {
        ok = false;
        break;
      }
      const double* Wv = &E.W[(size_t)iv * 54];
      double* Yv = &Y[(size_t)iv * 54];
      double* zv = &z[(size_t)iv * 6];
      for (int b = 0; b < 9; ++b) {
        double col[6];
        for (int a = 0; a < 6; ++a) col[a] = Wv[a * 9 + b];
        chol_solve(L, 6, col);
        for (int a = 0; a < 6; ++a) Yv[a * 9 + b] = col[a];
      }
      for (int a = 0; a < 6; ++a) zv[a] = E.gv[(size_t)iv * 6 + a];
      chol_solve(L, 6, zv);
      for (int a = 0; a < 9; ++a) {
        double s = 0.0;
        for (int k = 0; k < 6; ++k) s += Wv[k * 9 + a] * zv[k];
        rhs[a] += s;
        for (int b = 0; b < 9; ++b) {
          double w = 0.0;
          for (int k = 0; k < 6; ++k) w += Wv[k * 9 + a] * Yv[k * 9 + b];
          S[a * 9 + b] -= w;
        }
      }
    }
    if (ok) {
      for (int k = 0; k < 9; ++k)
        if ((mask >> k) & 1u) {
          for (int b = 0; b < 9; ++b) S[k * 9 + b] = S[b * 9 + k] = 0.0;
          S[k * 9 + k] = 1.0;
          rhs[k] = 0.0;
        }
      ok = chol(S, 9);
    }
    if (!ok) {
      rep.rejected++;
      radius /= nu;
      nu *= 2.0;
      if (radius < O.min_radius) {
        rep.termination = 5;
        break;
      }
      continue;
    }
    double du[9];
    std::memcpy(du, rhs, sizeof(du));
    chol_solve(S, 9, du);
    double gTd = 0.0, dDd = 0.0;
    for (int k = 0; k < 9; ++k) {
      if ((mask >> k) & 1u) continue;
      gTd += E.gu[k] * du[k];
      dDd += du[k] * du[k] * Du[k];
    }
    for (int64_t iv = 0; iv < nv; ++iv) {
      for (int a = 0; a < 6; ++a) {
        double s = -z[(size_t)iv * 6 + a];
        for (int b = 0; b < 9; ++b) s -= Y[(size_t)iv * 54 + a * 9 + b] * du[b];
        dv[(size_t)iv * 6 + a] = s;
        gTd += E.gv[(size_t)iv * 6 + a] * s;
        dDd += s * s * Dv[(size_t)iv * 6 + a];
      }
    }
    const double model = 0.5 * (lambda * dDd - gTd);
    // A step this small changes the cost by less than FP64 can resolve, so the
    // gain ratio cannot validate it: take it if the cost does not rise above
    // rounding noise, then stop.
    const bool tiny = model <= cost_noise(O, meas_scale, rnorm) ||
                      std::sqrt(dDd) <= O.parameter_tolerance * rnorm;
    double ic[9];
    for (int k = 0; k < 9; ++k) ic[k] = intr[k] + du[k];
    for (int64_t iv = 0; iv < nv; ++iv) {
      quaternion_plus(q + iv * 4, &dv[(size_t)iv * 6], &qc[(size_t)iv * 4]);
      for (int k = 0; k < 3; ++k) tc[(size_t)iv * 3 + k] = t[iv * 3 + k] + dv[(size_t)iv * 6 + 3 + k];
    }
    const double cc = cost_single(D, ic, qc.data(), tc.data());
    const double rho = (E.cost - cc) / model;
    if (O.verbose)
      std::printf("[oracle] it %3d cost %.17g cand %.17g rho %.3e radius %.3e |g| %.3e |d| %.3e\n", it,
                  E.cost, cc, rho, radius, gn / rnorm, std::sqrt(dDd) / rnorm);
    if (tiny || blind > 0) {
      // sub-resolution step: taken unless the cost rises; one more polishing step follows
      bool took = false;
      if (std::isfinite(cc) && cc <= E.cost * (1.0 + kNoise) + 4.0 * cost_noise(O, meas_scale, rnorm)) {
        std::memcpy(intr, ic, sizeof(ic));
        std::memcpy(q, qc.data(), sizeof(double) * (size_t)nv * 4);
        std::memcpy(t, tc.data(), sizeof(double) * (size_t)nv * 3);
        rep.accepted++;
        eval_single(D, intr, q, t, &E, nullptr);
        rep.sweeps++;
        took = true;
      } else {
        rep.rejected++;
      }
      ++blind;
      if (!took || blind >= kBlindSteps) {
        rep.termination = 3;
        break;
      }
      continue;
    }
    if (std::isfinite(cc) && model > 0.0 && rho > O.min_relative_decrease) {
      const double old = E.cost;
      std::memcpy(intr, ic, sizeof(ic));
      std::memcpy(q, qc.data(), sizeof(double) * (size_t)nv * 4);
      std::memcpy(t, tc.data(), sizeof(double) * (size_t)nv * 3);
      rep.accepted++;
      const double f = 2.0 * rho - 1.0;
      radius = std::min(O.max_radius, radius / std::max(1.0 / 3.0, 1.0 - f * f * f));
      nu = 2.0;
      eval_single(D, intr, q, t, &E, nullptr);
      rep.sweeps++;
      if (old - E.cost <= O.function_tolerance * old) {
        rep.termination = 2;
        break;
      }
    } else {
      rep.rejected++;
      radius /= nu;
      nu *= 2.0;
      if (radius < O.min_radius) {
        rep.termination = 5;
        break;
      }
    }
  }
  rep.final_cost = E.cost;
  const double rnorm = std::sqrt(2.0 * E.cost);
  rep.gradient_norm = rnorm > 0.0 ? grad_norm_single(E, nv, mask) / rnorm : 0.0;
  rep.rms = nobs > 0 ? std::sqrt(2.0 * E.cost / (double)nobs) : 0.0;
  if (R) *R = rep;
  return 0;
}

// =========================================================================
//                            rig problem
// =========================================================================
struct RigData {
  const float *u, *v;
  const int32_t* cam;
  const int64_t* pt;
  const int64_t* foff;
  int64_t nf;
  const float* Xw;
  int32_t nc;
  double huber_a;
  int64_t nobs() const { return foff[nf] - foff[0]; }
};

struct RigEval {
  double cost = 0.0;
  double sumsq = 0.0;           // raw sum ||r||^2 (for the RMS)
  std::vector<double> U, gu;    // (6C)^2, 6C
  std::vector<double> V, W, gv;  // 36, 6*6C, 6 per frame
};

static void eval_rig(const RigData& D, const double* qc, const double* tc, const double* qf,
                     const double* tf, RigEval* E, double* res_out) {
  const int nc = D.nc, n6 = 6 * nc;
  const int64_t nf = D.nf;
  E->V.assign((size_t)nf * 36, 0.0);
  E->W.assign((size_t)nf * 6 * n6, 0.0);
  E->gv.assign((size_t)nf * 6, 0.0);
  // per frame: block-diagonal U (36 per camera), gu (6 per camera), cost, sumsq
  const size_t ps = (size_t)nc * 42 + 2;
  std::vector<double> part((size_t)nf * ps, 0.0);
  typedef Jet<14> J;
#pragma omp parallel for schedule(dynamic, 8)
  for (int64_t f = 0; f < nf; ++f) {
    J jqf[4], jtf[3];
    for (int k = 0; k < 4; ++k) jqf[k] = J::variable(qf[f * 4 + k], k);
    for (int k = 0; k < 3; ++k) jtf[k] = J::variable(tf[f * 3 + k], 4 + k);
    double PJf[12];
    quaternion_plus_jacobian(qf + f * 4, PJf);
    double* Vv = &E->V[(size_t)f * 36];
    double* Wv = &E->W[(size_t)f * 6 * n6];
    double* gvv = &E->gv[(size_t)f * 6];
    double* P = &part[(size_t)f * ps];
    double cost = 0.0, sumsq = 0.0;
    for (int64_t o = D.foff[f]; o < D.foff[f + 1]; ++o) {
      const int c = D.cam[o];
      J jqc[4], jtc[3];
      for (int k = 0; k < 4; ++k) jqc[k] = J::variable(qc[c * 4 + k], 7 + k);
      for (int k = 0; k < 3; ++k) jtc[k] = J::variable(tc[c * 3 + k], 11 + k);
      double PJc[12];
      quaternion_plus_jacobian(qc + c * 4, PJc);
      const int64_t ip = D.pt[o];
      J Xw[3] = {J((double)D.Xw[ip * 3]), J((double)D.Xw[ip * 3 + 1]), J((double)D.Xw[ip * 3 + 2])};
      J r[2];
      residual_rig<J>(jqf, jtf, jqc, jtc, Xw, (double)D.u[o], (double)D.v[o], r);
      const double s = r[0].a * r[0].a + r[1].a * r[1].a;
      double rho = s, rho1 = 1.0;
      if (D.huber_a > 0.0) huber(D.huber_a, s, &rho, &rho1);
      const double w = std::sqrt(rho1);
      cost += 0.5 * rho;
      sumsq += s;
      double* Uc = P + (size_t)c * 42;
      double* gc = Uc + 36;
      for (int row = 0; row < 2; ++row) {
        double Jf[6], Jc[6];
        for (int k = 0; k < 3; ++k) {
          double a = 0.0, b = 0.0;
          for (int m = 0; m < 4; ++m) {
            a += r[row].v[m] * PJf[m * 3 + k];
            b += r[row].v[7 + m] * PJc[m * 3 + k];
          }
          Jf[k] = a * w;
          Jf[3 + k] = r[row].v[4 + k] * w;
          Jc[k] = b * w;
          Jc[3 + k] = r[row].v[11 + k] * w;
        }
        const double rr = r[row].a * w;
        if (res_out) res_out[2 * (o - D.foff[0]) + row] = rr;
        for (int a = 0; a < 6; ++a) {
          gvv[a] += Jf[a] * rr;
          gc[a] += Jc[a] * rr;
          for (int b = 0; b < 6; ++b) {
            Vv[a * 6 + b] += Jf[a] * Jf[b];
            Wv[a * n6 + 6 * c + b] += Jf[a] * Jc[b];
            Uc[a * 6 + b] += Jc[a] * Jc[b];
          }
        }
      }
    }
    P[(size_t)nc * 42] = cost;
    P[(size_t)nc * 42 + 1] = sumsq;
  }
  E->U.assign((size_t)n6 * n6, 0.0);
  E->gu.assign((size_t)n6, 0.0);
  E->cost = 0.0;
  E->sumsq = 0.0;
  for (int64_t f = 0; f < nf; ++f) {
    const double* P = &part[(size_t)f * ps];
    for (int c = 0; c < nc; ++c) {
      for (int a = 0; a < 6; ++a) {
        E->gu[6 * c + a] += P[(size_t)c * 42 + 36 + a];
        for (int b = 0; b < 6; ++b) E->U[(size_t)(6 * c + a) * n6 + 6 * c + b] += P[(size_t)c * 42 + a * 6 + b];
      }
    }
    E->cost += P[(size_t)nc * 42];
    E->sumsq += P[(size_t)nc * 42 + 1];
  }
}

static double cost_rig(const RigData& D, const double* qc, const double* tc, const double* qf,
                       const double* tf, double* obs_cost) {
  const int64_t nf = D.nf;
  std::vector<double> part((size_t)nf, 0.0);
#pragma omp parallel for schedule(dynamic, 32)
  for (int64_t f = 0; f < nf; ++f) {
    double cst = 0.0;
    for (int64_t o = D.foff[f]; o < D.foff[f + 1]; ++o) {
      const int c = D.cam[o];
      const int64_t ip = D.pt[o];
      const double Xw[3] = {(double)D.Xw[ip * 3], (double)D.Xw[ip * 3 + 1], (double)D.Xw[ip * 3 + 2]};
      double r[2];
      residual_rig<double>(qf + f * 4, tf + f * 3, qc + c * 4, tc + c * 3, Xw, (double)D.u[o],
                           (double)D.v[o], r);
      const double s = r[0] * r[0] + r[1] * r[1];
      double rho = s, rho1 = 1.0;
      if (D.huber_a > 0.0) huber(D.huber_a, s, &rho, &rho1);
      cst += 0.5 * rho;
      if (obs_cost) obs_cost[o - D.foff[0]] = 0.5 * rho;
    }
    part[(size_t)f] = cst;
  }
  double c = 0.0;
  for (int64_t f = 0; f < nf; ++f) c += part[(size_t)f];
  return c;
}

static double grad_norm_rig(const RigData& D, const RigEval& E, const uint8_t* frozen) {
  const int n6 = 6 * D.nc;
  double s = 0.0;
  for (int k = 0; k < n6; ++k) {
    if (frozen && frozen[k / 6]) continue;
    const double d = E.U[(size_t)k * n6 + k];
    if (d > 0.0) s += E.gu[k] * E.gu[k] / d;
  }
  for (int64_t f = 0; f < D.nf; ++f)
    for (int k = 0; k < 6; ++k) {
      const double d = E.V[(size_t)f * 36 + k * 6 + k];
      const double g = E.gv[(size_t)f * 6 + k];
      if (d > 0.0) s += g * g / d;
    }
  return std::sqrt(s);
}

static int solve_rig(const RigData& D, double* qc, double* tc, const uint8_t* frozen, double* qf,
                     double* tf, double* obs_cost, const oracle_opts& O, oracle_report* R) {
  const int nc = D.nc, n6 = 6 * nc;
  const int64_t nf = D.nf, nobs = D.nobs();
  const double meas_scale = std::max(max_abs(D.u + D.foff[0], nobs), max_abs(D.v + D.foff[0], nobs));
  RigEval E;
  eval_rig(D, qc, tc, qf, tf, &E, nullptr);
  oracle_report rep;
  std::memset(&rep, 0, sizeof(rep));
  rep.sweeps = 1;
  rep.initial_cost = E.cost;
  rep.termination = 4;
  double radius = O.initial_radius, nu = 2.0;
  int blind = 0;  // sub-resolution steps taken so far
  std::vector<double> Y((size_t)nf * 6 * n6), z((size_t)nf * 6), dv((size_t)nf * 6), Dv((size_t)nf * 6);
  std::vector<double> S((size_t)n6 * n6), rhs(n6), Du(n6), du(n6);
  std::vector<double> qcc((size_t)nc * 4), tcc((size_t)nc * 3), qfc((size_t)nf * 4), tfc((size_t)nf * 3);
  if (!std::isfinite(E.cost)) {
    rep.termination = -2;
    if (R) *R = rep;
    return -2;
  }
  for (int it = 0; it < O.max_iterations; ++it) {
    const double rnorm = std::sqrt(2.0 * E.cost);
    const double gn = grad_norm_rig(D, E, frozen);
    if (E.cost == 0.0 || gn <= O.gradient_tolerance * rnorm) {
      rep.termination = 1;
      break;
    }
    rep.iterations++;
    const double lambda = 1.0 / radius;
    for (int a = 0; a < n6; ++a) {
      Du[a] = clampd(E.U[(size_t)a * n6 + a], O.min_diagonal, O.max_diagonal);
      rhs[a] = -E.gu[a];
      for (int b = 0; b < n6; ++b) S[(size_t)a * n6 + b] = E.U[(size_t)a * n6 + b];
      S[(size_t)a * n6 + a] += lambda * Du[a];
    }
    bool ok = true;
    std::vector<double> col(6);
    for (int64_t f = 0; f < nf && ok; ++f) {
      double L[36];
      std::memcpy(L, &E.V[(size_t)f * 36], sizeof(L));
      for (int k = 0; k < 6; ++k) {
        Dv[(size_t)f * 6 + k] = clampd(L[k * 6 + k], O.min_diagonal, O.max_diagonal);
        L[k * 6 + k] += lambda * Dv[(size_t)f * 6 + k];
      }
      if (!chol(L, 6)) {
        ok = false;
        break;
      }
      const double* Wv = &E.W[(size_t)f * 6 * n6];
      double* Yv = &Y[(size_t)f * 6 * n6];
      double* zv = &z[(size_t)f * 6];
      for (int b = 0; b < n6; ++b) {
        for (int a = 0; a < 6; ++a) col[a] = Wv[a * n6 + b];
        chol_solve(L, 6, col.data());
        for (int a = 0; a < 6; ++a) Yv[a * n6 + b] = col[a];
      }
      for (int a = 0; a < 6; ++a) zv[a] = E.gv[(size_t)f * 6 + a];
      chol_solve(L, 6, zv);
      for (int a = 0; a < n6; ++a) {
        double s = 0.0;
        for (int k = 0; k < 6; ++k) s += Wv[k * n6 + a] * zv[k];
        rhs[a] += s;
        for (int b = 0; b < n6; ++b) {
          double w = 0.0;
          for (int k = 0; k < 6; ++k) w += Wv[k * n6 + a] * Yv[k * n6 + b];
          S[(size_t)a * n6 + b] -= w;
        }
      }
    }
    if (ok) {
      for (int k = 0; k < n6; ++k)
        if (frozen && frozen[k / 6]) {
          for (int b = 0; b < n6; ++b) S[(size_t)k * n6 + b] = S[(size_t)b * n6 + k] = 0.0;
          S[(size_t)k * n6 + k] = 1.0;
          rhs[k] = 0.0;
        }
      ok = chol(S.data(), n6);
    }
    if (!ok) {
      rep.rejected++;
      radius /= nu;
      nu *= 2.0;
      if (radius < O.min_radius) {
        rep.termination = 5;
        break;
      }
      continue;
    }
    du = rhs;
    chol_solve(S.data(), n6, du.data());
    double gTd = 0.0, dDd = 0.0;
    for (int k = 0; k < n6; ++k) {
      if (frozen && frozen[k / 6]) continue;
      gTd += E.gu[k] * du[k];
      dDd += du[k] * du[k] * Du[k];
    }
    for (int64_t f = 0; f < nf; ++f)
      for (int a = 0; a < 6; ++a) {
        double s = -z[(size_t)f * 6 + a];
        for (int b = 0; b < n6; ++b) s -= Y[(size_t)f * 6 * n6 + a * n6 + b] * du[b];
        dv[(size_t)f * 6 + a] = s;
        gTd += E.gv[(size_t)f * 6 + a] * s;
        dDd += s * s * Dv[(size_t)f * 6 + a];
      }
    const double model = 0.5 * (lambda * dDd - gTd);
    const bool tiny = model <= cost_noise(O, meas_scale, rnorm) ||
                      std::sqrt(dDd) <= O.parameter_tolerance * rnorm;  // see solve_single
    for (int c = 0; c < nc; ++c) {
      quaternion_plus(qc + c * 4, &du[6 * c], &qcc[(size_t)c * 4]);
      for (int k = 0; k < 3; ++k) tcc[(size_t)c * 3 + k] = tc[c * 3 + k] + du[6 * c + 3 + k];
    }
    for (int64_t f = 0; f < nf; ++f) {
      quaternion_plus(qf + f * 4, &dv[(size_t)f * 6], &qfc[(size_t)f * 4]);
      for (int k = 0; k < 3; ++k) tfc[(size_t)f * 3 + k] = tf[f * 3 + k] + dv[(size_t)f * 6 + 3 + k];
    }
    const double cc = cost_rig(D, qcc.data(), tcc.data(), qfc.data(), tfc.data(), nullptr);
    const double rho = (E.cost - cc) / model;
    if (O.verbose)
      std::printf("[oracle-rig] it %3d cost %.17g cand %.17g rho %.3e radius %.3e |g| %.3e |d| %.3e\n",
                  it, E.cost, cc, rho, radius, gn / rnorm, std::sqrt(dDd) / rnorm);
    if (tiny || blind > 0) {
      bool took = false;
      if (std::isfinite(cc) && cc <= E.cost * (1.0 + kNoise) + 4.0 * cost_noise(O, meas_scale, rnorm)) {
        std::memcpy(qc, qcc.data(), sizeof(double) * (size_t)nc * 4);
        std::memcpy(tc, tcc.data(), sizeof(double) * (size_t)nc * 3);
        std::memcpy(qf, qfc.data(), sizeof(double) * (size_t)nf * 4);
        std::memcpy(tf, tfc.data(), sizeof(double) * (size_t)nf * 3);
        rep.accepted++;
        eval_rig(D, qc, tc, qf, tf, &E, nullptr);
        rep.sweeps++;
        took = true;
      } else {
        rep.rejected++;
      }
      ++blind;
      if (!took || blind >= kBlindSteps) {
        rep.termination = 3;
        break;
      }
      continue;
    }
    if (std::isfinite(cc) && model > 0.0 && rho > O.min_relative_decrease) {
      const double old = E.cost;
      std::memcpy(qc, qcc.data(), sizeof(double) * (size_t)nc * 4);
      std::memcpy(tc, tcc.data(), sizeof(double) * (size_t)nc * 3);
      std::memcpy(qf, qfc.data(), sizeof(double) * (size_t)nf * 4);
      std::memcpy(tf, tfc.data(), sizeof(double) * (size_t)nf * 3);
      rep.accepted++;
      const double fq = 2.0 * rho - 1.0;
      radius = std::min(O.max_radius, radius / std::max(1.0 / 3.0, 1.0 - fq * fq * fq));
      nu = 2.0;
      eval_rig(D, qc, tc, qf, tf, &E, nullptr);
      rep.sweeps++;
      if (old - E.cost <= O.function_tolerance * old) {
        rep.termination = 2;
        break;
      }
    } else {
      rep.rejected++;
      radius /= nu;
      nu *= 2.0;
      if (radius < O.min_radius) {
        rep.termination = 5;
        break;
      }
    }
  }
  rep.final_cost = E.cost;
  const double rnorm = std::sqrt(2.0 * E.cost);
  rep.gradient_norm = rnorm > 0.0 ? grad_norm_rig(D, E, frozen) / rnorm : 0.0;
  rep.rms = nobs > 0 ? std::sqrt(E.sumsq / (double)nobs) : 0.0;
  if (obs_cost) cost_rig(D, qc, tc, qf, tf, obs_cost);
  if (R) *R = rep;
  return 0;
}

}  // namespace oracle

// =========================================================================
//                               C interface
// =========================================================================
using namespace oracle;

extern "C" {

void oracle_set_h_float(int on) { oracle::g_h_float = on; }
void oracle_set_g_float(int mode) { oracle::g_g_float = mode; }

void oracle_default_opts(oracle_opts* o) {
  o->max_iterations = 100;
  o->verbose = 0;
  o->initial_radius = 1e4;
  o->max_radius = 1e16;
  o->min_radius = 1e-32;
  o->min_relative_decrease = 1e-3;
  o->min_diagonal = 1e-6;
  o->max_diagonal = 1e32;
  o->function_tolerance = 1e-16;
  o->gradient_tolerance = 1e-12;
  o->parameter_tolerance = 1e-9;
  o->noise_factor = 4.0;
  o->huber_a = (double)(3.0f / 500.0f);  // extrinsics_calibrator.cpp:176 (float literal)
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int oracle_intrinsics_solve(const float* u, const float* v, const float* X, const float* Y,
                            const float* Z, const int64_t* view_offsets, int64_t n_views,
                            double* intr, double* q, double* t, uint32_t const_intr_mask,
                            const oracle_opts* opts, oracle_report* report) {
  oracle_opts O;
  if (opts) O = *opts; else oracle_default_opts(&O);
  SingleData D{u, v, X, Y, Z, view_offsets, n_views};
  return solve_single(D, intr, q, t, const_intr_mask, O, report);
}

int oracle_intrinsics_eval(const float* u, const float* v, const float* X, const float* Y,
                           const float* Z, const int64_t* view_offsets, int64_t n_views,
                           const double* intr, const double* q, const double* t, double* res,
                           double* Vb, double* Wb, double* gvb, double* U, double* gu,
                           double* cost) {
  SingleData D{u, v, X, Y, Z, view_offsets, n_views};
  SingleEval E;
  eval_single(D, intr, q, t, &E, res);
  if (Vb) std::memcpy(Vb, E.V.data(), sizeof(double) * E.V.size());
  if (Wb) std::memcpy(Wb, E.W.data(), sizeof(double) * E.W.size());
  if (gvb) std::memcpy(gvb, E.gv.data(), sizeof(double) * E.gv.size());
  if (U) std::memcpy(U, E.U, sizeof(E.U));
  if (gu) std::memcpy(gu, E.gu, sizeof(E.gu));
  if (cost) *cost = E.cost;
  return 0;
}

int oracle_intrinsics_solve_batch(const float* u, const float* v, const float* X, const float* Y,
                                  const float* Z, const int64_t* view_offsets,
                                  const int64_t* problem_offsets, int64_t n_problems, double* intr,
                                  double* q, double* t, uint32_t const_intr_mask,
                                  const oracle_opts* opts, oracle_report* reports) {
  oracle_opts O;
  if (opts) O = *opts; else oracle_default_opts(&O);
  int rc = 0;
  // problems are independent: run them one after another (each one is itself
  // parallel over views); the reference would loop over Calibrator objects.
  for (int64_t b = 0; b < n_problems; ++b) {
    const int64_t v0 = problem_offsets[b], v1 = problem_offsets[b + 1];
    SingleData D{u, v, X, Y, Z, view_offsets + v0, v1 - v0};
    const int r = solve_single(D, intr + 9 * b, q + 4 * v0, t + 3 * v0, const_intr_mask, O,
                               reports ? reports + b : nullptr);
    if (r != 0) rc = r;
  }
  return rc;
}

int oracle_rig_solve(const float* u, const float* v, const int32_t* obs_cam,
                     const int64_t* obs_point, const int64_t* frame_offsets, int64_t n_frames,
                     const float* Xw, int64_t n_points, int32_t n_cams, double* q_c, double* t_c,
                     const uint8_t* cam_frozen, double* q_f, double* t_f, double* obs_cost,
                     const oracle_opts* opts, oracle_report* report) {
  (void)n_points;
  oracle_opts O;
  if (opts) O = *opts; else { oracle_default_opts(&O); O.max_iterations = 1000; }
  RigData D{u, v, obs_cam, obs_point, frame_offsets, n_frames, Xw, n_cams, O.huber_a};
  return solve_rig(D, q_c, t_c, cam_frozen, q_f, t_f, obs_cost, O, report);
}

int oracle_rig_eval(const float* u, const float* v, const int32_t* obs_cam,
                    const int64_t* obs_point, const int64_t* frame_offsets, int64_t n_frames,
                    const float* Xw, int64_t n_points, int32_t n_cams, const double* q_c,
                    const double* t_c, const double* q_f, const double* t_f, double huber_a,
                    double* res, double* Vb, double* Wb, double* gvb, double* U, double* gu,
                    double* cost) {
  (void)n_points;
  RigData D{u, v, obs_cam, obs_point, frame_offsets, n_frames, Xw, n_cams, huber_a};
  RigEval E;
  eval_rig(D, q_c, t_c, q_f, t_f, &E, res);
  if (Vb) std::memcpy(Vb, E.V.data(), sizeof(double) * E.V.size());
  if (Wb) std::memcpy(Wb, E.W.data(), sizeof(double) * E.W.size());
  if (gvb) std::memcpy(gvb, E.gv.data(), sizeof(double) * E.gv.size());
  if (U) std::memcpy(U, E.U.data(), sizeof(double) * E.U.size());
  if (gu) std::memcpy(gu, E.gu.data(), sizeof(double) * E.gu.size());
  if (cost) *cost = E.cost;
  return 0;
}

}  // extern "C"
