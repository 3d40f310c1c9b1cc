// Single-camera hot path: residual + analytic Jacobian + normal-equation
// blocks + Schur elimination of the per-view poses + device-resident LM loop.
//
// Replaces, behind ccg_intrinsics_*:
//   the per-observation autodiff evaluation of ReprojectionError
//       reference src/calibrator.cpp:183-219 (model :70-95, layout :168-179)
//   the Ceres problem build and ceres::Solve call
//       reference src/calibrator.cpp:236-324
// Design:
//   * observations live in HBM as SoA float32 (u,v,X,Y,Z: 20 B/obs); per-view pose and blocks in FP64
//     (SoA as uploaded, and re-packed into 640-byte blocks of 32 for the default sweep)
//   * two sweep variants behind ccg_opts.hessian_fp64, both with FP64 residual, cost and gradient:
//       0 (default) k_intr_sweep_ws, intr_sweep.cuh: warp-specialised, J'J on the TF32 tensor cores
//       1           k_intr_sweep below: everything FP64, a warp PAIR per view (warp A accumulates
//                   V(21) g_v(6) W[:,fx fy px py](24) + cost, warp B W[:,k1 k2 p1 p2 k3](30) + U(35) +
//                   g_u(9); lane-private accumulators, one transposing shuffle reduction per view)
//   * the 9x9 reduced system is summed from per-CTA partials in a fixed order, so results are
//     bit-reproducible for a given launch geometry (and bit-identical on every rank of a partition)
//   * accept/reject, trust-region update and termination run on the device; the host replays one LM
//     iteration as a CUDA graph and polls one counter
//   * Calibrator::Estimate's closed-form initialisation runs on the device too (intr_init.cuh)
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace ccg {

constexpr int kBlk = 84;        // per-view block stride (doubles): V21 | gv6 | W4 24 | W5 30 | pad 3
constexpr int kYz = 72;         // per-view elimination record: Y 54 | z 6 | gv 6 | Dv 6
constexpr int kPose = 8;        // q4 t3 pad
constexpr int kSweepPart = 64;  // per-item sweep partial (doubles)
constexpr int kSchurPart = 64;  // per-item elimination partial: S upper 45 | rhs 9 | g'D^-1g | fails
constexpr int kRedSeg = 8;      // segments of the fixed-order reduction of the partials (block_reduce_items)
constexpr int kWarps = 8;       // warps per CTA (4 view pairs in the sweep, 8 views in the elimination)

// sweep partial layout
//  [0] sum r^2   [1] U fxfx [2] U fxpx [3] U pxpx(=n) [4] U fyfy [5] U fypy [6] U pypy
//  [7] g'd (poses) [8] d'Dd (poses)                                   <- warp A, slots 0..15
//  [16..24] g_u (fx fy px py k1 k2 p1 p2 k3)
//  [25..29] U fx,d  [30..34] U px,d  [35..39] U fy,d  [40..44] U py,d  [45..59] U d,d (packed upper)
//                                                                      <- warp B, slots 16..63

struct IntrState {
  LmState lm;
  double intr[2][9];
  double du[9];
  double U[81];  // at the current point, full symmetric
  double gu[9];
  int item0, item1;
  unsigned mask;
  int pad_;
};

struct IntrDev {
  const float *u, *v, *X, *Y, *Z;
  const float* obs_blk;  // the same observations as 640-byte blocks of 32: u[32] v[32] X[32] Y[32] Z[32]
  const int2* vmeta;     // per view (+ sentinel): first block, observations
  const int64_t* view_off;
  const WorkItem* items;
  double* pose[2];
  double* blk[2];
  double* yz;
  double* rt;         // candidate pose as rotation matrix + translation, 12 doubles per view (default sweep, FP64 role)
  float* rtf;         // the same in float32 (Jacobian role)
  double* sweep_part;
  double* schur_part;
  double* red_sweep;  // [B][64]
  double* red_schur;  // [B][96]
  IntrState* ps;
  int* n_done;
  int64_t obs_base;   // view_off[0]
  int schur_split;    // CTAs of the elimination kernel per work item
};

// --------------------------------------------------------------------------
// per-observation geometry shared by both warps of a pair
// --------------------------------------------------------------------------
struct ObsGeom {
  double xn, yn, r2, r4, r6, a1, a2, a3, xd, yd;
  double ju[6], jv[6];
};

__device__ __forceinline__ void obs_geometry(const double (&R)[9], const double (&t)[3],
                                             const double (&k)[9], float Xf, float Yf, float Zf,
                                             ObsGeom& g) {
  const double X = (double)Xf, Y = (double)Yf, Z = (double)Zf;
  // rotated point (before translation) and camera-frame point
  const double y0 = R[0] * X + R[1] * Y + R[2] * Z;
  const double y1 = R[3] * X + R[4] * Y + R[5] * Z;
  const double y2 = R[6] * X + R[7] * Y + R[8] * Z;
  const double iz = fast_rcp(y2 + t[2]);
  const double xn = (y0 + t[0]) * iz, yn = (y1 + t[1]) * iz;
  const double fx = k[0], fy = k[1], k1 = k[4], k2 = k[5], p1 = k[6], p2 = k[7], k3 = k[8];
  const double r2 = xn * xn + yn * yn, r4 = r2 * r2, r6 = r4 * r2;
  const double m = 1.0 + k1 * r2 + k2 * r4 + k3 * r6;
  const double a1 = 2.0 * xn * yn, a2 = r2 + 2.0 * xn * xn, a3 = r2 + 2.0 * yn * yn;
  g.xn = xn; g.yn = yn; g.r2 = r2; g.r4 = r4; g.r6 = r6; g.a1 = a1; g.a2 = a2; g.a3 = a3;
  g.xd = xn * m + p1 * a1 + p2 * a2;
  g.yd = yn * m + p2 * a1 + p1 * a3;
  // d(xd,yd)/d(xn,yn) (symmetric off-diagonal)
  const double mp = k1 + 2.0 * k2 * r2 + 3.0 * k3 * r4;
  const double F00 = m + 2.0 * xn * xn * mp + 2.0 * p1 * yn + 6.0 * p2 * xn;
  const double F01 = a1 * mp + 2.0 * p1 * xn + 2.0 * p2 * yn;
  const double F11 = m + 2.0 * yn * yn * mp + 2.0 * p2 * xn + 6.0 * p1 * yn;
  // d(xn,yn)/d(pose tangent): rotation part uses d y/d delta = -2 [y]x (half-angle tangent)
  const double A0 = fx * iz * F00, A1 = fx * iz * F01, B0 = fy * iz * F01, B1 = fy * iz * F11;
  const double c00 = -2.0 * xn * y1, c01 = 2.0 * (y2 + xn * y0), c02 = -2.0 * y1;
  const double c10 = -2.0 * (y2 + yn * y1), c11 = 2.0 * yn * y0, c12 = 2.0 * y0;
  g.ju[0] = A0 * c00 + A1 * c10;
  g.ju[1] = A0 * c01 + A1 * c11;
  g.ju[2] = A0 * c02 + A1 * c12;
  g.ju[3] = A0;
  g.ju[4] = A1;
  g.ju[5] = -(A0 * xn + A1 * yn);
  g.jv[0] = B0 * c00 + B1 * c10;
  g.jv[1] = B0 * c01 + B1 * c11;
  g.jv[2] = B0 * c02 + B1 * c12;
  g.jv[3] = B0;
  g.jv[4] = B1;
  g.jv[5] = -(B0 * xn + B1 * yn);
}

// one observation's five float32 fields, fetched ahead of use
struct ObsLoad {
  float mu, mv, X, Y, Z;
  int64_t o;
  bool valid;
  __device__ __forceinline__ void fetch(const IntrDev& d, int64_t idx, int64_t end) {
    o = idx;
    valid = idx < end;
    if (valid) {
      mu = __ldg(d.u + idx);
      mv = __ldg(d.v + idx);
      X = __ldg(d.X + idx);
      Y = __ldg(d.Y + idx);
      Z = __ldg(d.Z + idx);
    }
  }
};

// Candidate pose of a view: current pose (+) delta_v, delta_v = -z - Y du.
// Returns R, t of the candidate; lanes 0..5 of warp A also return their share
// of g'd and d'Dd; warp A stores the candidate pose.
__device__ __forceinline__ void view_prologue(const IntrDev& d, int par, int64_t v, int lane,
                                              bool is_a, const double* s_du, double (&R)[9],
                                              double (&t)[3], double& gTd, double& dDd) {
  const double* yz = d.yz + (size_t)v * kYz;
  double dl = 0.0;
  if (lane < 6) {
    dl = -yz[54 + lane];
#pragma unroll
    for (int j = 0; j < 9; ++j) dl -= yz[lane * 9 + j] * s_du[j];
    if (is_a) {
      gTd += yz[60 + lane] * dl;
      dDd += dl * dl * yz[66 + lane];
    }
  }
  const double d0 = __shfl_sync(0xffffffffu, dl, 0), d1 = __shfl_sync(0xffffffffu, dl, 1),
               d2 = __shfl_sync(0xffffffffu, dl, 2), d3 = __shfl_sync(0xffffffffu, dl, 3),
               d4 = __shfl_sync(0xffffffffu, dl, 4), d5 = __shfl_sync(0xffffffffu, dl, 5);
  const double* pc = d.pose[par] + (size_t)v * kPose;
  double q[4] = {pc[0], pc[1], pc[2], pc[3]};
  double qn[4];
  quat_plus(q, d0, d1, d2, qn);
  t[0] = pc[4] + d3;
  t[1] = pc[5] + d4;
  t[2] = pc[6] + d5;
  if (is_a && lane < 7) {
    double* pn = d.pose[par ^ 1] + (size_t)v * kPose;
    const double val = lane == 0 ? qn[0] : lane == 1 ? qn[1] : lane == 2 ? qn[2] : lane == 3 ? qn[3]
                     : lane == 4 ? t[0] : lane == 5 ? t[1] : t[2];
    pn[lane] = val;
  }
  quat_to_rot(qn, R);
}

// --------------------------------------------------------------------------
// sweep: residual + Jacobian + normal-equation blocks at the candidate point
// --------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarps * 32, 1) k_intr_sweep(IntrDev d, double* __restrict__ res_out) {
  __shared__ double s_part[kWarps][kSweepPart];
  __shared__ double s_k[9], s_du[9];
  __shared__ int s_meta[2];
  const WorkItem it = d.items[blockIdx.x];
  IntrState& ps = d.ps[it.problem];
  if (threadIdx.x == 0) {
    s_meta[0] = ps.lm.par;
    s_meta[1] = (ps.lm.status == 0 && ps.lm.solve_ok) ? 1 : 0;
  }
  __syncthreads();
  if (!s_meta[1]) return;
  const int par = s_meta[0];
  if (threadIdx.x < 9) {
    s_k[threadIdx.x] = ps.intr[par ^ 1][threadIdx.x];
    s_du[threadIdx.x] = ps.du[threadIdx.x];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = warp >> 1;
  const bool is_a = (warp & 1) == 0;
  double k[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) k[i] = s_k[i];

  if (is_a) {
    // ---------------- warp A: V, g_v, W[:,0..3], cost, U(fx,px,fy,py block)
    double pa[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) pa[i] = 0.0;
    for (int64_t v = it.v0 + pair; v < it.v1; v += kWarps / 2) {
      double R[9], t[3];
      view_prologue(d, par, v, lane, true, s_du, R, t, pa[7], pa[8]);
      double acc[51];
#pragma unroll
      for (int i = 0; i < 51; ++i) acc[i] = 0.0;
      const int64_t o1 = d.view_off[v + 1];
      // software-pipelined: the loads of the next 32 observations are in flight while
      // this lane works on the current one (8 warps per SM cannot hide HBM latency alone)
      ObsLoad nx;
      nx.fetch(d, d.view_off[v] + lane, o1);
      while (nx.valid) {
        const ObsLoad cur = nx;
        const int64_t o = cur.o;
        nx.fetch(d, o + 32, o1);
        const float mu = cur.mu, mv = cur.mv;
        ObsGeom g;
        obs_geometry(R, t, k, cur.X, cur.Y, cur.Z, g);
        const double ru = k[0] * g.xd + k[2] - (double)mu;
        const double rv = k[1] * g.yd + k[3] - (double)mv;
        if (res_out) {
          res_out[2 * (o - d.obs_base)] = ru;
          res_out[2 * (o - d.obs_base) + 1] = rv;
        }
        int idx = 0;
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
          for (int b = a; b < 6; ++b) {
            acc[idx] = fma(g.ju[a], g.ju[b], fma(g.jv[a], g.jv[b], acc[idx]));
            ++idx;
          }
#pragma unroll
        for (int a = 0; a < 6; ++a) {
          acc[21 + a] = fma(g.ju[a], ru, fma(g.jv[a], rv, acc[21 + a]));
          acc[27 + a * 4 + 0] = fma(g.ju[a], g.xd, acc[27 + a * 4 + 0]);
          acc[27 + a * 4 + 1] = fma(g.jv[a], g.yd, acc[27 + a * 4 + 1]);
          acc[27 + a * 4 + 2] += g.ju[a];
          acc[27 + a * 4 + 3] += g.jv[a];
        }
        pa[0] = fma(ru, ru, fma(rv, rv, pa[0]));
        pa[1] = fma(g.xd, g.xd, pa[1]);
        pa[2] += g.xd;
        pa[3] += 1.0;
        pa[4] = fma(g.yd, g.yd, pa[4]);
        pa[5] += g.yd;
        pa[6] += 1.0;
      }
      double out[2];
      warp_reduce_array<51>(acc, lane, out);
      double* blk = d.blk[par ^ 1] + (size_t)v * kBlk;
      blk[lane] = out[0];
      if (lane < 51 - 32) blk[32 + lane] = out[1];
    }
    // g'd and d'Dd are spread over lanes 0..5; the transposing reduction sums each slot over all lanes
    double out[1];
    warp_reduce_array<9>(pa, lane, out);
    if (lane < 16) s_part[warp][lane] = lane < 9 ? out[0] : 0.0;
  } else {
    // ---------------- warp B: W[:,4..8], g_u, U (everything involving distortion)
    double pb[44];
#pragma unroll
    for (int i = 0; i < 44; ++i) pb[i] = 0.0;
    for (int64_t v = it.v0 + pair; v < it.v1; v += kWarps / 2) {
      double R[9], t[3], dum0 = 0.0, dum1 = 0.0;
      view_prologue(d, par, v, lane, false, s_du, R, t, dum0, dum1);
      double acc[30];
#pragma unroll
      for (int i = 0; i < 30; ++i) acc[i] = 0.0;
      const int64_t o1 = d.view_off[v + 1];
      // software-pipelined: the loads of the next 32 observations are in flight while
      // this lane works on the current one (8 warps per SM cannot hide HBM latency alone)
      ObsLoad nx;
      nx.fetch(d, d.view_off[v] + lane, o1);
      while (nx.valid) {
        const ObsLoad cur = nx;
        const int64_t o = cur.o;
        nx.fetch(d, o + 32, o1);
        const float mu = cur.mu, mv = cur.mv;
        ObsGeom g;
        obs_geometry(R, t, k, cur.X, cur.Y, cur.Z, g);
        const double ru = k[0] * g.xd + k[2] - (double)mu;
        const double rv = k[1] * g.yd + k[3] - (double)mv;
        double au[5], av[5];
        au[0] = k[0] * g.xn * g.r2; au[1] = k[0] * g.xn * g.r4; au[2] = k[0] * g.a1;
        au[3] = k[0] * g.a2;        au[4] = k[0] * g.xn * g.r6;
        av[0] = k[1] * g.yn * g.r2; av[1] = k[1] * g.yn * g.r4; av[2] = k[1] * g.a3;
        av[3] = k[1] * g.a1;        av[4] = k[1] * g.yn * g.r6;
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
          for (int c = 0; c < 5; ++c)
            acc[a * 5 + c] = fma(g.ju[a], au[c], fma(g.jv[a], av[c], acc[a * 5 + c]));
        pb[0] = fma(g.xd, ru, pb[0]);
        pb[1] = fma(g.yd, rv, pb[1]);
        pb[2] += ru;
        pb[3] += rv;
        int idx = 29;
#pragma unroll
        for (int c = 0; c < 5; ++c) {
          pb[4 + c] = fma(au[c], ru, fma(av[c], rv, pb[4 + c]));
          pb[9 + c] = fma(g.xd, au[c], pb[9 + c]);
          pb[14 + c] += au[c];
          pb[19 + c] = fma(g.yd, av[c], pb[19 + c]);
          pb[24 + c] += av[c];
#pragma unroll
          for (int e = c; e < 5; ++e) {
            pb[idx] = fma(au[c], au[e], fma(av[c], av[e], pb[idx]));
            ++idx;
          }
        }
      }
      double out[1];
      warp_reduce_array<30>(acc, lane, out);
      if (lane < 30) d.blk[par ^ 1][(size_t)v * kBlk + 51 + lane] = out[0];
    }
    double out[2];
    warp_reduce_array<44>(pb, lane, out);
    s_part[warp][16 + lane] = out[0];
    if (lane < 16) s_part[warp][48 + lane] = lane < 12 ? out[1] : 0.0;
  }
  __syncthreads();
  if (threadIdx.x < kSweepPart) {
    const int c = threadIdx.x;
    const int w0 = c < 16 ? 0 : 1;  // slots 0..15 come from A warps, the rest from B warps
    double s = 0.0;
#pragma unroll
    for (int p = 0; p < kWarps / 2; ++p) s += s_part[2 * p + w0][c];
    d.sweep_part[(size_t)blockIdx.x * kSweepPart + c] = s;
  }
}

}  // namespace ccg
#include "intr_sweep.cuh"
#include "intr_init.cuh"
namespace ccg {

// --------------------------------------------------------------------------
// pose elimination, one THREAD per view: Cholesky of V + lambda D, T = L^-1 [W g_v],
// Y = L^-T T (stored for the back-substitution), contributions T'T and T't_g to the
// reduced system.  A lane reads its view's block (672 B, contiguous) straight into registers with
// 16-byte loads - all 41 in flight at once, the neighbouring half of every 32-byte sector is an L1
// hit for the next load - and writes the record the same way; the 56 reduced-system terms are summed
// across lanes with the transposing shuffle reduction.
// --------------------------------------------------------------------------
constexpr int kSchurWarps = 2;
static_assert(kBlk % 2 == 0 && kYz % 2 == 0, "16-byte rows");

// packed index of (j,k), j <= k, in the 9x9 upper triangle
__host__ __device__ constexpr int up9(int j, int k) { return j * 9 - j * (j - 1) / 2 + (k - j); }

// One warp eliminates the (up to 32) views base .. base + nv - 1, a lane per view; racc0 / racc1 accumulate the
// lane's two reduced-system terms (terms L and 32 + L of the 64-value partial).
__device__ __forceinline__ void schur_warp_pass(const IntrDev& d, const LmParams& P, int par, double lambda, int64_t base,
                                                int nv, int lane, double& racc0, double& racc1) {
  const bool has = lane < nv;
  // (an idle lane computes on a harmless unit block and contributes zeros)
  double s[kBlk];
#pragma unroll
  for (int i = 0; i < kBlk; ++i) s[i] = 0.0;
  s[0] = s[6] = s[11] = s[15] = s[18] = s[20] = 1.0;
  if (has) {
    const double2* src = reinterpret_cast<const double2*>(d.blk[par] + (size_t)(base + lane) * kBlk);
#pragma unroll
    for (int i = 0; i < 41; ++i) {
      const double2 t = src[i];
      s[2 * i] = t.x;
      s[2 * i + 1] = t.y;
    }
  }
  double A[6][6], Dv[6], tg[6], gv[6];
  {
    int idx = 0;
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
      for (int b = a; b < 6; ++b) {
        A[b][a] = s[idx];
        ++idx;
      }
  }
  double gDg = 0.0;
#pragma unroll
  for (int a = 0; a < 6; ++a) {
    tg[a] = gv[a] = s[21 + a];
    Dv[a] = clampd(A[a][a], P.min_diagonal, P.max_diagonal);
    if (A[a][a] > 0.0) gDg += tg[a] * tg[a] / A[a][a];
    A[a][a] += lambda * Dv[a];
  }
  const bool ok = chol6(A);
  double T[6][9];
#pragma unroll
  for (int a = 0; a < 6; ++a) {
#pragma unroll
    for (int c = 0; c < 4; ++c) T[a][c] = s[27 + a * 4 + c];
#pragma unroll
    for (int c = 0; c < 5; ++c) T[a][4 + c] = s[51 + a * 5 + c];
  }
  // forward substitution: T <- L^-1 W, tg <- L^-1 g_v
  double inv[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) inv[i] = 1.0 / A[i][i];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
#pragma unroll
    for (int k = 0; k < i; ++k) {
#pragma unroll
      for (int c = 0; c < 9; ++c) T[i][c] -= A[i][k] * T[k][c];
      tg[i] -= A[i][k] * tg[k];
    }
#pragma unroll
    for (int c = 0; c < 9; ++c) T[i][c] *= inv[i];
    tg[i] *= inv[i];
  }
  // reduced-system terms: (W' V'^-1 W)[j][k] = T[:,j].T[:,k] (packed upper, 45), (W' V'^-1 g)[j] = T[:,j].tg (9),
  // g'D^-1g, failures - summed over the warp's lanes 32 terms at a time (fewer live registers)
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    double vals[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) vals[i] = 0.0;
#pragma unroll
    for (int j = 0; j < 9; ++j) {
#pragma unroll
      for (int k = j; k < 9; ++k) {
        const int e = up9(j, k) - 32 * half;
        if (e >= 0 && e < 32) {
          double acc = 0.0;
#pragma unroll
          for (int a = 0; a < 6; ++a) acc = fma(T[a][j], T[a][k], acc);
          vals[e] = acc;
        }
      }
      const int e = 45 + j - 32 * half;
      if (e >= 0 && e < 32) {
        double acc = 0.0;
#pragma unroll
        for (int a = 0; a < 6; ++a) acc = fma(T[a][j], tg[a], acc);
        vals[e] = acc;
      }
    }
    if (half == 1) {
      vals[54 - 32] = gDg;
      vals[55 - 32] = ok ? 0.0 : 1.0;
    }
    if (!has) {
#pragma unroll
      for (int i = 0; i < 32; ++i) vals[i] = 0.0;
    }
    const double r = warp_reduce32(vals, lane);
    if (half == 0) racc0 += r;
    else racc1 += r;
  }
  // backward substitution: Y = L^-T T, z = L^-T tg
#pragma unroll
  for (int i = 5; i >= 0; --i) {
#pragma unroll
    for (int k = i + 1; k < 6; ++k) {
#pragma unroll
      for (int c = 0; c < 9; ++c) T[i][c] -= A[k][i] * T[k][c];
      tg[i] -= A[k][i] * tg[k];
    }
#pragma unroll
    for (int c = 0; c < 9; ++c) T[i][c] *= inv[i];
    tg[i] *= inv[i];
  }
  if (has) {
    // the elimination record: Y 54 | z 6 | g_v 6 | D_v 6
    double2* dst = reinterpret_cast<double2*>(d.yz + (size_t)(base + lane) * kYz);
#pragma unroll
    for (int i = 0; i < 27; ++i) dst[i] = make_double2(T[(2 * i) / 9][(2 * i) % 9], T[(2 * i + 1) / 9][(2 * i + 1) % 9]);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      dst[27 + i] = make_double2(tg[2 * i], tg[2 * i + 1]);
      dst[30 + i] = make_double2(gv[2 * i], gv[2 * i + 1]);
      dst[33 + i] = make_double2(Dv[2 * i], Dv[2 * i + 1]);
    }
  }
}

__global__ void __launch_bounds__(kSchurWarps * 32) k_intr_schur(IntrDev d, LmParams P) {
  __shared__ double s_part[kSchurWarps][kSchurPart];
  __shared__ int s_meta[2];
  __shared__ double s_lambda;
  // d.schur_split CTAs share one work item (the sweep wants few, long items; this kernel many)
  WorkItem it = d.items[blockIdx.x / d.schur_split];
  {
    const int sub = blockIdx.x % d.schur_split, nv = it.v1 - it.v0;
    const int a = it.v0 + (int)((int64_t)nv * sub / d.schur_split);
    const int b = it.v0 + (int)((int64_t)nv * (sub + 1) / d.schur_split);
    it.v0 = a;
    it.v1 = b;
  }
  IntrState& ps = d.ps[it.problem];
  if (threadIdx.x == 0) {
    s_meta[0] = ps.lm.par;
    s_meta[1] = (ps.lm.status == 0 || ps.lm.finalize) ? 1 : 0;
    s_lambda = ps.lm.lambda;
  }
  __syncthreads();
  if (!s_meta[1]) return;
  const int par = s_meta[0];
  const double lambda = s_lambda;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double racc0 = 0.0, racc1 = 0.0;  // lane L: reduced-system terms L and 32+L
  for (int64_t base = it.v0 + warp * 32; base < it.v1; base += kSchurWarps * 32) {
    schur_warp_pass(d, P, par, lambda, base, (int)min((int64_t)32, it.v1 - base), lane, racc0, racc1);
  }
  s_part[warp][lane] = racc0;
  s_part[warp][32 + lane] = racc1;
  __syncthreads();
  if (threadIdx.x < kSchurPart) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kSchurWarps; ++w) s += s_part[w][threadIdx.x];
    d.schur_part[(size_t)blockIdx.x * kSchurPart + threadIdx.x] = s;
  }
}

// --------------------------------------------------------------------------
// fixed-order reduction of per-item partials of one problem (blockIdx.x)
// --------------------------------------------------------------------------
// blockDim.x = nseg * STRIDE threads, nseg <= kRedSeg: thread (c, seg) sums items seg, seg + nseg, ..., the nseg
// segment sums are added in order.  The partials come from other CTAs (possibly of this launch): L2 loads.
template <int STRIDE>
__device__ __forceinline__ void block_reduce_items(const double* part, int item0, int item1,
                                                   double* s_buf /* [kRedSeg][STRIDE] */) {
  const int c = threadIdx.x % STRIDE, seg = threadIdx.x / STRIDE, nseg = blockDim.x / STRIDE;
  double s = 0.0;
  if (seg < nseg) {
    // eight independent loads in flight per thread (the chain of dependent L2 round trips used to dominate)
    for (int i = item0 + seg; i < item1; i += 8 * nseg) {
      double v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int j = i + k * nseg;
        v[k] = j < item1 ? __ldcg(part + (size_t)j * STRIDE + c) : 0.0;
      }
      s += ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
    }
  }
  if (seg < kRedSeg) s_buf[seg * STRIDE + c] = s;
  __syncthreads();
  if (seg == 0) {
    double tot = 0.0;
    for (int g = 0; g < nseg; ++g) tot += s_buf[g * STRIDE + c];
    s_buf[c] = tot;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kRedSeg * kSweepPart) k_intr_reduce_sweep(IntrDev d) {
  __shared__ double s_buf[kRedSeg * kSweepPart];
  IntrState& ps = d.ps[blockIdx.x];
  if (!(ps.lm.status == 0 && ps.lm.solve_ok)) return;
  block_reduce_items<kSweepPart>(d.sweep_part, ps.item0, ps.item1, s_buf);
  if (threadIdx.x < kSweepPart) d.red_sweep[(size_t)blockIdx.x * kSweepPart + threadIdx.x] = s_buf[threadIdx.x];
}

__global__ void __launch_bounds__(kRedSeg * kSchurPart) k_intr_reduce_schur(IntrDev d) {
  __shared__ double s_buf[kRedSeg * kSchurPart];
  IntrState& ps = d.ps[blockIdx.x];
  if (!(ps.lm.status == 0 || ps.lm.finalize)) return;
  block_reduce_items<kSchurPart>(d.schur_part, ps.item0 * d.schur_split, ps.item1 * d.schur_split, s_buf);
  if (threadIdx.x < kSchurPart) d.red_schur[(size_t)blockIdx.x * kSchurPart + threadIdx.x] = s_buf[threadIdx.x];
}

// index into the reduced sweep partial of U(a, b), a <= b (layout at the top of the file), or -1: structural zero
__device__ __forceinline__ int u_partial_index(int a, int b) {
  if (b < 4) {
    if (a == 0 && b == 0) return 1;
    if (a == 0 && b == 2) return 2;
    if (a == 2 && b == 2) return 3;
    if (a == 1 && b == 1) return 4;
    if (a == 1 && b == 3) return 5;
    if (a == 3 && b == 3) return 6;
    return -1;
  }
  const int c = b - 4;
  if (a == 0) return 25 + c;
  if (a == 2) return 30 + c;
  if (a == 1) return 35 + c;
  if (a == 3) return 40 + c;
  const int e = a - 4;
  return 45 + e * 5 - e * (e - 1) / 2 + (c - e);
}

// Accept / reject after a sweep, one CTA per problem (k_intr_decide; all threads call).  REDUCE: 1 sum the partials here (single GPU), 2 also all-reduce
// them over peer memory, 0 read the (NCCL all-reduced) red_sweep buffer.  The head of the LM state is staged in
// shared memory: thread 0 runs the state machine there (it used to walk global memory), all threads write it
// back and, on acceptance, unpack U and g_u.
template <int REDUCE>
__device__ __forceinline__ void decide_tail(const IntrDev& d, const LmParams& P, const PeerComm& pc, int problem,
                                            double* s_buf /* shared, kRedSeg * 64 doubles */) {
  __shared__ double s_lm[kLmHeadWords];
  __shared__ int s_accept;
  IntrState& ps = d.ps[problem];
  const int tid = threadIdx.x, nt = blockDim.x;
  const bool evaluated = ps.lm.solve_ok != 0;
  if (evaluated) {
    if (REDUCE) {
      block_reduce_items<kSweepPart>(d.sweep_part, ps.item0, ps.item1, s_buf);
      // partitioned problem: every rank holds the same LM state, so all of them are here
      if (REDUCE == 2) peer_allreduce64(s_buf, pc);
    } else {
      if (tid < kSweepPart) s_buf[tid] = d.red_sweep[(size_t)problem * kSweepPart + tid];
    }
  }
  if (tid < kLmHeadWords) s_lm[tid] = reinterpret_cast<const double*>(&ps.lm)[tid];
  __syncthreads();
  const double* r = s_buf;
  if (tid == 0) {
    LmState& s = *reinterpret_cast<LmState*>(s_lm);  // only the head is touched; the trace row goes to ps.lm.hist
    bool accept = false;
    if (!evaluated) {
      // the linear solve failed (non-PD): shrink the trust region and retry
      lm_reject(s, P, d.n_done);
      s.solve_ok = 1;
      if (s.status == 0 && s.iterations >= P.max_iterations) lm_finish(s, 4, d.n_done, true);
    } else if (r[kSweepPart - 1] < 0.0) {  // the peer-memory allreduce timed out
      lm_finish(s, CCG_ERR_NCCL, d.n_done, false);
    } else {
      accept = lm_decide(s, P, 0.5 * r[0], r[7] + s.gTd_u, r[8] + s.dDd_u, d.n_done, ps.lm.hist);
      if (accept) {
        s.n_obs = r[3];
        s.sumsq = r[0];
      }
    }
    s_accept = accept ? 1 : 0;
  }
  __syncthreads();
  if (tid < kLmHeadWords) reinterpret_cast<double*>(&ps.lm)[tid] = s_lm[tid];
  if (s_accept) {
    for (int e = tid; e < 90; e += nt) {
      if (e < 81) {
        const int a = e / 9, b = e - 9 * a;
        const int idx = u_partial_index(a < b ? a : b, a < b ? b : a);
        ps.U[e] = idx >= 0 ? r[idx] : 0.0;
      } else {
        ps.gu[e - 81] = r[16 + e - 81];
      }
    }
  }
}

template <int REDUCE>
__global__ void __launch_bounds__(kRedSeg * kSweepPart) k_intr_decide(IntrDev d, LmParams P, PeerComm pc) {
  __shared__ double s_buf[kRedSeg * kSweepPart];
  if (d.ps[blockIdx.x].lm.status != 0) return;
  decide_tail<REDUCE>(d, P, pc, blockIdx.x, s_buf);
}

// Reduced 9x9 solve, one CTA per problem (k_intr_solve; REDUCE as in decide_tail).  Also evaluates the gradient norm at the current point.  State
// head, U, g_u in shared memory; the system is assembled in parallel and solved by one thread in registers
// (chol_solve_reg) instead of one thread walking global and local memory.
template <int REDUCE>
__device__ __forceinline__ void solve_tail(const IntrDev& d, const LmParams& P, const PeerComm& pc, int problem,
                                           double* s_buf /* shared, kRedSeg * 64 doubles */) {
  __shared__ double s_lm[kLmHeadWords];
  __shared__ double s_U[81], s_S[81], s_gu[9], s_rhs[9], s_Du[9], s_x[9];
  __shared__ int s_go;
  IntrState& ps = d.ps[problem];
  const int tid = threadIdx.x, nt = blockDim.x;
  if (REDUCE) {
    block_reduce_items<kSchurPart>(d.schur_part, ps.item0 * d.schur_split, ps.item1 * d.schur_split, s_buf);
    if (REDUCE == 2) peer_allreduce64(s_buf, pc);
  } else {
    if (tid < kSchurPart) s_buf[tid] = d.red_schur[(size_t)problem * kSchurPart + tid];
  }
  if (tid < kLmHeadWords) s_lm[tid] = reinterpret_cast<const double*>(&ps.lm)[tid];
  for (int e = tid; e < 90; e += nt) {
    if (e < 81) s_U[e] = ps.U[e];
    else s_gu[e - 81] = ps.gu[e - 81];
  }
  __syncthreads();
  const double* r = s_buf;
  const unsigned mask = ps.mask;
  LmState& s = *reinterpret_cast<LmState*>(s_lm);  // only the head is touched
  if (tid == 0) {
    int go = 0;
    if (r[kSchurPart - 1] < 0.0) {  // the peer-memory allreduce timed out
      s.finalize = 0;
      if (s.status == 0) lm_finish(s, CCG_ERR_NCCL, d.n_done, false);
    } else {
      double gDg = r[54];
      for (int k = 0; k < 9; ++k)
        if (!((mask >> k) & 1u) && s_U[k * 9 + k] > 0.0) gDg += s_gu[k] * s_gu[k] / s_U[k * 9 + k];
      const double rnorm = sqrt(2.0 * s.cost);
      s.grad_norm = rnorm > 0.0 ? sqrt(gDg) / rnorm : 0.0;
      if (s.finalize) {
        s.finalize = 0;
      } else if (s.cost == 0.0 || sqrt(gDg) <= P.gradient_tolerance * rnorm) {
        lm_finish(s, 1, d.n_done, false);
      } else {
        s.iterations++;
        if (r[55] > 0.0) s.solve_ok = 0;
        else go = 1;
      }
    }
    s_go = go;
  }
  __syncthreads();
  if (s_go) {
    const double lambda = s.lambda;
    for (int e = tid; e < 90; e += nt) {
      if (e < 81) {
        const int a = e / 9, b = e - 9 * a;
        double v = s_U[e] - (a <= b ? r[up9(a, b)] : r[up9(b, a)]);  // packed upper triangle of W'V'^-1W
        if (a == b) v += lambda * clampd(s_U[e], P.min_diagonal, P.max_diagonal);
        if (((mask >> a) & 1u) || ((mask >> b) & 1u)) v = a == b ? 1.0 : 0.0;
        s_S[e] = v;
      } else {
        const int a = e - 81;
        s_Du[a] = clampd(s_U[a * 9 + a], P.min_diagonal, P.max_diagonal);
        s_rhs[a] = ((mask >> a) & 1u) ? 0.0 : -s_gu[a] + r[45 + a];
      }
    }
    __syncthreads();
    if (tid == 0) {
      // 9 x 9: one thread, everything in registers (a warp-cooperative solve through shared memory was slower:
      // it serialises 27 double-precision divisions and square roots behind warp barriers)
      double x[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) x[k] = s_rhs[k];
      const bool ok = chol_solve_reg<9>(s_S, x);
      if (ok) {
        double gTd = 0.0, dDd = 0.0;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          const double dk = ((mask >> k) & 1u) ? 0.0 : x[k];
          s_x[k] = dk;
          gTd += s_gu[k] * dk;
          dDd += dk * dk * s_Du[k];
        }
        s.gTd_u = gTd;
        s.dDd_u = dDd;
        s.solve_ok = 1;
      } else {
        s.solve_ok = 0;
      }
      s_go = ok ? 2 : 0;
    }
    __syncthreads();
    if (s_go == 2 && tid < 9) {
      const int par = s.par;
      ps.du[tid] = s_x[tid];
      ps.intr[par ^ 1][tid] = ps.intr[par][tid] + s_x[tid];
    }
    __syncthreads();
  }
  if (tid < kLmHeadWords) reinterpret_cast<double*>(&ps.lm)[tid] = s_lm[tid];
}

template <int REDUCE>
__global__ void __launch_bounds__(kRedSeg * kSchurPart) k_intr_solve(IntrDev d, LmParams P, PeerComm pc) {
  const IntrState& ps = d.ps[blockIdx.x];
  __shared__ double s_buf[kRedSeg * kSchurPart];
  if (!(ps.lm.status == 0 || ps.lm.finalize)) return;
  solve_tail<REDUCE>(d, P, pc, blockIdx.x, s_buf);
}

}  // namespace ccg
#include "intr_lm.cuh"
namespace ccg {

// max |measurement| per problem (scale of the cost's rounding noise, lm_decide): one CTA per work
// item, non-negative floats compared through their bit patterns
__global__ void __launch_bounds__(256) k_intr_meas_scale(IntrDev d, unsigned* __restrict__ bits) {
  const WorkItem it = d.items[blockIdx.x];
  const int64_t o0 = d.view_off[it.v0], o1 = d.view_off[it.v1];
  float m = 0.0f;
  for (int64_t o = o0 + threadIdx.x; o < o1; o += blockDim.x)
    m = fmaxf(m, fmaxf(fabsf(__ldg(d.u + o)), fabsf(__ldg(d.v + o))));
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
  if ((threadIdx.x & 31) == 0) atomicMax(bits + it.problem, __float_as_uint(m));
}

__global__ void k_intr_meas_finish(const unsigned* __restrict__ bits, double* __restrict__ out, int n) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < n) out[b] = (double)__uint_as_float(bits[b]);
}

__global__ void k_intr_init_state(IntrDev d, LmParams P, const double* intr0, const int2* item_ranges,
                                  const double* meas_scale, unsigned mask, int n_problems) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_problems) return;
  IntrState& ps = d.ps[b];
  lm_init(ps.lm, P);
  ps.lm.meas_scale = meas_scale[b];
  for (int k = 0; k < 9; ++k) {
    ps.intr[0][k] = intr0[(size_t)b * 9 + k];
    ps.intr[1][k] = intr0[(size_t)b * 9 + k];
    ps.du[k] = 0.0;
    ps.gu[k] = 0.0;
  }
  for (int k = 0; k < 81; ++k) ps.U[k] = 0.0;
  ps.item0 = item_ranges[b].x;
  ps.item1 = item_ranges[b].y;
  ps.mask = mask;
  ps.pad_ = 0;
  if (b == 0) *d.n_done = 0;
}

// results in the caller's layout: q[4V], t[3V] of the accepted pose buffer of each view's problem,
// intr[9B] and one ccg_report per problem (so the download is four plain copies, no host loops)
__global__ void __launch_bounds__(128) k_intr_gather(IntrDev d, const int64_t* __restrict__ problem_off, int n_problems,
                                                    double* __restrict__ q, double* __restrict__ t,
                                                    double* __restrict__ intr, ccg_report* __restrict__ rep) {
  // grid: (view blocks of the largest problem, problems)
  const int b = blockIdx.y;
  const IntrState& ps = d.ps[b];
  const int par = ps.lm.par;
  const double* pose = d.pose[par];
  const int64_t v = problem_off[b] + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v < problem_off[b + 1]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) q[v * 4 + k] = pose[v * kPose + k];
#pragma unroll
    for (int k = 0; k < 3; ++k) t[v * 3 + k] = pose[v * kPose + 4 + k];
  }
  if (blockIdx.x != 0) return;
  if (threadIdx.x < 9) intr[(size_t)b * 9 + threadIdx.x] = ps.intr[par][threadIdx.x];
  if (threadIdx.x == 0) {
    ccg_report r;
    r.iterations = ps.lm.iterations;
    r.accepted = ps.lm.accepted;
    r.rejected = ps.lm.rejected;
    r.sweeps = ps.lm.sweeps;
    r.termination = ps.lm.status;
    r.pad_ = 0;
    r.initial_cost = ps.lm.cost0;
    r.final_cost = ps.lm.cost;
    r.gradient_norm = ps.lm.grad_norm;
    r.rms = ps.lm.n_obs > 0.0 ? sqrt(2.0 * ps.lm.cost / ps.lm.n_obs) : 0.0;
    rep[b] = r;
  }
}

}  // namespace ccg

// ==========================================================================
// host side
// ==========================================================================
using namespace ccg;

struct ccg_intr_problem {
  ccg_ctx* ctx = nullptr;
  int64_t n_views = 0, n_obs = 0, n_problems = 0;
  int n_items = 0;
  int shared = 0;
  unsigned mask = 0;
  IntrDev dev;
  // owned device allocations
  std::vector<void*> allocs;
  int64_t bytes = 0;
  double *d_intr0 = nullptr, *d_pose0 = nullptr, *d_meas_scale = nullptr;
  double *d_out_q = nullptr, *d_out_t = nullptr, *d_out_intr = nullptr;  // results in the caller's layout
  ccg_report* d_out_rep = nullptr;
  int64_t* d_problem_off = nullptr;
  bool one_shot = false;             // created by ccg_intrinsics_solve*: runs once, no graph instantiation
  LmCtl* d_ctl = nullptr;            // persistent LM kernel: grid-barrier counter
  double* d_lm_schur = nullptr;      // persistent LM kernel: one elimination partial per work item
  int lm_mode = -1;                  // -1 the multi-launch loop, 0 the whole loop as one launch (k_intr_lm, intr_lm.cuh)
  int lm_grid = 0;
  int2* d_item_ranges = nullptr;
  float* d_obs = nullptr;
  bool single_item = false;          // every problem is one work item (batch of small problems)
  std::vector<int64_t> problem_off;  // host copy
  std::vector<int64_t> problem_obs;  // observations per problem
  LmParams params;
  bool params_set = false;
  // one LM iteration (elimination, reduced solve, candidate poses, sweep, accept/reject) as a CUDA
  // graph: five launches become one when no per-kernel events are requested
  cudaGraphExec_t graph_exec = nullptr;
  LmParams graph_params;
  int graph_kernels = 0, graph_allreduces = 0;
};

namespace {

template <typename T>
int dev_alloc(ccg_intr_problem* p, T** out, size_t n) {
  void* ptr = nullptr;
  const size_t bytes = std::max<size_t>(n, 1) * sizeof(T);
  cudaError_t e = pool_alloc(p->ctx, &ptr, bytes);
  if (e != cudaSuccess) return set_error(p->ctx, CCG_ERR_CUDA, "device allocation", e);
  p->allocs.push_back(ptr);
  p->bytes += (int64_t)bytes;
  *out = (T*)ptr;
  return CCG_OK;
}

#define CCG_TRY(x)            \
  do {                        \
    const int rc__ = (x);     \
    if (rc__ != CCG_OK) return rc__; \
  } while (0)

int build_problem(ccg_ctx* ctx, const float* u, const float* v, const float* X, const float* Y,
                  const float* Z, const int64_t* view_offsets, int64_t n_views,
                  const int64_t* problem_offsets, int64_t n_problems, const double* intr,
                  const double* q, const double* t, unsigned mask, int shared,
                  ccg_intr_problem* p) {
  p->ctx = ctx;
  p->n_views = n_views;
  p->n_problems = n_problems;
  p->shared = shared;
  p->mask = mask;
  const int64_t o0 = view_offsets[0];
  const int64_t n_obs = view_offsets[n_views] - o0;
  p->n_obs = n_obs;
  cudaStream_t st = ctx->stream;

  // ---- work items: consecutive views of one problem per CTA
  // one CTA per SM for a big problem (255-register kernels: one resident CTA per SM, so more
  // items only add wave-quantisation tail); small problems get one item each
  const int64_t target_items = (int64_t)ctx->sm_count;
  // (at least 32 views per item: four per warp pair of the default sweep; fewer only add CTAs that mostly wait)
  int64_t vpi = std::max<int64_t>(32, (n_views + target_items - 1) / std::max<int64_t>(target_items, 1));
  std::vector<WorkItem> items;
  std::vector<int2> ranges((size_t)n_problems);
  p->problem_off.assign(problem_offsets, problem_offsets + n_problems + 1);
  p->problem_obs.resize((size_t)n_problems);
  for (int64_t b = 0; b < n_problems; ++b) {
    const int64_t v0 = problem_offsets[b], v1 = problem_offsets[b + 1];
    if (v1 < v0 || v0 < 0 || v1 > n_views) return set_error(ctx, CCG_ERR_BAD_ARG, "problem_offsets");
    p->problem_obs[(size_t)b] = view_offsets[v1] - view_offsets[v0];
    ranges[(size_t)b].x = (int)items.size();
    const int64_t nv = v1 - v0;
    const int64_t n_it = std::max<int64_t>(1, (nv + vpi - 1) / vpi);
    for (int64_t i = 0; i < n_it; ++i) {
      WorkItem w;
      w.problem = (int)b;
      w.v0 = (int)(v0 + nv * i / n_it);
      w.v1 = (int)(v0 + nv * (i + 1) / n_it);
      w.pad = 0;
      items.push_back(w);
    }
    ranges[(size_t)b].y = (int)items.size();
  }
  p->n_items = (int)items.size();
  p->single_item = n_problems > 1 && (int64_t)items.size() == n_problems;

  // ---- device memory
  IntrDev& d = p->dev;
  std::memset(&d, 0, sizeof(d));
  float* obs = nullptr;
  CCG_TRY(dev_alloc(p, &obs, (size_t)n_obs * 5));
  p->d_obs = obs;
  int64_t* d_voff = nullptr;
  CCG_TRY(dev_alloc(p, &d_voff, (size_t)n_views + 1));
  WorkItem* d_items = nullptr;
  CCG_TRY(dev_alloc(p, &d_items, items.size()));
  CCG_TRY(dev_alloc(p, &p->d_item_ranges, (size_t)n_problems));
  CCG_TRY(dev_alloc(p, &d.pose[0], (size_t)n_views * kPose));
  CCG_TRY(dev_alloc(p, &d.pose[1], (size_t)n_views * kPose));
  CCG_TRY(dev_alloc(p, &p->d_pose0, (size_t)n_views * kPose));
  CCG_TRY(dev_alloc(p, &d.blk[0], (size_t)n_views * kBlk));
  CCG_TRY(dev_alloc(p, &d.blk[1], (size_t)n_views * kBlk));
  CCG_TRY(dev_alloc(p, &d.yz, (size_t)n_views * kYz));
  CCG_TRY(dev_alloc(p, &d.sweep_part, items.size() * kSweepPart));
  // elimination kernel: 64 views per CTA pass, one pass per CTA (a partition of a
  // large problem over 8 GPUs leaves ~85 views per item: two half-empty passes would double its latency)
  int64_t max_item_views = 1;
  for (const WorkItem& w : items) max_item_views = std::max<int64_t>(max_item_views, w.v1 - w.v0);
  // CTAs of the elimination kernel per work item: one 64-view pass each while that keeps the grid within ~4 CTAs
  // per SM (every CTA adds a partial to the reduction that follows), several passes per CTA beyond
  const int per_pass = kSchurWarps * 32;
  const int64_t cap = std::max<int64_t>(1, (4 * (int64_t)ctx->sm_count) / std::max<int64_t>(1, (int64_t)items.size()));
  const int schur_split = (int)std::max<int64_t>(1, std::min<int64_t>((max_item_views + per_pass - 1) / per_pass, cap));
  CCG_TRY(dev_alloc(p, &d.schur_part, items.size() * (size_t)schur_split * kSchurPart));
  CCG_TRY(dev_alloc(p, &d.rt, (size_t)n_views * 12));
  CCG_TRY(dev_alloc(p, &d.rtf, (size_t)n_views * 12));
  CCG_TRY(dev_alloc(p, &d.red_sweep, (size_t)n_problems * kSweepPart));
  CCG_TRY(dev_alloc(p, &d.red_schur, (size_t)n_problems * kSchurPart));
  CCG_TRY(dev_alloc(p, &d.ps, (size_t)n_problems));
  CCG_TRY(dev_alloc(p, &d.n_done, 1));
  CCG_TRY(dev_alloc(p, &p->d_intr0, (size_t)n_problems * 9));
  CCG_TRY(dev_alloc(p, &p->d_meas_scale, (size_t)n_problems));

  // observations are re-based so that view_off[0] == 0 on the device
  d.u = obs;
  d.v = obs + n_obs;
  d.X = obs + 2 * n_obs;
  d.Y = obs + 3 * n_obs;
  d.Z = obs + 4 * n_obs;
  d.view_off = d_voff;
  d.items = d_items;
  d.obs_base = 0;
  d.schur_split = schur_split;
  CCG_CUDA(ctx, cudaMemcpyAsync(obs, u + o0, sizeof(float) * n_obs, cudaMemcpyHostToDevice, st));
  CCG_CUDA(ctx, cudaMemcpyAsync(obs + n_obs, v + o0, sizeof(float) * n_obs, cudaMemcpyHostToDevice, st));
  CCG_CUDA(ctx, cudaMemcpyAsync(obs + 2 * n_obs, X + o0, sizeof(float) * n_obs, cudaMemcpyHostToDevice, st));
  CCG_CUDA(ctx, cudaMemcpyAsync(obs + 3 * n_obs, Y + o0, sizeof(float) * n_obs, cudaMemcpyHostToDevice, st));
  CCG_CUDA(ctx, cudaMemcpyAsync(obs + 4 * n_obs, Z + o0, sizeof(float) * n_obs, cudaMemcpyHostToDevice, st));
  // ---- per-view arrays: packed into the context's pinned staging arena (while the observation arrays above are on
  // the wire) and uploaded behind them; pageable vectors only if the arena cannot be had
  const size_t nv1 = (size_t)n_views + 1;
  const size_t b_voff = sizeof(int64_t) * nv1, b_vmeta = sizeof(int2) * nv1, b_pose = sizeof(double) * (size_t)n_views * kPose,
               b_items = sizeof(WorkItem) * items.size(), b_ranges = sizeof(int2) * ranges.size(),
               b_intr = sizeof(double) * 9 * (size_t)n_problems, b_poff = sizeof(int64_t) * ((size_t)n_problems + 1);
  auto up8 = [](size_t x) { return (x + 15) & ~(size_t)15; };
  const size_t stage_bytes = up8(b_voff) + up8(b_vmeta) + up8(b_pose) + up8(b_items) + up8(b_ranges) + up8(b_intr) + up8(b_poff);
  std::vector<unsigned char> pageable;
  unsigned char* stage = static_cast<unsigned char*>(host_stage(ctx, stage_bytes));
  if (!stage) {
    pageable.resize(stage_bytes);
    stage = pageable.data();
  }
  size_t cur = 0;
  auto carve = [&](size_t bytes) {
    unsigned char* q = stage + cur;
    cur += up8(bytes);
    return q;
  };
  int64_t* voff = reinterpret_cast<int64_t*>(carve(b_voff));
  int2* vmeta = reinterpret_cast<int2*>(carve(b_vmeta));
  double* pose = reinterpret_cast<double*>(carve(b_pose));
  WorkItem* h_items = reinterpret_cast<WorkItem*>(carve(b_items));
  int2* h_ranges = reinterpret_cast<int2*>(carve(b_ranges));
  double* h_intr = reinterpret_cast<double*>(carve(b_intr));
  int64_t* h_poff = reinterpret_cast<int64_t*>(carve(b_poff));
  for (int64_t i = 0; i <= n_views; ++i) voff[i] = view_offsets[i] - o0;
  // block layout of the default sweep: every view padded to whole 32-observation blocks
  int64_t n_blocks = 0;
  for (int64_t i = 0; i < n_views; ++i) {
    const int64_t n = view_offsets[i + 1] - view_offsets[i];
    if (n > 0x7fffffff || n_blocks > 0x7fffffff) return set_error(ctx, CCG_ERR_UNSUPPORTED, "view or problem too large");
    vmeta[i].x = (int)n_blocks;
    vmeta[i].y = (int)n;
    n_blocks += (n + 31) / 32;
  }
  if (n_blocks > 0x7fffffff) return set_error(ctx, CCG_ERR_UNSUPPORTED, "problem too large");
  vmeta[n_views].x = (int)n_blocks;
  vmeta[n_views].y = 0;
  int2* d_vmeta = nullptr;
  float* d_blk = nullptr;
  CCG_TRY(dev_alloc(p, &d_vmeta, nv1));
  CCG_TRY(dev_alloc(p, &d_blk, (size_t)std::max<int64_t>(n_blocks, 1) * kObsBlock));
  d.vmeta = d_vmeta;
  d.obs_blk = d_blk;
  for (int64_t i = 0; i < n_views; ++i) {
    double* pp = pose + (size_t)i * kPose;
    for (int k = 0; k < 4; ++k) pp[k] = q[i * 4 + k];
    for (int k = 0; k < 3; ++k) pp[4 + k] = t[i * 3 + k];
    for (int k = 7; k < kPose; ++k) pp[k] = 0.0;
  }
  std::memcpy(h_items, items.data(), b_items);
  std::memcpy(h_ranges, ranges.data(), b_ranges);
  std::memcpy(h_intr, intr, b_intr);
  std::memcpy(h_poff, p->problem_off.data(), b_poff);
  CCG_TRY(dev_alloc(p, &p->d_problem_off, (size_t)n_problems + 1));
  CCG_CUDA(ctx, cudaMemcpyAsync(d_vmeta, vmeta, b_vmeta, cudaMemcpyHostToDevice, st));
  CCG_CUDA(ctx, cudaMemcpyAsync(d_voff, voff, b_voff, cudaMemcpyHostToDevice, st));
  CCG_CUDA(ctx, cudaMemcpyAsync(d_items, h_items, b_items, cudaMemcpyHostToDevice, st));
  CCG_CUDA(ctx, cudaMemcpyAsync(p->d_item_ranges, h_ranges, b_ranges, cudaMemcpyHostToDevice, st));
  if (n_views > 0) CCG_CUDA(ctx, cudaMemcpyAsync(p->d_pose0, pose, b_pose, cudaMemcpyHostToDevice, st));
  CCG_CUDA(ctx, cudaMemcpyAsync(p->d_intr0, h_intr, b_intr, cudaMemcpyHostToDevice, st));
  CCG_CUDA(ctx, cudaMemcpyAsync(p->d_problem_off, h_poff, b_poff, cudaMemcpyHostToDevice, st));
  if (n_views > 0) {
    k_intr_pack<<<(unsigned)((n_views * 32 + 255) / 256), 256, 0, st>>>(d, n_views, d_blk);
    ctx->launches++;
  }
  // max |measurement| per problem, computed on the device from the uploaded observations
  {
    unsigned* d_bits = nullptr;
    CCG_TRY(dev_alloc(p, &d_bits, (size_t)n_problems));
    CCG_CUDA(ctx, cudaMemsetAsync(d_bits, 0, sizeof(unsigned) * n_problems, st));
    k_intr_meas_scale<<<p->n_items, 256, 0, st>>>(d, d_bits);
    k_intr_meas_finish<<<(int)((n_problems + 127) / 128), 128, 0, st>>>(d_bits, p->d_meas_scale, (int)n_problems);
    ctx->launches += 2;
    // a partitioned calibration: the scale feeds the accept/reject thresholds of lm_decide, and every rank
    // must take the same branch - the maximum is over ALL ranks' observations
    if (shared && ctx->world > 1) CCG_TRY(nccl_allreduce_max(ctx, p->d_meas_scale, (size_t)n_problems));
  }
  CCG_TRY(dev_alloc(p, &p->d_ctl, (size_t)1));
  CCG_TRY(dev_alloc(p, &p->d_lm_schur, items.size() * (size_t)kSchurPart));
  // the whole LM loop as one launch (intr_lm.cuh) when every CTA of the problem can be resident at once (one
  // problem, one CTA per SM) or when every problem is a single work item (batches of small problems)
  // Measured (B200): one launch beats the four-launches-per-iteration loop for small problems (C1 0.21 vs 0.34 ms, C2
  // 0.30 vs 0.40 ms per solve) and loses from a few million observations on (12,500 x 400: 1.20 vs 1.13 ms; C5: 7.07
  // vs 6.69 ms), where its elimination phase - 8 warps per SM, register-starved next to the sweep's split - costs
  // more than the launches it saves.  CCG_PERSIST_MAX_OBS overrides the threshold (experiments).
  p->lm_mode = -1;
  if (!getenv("CCG_NO_PERSIST")) {
    int coop = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device);
    const char* e = getenv("CCG_PERSIST_MAX_OBS");
    const int64_t max_obs = e ? atoll(e) : 1500000;
    if (n_problems == 1 && coop && n_obs <= max_obs && p->n_items <= ctx->sm_count &&
        (!(shared && ctx->world > 1) || ctx->peer_ok)) {
      p->lm_mode = 0;
      p->lm_grid = p->n_items;
    }
  }
  CCG_CUDA(ctx, cudaStreamSynchronize(st));  // host staging vectors go out of scope
  return CCG_OK;
}

int reset_problem(ccg_intr_problem* p, const LmParams& P) {
  ccg_ctx* ctx = p->ctx;
  IntrDev& d = p->dev;
  cudaStream_t st = ctx->stream;
  CCG_CUDA(ctx, cudaMemcpyAsync(d.pose[0], p->d_pose0, sizeof(double) * p->n_views * kPose, cudaMemcpyDeviceToDevice, st));
  CCG_CUDA(ctx, cudaMemsetAsync(d.yz, 0, sizeof(double) * std::max<int64_t>(p->n_views, 1) * kYz, st));
  const int nb = (int)p->n_problems;
  k_intr_init_state<<<(nb + 127) / 128, 128, 0, st>>>(d, P, p->d_intr0, p->d_item_ranges, p->d_meas_scale, p->mask, nb);
  ctx->launches++;
  CCG_CUDA(ctx, cudaGetLastError());
  p->params = P;
  p->params_set = true;
  return CCG_OK;
}

enum { TAG_SWEEP = 0, TAG_SCHUR = 1, TAG_SMALL = 2 };

template <bool RES, int PAIRS>
void launch_sweep_ws_tp(ccg_intr_problem* p, double* res_out) {
  ccg_ctx* ctx = p->ctx;
  constexpr int slot = (RES ? 1 : 0) + (PAIRS == 4 ? 2 : 0);
  if (!ctx->sweep_attr[slot]) {  // per context: the attribute is per device
    // (no carveout preference: with cudaSharedmemCarveoutMaxShared the few spilled registers of the roles lose their
    // L1 and the launch is 8 % slower - measured)
    cudaFuncSetAttribute(k_intr_sweep_ws<RES, PAIRS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sw_smem_bytes(PAIRS));
    ctx->sweep_attr[slot] = true;
  }
  k_intr_sweep_ws<RES, PAIRS><<<p->n_items, 64 * PAIRS, sw_smem_bytes(PAIRS), ctx->stream>>>(p->dev, res_out);
}
template <bool RES>
void launch_sweep_ws_t(ccg_intr_problem* p, double* res_out) {
  // a batch of one-CTA problems: two CTAs of 4 warp pairs per SM (see kSwPairs in intr_sweep.cuh)
  if (p->single_item) launch_sweep_ws_tp<RES, 4>(p, res_out);
  else launch_sweep_ws_tp<RES, 8>(p, res_out);
}

// one residual+Jacobian sweep followed by the accept/reject step
int launch_sweep_decide(ccg_intr_problem* p, const LmParams& P, Profiler& prof, double* res_out, int* n_allreduce) {
  ccg_ctx* ctx = p->ctx;
  IntrDev& d = p->dev;
  cudaStream_t st = ctx->stream;
  const int nb = (int)p->n_problems;
  const PeerComm pc = peer_comm(ctx);
  const bool partitioned = p->shared && ctx->world > 1;
  prof.begin(TAG_SWEEP);
  if (P.hessian_fp64 == 0) {
    // default: candidate poses + warp-specialised tensor-core sweep in one launch
    if (res_out) launch_sweep_ws_t<true>(p, res_out);
    else launch_sweep_ws_t<false>(p, res_out);
  } else {
    k_intr_sweep<<<p->n_items, kWarps * 32, 0, st>>>(d, res_out);  // everything in FP64, a warp pair per view
  }
  ctx->launches++;
  prof.end();
  prof.begin(TAG_SMALL);
  if (partitioned && ctx->peer_ok) {
    k_intr_decide<2><<<nb, kRedSeg * kSweepPart, 0, st>>>(d, P, pc);
    (*n_allreduce)++;
    ctx->launches++;
  } else if (partitioned) {
    k_intr_reduce_sweep<<<nb, kRedSeg * kSweepPart, 0, st>>>(d);
    CCG_TRY(nccl_allreduce_sum(ctx, d.red_sweep, (size_t)nb * kSweepPart));
    (*n_allreduce)++;
    k_intr_decide<0><<<nb, kRedSeg * kSweepPart, 0, st>>>(d, P, pc);
    ctx->launches += 2;
  } else if (p->single_item) {
    // batch of small problems (one work item each): a 64-thread CTA per problem
    k_intr_decide<1><<<nb, kSweepPart, 0, st>>>(d, P, pc);
    ctx->launches++;
  } else {
    k_intr_decide<1><<<nb, kRedSeg * kSweepPart, 0, st>>>(d, P, pc);
    ctx->launches++;
  }
  prof.end();
  return CCG_OK;
}

int launch_schur_solve(ccg_intr_problem* p, const LmParams& P, Profiler& prof, int* n_allreduce) {
  ccg_ctx* ctx = p->ctx;
  IntrDev& d = p->dev;
  cudaStream_t st = ctx->stream;
  const int nb = (int)p->n_problems;
  const PeerComm pc = peer_comm(ctx);
  const bool partitioned = p->shared && ctx->world > 1;
  prof.begin(TAG_SCHUR);
  k_intr_schur<<<p->n_items * d.schur_split, kSchurWarps * 32, 0, st>>>(d, P);
  prof.end();
  ctx->launches++;
  prof.begin(TAG_SMALL);
  if (partitioned && ctx->peer_ok) {
    k_intr_solve<2><<<nb, kRedSeg * kSchurPart, 0, st>>>(d, P, pc);
    (*n_allreduce)++;
    ctx->launches++;
  } else if (partitioned) {
    k_intr_reduce_schur<<<nb, kRedSeg * kSchurPart, 0, st>>>(d);
    CCG_TRY(nccl_allreduce_sum(ctx, d.red_schur, (size_t)nb * kSchurPart));
    (*n_allreduce)++;
    k_intr_solve<0><<<nb, kRedSeg * kSchurPart, 0, st>>>(d, P, pc);
    ctx->launches += 2;
  } else if (p->single_item) {
    k_intr_solve<1><<<nb, kSchurPart, 0, st>>>(d, P, pc);
    ctx->launches++;
  } else {
    k_intr_solve<1><<<nb, kRedSeg * kSchurPart, 0, st>>>(d, P, pc);
    ctx->launches++;
  }
  prof.end();
  return CCG_OK;
}

int poll_done(ccg_intr_problem* p, int* done) {
  ccg_ctx* ctx = p->ctx;
  CCG_CUDA(ctx, cudaMemcpyAsync(ctx->h_poll, p->dev.n_done, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CCG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *done = ctx->h_poll[0];
  return CCG_OK;
}

int run_problem(ccg_intr_problem* p, const ccg_opts& opts, ccg_timing* timing) {
  ccg_ctx* ctx = p->ctx;
  CCG_CUDA(ctx, cudaSetDevice(ctx->device));
  const LmParams P = to_params(opts);
  Profiler prof;
  prof.ctx = ctx;
  prof.on = opts.profile != 0;
  int n_allreduce = 0;
  size_t ev_total0 = 0, ev_total1 = 0;
  {
    // total-loop events are always recorded (cheap)
    Profiler tot;
    tot.ctx = ctx;
    tot.on = true;
    tot.used = 0;
    ev_total0 = tot.grab();
    ev_total1 = tot.grab();
    prof.used = tot.used;
  }
  cudaEventRecord(ctx->events[ev_total0], ctx->stream);
  CCG_TRY(reset_problem(p, P));  // restoring the initial parameters is part of the timed solve
  if (p->lm_mode >= 0 && P.hessian_fp64 == 0 && !prof.on) {
    // one launch for the whole loop
    const bool partitioned = p->shared && ctx->world > 1;
    if (partitioned) {
      CCG_TRY(peer_resync(ctx, p->dev.red_sweep));  // lines the ranks up and agrees on the allreduce sequence number
      n_allreduce++;
    }
    CCG_CUDA(ctx, cudaMemsetAsync(p->d_ctl, 0, sizeof(LmCtl), ctx->stream));
    IntrDev dev = p->dev;
    LmParams Pk = P;
    PeerComm pck = peer_comm(ctx);
    int reduce_mode = partitioned ? 2 : 1;
    LmCtl* ctl = p->d_ctl;
    double* lms = p->d_lm_schur;
    void* args[] = {&dev, &Pk, &pck, &reduce_mode, &ctl, &lms};
    if (!ctx->lm_attr) {
      CCG_CUDA(ctx, cudaFuncSetAttribute(k_intr_lm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSwSmem));
      ctx->lm_attr = true;
    }
    CCG_CUDA(ctx, cudaLaunchCooperativeKernel((void*)k_intr_lm, dim3(p->lm_grid), dim3(kSwThreads), args, kSwSmem,
                                              ctx->stream));
    ctx->launches++;
    cudaEventRecord(ctx->events[ev_total1], ctx->stream);
    int counts[2] = {0, 0};  // iterations, sweeps of problem 0: the allreduce count of a partitioned solve
    if (partitioned)
      CCG_CUDA(ctx, cudaMemcpyAsync(counts, reinterpret_cast<const char*>(&p->dev.ps->lm) + offsetof(LmState, iterations),
                                    sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CCG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    CCG_CUDA(ctx, cudaGetLastError());
    if (timing) {
      std::memset(timing, 0, sizeof(*timing));
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ctx->events[ev_total0], ctx->events[ev_total1]);
      timing->total_ms = ms;
      timing->allreduces = n_allreduce + (partitioned ? 2 * counts[0] + 3 : 0);
      timing->sweep_obs = p->n_obs;
      timing->sweep_views = p->n_views;
    }
    return CCG_OK;
  }
  if (p->shared && ctx->world > 1 && ctx->peer_ok) {
    // line the ranks up once per solve with a collective that waits without a deadline; the
    // peer-memory allreduces inside the kernels then only see the skew of one LM iteration
    // (the same collective re-synchronises the sequence numbers of those allreduces)
    CCG_TRY(peer_resync(ctx, p->dev.red_sweep));
    n_allreduce++;
  }
  CCG_TRY(launch_sweep_decide(p, P, prof, nullptr, &n_allreduce));
  // per-kernel events stay outside graphs; the NCCL allreduces of a partitioned problem are captured
  // with the kernels (CCG_NO_GRAPH=1 disables graphs altogether)
  static const bool no_graph = getenv("CCG_NO_GRAPH") != nullptr;
  const bool use_graph = !prof.on && !no_graph && !p->one_shot;
  if (use_graph && (!p->graph_exec || std::memcmp(&p->graph_params, &P, sizeof(P)) != 0)) {
    if (p->graph_exec) cudaGraphExecDestroy(p->graph_exec);
    p->graph_exec = nullptr;
    cudaGraph_t g = nullptr;
    const int64_t l0 = ctx->launches;
    const int a0 = n_allreduce;
    CCG_CUDA(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed));
    int rc = launch_schur_solve(p, P, prof, &n_allreduce);
    if (rc == CCG_OK) rc = launch_sweep_decide(p, P, prof, nullptr, &n_allreduce);
    const cudaError_t e = cudaStreamEndCapture(ctx->stream, &g);
    p->graph_kernels = (int)(ctx->launches - l0);
    p->graph_allreduces = n_allreduce - a0;
    ctx->launches = l0;
    n_allreduce = a0;
    if (rc != CCG_OK) return rc;
    if (e != cudaSuccess) return set_error(ctx, CCG_ERR_CUDA, "cudaStreamEndCapture", e);
    CCG_CUDA(ctx, cudaGraphInstantiate(&p->graph_exec, g, 0));
    cudaGraphDestroy(g);
    p->graph_params = P;
  }
  const int poll_every = 4;
  int done = 0;
  for (int it = 0; it <= opts.max_iterations * 2 + 64; ++it) {
    if (use_graph && done < (int)p->n_problems) {
      CCG_CUDA(ctx, cudaGraphLaunch(p->graph_exec, ctx->stream));
      ctx->launches += p->graph_kernels;
      n_allreduce += p->graph_allreduces;
    } else {
      CCG_TRY(launch_schur_solve(p, P, prof, &n_allreduce));
      if (done >= (int)p->n_problems) break;  // that was the finalize pass
      CCG_TRY(launch_sweep_decide(p, P, prof, nullptr, &n_allreduce));
    }
    if ((it % poll_every) == poll_every - 1) CCG_TRY(poll_done(p, &done));
  }
  cudaEventRecord(ctx->events[ev_total1], ctx->stream);
  CCG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  CCG_CUDA(ctx, cudaGetLastError());
  if (timing) {
    std::memset(timing, 0, sizeof(*timing));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->events[ev_total0], ctx->events[ev_total1]);
    timing->total_ms = ms;
    double tms[4];
    int tn[4];
    prof.collect(tms, tn);
    timing->sweep_ms = tms[TAG_SWEEP];
    timing->schur_ms = tms[TAG_SCHUR];
    timing->small_ms = tms[TAG_SMALL];
    timing->sweep_launches = tn[TAG_SWEEP];
    timing->schur_launches = tn[TAG_SCHUR];
    timing->small_launches = tn[TAG_SMALL];
    timing->allreduces = n_allreduce;
    timing->sweep_obs = p->n_obs;
    timing->sweep_views = p->n_views;
  }
  return CCG_OK;
}

int download_problem(ccg_intr_problem* p, double* intr, double* q, double* t, ccg_report* reports) {
  ccg_ctx* ctx = p->ctx;
  CCG_CUDA(ctx, cudaSetDevice(ctx->device));
  const int64_t nb = p->n_problems, nv = p->n_views;
  cudaStream_t st = ctx->stream;
  if (!p->d_out_q) {
    CCG_TRY(dev_alloc(p, &p->d_out_q, (size_t)nv * 4));
    CCG_TRY(dev_alloc(p, &p->d_out_t, (size_t)nv * 3));
    CCG_TRY(dev_alloc(p, &p->d_out_intr, (size_t)nb * 9));
    CCG_TRY(dev_alloc(p, &p->d_out_rep, (size_t)nb));
  }
  int64_t max_views = 1;
  for (int64_t b = 0; b < nb; ++b)
    max_views = std::max(max_views, p->problem_off[(size_t)b + 1] - p->problem_off[(size_t)b]);
  if (nb > 65535) return set_error(ctx, CCG_ERR_UNSUPPORTED, "more than 65535 problems in one batch");
  k_intr_gather<<<dim3((unsigned)((max_views + 127) / 128), (unsigned)nb), 128, 0, st>>>(
      p->dev, p->d_problem_off, (int)nb, p->d_out_q, p->d_out_t, p->d_out_intr, p->d_out_rep);
  ctx->launches++;
  std::vector<ccg_report> rep_local;
  if (!reports) {
    rep_local.resize((size_t)nb);
    reports = rep_local.data();
  }
  if (q && nv > 0) CCG_CUDA(ctx, cudaMemcpyAsync(q, p->d_out_q, sizeof(double) * nv * 4, cudaMemcpyDeviceToHost, st));
  if (t && nv > 0) CCG_CUDA(ctx, cudaMemcpyAsync(t, p->d_out_t, sizeof(double) * nv * 3, cudaMemcpyDeviceToHost, st));
  if (intr) CCG_CUDA(ctx, cudaMemcpyAsync(intr, p->d_out_intr, sizeof(double) * nb * 9, cudaMemcpyDeviceToHost, st));
  CCG_CUDA(ctx, cudaMemcpyAsync(reports, p->d_out_rep, sizeof(ccg_report) * nb, cudaMemcpyDeviceToHost, st));
  CCG_CUDA(ctx, cudaStreamSynchronize(st));
  CCG_CUDA(ctx, cudaGetLastError());
  int rc = CCG_OK;
  for (int64_t b = 0; b < nb; ++b)
    if (reports[b].termination < 0) rc = reports[b].termination;
  return rc;
}

void free_problem(ccg_intr_problem* p) {
  if (!p) return;
  if (p->ctx) cudaSetDevice(p->ctx->device);
  if (p->graph_exec) cudaGraphExecDestroy(p->graph_exec);
  for (void* a : p->allocs) pool_free(p->ctx, a);
  delete p;
}

bool bad_inputs(const float* u, const float* v, const float* X, const float* Y, const float* Z,
                const int64_t* view_offsets, int64_t n_views) {
  if (!view_offsets || n_views < 0) return true;
  const int64_t n = view_offsets[n_views] - view_offsets[0];
  if (n < 0) return true;
  if (n > 0 && (!u || !v || !X || !Y || !Z)) return true;
  for (int64_t i = 0; i < n_views; ++i)
    if (view_offsets[i + 1] < view_offsets[i]) return true;
  return false;
}

}  // namespace

extern "C" {

int ccg_intrinsics_create(ccg_ctx* ctx, const float* u, const float* v, const float* X,
                          const float* Y, const float* Z, const int64_t* view_offsets,
                          int64_t n_views, const int64_t* problem_offsets, int64_t n_problems,
                          const double* intr, const double* q, const double* t,
                          uint32_t const_intr_mask, int shared, ccg_intr_problem** out) {
  if (!ctx || !out) return CCG_ERR_BAD_ARG;
  *out = nullptr;
  if (bad_inputs(u, v, X, Y, Z, view_offsets, n_views) || !intr || n_problems < 1 ||
      (n_views > 0 && (!q || !t)) || n_views > 0x7fffffff)
    return set_error(ctx, CCG_ERR_BAD_ARG, "ccg_intrinsics_create: bad argument");
  if (shared && n_problems != 1) return set_error(ctx, CCG_ERR_UNSUPPORTED, "shared requires n_problems == 1");
  const int64_t one[2] = {0, n_views};
  if (!problem_offsets) {
    if (n_problems != 1) return set_error(ctx, CCG_ERR_BAD_ARG, "problem_offsets required");
    problem_offsets = one;
  }
  CCG_CUDA(ctx, cudaSetDevice(ctx->device));
  ccg_intr_problem* p = new ccg_intr_problem();
  const int rc = build_problem(ctx, u, v, X, Y, Z, view_offsets, n_views, problem_offsets, n_problems,
                               intr, q, t, const_intr_mask, shared, p);
  if (rc != CCG_OK) {
    free_problem(p);
    return rc;
  }
  *out = p;
  return CCG_OK;
}

int ccg_intrinsics_reset(ccg_intr_problem* p) {
  if (!p) return CCG_ERR_BAD_ARG;
  ccg_opts o;
  ccg_default_opts(&o);
  CCG_CUDA(p->ctx, cudaSetDevice(p->ctx->device));
  return reset_problem(p, p->params_set ? p->params : to_params(o));
}

int ccg_intrinsics_run(ccg_intr_problem* p, const ccg_opts* opts, ccg_timing* timing) {
  if (!p) return CCG_ERR_BAD_ARG;
  ccg_opts o;
  if (opts) o = *opts; else ccg_default_opts(&o);
  return run_problem(p, o, timing);
}

int ccg_intrinsics_time_sweep(ccg_intr_problem* p, int n, float* ms_per_launch) {
  if (!p || n < 1 || !ms_per_launch) return CCG_ERR_BAD_ARG;
  ccg_ctx* ctx = p->ctx;
  CCG_CUDA(ctx, cudaSetDevice(ctx->device));
  ccg_opts o;
  ccg_default_opts(&o);
  const LmParams P = p->params_set ? p->params : to_params(o);
  CCG_TRY(reset_problem(p, P));  // status 0, first = 1: every launch evaluates the initial point
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  // the sweep alone (no accept/reject in between): every launch evaluates the same (initial) point
  auto sweep = [&]() {
    if (P.hessian_fp64 == 0) launch_sweep_ws_t<false>(p, nullptr);
    else k_intr_sweep<<<p->n_items, kWarps * 32, 0, ctx->stream>>>(p->dev, nullptr);
    ctx->launches++;
  };
  sweep();  // warm-up
  cudaEventRecord(e0, ctx->stream);
  for (int i = 0; i < n; ++i) sweep();
  cudaEventRecord(e1, ctx->stream);
  CCG_CUDA(ctx, cudaEventSynchronize(e1));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *ms_per_launch = ms / (float)n;
  CCG_CUDA(ctx, cudaGetLastError());
  return CCG_OK;
}

int ccg_intrinsics_download(ccg_intr_problem* p, double* intr, double* q, double* t,
                            ccg_report* reports) {
  if (!p) return CCG_ERR_BAD_ARG;
  return download_problem(p, intr, q, t, reports);
}

int ccg_intrinsics_history(ccg_intr_problem* p, int64_t b, double* rows, int cap, int* n_rows) {
  if (!p || b < 0 || b >= p->n_problems || !rows || !n_rows) return CCG_ERR_BAD_ARG;
  ccg_ctx* ctx = p->ctx;
  CCG_CUDA(ctx, cudaSetDevice(ctx->device));
  IntrState st;
  CCG_CUDA(ctx, cudaMemcpy(&st, p->dev.ps + b, sizeof(IntrState), cudaMemcpyDeviceToHost));
  const int total = st.lm.n_hist, n = std::min(std::min(total, kHist), cap);
  for (int i = 0; i < n; ++i) {
    const int src = (total - n + i) % kHist;
    for (int k = 0; k < 6; ++k) rows[i * 6 + k] = st.lm.hist[src][k];
  }
  *n_rows = n;
  return CCG_OK;
}

int64_t ccg_intrinsics_device_bytes(const ccg_intr_problem* p) { return p ? p->bytes : 0; }

void ccg_intrinsics_destroy(ccg_intr_problem* p) { free_problem(p); }

int ccg_intrinsics_solve_batch(ccg_ctx* ctx, const float* u, const float* v, const float* X,
                               const float* Y, const float* Z, const int64_t* view_offsets,
                               int64_t n_views, const int64_t* problem_offsets,
                               int64_t n_problems, double* intr, double* q, double* t,
                               uint32_t const_intr_mask, const ccg_opts* opts,
                               ccg_report* reports) {
  ccg_intr_problem* p = nullptr;
  int rc = ccg_intrinsics_create(ctx, u, v, X, Y, Z, view_offsets, n_views, problem_offsets, n_problems,
                                 intr, q, t, const_intr_mask, 0, &p);
  if (rc != CCG_OK) return rc;
  p->one_shot = true;
  rc = ccg_intrinsics_run(p, opts, nullptr);
  if (rc == CCG_OK) rc = download_problem(p, intr, q, t, reports);
  free_problem(p);
  return rc;
}

int ccg_intrinsics_solve(ccg_ctx* ctx, const float* u, const float* v, const float* X,
                         const float* Y, const float* Z, const int64_t* view_offsets,
                         int64_t n_views, double* intr, double* q, double* t,
                         uint32_t const_intr_mask, const ccg_opts* opts, ccg_report* report) {
  const int64_t poff[2] = {0, n_views};
  return ccg_intrinsics_solve_batch(ctx, u, v, X, Y, Z, view_offsets, n_views, poff, 1, intr, q, t,
                                    const_intr_mask, opts, report);
}

// Estimate = closed-form initialisation on the device + LM solve, the whole of Calibrator::Estimate
// (reference src/calibrator.cpp:47-68) behind one call: the observations are uploaded once.
int ccg_intrinsics_estimate_batch(ccg_ctx* ctx, const float* u, const float* v, const float* X,
                                  const float* Y, const float* Z, const int64_t* view_offsets,
                                  int64_t n_views, const int64_t* problem_offsets, int64_t n_problems,
                                  double* intr, double* q, double* t, double* intr_init,
                                  uint32_t const_intr_mask, const ccg_opts* opts, ccg_report* reports) {
  if (!ctx || !intr || n_problems < 1 || n_views < 0) return CCG_ERR_BAD_ARG;
  if (n_problems > 65535) return set_error(ctx, CCG_ERR_UNSUPPORTED, "more than 65535 problems in one batch");
  const int64_t one[2] = {0, n_views};
  if (!problem_offsets) {
    if (n_problems != 1) return set_error(ctx, CCG_ERR_BAD_ARG, "problem_offsets required");
    problem_offsets = one;
  }
  // placeholders: the init kernels overwrite the starting point on the device
  std::vector<double> intr0((size_t)n_problems * 9, 0.0), q0((size_t)std::max<int64_t>(n_views, 1) * 4, 0.0),
      t0((size_t)std::max<int64_t>(n_views, 1) * 3, 0.0);
  ccg_intr_problem* p = nullptr;
  int rc = ccg_intrinsics_create(ctx, u, v, X, Y, Z, view_offsets, n_views, problem_offsets, n_problems,
                                 intr0.data(), q0.data(), t0.data(), const_intr_mask, 0, &p);
  if (rc != CCG_OK) return rc;
  p->one_shot = true;
  auto body = [&]() -> int {
    cudaStream_t st = ctx->stream;
    double* d_H = nullptr;
    int* d_status = nullptr;
    CCG_TRY(dev_alloc(p, &d_H, (size_t)std::max<int64_t>(n_views, 1) * 9));
    CCG_TRY(dev_alloc(p, &d_status, (size_t)n_problems));
    int64_t max_views = 1;
    for (int64_t b = 0; b < n_problems; ++b)
      max_views = std::max(max_views, problem_offsets[b + 1] - problem_offsets[b]);
    if (n_views > 0)
      k_init_homography<<<(int)((n_views + kInitWarps - 1) / kInitWarps), kInitWarps * 32, 0, st>>>(p->dev, n_views, d_H);
    k_init_zhang<<<(int)n_problems, 256, 0, st>>>(d_H, p->d_problem_off, p->d_meas_scale, p->d_intr0, d_status);
    k_init_extrinsics<<<dim3((unsigned)((max_views + 127) / 128), (unsigned)n_problems), 128, 0, st>>>(
        d_H, p->d_problem_off, (int)n_problems, p->d_intr0, p->d_pose0);
    ctx->launches += 3;
    std::vector<int> status((size_t)n_problems, 0);
    CCG_CUDA(ctx, cudaMemcpyAsync(status.data(), d_status, sizeof(int) * n_problems, cudaMemcpyDeviceToHost, st));
    if (intr_init)
      CCG_CUDA(ctx, cudaMemcpyAsync(intr_init, p->d_intr0, sizeof(double) * 9 * n_problems, cudaMemcpyDeviceToHost, st));
    CCG_CUDA(ctx, cudaStreamSynchronize(st));
    CCG_CUDA(ctx, cudaGetLastError());
    for (int64_t b = 0; b < n_problems; ++b)
      if (status[(size_t)b] != 0)
        return set_error(ctx, status[(size_t)b], "closed-form initialisation failed (fewer than 3 usable views, or no real solution for K)");
    CCG_TRY(ccg_intrinsics_run(p, opts, nullptr));
    return download_problem(p, intr, q, t, reports);
  };
  rc = body();
  free_problem(p);
  return rc;
}

// EstimateHomography (reference src/geometry.cpp:70-107) for many point sets at once, on the device: the kernel the
// initialisation of ccg_intrinsics_estimate runs, exposed so that it can be checked against the reference's own
// known-answer test (src/test_geometry.cpp:172-194).
int ccg_homography_batch(ccg_ctx* ctx, const float* x1, const float* y1, const float* x2, const float* y2,
                         const int64_t* set_offsets, int64_t n_sets, double* H) {
  if (!ctx || !set_offsets || n_sets < 0 || (n_sets > 0 && (!x1 || !y1 || !x2 || !y2 || !H))) return CCG_ERR_BAD_ARG;
  if (n_sets == 0) return CCG_OK;
  CCG_CUDA(ctx, cudaSetDevice(ctx->device));
  const int64_t o0 = set_offsets[0], n = set_offsets[n_sets] - o0;
  for (int64_t i = 0; i < n_sets; ++i)
    if (set_offsets[i + 1] < set_offsets[i]) return set_error(ctx, CCG_ERR_BAD_ARG, "set_offsets");
  cudaStream_t st = ctx->stream;
  float* d_pts = nullptr;
  int64_t* d_off = nullptr;
  double* d_H = nullptr;
  auto body = [&]() -> int {
    cudaError_t e = pool_alloc(ctx, (void**)&d_pts, sizeof(float) * 4 * std::max<int64_t>(n, 1));
    if (e == cudaSuccess) e = pool_alloc(ctx, (void**)&d_off, sizeof(int64_t) * (n_sets + 1));
    if (e == cudaSuccess) e = pool_alloc(ctx, (void**)&d_H, sizeof(double) * 9 * n_sets);
    if (e != cudaSuccess) return set_error(ctx, CCG_ERR_CUDA, "device allocation", e);
    std::vector<int64_t> off((size_t)n_sets + 1);
    for (int64_t i = 0; i <= n_sets; ++i) off[(size_t)i] = set_offsets[i] - o0;
    IntrDev d;
    std::memset(&d, 0, sizeof(d));
    d.X = d_pts;
    d.Y = d_pts + n;
    d.u = d_pts + 2 * n;
    d.v = d_pts + 3 * n;
    d.view_off = d_off;
    CCG_CUDA(ctx, cudaMemcpyAsync(d_pts, x1 + o0, sizeof(float) * n, cudaMemcpyHostToDevice, st));
    CCG_CUDA(ctx, cudaMemcpyAsync(d_pts + n, y1 + o0, sizeof(float) * n, cudaMemcpyHostToDevice, st));
    CCG_CUDA(ctx, cudaMemcpyAsync(d_pts + 2 * n, x2 + o0, sizeof(float) * n, cudaMemcpyHostToDevice, st));
    CCG_CUDA(ctx, cudaMemcpyAsync(d_pts + 3 * n, y2 + o0, sizeof(float) * n, cudaMemcpyHostToDevice, st));
    CCG_CUDA(ctx, cudaMemcpyAsync(d_off, off.data(), sizeof(int64_t) * off.size(), cudaMemcpyHostToDevice, st));
    k_init_homography<<<(int)((n_sets + kInitWarps - 1) / kInitWarps), kInitWarps * 32, 0, st>>>(d, n_sets, d_H);
    ctx->launches++;
    CCG_CUDA(ctx, cudaMemcpyAsync(H, d_H, sizeof(double) * 9 * n_sets, cudaMemcpyDeviceToHost, st));
    CCG_CUDA(ctx, cudaStreamSynchronize(st));
    CCG_CUDA(ctx, cudaGetLastError());
    return CCG_OK;
  };
  const int rc = body();
  if (d_pts) pool_free(ctx, d_pts);
  if (d_off) pool_free(ctx, d_off);
  if (d_H) pool_free(ctx, d_H);
  return rc;
}

int ccg_intrinsics_estimate(ccg_ctx* ctx, const float* u, const float* v, const float* X, const float* Y,
                            const float* Z, const int64_t* view_offsets, int64_t n_views, double* intr,
                            double* q, double* t, double* intr_init, uint32_t const_intr_mask,
                            const ccg_opts* opts, ccg_report* report) {
  const int64_t poff[2] = {0, n_views};
  return ccg_intrinsics_estimate_batch(ctx, u, v, X, Y, Z, view_offsets, n_views, poff, 1, intr, q, t, intr_init,
                                       const_intr_mask, opts, report);
}

int ccg_intrinsics_eval(ccg_ctx* ctx, const float* u, const float* v, const float* X,
                        const float* Y, const float* Z, const int64_t* view_offsets,
                        int64_t n_views, const double* intr, const double* q, const double* t,
                        double* res, double* Vb, double* Wb, double* gvb, double* U, double* gu,
                        double* cost, int hessian_fp64) {
  ccg_intr_problem* p = nullptr;
  const int64_t poff[2] = {0, n_views};
  int rc = ccg_intrinsics_create(ctx, u, v, X, Y, Z, view_offsets, n_views, poff, 1, intr, q, t, 0, 0, &p);
  if (rc != CCG_OK) return rc;
  auto body = [&]() -> int {
    ccg_opts o;
    ccg_default_opts(&o);
    o.hessian_fp64 = hessian_fp64;
    const LmParams P = to_params(o);
    CCG_TRY(reset_problem(p, P));
    double* d_res = nullptr;
    if (res) CCG_TRY(dev_alloc(p, &d_res, (size_t)p->n_obs * 2));
    Profiler prof;
    prof.ctx = ctx;
    int nar = 0;
    CCG_TRY(launch_sweep_decide(p, P, prof, d_res, &nar));
    CCG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    CCG_CUDA(ctx, cudaGetLastError());
    // the first sweep wrote the blocks of the (accepted) initial point into buffer 1
    std::vector<double> blk((size_t)n_views * kBlk);
    CCG_CUDA(ctx, cudaMemcpy(blk.data(), p->dev.blk[1], sizeof(double) * blk.size(), cudaMemcpyDeviceToHost));
    IntrState st;
    CCG_CUDA(ctx, cudaMemcpy(&st, p->dev.ps, sizeof(IntrState), cudaMemcpyDeviceToHost));
    if (res) CCG_CUDA(ctx, cudaMemcpy(res, d_res, sizeof(double) * 2 * p->n_obs, cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < n_views; ++i) {
      const double* b = &blk[(size_t)i * kBlk];
      if (Vb) {
        int idx = 0;
        for (int a = 0; a < 6; ++a)
          for (int c = a; c < 6; ++c) {
            Vb[i * 36 + a * 6 + c] = b[idx];
            Vb[i * 36 + c * 6 + a] = b[idx];
            ++idx;
          }
      }
      if (gvb) for (int a = 0; a < 6; ++a) gvb[i * 6 + a] = b[21 + a];
      if (Wb)
        for (int a = 0; a < 6; ++a) {
          for (int c = 0; c < 4; ++c) Wb[i * 54 + a * 9 + c] = b[27 + a * 4 + c];
          for (int c = 0; c < 5; ++c) Wb[i * 54 + a * 9 + 4 + c] = b[51 + a * 5 + c];
        }
    }
    if (U) std::memcpy(U, st.U, sizeof(double) * 81);
    if (gu) std::memcpy(gu, st.gu, sizeof(double) * 9);
    if (cost) *cost = st.lm.cost;
    return CCG_OK;
  };
  rc = body();
  free_problem(p);
  return rc;
}

}  // extern "C"
