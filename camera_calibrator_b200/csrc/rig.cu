// Rig hot path: residual + analytic Jacobian + Huber corrector + normal-equation
// blocks + Schur elimination of the per-frame rig poses onto the camera
// extrinsics + device-resident LM loop.
//
// Replaces, behind ccg_rig_*:
//   the per-observation autodiff evaluation of ReprojectionErrorExtrinsics
//       reference src/extrinsics_calibrator.cpp:51-84
//   HuberLoss(3.0f/500.0f) per residual block      reference :175-176  [ext Ceres corrector]
//   the Ceres problem build, ceres::Solve, per-observation final cost
//       reference src/extrinsics_calibrator.cpp:92-225
// Layout: the host sorts the observations of every frame by camera (stable) so
// that a (frame, camera) SEGMENT is contiguous; a warp pair owns a frame and
// walks its segments: warp A accumulates V_f(21) and U_cc(21), warp B W_fc(36),
// g_f(6), g_c(6).  Up to kMaxCams cameras (the reference has no bound, extrinsics_calibrator.cpp:9-17;
// the reduced system is dense (6C)^2 and is factored in shared memory, which sets the cap).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>

#include "common.cuh"

namespace ccg {

constexpr int kMaxCams = 16;
constexpr int kMaxN6 = 6 * kMaxCams;          // 96 reduced unknowns at most
constexpr int kRigWarps = 8;                  // sweep: 4 frame pairs per CTA
constexpr int kRigPart = 27 * kMaxCams + 9;   // 441: U_cc 21 + g_c 6 per camera | rho sum | sumsq | g'd | d'Dd | n
constexpr int kRigSchurCols = (kMaxN6 + 1 + 31) / 32;  // columns of [W | g_f] a lane of the elimination kernel owns
// elimination partial of an item, sized by the actual camera count: (6C + 1) columns x 6C rows of (W'Y | W'z),
// then g'D^-1g and the failure count
__host__ __device__ inline int rig_schur_part(int n6) { return (n6 + 1) * n6 + 2; }

struct RigState {
  LmState lm;
  double cam[2][kMaxCams][8];  // q4 t3 pad, [par] = current
  double du[kMaxN6];
  double Ucc[kMaxCams][21];    // block-diagonal U at the current point (packed upper 6x6)
  double gc[kMaxCams][6];
  int frozen[kMaxCams];
  int n_cams, item0, item1, pad_;
};

struct RigDev {
  const float *u, *v;       // sorted by (frame, camera)
  const int* pt;            // world point id per observation
  const float* Xw;          // 3 per world point
  const int64_t* seg_off;   // [F*C + 1]
  const WorkItem* items;    // frame ranges
  double* pose[2];          // per frame q4 t3 pad
  double* blk[2];           // per frame: V21 | g6 | W 6 x 6C
  double* yz;               // per frame: Y 6 x 6C | z 6 | g 6 | D 6
  double* sweep_part;
  double* schur_part;
  double* red_sweep;
  double* red_schur;
  RigState* ps;
  int* n_done;
  int n_cams, blk_stride, yz_stride;
  int schur_stride;  // rig_schur_part(6 n_cams)
};

// 2 (y x m)
__device__ __forceinline__ void cross2(const double (&y)[3], const double (&m)[3], double* out) {
  out[0] = 2.0 * (y[1] * m[2] - y[2] * m[1]);
  out[1] = 2.0 * (y[2] * m[0] - y[0] * m[2]);
  out[2] = 2.0 * (y[0] * m[1] - y[1] * m[0]);
}

struct RigGeom {
  double ru, rv;        // residual scaled by sqrt(rho')
  double rho, s;        // robust cost term (rho(s)) and raw squared norm
  double jfu[6], jfv[6], jcu[6], jcv[6];
};

__device__ __forceinline__ void rig_geometry(const double (&Rf)[9], const double (&tf)[3],
                                             const double (&Rc)[9], const double (&tc)[3], double X,
                                             double Y, double Z, double mu, double mv, double huber_a,
                                             RigGeom& g) {
  double yf[3], yc[3];
  yf[0] = Rf[0] * X + Rf[1] * Y + Rf[2] * Z;
  yf[1] = Rf[3] * X + Rf[4] * Y + Rf[5] * Z;
  yf[2] = Rf[6] * X + Rf[7] * Y + Rf[8] * Z;
  const double p0 = yf[0] + tf[0], p1 = yf[1] + tf[1], p2 = yf[2] + tf[2];
  yc[0] = Rc[0] * p0 + Rc[1] * p1 + Rc[2] * p2;
  yc[1] = Rc[3] * p0 + Rc[4] * p1 + Rc[5] * p2;
  yc[2] = Rc[6] * p0 + Rc[7] * p1 + Rc[8] * p2;
  const double iz = fast_rcp(yc[2] + tc[2]);
  const double xn = (yc[0] + tc[0]) * iz, yn = (yc[1] + tc[1]) * iz;
  double ru = xn - mu, rv = yn - mv;
  const double s = ru * ru + rv * rv;
  double rho = s, w = 1.0;
  if (huber_a > 0.0 && s > huber_a * huber_a) {
    const double r = sqrt(s);
    rho = 2.0 * huber_a * r - huber_a * huber_a;
    w = sqrt(huber_a / r);  // sqrt(rho'): [ext] Ceres corrector for rho'' <= 0
  }
  g.s = s;
  g.rho = rho;
  g.ru = ru * w;
  g.rv = rv * w;
  const double wz = w * iz;
  // rows of d(xn,yn)/d(camera-frame point), scaled by sqrt(rho')
  const double m0[3] = {wz, 0.0, -wz * xn};
  const double m1[3] = {0.0, wz, -wz * yn};
  cross2(yc, m0, g.jcu);
  cross2(yc, m1, g.jcv);
  g.jcu[3] = m0[0]; g.jcu[4] = m0[1]; g.jcu[5] = m0[2];
  g.jcv[3] = m1[0]; g.jcv[4] = m1[1]; g.jcv[5] = m1[2];
  // chain through the camera rotation: row * Rc
  double f0[3], f1[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    f0[j] = m0[0] * Rc[j] + m0[2] * Rc[6 + j];
    f1[j] = m1[1] * Rc[3 + j] + m1[2] * Rc[6 + j];
  }
  cross2(yf, f0, g.jfu);
  cross2(yf, f1, g.jfv);
  g.jfu[3] = f0[0]; g.jfu[4] = f0[1]; g.jfu[5] = f0[2];
  g.jfv[3] = f1[0]; g.jfv[4] = f1[1]; g.jfv[5] = f1[2];
}

// candidate frame pose = current (+) (-z - Y du)
__device__ __forceinline__ void frame_prologue(const RigDev& d, int par, int64_t f, int lane, bool is_a,
                                               const double* s_du, int n6, double (&R)[9],
                                               double (&t)[3], double& gTd, double& dDd) {
  const double* yz = d.yz + (size_t)f * d.yz_stride;
  double dl = 0.0;
  if (lane < 6) {
    dl = -yz[6 * n6 + lane];
    for (int j = 0; j < n6; ++j) dl -= yz[lane * n6 + j] * s_du[j];
    if (is_a) {
      gTd += yz[6 * n6 + 6 + lane] * dl;
      dDd += dl * dl * yz[6 * n6 + 12 + lane];
    }
  }
  const double d0 = __shfl_sync(0xffffffffu, dl, 0), d1 = __shfl_sync(0xffffffffu, dl, 1),
               d2 = __shfl_sync(0xffffffffu, dl, 2), d3 = __shfl_sync(0xffffffffu, dl, 3),
               d4 = __shfl_sync(0xffffffffu, dl, 4), d5 = __shfl_sync(0xffffffffu, dl, 5);
  const double* pc = d.pose[par] + (size_t)f * 8;
  double q[4] = {pc[0], pc[1], pc[2], pc[3]};
  double qn[4];
  quat_plus(q, d0, d1, d2, qn);
  t[0] = pc[4] + d3;
  t[1] = pc[5] + d4;
  t[2] = pc[6] + d5;
  if (is_a && lane < 7) {
    double* pn = d.pose[par ^ 1] + (size_t)f * 8;
    const double val = lane == 0 ? qn[0] : lane == 1 ? qn[1] : lane == 2 ? qn[2] : lane == 3 ? qn[3]
                     : lane == 4 ? t[0] : lane == 5 ? t[1] : t[2];
    pn[lane] = val;
  }
  quat_to_rot(qn, R);
}

__global__ void __launch_bounds__(kRigWarps * 32, 1)
k_rig_sweep(RigDev d, double huber_a, double* __restrict__ res_out, double* __restrict__ cost_out) {
  __shared__ double s_cam[kMaxCams][12];                     // candidate camera R (9) t (3)
  __shared__ double s_du[kMaxN6];
  __shared__ double s_acc[kRigWarps / 2][kMaxCams][27];      // per pair: U_cc 21 (warp A) | g_c 6 (warp B)
  __shared__ double s_sc[kRigWarps / 2][8];                  // per pair scalars
  __shared__ int s_meta[2];
  const WorkItem it = d.items[blockIdx.x];
  RigState& ps = d.ps[0];
  const int C = d.n_cams, n6 = 6 * C;
  if (threadIdx.x == 0) {
    s_meta[0] = ps.lm.par;
    s_meta[1] = (ps.lm.status == 0 && ps.lm.solve_ok) ? 1 : 0;
  }
  __syncthreads();
  if (!s_meta[1] && !cost_out) return;
  const int par = s_meta[0];
  // cost_out != NULL: final per-observation cost pass at the CURRENT point (no step)
  const int src = cost_out ? par : par ^ 1;
  if (threadIdx.x < C) {
    double q[4] = {ps.cam[src][threadIdx.x][0], ps.cam[src][threadIdx.x][1], ps.cam[src][threadIdx.x][2],
                   ps.cam[src][threadIdx.x][3]};
    double R[9];
    quat_to_rot(q, R);
    for (int i = 0; i < 9; ++i) s_cam[threadIdx.x][i] = R[i];
    for (int i = 0; i < 3; ++i) s_cam[threadIdx.x][9 + i] = ps.cam[src][threadIdx.x][4 + i];
  }
  if (threadIdx.x < n6) s_du[threadIdx.x] = cost_out ? 0.0 : ps.du[threadIdx.x];
  for (int i = threadIdx.x; i < (kRigWarps / 2) * kMaxCams * 27; i += blockDim.x) (&s_acc[0][0][0])[i] = 0.0;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = warp >> 1;
  const bool is_a = (warp & 1) == 0;
  double rho_sum = 0.0, sumsq = 0.0, nobs = 0.0, gTd = 0.0, dDd = 0.0;

  for (int64_t f = it.v0 + pair; f < it.v1; f += kRigWarps / 2) {
    double Rf[9], tf[3];
    if (cost_out) {
      const double* pc = d.pose[par] + (size_t)f * 8;
      double q[4] = {pc[0], pc[1], pc[2], pc[3]};
      quat_to_rot(q, Rf);
      tf[0] = pc[4]; tf[1] = pc[5]; tf[2] = pc[6];
    } else {
      frame_prologue(d, par, f, lane, is_a, s_du, n6, Rf, tf, gTd, dDd);
    }
    double* blk = d.blk[par ^ 1] + (size_t)f * d.blk_stride;
    double accf[21];  // warp A: V_f ; warp B: g_f in accf[0..5]
#pragma unroll
    for (int i = 0; i < 21; ++i) accf[i] = 0.0;
    // Loads run two steps ahead of the arithmetic (the world point is a gather through the point id, two dependent
    // L2 round trips): step k computes on registers loaded during step k - 1 while the gather of step k + 1 and the
    // point id of step k + 2 are in flight.  A segment's first two steps are primed before the reductions of the
    // segment before it.  Out-of-range steps re-read the segment's last observation (never used).
    int64_t o0 = d.seg_off[f * C], o1 = d.seg_off[f * C + 1];
    int ip_a = 0, ip_b = 0;         // point ids of this lane's steps 0 and 1 of the segment
    float mu_n = 0.f, mv_n = 0.f;   // measurement of step 0
    auto prime_ids = [&](int64_t a0, int64_t a1) {
      if (a0 < a1) {
        const int64_t oa = min(a0 + lane, a1 - 1), ob = min(a0 + lane + 32, a1 - 1);
        ip_a = __ldg(d.pt + oa);
        ip_b = __ldg(d.pt + ob);
        mu_n = __ldg(d.u + oa);
        mv_n = __ldg(d.v + oa);
      }
    };
    prime_ids(o0, o1);
    for (int c = 0; c < C; ++c) {
      double Rc[9], tc[3];
#pragma unroll
      for (int i = 0; i < 9; ++i) Rc[i] = s_cam[c][i];
#pragma unroll
      for (int i = 0; i < 3; ++i) tc[i] = s_cam[c][9 + i];
      const int64_t o2 = c + 1 < C ? d.seg_off[f * C + c + 2] : o1;  // end of the NEXT segment
      double acc[42];  // warp A: U_cc in acc[0..20] ; warp B: W_fc 36 | g_c 6
#pragma unroll
      for (int i = 0; i < 42; ++i) acc[i] = 0.0;
      float Xn = 0.f, Yn = 0.f, Zn = 0.f;
      int ip_nn = ip_b;  // point id two steps ahead
      if (o0 < o1) {
        Xn = __ldg(d.Xw + 3 * (size_t)ip_a);
        Yn = __ldg(d.Xw + 3 * (size_t)ip_a + 1);
        Zn = __ldg(d.Xw + 3 * (size_t)ip_a + 2);
      }
      // one step: the observation at o from registers
      auto step = [&](int64_t o, float mu, float mv, float Xf, float Yf, float Zf) {
        RigGeom g;
        rig_geometry(Rf, tf, Rc, tc, (double)Xf, (double)Yf, (double)Zf, (double)mu, (double)mv, huber_a, g);
        if (cost_out) {
          if (is_a) cost_out[o] = 0.5 * g.rho;
          return;
        }
        if (is_a) {
          if (res_out) {
            res_out[2 * o] = g.ru;
            res_out[2 * o + 1] = g.rv;
          }
          int idx = 0;
#pragma unroll
          for (int a = 0; a < 6; ++a)
#pragma unroll
            for (int b = a; b < 6; ++b) {
              accf[idx] = fma(g.jfu[a], g.jfu[b], fma(g.jfv[a], g.jfv[b], accf[idx]));
              acc[idx] = fma(g.jcu[a], g.jcu[b], fma(g.jcv[a], g.jcv[b], acc[idx]));
              ++idx;
            }
          rho_sum += g.rho;
          sumsq += g.s;
          nobs += 1.0;
        } else {
#pragma unroll
          for (int a = 0; a < 6; ++a) {
#pragma unroll
            for (int b = 0; b < 6; ++b)
              acc[a * 6 + b] = fma(g.jfu[a], g.jcu[b], fma(g.jfv[a], g.jcv[b], acc[a * 6 + b]));
            accf[a] = fma(g.jfu[a], g.ru, fma(g.jfv[a], g.rv, accf[a]));
            acc[36 + a] = fma(g.jcu[a], g.ru, fma(g.jcv[a], g.rv, acc[36 + a]));
          }
        }
      };
      // the loads of step k + 1 (gather through the id fetched a step earlier) and the id of step k + 2
      auto fetch = [&](int64_t o, int& ip_next, float& mu, float& mv, float& Xf, float& Yf, float& Zf) {
        const int64_t on = min(o + 32, o1 - 1), onn = min(o + 64, o1 - 1);
        const int ip1 = ip_next;
        ip_next = __ldg(d.pt + onn);
        mu = __ldg(d.u + on);
        mv = __ldg(d.v + on);
        Xf = __ldg(d.Xw + 3 * (size_t)ip1);
        Yf = __ldg(d.Xw + 3 * (size_t)ip1 + 1);
        Zf = __ldg(d.Xw + 3 * (size_t)ip1 + 2);
      };
      // two steps per trip on two register sets (A: mu_n .. Zn, B): the sets swap roles instead of being copied -
      // with a copy, ptxas moved the "next" registers into the "current" ones at the TOP of the step and so waited
      // for the loads it had just issued (14 % of all stall samples on that one move)
      {
        float mu_b = 0.f, mv_b = 0.f, Xb = 0.f, Yb = 0.f, Zb = 0.f;
        int64_t o = o0 + lane;
        while (o < o1) {
          fetch(o, ip_nn, mu_b, mv_b, Xb, Yb, Zb);
          step(o, mu_n, mv_n, Xn, Yn, Zn);
          o += 32;
          if (!(o < o1)) break;
          fetch(o, ip_nn, mu_n, mv_n, Xn, Yn, Zn);
          step(o, mu_b, mv_b, Xb, Yb, Zb);
          o += 32;
        }
      }
      // the next segment's first point ids and measurement, in flight during the reductions below
      o0 = o1;
      o1 = o2;
      if (c + 1 < C) prime_ids(o0, o1);
      if (cost_out) continue;
      if (is_a) {
        double tmp[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) tmp[i] = i < 21 ? acc[i] : 0.0;
        const double r = warp_reduce32(tmp, lane);
        if (lane < 21) s_acc[pair][c][lane] += r;
      } else {
        double out[2];
        warp_reduce_array<42>(acc, lane, out);
        // element e = a*6+b of W_fc -> blk[27 + a*n6 + 6c + b]
        {
          const int a = lane / 6, b = lane - 6 * a;
          blk[27 + a * n6 + 6 * c + b] = out[0];
        }
        if (lane < 4) {
          const int e = 32 + lane, a = e / 6, b = e - 6 * a;
          blk[27 + a * n6 + 6 * c + b] = out[1];
        } else if (lane < 10) {
          s_acc[pair][c][21 + lane - 4] += out[1];
        }
      }
    }
    if (cost_out) continue;
    {
      double tmp[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) tmp[i] = i < 21 ? accf[i] : 0.0;
      const double r = warp_reduce32(tmp, lane);
      if (is_a) {
        if (lane < 21) blk[lane] = r;
      } else {
        if (lane < 6) blk[21 + lane] = r;
      }
    }
  }
  if (cost_out) return;
  if (is_a) {
    double tmp[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) tmp[i] = 0.0;
    tmp[0] = rho_sum; tmp[1] = sumsq; tmp[2] = gTd; tmp[3] = dDd; tmp[4] = nobs;
    const double r = warp_reduce32(tmp, lane);
    if (lane < 8) s_sc[pair][lane] = lane < 5 ? r : 0.0;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < kRigPart; c += blockDim.x) {
    double s = 0.0;
    if (c < 27 * kMaxCams) {
      const int cam = c / 27, e = c - 27 * cam;
      if (cam < C)
        for (int p = 0; p < kRigWarps / 2; ++p) s += s_acc[p][cam][e];
    } else if (c < 27 * kMaxCams + 5) {
      for (int p = 0; p < kRigWarps / 2; ++p) s += s_sc[p][c - 27 * kMaxCams];
    }
    d.sweep_part[(size_t)blockIdx.x * kRigPart + c] = s;
  }
}

// --------------------------------------------------------------------------
// Default rig sweep: FP64 residual, cost and gradient; J'J on the tensor cores.
//
// Same split as the single-camera sweep (intr_sweep.cuh): the residual, the robust cost and the gradients
// g_f, g_c = J'r decide WHERE the LM iteration converges and stay FP64; the Gauss-Newton blocks V_f, W_fc, U_cc =
// J'J only decide how fast, and per (frame, camera) segment they are one dense contraction
//     [V W; W' U] = sum over the segment's 2 n residual rows of (row)'(row),   row = [J_frame(6) | J_camera(6)],
// a 12 x 12 x 2n product that fits one 16 x 16 tile.  One WARP per frame (four per CTA): a lane owns an observation,
// computes the FP64 geometry once (the all-FP64 kernel computes it in both warps of a pair), accumulates its share of
// g_f / g_c, rounds the two Jacobian rows to FP16 pairs (u row, v row) per column - 11 significant bits, as TF32;
// the entries are O(1) in normalised image coordinates, no scaling needed - into a warp-private shared-memory tile, and
// the warp contracts the 32 observations with 8 mma.sync.m16n8k16 (FP32 accumulation); the A fragment doubles as the
// B fragment because the product is J'J.  84 + 96 FP64 accumulation FMAs per observation and warp pair become
// 24, the 90 FP64 accumulators 16 floats, and the kernel needs 130 registers less (more warps per SM).
// ccg_opts.hessian_fp64 = 1 keeps the all-FP64 kernel (k_rig_sweep), which also serves ccg_rig_eval and the final
// per-observation cost pass.
// --------------------------------------------------------------------------
constexpr int kRtWarps = 4;             // warps (frames in flight) per CTA
constexpr int kRtRow = 20;              // 32-bit words per tile row: 16 + 4 (conflict-free STS.128 / LDS.64)
constexpr int kRtTile = 32 * kRtRow;    // words per warp tile: a row per observation

__device__ __forceinline__ unsigned rt_pack(float lo, float hi) {  // (lo, hi) -> f16x2, round to nearest, saturating
  unsigned r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void rt_mma(float (&c)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0,
                                       unsigned b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// candidate frame pose = current (+) (-z - Y du), all 32 lanes: lane j takes column j of Y (the one-lane-per-row
// version of frame_prologue walks 6 C dependent loads per row)
__device__ __forceinline__ void frame_prologue_warp(const RigDev& d, int par, int64_t f, int lane, const double* s_du,
                                                    int n6, double (&R)[9], double (&t)[3], double& gTd, double& dDd) {
  const double* yz = d.yz + (size_t)f * d.yz_stride;
  double part[8];
#pragma unroll
  for (int a = 0; a < 8; ++a) part[a] = 0.0;
  for (int j = lane; j < n6; j += 32) {
    const double duj = s_du[j];
#pragma unroll
    for (int a = 0; a < 6; ++a) part[a] = fma(yz[a * n6 + j], duj, part[a]);
  }
  double r = group_reduce<8, double>(part, lane & 7);
  r += __shfl_xor_sync(0xffffffffu, r, 8);
  r += __shfl_xor_sync(0xffffffffu, r, 16);
  double dl = 0.0;
  if (lane < 6) {
    dl = -yz[6 * n6 + lane] - r;
    gTd += yz[6 * n6 + 6 + lane] * dl;
    dDd += dl * dl * yz[6 * n6 + 12 + lane];
  }
  const double d0 = __shfl_sync(0xffffffffu, dl, 0), d1 = __shfl_sync(0xffffffffu, dl, 1),
               d2 = __shfl_sync(0xffffffffu, dl, 2), d3 = __shfl_sync(0xffffffffu, dl, 3),
               d4 = __shfl_sync(0xffffffffu, dl, 4), d5 = __shfl_sync(0xffffffffu, dl, 5);
  const double* pc = d.pose[par] + (size_t)f * 8;
  double q[4] = {pc[0], pc[1], pc[2], pc[3]};
  double qn[4];
  quat_plus(q, d0, d1, d2, qn);
  t[0] = pc[4] + d3;
  t[1] = pc[5] + d4;
  t[2] = pc[6] + d5;
  if (lane < 7) {
    double* pn = d.pose[par ^ 1] + (size_t)f * 8;
    const double val = lane == 0 ? qn[0] : lane == 1 ? qn[1] : lane == 2 ? qn[2] : lane == 3 ? qn[3]
                     : lane == 4 ? t[0] : lane == 5 ? t[1] : t[2];
    pn[lane] = val;
  }
  quat_to_rot(qn, R);
}

__global__ void __launch_bounds__(kRtWarps * 32, 2) k_rig_sweep_tc(RigDev d, double huber_a) {
  __shared__ double s_cam[kMaxCams][12];
  __shared__ double s_du[kMaxN6];
  __shared__ double s_acc[kRtWarps][kMaxCams][27];  // per warp: U_cc 21 | g_c 6
  __shared__ double s_sc[kRtWarps][8];
  __shared__ __align__(16) unsigned s_tile[kRtWarps][kRtTile];
  __shared__ int s_meta[2];
  const WorkItem it = d.items[blockIdx.x];
  RigState& ps = d.ps[0];
  const int C = d.n_cams, n6 = 6 * C;
  if (threadIdx.x == 0) {
    s_meta[0] = ps.lm.par;
    s_meta[1] = (ps.lm.status == 0 && ps.lm.solve_ok) ? 1 : 0;
  }
  __syncthreads();
  if (!s_meta[1]) return;
  const int par = s_meta[0];
  if (threadIdx.x < C) {
    double q[4] = {ps.cam[par ^ 1][threadIdx.x][0], ps.cam[par ^ 1][threadIdx.x][1], ps.cam[par ^ 1][threadIdx.x][2],
                   ps.cam[par ^ 1][threadIdx.x][3]};
    double R[9];
    quat_to_rot(q, R);
    for (int i = 0; i < 9; ++i) s_cam[threadIdx.x][i] = R[i];
    for (int i = 0; i < 3; ++i) s_cam[threadIdx.x][9 + i] = ps.cam[par ^ 1][threadIdx.x][4 + i];
  }
  if (threadIdx.x < n6) s_du[threadIdx.x] = ps.du[threadIdx.x];
  for (int i = threadIdx.x; i < kRtWarps * kMaxCams * 27; i += blockDim.x) (&s_acc[0][0][0])[i] = 0.0;
  for (int i = threadIdx.x; i < kRtWarps * kRtTile; i += blockDim.x) (&s_tile[0][0])[i] = 0u;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tig = lane & 3;
  const unsigned tile = (unsigned)__cvta_generic_to_shared(&s_tile[warp][0]);
  const unsigned row_w = tile + (unsigned)lane * (kRtRow * 4);                  // this lane's observation row
  const unsigned frag = tile + (unsigned)(2 * tig * kRtRow + 2 * g) * 4;       // its fragment of k-step 0
  // where this lane's eight product elements go: element i = (n-block i >> 2, slot i & 3) is (row, col) with
  // row = g (+ 8), col = 8 (i >> 2) + 2 tig (+ 1).  kind 1: V (upper), 2: W, 3: U (upper); ofs as used below.
  int kind[8], ofs[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int e = i & 3, row = g + ((e & 2) ? 8 : 0), col = (i >> 2) * 8 + 2 * tig + (e & 1);
    kind[i] = 0;
    ofs[i] = 0;
    if (row < 6 && col < 6 && col >= row) {
      kind[i] = 1;
      ofs[i] = row * 6 - row * (row - 1) / 2 + (col - row);
    } else if (row < 6 && col >= 6 && col < 12) {
      kind[i] = 2;
      ofs[i] = 27 + row * n6 + (col - 6);
    } else if (row >= 6 && row < 12 && col >= row && col < 12) {
      const int a = row - 6, b = col - 6;
      kind[i] = 3;
      ofs[i] = a * 6 - a * (a - 1) / 2 + (b - a);
    }
  }
  double rho_sum = 0.0, sumsq = 0.0, nobs = 0.0, gTd = 0.0, dDd = 0.0;

  for (int64_t f = it.v0 + warp; f < it.v1; f += kRtWarps) {
    double Rf[9], tf[3];
    frame_prologue_warp(d, par, f, lane, s_du, n6, Rf, tf, gTd, dDd);
    double* blk = d.blk[par ^ 1] + (size_t)f * d.blk_stride;
    double gf[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) gf[i] = 0.0;
    float vsum[2] = {0.f, 0.f};  // V elements of this lane (slots 0, 1 of n-block 0), summed over the cameras
    int64_t o0 = d.seg_off[f * C], o1 = d.seg_off[f * C + 1];
    int ip_a = 0, ip_b = 0;
    float mu_n = 0.f, mv_n = 0.f;
    auto prime_ids = [&](int64_t a0, int64_t a1) {
      if (a0 < a1) {
        const int64_t oa = min(a0 + lane, a1 - 1), ob = min(a0 + lane + 32, a1 - 1);
        ip_a = __ldg(d.pt + oa);
        ip_b = __ldg(d.pt + ob);
        mu_n = __ldg(d.u + oa);
        mv_n = __ldg(d.v + oa);
      }
    };
    prime_ids(o0, o1);
    for (int c = 0; c < C; ++c) {
      double Rc[9], tc[3];
#pragma unroll
      for (int i = 0; i < 9; ++i) Rc[i] = s_cam[c][i];
#pragma unroll
      for (int i = 0; i < 3; ++i) tc[i] = s_cam[c][9 + i];
      const int64_t o2 = c + 1 < C ? d.seg_off[f * C + c + 2] : o1;
      double gc[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) gc[i] = 0.0;
      float c0a[4] = {0.f, 0.f, 0.f, 0.f}, c0b[4] = {0.f, 0.f, 0.f, 0.f};
      float c1a[4] = {0.f, 0.f, 0.f, 0.f}, c1b[4] = {0.f, 0.f, 0.f, 0.f};
      float Xn = 0.f, Yn = 0.f, Zn = 0.f;
      int ip_nn = ip_b;
      if (o0 < o1) {
        Xn = __ldg(d.Xw + 3 * (size_t)ip_a);
        Yn = __ldg(d.Xw + 3 * (size_t)ip_a + 1);
        Zn = __ldg(d.Xw + 3 * (size_t)ip_a + 2);
      }
      auto fetch = [&](int64_t o, int& ip_next, float& mu, float& mv, float& Xf, float& Yf, float& Zf) {
        const int64_t on = min(o + 32, o1 - 1), onn = min(o + 64, o1 - 1);
        const int ip1 = ip_next;
        ip_next = __ldg(d.pt + onn);
        mu = __ldg(d.u + on);
        mv = __ldg(d.v + on);
        Xf = __ldg(d.Xw + 3 * (size_t)ip1);
        Yf = __ldg(d.Xw + 3 * (size_t)ip1 + 1);
        Zf = __ldg(d.Xw + 3 * (size_t)ip1 + 2);
      };
      // one step: 32 observations, lane's own at o (valid or padding)
      auto step = [&](int64_t o, float mu, float mv, float Xf, float Yf, float Zf) {
        const bool valid = o < o1;
        RigGeom gm;
        rig_geometry(Rf, tf, Rc, tc, (double)Xf, (double)Yf, (double)Zf, (double)mu, (double)mv, huber_a, gm);
        const double ru = valid ? gm.ru : 0.0, rv = valid ? gm.rv : 0.0;
        rho_sum += valid ? gm.rho : 0.0;
        sumsq += valid ? gm.s : 0.0;
        nobs += valid ? 1.0 : 0.0;
#pragma unroll
        for (int a = 0; a < 6; ++a) {
          gf[a] = fma(gm.jfu[a], ru, fma(gm.jfv[a], rv, gf[a]));
          gc[a] = fma(gm.jcu[a], ru, fma(gm.jcv[a], rv, gc[a]));
        }
        // tile row: product column m (0..5 frame, 6..11 camera, 12..15 zero) at word (m % 8) * 2 + m / 8
        const float vz = valid ? 1.0f : 0.0f;
        unsigned w[16];
#pragma unroll
        for (int m = 0; m < 16; ++m) {
          const int pos = (m % 8) * 2 + m / 8;
          if (m < 6) w[pos] = rt_pack((float)gm.jfu[m] * vz, (float)gm.jfv[m] * vz);
          else if (m < 12) w[pos] = rt_pack((float)gm.jcu[m - 6] * vz, (float)gm.jcv[m - 6] * vz);
          else w[pos] = 0u;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_w + 16 * i), "r"(w[4 * i]), "r"(w[4 * i + 1]),
                       "r"(w[4 * i + 2]), "r"(w[4 * i + 3])
                       : "memory");
        __syncwarp();
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          unsigned a0, a1, a2, a3;
          asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(a0), "=r"(a1) : "r"(frag + (8 * s) * kRtRow * 4) : "memory");
          asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(a2), "=r"(a3) : "r"(frag + (8 * s + 1) * kRtRow * 4) : "memory");
          if (s & 1) {
            rt_mma(c0b, a0, a1, a2, a3, a0, a2);
            rt_mma(c1b, a0, a1, a2, a3, a1, a3);
          } else {
            rt_mma(c0a, a0, a1, a2, a3, a0, a2);
            rt_mma(c1a, a0, a1, a2, a3, a1, a3);
          }
        }
        __syncwarp();
      };
      {
        float mu_b = 0.f, mv_b = 0.f, Xb = 0.f, Yb = 0.f, Zb = 0.f;
        for (int64_t base = o0; base < o1; base += 64) {
          const int64_t o = base + lane;
          fetch(min(o, o1 - 1), ip_nn, mu_b, mv_b, Xb, Yb, Zb);
          step(o, mu_n, mv_n, Xn, Yn, Zn);
          if (base + 32 >= o1) break;
          fetch(min(o + 32, o1 - 1), ip_nn, mu_n, mv_n, Xn, Yn, Zn);
          step(o + 32, mu_b, mv_b, Xb, Yb, Zb);
        }
      }
      o0 = o1;
      o1 = o2;
      if (c + 1 < C) prime_ids(o0, o1);
      // ---- per-segment epilogue
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float val = (i >> 2) == 0 ? c0a[i & 3] + c0b[i & 3] : c1a[i & 3] + c1b[i & 3];
        if (i < 2) vsum[i] += val;
        if (kind[i] == 2) blk[ofs[i] + 6 * c] = (double)val;
        if (kind[i] == 3) s_acc[warp][c][ofs[i]] += (double)val;  // one owner lane per element
      }
      {
        double tmp[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) tmp[i] = i < 6 ? gc[i] : 0.0;
        double r = group_reduce<8, double>(tmp, lane & 7);
        r += __shfl_xor_sync(0xffffffffu, r, 8);
        r += __shfl_xor_sync(0xffffffffu, r, 16);
        if (lane < 6) s_acc[warp][c][21 + lane] += r;
      }
    }
    // ---- per-frame epilogue: V_f (upper triangle) and g_f
#pragma unroll
    for (int i = 0; i < 2; ++i)
      if (kind[i] == 1) blk[ofs[i]] = (double)vsum[i];
    {
      double tmp[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) tmp[i] = i < 6 ? gf[i] : 0.0;
      double r = group_reduce<8, double>(tmp, lane & 7);
      r += __shfl_xor_sync(0xffffffffu, r, 8);
      r += __shfl_xor_sync(0xffffffffu, r, 16);
      if (lane < 6) blk[21 + lane] = r;
    }
  }
  {
    double tmp[8] = {rho_sum, sumsq, gTd, dDd, nobs, 0.0, 0.0, 0.0};
    double r = group_reduce<8, double>(tmp, lane & 7);
    r += __shfl_xor_sync(0xffffffffu, r, 8);
    r += __shfl_xor_sync(0xffffffffu, r, 16);
    if (lane < 8) s_sc[warp][lane] = lane < 5 ? r : 0.0;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < kRigPart; c += blockDim.x) {
    double s = 0.0;
    if (c < 27 * kMaxCams) {
      const int cam = c / 27, e = c - 27 * cam;
      if (cam < C)
        for (int p = 0; p < kRtWarps; ++p) s += s_acc[p][cam][e];
    } else if (c < 27 * kMaxCams + 5) {
      for (int p = 0; p < kRtWarps; ++p) s += s_sc[p][c - 27 * kMaxCams];
    }
    d.sweep_part[(size_t)blockIdx.x * kRigPart + c] = s;
  }
}

// --------------------------------------------------------------------------
// frame elimination, one warp per frame; a lane owns the columns lane, lane + 32, ... of [W | g_f]
// (6C + 1 columns).  The frame's columns are staged in shared memory so that every lane can read
// column i for the products (W'Y)[i][j]; the warp's share of the reduced system accumulates in
// shared memory (dynamic: warps x ((6C + 1) 6C + 2) doubles - 7.4 KB per warp at 5 cameras, 74.5 KB
// at 16), in a fixed order, so the sums are bit-reproducible.
// --------------------------------------------------------------------------
__global__ void k_rig_schur(RigDev d, LmParams P) {
  extern __shared__ double s_dyn[];
  __shared__ int s_meta[2];
  __shared__ double s_lambda;
  const WorkItem it = d.items[blockIdx.x];
  RigState& ps = d.ps[0];
  const int C = d.n_cams, n6 = 6 * C, ncol = n6 + 1, part = d.schur_stride;
  const int nwarps = blockDim.x >> 5;
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
  double* s_acc = s_dyn + (size_t)warp * part;                           // this warp's partial
  double* s_col = s_dyn + (size_t)nwarps * part + (size_t)warp * 6 * ncol;  // the frame's [W | g_f], row a at a * ncol
  for (int i = lane; i < part; i += 32) s_acc[i] = 0.0;
  double gDg = 0.0, fails = 0.0;
  for (int64_t f = it.v0 + warp; f < it.v1; f += nwarps) {
    const double* blk = d.blk[par] + (size_t)f * d.blk_stride;
    double A[6][6], Dv[6], gv[6];
    {
      int idx = 0;
#pragma unroll
      for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int b = a; b < 6; ++b) {
          A[b][a] = blk[idx];
          ++idx;
        }
    }
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      gv[a] = blk[21 + a];
      Dv[a] = clampd(A[a][a], P.min_diagonal, P.max_diagonal);
      if (lane == 0 && A[a][a] > 0.0) gDg += gv[a] * gv[a] / A[a][a];
      A[a][a] += lambda * Dv[a];
    }
    const bool ok = chol6(A);
    if (!ok && lane == 0) fails += 1.0;
    __syncwarp();  // the previous frame's columns are no longer being read
    double y[kRigSchurCols][6];
#pragma unroll
    for (int cc = 0; cc < kRigSchurCols; ++cc) {
      const int j = lane + 32 * cc;
      double col[6];
#pragma unroll
      for (int a = 0; a < 6; ++a) {
        double w = 0.0;
        if (j < n6) w = blk[27 + a * n6 + j];
        else if (j == n6) w = gv[a];
        col[a] = w;
        y[cc][a] = w;
        if (j < ncol) s_col[a * ncol + j] = w;
      }
      if (j < ncol) chol6_solve(A, y[cc]);
    }
    double* yz = d.yz + (size_t)f * d.yz_stride;
#pragma unroll
    for (int cc = 0; cc < kRigSchurCols; ++cc) {
      const int j = lane + 32 * cc;
      if (j < n6) {
#pragma unroll
        for (int a = 0; a < 6; ++a) yz[a * n6 + j] = y[cc][a];
      } else if (j == n6) {
#pragma unroll
        for (int a = 0; a < 6; ++a) {
          yz[6 * n6 + a] = y[cc][a];
          yz[6 * n6 + 6 + a] = gv[a];
          yz[6 * n6 + 12 + a] = Dv[a];
        }
      }
    }
    __syncwarp();
    // (W'Y)[i][j] = sum_a W[a][i] Y[a][j]: this lane's columns j, every row i
#pragma unroll
    for (int cc = 0; cc < kRigSchurCols; ++cc) {
      const int j = lane + 32 * cc;
      if (j < ncol) {
        for (int i = 0; i < n6; ++i) {
          double sum = 0.0;
#pragma unroll
          for (int a = 0; a < 6; ++a) sum = fma(s_col[a * ncol + i], y[cc][a], sum);
          s_acc[j * n6 + i] += sum;
        }
      }
    }
  }
  if (lane == 0) {
    s_acc[part - 2] = gDg;
    s_acc[part - 1] = fails;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < part; c += blockDim.x) {
    double sum = 0.0;
    for (int w = 0; w < nwarps; ++w) sum += s_dyn[(size_t)w * part + c];
    d.schur_part[(size_t)blockIdx.x * part + c] = sum;
  }
}

// fixed-order sum of the item partials: block b owns components [32b, 32b+32), thread
// (c, seg) sums items seg, seg+32, ... (independent loads, all in flight at once) and the 32 segment
// sums are combined in order
constexpr int kRigRedSeg = 32;
__global__ void __launch_bounds__(32 * kRigRedSeg) k_rig_reduce(const double* __restrict__ part, int stride, int n_items,
                                                                double* __restrict__ red) {
  __shared__ double s_buf[kRigRedSeg][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), seg = threadIdx.x >> 5;
  double s = 0.0;
  if (c < stride) {
#pragma unroll 8
    for (int i = seg; i < n_items; i += kRigRedSeg) s += part[(size_t)i * stride + c];
  }
  s_buf[seg][threadIdx.x & 31] = s;
  __syncthreads();
  if (seg == 0 && c < stride) {
    double tot = 0.0;
#pragma unroll
    for (int g = 0; g < kRigRedSeg; ++g) tot += s_buf[g][threadIdx.x];
    red[c] = tot;
  }
}

// The LM state lives in global memory between launches; a single-CTA kernel stages it in shared memory, lets its
// leader thread run the state machine there, and writes it back - a lone thread working on global memory pays an L2
// round trip per dependent access (measured: 38 us per reduced solve of 18 unknowns, nearly all of it in those accesses).
static_assert(sizeof(RigState) % sizeof(double) == 0, "RigState is staged as doubles");
__device__ __forceinline__ void rig_state_in(RigState& sh, const RigState* g) {
  for (int i = threadIdx.x; i < (int)(sizeof(RigState) / sizeof(double)); i += blockDim.x)
    reinterpret_cast<double*>(&sh)[i] = reinterpret_cast<const double*>(g)[i];
  __syncthreads();
}
__device__ __forceinline__ void rig_state_out(RigState* g, const RigState& sh) {
  __syncthreads();
  for (int i = threadIdx.x; i < (int)(sizeof(RigState) / sizeof(double)); i += blockDim.x)
    reinterpret_cast<double*>(g)[i] = reinterpret_cast<const double*>(&sh)[i];
}

// FUSED: the kernel also reduces the per-item sweep partials itself (one launch instead of k_rig_reduce + k_rig_decide;
// a problem that is not partitioned over GPUs needs nothing in between).  1024 threads: thread (c, seg) sums items
// seg, seg + 8, ... of component c, the 8 segment sums are combined in order - only the 27 C + 5 components in use.
// peer != 0 (frames partitioned over GPUs, FUSED only): the reduced partial is all-reduced over peer memory inside this
// launch (peer_allreduce_n, common.cuh) - every rank holds the same LM state, so all of them are here - instead of
// k_rig_reduce + ncclAllReduce + k_rig_decide<false>.
template <bool FUSED>
__global__ void __launch_bounds__(FUSED ? 1024 : 256) k_rig_decide(RigDev d, LmParams P, int n_items, PeerComm pc, int peer) {
  __shared__ double s_buf[kRigPart];
  __shared__ double s_seg[8][128];
  __shared__ RigState ps;
  __shared__ int s_comm_ok;
  if (d.ps[0].lm.status != 0) return;
  if (threadIdx.x == 0) s_comm_ok = 1;
  rig_state_in(ps, d.ps);
  const bool evaluated = ps.lm.solve_ok != 0;
  if (evaluated && FUSED) {
    const int C = d.n_cams, ncomp = 27 * C + 5;
    for (int c = threadIdx.x; c < kRigPart; c += blockDim.x) s_buf[c] = 0.0;
    const int lc = threadIdx.x & 127, seg = threadIdx.x >> 7;
    for (int c0 = 0; c0 < ncomp; c0 += 128) {
      const int idx = c0 + lc;
      const int comp = idx < 27 * C ? idx : 27 * kMaxCams + (idx - 27 * C);
      double sum = 0.0;
      if (idx < ncomp) {
#pragma unroll 8
        for (int i = seg; i < n_items; i += 8) sum += d.sweep_part[(size_t)i * kRigPart + comp];
      }
      s_seg[seg][lc] = sum;
      __syncthreads();
      if (seg == 0 && idx < ncomp) {
        double tot = 0.0;
#pragma unroll
        for (int g = 0; g < 8; ++g) tot += s_seg[g][lc];
        s_buf[comp] = tot;
      }
      __syncthreads();
    }
    if (peer) {
      const bool ok = peer_allreduce_n(s_buf, kRigPart, pc);
      if (threadIdx.x == 0) s_comm_ok = ok ? 1 : 0;
    }
  } else if (evaluated) {
    for (int c = threadIdx.x; c < kRigPart; c += blockDim.x) s_buf[c] = d.red_sweep[c];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    LmState& s = ps.lm;
    if (!evaluated) {
      lm_reject(s, P, d.n_done);
      s.solve_ok = 1;
      if (s.status == 0 && s.iterations >= P.max_iterations) lm_finish(s, 4, d.n_done, true);
    } else if (!s_comm_ok) {  // the peer-memory allreduce timed out
      lm_finish(s, CCG_ERR_NCCL, d.n_done, false);
    } else {
      const double* r = s_buf;
      const int base = 27 * kMaxCams;
      const bool accept = lm_decide(s, P, 0.5 * r[base], r[base + 2] + s.gTd_u, r[base + 3] + s.dDd_u, d.n_done);
      if (accept) {
        for (int c = 0; c < ps.n_cams; ++c) {
          for (int e = 0; e < 21; ++e) ps.Ucc[c][e] = r[27 * c + e];
          for (int e = 0; e < 6; ++e) ps.gc[c][e] = r[27 * c + 21 + e];
        }
        s.sumsq = r[base + 1];
        s.n_obs = r[base + 4];
      }
    }
  }
  rig_state_out(d.ps, ps);
}

// the reduced step -> candidate camera extrinsics and the model terms g'd, d'Dd (one thread)
__device__ __forceinline__ void rig_apply_step(RigState& ps, LmState& s, const double* x, const double* Du, int C, int par) {
  double gTd = 0.0, dDd = 0.0;
  for (int c = 0; c < C; ++c) {
    double dc[6];
    for (int i = 0; i < 6; ++i) {
      dc[i] = ps.frozen[c] ? 0.0 : x[6 * c + i];
      ps.du[6 * c + i] = dc[i];
      gTd += ps.gc[c][i] * dc[i];
      dDd += dc[i] * dc[i] * Du[6 * c + i];
    }
    double q[4] = {ps.cam[par][c][0], ps.cam[par][c][1], ps.cam[par][c][2], ps.cam[par][c][3]};
    double qn[4];
    quat_plus(q, dc[0], dc[1], dc[2], qn);
    for (int i = 0; i < 4; ++i) ps.cam[par ^ 1][c][i] = qn[i];
    for (int i = 0; i < 3; ++i) ps.cam[par ^ 1][c][4 + i] = ps.cam[par][c][4 + i] + dc[3 + i];
  }
  s.gTd_u = gTd;
  s.dDd_u = dDd;
  s.solve_ok = 1;
}

// S x = rhs for n <= 32 by ONE warp, warp barriers only.  Right-looking Cholesky on a copy T of S (row stride n | 1:
// a column is conflict-free): lane k owns COLUMN k of the trailing matrix and updates its rows k .. n-1 in batches of
// four (loads of a batch in flight together); then the substitutions with unknown i in a register of lane i and the
// pivots broadcast by shuffles.  Compact loops on purpose: a fully unrolled register-resident version (tried) is
// 250 KB of straight-line code and runs at instruction-fetch speed.  The block-cooperative factorisation below pays
// three block barriers per column and two per substituted unknown: 30 us at n = 24.  x is returned in rhs.
__device__ __forceinline__ bool rig_warp_solve(const double* S, double* rhs, double* T, int n, int lane) {
  const unsigned full = 0xffffffffu;
  const int ns = n | 1;
  if (lane < n)
    for (int k = 0; k < n; ++k) T[lane * ns + k] = S[lane * n + k];
  __syncwarp();
  double inv_ll = 1.0;  // 1 / L[lane][lane]
  for (int j = 0; j < n; ++j) {
    const double dd = T[j * ns + j];
    if (!(dd > 0.0) || !isfinite(dd)) return false;  // (uniform)
    const double inv = rsqrt(dd);
    if (lane == j) inv_ll = inv;
    __syncwarp();
    if (lane > j && lane < n) T[lane * ns + j] *= inv;
    __syncwarp();
    if (lane > j && lane < n) {
      const double lkj = T[lane * ns + j];
      int i = lane;
      for (; i + 3 < n; i += 4) {
        const double a0 = T[i * ns + j], a1 = T[(i + 1) * ns + j], a2 = T[(i + 2) * ns + j], a3 = T[(i + 3) * ns + j];
        const double c0 = T[i * ns + lane], c1 = T[(i + 1) * ns + lane], c2 = T[(i + 2) * ns + lane],
                     c3 = T[(i + 3) * ns + lane];
        T[i * ns + lane] = fma(-a0, lkj, c0);
        T[(i + 1) * ns + lane] = fma(-a1, lkj, c1);
        T[(i + 2) * ns + lane] = fma(-a2, lkj, c2);
        T[(i + 3) * ns + lane] = fma(-a3, lkj, c3);
      }
      for (; i < n; ++i) T[i * ns + lane] = fma(-T[i * ns + j], lkj, T[i * ns + lane]);
    }
    __syncwarp();
  }
  double b = lane < n ? rhs[lane] : 0.0;
#pragma unroll 4
  for (int j = 0; j < n; ++j) {  // L y = rhs
    const double lij = (lane > j && lane < n) ? T[lane * ns + j] : 0.0;
    const double yj = __shfl_sync(full, b * inv_ll, j);
    b = lane == j ? yj : fma(-lij, yj, b);
  }
#pragma unroll 4
  for (int j = n - 1; j >= 0; --j) {  // L' x = y
    const double lji = lane < j ? T[j * ns + lane] : 0.0;
    const double xj = __shfl_sync(full, b * inv_ll, j);
    b = lane == j ? xj : fma(-lji, xj, b);
  }
  if (lane < n) rhs[lane] = b;
  __syncwarp();
  return true;
}

// reduced (6C x 6C) solve, block-cooperative Cholesky and substitutions in (dynamic) shared memory:
// S n x n | rhs n | Du n doubles, n = 6C <= 96
__device__ __forceinline__ void rig_solve_body(const RigDev& d, const LmParams& P, RigState& ps, double* s_dyn, int* s_flag,
                                               double& s_piv) {
  LmState& s = ps.lm;
  const int C = ps.n_cams, n = 6 * C, par = s.par, part = d.schur_stride;
  double* S = s_dyn;
  double* rhs = S + n * n;
  double* Du = rhs + n;
  const double* r = d.red_schur;  // (W'Y | W'z): column j at j * n, then g'D^-1g and the failure count
  if (threadIdx.x == 0) {
    double gDg = r[part - 2];
    for (int c = 0; c < C; ++c) {
      if (ps.frozen[c]) continue;
      int idx = 0;
      for (int a = 0; a < 6; ++a) {
        const double dd = ps.Ucc[c][idx];
        if (dd > 0.0) gDg += ps.gc[c][a] * ps.gc[c][a] / dd;
        idx += 6 - a;
      }
    }
    const double rnorm = sqrt(2.0 * s.cost);
    s.grad_norm = rnorm > 0.0 ? sqrt(gDg) / rnorm : 0.0;
    int go = 1;
    if (s.finalize) {
      s.finalize = 0;
      go = 0;
    } else if (s.cost == 0.0 || sqrt(gDg) <= P.gradient_tolerance * rnorm) {
      lm_finish(s, 1, d.n_done, false);
      go = 0;
    } else {
      s.iterations++;
      if (r[part - 1] > 0.0) {
        s.solve_ok = 0;
        go = 0;
      }
    }
    s_flag[0] = go;
    s_flag[1] = 1;
  }
  __syncthreads();
  if (!s_flag[0]) return;
  const double lambda = s.lambda;
  // assemble S = U + lambda D - W'Y (upper from the partial, symmetrised), rhs = -g + W'z
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
    const int a = e / n, b = e - n * a;
    const int lo = a <= b ? a : b, hi = a <= b ? b : a;
    double val = -r[hi * n + lo];  // (W'Y)[lo][hi] lives in column hi, row lo
    const int ca = a / 6, cb = b / 6;
    if (ca == cb) {
      const int i = lo - 6 * ca, j = hi - 6 * ca;
      val += ps.Ucc[ca][i * 6 - i * (i - 1) / 2 + (j - i)];
    }
    S[a * n + b] = val;
  }
  __syncthreads();
  for (int a = threadIdx.x; a < n; a += blockDim.x) {
    const int c = a / 6, i = a - 6 * c;
    const double dd = ps.Ucc[c][i * 6 - i * (i - 1) / 2];
    Du[a] = clampd(dd, P.min_diagonal, P.max_diagonal);
    rhs[a] = -ps.gc[c][i] + r[n * n + a];
    S[a * n + a] += lambda * Du[a];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
    const int a = e / n, b = e - n * a;
    if (ps.frozen[a / 6] || ps.frozen[b / 6]) S[e] = (a == b) ? 1.0 : 0.0;
  }
  for (int a = threadIdx.x; a < n; a += blockDim.x)
    if (ps.frozen[a / 6]) rhs[a] = 0.0;
  __syncthreads();
  if (n <= 32) {
    // up to 5 cameras: ONE warp (see rig_warp_solve)
    if (threadIdx.x >= 32) return;
    double* T = Du + n;  // n x (n | 1) scratch behind S | rhs | Du: the transpose access of the back substitution
    const bool ok = rig_warp_solve(S, rhs, T, n, threadIdx.x);
    if (threadIdx.x == 0) {
      if (ok) rig_apply_step(ps, s, rhs, Du, C, par);
      else s.solve_ok = 0;
    }
    return;
  }
  // right-looking Cholesky, lower triangle, all threads
  for (int j = 0; j < n; ++j) {
    if (threadIdx.x == 0) {
      const double dd = S[j * n + j];
      if (!(dd > 0.0) || !isfinite(dd)) s_flag[1] = 0;
      S[j * n + j] = sqrt(dd);
    }
    __syncthreads();
    if (!s_flag[1]) break;
    const double inv = 1.0 / S[j * n + j];
    for (int i = j + 1 + threadIdx.x; i < n; i += blockDim.x) S[i * n + j] *= inv;
    __syncthreads();
    const int m = n - j - 1;
    for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
      const int i = j + 1 + e / m, k = j + 1 + e % m;
      if (k <= i) S[i * n + k] -= S[i * n + j] * S[k * n + j];
    }
    __syncthreads();
  }
  if (!s_flag[1]) {
    if (threadIdx.x == 0) s.solve_ok = 0;
    return;
  }
  // forward L y = rhs and backward L' x = y, column-oriented: thread 0 finishes unknown i, all threads update the rest
  for (int i = 0; i < n; ++i) {
    if (threadIdx.x == 0) {
      s_piv = rhs[i] / S[i * n + i];
      rhs[i] = s_piv;
    }
    __syncthreads();
    const double xi = s_piv;
    for (int k = i + 1 + threadIdx.x; k < n; k += blockDim.x) rhs[k] -= S[k * n + i] * xi;
    __syncthreads();
  }
  for (int i = n - 1; i >= 0; --i) {
    if (threadIdx.x == 0) {
      s_piv = rhs[i] / S[i * n + i];
      rhs[i] = s_piv;
    }
    __syncthreads();
    const double xi = s_piv;
    for (int k = threadIdx.x; k < i; k += blockDim.x) rhs[k] -= S[i * n + k] * xi;
    __syncthreads();
  }
  if (threadIdx.x != 0) return;
  rig_apply_step(ps, s, rhs, Du, C, par);
}

// peer != 0: the rank's reduced system in d.red_schur is first all-reduced over peer memory, in place.
__global__ void __launch_bounds__(256) k_rig_solve(RigDev d, LmParams P, PeerComm pc, int peer) {
  extern __shared__ double s_dyn[];
  __shared__ int s_flag[2];
  __shared__ double s_piv;
  __shared__ RigState ps;
  if (!(d.ps[0].lm.status == 0 || d.ps[0].lm.finalize)) return;
  rig_state_in(ps, d.ps);
  if (peer && !peer_allreduce_n(d.red_schur, d.schur_stride, pc)) {  // (uniform over the CTA)
    if (threadIdx.x == 0) {
      ps.lm.finalize = 0;
      if (ps.lm.status == 0) lm_finish(ps.lm, CCG_ERR_NCCL, d.n_done, false);
    }
    rig_state_out(d.ps, ps);
    return;
  }
  rig_solve_body(d, P, ps, s_dyn, s_flag, s_piv);  // (threads leave the body at different points; they meet again here)
  rig_state_out(d.ps, ps);
}

__global__ void k_rig_init_state(RigDev d, LmParams P, const double* cam0, const int* frozen, int n_items,
                                 double meas_scale) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  RigState& ps = d.ps[0];
  lm_init(ps.lm, P);
  ps.lm.meas_scale = meas_scale;
  ps.n_cams = d.n_cams;
  ps.item0 = 0;
  ps.item1 = n_items;
  for (int c = 0; c < kMaxCams; ++c) {
    for (int k = 0; k < 8; ++k) {
      const double v = c < d.n_cams ? cam0[c * 8 + k] : 0.0;
      ps.cam[0][c][k] = v;
      ps.cam[1][c][k] = v;
    }
    ps.frozen[c] = c < d.n_cams ? frozen[c] : 1;
    for (int e = 0; e < 21; ++e) ps.Ucc[c][e] = 0.0;
    for (int e = 0; e < 6; ++e) ps.gc[c][e] = 0.0;
  }
  for (int k = 0; k < kMaxN6; ++k) ps.du[k] = 0.0;
  *d.n_done = 0;
}

}  // namespace ccg

// ==========================================================================
// host side
// ==========================================================================
using namespace ccg;

struct ccg_rig_problem {
  ccg_ctx* ctx = nullptr;
  int64_t n_frames = 0, n_obs = 0, n_points = 0;
  int n_cams = 0, n_items = 0, shared = 0;
  RigDev dev;
  std::vector<void*> allocs;
  double *d_cam0 = nullptr, *d_pose0 = nullptr;
  int* d_frozen = nullptr;
  std::vector<int64_t> perm;  // sorted position -> caller's observation index
  double meas_scale = 0.0;
  double huber_a = (double)(3.0f / 500.0f);  // of the last run (final per-observation cost pass)
  cudaGraphExec_t graph_exec = nullptr;      // one LM iteration, see rig_run
  LmParams graph_params;
  int graph_kernels = 0, graph_allreduces = 0;
  bool one_shot = false;                     // created by ccg_rig_solve: runs once, no graph instantiation
  int schur_warps = 4;                       // elimination kernel: warps per CTA and dynamic shared memory
  size_t schur_smem = 0, solve_smem = 0;     // (sized by the camera count, rig_build)
};

namespace {

#define CCG_TRY(x)                   \
  do {                               \
    const int rc__ = (x);            \
    if (rc__ != CCG_OK) return rc__; \
  } while (0)

template <typename T>
int rig_alloc(ccg_rig_problem* p, T** out, size_t n) {
  void* ptr = nullptr;
  cudaError_t e = pool_alloc(p->ctx, &ptr, std::max<size_t>(n, 1) * sizeof(T));
  if (e != cudaSuccess) return set_error(p->ctx, CCG_ERR_CUDA, "device allocation", e);
  p->allocs.push_back(ptr);
  *out = (T*)ptr;
  return CCG_OK;
}

void rig_free(ccg_rig_problem* p) {
  if (!p) return;
  if (p->ctx) cudaSetDevice(p->ctx->device);
  if (p->graph_exec) cudaGraphExecDestroy(p->graph_exec);
  for (void* a : p->allocs) pool_free(p->ctx, a);
  delete p;
}

int rig_build(ccg_ctx* ctx, const float* u, const float* v, const int32_t* obs_cam, const int64_t* obs_point,
              const int64_t* frame_offsets, int64_t n_frames, const float* Xw, int64_t n_points,
              int32_t n_cams, const double* q_c, const double* t_c, const uint8_t* cam_frozen,
              const double* q_f, const double* t_f, int shared, ccg_rig_problem* p) {
  p->ctx = ctx;
  p->n_frames = n_frames;
  p->n_cams = n_cams;
  p->n_points = n_points;
  p->shared = shared;
  const int C = n_cams;
  const int64_t o0 = frame_offsets[0], n_obs = frame_offsets[n_frames] - o0;
  p->n_obs = n_obs;
  // ---- stable sort of every frame's observations by camera -> contiguous (frame, camera) segments
  std::vector<int64_t> seg((size_t)n_frames * C + 1, 0);
  p->perm.resize((size_t)n_obs);
  std::vector<float> su((size_t)n_obs), sv((size_t)n_obs);
  std::vector<int> spt((size_t)n_obs);
  float mscale = 0.0f;
  int64_t w = 0;
  for (int64_t f = 0; f < n_frames; ++f) {
    for (int c = 0; c < C; ++c) {
      seg[(size_t)f * C + c] = w;
      for (int64_t o = frame_offsets[f]; o < frame_offsets[f + 1]; ++o) {
        if (obs_cam[o] != c) continue;
        p->perm[(size_t)w] = o - o0;
        su[(size_t)w] = u[o];
        sv[(size_t)w] = v[o];
        spt[(size_t)w] = (int)obs_point[o];
        mscale = std::max(mscale, std::max(std::fabs(u[o]), std::fabs(v[o])));
        ++w;
      }
    }
  }
  seg[(size_t)n_frames * C] = w;
  if (w != n_obs) return set_error(ctx, CCG_ERR_BAD_ARG, "camera id out of range");
  p->meas_scale = (double)mscale;

  // ---- work items over frames
  const int64_t target = (int64_t)ctx->sm_count * 4;
  const int64_t fpi = std::max<int64_t>(kRigWarps / 2, (n_frames + target - 1) / std::max<int64_t>(target, 1));
  std::vector<WorkItem> items;
  const int64_t n_it = std::max<int64_t>(1, (n_frames + fpi - 1) / fpi);
  for (int64_t i = 0; i < n_it; ++i) {
    WorkItem it;
    it.problem = 0;
    it.v0 = (int)(n_frames * i / n_it);
    it.v1 = (int)(n_frames * (i + 1) / n_it);
    it.pad = 0;
    items.push_back(it);
  }
  p->n_items = (int)items.size();

  RigDev& d = p->dev;
  std::memset(&d, 0, sizeof(d));
  d.n_cams = C;
  d.schur_stride = rig_schur_part(6 * C);
  d.blk_stride = ((27 + 36 * C) + 3) / 4 * 4;
  d.yz_stride = ((36 * C + 18) + 3) / 4 * 4;
  float *du_ = nullptr, *dv_ = nullptr, *dX = nullptr;
  int* dpt = nullptr;
  int64_t* dseg = nullptr;
  WorkItem* ditems = nullptr;
  CCG_TRY(rig_alloc(p, &du_, (size_t)n_obs));
  CCG_TRY(rig_alloc(p, &dv_, (size_t)n_obs));
  CCG_TRY(rig_alloc(p, &dpt, (size_t)n_obs));
  CCG_TRY(rig_alloc(p, &dX, (size_t)n_points * 3));
  CCG_TRY(rig_alloc(p, &dseg, seg.size()));
  CCG_TRY(rig_alloc(p, &ditems, items.size()));
  CCG_TRY(rig_alloc(p, &d.pose[0], (size_t)n_frames * 8));
  CCG_TRY(rig_alloc(p, &d.pose[1], (size_t)n_frames * 8));
  CCG_TRY(rig_alloc(p, &p->d_pose0, (size_t)n_frames * 8));
  CCG_TRY(rig_alloc(p, &d.blk[0], (size_t)n_frames * d.blk_stride));
  CCG_TRY(rig_alloc(p, &d.blk[1], (size_t)n_frames * d.blk_stride));
  CCG_TRY(rig_alloc(p, &d.yz, (size_t)n_frames * d.yz_stride));
  CCG_TRY(rig_alloc(p, &d.sweep_part, items.size() * kRigPart));
  CCG_TRY(rig_alloc(p, &d.schur_part, items.size() * (size_t)d.schur_stride));
  CCG_TRY(rig_alloc(p, &d.red_sweep, (size_t)kRigPart));
  CCG_TRY(rig_alloc(p, &d.red_schur, (size_t)d.schur_stride));
  CCG_TRY(rig_alloc(p, &d.ps, (size_t)1));
  CCG_TRY(rig_alloc(p, &d.n_done, (size_t)1));
  CCG_TRY(rig_alloc(p, &p->d_cam0, (size_t)kMaxCams * 8));
  CCG_TRY(rig_alloc(p, &p->d_frozen, (size_t)kMaxCams));
  d.u = du_;
  d.v = dv_;
  d.pt = dpt;
  d.Xw = dX;
  d.seg_off = dseg;
  d.items = ditems;
  cudaStream_t st = ctx->stream;
  std::vector<double> pose((size_t)n_frames * 8, 0.0), cam((size_t)kMaxCams * 8, 0.0);
  for (int64_t f = 0; f < n_frames; ++f) {
    for (int k = 0; k < 4; ++k) pose[(size_t)f * 8 + k] = q_f[f * 4 + k];
    for (int k = 0; k < 3; ++k) pose[(size_t)f * 8 + 4 + k] = t_f[f * 3 + k];
  }
  std::vector<int> frozen(kMaxCams, 1);
  for (int c = 0; c < C; ++c) {
    for (int k = 0; k < 4; ++k) cam[(size_t)c * 8 + k] = q_c[c * 4 + k];
    for (int k = 0; k < 3; ++k) cam[(size_t)c * 8 + 4 + k] = t_c[c * 3 + k];
    frozen[c] = cam_frozen && cam_frozen[c] ? 1 : 0;
  }
  CCG_CUDA(ctx, cudaMemcpyAsync(du_, su.data(), sizeof(float) * n_obs, cudaMemcpyHostToDevice, st));
  CCG_CUDA(ctx, cudaMemcpyAsync(dv_, sv.data(), sizeof(float) * n_obs, cudaMemcpyHostToDevice, st));
  CCG_CUDA(ctx, cudaMemcpyAsync(dpt, spt.data(), sizeof(int) * n_obs, cudaMemcpyHostToDevice, st));
  CCG_CUDA(ctx, cudaMemcpyAsync(dX, Xw, sizeof(float) * 3 * n_points, cudaMemcpyHostToDevice, st));
  CCG_CUDA(ctx, cudaMemcpyAsync(dseg, seg.data(), sizeof(int64_t) * seg.size(), cudaMemcpyHostToDevice, st));
  CCG_CUDA(ctx, cudaMemcpyAsync(ditems, items.data(), sizeof(WorkItem) * items.size(), cudaMemcpyHostToDevice, st));
  CCG_CUDA(ctx, cudaMemcpyAsync(p->d_pose0, pose.data(), sizeof(double) * pose.size(), cudaMemcpyHostToDevice, st));
  CCG_CUDA(ctx, cudaMemcpyAsync(p->d_cam0, cam.data(), sizeof(double) * cam.size(), cudaMemcpyHostToDevice, st));
  CCG_CUDA(ctx, cudaMemcpyAsync(p->d_frozen, frozen.data(), sizeof(int) * kMaxCams, cudaMemcpyHostToDevice, st));
  CCG_CUDA(ctx, cudaStreamSynchronize(st));
  // a partitioned rig: the scale feeds the accept/reject thresholds of lm_decide - every rank needs the global maximum
  if (shared && ctx->world > 1) {
    double* d_ms = nullptr;
    CCG_TRY(rig_alloc(p, &d_ms, (size_t)1));
    CCG_CUDA(ctx, cudaMemcpyAsync(d_ms, &p->meas_scale, sizeof(double), cudaMemcpyHostToDevice, st));
    CCG_TRY(nccl_allreduce_max(ctx, d_ms, 1));
    CCG_CUDA(ctx, cudaMemcpyAsync(&p->meas_scale, d_ms, sizeof(double), cudaMemcpyDeviceToHost, st));
    CCG_CUDA(ctx, cudaStreamSynchronize(st));
  }
  // dynamic shared memory of the elimination and reduced-solve kernels (opt-in above 48 KB, per device)
  {
    const size_t part = (size_t)d.schur_stride, cols = (size_t)6 * (6 * C + 1);
    p->schur_warps = 4;
    while (p->schur_warps > 1 && sizeof(double) * p->schur_warps * (part + cols) > 200 * 1024) p->schur_warps /= 2;
    p->schur_smem = sizeof(double) * p->schur_warps * (part + cols);
    p->solve_smem = sizeof(double) * ((size_t)36 * C * C + 12 * C + (6 * C <= 32 ? (size_t)6 * C * ((6 * C) | 1) : 0));
    CCG_CUDA(ctx, cudaFuncSetAttribute(k_rig_schur, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->schur_smem));
    CCG_CUDA(ctx, cudaFuncSetAttribute(k_rig_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->solve_smem));
  }
  return CCG_OK;
}

int rig_reset(ccg_rig_problem* p, const LmParams& P) {
  ccg_ctx* ctx = p->ctx;
  RigDev& d = p->dev;
  CCG_CUDA(ctx, cudaMemcpyAsync(d.pose[0], p->d_pose0, sizeof(double) * p->n_frames * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  CCG_CUDA(ctx, cudaMemsetAsync(d.yz, 0, sizeof(double) * std::max<int64_t>(p->n_frames, 1) * d.yz_stride, ctx->stream));
  k_rig_init_state<<<1, 32, 0, ctx->stream>>>(d, P, p->d_cam0, p->d_frozen, p->n_items, p->meas_scale);
  ctx->launches++;
  CCG_CUDA(ctx, cudaGetLastError());
  return CCG_OK;
}

enum { TAG_SWEEP = 0, TAG_SCHUR = 1, TAG_SMALL = 2 };

int rig_sweep_decide(ccg_rig_problem* p, const LmParams& P, Profiler& prof, double* res_out, int* nar) {
  ccg_ctx* ctx = p->ctx;
  RigDev& d = p->dev;
  cudaStream_t st = ctx->stream;
  prof.begin(TAG_SWEEP);
  if (P.hessian_fp64 == 0 && !res_out) k_rig_sweep_tc<<<p->n_items, kRtWarps * 32, 0, st>>>(d, P.huber_a);
  else k_rig_sweep<<<p->n_items, kRigWarps * 32, 0, st>>>(d, P.huber_a, res_out, nullptr);
  prof.end();
  ctx->launches++;
  prof.begin(TAG_SMALL);
  const PeerComm pc = peer_comm(ctx);
  if (p->shared && ctx->world > 1 && !ctx->peer_ok) {
    k_rig_reduce<<<(kRigPart + 31) / 32, 32 * kRigRedSeg, 0, st>>>(d.sweep_part, kRigPart, p->n_items, d.red_sweep);
    CCG_TRY(nccl_allreduce_sum(ctx, d.red_sweep, kRigPart));
    (*nar)++;
    k_rig_decide<false><<<1, 256, 0, st>>>(d, P, p->n_items, pc, 0);
    ctx->launches += 2;
  } else {
    // frames partitioned over GPUs: the allreduce runs over peer memory inside the launch
    const int peer = (p->shared && ctx->world > 1) ? 1 : 0;
    k_rig_decide<true><<<1, 1024, 0, st>>>(d, P, p->n_items, pc, peer);
    ctx->launches++;
    *nar += peer;
  }
  prof.end();
  return CCG_OK;
}

int rig_schur_solve(ccg_rig_problem* p, const LmParams& P, Profiler& prof, int* nar) {
  ccg_ctx* ctx = p->ctx;
  RigDev& d = p->dev;
  cudaStream_t st = ctx->stream;
  prof.begin(TAG_SCHUR);
  k_rig_schur<<<p->n_items, p->schur_warps * 32, p->schur_smem, st>>>(d, P);
  prof.end();
  ctx->launches++;
  prof.begin(TAG_SMALL);
  k_rig_reduce<<<(d.schur_stride + 31) / 32, 32 * kRigRedSeg, 0, st>>>(d.schur_part, d.schur_stride, p->n_items, d.red_schur);
  int peer = 0;
  if (p->shared && ctx->world > 1) {
    if (ctx->peer_ok) peer = 1;  // all-reduced over peer memory inside k_rig_solve
    else CCG_TRY(nccl_allreduce_sum(ctx, d.red_schur, (size_t)d.schur_stride));
    (*nar)++;
  }
  k_rig_solve<<<1, 256, p->solve_smem, st>>>(d, P, peer_comm(ctx), peer);
  ctx->launches += 2;
  prof.end();
  return CCG_OK;
}

int rig_run(ccg_rig_problem* p, const ccg_opts& opts, ccg_timing* timing) {
  ccg_ctx* ctx = p->ctx;
  CCG_CUDA(ctx, cudaSetDevice(ctx->device));
  const LmParams P = to_params(opts);
  Profiler prof;
  prof.ctx = ctx;
  prof.on = opts.profile != 0;
  Profiler tot;
  tot.ctx = ctx;
  const size_t e0 = tot.grab(), e1 = tot.grab();
  prof.used = tot.used;
  int nar = 0;
  cudaEventRecord(ctx->events[e0], ctx->stream);
  CCG_TRY(rig_reset(p, P));
  if (p->shared && ctx->world > 1 && ctx->peer_ok) {
    // line the ranks up once per solve with a collective that waits without a deadline (the peer-memory allreduces
    // inside the kernels then only see the skew of one LM iteration) and agree on their sequence number
    CCG_TRY(peer_resync(ctx, p->dev.red_sweep));
    nar++;
  }
  CCG_TRY(rig_sweep_decide(p, P, prof, nullptr, &nar));
  // one LM iteration (six launches, allreduces included) replayed as a CUDA graph when no
  // per-kernel events are requested and the problem object is reused (CCG_NO_GRAPH=1 disables)
  static const bool no_graph = getenv("CCG_NO_GRAPH") != nullptr;
  const bool use_graph = !prof.on && !no_graph && !p->one_shot;
  if (use_graph && (!p->graph_exec || std::memcmp(&p->graph_params, &P, sizeof(P)) != 0)) {
    if (p->graph_exec) cudaGraphExecDestroy(p->graph_exec);
    p->graph_exec = nullptr;
    cudaGraph_t g = nullptr;
    const int64_t l0 = ctx->launches;
    const int a0 = nar;
    CCG_CUDA(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed));
    int rc = rig_schur_solve(p, P, prof, &nar);
    if (rc == CCG_OK) rc = rig_sweep_decide(p, P, prof, nullptr, &nar);
    const cudaError_t e = cudaStreamEndCapture(ctx->stream, &g);
    p->graph_kernels = (int)(ctx->launches - l0);
    p->graph_allreduces = nar - a0;
    ctx->launches = l0;
    nar = a0;
    if (rc != CCG_OK) return rc;
    if (e != cudaSuccess) return set_error(ctx, CCG_ERR_CUDA, "cudaStreamEndCapture", e);
    CCG_CUDA(ctx, cudaGraphInstantiate(&p->graph_exec, g, 0));
    cudaGraphDestroy(g);
    p->graph_params = P;
  }
  int done = 0;
  for (int it = 0; it <= opts.max_iterations * 2 + 64; ++it) {
    if (use_graph && done < 1) {
      CCG_CUDA(ctx, cudaGraphLaunch(p->graph_exec, ctx->stream));
      ctx->launches += p->graph_kernels;
      nar += p->graph_allreduces;
    } else {
      CCG_TRY(rig_schur_solve(p, P, prof, &nar));
      if (done >= 1) break;
      CCG_TRY(rig_sweep_decide(p, P, prof, nullptr, &nar));
    }
    if ((it & 3) == 3) {
      CCG_CUDA(ctx, cudaMemcpyAsync(ctx->h_poll, p->dev.n_done, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
      CCG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      done = ctx->h_poll[0];
    }
  }
  cudaEventRecord(ctx->events[e1], ctx->stream);
  CCG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  CCG_CUDA(ctx, cudaGetLastError());
  if (timing) {
    std::memset(timing, 0, sizeof(*timing));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->events[e0], ctx->events[e1]);
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
    timing->allreduces = nar;
    timing->sweep_obs = p->n_obs;
    timing->sweep_views = p->n_frames;
  }
  return CCG_OK;
}

int rig_download(ccg_rig_problem* p, double* q_c, double* t_c, double* q_f, double* t_f, double* obs_cost,
                 ccg_report* report) {
  ccg_ctx* ctx = p->ctx;
  CCG_CUDA(ctx, cudaSetDevice(ctx->device));
  RigState st;
  CCG_CUDA(ctx, cudaMemcpy(&st, p->dev.ps, sizeof(RigState), cudaMemcpyDeviceToHost));
  const int par = st.lm.par;
  for (int c = 0; c < p->n_cams; ++c) {
    if (q_c) for (int k = 0; k < 4; ++k) q_c[c * 4 + k] = st.cam[par][c][k];
    if (t_c) for (int k = 0; k < 3; ++k) t_c[c * 3 + k] = st.cam[par][c][4 + k];
  }
  if (q_f || t_f) {
    std::vector<double> pose((size_t)p->n_frames * 8);
    CCG_CUDA(ctx, cudaMemcpy(pose.data(), p->dev.pose[par], sizeof(double) * pose.size(), cudaMemcpyDeviceToHost));
    for (int64_t f = 0; f < p->n_frames; ++f) {
      if (q_f) for (int k = 0; k < 4; ++k) q_f[f * 4 + k] = pose[(size_t)f * 8 + k];
      if (t_f) for (int k = 0; k < 3; ++k) t_f[f * 3 + k] = pose[(size_t)f * 8 + 4 + k];
    }
  }
  if (obs_cost && p->n_obs > 0) {
    // per-observation 0.5*rho at the returned point (extrinsics_calibrator.cpp:219-225)
    double* d_cost = nullptr;
    CCG_CUDA(ctx, pool_alloc(ctx, (void**)&d_cost, sizeof(double) * p->n_obs));
    k_rig_sweep<<<p->n_items, kRigWarps * 32, 0, ctx->stream>>>(p->dev, p->huber_a, nullptr, d_cost);
    ctx->launches++;
    std::vector<double> sorted((size_t)p->n_obs);
    cudaError_t e = cudaMemcpyAsync(sorted.data(), d_cost, sizeof(double) * p->n_obs, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    pool_free(ctx, d_cost);
    if (e != cudaSuccess) return set_error(ctx, CCG_ERR_CUDA, "obs cost", e);
    for (int64_t i = 0; i < p->n_obs; ++i) obs_cost[p->perm[(size_t)i]] = sorted[(size_t)i];
  }
  if (report) {
    report->iterations = st.lm.iterations;
    report->accepted = st.lm.accepted;
    report->rejected = st.lm.rejected;
    report->sweeps = st.lm.sweeps;
    report->termination = st.lm.status;
    report->pad_ = 0;
    report->initial_cost = st.lm.cost0;
    report->final_cost = st.lm.cost;
    report->gradient_norm = st.lm.grad_norm;
    report->rms = st.lm.n_obs > 0.0 ? std::sqrt(st.lm.sumsq / st.lm.n_obs) : 0.0;
  }
  return st.lm.status < 0 ? st.lm.status : CCG_OK;
}

bool rig_bad(const float* u, const float* v, const int32_t* obs_cam, const int64_t* obs_point,
             const int64_t* frame_offsets, int64_t n_frames, const float* Xw, int64_t n_points, int n_cams) {
  if (!frame_offsets || n_frames < 0 || n_cams < 1 || n_points < 0) return true;
  const int64_t n = frame_offsets[n_frames] - frame_offsets[0];
  if (n < 0 || (n > 0 && (!u || !v || !obs_cam || !obs_point || !Xw))) return true;
  for (int64_t f = 0; f < n_frames; ++f)
    if (frame_offsets[f + 1] < frame_offsets[f]) return true;
  for (int64_t o = frame_offsets[0]; o < frame_offsets[n_frames]; ++o)
    if (obs_cam[o] < 0 || obs_cam[o] >= n_cams || obs_point[o] < 0 || obs_point[o] >= n_points) return true;
  return false;
}

}  // namespace

extern "C" {

int ccg_rig_create(ccg_ctx* ctx, const float* u, const float* v, const int32_t* obs_cam,
                   const int64_t* obs_point, const int64_t* frame_offsets, int64_t n_frames,
                   const float* Xw, int64_t n_points, int32_t n_cams, const double* q_c,
                   const double* t_c, const uint8_t* cam_frozen, const double* q_f,
                   const double* t_f, int shared, ccg_rig_problem** out) {
  if (!ctx || !out) return CCG_ERR_BAD_ARG;
  *out = nullptr;
  if (rig_bad(u, v, obs_cam, obs_point, frame_offsets, n_frames, Xw, n_points, n_cams) || !q_c || !t_c ||
      (n_frames > 0 && (!q_f || !t_f)))
    return set_error(ctx, CCG_ERR_BAD_ARG, "ccg_rig_create: bad argument");
  if (n_cams > kMaxCams) return set_error(ctx, CCG_ERR_UNSUPPORTED, "at most 16 cameras per rig in this build");
  CCG_CUDA(ctx, cudaSetDevice(ctx->device));
  ccg_rig_problem* p = new ccg_rig_problem();
  const int rc = rig_build(ctx, u, v, obs_cam, obs_point, frame_offsets, n_frames, Xw, n_points, n_cams, q_c,
                           t_c, cam_frozen, q_f, t_f, shared, p);
  if (rc != CCG_OK) {
    rig_free(p);
    return rc;
  }
  *out = p;
  return CCG_OK;
}

int ccg_rig_reset(ccg_rig_problem* p) {
  if (!p) return CCG_ERR_BAD_ARG;
  ccg_opts o;
  ccg_default_opts(&o);
  CCG_CUDA(p->ctx, cudaSetDevice(p->ctx->device));
  return rig_reset(p, to_params(o));
}

int ccg_rig_run(ccg_rig_problem* p, const ccg_opts* opts, ccg_timing* timing) {
  if (!p) return CCG_ERR_BAD_ARG;
  ccg_opts o;
  if (opts) o = *opts; else { ccg_default_opts(&o); o.max_iterations = 1000; }
  p->huber_a = o.huber_a;
  return rig_run(p, o, timing);
}

int ccg_rig_download(ccg_rig_problem* p, double* q_c, double* t_c, double* q_f, double* t_f,
                     double* obs_cost, ccg_report* report) {
  if (!p) return CCG_ERR_BAD_ARG;
  return rig_download(p, q_c, t_c, q_f, t_f, obs_cost, report);
}

void ccg_rig_destroy(ccg_rig_problem* p) { rig_free(p); }

int ccg_rig_solve(ccg_ctx* ctx, const float* u, const float* v, const int32_t* obs_cam,
                  const int64_t* obs_point, const int64_t* frame_offsets, int64_t n_frames,
                  const float* Xw, int64_t n_points, int32_t n_cams, double* q_c, double* t_c,
                  const uint8_t* cam_frozen, double* q_f, double* t_f, double* obs_cost,
                  const ccg_opts* opts, ccg_report* report) {
  ccg_rig_problem* p = nullptr;
  int rc = ccg_rig_create(ctx, u, v, obs_cam, obs_point, frame_offsets, n_frames, Xw, n_points, n_cams, q_c,
                          t_c, cam_frozen, q_f, t_f, 0, &p);
  if (rc != CCG_OK) return rc;
  p->one_shot = true;
  rc = ccg_rig_run(p, opts, nullptr);
  if (rc == CCG_OK) rc = rig_download(p, q_c, t_c, q_f, t_f, obs_cost, report);
  rig_free(p);
  return rc;
}

int ccg_rig_eval(ccg_ctx* ctx, const float* u, const float* v, const int32_t* obs_cam,
                 const int64_t* obs_point, const int64_t* frame_offsets, int64_t n_frames,
                 const float* Xw, int64_t n_points, int32_t n_cams, const double* q_c,
                 const double* t_c, const double* q_f, const double* t_f, double huber_a,
                 double* res, double* Vb, double* Wb, double* gvb, double* U, double* gu,
                 double* cost) {
  ccg_rig_problem* p = nullptr;
  int rc = ccg_rig_create(ctx, u, v, obs_cam, obs_point, frame_offsets, n_frames, Xw, n_points, n_cams, q_c,
                          t_c, nullptr, q_f, t_f, 0, &p);
  if (rc != CCG_OK) return rc;
  auto body = [&]() -> int {
    ccg_opts o;
    ccg_default_opts(&o);
    o.huber_a = huber_a;
    o.hessian_fp64 = 1;  // the evaluation entry returns the exact (all-FP64) blocks
    const LmParams P = to_params(o);
    CCG_TRY(rig_reset(p, P));
    double* d_res = nullptr;
    if (res) CCG_TRY(rig_alloc(p, &d_res, (size_t)p->n_obs * 2));
    Profiler prof;
    prof.ctx = ctx;
    int nar = 0;
    CCG_TRY(rig_sweep_decide(p, P, prof, d_res, &nar));
    CCG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    CCG_CUDA(ctx, cudaGetLastError());
    const int C = n_cams, n6 = 6 * C, bs = p->dev.blk_stride;
    std::vector<double> blk((size_t)n_frames * bs);
    CCG_CUDA(ctx, cudaMemcpy(blk.data(), p->dev.blk[1], sizeof(double) * blk.size(), cudaMemcpyDeviceToHost));
    RigState st;
    CCG_CUDA(ctx, cudaMemcpy(&st, p->dev.ps, sizeof(RigState), cudaMemcpyDeviceToHost));
    if (res) {
      std::vector<double> sorted((size_t)p->n_obs * 2);
      CCG_CUDA(ctx, cudaMemcpy(sorted.data(), d_res, sizeof(double) * sorted.size(), cudaMemcpyDeviceToHost));
      for (int64_t i = 0; i < p->n_obs; ++i) {
        res[2 * p->perm[(size_t)i]] = sorted[(size_t)2 * i];
        res[2 * p->perm[(size_t)i] + 1] = sorted[(size_t)2 * i + 1];
      }
    }
    for (int64_t f = 0; f < n_frames; ++f) {
      const double* b = &blk[(size_t)f * bs];
      if (Vb) {
        int idx = 0;
        for (int a = 0; a < 6; ++a)
          for (int c = a; c < 6; ++c) {
            Vb[f * 36 + a * 6 + c] = b[idx];
            Vb[f * 36 + c * 6 + a] = b[idx];
            ++idx;
          }
      }
      if (gvb) for (int a = 0; a < 6; ++a) gvb[f * 6 + a] = b[21 + a];
      if (Wb) for (int e = 0; e < 6 * n6; ++e) Wb[f * 6 * n6 + e] = b[27 + e];
    }
    if (U) {
      std::fill(U, U + (size_t)n6 * n6, 0.0);
      for (int c = 0; c < C; ++c) {
        int idx = 0;
        for (int a = 0; a < 6; ++a)
          for (int b2 = a; b2 < 6; ++b2) {
            U[(size_t)(6 * c + a) * n6 + 6 * c + b2] = st.Ucc[c][idx];
            U[(size_t)(6 * c + b2) * n6 + 6 * c + a] = st.Ucc[c][idx];
            ++idx;
          }
      }
    }
    if (gu) for (int c = 0; c < C; ++c) for (int a = 0; a < 6; ++a) gu[6 * c + a] = st.gc[c][a];
    if (cost) *cost = st.lm.cost;
    return CCG_OK;
  };
  rc = body();
  rig_free(p);
  return rc;
}

}  // extern "C"
