// Warp-specialised single-camera sweep (included by intr.cu) - the default sweep.
//
// Per observation the sweep needs (a) the residual, the cost and the gradient g = J'r, which
// decide WHERE the LM iteration converges, and (b) the Gauss-Newton blocks V, W, U = J'J, which
// only decide HOW FAST (a perturbed J'J still has g = 0 as its fixed point and every step is
// validated against the FP64 cost).  (b) is 2/3 of the arithmetic and is a dense contraction:
// per view  [V W; W' U] = sum over its 2P residual rows of (row)'(row), a 16 x 16 x 2P product.
//   FP64 pipe  : forward model, residual, cost, and g_v / g_u by REVERSE mode (no explicit J)
//   FP32 pipe  : a second forward pass and the analytic 2 x 15 Jacobian rows
//   tensor pipe: J'J by mma.sync.m16n8k16 with FP16 operands and FP32 accumulation - one lane owns one
//                observation, writes its two Jacobian rows (FP16) to a warp-private shared-memory tile,
//                and the warp contracts 32 observations with 8 MMAs; the B fragments are the A
//                fragment's registers regrouped (the product is J'J), and two ldmatrix.x4.trans per
//                k-step deliver both groupings without register moves.
//
// Roles (one work item = a range of views per CTA; template parameter PAIRS = warp pairs per CTA):
//   * an SM runs 8 warp PAIRS, two per SM sub-partition, as one CTA of 8 or two CTAs of 4: an FP64-role
//     warp (the higher warp ids) and a Jacobian-role warp that walk the SAME contiguous range of views,
//     each on its own - they share nothing but the read-only inputs, so there is no synchronisation inside
//     the loops; once per view the Jacobian warp checks that it is not too far ahead of its partner
//     (kSwLead), so that the second reader of an observation block hits L2 (HBM traffic 1.19 x the 20 B
//     per observation).  setmaxnreg gives the FP64 warps kSwRegD registers and leaves the Jacobian
//     warps kSwRegJ.
//
// Observation stream: the observations live in HBM as 640-byte blocks of 32 observations
// (u[32] v[32] X[32] Y[32] Z[32], views padded to whole blocks with zeros), so a 32-observation
// chunk is ONE aligned block and the blocks of a warp's view range are one contiguous stream:
// two 16-byte cp.async per lane (one for the Jacobian role, which reads X Y Z only) fetch block
// k + kSwAhead* into a small per-warp ring while block k is consumed - no view logic on that path.
// A warp loads the NEXT chunks' observations into registers before it computes the current one
// (shared-memory latency off the dependent chain).  Loops: views outside (pose and size of the next
// view are fetched one view ahead with plain loads), full chunks inside, the partial last chunk of
// a view peeled off, so the inner loops carry no masking and no view bookkeeping.
//
// Operand precision: the float32 rows are rounded to FP16 (11 significant bits, as TF32; the pose columns
// are scaled by a power of two so that nothing leaves the FP16 range, and the conversion saturates).
// tools/h_precision_study.py shows that the exact-Schur LM takes the same iterations with such
// blocks and lands within 1e-10 of the all-FP64 solve (oracle experiment mode 311); CCG_SW_F16=0 keeps the
// TF32 m16n8k8 variant.
// Why legacy mma.sync and not tcgen05: the tile is 16 x 16 per VIEW; tcgen05's smallest M is 64,
// so four views would share an instruction through a block-diagonal A with 3/4 zeros, which makes
// it slower per observation than mma.sync (measured 8.6 cycles per mma.sync per SM sub-partition,
// tools/ubench_pipes.cu) while the kernel is bound by instruction issue, not by the tensor pipe.
#pragma once

namespace ccg {

// (FP64 warp, Jacobian warp) pairs per CTA: a template parameter PAIRS of everything below.  An SM always holds 8
// pairs, two per SM sub-partition: ONE CTA of 8 pairs or TWO CTAs of 4 pairs (one warpgroup per role, which is what
// setmaxnreg needs).  Two smaller CTAs overlap one CTA's latency-bound prologue and epilogue (LM state, candidate
// poses, first blocks of the streams, final reduction) with the other's streaming - what a batch of 50-view problems,
// one CTA each, spends most of its time on (C4: 0.59 -> 0.48 ms per sweep launch); a big problem with hundreds of
// views per CTA is 2 % faster with one CTA of 8 pairs (build_problem chooses).
constexpr int kSwPairs = 8;              // the larger configuration (the persistent LM kernel uses it)
constexpr int kSwWarps = 2 * kSwPairs;   // warps 0 .. PAIRS-1: Jacobian role; the rest: FP64 role
constexpr int kSwThreads = kSwWarps * 32;
// register budgets: the CTA is launched with 128 registers per thread (8 pairs per SM in either configuration); the
// Jacobian warpgroups give registers back (setmaxnreg.dec) and the FP64 warpgroups take them (.inc)
constexpr int kSwRegLaunch = (65536 / kSwThreads) & ~7;
#ifndef CCG_SW_REGJ
#define CCG_SW_REGJ 88
#endif
constexpr int kSwRegJ = CCG_SW_REGJ;
constexpr int kSwRegD = 2 * kSwRegLaunch - kSwRegJ;
static_assert(kSwRegD >= kSwRegLaunch && kSwRegD <= 256 && kSwRegJ % 8 == 0, "register split");

#ifndef CCG_SW_AHEAD_D
#define CCG_SW_AHEAD_D 6
#endif
#ifndef CCG_SW_AHEAD_J
#define CCG_SW_AHEAD_J 4
#endif
constexpr int kSwAheadD = CCG_SW_AHEAD_D, kSwStagesD = kSwAheadD + 2;  // blocks in flight / ring size
constexpr int kSwAheadJ = CCG_SW_AHEAD_J, kSwStagesJ = kSwAheadJ + 2;
// CCG_SW_D2: two chunks per trip INTERLEAVED statement by statement in the FP64 role.  Measured on C5: no effect -
// ptxas serialises the two chains under the register budget (profiles/r02_sweep_experiments.md).  CCG_SW_DU2 (below)
// is the variant that pays: two chunks per trip one AFTER the other, which only removes the register moves.
#ifndef CCG_SW_D2
#define CCG_SW_D2 0
#endif
#ifndef CCG_SW_J2
#define CCG_SW_J2 1
#endif
#ifndef CCG_SW_DU2
#define CCG_SW_DU2 1
#endif
#ifndef CCG_SW_JPIPE
#define CCG_SW_JPIPE 1
#endif
constexpr bool kSwD2 = CCG_SW_D2 != 0, kSwJ2 = CCG_SW_J2 != 0, kSwDU2 = CCG_SW_DU2 != 0, kSwJPipe = CCG_SW_JPIPE != 0;
// A Jacobian warp can also run the FP64 role on the last kSwTail / 64 of its pair's views once its own work is
// done (static split: sums stay bit-reproducible).  Measured on C5: 0 .. 22 / 64 all within 2 % - the FP64 role is
// bound by its pipe, not by the number of warps that feed it - so it is off by default.
#ifndef CCG_SW_TAIL
#define CCG_SW_TAIL 0
#endif
constexpr int kSwTail = CCG_SW_TAIL;
constexpr int kSwTiles = (kSwJ2 || kSwJPipe) ? 2 : 1;  // Jacobian-row tiles per Jacobian warp
constexpr int kObsBlock = 160;   // floats per 32-observation block in HBM: u v X Y Z
constexpr int kSwStageD = 160;   // floats per stage: the whole block
constexpr int kSwStageJ = 96;    // floats per stage: X Y Z

// Operand format of the J'J products: 1 = FP16 pairs (u row, v row) per column, mma m16n8k16 (the legacy tensor
// path issues one m16n8k8 or m16n8k16 per 8.6 cycles and SM sub-partition whatever the shape, so half the
// instructions is half the tensor-pipe time - measured 0.845 -> 0.735 ms per C5 launch; FP16 carries the same 11
// significant bits as TF32; the pose columns are
// scaled by a power of two so that no entry leaves the FP16 range), 0 = TF32 rows, mma m16n8k8.
#ifndef CCG_SW_F16
#define CCG_SW_F16 1
#endif
constexpr bool kSwF16 = CCG_SW_F16 != 0;
// CCG_SW_LDSM (FP16 operands only): fragments by ldmatrix instead of LDS.64.  The product is J'J, so the B fragments of a
// k-step are the A fragment's registers regrouped - (a0, a2) and (a1, a3) - but HMMA wants each operand in consecutive
// registers, and ptxas pays two MOVs per HMMA for the regrouping (16 of the role's 193 instructions per chunk).  Two
// ldmatrix.x4.trans of the same eight 8 x 8 blocks, the second with the blocks in B order, deliver both groupings
// directly.  Tile: a row per RESIDUAL (u rows 0..31, v rows 32..63), 16 FP16 in natural column order, row stride 48 B
// (conflict-free for the lane's two STS.128 per row and for the eight row addresses of a block).
#ifndef CCG_SW_LDSM
#define CCG_SW_LDSM 1
#endif
constexpr bool kSwLdsm = kSwF16 && CCG_SW_LDSM != 0;
constexpr int kTcRow = kSwLdsm ? 12 : 20;  // 32-bit words per tile row (LDS.64 layout: 16 + 4, conflict-free STS.128 and LDS.64)
constexpr int kTcTile = ((kSwF16 && !kSwLdsm) ? 32 : 64) * kTcRow;  // one 32-observation tile: a row per observation (FP16 pairs) or per residual

// The Jacobian warp of a pair is the faster one; left alone it runs ahead of its FP64 partner by more than L2 holds
// (all pairs of the GPU together), and every block is then fetched from HBM twice (measured: 1.70 x the algorithmic
// bytes).  Paced, it starts a view only while it is within kSwLead blocks of what its partner has finished (checked
// once per view; the partner publishes its block count once per view).  0 = unpaced.
#ifndef CCG_SW_LEAD
#define CCG_SW_LEAD 64
#endif
constexpr int kSwLead = CCG_SW_LEAD;
constexpr size_t sw_smem_bytes(int pairs) {
  return sizeof(int) * 16 + sizeof(double) * 2 * pairs * kSweepPart +
         sizeof(float) * pairs * (kSwStagesD * kSwStageD + kSwStagesJ * kSwStageJ + kSwTiles * kTcTile);
}
constexpr size_t kSwSmem = sw_smem_bytes(kSwPairs);
static_assert(kSwSmem <= 227 * 1024 && 2 * (sw_smem_bytes(4) + 1024) <= 227 * 1024, "shared memory");
static_assert(kSwTail == 0 || kSwTiles * kTcTile >= kSwStagesD * kSwStageD, "the FP64 tail's ring lives in the tile(s)");

__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], unsigned a0, unsigned a1, unsigned a2,
                                                unsigned a3, unsigned b0, unsigned b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void mma_f16_16x8x16(float (&c)[4], unsigned a0, unsigned a1, unsigned a2,
                                                unsigned a3, unsigned b0, unsigned b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// (lo, hi) -> two FP16 in one register, round to nearest, saturating (an out-of-range entry must not become Inf)
__device__ __forceinline__ unsigned pack_f16x2(float lo, float hi) {
  unsigned d;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

// destination of element (row, col) of the per-view 16 x 16 product inside the view block
// (V21 | gv6 | W[:,fx fy px py] 24 | W[:,k1 k2 p1 p2 k3] 30), or -1.  Parameter order of the
// product: 0..5 pose tangent, 6 fx, 7 fy, 8 px, 9 py, 10..14 k1 k2 p1 p2 k3, 15 unused.
__device__ __forceinline__ int tc_blk_slot(int row, int col) {
  if (row >= 6 || col < row || col >= 15) return -1;
  if (col < 6) return row * 6 - row * (row - 1) / 2 + (col - row);
  const int c = col - 6;
  return c < 4 ? 27 + row * 4 + c : 51 + row * 5 + (c - 4);
}

// slot of U(row, col) in the per-item sweep partial (layout at the top of intr.cu), or -1 for the
// lower triangle / structurally zero pairs.  row, col are product indices (6..14).
__device__ __forceinline__ int tc_u_slot(int row, int col) {
  if (row < 6 || col < row || col >= 15) return -1;
  const int a = row - 6, b = col - 6;  // fx fy px py k1 k2 p1 p2 k3
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
  const int e = a - 4;  // both distortion: packed upper triangle of the 5 x 5 block
  return 45 + e * 5 - e * (e - 1) / 2 + (c - e);
}

// Per-CTA constants of the sweep (registers): intrinsics at the candidate point.
struct TcConst64 {
  double fx, fy, cx, cy, k1, k2, p1, p2, k3;
};
struct TcConst32 {
  float fx, fy, k1, k2, p1, p2, k3;
  float sp;  // power-of-two scale of the six pose columns (FP16 operands), 1 otherwise
};

// ---------------------------------------------------------------------------------------------
// FP64 role: forward model, residual, cost, and the gradients g_u / g_v by reverse mode, one
// observation per lane.  Branch-free: an invalid lane (padding of the view's last block) sees a
// zero observation (finite arithmetic) and its residual is forced to zero, so it contributes
// nothing.  res (RES only): where this lane's residual pair goes.
// ---------------------------------------------------------------------------------------------
template <bool RES, bool MASK>
__device__ __forceinline__ void tc_fp64_observation(const TcConst64& K, const double (&R)[9], const double (&t)[3],
                                                    float mu, float mv, float Xf, float Yf, float Zf, bool valid,
                                                    double (&gu)[9], double (&gv)[6], double& cost,
                                                    double* __restrict__ res) {
  const double X = (double)Xf, Y = (double)Yf, Z = (double)Zf;
  const double dmu = (double)mu, dmv = (double)mv;
  const double pt0 = fma(R[0], X, fma(R[1], Y, fma(R[2], Z, t[0])));
  const double pt1 = fma(R[3], X, fma(R[4], Y, fma(R[5], Z, t[1])));
  const double pt2 = fma(R[6], X, fma(R[7], Y, fma(R[8], Z, t[2])));
  const double iz = fast_rcp(pt2);
  const double xn = pt0 * iz, yn = pt1 * iz;
  const double tx = xn + xn, ty = yn + yn;
  const double r2 = fma(xn, xn, yn * yn), r4 = r2 * r2, r6 = r4 * r2;
  // m = 1 + k1 r2 + k2 r4 + k3 r6 and mp = dm/dr2 by Horner (no derived constants to keep in registers)
  const double h2 = fma(r2, K.k3, K.k2), h1 = fma(r2, h2, K.k1);
  const double m = fma(r2, h1, 1.0), mp = fma(r2, fma(r2, K.k3, h2), h1);
  // a1 = 2 xn yn, a2 = r2 + 2 xn^2, a3 = r2 + 2 yn^2 through tx = 2 xn, ty = 2 yn (the reverse pass needs them anyway)
  const double a1 = tx * yn, a2 = fma(tx, xn, r2), a3 = fma(ty, yn, r2);
  const double xd = fma(xn, m, fma(K.p1, a1, K.p2 * a2));
  const double yd = fma(yn, m, fma(K.p2, a1, K.p1 * a3));
  double ru = fma(K.fx, xd, K.cx - dmu);
  double rv = fma(K.fy, yd, K.cy - dmv);
  if (MASK) {
    ru = valid ? ru : 0.0;
    rv = valid ? rv : 0.0;
  }
  if (RES) {
    if (valid) {
      res[0] = ru;
      res[1] = rv;
    }
  }
  cost = fma(ru, ru, fma(rv, rv, cost));
  const double tu = K.fx * ru, tv = K.fy * rv;  // d cost / d(xd, yd)
  gu[0] = fma(xd, ru, gu[0]);
  gu[1] = fma(yd, rv, gu[1]);
  gu[2] += ru;
  gu[3] += rv;
  const double sr = fma(xn, tu, yn * tv);
  gu[4] = fma(r2, sr, gu[4]);
  gu[5] = fma(r4, sr, gu[5]);
  gu[8] = fma(r6, sr, gu[8]);
  gu[6] = fma(a1, tu, fma(a3, tv, gu[6]));
  gu[7] = fma(a2, tu, fma(a1, tv, gu[7]));
  // d(xd,yd)/d(xn,yn): F00 = m + 2xx mp + 2p1 yn + 6p2 xn, F01 = a1 mp + 2p1 xn + 2p2 yn,
  // F11 = m + 2yy mp + 2p2 xn + 6p1 yn
  const double u1 = K.p1 * ty, u2 = K.p2 * tx, base = m + (u1 + u2);
  const double F00 = fma(2.0, u2, fma(a2 - r2, mp, base));
  const double F01 = fma(a1, mp, fma(K.p1, tx, K.p2 * ty));
  const double F11 = fma(2.0, u1, fma(a3 - r2, mp, base));
  const double w0 = fma(F00, tu, F01 * tv) * iz, w1 = fma(F01, tu, F11 * tv) * iz;  // d cost / d(pt0, pt1)
  const double w2 = -fma(xn, w0, yn * w1);                                          // d cost / d pt2
  gv[3] += w0;
  gv[4] += w1;
  gv[5] += w2;
  // rotation tangent: 2 (y x w) with y = pt - t; accumulated as pt x w, fixed up per view
  gv[0] = fma(pt1, w2, fma(-pt2, w1, gv[0]));
  gv[1] = fma(pt2, w0, fma(-pt0, w2, gv[1]));
  gv[2] = fma(pt0, w1, fma(-pt1, w0, gv[2]));
}

// N observations per lane at a time (N full chunks), written statement by statement across the N
// observations so that the instruction stream interleaves N independent dependency chains (ptxas keeps
// two separately written observations one after the other; DFMA has a latency of 8 cycles and the FP64
// pipe issues a warp instruction every 2.2 cycles, tools/ubench_dfma_latency.cu).  No masking.
template <bool RES, int N>
__device__ __forceinline__ void tc_fp64_observations(const TcConst64& K, const double (&R)[9], const double (&t)[3],
                                                     const float (&muf)[N], const float (&mvf)[N], const float (&Xf)[N],
                                                     const float (&Yf)[N], const float (&Zf)[N], double (&gu)[9],
                                                     double (&gv)[6], double& cost, double* __restrict__ res) {
#define CCG_EACH for (int k = 0; k < N; ++k)
  double X[N], Y[N], Z[N], dmu[N], dmv[N], pt0[N], pt1[N], pt2[N], iz[N], xn[N], yn[N], xx[N], yy[N], xy[N], r2[N], r4[N],
      r6[N], h2[N], h1[N], m[N], mp[N], a1[N], a2[N], a3[N], xd[N], yd[N], ru[N], rv[N], tu[N], tv[N], sr[N], tx[N], ty[N],
      u1[N], u2[N], base[N], F00[N], F01[N], F11[N], w0[N], w1[N], w2[N];
#pragma unroll
  CCG_EACH { X[k] = (double)Xf[k]; Y[k] = (double)Yf[k]; Z[k] = (double)Zf[k]; }
#pragma unroll
  CCG_EACH pt2[k] = fma(R[6], X[k], fma(R[7], Y[k], fma(R[8], Z[k], t[2])));
#pragma unroll
  CCG_EACH pt0[k] = fma(R[0], X[k], fma(R[1], Y[k], fma(R[2], Z[k], t[0])));
#pragma unroll
  CCG_EACH pt1[k] = fma(R[3], X[k], fma(R[4], Y[k], fma(R[5], Z[k], t[1])));
#pragma unroll
  CCG_EACH iz[k] = fast_rcp(pt2[k]);
#pragma unroll
  CCG_EACH { dmu[k] = (double)muf[k]; dmv[k] = (double)mvf[k]; }
#pragma unroll
  CCG_EACH { xn[k] = pt0[k] * iz[k]; yn[k] = pt1[k] * iz[k]; }
#pragma unroll
  CCG_EACH { xx[k] = xn[k] * xn[k]; yy[k] = yn[k] * yn[k]; xy[k] = xn[k] * yn[k]; }
#pragma unroll
  CCG_EACH { r2[k] = xx[k] + yy[k]; a1[k] = xy[k] + xy[k]; }
#pragma unroll
  CCG_EACH { h2[k] = fma(r2[k], K.k3, K.k2); r4[k] = r2[k] * r2[k]; a2[k] = fma(2.0, xx[k], r2[k]); a3[k] = fma(2.0, yy[k], r2[k]); }
#pragma unroll
  CCG_EACH { h1[k] = fma(r2[k], h2[k], K.k1); r6[k] = r4[k] * r2[k]; }
#pragma unroll
  CCG_EACH { m[k] = fma(r2[k], h1[k], 1.0); mp[k] = fma(r2[k], fma(r2[k], K.k3, h2[k]), h1[k]); }
#pragma unroll
  CCG_EACH {
    xd[k] = fma(xn[k], m[k], fma(K.p1, a1[k], K.p2 * a2[k]));
    yd[k] = fma(yn[k], m[k], fma(K.p2, a1[k], K.p1 * a3[k]));
  }
#pragma unroll
  CCG_EACH { ru[k] = fma(K.fx, xd[k], K.cx - dmu[k]); rv[k] = fma(K.fy, yd[k], K.cy - dmv[k]); }
  if (RES) {
#pragma unroll
    CCG_EACH { res[64 * k] = ru[k]; res[64 * k + 1] = rv[k]; }
  }
#pragma unroll
  CCG_EACH { tu[k] = K.fx * ru[k]; tv[k] = K.fy * rv[k]; }  // d cost / d(xd, yd)
#pragma unroll
  CCG_EACH { tx[k] = xn[k] + xn[k]; ty[k] = yn[k] + yn[k]; }
#pragma unroll
  CCG_EACH { u1[k] = K.p1 * ty[k]; u2[k] = K.p2 * tx[k]; sr[k] = fma(xn[k], tu[k], yn[k] * tv[k]); }
#pragma unroll
  CCG_EACH base[k] = m[k] + (u1[k] + u2[k]);
  // d(xd,yd)/d(xn,yn): F00 = m + 2xx mp + 2p1 yn + 6p2 xn, F01 = a1 mp + 2p1 xn + 2p2 yn, F11 = m + 2yy mp + 2p2 xn + 6p1 yn
#pragma unroll
  CCG_EACH {
    F00[k] = fma(2.0, u2[k], fma(a2[k] - r2[k], mp[k], base[k]));
    F01[k] = fma(a1[k], mp[k], fma(K.p1, tx[k], K.p2 * ty[k]));
    F11[k] = fma(2.0, u1[k], fma(a3[k] - r2[k], mp[k], base[k]));
  }
#pragma unroll
  CCG_EACH {  // d cost / d(pt0, pt1, pt2)
    w0[k] = fma(F00[k], tu[k], F01[k] * tv[k]) * iz[k];
    w1[k] = fma(F01[k], tu[k], F11[k] * tv[k]) * iz[k];
  }
#pragma unroll
  CCG_EACH w2[k] = -fma(xn[k], w0[k], yn[k] * w1[k]);
  // accumulation: the sums are taken in observation order, as the one-observation path does
#pragma unroll
  CCG_EACH {
    cost = fma(ru[k], ru[k], fma(rv[k], rv[k], cost));
    gu[0] = fma(xd[k], ru[k], gu[0]);
    gu[1] = fma(yd[k], rv[k], gu[1]);
    gu[2] += ru[k];
    gu[3] += rv[k];
    gu[4] = fma(r2[k], sr[k], gu[4]);
    gu[5] = fma(r4[k], sr[k], gu[5]);
    gu[8] = fma(r6[k], sr[k], gu[8]);
    gu[6] = fma(a1[k], tu[k], fma(a3[k], tv[k], gu[6]));
    gu[7] = fma(a2[k], tu[k], fma(a1[k], tv[k], gu[7]));
    gv[3] += w0[k];
    gv[4] += w1[k];
    gv[5] += w2[k];
    // rotation tangent: 2 (y x w) with y = pt - t; accumulated as pt x w, fixed up per view
    gv[0] = fma(pt1[k], w2[k], fma(-pt2[k], w1[k], gv[0]));
    gv[1] = fma(pt2[k], w0[k], fma(-pt0[k], w2[k], gv[1]));
    gv[2] = fma(pt0[k], w1[k], fma(-pt1[k], w0[k], gv[2]));
  }
#undef CCG_EACH
}

// ---------------------------------------------------------------------------------------------
// Jacobian role: the two float32 Jacobian rows of one observation per lane, in tile order.
// Invalid lanes produce all-zero rows.
// ---------------------------------------------------------------------------------------------
template <bool MASK>
__device__ __forceinline__ void tc_fp32_rows(const TcConst32& K, const float (&Rf)[9], const float (&tf)[3],
                                             float X, float Y, float Z, bool valid, float4 (&ur)[4],
                                             float4 (&vr)[4]) {
  const float one = (!MASK || valid) ? 1.0f : 0.0f;
  const float fxv = (!MASK || valid) ? K.fx : 0.0f, fyv = (!MASK || valid) ? K.fy : 0.0f;
  const float y0 = fmaf(Rf[0], X, fmaf(Rf[1], Y, Rf[2] * Z));
  const float y1 = fmaf(Rf[3], X, fmaf(Rf[4], Y, Rf[5] * Z));
  const float y2 = fmaf(Rf[6], X, fmaf(Rf[7], Y, Rf[8] * Z));
  const float iz = fast_rcpf(y2 + tf[2]);
  const float xn = (y0 + tf[0]) * iz, yn = (y1 + tf[1]) * iz;
  const float xx = xn * xn, yy = yn * yn, xy = xn * yn;
  const float r2 = xx + yy, r4 = r2 * r2, r6 = r4 * r2;
  const float h2 = fmaf(r2, K.k3, K.k2), h1 = fmaf(r2, h2, K.k1);
  const float m = fmaf(r2, h1, 1.0f), mp = fmaf(r2, fmaf(r2, K.k3, h2), h1);
  const float a1 = xy + xy, a2 = fmaf(2.0f, xx, r2), a3 = fmaf(2.0f, yy, r2);
  const float xd = fmaf(xn, m, fmaf(K.p1, a1, K.p2 * a2));
  const float yd = fmaf(yn, m, fmaf(K.p2, a1, K.p1 * a3));
  const float tx = xn + xn, ty = yn + yn;
  const float u1 = K.p1 * ty, u2 = K.p2 * tx, base = m + (u1 + u2);
  const float F00 = fmaf(2.0f, u2, fmaf(a2 - r2, mp, base));
  const float F01 = fmaf(a1, mp, fmaf(K.p1, tx, K.p2 * ty));
  const float F11 = fmaf(2.0f, u1, fmaf(a3 - r2, mp, base));
  const float fiz = fxv * iz * K.sp, giz = fyv * iz * K.sp;
  // d(u,v)/d(camera point): rows (A0 A1 A2), (B0 B1 B2) (times the pose-column scale)
  const float A0 = fiz * F00, A1 = fiz * F01, B0 = giz * F01, B1 = giz * F11;
  const float A2 = -fmaf(A0, xn, A1 * yn), B2 = -fmaf(B0, xn, B1 * yn);
  // d(camera point)/d(rotation tangent) = 2 (y x .)
  const float z0 = y0 + y0, z1 = y1 + y1, z2 = y2 + y2;
  const float ju0 = fmaf(z1, A2, -z2 * A1), ju1 = fmaf(z2, A0, -z0 * A2), ju2 = fmaf(z0, A1, -z1 * A0);
  const float jv0 = fmaf(z1, B2, -z2 * B1), jv1 = fmaf(z2, B0, -z0 * B2), jv2 = fmaf(z0, B1, -z1 * B0);
  const float fxn = fxv * xn, fyn = fyv * yn;
  // tile order of a row: product column m sits at position (m % 8) * 2 + m / 8, so one
  // LDS.64 at float 2g returns columns g and g + 8:
  //   [c0 c8 | c1 c9 | c2 c10 | c3 c11 | c4 c12 | c5 c13 | c6 c14 | c7 c15]
  // u row: c0..5 pose, c6 = xd (fx), c7 = 0 (fy), c8 = 1 (px), c9 = 0 (py), c10..14 = k1 k2 p1 p2 k3
  ur[0] = make_float4(ju0, one, ju1, 0.0f);
  ur[1] = make_float4(ju2, fxn * r2, A0, fxn * r4);
  ur[2] = make_float4(A1, fxv * a1, A2, fxv * a2);
  ur[3] = make_float4(xd * one, fxn * r6, 0.0f, 0.0f);
  // v row: c6 = 0, c7 = yd (fy), c8 = 0, c9 = 1 (py)
  vr[0] = make_float4(jv0, 0.0f, jv1, one);
  vr[1] = make_float4(jv2, fyn * r2, B0, fyn * r4);
  vr[2] = make_float4(B1, fyv * a3, B2, fyv * a1);
  vr[3] = make_float4(0.0f, fyn * r6, yd * one, 0.0f);
}

// rows 0..31 of a tile = u rows of the 32 observations, rows 32..63 = their v rows; row_u = shared
// address of this lane's u row
__device__ __forceinline__ void tc_store_rows(unsigned row_u, const float4 (&ur)[4], const float4 (&vr)[4]);

// [V W; W' U] += rows' rows over one tile.  k-step s contracts tile rows 8s..8s+7 (k = tig -> row
// 8s + 2 tig, k = tig + 4 -> the next row); two accumulator sets per column block halve the
// depth of the dependent HMMA chain.  FULL: all 32 observations are valid.  frag = shared address
// of element (row 2 tig, float 2 g) of the tile.
template <bool FULL>
__device__ __forceinline__ void tc_contract(unsigned frag, unsigned frag_b, int nvalid, float (&c0a)[4], float (&c0b)[4],
                                            float (&c1a)[4], float (&c1b)[4]);

// ---------------------------------------------------------------------------------------------
// Shared-memory access by 32-bit shared-window address (the hot loops keep these addresses in
// registers; generic pointers made ptxas rebuild them from special registers at every use).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float lds_f32(unsigned a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ int lds_s32(unsigned a) {
  int v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ int2 lds_v2s32(unsigned a) {
  int2 v;
  asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ double2 lds_v2f64(unsigned a) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ float2 lds_v2f32(unsigned a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_s32(unsigned a, int v) { asm volatile("st.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_v4f32(unsigned a, const float4& v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void cp_async16(unsigned dst, const void* src) {
#ifdef CCG_SW_CG
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
#else
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
#endif
}
// the same under a predicate instead of a branch (the request of a block stream sits in the hot loops: a branch
// there costs a BSSY/BRA/BSYNC triple and a branch-resolving stall of the instruction that follows)
__device__ __forceinline__ void cp_async16_if(bool on, unsigned dst, const void* src) {
#ifdef CCG_SW_CG
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p cp.async.cg.shared.global [%0], [%1], 16;\n\t}" ::"r"(dst), "l"(src), "r"((int)on) : "memory");
#else
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p cp.async.ca.shared.global [%0], [%1], 16;\n\t}" ::"r"(dst), "l"(src), "r"((int)on) : "memory");
#endif
}

__device__ __forceinline__ void tc_store_rows(unsigned row_u, const float4 (&ur)[4], const float4 (&vr)[4]) {
  if (kSwLdsm) {
    // a row per residual in natural column order: ur / vr hold [c0 c8 | c1 c9], [c2 c10 | c3 c11], ... (tc_fp32_rows)
    uint4 lo, hi;
    lo.x = pack_f16x2(ur[0].x, ur[0].z); lo.y = pack_f16x2(ur[1].x, ur[1].z);
    lo.z = pack_f16x2(ur[2].x, ur[2].z); lo.w = pack_f16x2(ur[3].x, ur[3].z);
    hi.x = pack_f16x2(ur[0].y, ur[0].w); hi.y = pack_f16x2(ur[1].y, ur[1].w);
    hi.z = pack_f16x2(ur[2].y, ur[2].w); hi.w = pack_f16x2(ur[3].y, ur[3].w);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_u), "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_u + 16), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w) : "memory");
    lo.x = pack_f16x2(vr[0].x, vr[0].z); lo.y = pack_f16x2(vr[1].x, vr[1].z);
    lo.z = pack_f16x2(vr[2].x, vr[2].z); lo.w = pack_f16x2(vr[3].x, vr[3].z);
    hi.x = pack_f16x2(vr[0].y, vr[0].w); hi.y = pack_f16x2(vr[1].y, vr[1].w);
    hi.z = pack_f16x2(vr[2].y, vr[2].w); hi.w = pack_f16x2(vr[3].y, vr[3].w);
    const unsigned row_v = row_u + 32 * kTcRow * 4;
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_v), "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_v + 16), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w) : "memory");
    return;
  }
  if (kSwF16) {
    // one row per observation: word p = (u row, v row) at tile position p, as FP16
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 w;
      w.x = pack_f16x2(ur[i].x, vr[i].x);
      w.y = pack_f16x2(ur[i].y, vr[i].y);
      w.z = pack_f16x2(ur[i].z, vr[i].z);
      w.w = pack_f16x2(ur[i].w, vr[i].w);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_u + 16 * i), "r"(w.x), "r"(w.y), "r"(w.z), "r"(w.w)
                   : "memory");
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      sts_v4f32(row_u + 16 * i, ur[i]);
      sts_v4f32(row_u + 32 * kTcRow * 4 + 16 * i, vr[i]);
    }
  }
}

__device__ __forceinline__ void ldsm_x4_trans(unsigned addr, unsigned& r0, unsigned& r1, unsigned& r2, unsigned& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr)
               : "memory");
}

template <bool FULL>
__device__ __forceinline__ void tc_contract(unsigned frag, unsigned frag_b, int nvalid, float (&c0a)[4], float (&c0b)[4],
                                            float (&c1a)[4], float (&c1b)[4]) {
#ifdef CCG_SW_NO_MMA  // timing experiment: no contraction (results are then wrong)
  return;
#endif
  if (kSwLdsm) {
    // k-step s contracts observations 8s..8s+7: k = 0..7 -> their u rows, k = 8..15 -> their v rows
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      if (FULL || 8 * s < nvalid) {
        unsigned a0, a1, a2, a3, b0, b1, b2, b3;
        ldsm_x4_trans(frag + (8 * s) * kTcRow * 4, a0, a1, a2, a3);    // A = J' (16 parameters x 16 residual rows)
        ldsm_x4_trans(frag_b + (8 * s) * kTcRow * 4, b0, b1, b2, b3);  // B: columns 0..7 (b0, b1), columns 8..15 (b2, b3)
        if (s & 1) {
          mma_f16_16x8x16(c0b, a0, a1, a2, a3, b0, b1);
          mma_f16_16x8x16(c1b, a0, a1, a2, a3, b2, b3);
        } else {
          mma_f16_16x8x16(c0a, a0, a1, a2, a3, b0, b1);
          mma_f16_16x8x16(c1a, a0, a1, a2, a3, b2, b3);
        }
      }
    }
    return;
  }
  if (kSwF16) {
    // k-step s contracts observations 8s..8s+7: k = 2 tig, 2 tig + 1 -> (u, v) of observation 8s + 2 tig,
    // k = 2 tig + 8, + 9 -> the next observation; a fragment register is one tile word
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      if (FULL || 8 * s < nvalid) {
        const float2 lo = lds_v2f32(frag + (8 * s) * kTcRow * 4);
        const float2 hi = lds_v2f32(frag + (8 * s + 1) * kTcRow * 4);
        const unsigned a0 = __float_as_uint(lo.x), a1 = __float_as_uint(lo.y);
        const unsigned a2 = __float_as_uint(hi.x), a3 = __float_as_uint(hi.y);
        if (s & 1) {
          mma_f16_16x8x16(c0b, a0, a1, a2, a3, a0, a2);
          mma_f16_16x8x16(c1b, a0, a1, a2, a3, a1, a3);
        } else {
          mma_f16_16x8x16(c0a, a0, a1, a2, a3, a0, a2);
          mma_f16_16x8x16(c1a, a0, a1, a2, a3, a1, a3);
        }
      }
    }
    return;
  }
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    if (FULL || 8 * (s & 3) < nvalid) {
      const float2 lo = lds_v2f32(frag + (8 * s) * kTcRow * 4);
      const float2 hi = lds_v2f32(frag + (8 * s + 1) * kTcRow * 4);
      const unsigned a0 = __float_as_uint(lo.x), a1 = __float_as_uint(lo.y);
      const unsigned a2 = __float_as_uint(hi.x), a3 = __float_as_uint(hi.y);
#ifdef CCG_SW_NO_HMMA  // timing experiment: fragment loads without the MMAs
      c0a[0] += __uint_as_float(a0 ^ a1 ^ a2 ^ a3);
      continue;
#endif
      if (s & 1) {
        mma_tf32_16x8x8(c0b, a0, a1, a2, a3, a0, a2);
        mma_tf32_16x8x8(c1b, a0, a1, a2, a3, a1, a3);
      } else {
        mma_tf32_16x8x8(c0a, a0, a1, a2, a3, a0, a2);
        mma_tf32_16x8x8(c1a, a0, a1, a2, a3, a1, a3);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Block stream of a role warp: the blocks of its view range are contiguous in HBM.  next() requests
// block k + AHEAD (if there is one) into the ring stage that was read two calls ago - every lane has
// read it, the __syncwarp of the previous call came after those reads - commits one cp.async group
// (so cp.async.wait_group counts blocks), waits for block k and returns the shared address of this
// lane's observation in it; the caller reads its registers from there at once.  The ring has
// AHEAD + 2 stages: AHEAD blocks in flight, the one being read, the one read one call ago.
// ---------------------------------------------------------------------------------------------
template <int NARR, int AHEAD>
struct SwStream {
  static constexpr unsigned kStage = (NARR == 5 ? kSwStageD : kSwStageJ) * 4, kRing = (AHEAD + 2) * kStage;
  const char* src;        // this lane's 16-byte piece of the next block to request
  int left;               // blocks not yet requested
  unsigned ring, lane12;  // shared address of the ring; 12 * lane (piece offset - observation offset)
  unsigned fill, rd, rd_prev;  // this lane's piece of the stage to fill | its observation in the stage to read | read last
  __device__ __forceinline__ void request(int lane) {
#ifdef CCG_SW_BRANCHY_REQUEST
    if (left > 0) {
      if (NARR == 5) {
        cp_async16(fill, src);
        if (lane < 8) cp_async16(fill + 512, src + 512);
      } else {
        if (lane < 24) cp_async16(fill, src);
      }
    }
#else
    if (NARR == 5) {
      cp_async16_if(left > 0, fill, src);
      cp_async16_if(left > 0 && lane < 8, fill + 512, src + 512);
    } else {
      cp_async16_if(left > 0 && lane < 24, fill, src);
    }
#endif
    asm volatile("cp.async.commit_group;" ::: "memory");
    src += kObsBlock * 4;
    --left;
  }
  __device__ __forceinline__ void init(const float* obs_blk, int b0, int b1, unsigned ring_, int lane) {
    ring = ring_;
    lane12 = 12u * (unsigned)lane;
    src = reinterpret_cast<const char*>(obs_blk) + (size_t)b0 * (kObsBlock * 4) + (NARR == 5 ? 0 : 256) + 16 * lane;
    left = b1 - b0;
    fill = ring + 16u * (unsigned)lane;
#pragma unroll
    for (int i = 0; i < AHEAD; ++i) {
      request(lane);
      fill += kStage;
    }
    rd = ring + 4u * (unsigned)lane;
    rd_prev = rd + (AHEAD + 1) * kStage;  // the last stage of the ring: free
  }
  __device__ __forceinline__ unsigned next(int lane) {
    request(lane);
    asm volatile("cp.async.wait_group %0;" ::"n"(AHEAD) : "memory");
    __syncwarp();
    const unsigned at = rd;
    fill = rd_prev + lane12;
    rd_prev = rd;
    rd += kStage;
    if (rd == ring + kRing + (lane12 / 3u)) rd -= kRing;
    return at;
  }
};

// keeps a lane-derived value in a register (ptxas otherwise rebuilds it from SR_TID at every use)
__device__ __forceinline__ int opaque(int x) {
  asm volatile("mov.u32 %0, %0;" : "+r"(x));
  return x;
}

template <typename T>
__device__ __forceinline__ T* opaque_ptr(T* p) {
  unsigned long long x = reinterpret_cast<unsigned long long>(p);
  asm volatile("mov.u64 %0, %0;" : "+l"(x));
  return reinterpret_cast<T*>(x);
}

// ---------------------------------------------------------------------------------------------
// Candidate poses of the item's views, by all threads of the sweep CTA before the roles start (one thread
// per view): pose (+) delta_v with delta_v = -z - Y du (the back-substitution of the reduced solve), written
// as the candidate pose and as the rotation matrix + translation the roles consume (double and float);
// s_step[0..1] receive the item's share of the model terms g'd and d'Dd (fixed-order sums).
// ---------------------------------------------------------------------------------------------
template <int PAIRS>
__device__ __forceinline__ void sw_candidate_poses(const IntrDev& d, const WorkItem& it, int par, const double* s_du,
                                                   double (*s_red)[2 * PAIRS], double* s_step) {
  double gTd = 0.0, dDd = 0.0;
  for (int64_t v = it.v0 + threadIdx.x; v < it.v1; v += 64 * PAIRS) {
    const double* yz = d.yz + (size_t)v * kYz;
    double dl[6];
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      double s = -yz[54 + a];
#pragma unroll
      for (int j = 0; j < 9; ++j) s -= yz[a * 9 + j] * s_du[j];
      dl[a] = s;
      gTd += yz[60 + a] * s;
      dDd += s * s * yz[66 + a];
    }
    const double* pc = d.pose[par] + (size_t)v * kPose;
    const double q[4] = {pc[0], pc[1], pc[2], pc[3]};
    double qn[4], R[9];
    quat_plus(q, dl[0], dl[1], dl[2], qn);
    const double t0 = pc[4] + dl[3], t1 = pc[5] + dl[4], t2 = pc[6] + dl[5];
    double* pn = d.pose[par ^ 1] + (size_t)v * kPose;
    pn[0] = qn[0]; pn[1] = qn[1]; pn[2] = qn[2]; pn[3] = qn[3];
    pn[4] = t0; pn[5] = t1; pn[6] = t2;
    quat_to_rot(qn, R);
    double2* rt = reinterpret_cast<double2*>(d.rt + (size_t)v * 12);
    rt[0] = make_double2(R[0], R[1]);
    rt[1] = make_double2(R[2], R[3]);
    rt[2] = make_double2(R[4], R[5]);
    rt[3] = make_double2(R[6], R[7]);
    rt[4] = make_double2(R[8], t0);
    rt[5] = make_double2(t1, t2);
    float4* rtf = reinterpret_cast<float4*>(d.rtf + (size_t)v * 12);  // the Jacobian role's copy
    rtf[0] = make_float4((float)R[0], (float)R[1], (float)R[2], (float)R[3]);
    rtf[1] = make_float4((float)R[4], (float)R[5], (float)R[6], (float)R[7]);
    rtf[2] = make_float4((float)R[8], (float)t0, (float)t1, (float)t2);
  }
  // fixed-order block reduction (bit-reproducible)
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    gTd += __shfl_xor_sync(0xffffffffu, gTd, s);
    dDd += __shfl_xor_sync(0xffffffffu, dDd, s);
  }
  if ((threadIdx.x & 31) == 0) {
    s_red[0][threadIdx.x >> 5] = gTd;
    s_red[1][threadIdx.x >> 5] = dDd;
  }
  __syncthreads();  // also: the poses are visible to the whole CTA
  if (threadIdx.x < 2) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < 2 * PAIRS; ++w) s += s_red[threadIdx.x][w];
    s_step[threadIdx.x] = s;
  }
}


// ---------------------------------------------------------------------------------------------
// The FP64 role over the views [va, vb) (absolute view indices): residuals, cost, g_u (into the
// warp's partial) and g_v (into the view blocks).  `stages` = shared address of the warp's ring.
// ---------------------------------------------------------------------------------------------
struct SwObs5 {
  float mu, mv, X, Y, Z;
};
struct SwObs3 {
  float X, Y, Z;
};

// LEAN (the Jacobian warps' share, fewer registers): one chunk at a time, the pose is loaded at the view's start.
// strm: the warp's block stream over [va, vb), already initialised (its first blocks are in flight: sw_sweep_item
// starts the streams BEFORE the candidate poses are computed, so the first HBM round trip hides behind that phase).
template <bool RES, bool LEAN>
__device__ __forceinline__ void sw_run_fp64(const IntrDev& d, SwStream<5, kSwAheadD>& strm, unsigned pace, const double* s_k,
                                            int par, int va, int vb, int lane, double* upart, double* __restrict__ res_out) {
  TcConst64 K;
  K.fx = s_k[0]; K.fy = s_k[1]; K.cx = s_k[2]; K.cy = s_k[3];
  K.k1 = s_k[4]; K.k2 = s_k[5]; K.p1 = s_k[6]; K.p2 = s_k[7]; K.k3 = s_k[8];
  double gu[9], cost = 0.0;
#pragma unroll
  for (int i = 0; i < 9; ++i) gu[i] = 0.0;
  if (va < vb) {
    lane = opaque(lane);
    const int2* vmeta = opaque_ptr(d.vmeta);
    const double2* rt = reinterpret_cast<const double2*>(opaque_ptr(d.rt));
    double* gv_out = opaque_ptr(d.blk[par ^ 1] + 21 + lane);
    // the lane keeps the observations of the next two chunks of its stream in registers (o0: next, o1: the one
    // after): the shared-memory latency of a chunk hides behind the arithmetic of the chunks before it
    int consumed = 0;  // blocks read so far, published for the partner's pacing
    auto load = [&]() -> SwObs5 {
      const unsigned a = strm.next(lane);
      SwObs5 o;
      o.mu = lds_f32(a); o.mv = lds_f32(a + 128); o.X = lds_f32(a + 256); o.Y = lds_f32(a + 384); o.Z = lds_f32(a + 512);
      return o;
    };
    SwObs5 o0 = load(), o1 = load();
    double2 pn[6];
    if (!LEAN) {
#pragma unroll
      for (int i = 0; i < 6; ++i) pn[i] = rt[(size_t)va * 6 + i];
    }
    int n_next = vmeta[va].y;
    for (int v = va; v < vb; ++v) {
      if (LEAN) {
#pragma unroll
        for (int i = 0; i < 6; ++i) pn[i] = rt[(size_t)v * 6 + i];
      }
      const double R[9] = {pn[0].x, pn[0].y, pn[1].x, pn[1].y, pn[2].x, pn[2].y, pn[3].x, pn[3].y, pn[4].x};
      const double t[3] = {pn[4].y, pn[5].x, pn[5].y};
      const int n = n_next;
      {
        // next view's pose and size, one view ahead (the last view re-reads itself)
        const int vn = min(v + 1, vb - 1);
        if (!LEAN) {
#pragma unroll
          for (int i = 0; i < 6; ++i) pn[i] = rt[(size_t)vn * 6 + i];
        }
        n_next = vmeta[vn].y;
      }
      double gv[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) gv[i] = 0.0;
      double* res = nullptr;
      if (RES) res = res_out + 2 * (d.view_off[v] - d.obs_base + lane);
      int c = n >> 5;  // full chunks: no masking
      if (kSwD2 && !LEAN) {
        for (; c >= 2; c -= 2) {
          // two chunks at a time: two independent chains per lane
          const SwObs5 a = o0, b = o1;
          o0 = load();
          o1 = load();
          const float mu2[2] = {a.mu, b.mu}, mv2[2] = {a.mv, b.mv}, X2[2] = {a.X, b.X}, Y2[2] = {a.Y, b.Y}, Z2[2] = {a.Z, b.Z};
          tc_fp64_observations<RES, 2>(K, R, t, mu2, mv2, X2, Y2, Z2, gu, gv, cost, res);
          if (RES) res += 128;
        }
      }
      if (kSwDU2 && !LEAN) {
        // two chunks per trip, one after the other: the two-deep register queue (o0, o1) turns over in place, so the
        // loop carries no register moves (the one-chunk loop below spends 10 of its 136 instructions on them)
        for (; c >= 2; c -= 2) {
          tc_fp64_observation<RES, false>(K, R, t, o0.mu, o0.mv, o0.X, o0.Y, o0.Z, true, gu, gv, cost, res);
          o0 = load();
          tc_fp64_observation<RES, false>(K, R, t, o1.mu, o1.mv, o1.X, o1.Y, o1.Z, true, gu, gv, cost, RES ? res + 64 : res);
          o1 = load();
          if (RES) res += 128;
        }
      }
      for (; c >= 1; --c) {
        const SwObs5 a = o0;
        o0 = o1;
        o1 = load();
        tc_fp64_observation<RES, false>(K, R, t, a.mu, a.mv, a.X, a.Y, a.Z, true, gu, gv, cost, res);
        if (RES) res += 64;
      }
      if (n & 31) {
        const SwObs5 a = o0;
        o0 = o1;
        o1 = load();
        tc_fp64_observation<RES, true>(K, R, t, a.mu, a.mv, a.X, a.Y, a.Z, lane < (n & 31), gu, gv, cost, res);
      }
      // per-view gradient: reduce over lanes, fix up the rotation part, store
      double tmp[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) tmp[i] = i < 6 ? gv[i] : 0.0;
      double r = group_reduce<8, double>(tmp, lane & 7);
      r += __shfl_xor_sync(0xffffffffu, r, 8);
      r += __shfl_xor_sync(0xffffffffu, r, 16);
      // lanes 0..5 hold sum(pt x w) (0..2) and sum(w) (3..5); y x w = pt x w - t x sum(w)
      const double sw0 = __shfl_sync(0xffffffffu, r, 3), sw1 = __shfl_sync(0xffffffffu, r, 4),
                   sw2 = __shfl_sync(0xffffffffu, r, 5);
      double out = r;
      if (lane == 0) out = 2.0 * (r - (t[1] * sw2 - t[2] * sw1));
      if (lane == 1) out = 2.0 * (r - (t[2] * sw0 - t[0] * sw2));
      if (lane == 2) out = 2.0 * (r - (t[0] * sw1 - t[1] * sw0));
      if (lane < 6) gv_out[(size_t)v * kBlk] = out;
      if (!LEAN && kSwLead > 0) {  // blocks consumed so far, published once per view (only when the partner is paced)
        consumed += (n + 31) >> 5;
        if (lane == 0) asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(pace), "r"(consumed) : "memory");
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  if (!LEAN && kSwLead > 0 && lane == 0) asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(pace), "r"(0x7fffffff) : "memory");
  // lane-private sums -> the warp's partial: [0] sum r^2 | [16..24] g_u
  double tmp[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) tmp[i] = 0.0;
#pragma unroll
  for (int i = 0; i < 9; ++i) tmp[i] = gu[i];
  tmp[9] = cost;
  double r = group_reduce<16, double>(tmp, lane & 15);
  r += __shfl_xor_sync(0xffffffffu, r, 16);
  if (lane < 9) upart[16 + lane] = r;
  else if (lane == 9) upart[0] = r;
}

// ---------------------------------------------------------------------------------------------
// The Jacobian + tensor role over the views [va, vb): V, W (into the view blocks) and U (into the
// warp's partial).  `stages`, `tile` = shared addresses.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void sw_run_jac(const IntrDev& d, SwStream<3, kSwAheadJ>& strm, unsigned tile, unsigned pace,
                                           const double* s_k, int par, int va, int vb, int lane, double* upart) {
  if (va >= vb) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    return;
  }
  lane = opaque(lane);
  tile = (unsigned)opaque((int)tile);
  const int g = lane >> 2, tig = lane & 3;
  TcConst32 K;
  K.fx = (float)s_k[0]; K.fy = (float)s_k[1];
  K.k1 = (float)s_k[4]; K.k2 = (float)s_k[5]; K.p1 = (float)s_k[6]; K.p2 = (float)s_k[7]; K.k3 = (float)s_k[8];
  // FP16 operands: scale the pose columns so that the focal length they carry is below 512 (their entries are a
  // few times fx / depth); undone exactly (powers of two) when the blocks are written
  const int sp_e = kSwF16 ? max(0, ilogbf(fmaxf(fabsf(K.fx), fabsf(K.fy))) - 8) : 0;
  K.sp = ldexpf(1.0f, -sp_e);
  const double isp = (double)ldexpf(1.0f, sp_e);
  const int2* vmeta = opaque_ptr(d.vmeta);
  const float4* rtf = reinterpret_cast<const float4*>(opaque_ptr(d.rtf));
  double* blk0 = opaque_ptr(d.blk[par ^ 1]);
  int started = 0;  // blocks of the views this warp has started (pacing)
  auto load = [&]() -> SwObs3 {
    const unsigned a = strm.next(lane);
    SwObs3 o;
    o.X = lds_f32(a); o.Y = lds_f32(a + 128); o.Z = lds_f32(a + 256);
    return o;
  };
  SwObs3 o0 = load(), o1 = load();
  // destinations of this lane's eight product elements (view block slot, U slot; 0xff = none), packed four to a
  // register once per work item: derived per view they cost more integer instructions than the epilogue itself
  unsigned bsp[2] = {0u, 0u}, usp[2] = {0u, 0u};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int nt = i >> 2, e = i & 3;
    const int row = g + ((e & 2) ? 8 : 0), col = nt * 8 + 2 * tig + (e & 1);
    bsp[i >> 2] |= (unsigned)(tc_blk_slot(row, col) & 0xff) << (8 * (i & 3));
    usp[i >> 2] |= (unsigned)(tc_u_slot(row, col) & 0xff) << (8 * (i & 3));
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    bsp[i] = (unsigned)opaque((int)bsp[i]);
    usp[i] = (unsigned)opaque((int)usp[i]);
  }
  const unsigned row_u = tile + (unsigned)lane * (kTcRow * 4);            // this lane's u row (v row: + 32 rows)
  // this lane's fragment of k-step 0 (LDS.64 layout) | its row address for the two ldmatrix of k-step 0: lanes 8i..8i+7
  // address block i; A order: (u rows, columns 0..7), (u, 8..15), (v, 0..7), (v, 8..15); B order: (u, 0..7), (v, 0..7),
  // (u, 8..15), (v, 8..15)
  const unsigned frag = kSwLdsm ? tile + (unsigned)(((lane & 7) + 32 * ((lane >> 4) & 1)) * kTcRow * 4 + ((lane >> 3) & 1) * 16)
                                : tile + (unsigned)(2 * tig * kTcRow + 2 * g) * 4;
  const unsigned frag_b = (unsigned)opaque((int)(tile + (unsigned)(((lane & 7) + 32 * ((lane >> 3) & 1)) * kTcRow * 4 + ((lane >> 4) & 1) * 16)));
  float4 pn[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) pn[i] = rtf[(size_t)va * 3 + i];
  int n_next = vmeta[va].y;
  for (int v = va; v < vb; ++v) {
    const float Rf[9] = {pn[0].x, pn[0].y, pn[0].z, pn[0].w, pn[1].x, pn[1].y, pn[1].z, pn[1].w, pn[2].x};
    const float tf[3] = {pn[2].y, pn[2].z, pn[2].w};
    const int n = n_next;
    {
      const int vn = min(v + 1, vb - 1);
#pragma unroll
      for (int i = 0; i < 3; ++i) pn[i] = rtf[(size_t)vn * 3 + i];
      n_next = vmeta[vn].y;
    }
#ifndef CCG_SW_ONLY_J
    if (kSwLead > 0 && kSwTail == 0) {
      // pacing, once per view: start a view only within kSwLead blocks of what the FP64 partner has finished (see kSwLead)
      for (;;) {
        int seen;
        asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(seen) : "r"(pace) : "memory");
        if (started - kSwLead <= seen) break;  // (seen = INT_MAX once the partner is done)
        __nanosleep(128);
      }
      started += (n + 31) >> 5;
    }
#endif
    // product columns 0..7 (c0*) and 8..15 (c1*), two accumulator sets each
    float c0a[4] = {0.f, 0.f, 0.f, 0.f}, c0b[4] = {0.f, 0.f, 0.f, 0.f};
    float c1a[4] = {0.f, 0.f, 0.f, 0.f}, c1b[4] = {0.f, 0.f, 0.f, 0.f};
    // per chunk: rows -> tile -> contraction.  The __syncwarp inside load() orders the previous chunk's tile
    // reads before this chunk's tile writes.
    int c = n >> 5;  // full chunks: no masking
    if (kSwJPipe && c >= 2) {
      // software pipeline inside the view: the rows of chunk j are computed while the tile of chunk j - 1 is
      // contracted - the FP32 arithmetic has no dependence on the fragment loads and the HMMAs and is written
      // BEFORE the tile store, so ptxas is free to fill the 8.6-cycle issue interval of the legacy tensor path
      // with it.  Two tiles; the __syncwarp inside load() separates a tile's store from its fragment loads, and
      // the fragment loads of one trip from the store into the same tile two trips later.
      const unsigned t0 = 0u, t1 = kTcTile * 4;
      float4 ur[4], vr[4];
      tc_fp32_rows<false>(K, Rf, tf, o0.X, o0.Y, o0.Z, true, ur, vr);
      tc_store_rows(row_u + t0, ur, vr);
      o0 = load();  // queue now: o1 next, o0 after it
      for (c -= 1; c >= 2; c -= 2) {
        tc_fp32_rows<false>(K, Rf, tf, o1.X, o1.Y, o1.Z, true, ur, vr);
        tc_contract<true>(frag + t0, frag_b + t0, 32, c0a, c0b, c1a, c1b);
        tc_store_rows(row_u + t1, ur, vr);
        o1 = load();
        tc_fp32_rows<false>(K, Rf, tf, o0.X, o0.Y, o0.Z, true, ur, vr);
        tc_contract<true>(frag + t1, frag_b + t1, 32, c0a, c0b, c1a, c1b);
        tc_store_rows(row_u + t0, ur, vr);
        o0 = load();
      }
      if (c == 1) {
        tc_fp32_rows<false>(K, Rf, tf, o1.X, o1.Y, o1.Z, true, ur, vr);
        tc_contract<true>(frag + t0, frag_b + t0, 32, c0a, c0b, c1a, c1b);
        tc_store_rows(row_u + t1, ur, vr);
        o1 = load();  // queue back in order: o0 next, o1 after it
        tc_contract<true>(frag + t1, frag_b + t1, 32, c0a, c0b, c1a, c1b);
      } else {
        tc_contract<true>(frag + t0, frag_b + t0, 32, c0a, c0b, c1a, c1b);
        const SwObs3 x = o0;
        o0 = o1;
        o1 = x;
      }
      c = 0;
    }
    if (kSwJ2) {
      for (; c >= 2; c -= 2) {
        const SwObs3 a = o0, b = o1;
        o0 = load();
        o1 = load();
        float4 ur[4], vr[4];
        tc_fp32_rows<false>(K, Rf, tf, a.X, a.Y, a.Z, true, ur, vr);
        tc_store_rows(row_u, ur, vr);
        tc_fp32_rows<false>(K, Rf, tf, b.X, b.Y, b.Z, true, ur, vr);
        tc_store_rows(row_u + kTcTile * 4, ur, vr);
        __syncwarp();
        tc_contract<true>(frag, frag_b, 32, c0a, c0b, c1a, c1b);
        tc_contract<true>(frag + kTcTile * 4, frag_b + kTcTile * 4, 32, c0a, c0b, c1a, c1b);
      }
    }
    for (; c >= 1; --c) {
      const SwObs3 a = o0;
      o0 = o1;
      o1 = load();
      float4 ur[4], vr[4];
      tc_fp32_rows<false>(K, Rf, tf, a.X, a.Y, a.Z, true, ur, vr);
      tc_store_rows(row_u, ur, vr);
      __syncwarp();
      tc_contract<true>(frag, frag_b, 32, c0a, c0b, c1a, c1b);
    }
    if (n & 31) {
      const SwObs3 a = o0;
      o0 = o1;
      o1 = load();
      float4 ur[4], vr[4];
      tc_fp32_rows<true>(K, Rf, tf, a.X, a.Y, a.Z, lane < (n & 31), ur, vr);
      tc_store_rows(row_u, ur, vr);
      __syncwarp();
      tc_contract<false>(frag, frag_b, n & 31, c0a, c0b, c1a, c1b);
    }
    // per-view epilogue: V, W blocks and the view's contribution to U
    double* blk = blk0 + (size_t)v * kBlk;
    const double sc0 = tig < 3 ? isp * isp : isp;  // columns 0..7 of this lane: pose x pose (V) below column 6, else W
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int nt = i >> 2, e = i & 3;
      const float val = nt == 0 ? c0a[e] + c0b[e] : c1a[e] + c1b[e];
      const unsigned bs = (bsp[i >> 2] >> (8 * (i & 3))) & 0xffu, us = (usp[i >> 2] >> (8 * (i & 3))) & 0xffu;
      if (bs != 0xffu) blk[bs] = (double)val * (nt == 0 ? sc0 : isp);  // V: pose x pose, W: pose x intrinsics
      if (us != 0xffu) upart[us] += (double)val;  // one owner lane per slot within the warp
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// The sweep of one work item by one CTA (all kSwThreads threads call): candidate poses, then the two
// roles, then the item's 64-value partial into `out` (global or shared memory).  s_k: intrinsics at the
// candidate point, s_du: the reduced step (both in shared memory).
// ---------------------------------------------------------------------------------------------
template <int PAIRS>
struct SwSmem {  // dynamic shared memory of a sweep CTA
  int* pace;  // per pair: blocks the FP64 warp has consumed (the Jacobian warp's pacing)
  double (*part)[kSweepPart];
  float *stage_d, *stage_j, *tiles;
  __device__ __forceinline__ explicit SwSmem(unsigned char* dyn) {
    pace = reinterpret_cast<int*>(dyn);
    dyn += sizeof(int) * 16;
    part = reinterpret_cast<double(*)[kSweepPart]>(dyn);
    stage_d = reinterpret_cast<float*>(dyn + sizeof(double) * 2 * PAIRS * kSweepPart);
    stage_j = stage_d + PAIRS * kSwStagesD * kSwStageD;
    tiles = stage_j + PAIRS * kSwStagesJ * kSwStageJ;
  }
};

// ROLE: 0 = the calling CTA's warps take their roles here and split the registers (the sweep as a launch of
// its own); 1 / 2 = the caller is an FP64 / a Jacobian warp whose registers are already set (the persistent LM
// kernel instantiates its whole loop once per role, so that no code is shared between the two register budgets).
template <bool RES, int ROLE, int PAIRS>
__device__ __forceinline__ void sw_sweep_item(const IntrDev& d, const WorkItem& it, int par, const double* s_k,
                                              const double* s_du, const SwSmem<PAIRS>& sm, double (*s_red)[2 * PAIRS],
                                              double* s_step, double* out, double* __restrict__ res_out) {
  constexpr int kSwPairs = PAIRS, kSwWarps = 2 * PAIRS, kSwThreads = 64 * PAIRS;  // (shadow the 8-pair constants)
  for (int i = threadIdx.x; i < kSwWarps * kSweepPart; i += kSwThreads) (&sm.part[0][0])[i] = 0.0;
  if (threadIdx.x < kSwPairs) sm.pace[threadIdx.x] = 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* upart = sm.part[warp];
  // pair p (Jacobian warp p, FP64 warp kSwPairs + p, same SM sub-partition) owns the p-th part of the item's views
  const int pair = warp & (kSwPairs - 1), nv = it.v1 - it.v0;
  const int va = it.v0 + (int)((int64_t)nv * pair / kSwPairs), vb = it.v0 + (int)((int64_t)nv * (pair + 1) / kSwPairs);
  const int vsplit = va + ((vb - va) * (64 - kSwTail) + 32) / 64;  // FP64 role: [va, vsplit) the FP64 warp, the rest its partner
  const bool is_d = ROLE == 1 || (ROLE == 0 && warp >= kSwPairs);
  // The block streams start here: they depend on the observations only, so their first requests are on the way to HBM
  // while the candidate poses are computed (per CTA a microsecond or two - 8 % of the sweep of a 50-view problem).
  // (Measured: -2.5 % per launch on C4's 50-view problems, nothing at C5; the persistent LM kernel, ROLE 1 / 2, starts
  // its streams after the poses - there the early start measured slightly slower.)
  SwStream<5, kSwAheadD> strm_d;
  SwStream<3, kSwAheadJ> strm_j;
#ifndef CCG_SW_LATE_STREAMS
  constexpr bool kEarly = ROLE == 0;
#else
  constexpr bool kEarly = false;
#endif
  if (kEarly) {
    if (is_d) strm_d.init(d.obs_blk, d.vmeta[va].x, d.vmeta[vsplit].x, smem_u32(sm.stage_d + pair * kSwStagesD * kSwStageD), lane);
    else strm_j.init(d.obs_blk, d.vmeta[va].x, d.vmeta[vb].x, smem_u32(sm.stage_j + pair * kSwStagesJ * kSwStageJ), lane);
  }
  sw_candidate_poses<PAIRS>(d, it, par, s_du, s_red, s_step);  // (its barrier also covers the zeroing above)
  if (!kEarly) {
    if (is_d) strm_d.init(d.obs_blk, d.vmeta[va].x, d.vmeta[vsplit].x, smem_u32(sm.stage_d + pair * kSwStagesD * kSwStageD), lane);
    else strm_j.init(d.obs_blk, d.vmeta[va].x, d.vmeta[vb].x, smem_u32(sm.stage_j + pair * kSwStagesJ * kSwStageJ), lane);
  }
  if (is_d) {
    // =========================== FP64 role ===========================
    if (ROLE == 0 && kSwRegD > kSwRegLaunch) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kSwRegD));
#ifndef CCG_SW_ONLY_J  // timing experiments: one role alone (results are then incomplete)
    sw_run_fp64<RES, false>(d, strm_d, smem_u32(sm.pace + pair), s_k, par, va, vsplit, lane, upart, res_out);
#else
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
  } else {
    // =========================== Jacobian + tensor role ===========================
    if (ROLE == 0 && kSwRegJ < kSwRegLaunch) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kSwRegJ));
#ifndef CCG_SW_ONLY_D
    sw_run_jac(d, strm_j, smem_u32(sm.tiles + pair * kSwTiles * kTcTile), smem_u32(sm.pace + pair), s_k, par, va, vb, lane,
               upart);
#else
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
#if !defined(CCG_SW_ONLY_J) || defined(CCG_SW_TAIL_ANYWAY)
    // the tail of the pair's FP64 work; the ring lives in the tile(s), free by now (every lane is past its last
    // tile read: the warp is synchronised before the first ring write lands)
    if (kSwTail > 0) {
      __syncwarp();
      SwStream<5, kSwAheadD> strm_t;
      strm_t.init(d.obs_blk, d.vmeta[vsplit].x, d.vmeta[vb].x, smem_u32(sm.tiles + pair * kSwTiles * kTcTile), lane);
      sw_run_fp64<RES, true>(d, strm_t, 0u, s_k, par, vsplit, vb, lane, upart, res_out);
    }
#endif
  }
  __syncthreads();
  if (threadIdx.x < kSweepPart) {  // (Jacobian-role threads: warps 0 and 1)
    double s = 0.0;
    if (threadIdx.x == 7 || threadIdx.x == 8) {
      s = s_step[threadIdx.x - 7];  // the model terms g'd, d'Dd of this item
    } else {
#pragma unroll
      for (int w = 0; w < kSwWarps; ++w) s += sm.part[w][threadIdx.x];
    }
    out[threadIdx.x] = s;
  }
}

// The sweep as a launch of its own: one CTA per work item, state read from global memory.
template <bool RES, int PAIRS>
__global__ void __launch_bounds__(64 * PAIRS, 8 / PAIRS) k_intr_sweep_ws(IntrDev d, double* __restrict__ res_out) {
  extern __shared__ __align__(16) unsigned char s_dyn[];
  __shared__ double s_k[9], s_du[9], s_step[2];
  __shared__ double s_red[2][2 * PAIRS];
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
  sw_sweep_item<RES, 0, PAIRS>(d, it, par, s_k, s_du, SwSmem<PAIRS>(s_dyn), s_red, s_step,
                               d.sweep_part + (size_t)blockIdx.x * kSweepPart, res_out);
}

// Observations SoA -> 640-byte blocks of 32 (views padded to whole blocks with zeros); one warp per view.
__global__ void __launch_bounds__(256) k_intr_pack(IntrDev d, int64_t n_views, float* __restrict__ out) {
  const int64_t v = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (v >= n_views) return;
  const int2 m = d.vmeta[v];
  const int64_t o0 = d.view_off[v];
  for (int c = 0; 32 * c < m.y; ++c) {
    const int i = 32 * c + lane;
    const bool ok = i < m.y;
    float* dst = out + (size_t)(m.x + c) * kObsBlock + lane;
    dst[0] = ok ? d.u[o0 + i] : 0.0f;
    dst[32] = ok ? d.v[o0 + i] : 0.0f;
    dst[64] = ok ? d.X[o0 + i] : 0.0f;
    dst[96] = ok ? d.Y[o0 + i] : 0.0f;
    dst[128] = ok ? d.Z[o0 + i] : 0.0f;
  }
}

}  // namespace ccg
