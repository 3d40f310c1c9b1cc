// Warp-specialised single-camera sweep (included by intr.cu) - the default sweep.
//
// Per observation the sweep needs (a) the residual, the cost and the gradient g = J'r, which
// decide WHERE the LM iteration converges, and (b) the Gauss-Newton blocks V, W, U = J'J, which
// only decide HOW FAST (a perturbed J'J still has g = 0 as its fixed point and every step is
// validated against the FP64 cost).  (b) is 2/3 of the arithmetic and is a dense contraction:
// per view  [V W; W' U] = sum over its 2P residual rows of (row)'(row), a 16 x 16 x 2P product.
//   FP64 pipe  : forward model, residual, cost, and g_v / g_u by REVERSE mode (no explicit J)
//   FP32 pipe  : a second forward pass and the analytic 2 x 15 Jacobian rows
//   tensor pipe: J'J by mma.sync.m16n8k8 TF32 with FP32 accumulation - one lane owns one
//                observation, writes its two Jacobian rows to a warp-private shared-memory tile,
//                and the warp contracts 32 observations with 16 MMAs; the A fragment doubles as
//                the B fragment (the product is J'J), so a k-step costs two LDS.64.
//
// Roles (one CTA per SM, one work item = a range of views per CTA):
//   * every SM sub-partition runs ONE FP64-role warp (the highest warp ids: the issue arbiter
//     prefers them) and kSwJPD Jacobian-role warps.  The FP64 role is bound by the FP64 pipe
//     (~90 two-cycle instructions per 32 observations) and needs registers, not warps; the Jacobian
//     role is a long chain of dependent FP32 -> STS -> LDS -> HMMA and needs warps, not registers.
//     setmaxnreg moves the registers accordingly (FP64 warps kSwRegD, Jacobian warps kSwRegJ).
//   * the four groups of a CTA split the item's views into four ranges; inside a group the FP64
//     warp walks every view of the range and Jacobian warp j the views j, j + kSwJPD, ... - so all
//     warps of a group stay within a few views of each other and the second reader of an
//     observation block hits L1/L2 (HBM traffic stays 20 B per observation).  Roles share nothing
//     but the read-only inputs: no synchronisation inside the loops.
//
// Observation stream: the observations live in HBM as 640-byte blocks of 32 observations
// (u[32] v[32] X[32] Y[32] Z[32], views padded to whole blocks with zeros), so a 32-observation
// chunk is ONE aligned block: two 16-byte cp.async per lane fetch it (one for the Jacobian role,
// which reads X Y Z only) into a small per-warp ring of stages, kSwAhead* chunks ahead of the
// one being consumed.  A stage also carries the pose of the chunk's view when the chunk is the
// first of its view (rotation matrix + translation, written by k_intr_step) and a descriptor
// (view, valid observations, first/last flags) written by the prefetch cursor, so the consumer
// loop is flat: wait, read the descriptor, compute, and at a view's last chunk run the epilogue.
//
// Operand precision: the raw float32 rows are fed to the TF32 unit (10 explicit mantissa bits).
// tools/h_precision_study.py shows that the exact-Schur LM takes the same iterations with such
// blocks and lands within 1e-10 of the all-FP64 solve (oracle experiment mode 311).
// Why legacy mma.sync and not tcgen05: the tile is 16 x 16 per VIEW; tcgen05's smallest M is 64,
// so four views would share an instruction through a block-diagonal A with 3/4 zeros, which makes
// it slower per observation than mma.sync (measured 8.6 cycles per m16n8k8 per SM sub-partition,
// tools/ubench_pipes.cu) while the kernel is bound by instruction issue, not by the tensor pipe.
#pragma once

namespace ccg {

#ifndef CCG_SW_JPD
#define CCG_SW_JPD 4
#endif
constexpr int kSwD = 4;               // FP64-role warps, one per SM sub-partition
constexpr int kSwJPD = CCG_SW_JPD;    // Jacobian-role warps per FP64-role warp
constexpr int kSwJ = kSwD * kSwJPD;   // Jacobian-role warps (warps 0 .. kSwJ-1)
constexpr int kSwWarps = kSwJ + kSwD;
constexpr int kSwThreads = kSwWarps * 32;
// register budgets: the CTA is launched with kSwRegLaunch registers per thread; the Jacobian
// warpgroups give registers back (setmaxnreg.dec) and the FP64 warpgroup takes them (.inc)
constexpr int kSwRegLaunch = (65536 / kSwThreads) & ~7;
#ifndef CCG_SW_REGJ
#define CCG_SW_REGJ (CCG_SW_JPD >= 4 ? 80 : 96)
#endif
constexpr int kSwRegJ = CCG_SW_REGJ;
constexpr int kSwRegD = ((kSwRegLaunch * kSwThreads - kSwRegJ * kSwJ * 32) / (kSwD * 32)) & ~7;
static_assert(kSwRegD >= 128 && kSwRegD <= 256 && kSwRegJ % 8 == 0, "register split");
static_assert(kSwD == 4, "one setmaxnreg warpgroup of FP64 warps");

#ifndef CCG_SW_AHEAD_D
#define CCG_SW_AHEAD_D 6
#endif
#ifndef CCG_SW_AHEAD_J
#define CCG_SW_AHEAD_J 2
#endif
constexpr int kSwAheadD = CCG_SW_AHEAD_D, kSwStagesD = kSwAheadD + 2;  // chunks in flight / ring size
constexpr int kSwAheadJ = CCG_SW_AHEAD_J, kSwStagesJ = kSwAheadJ + 2;
constexpr int kSwStageD = 192;  // floats: u v X Y Z (160) | pose (24) | descriptor (4) | pad (4)
constexpr int kSwStageJ = 128;  // floats: X Y Z (96) | pose (24) | descriptor (4) | pad (4)
constexpr int kSwMaxViews = 2048;  // views per work item (per-view table in shared memory)
constexpr int kObsBlock = 160;     // floats per 32-observation block in HBM

constexpr int kTcRow = 20;            // floats per Jacobian row in the tile (16 + 4: conflict-free STS.128 and LDS.64)
constexpr int kTcTile = 64 * kTcRow;  // one 32-observation tile
#ifndef CCG_SW_TILES
#define CCG_SW_TILES 1
#endif
constexpr int kSwTiles = CCG_SW_TILES;  // tiles per Jacobian warp (2: chunk j's rows are computed while j-1 is contracted)

constexpr size_t kSwSmem = sizeof(double) * kSwWarps * kSweepPart + sizeof(int2) * kSwMaxViews +
                           sizeof(float) * (kSwD * kSwStagesD * kSwStageD + kSwJ * kSwStagesJ * kSwStageJ +
                                            kSwJ * kSwTiles * kTcTile);
static_assert(kSwSmem <= 227 * 1024, "shared memory");

__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], unsigned a0, unsigned a1, unsigned a2,
                                                unsigned a3, unsigned b0, unsigned b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
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
};

// ---------------------------------------------------------------------------------------------
// FP64 role: forward model, residual, cost, and the gradients g_u / g_v by reverse mode, one
// observation per lane.  Branch-free: an invalid lane (padding of the view's last block) sees a
// zero observation (finite arithmetic) and its residual is forced to zero, so it contributes
// nothing.  res (RES only): where this lane's residual pair goes.
// ---------------------------------------------------------------------------------------------
template <bool RES>
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
  const double xx = xn * xn, yy = yn * yn, xy = xn * yn;
  const double r2 = xx + yy, r4 = r2 * r2, r6 = r4 * r2;
  // m = 1 + k1 r2 + k2 r4 + k3 r6 and mp = dm/dr2 by Horner (no derived constants to keep in registers)
  const double h2 = fma(r2, K.k3, K.k2), h1 = fma(r2, h2, K.k1);
  const double m = fma(r2, h1, 1.0), mp = fma(r2, fma(r2, K.k3, h2), h1);
  const double a1 = xy + xy, a2 = fma(2.0, xx, r2), a3 = fma(2.0, yy, r2);
  const double xd = fma(xn, m, fma(K.p1, a1, K.p2 * a2));
  const double yd = fma(yn, m, fma(K.p2, a1, K.p1 * a3));
  double ru = fma(K.fx, xd, K.cx - dmu);
  double rv = fma(K.fy, yd, K.cy - dmv);
  ru = valid ? ru : 0.0;
  rv = valid ? rv : 0.0;
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
  const double tx = xn + xn, ty = yn + yn;
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

// ---------------------------------------------------------------------------------------------
// Jacobian role: the two float32 Jacobian rows of one observation per lane, in tile order.
// Invalid lanes produce all-zero rows.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fp32_rows(const TcConst32& K, const float (&Rf)[9], const float (&tf)[3],
                                             float X, float Y, float Z, bool valid, float4 (&ur)[4],
                                             float4 (&vr)[4]) {
  const float one = valid ? 1.0f : 0.0f;
  const float fxv = valid ? K.fx : 0.0f, fyv = valid ? K.fy : 0.0f;
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
  const float fiz = fxv * iz, giz = fyv * iz;
  // d(u,v)/d(camera point): rows (A0 A1 A2), (B0 B1 B2)
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
__device__ __forceinline__ void tc_contract(unsigned frag, int nvalid, float (&c0a)[4], float (&c0b)[4],
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

__device__ __forceinline__ void tc_store_rows(unsigned row_u, const float4 (&ur)[4], const float4 (&vr)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    sts_v4f32(row_u + 16 * i, ur[i]);
    sts_v4f32(row_u + 32 * kTcRow * 4 + 16 * i, vr[i]);
  }
}

template <bool FULL>
__device__ __forceinline__ void tc_contract(unsigned frag, int nvalid, float (&c0a)[4], float (&c0b)[4],
                                            float (&c1a)[4], float (&c1b)[4]) {
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    if (FULL || 8 * (s & 3) < nvalid) {
      const float2 lo = lds_v2f32(frag + (8 * s) * kTcRow * 4);
      const float2 hi = lds_v2f32(frag + (8 * s + 1) * kTcRow * 4);
      const unsigned a0 = __float_as_uint(lo.x), a1 = __float_as_uint(lo.y);
      const unsigned a2 = __float_as_uint(hi.x), a3 = __float_as_uint(hi.y);
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
// Prefetch cursor of a role warp over its views (v0, v0 + STEP, ... < v_end, relative to the work
// item's first view) and their 32-observation chunks.  An empty view yields one chunk with zero
// valid observations, so its (all-zero) blocks are still written.  All fields are warp-uniform.
// ---------------------------------------------------------------------------------------------
template <int STEP>
struct SwCursor {
  unsigned vm;  // shared address of the item's per-view table (first block, observations)
  int v, v_end;
  int blk, rem;  // block of the cursor's chunk; observations of the view from this chunk on
  bool first, ok;
  __device__ __forceinline__ void load() {
    const int2 m = lds_v2s32(vm + 8u * (unsigned)v);
    blk = m.x;
    rem = m.y;
    first = true;
  }
  __device__ __forceinline__ void init(unsigned vm_, int v0, int v_end_) {
    vm = vm_;
    v = v0;
    v_end = v_end_;
    blk = rem = 0;
    first = false;
    ok = v < v_end;
    if (ok) load();
  }
  __device__ __forceinline__ void next() {
    if (rem > 32) {
      rem -= 32;
      ++blk;
      first = false;
    } else {
      v += STEP;
      ok = v < v_end;
      if (ok) load();
    }
  }
};

// chunk descriptor, one int: view (relative to the item) | valid observations << 12 | first << 20 | last << 21;
// -1 = end of the warp's stream
constexpr int kDescFirst = 1 << 20, kDescLast = 1 << 21;

// Request the cursor's chunk into the stage at shared address `st` and advance the cursor: the
// observation block (the X Y Z part only for the Jacobian role), the view's pose when the chunk
// opens a view, and the descriptor.
//   default: one bulk copy (TMA, cp.async.bulk) per block / pose issued by lane 0, completing on the
//     stage's mbarrier (expect_tx bytes); the descriptor store precedes the arrive, so a lane that
//     has seen the phase complete sees the descriptor too.
//   CCG_SW_LDGSTS: 16-byte cp.async by the lanes, one commit group per call (wait_group counts chunks).
// src = obs_blk (+ 256 for the Jacobian role), rt = rt(item's first view) (bytes).
template <int NARR, int STEP>
__device__ __forceinline__ void sw_issue(const char* src, const char* rt, unsigned st, SwCursor<STEP>& cur, int lane) {
  constexpr unsigned kPoseOff = (NARR == 5 ? 160 : 96) * 4, kDescOff = kPoseOff + 96, kBarOff = kDescOff + 8;
  constexpr unsigned kBytes = NARR == 5 ? 640 : 384;
#ifdef CCG_SW_LDGSTS
  if (cur.ok) {
    const int nvalid = min(32, cur.rem);  // 0 for an empty view
    const unsigned dst = st + 16u * (unsigned)lane;
    if (nvalid > 0) {
      const char* s16 = src + (size_t)(unsigned)cur.blk * (kObsBlock * 4) + 16 * lane;
      if (NARR == 5) {
        cp_async16(dst, s16);
        if (lane < 8) cp_async16(dst + 512, s16 + 512);
      } else {
        if (lane < 24) cp_async16(dst, s16);
      }
    }
    if (cur.first && lane < 6) cp_async16(dst + kPoseOff, rt + (size_t)(unsigned)cur.v * 96 + 16 * lane);
    if (lane == 0)
      sts_s32(st + kDescOff, cur.v | (nvalid << 12) | (cur.first ? kDescFirst : 0) | (cur.rem <= 32 ? kDescLast : 0));
    cur.next();
  } else if (lane == 0) {
    sts_s32(st + kDescOff, -1);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
#else
  if (lane == 0) {
    const unsigned bar = st + kBarOff;
    if (cur.ok) {
      const int nvalid = min(32, cur.rem);  // 0 for an empty view
      sts_s32(st + kDescOff, cur.v | (nvalid << 12) | (cur.first ? kDescFirst : 0) | (cur.rem <= 32 ? kDescLast : 0));
      const unsigned tx = (nvalid > 0 ? kBytes : 0u) + (cur.first ? 96u : 0u);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(tx) : "memory");
      if (nvalid > 0)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(st),
                     "l"(src + (size_t)(unsigned)cur.blk * (kObsBlock * 4)), "n"(kBytes), "r"(bar)
                     : "memory");
      if (cur.first)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 96, [%2];" ::"r"(
                         st + kPoseOff),
                     "l"(rt + (size_t)(unsigned)cur.v * 96), "r"(bar)
                     : "memory");
    } else {
      sts_s32(st + kDescOff, -1);
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
    }
  }
  if (cur.ok) cur.next();
#endif
}

// wait until the chunk in the stage at `st` has landed (AHEAD = chunks allowed to stay in flight)
template <int NARR, int AHEAD>
__device__ __forceinline__ void sw_wait(unsigned st, unsigned phase) {
#ifdef CCG_SW_LDGSTS
  asm volatile("cp.async.wait_group %0;" ::"n"(AHEAD) : "memory");
#else
  constexpr unsigned kBarOff = (NARR == 5 ? 160 : 96) * 4 + 96 + 8;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(st + kBarOff),
      "r"(phase)
      : "memory");
#endif
}

// per-warp setup of a stage ring (mbarriers, one arrival per phase: lane 0's)
template <int NARR>
__device__ __forceinline__ void sw_ring_init(unsigned stages, int n_stages, int lane) {
#ifndef CCG_SW_LDGSTS
  constexpr unsigned kStride = (NARR == 5 ? kSwStageD : kSwStageJ) * 4, kBarOff = (NARR == 5 ? 160 : 96) * 4 + 96 + 8;
  if (lane < n_stages)
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(stages + (unsigned)lane * kStride + kBarOff) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
#endif
}

// keeps a lane-derived value in a register (ptxas otherwise rebuilds it from SR_TID at every use)
__device__ __forceinline__ int opaque(int x) {
  asm volatile("mov.u32 %0, %0;" : "+r"(x));
  return x;
}

// ---------------------------------------------------------------------------------------------
// Candidate poses, one thread per view: pose (+) delta_v with delta_v = -z - Y du (the
// back-substitution of the reduced solve), written as the candidate pose and as the rotation
// matrix + translation the sweep consumes; per-CTA partial sums of the model terms g'd and d'Dd.
// Grid: schur_split CTAs per work item (the elimination kernel's split).
// ---------------------------------------------------------------------------------------------
constexpr int kStepThreads = 128;
__global__ void __launch_bounds__(kStepThreads) k_intr_step(IntrDev d) {
  __shared__ double s_du[9];
  __shared__ double s_red[2][kStepThreads / 32];
  __shared__ int s_meta[2];
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
    s_meta[1] = (ps.lm.status == 0 && ps.lm.solve_ok) ? 1 : 0;
  }
  if (threadIdx.x < 9) s_du[threadIdx.x] = ps.du[threadIdx.x];
  __syncthreads();
  if (!s_meta[1]) return;
  const int par = s_meta[0];
  double gTd = 0.0, dDd = 0.0;
  for (int64_t v = it.v0 + threadIdx.x; v < it.v1; v += kStepThreads) {
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
  __syncthreads();
  if (threadIdx.x < 2) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kStepThreads / 32; ++w) s += s_red[threadIdx.x][w];
    d.step_part[(size_t)blockIdx.x * 2 + threadIdx.x] = s;
  }
}

// ---------------------------------------------------------------------------------------------
// The FP64 role over the views va, va + 1, ... < vb of the item: residuals, cost, g_u (into the
// warp's partial) and g_v (into the view blocks).  `stages` = shared address of the warp's ring.
// ---------------------------------------------------------------------------------------------
template <bool RES>
__device__ __forceinline__ void sw_run_fp64(const IntrDev& d, int item_v0, unsigned vm, unsigned stages,
                                            const double* s_k, int par, int va, int vb, int lane, double* upart,
                                            double* __restrict__ res_out) {
  TcConst64 K;
  K.fx = s_k[0]; K.fy = s_k[1]; K.cx = s_k[2]; K.cy = s_k[3];
  K.k1 = s_k[4]; K.k2 = s_k[5]; K.p1 = s_k[6]; K.p2 = s_k[7]; K.k3 = s_k[8];
  double gu[9], cost = 0.0;
#pragma unroll
  for (int i = 0; i < 9; ++i) gu[i] = 0.0;
  lane = opaque(lane);
  const char* src_lane = reinterpret_cast<const char*>(d.obs_blk);
  const char* rt_lane = reinterpret_cast<const char*>(d.rt + (size_t)item_v0 * 12);
  sw_ring_init<5>(stages, kSwStagesD, lane);
  double* gv_out = d.blk[par ^ 1] + (size_t)item_v0 * kBlk + 21 + lane;
  SwCursor<1> cur;
  cur.init(vm, va, vb);
#pragma unroll
  for (int i = 0; i <= kSwAheadD; ++i) sw_issue<5, 1>(src_lane, rt_lane, stages + i * (kSwStageD * 4), cur, lane);
  double R[9], t[3], gv[6];
#pragma unroll
  for (int i = 0; i < 9; ++i) R[i] = 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i) t[i] = 0.0;
#pragma unroll
  for (int i = 0; i < 6; ++i) gv[i] = 0.0;
  const unsigned ring_end = stages + kSwStagesD * (kSwStageD * 4);
  unsigned st = stages, st_prev = ring_end - kSwStageD * 4;  // stage being consumed / consumed one iteration ago
  const unsigned lane4 = 4u * (unsigned)lane;
  unsigned phase = 0;  // parity of the ring pass (mbarrier phase of the stage being consumed)
  int chunk = 0;       // chunk index inside the view (residual output only)
  for (;;) {
    sw_wait<5, kSwAheadD>(st, phase);
    __syncwarp();  // a lane reads what other lanes copied; the previous stage is no longer being read
    const int desc = lds_s32(st + 184 * 4);
    // refill the stage consumed one iteration ago (the ring has one stage more than chunks in flight)
    sw_issue<5, 1>(src_lane, rt_lane, st_prev, cur, lane);
    if (desc < 0) break;
    const int nvalid = (desc >> 12) & 63;
    if (desc & kDescFirst) {
      const double2 a = lds_v2f64(st + 640), b = lds_v2f64(st + 656), c = lds_v2f64(st + 672),
                    e = lds_v2f64(st + 688), f = lds_v2f64(st + 704), h = lds_v2f64(st + 720);
      R[0] = a.x; R[1] = a.y; R[2] = b.x; R[3] = b.y; R[4] = c.x; R[5] = c.y;
      R[6] = e.x; R[7] = e.y; R[8] = f.x; t[0] = f.y; t[1] = h.x; t[2] = h.y;
#pragma unroll
      for (int i = 0; i < 6; ++i) gv[i] = 0.0;
      chunk = 0;
    }
    if (nvalid > 0) {
      const unsigned a = st + lane4;
      const float mu = lds_f32(a), mv = lds_f32(a + 128), X = lds_f32(a + 256), Y = lds_f32(a + 384),
                  Z = lds_f32(a + 512);
      double* res = nullptr;
      if (RES) res = res_out + 2 * (d.view_off[item_v0 + (desc & 4095)] - d.obs_base + 32 * chunk + lane);
      tc_fp64_observation<RES>(K, R, t, mu, mv, X, Y, Z, lane < nvalid, gu, gv, cost, res);
    }
    ++chunk;
    if (desc & kDescLast) {
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
      if (lane < 6) gv_out[(size_t)(desc & 4095) * kBlk] = out;
    }
    st_prev = st;
    st += kSwStageD * 4;
    if (st == ring_end) {
      st = stages;
      phase ^= 1u;
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
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
// The Jacobian + tensor role over the views va, va + kSwJPD, ... < vb of the item: V, W (into the
// view blocks) and U (into the warp's partial).  `stages`, `tiles` = shared addresses.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void sw_run_jac(const IntrDev& d, int item_v0, unsigned vm, unsigned stages, unsigned tiles,
                                           const double* s_k, int par, int va, int vb, int lane, double* upart) {
  const int g = lane >> 2, tig = lane & 3;
  TcConst32 K;
  K.fx = (float)s_k[0]; K.fy = (float)s_k[1];
  K.k1 = (float)s_k[4]; K.k2 = (float)s_k[5]; K.p1 = (float)s_k[6]; K.p2 = (float)s_k[7]; K.k3 = (float)s_k[8];
  lane = opaque(lane);
  const char* src_lane = reinterpret_cast<const char*>(d.obs_blk) + 256;
  const char* rt_lane = reinterpret_cast<const char*>(d.rt + (size_t)item_v0 * 12);
  sw_ring_init<3>(stages, kSwStagesJ, lane);
  double* blk0 = d.blk[par ^ 1] + (size_t)item_v0 * kBlk;
  SwCursor<kSwJPD> cur;
  cur.init(vm, va, vb);
#pragma unroll
  for (int i = 0; i <= kSwAheadJ; ++i) sw_issue<3, kSwJPD>(src_lane, rt_lane, stages + i * (kSwStageJ * 4), cur, lane);
  float Rf[9], tf[3];
#pragma unroll
  for (int i = 0; i < 9; ++i) Rf[i] = 0.0f;
#pragma unroll
  for (int i = 0; i < 3; ++i) tf[i] = 0.0f;
  // product columns 0..7 (c0*) and 8..15 (c1*), two accumulator sets each
  float c0a[4] = {0.f, 0.f, 0.f, 0.f}, c0b[4] = {0.f, 0.f, 0.f, 0.f};
  float c1a[4] = {0.f, 0.f, 0.f, 0.f}, c1b[4] = {0.f, 0.f, 0.f, 0.f};
  const unsigned ring_end = stages + kSwStagesJ * (kSwStageJ * 4);
  unsigned st = stages, st_prev = ring_end - kSwStageJ * 4;
  const unsigned lane4 = 4u * (unsigned)lane;
  const unsigned row_u = tiles + (unsigned)lane * (kTcRow * 4);               // this lane's u row (v row: + 32 rows)
  const unsigned frag = tiles + (unsigned)(2 * tig * kTcRow + 2 * g) * 4;  // this lane's fragment of k-step 0
  unsigned tsel = 0;   // byte offset of the tile in use (kSwTiles == 2: alternates)
  unsigned phase = 0;  // parity of the ring pass (mbarrier phase of the stage being consumed)
  for (;;) {
    sw_wait<3, kSwAheadJ>(st, phase);
    __syncwarp();  // stage visible to every lane; the tile written two chunks ago has been read
    const int desc = lds_s32(st + 120 * 4);
    sw_issue<3, kSwJPD>(src_lane, rt_lane, st_prev, cur, lane);
    if (desc < 0) break;
    const int nvalid = (desc >> 12) & 63;
    if (desc & kDescFirst) {
      const double2 a = lds_v2f64(st + 384), b = lds_v2f64(st + 400), c = lds_v2f64(st + 416),
                    e = lds_v2f64(st + 432), f = lds_v2f64(st + 448), h = lds_v2f64(st + 464);
      Rf[0] = (float)a.x; Rf[1] = (float)a.y; Rf[2] = (float)b.x; Rf[3] = (float)b.y; Rf[4] = (float)c.x;
      Rf[5] = (float)c.y; Rf[6] = (float)e.x; Rf[7] = (float)e.y; Rf[8] = (float)f.x;
      tf[0] = (float)f.y; tf[1] = (float)h.x; tf[2] = (float)h.y;
#pragma unroll
      for (int i = 0; i < 4; ++i) c0a[i] = c0b[i] = c1a[i] = c1b[i] = 0.0f;
    }
    if (nvalid > 0) {
      const unsigned a = st + lane4;
      const float X = lds_f32(a), Y = lds_f32(a + 128), Z = lds_f32(a + 256);
      float4 ur[4], vr[4];
      tc_fp32_rows(K, Rf, tf, X, Y, Z, lane < nvalid, ur, vr);
      tc_store_rows(row_u + tsel, ur, vr);
      __syncwarp();
      if (nvalid == 32) tc_contract<true>(frag + tsel, 32, c0a, c0b, c1a, c1b);
      else tc_contract<false>(frag + tsel, nvalid, c0a, c0b, c1a, c1b);
      if (kSwTiles > 1) tsel ^= kTcTile * 4;
    }
    if (desc & kDescLast) {
      // per-view epilogue: V, W blocks and the view's contribution to U
      double* blk = blk0 + (size_t)(desc & 4095) * kBlk;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int nt = i >> 2, e = i & 3;
        const int row = g + ((e & 2) ? 8 : 0), col = nt * 8 + 2 * tig + (e & 1);
        const float val = nt == 0 ? c0a[e] + c0b[e] : c1a[e] + c1b[e];
        const int bs = tc_blk_slot(row, col), us = tc_u_slot(row, col);
        if (bs >= 0) blk[bs] = (double)val;
        if (us >= 0) upart[us] += (double)val;  // one owner lane per slot within the warp
      }
    }
    st_prev = st;
    st += kSwStageJ * 4;
    if (st == ring_end) {
      st = stages;
      phase ^= 1u;
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// The sweep.  One CTA per work item; see the top of the file for the role layout.
// ---------------------------------------------------------------------------------------------
template <bool RES>
__global__ void __launch_bounds__(kSwThreads, 1) k_intr_sweep_ws(IntrDev d, double* __restrict__ res_out) {
  extern __shared__ __align__(16) unsigned char s_dyn[];
  double(*s_part)[kSweepPart] = reinterpret_cast<double(*)[kSweepPart]>(s_dyn);
  int2* s_vm = reinterpret_cast<int2*>(s_dyn + sizeof(double) * kSwWarps * kSweepPart);
  float* s_stage_d = reinterpret_cast<float*>(s_vm + kSwMaxViews);
  float* s_stage_j = s_stage_d + kSwD * kSwStagesD * kSwStageD;
  float* s_tiles = s_stage_j + kSwJ * kSwStagesJ * kSwStageJ;
  __shared__ double s_k[9];
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
  const int nv = it.v1 - it.v0;  // <= kSwMaxViews (build_problem)
  if (threadIdx.x < 9) s_k[threadIdx.x] = ps.intr[par ^ 1][threadIdx.x];
  for (int i = threadIdx.x; i < kSwWarps * kSweepPart; i += kSwThreads) (&s_part[0][0])[i] = 0.0;
  for (int i = threadIdx.x; i < nv; i += kSwThreads) s_vm[i] = d.vmeta[it.v0 + i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* upart = s_part[warp];
  const unsigned vm = smem_u32(s_vm);
  if (warp >= kSwJ) {
    // =========================== FP64 role ===========================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kSwRegD));
    const int grp = warp - kSwJ;
    const int va = (int)((int64_t)nv * grp / kSwD), vb = (int)((int64_t)nv * (grp + 1) / kSwD);
    sw_run_fp64<RES>(d, it.v0, vm, smem_u32(s_stage_d + grp * kSwStagesD * kSwStageD), s_k, par, va, vb, lane, upart,
                     res_out);
  } else {
    // =========================== Jacobian + tensor role ===========================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kSwRegJ));
    const int grp = warp % kSwD, j = warp / kSwD;
    const int va = (int)((int64_t)nv * grp / kSwD), vb = (int)((int64_t)nv * (grp + 1) / kSwD);
    sw_run_jac(d, it.v0, vm, smem_u32(s_stage_j + warp * kSwStagesJ * kSwStageJ),
               smem_u32(s_tiles + warp * kSwTiles * kTcTile), s_k, par, va + j, vb, lane, upart);
  }
  __syncthreads();
  if (threadIdx.x < kSweepPart) {
    // the model terms g'd, d'Dd of this item come from the candidate-pose kernel
    double s = 0.0;
    if (threadIdx.x == 7 || threadIdx.x == 8) {
      const int c = threadIdx.x - 7;
      for (int i = 0; i < d.schur_split; ++i) s += d.step_part[((size_t)blockIdx.x * d.schur_split + i) * 2 + c];
    } else {
#pragma unroll
      for (int w = 0; w < kSwWarps; ++w) s += s_part[w][threadIdx.x];
    }
    d.sweep_part[(size_t)blockIdx.x * kSweepPart + threadIdx.x] = s;
  }
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
