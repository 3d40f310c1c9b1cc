// Warp-specialised tensor-core single-camera sweep (included by intr.cu) - the default sweep.
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
// Operand precision: the raw float32 rows are fed to the TF32 unit (10 explicit mantissa bits).
// tools/h_precision_study.py shows that the exact-Schur LM takes the same iterations with such
// blocks and lands within 1e-10 of the all-FP64 solve (oracle experiment mode 311).
// Why legacy mma.sync and not tcgen05: the tile is 16 x 16 per VIEW; tcgen05's smallest M is 64,
// so four views would share an instruction through a block-diagonal A with 3/4 zeros, which makes
// it slower per observation than mma.sync (measured 8.6 cycles per m16n8k8 per SM sub-partition,
// tools/ubench_pipes.cu) while the kernel is bound by instruction issue, not by the tensor pipe.
#pragma once

namespace ccg {

#ifndef CCG_TC_WARPS
#define CCG_TC_WARPS 16
#endif
constexpr int kTcWarps = CCG_TC_WARPS;  // warps per CTA, one CTA per SM
#ifndef CCG_TC_D
#define CCG_TC_D (kTcWarps / 2)
#endif
constexpr int kTcD = CCG_TC_D;     // FP64-role warps (the first kTcD warps of the CTA)
constexpr int kTcJ = kTcWarps - kTcD;  // Jacobian/tensor-role warps
static_assert(kTcD == kTcJ, "a Jacobian warp shares its view range with one FP64 warp");
#ifndef CCG_TC_TAIL
#define CCG_TC_TAIL 10
#endif
constexpr int kTcTail = CCG_TC_TAIL;  // 64ths of a range's FP64 work done by the Jacobian warp afterwards
constexpr int kTcRow = 20;         // floats per Jacobian row in the tile (16 + 4: conflict-free STS.128 and LDS.64)
constexpr int kTcTile = 64 * kTcRow;  // one 32-observation tile; every warp owns two (double buffer)
constexpr int kTcSlots = 256;         // observations the circular staging buffer of a warp holds (8 granules of 32)
#ifndef CCG_TC_AHEAD
#define CCG_TC_AHEAD 2
#endif
constexpr int kTcAhead = CCG_TC_AHEAD;  // granules requested beyond the one being consumed (ring holds 8)
constexpr int kTcBatch = 16;          // views whose poses and offsets a warp stages at a time
// floats per warp: u v X Y Z circular buffers | 16 poses (12 doubles each) | 17 view offsets (int)
constexpr int kTcRing = 5 * kTcSlots + kTcBatch * 24 + 32;
// dynamic shared memory: per warp a 64-double partial and a 3-stage observation ring; per
// Jacobian warp two tiles
constexpr size_t kTcSmem = sizeof(double) * kTcWarps * kSweepPart +
                           sizeof(float) * (kTcWarps * kTcRing + kTcJ * 2 * kTcTile);

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

struct ObsIn {
  float mu, mv, X, Y, Z;
  int64_t o;
  bool valid;
};

// ---------------------------------------------------------------------------------------------
// FP64 role: forward model, residual, cost, and the gradients g_u / g_v by reverse mode, one
// observation per lane.  Branch-free: an invalid lane (past the end of the view) sees a
// zero-filled observation and its residual is forced to zero, so it contributes nothing.
// ---------------------------------------------------------------------------------------------
template <bool RES>
__device__ __forceinline__ void tc_fp64_observation(const IntrDev& d, const TcConst64& K, const double (&R)[9],
                                                    const double (&t)[3], const ObsIn& cur, double (&gu)[9],
                                                    double (&gv)[6], double& cost, double* __restrict__ res_out) {
  const double X = (double)cur.X, Y = (double)cur.Y, Z = (double)cur.Z;
  const double dmu = (double)cur.mu, dmv = (double)cur.mv;
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
  ru = cur.valid ? ru : 0.0;
  rv = cur.valid ? rv : 0.0;
  if (RES) {
    if (cur.valid) {
      res_out[2 * (cur.o - d.obs_base)] = ru;
      res_out[2 * (cur.o - d.obs_base) + 1] = rv;
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

// rows 0..31 of a tile = u rows of the 32 observations, rows 32..63 = their v rows
__device__ __forceinline__ void tc_store_rows(float* tile, int lane, const float4 (&ur)[4], const float4 (&vr)[4]) {
  float4* tu4 = reinterpret_cast<float4*>(tile + lane * kTcRow);
  float4* tv4 = reinterpret_cast<float4*>(tile + (32 + lane) * kTcRow);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    tu4[i] = ur[i];
    tv4[i] = vr[i];
  }
}

// [V W; W' U] += rows' rows over one tile.  k-step s contracts tile rows 8s..8s+7 (k = tig -> row
// 8s + 2 tig, k = tig + 4 -> the next row); two accumulator sets per column block halve the
// depth of the dependent HMMA chain.  FULL: all 32 observations are valid.
template <bool FULL>
__device__ __forceinline__ void tc_contract(const float* tile, int g, int tig, int nvalid, float (&c0a)[4],
                                            float (&c0b)[4], float (&c1a)[4], float (&c1b)[4]) {
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    if (FULL || 8 * (s & 3) < nvalid) {
      const float2 lo = *reinterpret_cast<const float2*>(tile + (8 * s + 2 * tig) * kTcRow + 2 * g);
      const float2 hi = *reinterpret_cast<const float2*>(tile + (8 * s + 2 * tig + 1) * kTcRow + 2 * g);
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
// Observation stream of a role warp.  A warp owns a CONTIGUOUS range of views, so its observations
// are one contiguous stream in the SoA arrays.  The stream is fetched in 32-observation granules
// (cp.async, LDGSTS: no registers are held while the loads are in flight - ptxas sank plain
// prefetch loads next to their use) into a circular buffer of 256 observations per array, indexed
// by the position in the stream modulo 256 - independent of where views begin and end.  The
// consumer works on view-aligned chunks; before a chunk it makes sure the stream is requested
// kTcAhead granules beyond the chunk's last observation and that the granule holding that
// observation has landed (one commit group per granule, groups complete in order).  The HBM
// latency is therefore exposed once per work item and warp, not once per view.
// NARR = 5: u v X Y Z (FP64 role), NARR = 3: X Y Z (Jacobian role; the other role's copy of the
// same lines is an L1/L2 hit, HBM traffic stays 20 B per observation).
// ---------------------------------------------------------------------------------------------
template <int NARR>
struct TcStream {
  const IntrDev& d;
  float* ring;
  int64_t base;   // first observation of the warp's range
  int n;          // observations in the range
  int g_issued;   // granules requested so far
  int lane;
  __device__ __forceinline__ TcStream(const IntrDev& d_, float* ring_, int64_t o_first, int64_t o_last, int lane_)
      : d(d_), ring(ring_), base(o_first), n((int)(o_last - o_first)), g_issued(0), lane(lane_) {
    // the head of the stream is requested at once: it travels together with the first batch of
    // poses and offsets (one round trip instead of two before the first observation is touched)
#pragma unroll
    for (int i = 0; i <= kTcAhead; ++i) issue();
  }
  __device__ __forceinline__ void issue() {
    const int rel = 32 * g_issued + lane;
    const bool valid = rel < n;
    const int64_t o = base + (valid ? rel : 0);
    const int sz = valid ? 4 : 0;
    const unsigned dst = (unsigned)__cvta_generic_to_shared(ring + (rel & (kTcSlots - 1)));
    if (NARR == 5) {
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(d.u + o), "r"(sz) : "memory");
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst + 4 * kTcSlots), "l"(d.v + o), "r"(sz) : "memory");
    }
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst + 8 * kTcSlots), "l"(d.X + o), "r"(sz) : "memory");
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst + 12 * kTcSlots), "l"(d.Y + o), "r"(sz) : "memory");
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst + 16 * kTcSlots), "l"(d.Z + o), "r"(sz) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
    ++g_issued;
  }
  // the chunk whose last observation has stream position rel_last is about to be consumed
  __device__ __forceinline__ void ensure(int rel_last) {
    const int g_need = rel_last >> 5;
    while (g_issued <= g_need + kTcAhead) issue();  // warp-uniform
    asm volatile("cp.async.wait_group %0;" ::"n"(kTcAhead) : "memory");
    __syncwarp();  // a lane reads what other lanes copied
  }
  __device__ __forceinline__ void take(ObsIn& L, int rel, bool valid) const {
    const int slot = rel & (kTcSlots - 1);
    L.valid = valid;
    L.o = base + rel;
    if (NARR == 5) {
      L.mu = ring[slot];
      L.mv = ring[kTcSlots + slot];
    }
    L.X = ring[2 * kTcSlots + slot];
    L.Y = ring[3 * kTcSlots + slot];
    L.Z = ring[4 * kTcSlots + slot];
  }
};

// Poses (rotation + translation, 12 doubles) and stream-relative observation offsets of the next
// kTcBatch views of the warp, staged in shared memory with one blocking round trip per batch.
__device__ __forceinline__ void tc_stage_batch(const IntrDev& d, float* ring, int vb, int nb, int64_t base, int lane) {
  double* s_pose = reinterpret_cast<double*>(ring + 5 * kTcSlots);
  int* s_off = reinterpret_cast<int*>(ring + 5 * kTcSlots + kTcBatch * 24);
  __syncwarp();  // the previous batch is no longer being read
  // 16 poses = 1536 B = 3 x 16 B per lane
  const char* src = reinterpret_cast<const char*>(d.rt + (size_t)vb * 12);
  const unsigned dst = (unsigned)__cvta_generic_to_shared(s_pose);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int byte = (k * 32 + lane) * 16;
    if (byte < nb * 96)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + byte), "l"(src + byte) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  if (lane <= nb) s_off[lane] = (int)(d.view_off[vb + lane] - base);
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncwarp();
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
// The FP64 role over the contiguous view range [vw0, vw1): residuals, cost, g_u (into the warp's
// partial) and g_v (into the view blocks).
// ---------------------------------------------------------------------------------------------
template <bool RES>
__device__ __forceinline__ void tc_run_fp64_role(const IntrDev& d, float* ring, const double* s_k, int par, int vw0,
                                                 int vw1, int lane, double* upart, double* __restrict__ res_out) {
  {
    TcConst64 K;
    K.fx = s_k[0]; K.fy = s_k[1]; K.cx = s_k[2]; K.cy = s_k[3];
    K.k1 = s_k[4]; K.k2 = s_k[5]; K.p1 = s_k[6]; K.p2 = s_k[7]; K.k3 = s_k[8];
    double gu[9], cost = 0.0;
#pragma unroll
    for (int i = 0; i < 9; ++i) gu[i] = 0.0;
    const double* s_pose = reinterpret_cast<const double*>(ring + 5 * kTcSlots);
    const int* s_off = reinterpret_cast<const int*>(ring + 5 * kTcSlots + kTcBatch * 24);
    TcStream<5> strm(d, ring, d.view_off[vw0], d.view_off[vw1], lane);
    for (int vb = vw0; vb < vw1; vb += kTcBatch) {
     const int nb = min(kTcBatch, vw1 - vb);
     tc_stage_batch(d, ring, vb, nb, strm.base, lane);
     for (int vi = 0; vi < nb; ++vi) {
      const int v = vb + vi;
      const int r0 = s_off[vi], r1 = s_off[vi + 1];
      double R[9], t[3], gv[6];
      {
        const double2* rt = reinterpret_cast<const double2*>(s_pose + vi * 12);
        const double2 a = rt[0], b = rt[1], c = rt[2], e = rt[3], f = rt[4], h = rt[5];
        R[0] = a.x; R[1] = a.y; R[2] = b.x; R[3] = b.y; R[4] = c.x; R[5] = c.y;
        R[6] = e.x; R[7] = e.y; R[8] = f.x; t[0] = f.y; t[1] = h.x; t[2] = h.y;
      }
#pragma unroll
      for (int i = 0; i < 6; ++i) gv[i] = 0.0;
      for (int rb = r0; rb < r1; rb += 32) {
        strm.ensure(min(rb + 31, r1 - 1));
        ObsIn cur;
        strm.take(cur, min(rb + lane, r1 - 1), rb + lane < r1);
        tc_fp64_observation<RES>(d, K, R, t, cur, gu, gv, cost, res_out);
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
      if (lane < 6) d.blk[par ^ 1][(size_t)v * kBlk + 21 + lane] = out;
     }
    }
    // lane-private sums -> the team's partial: [0] sum r^2 | [16..24] g_u
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
}

// ---------------------------------------------------------------------------------------------
// The sweep.  One CTA per work item; the first kTcD warps run the FP64 role, the others the
// Jacobian role, so the FP64 pipe and the FP32 + tensor pipes of every SM sub-partition are fed
// by different warps at the same time.  Each role walks ALL views of the item on its own
// (view v goes to role-warp v mod role size: static, so sums are bit-reproducible); the roles
// share nothing, so there is no synchronisation inside the loops, and the second reader of an
// observation line hits L1/L2 (HBM traffic stays 20 B per observation).  The split 7 : 9 balances
// the measured role times (FP64 role 0.78 of the Jacobian role per observation).
// ---------------------------------------------------------------------------------------------
template <bool RES>
__global__ void __launch_bounds__(kTcWarps * 32, 1) k_intr_sweep_tc(IntrDev d, double* __restrict__ res_out) {
  extern __shared__ __align__(16) unsigned char s_dyn[];
  double(*s_part)[kSweepPart] = reinterpret_cast<double(*)[kSweepPart]>(s_dyn);
  float* s_rings = reinterpret_cast<float*>(s_dyn + sizeof(double) * kTcWarps * kSweepPart);
  float* s_tiles = s_rings + kTcWarps * kTcRing;
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
  if (threadIdx.x < 9) s_k[threadIdx.x] = ps.intr[par ^ 1][threadIdx.x];
  for (int i = threadIdx.x; i < kTcWarps * kSweepPart; i += blockDim.x) (&s_part[0][0])[i] = 0.0;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool jac_role = warp >= kTcD;
  float* ring = s_rings + warp * kTcRing;
  double* upart = s_part[warp];

  // The FP64 role is the slower one when both run (the Jacobian warps used to wait ~12 % of the
  // kernel at the final barrier), so a Jacobian warp also takes the last kTcTail / 64 of its partner's
  // FP64 work once its own range is done.  The split is static: sums stay bit-reproducible.
  const int nvw_item = it.v1 - it.v0, rw = jac_role ? warp - kTcD : warp;
  const int vw0 = it.v0 + (int)((int64_t)nvw_item * rw / kTcD), vw1 = it.v0 + (int)((int64_t)nvw_item * (rw + 1) / kTcD);
  const int vsplit = vw0 + ((vw1 - vw0) * (64 - kTcTail) + 32) / 64;
  if (!jac_role) {
    // =========================== FP64 role ===========================
    tc_run_fp64_role<RES>(d, ring, s_k, par, vw0, vsplit, lane, upart, res_out);
  } else {
    // =========================== Jacobian + tensor role ===========================
    const int g = lane >> 2, tig = lane & 3;
    float* tiles = s_tiles + (warp - kTcD) * 2 * kTcTile;
    TcConst32 K;
    K.fx = (float)s_k[0]; K.fy = (float)s_k[1];
    K.k1 = (float)s_k[4]; K.k2 = (float)s_k[5]; K.p1 = (float)s_k[6]; K.p2 = (float)s_k[7]; K.k3 = (float)s_k[8];
    const double* s_pose = reinterpret_cast<const double*>(ring + 5 * kTcSlots);
    const int* s_off = reinterpret_cast<const int*>(ring + 5 * kTcSlots + kTcBatch * 24);
    TcStream<3> strm(d, ring, d.view_off[vw0], d.view_off[vw1], lane);
    for (int vb = vw0; vb < vw1; vb += kTcBatch) {
     const int nb = min(kTcBatch, vw1 - vb);
     tc_stage_batch(d, ring, vb, nb, strm.base, lane);
     for (int vi = 0; vi < nb; ++vi) {
      const int v = vb + vi;
      const int r0 = s_off[vi], r1 = s_off[vi + 1];
      float Rf[9], tf[3];
      {
        const double2* rt = reinterpret_cast<const double2*>(s_pose + vi * 12);
        const double2 a = rt[0], b = rt[1], c = rt[2], e = rt[3], f = rt[4], h = rt[5];
        Rf[0] = (float)a.x; Rf[1] = (float)a.y; Rf[2] = (float)b.x; Rf[3] = (float)b.y; Rf[4] = (float)c.x;
        Rf[5] = (float)c.y; Rf[6] = (float)e.x; Rf[7] = (float)e.y; Rf[8] = (float)f.x;
        tf[0] = (float)f.y; tf[1] = (float)h.x; tf[2] = (float)h.y;
      }
      // product columns 0..7 (c0*) and 8..15 (c1*), two accumulator sets each
      float c0a[4] = {0.f, 0.f, 0.f, 0.f}, c0b[4] = {0.f, 0.f, 0.f, 0.f};
      float c1a[4] = {0.f, 0.f, 0.f, 0.f}, c1b[4] = {0.f, 0.f, 0.f, 0.f};
      if (r1 > r0) {
        // software pipeline over 32-observation chunks: while chunk j's rows are being computed,
        // chunk j-1's tile is contracted (LDS + tensor pipe) and the stream runs two granules ahead
        ObsIn cur;
        float4 ur[4], vr[4];
        strm.ensure(min(r0 + 31, r1 - 1));
        strm.take(cur, min(r0 + lane, r1 - 1), r0 + lane < r1);
        tc_fp32_rows(K, Rf, tf, cur.X, cur.Y, cur.Z, cur.valid, ur, vr);
        tc_store_rows(tiles, lane, ur, vr);
        __syncwarp();
        int buf = 0;  // tile that holds chunk j-1
        int rb = r0 + 32;
        for (; rb < r1; rb += 32) {
          strm.ensure(min(rb + 31, r1 - 1));
          strm.take(cur, min(rb + lane, r1 - 1), rb + lane < r1);
          tc_contract<true>(tiles + buf * kTcTile, g, tig, 32, c0a, c0b, c1a, c1b);
          tc_fp32_rows(K, Rf, tf, cur.X, cur.Y, cur.Z, cur.valid, ur, vr);
          buf ^= 1;
          tc_store_rows(tiles + buf * kTcTile, lane, ur, vr);
          __syncwarp();
        }
        tc_contract<false>(tiles + buf * kTcTile, g, tig, r1 - (rb - 32), c0a, c0b, c1a, c1b);
      }
      __syncwarp();
      // per-view epilogue: V, W blocks and the view's contribution to U
      double* blk = d.blk[par ^ 1] + (size_t)v * kBlk;
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
    }
    // the tail of the partner's FP64 work (the ring is free again: every granule has been consumed)
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    tc_run_fp64_role<RES>(d, ring, s_k, par, vsplit, vw1, lane, upart, res_out);
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
      for (int w = 0; w < kTcWarps; ++w) s += s_part[w][threadIdx.x];
    }
    d.sweep_part[(size_t)blockIdx.x * kSweepPart + threadIdx.x] = s;
  }
}

}  // namespace ccg
