// Mixed-precision single-camera sweep (included by intr.cu).
//
// Why: the all-FP64 sweep is FP64-pipe bound at ~716 flop/observation (36 flop/B).  Only the
// residual r, the cost and the gradient g = J'r decide WHERE the LM iteration converges; the
// normal-equation blocks V, W, U (= J'J) only decide HOW FAST (a perturbed Gauss-Newton matrix
// still has g = 0 as its fixed point, and every step is validated against the FP64 cost).  So:
//   FP64 : forward model, residual, cost, and g_v / g_u by REVERSE mode (no explicit Jacobian)
//   FP32 : a second forward pass, the analytic Jacobian and the J'J accumulators
// One warp (or a G-lane sub-warp group, G = 32/16/8 chosen from the view size) owns a view, so no
// redundant geometry; FP64 and FP32 pipes overlap.  Convergence equivalence was checked in the
// oracle (ORACLE_H_FLOAT experiment) and is enforced by the GPU parity tests.
#pragma once

namespace ccg {

constexpr int kNH = 116;  // FP32 accumulators per lane: V 21 | W[:,0..3] 24 | W[:,4..8] 30 | U 41
constexpr int kUOff = 75; // first U element in that array

// transposing reduction over a group of G lanes (G = 8, 16, 32): a[0..G) per lane in, the sum over
// the group's lanes of a[gl] out (gl = lane within the group)
template <int G, typename T>
__device__ __forceinline__ T group_reduce(T (&a)[G], int gl) {
#pragma unroll
  for (int half = G / 2; half >= 1; half >>= 1) {
    const bool up = (gl & half) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const T send = up ? a[i] : a[i + half];
      const T keep = up ? a[i + half] : a[i];
      a[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
    }
  }
  return a[0];
}

struct PoseF {
  double R[9], t[3];
  float Rf[9], tf[3];
};

template <int G>
__global__ void __launch_bounds__(kWarps * 32, 1) k_intr_sweep_mixed(IntrDev d, double* __restrict__ res_out) {
  constexpr int NG = 32 / G;             // views in flight per warp
  constexpr int NCH = (kNH + G - 1) / G; // FP32 chunks per view
  constexpr int CU0 = kUOff / G;         // first chunk that contains U elements
  constexpr int NU = NCH - CU0;          // U values a lane can own
  __shared__ double s_k[9], s_du[9];
  __shared__ float s_kf[9];
  __shared__ int s_meta[2];
  __shared__ double s_u[kWarps][NG][41];
  __shared__ double s_g[kWarps][NG][12];  // g_u 9 | sum r^2 | g'd | d'Dd
  __shared__ float s_pose[kWarps][NG][12];  // float copy of the candidate R (9) and t (3) of the view in flight
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
    const double kv = ps.intr[par ^ 1][threadIdx.x];
    s_k[threadIdx.x] = kv;
    s_kf[threadIdx.x] = (float)kv;
    s_du[threadIdx.x] = ps.du[threadIdx.x];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = lane / G, gl = lane % G;
  double k[9];
  float kf[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    k[i] = s_k[i];
    kf[i] = s_kf[i];
  }
  double gu[9], cost = 0.0, gTd = 0.0, dDd = 0.0, uacc[NU];
#pragma unroll
  for (int i = 0; i < 9; ++i) gu[i] = 0.0;
#pragma unroll
  for (int i = 0; i < NU; ++i) uacc[i] = 0.0;

  const int slots = kWarps * NG;
  for (int64_t vb = it.v0; vb < it.v1; vb += slots) {  // warp-uniform trip count
    const int64_t v = vb + warp * NG + grp;
    const bool has_view = v < it.v1;
    // ---------------- candidate pose of this group's view: current (+) (-z - Y du)
    double R[9], t[3];
    {
      double dl = 0.0;
      const int64_t vv = has_view ? v : it.v0;
      const double* yz = d.yz + (size_t)vv * kYz;
      if (gl < 6) {
        dl = -yz[54 + gl];
#pragma unroll
        for (int j = 0; j < 9; ++j) dl -= yz[gl * 9 + j] * s_du[j];
        if (has_view) {
          gTd += yz[60 + gl] * dl;
          dDd += dl * dl * yz[66 + gl];
        }
      }
      const int base = grp * G;
      const double d0 = __shfl_sync(0xffffffffu, dl, base + 0), d1 = __shfl_sync(0xffffffffu, dl, base + 1),
                   d2 = __shfl_sync(0xffffffffu, dl, base + 2), d3 = __shfl_sync(0xffffffffu, dl, base + 3),
                   d4 = __shfl_sync(0xffffffffu, dl, base + 4), d5 = __shfl_sync(0xffffffffu, dl, base + 5);
      const double* pc = d.pose[par] + (size_t)vv * kPose;
      double q[4] = {pc[0], pc[1], pc[2], pc[3]};
      double qn[4];
      quat_plus(q, d0, d1, d2, qn);
      t[0] = pc[4] + d3;
      t[1] = pc[5] + d4;
      t[2] = pc[6] + d5;
      if (has_view && gl < 7) {
        double* pn = d.pose[par ^ 1] + (size_t)v * kPose;
        const double val = gl == 0 ? qn[0] : gl == 1 ? qn[1] : gl == 2 ? qn[2] : gl == 3 ? qn[3]
                         : gl == 4 ? t[0] : gl == 5 ? t[1] : t[2];
        pn[gl] = val;
      }
      quat_to_rot(qn, R);
      // the FP32 pass reads its pose from shared memory: keeping a second register copy forces
      // the compiler to re-convert R every iteration (12 F2F per observation)
      __syncwarp();
      for (int e = gl; e < 12; e += G) {  // G may be 8: a lane can own two entries
        double val = t[0];
#pragma unroll
        for (int i = 0; i < 9; ++i) val = (e == i) ? R[i] : val;
        val = (e == 10) ? t[1] : (e == 11) ? t[2] : val;
        s_pose[warp][grp][e] = (float)val;
      }
      __syncwarp();
    }
    const float* Rf = s_pose[warp][grp];
    const float* tf = Rf + 9;
    float acc[kNH];
#pragma unroll
    for (int i = 0; i < kNH; ++i) acc[i] = 0.0f;
    double gv[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) gv[i] = 0.0;

    const int64_t o0 = has_view ? d.view_off[v] : 0, o1 = has_view ? d.view_off[v + 1] : 0;
    // software-pipelined observation stream: the next fetch is in flight during the math
    ObsLoad nx;
    nx.fetch(d, o0 + gl, o1);
    while (__any_sync(0xffffffffu, nx.valid)) {
      const ObsLoad cur = nx;
      nx.fetch(d, cur.o + G, o1);
      if (!cur.valid) continue;
      // ================= FP64: forward, residual, cost, gradient by reverse mode =================
      {
        const double X = (double)cur.X, Y = (double)cur.Y, Z = (double)cur.Z;
        const double y0 = R[0] * X + R[1] * Y + R[2] * Z;
        const double y1 = R[3] * X + R[4] * Y + R[5] * Z;
        const double y2 = R[6] * X + R[7] * Y + R[8] * Z;
        const double iz = fast_rcp(y2 + t[2]);
        const double xn = (y0 + t[0]) * iz, yn = (y1 + t[1]) * iz;
        const double r2 = xn * xn + yn * yn, r4 = r2 * r2, r6 = r4 * r2;
        const double m = 1.0 + k[4] * r2 + k[5] * r4 + k[8] * r6;
        const double a1 = 2.0 * xn * yn, a2 = r2 + 2.0 * xn * xn, a3 = r2 + 2.0 * yn * yn;
        const double xd = xn * m + k[6] * a1 + k[7] * a2;
        const double yd = yn * m + k[7] * a1 + k[6] * a3;
        const double ru = k[0] * xd + k[2] - (double)cur.mu;
        const double rv = k[1] * yd + k[3] - (double)cur.mv;
        if (res_out) {
          res_out[2 * (cur.o - d.obs_base)] = ru;
          res_out[2 * (cur.o - d.obs_base) + 1] = rv;
        }
        cost = fma(ru, ru, fma(rv, rv, cost));
        const double tu = k[0] * ru, tv = k[1] * rv;  // d cost / d(xd, yd)
        gu[0] = fma(xd, ru, gu[0]);
        gu[1] = fma(yd, rv, gu[1]);
        gu[2] += ru;
        gu[3] += rv;
        const double sr = xn * tu + yn * tv;
        gu[4] = fma(r2, sr, gu[4]);
        gu[5] = fma(r4, sr, gu[5]);
        gu[8] = fma(r6, sr, gu[8]);
        gu[6] = fma(a1, tu, fma(a3, tv, gu[6]));
        gu[7] = fma(a2, tu, fma(a1, tv, gu[7]));
        const double mp = k[4] + 2.0 * k[5] * r2 + 3.0 * k[8] * r4;
        const double F00 = m + 2.0 * xn * xn * mp + 2.0 * k[6] * yn + 6.0 * k[7] * xn;
        const double F01 = a1 * mp + 2.0 * k[6] * xn + 2.0 * k[7] * yn;
        const double F11 = m + 2.0 * yn * yn * mp + 2.0 * k[7] * xn + 6.0 * k[6] * yn;
        const double w0 = (F00 * tu + F01 * tv) * iz, w1 = (F01 * tu + F11 * tv) * iz;  // d cost / d(p0,p1)
        const double w2 = -(xn * w0 + yn * w1);                                        // d cost / d p2
        gv[3] += w0;
        gv[4] += w1;
        gv[5] += w2;
        gv[0] = fma(2.0, y1 * w2 - y2 * w1, gv[0]);  // rotation tangent: 2 (y x w)
        gv[1] = fma(2.0, y2 * w0 - y0 * w2, gv[1]);
        gv[2] = fma(2.0, y0 * w1 - y1 * w0, gv[2]);
      }
      // ================= FP32: Jacobian and J'J accumulators =================
      {
        const float X = cur.X, Y = cur.Y, Z = cur.Z;
        const float y0 = Rf[0] * X + Rf[1] * Y + Rf[2] * Z;
        const float y1 = Rf[3] * X + Rf[4] * Y + Rf[5] * Z;
        const float y2 = Rf[6] * X + Rf[7] * Y + Rf[8] * Z;
        const float iz = fast_rcpf(y2 + tf[2]);
        const float xn = (y0 + tf[0]) * iz, yn = (y1 + tf[1]) * iz;
        const float r2 = xn * xn + yn * yn, r4 = r2 * r2, r6 = r4 * r2;
        const float m = 1.0f + kf[4] * r2 + kf[5] * r4 + kf[8] * r6;
        const float a1 = 2.0f * xn * yn, a2 = r2 + 2.0f * xn * xn, a3 = r2 + 2.0f * yn * yn;
        const float xd = xn * m + kf[6] * a1 + kf[7] * a2;
        const float yd = yn * m + kf[7] * a1 + kf[6] * a3;
        const float mp = kf[4] + 2.0f * kf[5] * r2 + 3.0f * kf[8] * r4;
        const float F00 = m + 2.0f * xn * xn * mp + 2.0f * kf[6] * yn + 6.0f * kf[7] * xn;
        const float F01 = a1 * mp + 2.0f * kf[6] * xn + 2.0f * kf[7] * yn;
        const float F11 = m + 2.0f * yn * yn * mp + 2.0f * kf[7] * xn + 6.0f * kf[6] * yn;
        const float A0 = kf[0] * iz * F00, A1 = kf[0] * iz * F01, B0 = kf[1] * iz * F01, B1 = kf[1] * iz * F11;
        const float c00 = -2.0f * xn * y1, c01 = 2.0f * (y2 + xn * y0), c02 = -2.0f * y1;
        const float c10 = -2.0f * (y2 + yn * y1), c11 = 2.0f * yn * y0, c12 = 2.0f * y0;
        float ju[6], jv[6], au[5], av[5];
        ju[0] = A0 * c00 + A1 * c10; ju[1] = A0 * c01 + A1 * c11; ju[2] = A0 * c02 + A1 * c12;
        ju[3] = A0; ju[4] = A1; ju[5] = -(A0 * xn + A1 * yn);
        jv[0] = B0 * c00 + B1 * c10; jv[1] = B0 * c01 + B1 * c11; jv[2] = B0 * c02 + B1 * c12;
        jv[3] = B0; jv[4] = B1; jv[5] = -(B0 * xn + B1 * yn);
        au[0] = kf[0] * xn * r2; au[1] = kf[0] * xn * r4; au[2] = kf[0] * a1; au[3] = kf[0] * a2; au[4] = kf[0] * xn * r6;
        av[0] = kf[1] * yn * r2; av[1] = kf[1] * yn * r4; av[2] = kf[1] * a3; av[3] = kf[1] * a1; av[4] = kf[1] * yn * r6;
        int idx = 0;
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
          for (int b = a; b < 6; ++b) {
            acc[idx] = fmaf(ju[a], ju[b], fmaf(jv[a], jv[b], acc[idx]));
            ++idx;
          }
#pragma unroll
        for (int a = 0; a < 6; ++a) {
          acc[21 + a * 4 + 0] = fmaf(ju[a], xd, acc[21 + a * 4 + 0]);
          acc[21 + a * 4 + 1] = fmaf(jv[a], yd, acc[21 + a * 4 + 1]);
          acc[21 + a * 4 + 2] += ju[a];
          acc[21 + a * 4 + 3] += jv[a];
#pragma unroll
          for (int c = 0; c < 5; ++c)
            acc[45 + a * 5 + c] = fmaf(ju[a], au[c], fmaf(jv[a], av[c], acc[45 + a * 5 + c]));
        }
        // U: [0..5] fxfx fxpx pxpx fyfy fypy pypy | fx,d 5 | px,d 5 | fy,d 5 | py,d 5 | d,d 15
        float* U = acc + kUOff;
        U[0] = fmaf(xd, xd, U[0]);
        U[1] += xd;
        U[2] += 1.0f;
        U[3] = fmaf(yd, yd, U[3]);
        U[4] += yd;
        U[5] += 1.0f;
        int ui = 26;
#pragma unroll
        for (int c = 0; c < 5; ++c) {
          U[6 + c] = fmaf(xd, au[c], U[6 + c]);
          U[11 + c] += au[c];
          U[16 + c] = fmaf(yd, av[c], U[16 + c]);
          U[21 + c] += av[c];
#pragma unroll
          for (int e = c; e < 5; ++e) {
            U[ui] = fmaf(au[c], au[e], fmaf(av[c], av[e], U[ui]));
            ++ui;
          }
        }
      }
    }
    // ---------------- per-view reduction over the group's lanes and block store
    double* blk = d.blk[par ^ 1] + (size_t)(has_view ? v : it.v0) * kBlk;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      float tmp[G];
#pragma unroll
      for (int i = 0; i < G; ++i) tmp[i] = (c * G + i < kNH) ? acc[c * G + i] : 0.0f;
      const float r = group_reduce<G, float>(tmp, gl);
      const int e = c * G + gl;
      if (c < CU0 || (c == CU0 && e < kUOff)) {
        if (has_view && e < kUOff) blk[e < 21 ? e : e + 6] = (double)r;
      }
      if (c >= CU0) {
        if (e >= kUOff && e < kNH) uacc[c - CU0] += (double)r;
      }
    }
    {
      double tmp[G];
#pragma unroll
      for (int i = 0; i < G; ++i) tmp[i] = i < 6 ? gv[i] : 0.0;
      const double r = group_reduce<G, double>(tmp, gl);
      if (has_view && gl < 6) blk[21 + gl] = r;
    }
  }
  // ---------------- end of item: persistent sums -> shared -> one partial per CTA
  {
    const double vals[12] = {gu[0], gu[1], gu[2], gu[3], gu[4], gu[5], gu[6], gu[7], gu[8], cost, gTd, dDd};
    if constexpr (G >= 16) {  // all 12 values fit one chunk
      double tmp[G];
#pragma unroll
      for (int i = 0; i < G; ++i) tmp[i] = i < 12 ? vals[i] : 0.0;
      const double r = group_reduce<G, double>(tmp, gl);
      if (gl < 12) s_g[warp][grp][gl] = r;
    } else {  // G == 8: two chunks
      double tmp[G], tmp2[G];
#pragma unroll
      for (int i = 0; i < G; ++i) {
        tmp[i] = vals[i];
        tmp2[i] = i < 4 ? vals[8 + i] : 0.0;
      }
      const double r0 = group_reduce<G, double>(tmp, gl);
      const double r1 = group_reduce<G, double>(tmp2, gl);
      s_g[warp][grp][gl] = r0;
      if (gl < 4) s_g[warp][grp][8 + gl] = r1;
    }
#pragma unroll
    for (int c = CU0; c < NCH; ++c) {
      const int e = c * G + gl;
      if (e >= kUOff && e < kNH) s_u[warp][grp][e - kUOff] = uacc[c - CU0];
    }
  }
  __syncthreads();
  if (threadIdx.x < kSweepPart) {
    // partial layout shared with the FP64 sweep: [0] sum r^2 | [1..6] U(fx,px,fy,py block) | [7] g'd |
    // [8] d'Dd | [16..24] g_u | [25..59] U rows that involve distortion
    const int c = threadIdx.x;
    double s = 0.0;
    for (int w = 0; w < kWarps; ++w)
      for (int g = 0; g < NG; ++g) {
        double val = 0.0;
        if (c == 0) val = s_g[w][g][9];
        else if (c >= 1 && c <= 6) val = s_u[w][g][c - 1];
        else if (c == 7) val = s_g[w][g][10];
        else if (c == 8) val = s_g[w][g][11];
        else if (c >= 16 && c < 25) val = s_g[w][g][c - 16];
        else if (c >= 25 && c < 60) val = s_u[w][g][6 + (c - 25)];
        s += val;
      }
    d.sweep_part[(size_t)blockIdx.x * kSweepPart + c] = s;
  }
}

}  // namespace ccg
