// The FP64 role of the sweep as a standalone streaming kernel, to find what separates it from the
// bare FP64 block (tools/ubench_fp64block.cu: 220 cycles per 32-observation chunk per SM
// sub-partition at 2 warps).  Flags switch the features of the real loop on one by one.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o ubench_fp64role ubench_fp64role.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = fma(fma(-x, r, 1.0), r, r);
  r = fma(fma(-x, r, 1.0), r, r);
  return r;
}
struct K64 { double fx, fy, cx, cy, k1, k2, p1, p2, k3; };

__device__ __forceinline__ void obs(const K64& K, const double (&R)[9], const double (&t)[3], double X, double Y,
                                    double Z, double mu, double mv, bool valid, double (&gu)[9], double (&gv)[6],
                                    double& cost) {
  const double pt0 = fma(R[0], X, fma(R[1], Y, fma(R[2], Z, t[0])));
  const double pt1 = fma(R[3], X, fma(R[4], Y, fma(R[5], Z, t[1])));
  const double pt2 = fma(R[6], X, fma(R[7], Y, fma(R[8], Z, t[2])));
  const double iz = fast_rcp(pt2);
  const double xn = pt0 * iz, yn = pt1 * iz;
  const double xx = xn * xn, yy = yn * yn, xy = xn * yn;
  const double r2 = xx + yy, r4 = r2 * r2, r6 = r4 * r2;
  const double h2 = fma(r2, K.k3, K.k2), h1 = fma(r2, h2, K.k1);
  const double m = fma(r2, h1, 1.0), mp = fma(r2, fma(r2, K.k3, h2), h1);
  const double a1 = xy + xy, a2 = fma(2.0, xx, r2), a3 = fma(2.0, yy, r2);
  const double xd = fma(xn, m, fma(K.p1, a1, K.p2 * a2));
  const double yd = fma(yn, m, fma(K.p2, a1, K.p1 * a3));
  double ru = fma(K.fx, xd, K.cx - mu);
  double rv = fma(K.fy, yd, K.cy - mv);
  ru = valid ? ru : 0.0;
  rv = valid ? rv : 0.0;
  cost = fma(ru, ru, fma(rv, rv, cost));
  const double tu = K.fx * ru, tv = K.fy * rv;
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
  const double tx = xn + xn, ty = yn + yn;
  const double u1 = K.p1 * ty, u2 = K.p2 * tx, base = m + (u1 + u2);
  const double F00 = fma(2.0, u2, fma(a2 - r2, mp, base));
  const double F01 = fma(a1, mp, fma(K.p1, tx, K.p2 * ty));
  const double F11 = fma(2.0, u1, fma(a3 - r2, mp, base));
  const double w0 = fma(F00, tu, F01 * tv) * iz, w1 = fma(F01, tu, F11 * tv) * iz;
  const double w2 = -fma(xn, w0, yn * w1);
  gv[3] += w0;
  gv[4] += w1;
  gv[5] += w2;
  gv[0] = fma(pt1, w2, fma(-pt2, w1, gv[0]));
  gv[1] = fma(pt2, w0, fma(-pt0, w2, gv[1]));
  gv[2] = fma(pt0, w1, fma(-pt1, w0, gv[2]));
}

// MODE 0: observations from registers; 1: plain LDG (one chunk ahead); 2: cp.async ring (two chunks ahead)
template <int MODE, bool VALID, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) k_role(double* out, const double* in, const float* u, const float* v,
                                                       const float* X, const float* Y, const float* Z,
                                                       long chunks_per_warp) {
  __shared__ float s_ring[WARPS][3][160];
  K64 K = {in[0], in[1], in[2], in[3], in[4], in[5], in[6], in[7], in[8]};
  double R[9], t[3];
  for (int i = 0; i < 9; ++i) R[i] = in[9 + i];
  for (int i = 0; i < 3; ++i) t[i] = in[18 + i];
  double gu[9], gv[6], cost = 0.0;
  for (int i = 0; i < 9; ++i) gu[i] = 0.0;
  for (int i = 0; i < 6; ++i) gv[i] = 0.0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long w = (long)blockIdx.x * WARPS + warp;
  const long o0 = w * chunks_per_warp * 32, o1 = o0 + chunks_per_warp * 32 - (VALID ? 7 : 0);
  float* ring = &s_ring[warp][0][0];
  auto prefetch = [&](int st, long idx) {
    const bool ok = idx < o1;
    const long o = ok ? idx : o0;
    const int sz = ok ? 4 : 0;
    const unsigned dst = (unsigned)__cvta_generic_to_shared(ring + st * 160 + lane);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(u + o), "r"(sz) : "memory");
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst + 128), "l"(v + o), "r"(sz) : "memory");
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst + 256), "l"(X + o), "r"(sz) : "memory");
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst + 384), "l"(Y + o), "r"(sz) : "memory");
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst + 512), "l"(Z + o), "r"(sz) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  float fu = 800.f + lane, fv = 500.f + lane, fX = 1e-3f * lane, fY = 2e-3f * lane, fZ = 0.01f;
  if (MODE == 2) {
    prefetch(0, o0 + lane);
    prefetch(1, o0 + 32 + lane);
  }
  float nu = 0, nv = 0, nX = 0, nY = 0, nZ = 0;
  if (MODE == 1) {
    nu = u[o0 + lane]; nv = v[o0 + lane]; nX = X[o0 + lane]; nY = Y[o0 + lane]; nZ = Z[o0 + lane];
  }
  int st = 0;
  for (long ob = o0; ob < o1; ob += 32) {
    bool valid = true;
    if (MODE == 2) {
      const int st_free = st == 0 ? 2 : st - 1;
      prefetch(st_free, ob + 64 + lane);
      asm volatile("cp.async.wait_group 2;" ::: "memory");
      const float* s = ring + st * 160;
      fu = s[lane]; fv = s[32 + lane]; fX = s[64 + lane]; fY = s[96 + lane]; fZ = s[128 + lane];
      st = st == 2 ? 0 : st + 1;
    } else if (MODE == 1) {
      fu = nu; fv = nv; fX = nX; fY = nY; fZ = nZ;
      const long nx = ob + 32 + lane < o1 ? ob + 32 + lane : o0;
      nu = u[nx]; nv = v[nx]; nX = X[nx]; nY = Y[nx]; nZ = Z[nx];
    } else {
      fX += 1e-5f;
      fY -= 1e-5f;
    }
    if (VALID) valid = ob + lane < o1;
    obs(K, R, t, (double)fX, (double)fY, (double)fZ, (double)fu, (double)fv, valid, gu, gv, cost);
  }
  double s = cost;
  for (int i = 0; i < 9; ++i) s += gu[i];
  for (int i = 0; i < 6; ++i) s += gv[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double ghz = khz * 1e-6;
  double h[21] = {1000, 1000, 800, 500, -0.04, 5e-4, 1e-3, 2e-5, -3e-4, 1, 0, 0, 0, 1, 0, 0, 0, 1, 0.01, 0.02, 0.6};
  double *din, *dout;
  cudaMalloc(&din, sizeof(h));
  cudaMemcpy(din, h, sizeof(h), cudaMemcpyHostToDevice);
  cudaMalloc(&dout, sizeof(double) * sms * 1024);
  const long n = 40000000;
  std::vector<float> hu(n), hv(n), hX(n), hY(n), hZ(n);
  for (long i = 0; i < n; ++i) {
    hu[i] = 800.f + (i % 977) * 0.3f;
    hv[i] = 500.f + (i % 613) * 0.3f;
    hX[i] = 1e-3f * (i % 311);
    hY[i] = 1e-3f * (i % 419);
    hZ[i] = 0.f;
  }
  float *u, *v, *X, *Y, *Z;
  for (float** p : {&u, &v, &X, &Y, &Z}) cudaMalloc(p, sizeof(float) * n);
  cudaMemcpy(u, hu.data(), sizeof(float) * n, cudaMemcpyHostToDevice);
  cudaMemcpy(v, hv.data(), sizeof(float) * n, cudaMemcpyHostToDevice);
  cudaMemcpy(X, hX.data(), sizeof(float) * n, cudaMemcpyHostToDevice);
  cudaMemcpy(Y, hY.data(), sizeof(float) * n, cudaMemcpyHostToDevice);
  cudaMemcpy(Z, hZ.data(), sizeof(float) * n, cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  auto time = [&](const char* name, int warps, auto launch) {
    const long cpw = n / 32 / ((long)sms * warps);
    launch(cpw);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    launch(cpw);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double chunks_per_smsp = (double)cpw * warps / 4.0;
    printf("%-52s %2d warps/SM: %.3f ms, %.0f cycles per chunk per SMSP  (%s)\n", name, warps, ms,
           ms * 1e-3 * ghz * 1e9 / chunks_per_smsp, cudaGetErrorString(cudaGetLastError()));
  };
#define RUN(MODE, VALID, W, NAME) \
  time(NAME, W, [&](long cpw) { k_role<MODE, VALID, W><<<sms, W * 32>>>(dout, din, u, v, X, Y, Z, cpw); })
  RUN(0, false, 8, "registers, no validity select");
  RUN(0, true, 8, "registers, validity select");
  RUN(1, false, 8, "plain LDG one chunk ahead");
  RUN(2, false, 8, "cp.async ring two chunks ahead");
  RUN(2, true, 8, "cp.async ring + validity select (= the role's loop)");
  RUN(2, true, 4, "cp.async ring + validity select");
  RUN(2, true, 16, "cp.async ring + validity select");
  return 0;
}
