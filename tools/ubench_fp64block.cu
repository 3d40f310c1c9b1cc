// How fast can the FP64 block of the sweep (forward model + reverse-mode gradient, 88 FP64
// instructions per observation) run as a function of warps per SM sub-partition and of the
// number of independent observations per lane?  (Design evidence for the sweep kernel.)
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o ubench_fp64block ubench_fp64block.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = fma(fma(-x, r, 1.0), r, r);
  r = fma(fma(-x, r, 1.0), r, r);
  return r;
}
struct K64 { double fx, fy, cx, cy, k1, k2, p1, p2, k3; };

template <bool SEPACC>
__device__ __forceinline__ void obs(const K64& K, const double (&R)[9], const double (&t)[3], double X, double Y,
                                    double Z, double mu, double mv, double (&gu)[9], double (&gv)[6], double& cost) {
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
  const double ru = fma(K.fx, xd, K.cx - mu);
  const double rv = fma(K.fy, yd, K.cy - mv);
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

// One warp per SM sub-partition and CTA (128 threads); W CTAs resident per SM give W warps per sub-partition, and
// __launch_bounds__(128, W) caps the registers accordingly (W = 4: 128 registers, as in the sweep's launch).
template <int ILP, int W>
__global__ void __launch_bounds__(128, W) k_block(double* out, const double* in, int iters) {
  K64 K = {in[0], in[1], in[2], in[3], in[4], in[5], in[6], in[7], in[8]};
  double R[9], t[3];
  for (int i = 0; i < 9; ++i) R[i] = in[9 + i];
  for (int i = 0; i < 3; ++i) t[i] = in[18 + i];
  double gu[9], gv[6], cost = 0.0;
  for (int i = 0; i < 9; ++i) gu[i] = 0.0;
  for (int i = 0; i < 6; ++i) gv[i] = 0.0;
  double X[ILP], Y[ILP];
  for (int e = 0; e < ILP; ++e) {
    X[e] = 1e-3 * (threadIdx.x + e);
    Y[e] = 2e-3 * (threadIdx.x + 3 * e);
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int e = 0; e < ILP; ++e) {
      obs<false>(K, R, t, X[e], Y[e], 0.01, 800.0 + X[e], 500.0 + Y[e], gu, gv, cost);
      X[e] += 1e-5;
      Y[e] -= 1e-5;
    }
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
  cudaMalloc(&dout, sizeof(double) * sms * 1024);  // (4 CTAs x 128 threads per SM at most)
  const int iters = 2048;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  auto run = [&](int ilp, int w) {
    auto launch = [&] {
#define CCG_CASE(I, WW) if (ilp == I && w == WW) k_block<I, WW><<<sms * WW, 128>>>(dout, din, iters)
      CCG_CASE(1, 1); CCG_CASE(1, 2); CCG_CASE(1, 3); CCG_CASE(1, 4);
      CCG_CASE(2, 1); CCG_CASE(2, 2); CCG_CASE(2, 3); CCG_CASE(2, 4);
      CCG_CASE(4, 1); CCG_CASE(4, 2);
#undef CCG_CASE
    };
    launch();
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) {
      printf("ILP %d, %d warps/SMSP: launch failed (%s)\n", ilp, w, cudaGetErrorString(err));
      return;
    }
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    err = cudaGetLastError();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double obs_per_smsp = (double)w * iters * ilp;  // warp-observations (32 obs each)
    printf("ILP %d, %d warps/SMSP: %.1f cycles per 32-observation chunk per SMSP (FP64 pipe floor 176)%s\n", ilp, w,
           ms * 1e-3 * ghz * 1e9 / obs_per_smsp, err == cudaSuccess ? "" : "  [launch error]");
  };
  for (int ilp : {1, 2, 4})
    for (int w : {1, 2, 3, 4}) {
      if (ilp == 4 && w > 2) continue;  // (four observations in flight per lane do not fit 128 registers)
      run(ilp, w);
    }
  return 0;
}
