// Batched Undistort / Distort of calibrator::Calibrator on the device.
//
// Replaces, behind ccg_undistort / ccg_distort:
//   Calibrator::Undistort  reference src/calibrator.cpp:97-155 - per point a 2-unknown Ceres problem
//                          (DENSE_QR, <= 1000 iterations) whose minimum is the root of
//                          distort(x, y) = K^-1 pixel; here one thread per point takes Newton steps on
//                          the 2 x 2 system in double until the step is below 1e-16 relative
//   Calibrator::Distort    reference src/calibrator.cpp:157-166 - float32 lens model + K
// The caller passes K^-1 (float32, as the class forms it with Matrix3f::inverse) so that host
// and device share the same inverse.
#include "common.cuh"

namespace ccg {

__global__ void __launch_bounds__(256) k_undistort(const float* __restrict__ px, const float* __restrict__ py, int64_t n,
                                                  float i00, float i01, float i02, float i10, float i11, float i12,
                                                  float i20, float i21, float i22, double k1, double k2, double p1,
                                                  double p2, double k3, float* __restrict__ xo, float* __restrict__ yo) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float u = px[i], v = py[i];
  // homogeneous normalisation in float32, then double (calibrator.cpp:126-131)
  const float w = __fadd_rn(__fadd_rn(__fmul_rn(i20, u), __fmul_rn(i21, v)), i22);
  const double xt = (double)__fadd_rn(__fadd_rn(__fmul_rn(i00, u), __fmul_rn(i01, v)), i02) / (double)w;
  const double yt = (double)__fadd_rn(__fadd_rn(__fmul_rn(i10, u), __fmul_rn(i11, v)), i12) / (double)w;
  double x = xt, y = yt;
  for (int it = 0; it < 1000; ++it) {
    const double r2 = x * x + y * y, r4 = r2 * r2;
    const double m = 1.0 + k1 * r2 + k2 * r4 + k3 * r4 * r2, mp = k1 + 2.0 * k2 * r2 + 3.0 * k3 * r4;
    const double xd = x * m + 2.0 * p1 * x * y + p2 * (r2 + 2.0 * x * x);
    const double yd = y * m + 2.0 * p2 * x * y + p1 * (r2 + 2.0 * y * y);
    const double ex = xd - xt, ey = yd - yt;
    const double a = m + 2.0 * x * x * mp + 2.0 * p1 * y + 6.0 * p2 * x;
    const double b = 2.0 * x * y * mp + 2.0 * p1 * x + 2.0 * p2 * y;
    const double dd = m + 2.0 * y * y * mp + 2.0 * p2 * x + 6.0 * p1 * y;
    const double det = a * dd - b * b;
    if (!(fabs(det) > 1e-300)) break;
    const double dx = (dd * ex - b * ey) / det, dy = (a * ey - b * ex) / det;
    x -= dx;
    y -= dy;
    if (fabs(dx) + fabs(dy) <= 1e-16 * (1.0 + fabs(x) + fabs(y))) break;
  }
  xo[i] = (float)x;
  yo[i] = (float)y;
}

__global__ void __launch_bounds__(256) k_distort(const float* __restrict__ xn, const float* __restrict__ yn, int64_t n,
                                                float fx, float fy, float cx, float cy, float k1, float k2, float p1,
                                                float p2, float k3, float* __restrict__ uo, float* __restrict__ vo) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // float32 arithmetic without contraction, the way the host class evaluates the same template
  const float x = xn[i], y = yn[i];
  const float r2 = __fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), r4 = __fmul_rn(r2, r2), r6 = __fmul_rn(r4, r2);
  const float m = __fadd_rn(__fadd_rn(__fadd_rn(1.0f, __fmul_rn(k1, r2)), __fmul_rn(k2, r4)), __fmul_rn(k3, r6));
  const float xd = __fadd_rn(__fadd_rn(__fmul_rn(x, m), __fmul_rn(__fmul_rn(__fmul_rn(2.0f, p1), x), y)),
                             __fmul_rn(p2, __fadd_rn(r2, __fmul_rn(__fmul_rn(2.0f, x), x))));
  const float yd = __fadd_rn(__fadd_rn(__fmul_rn(y, m), __fmul_rn(__fmul_rn(__fmul_rn(2.0f, p2), x), y)),
                             __fmul_rn(p1, __fadd_rn(r2, __fmul_rn(__fmul_rn(2.0f, y), y))));
  uo[i] = __fadd_rn(__fmul_rn(fx, xd), cx);
  vo[i] = __fadd_rn(__fmul_rn(fy, yd), cy);
}

}  // namespace ccg

using namespace ccg;

namespace {
// two float inputs up, one kernel, two float outputs down; buffers from the context's pool
template <typename Launch>
int lens_call(ccg_ctx* ctx, const float* a, const float* b, int64_t n, float* oa, float* ob, Launch launch) {
  if (!ctx || n < 0 || (n > 0 && (!a || !b || !oa || !ob))) return CCG_ERR_BAD_ARG;
  if (n == 0) return CCG_OK;
  CCG_CUDA(ctx, cudaSetDevice(ctx->device));
  float* d = nullptr;
  CCG_CUDA(ctx, pool_alloc(ctx, (void**)&d, sizeof(float) * 4 * (size_t)n));
  cudaStream_t st = ctx->stream;
  cudaError_t e = cudaMemcpyAsync(d, a, sizeof(float) * n, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d + n, b, sizeof(float) * n, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) {
    launch(d, d + n, d + 2 * n, d + 3 * n, st);
    ctx->launches++;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(oa, d + 2 * n, sizeof(float) * n, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(ob, d + 3 * n, sizeof(float) * n, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  pool_free(ctx, d);
  if (e != cudaSuccess) return set_error(ctx, CCG_ERR_CUDA, "ccg lens kernel", e);
  return CCG_OK;
}
}  // namespace

extern "C" {

int ccg_undistort(ccg_ctx* ctx, const float* px, const float* py, int64_t n, const float K_inv[9],
                  const float dist[5], float* xn, float* yn) {
  if (!K_inv || !dist) return CCG_ERR_BAD_ARG;
  const int grid = (int)((n + 255) / 256);
  return lens_call(ctx, px, py, n, xn, yn, [&](float* a, float* b, float* oa, float* ob, cudaStream_t st) {
    k_undistort<<<grid, 256, 0, st>>>(a, b, n, K_inv[0], K_inv[1], K_inv[2], K_inv[3], K_inv[4], K_inv[5], K_inv[6],
                                      K_inv[7], K_inv[8], (double)dist[0], (double)dist[1], (double)dist[2],
                                      (double)dist[3], (double)dist[4], oa, ob);
  });
}

int ccg_distort(ccg_ctx* ctx, const float* xn, const float* yn, int64_t n, const float K[9], const float dist[5],
                float* px, float* py) {
  if (!K || !dist) return CCG_ERR_BAD_ARG;
  const int grid = (int)((n + 255) / 256);
  return lens_call(ctx, xn, yn, n, px, py, [&](float* a, float* b, float* oa, float* ob, cudaStream_t st) {
    k_distort<<<grid, 256, 0, st>>>(a, b, n, K[0], K[4], K[2], K[5], dist[0], dist[1], dist[2], dist[3], dist[4], oa, ob);
  });
}

}  // extern "C"
