// Latency / throughput of DFMA on sm_100a: NCH independent dependent-chains per lane, W warps per SM sub-partition.
#include <cuda_runtime.h>
#include <cstdio>
template <int NCH>
__global__ void k(double* out, int n, double b, double c, long long* cyc) {
  double a[NCH];
  #pragma unroll
  for (int i = 0; i < NCH; ++i) a[i] = threadIdx.x + i;
  long long t0 = clock64();
  for (int k = 0; k < n; ++k) {
    #pragma unroll
    for (int r = 0; r < 8; ++r) {
      #pragma unroll
      for (int i = 0; i < NCH; ++i) a[i] = fma(a[i], b, c);
    }
  }
  long long t1 = clock64();
  double s = 0;
  #pragma unroll
  for (int i = 0; i < NCH; ++i) s += a[i];
  if (s == 1.2345) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int NCH> void run(double* out, long long* cyc) {
  for (int wps = 1; wps <= 4; wps *= 2) {
    const int n = 2000;
    k<NCH><<<148, 128 * wps>>>(out, n, 1.0000001, 1e-9, cyc);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("chains/lane %d warps/SMSP %d: %.2f cycles per DFMA per warp (%.2f cycles per DFMA per SMSP)\n", NCH, wps,
           (double)c / (n * 8.0 * NCH), (double)c / (n * 8.0 * NCH * wps));
  }
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 64); cudaMalloc(&cyc, 8);
  run<1>(out, cyc); run<2>(out, cyc); run<3>(out, cyc); run<4>(out, cyc); run<6>(out, cyc); run<8>(out, cyc);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
