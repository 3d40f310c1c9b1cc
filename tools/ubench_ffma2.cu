#include <cuda_runtime.h>
#include <cstdio>
__global__ void k2(float2* out, const float2* in, int n) {
  float2 a[8];
  #pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = in[threadIdx.x * 8 + i];
  const float2 b = in[1000], c = in[1001];
  long long t0 = clock64();
  for (int k = 0; k < n; ++k) {
    #pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = __ffma2_rn(a[i], b, c);
  }
  long long t1 = clock64();
  float2 s = a[0];
  #pragma unroll
  for (int i = 1; i < 8; ++i) s = __fadd2_rn(s, a[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[100000].x = (float)(t1 - t0);
}
__global__ void k1(float* out, const float* in, int n) {
  float a[16];
  #pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = in[threadIdx.x * 16 + i];
  const float b = in[1000], c = in[1001];
  long long t0 = clock64();
  for (int k = 0; k < n; ++k) {
    #pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], b, c);
  }
  long long t1 = clock64();
  float s = a[0];
  #pragma unroll
  for (int i = 1; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[200000] = (float)(t1 - t0);
}
int main() {
  float *in, *out;
  cudaMalloc(&in, 1 << 20); cudaMalloc(&out, 1 << 22);
  cudaMemset(in, 0, 1 << 20);
  for (int warps = 1; warps <= 16; warps *= 2) {
    const int n = 10000;
    k1<<<148, 32 * warps>>>(out, in, n); cudaDeviceSynchronize();
    float c1; cudaMemcpy(&c1, out + 200000, 4, cudaMemcpyDeviceToHost);
    k2<<<148, 32 * warps>>>((float2*)out, (const float2*)in, n); cudaDeviceSynchronize();
    float c2; cudaMemcpy(&c2, out + 200000, 4, cudaMemcpyDeviceToHost);
    // per SMSP: warps/4 (min 1) warps each issuing 16 FFMA (k1) or 8 FFMA2 (k2) per iteration = 16 lane-FMAs
    printf("warps/CTA %2d: FFMA  %.2f cycles per 16 FMA/lane-iter,  FFMA2 %.2f cycles\n", warps, c1 / n, c2 / n);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
