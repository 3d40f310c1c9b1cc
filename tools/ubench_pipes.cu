// Pipe-throughput micro-benchmarks behind the sweep-kernel design:
//   legacy mma.sync (tf32 m16n8k8, bf16 m16n8k16) rate, DFMA / FFMA / FFMA2 rates, whether FP64 and
//   FP32 work overlap, and the cost of F2F conversions.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o ubench_pipes ubench_pipes.cu
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                   \
  do {                                                                          \
    cudaError_t e = (x);                                                        \
    if (e != cudaSuccess) {                                                     \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(1);                                                                  \
    }                                                                           \
  } while (0)

constexpr int kIters = 4096;

__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__device__ __forceinline__ void mma_tf32_k4(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(b[0]));
}

template <int KIND, int CHAINS>
__global__ void __launch_bounds__(256) k_mma(float* out, unsigned seed) {
  unsigned a[4], b[2];
  for (int i = 0; i < 4; ++i) a[i] = seed + threadIdx.x * 4 + i;
  for (int i = 0; i < 2; ++i) b[i] = seed * 3 + threadIdx.x * 2 + i;
  float c[CHAINS][4];
  for (int j = 0; j < CHAINS; ++j)
    for (int i = 0; i < 4; ++i) c[j][i] = 0.f;
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int j = 0; j < CHAINS; ++j) {
      if (KIND == 0) mma_tf32(c[j], a, b);
      else if (KIND == 1) mma_bf16(c[j], a, b);
      else mma_tf32_k4(c[j], a, b);
    }
  }
  float s = 0.f;
  for (int j = 0; j < CHAINS; ++j)
    for (int i = 0; i < 4; ++i) s += c[j][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ND independent DFMA chains and NF independent FFMA chains per iteration, NP packed FFMA2 chains
template <int ND, int NF, int NP>
__global__ void __launch_bounds__(256) k_mix(double* out, double a, double b, float fa, float fb) {
  double x[ND > 0 ? ND : 1];
  float y[NF > 0 ? NF : 1];
  unsigned long long z[NP > 0 ? NP : 1];
  for (int i = 0; i < ND; ++i) x[i] = threadIdx.x + i;
  for (int i = 0; i < NF; ++i) y[i] = threadIdx.x + i;
  for (int i = 0; i < NP; ++i) z[i] = ((unsigned long long)__float_as_uint((float)threadIdx.x + i) << 32) | __float_as_uint(1.0f + i);
  unsigned long long pa = ((unsigned long long)__float_as_uint(fa) << 32) | __float_as_uint(fa);
  unsigned long long pb = ((unsigned long long)__float_as_uint(fb) << 32) | __float_as_uint(fb);
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < ND; ++i) x[i] = fma(x[i], a, b);
#pragma unroll
      for (int i = 0; i < NF; ++i) y[i] = fmaf(y[i], fa, fb);
#pragma unroll
      for (int i = 0; i < NP; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(z[i]) : "l"(pa), "l"(pb));
    }
  }
  double s = 0.0;
  for (int i = 0; i < ND; ++i) s += x[i];
  for (int i = 0; i < NF; ++i) s += y[i];
  for (int i = 0; i < NP; ++i) s += (double)(z[i] & 0xffff);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// conversions: KIND 0 = f32->f64, 1 = f64->f32, 2 = cvt.rna.tf32
template <int KIND>
__global__ void __launch_bounds__(256) k_cvt(double* out, float f0) {
  float f[8];
  double d[8];
  for (int i = 0; i < 8; ++i) {
    f[i] = f0 + threadIdx.x + i;
    d[i] = f[i];
  }
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (KIND == 0) {
        double t;
        asm volatile("cvt.f64.f32 %0, %1;" : "=d"(t) : "f"(f[i]));
        asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f[i]) : "d"(t));
      } else if (KIND == 1) {
        asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f[i]) : "d"(d[i]));
        d[i] += f[i];
      } else {
        unsigned t;
        asm volatile("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(f[i]));
        f[i] = __uint_as_float(t) + 1.0f;
      }
    }
  }
  double s = 0.0;
  for (int i = 0; i < 8; ++i) s += f[i] + d[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
double time_ms(F launch) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  launch();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < 5; ++i) launch();
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms / 5.0;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double ghz = khz * 1e-6;
  printf("device %s, %d SMs, max clock %.3f GHz (cycles below assume it)\n", prop.name, sms, ghz);
  void* buf;
  CK(cudaMalloc(&buf, (size_t)sms * 8 * 1024 * 8));
  const int grid = sms * 4;  // 4 CTAs x 256 threads per SM = 32 warps/SM = 8 warps/SMSP
  auto rep = [&](const char* name, double ms, double warp_instr_per_thread_iter) {
    // warp-instructions per SMSP: 8 warps/SMSP x kIters x per-iteration count
    const double instr = 8.0 * kIters * warp_instr_per_thread_iter;
    const double cyc = ms * 1e-3 * ghz * 1e9;
    printf("%-44s %8.3f ms  %7.3f cycles / warp-instr / SMSP\n", name, ms, cyc / instr);
  };
  rep("mma tf32 m16n8k8, 1 chain", time_ms([&] { k_mma<0, 1><<<grid, 256>>>((float*)buf, 1); }), 1);
  rep("mma tf32 m16n8k8, 2 chains", time_ms([&] { k_mma<0, 2><<<grid, 256>>>((float*)buf, 1); }), 2);
  rep("mma tf32 m16n8k8, 4 chains", time_ms([&] { k_mma<0, 4><<<grid, 256>>>((float*)buf, 1); }), 4);
  rep("mma bf16 m16n8k16, 1 chain", time_ms([&] { k_mma<1, 1><<<grid, 256>>>((float*)buf, 1); }), 1);
  rep("mma bf16 m16n8k16, 2 chains", time_ms([&] { k_mma<1, 2><<<grid, 256>>>((float*)buf, 1); }), 2);
  rep("mma bf16 m16n8k16, 4 chains", time_ms([&] { k_mma<1, 4><<<grid, 256>>>((float*)buf, 1); }), 4);
  rep("mma tf32 m16n8k4, 1 chain", time_ms([&] { k_mma<2, 1><<<grid, 256>>>((float*)buf, 1); }), 1);
  rep("mma tf32 m16n8k4, 4 chains", time_ms([&] { k_mma<2, 4><<<grid, 256>>>((float*)buf, 1); }), 4);
  // one warp per SMSP, one dependent chain: latency per instruction
  rep("mma tf32 m16n8k8 latency (1 warp/SMSP)", 8.0 * time_ms([&] { k_mma<0, 1><<<sms, 128>>>((float*)buf, 1); }), 1);
  rep("DFMA latency (1 warp/SMSP, 1 chain)", 8.0 * time_ms([&] { k_mix<1, 0, 0><<<sms, 128>>>((double*)buf, 1.0001, 0.5, 1.0001f, 0.5f); }), 4);
  rep("FFMA latency (1 warp/SMSP, 1 chain)", 8.0 * time_ms([&] { k_mix<0, 1, 0><<<sms, 128>>>((double*)buf, 1.0001, 0.5, 1.0001f, 0.5f); }), 4);
  // does one warp saturate the FP64 pipe?  W warps per SMSP, 8 independent chains each
  rep("DFMA x8, 1 warp/SMSP (per instr)", 8.0 * time_ms([&] { k_mix<8, 0, 0><<<sms, 128>>>((double*)buf, 1.0001, 0.5, 1.0001f, 0.5f); }), 32);
  rep("DFMA x8, 2 warps/SMSP (per instr)", 4.0 * time_ms([&] { k_mix<8, 0, 0><<<sms, 256>>>((double*)buf, 1.0001, 0.5, 1.0001f, 0.5f); }), 32);
  rep("DFMA x8, 4 warps/SMSP (per instr)", 2.0 * time_ms([&] { k_mix<8, 0, 0><<<sms * 2, 256>>>((double*)buf, 1.0001, 0.5, 1.0001f, 0.5f); }), 32);
  rep("FFMA x16, 1 warp/SMSP (per instr)", 8.0 * time_ms([&] { k_mix<0, 16, 0><<<sms, 128>>>((double*)buf, 1.0001, 0.5, 1.0001f, 0.5f); }), 64);
  rep("DFMA x8 + FFMA x16, 1 warp/SMSP (per instr of 24)", 8.0 * time_ms([&] { k_mix<8, 16, 0><<<sms, 128>>>((double*)buf, 1.0001, 0.5, 1.0001f, 0.5f); }), 96);
  rep("DFMA x8 + FFMA x16, 2 warps/SMSP (per instr of 24)", 4.0 * time_ms([&] { k_mix<8, 16, 0><<<sms, 256>>>((double*)buf, 1.0001, 0.5, 1.0001f, 0.5f); }), 96);
  rep("DFMA x8", time_ms([&] { k_mix<8, 0, 0><<<grid, 256>>>((double*)buf, 1.0001, 0.5, 1.0001f, 0.5f); }), 32);
  rep("FFMA x16", time_ms([&] { k_mix<0, 16, 0><<<grid, 256>>>((double*)buf, 1.0001, 0.5, 1.0001f, 0.5f); }), 64);
  rep("FFMA2 x8 (per packed instr)", time_ms([&] { k_mix<0, 0, 8><<<grid, 256>>>((double*)buf, 1.0001, 0.5, 1.0001f, 0.5f); }), 32);
  rep("DFMA x8 + FFMA x16 (per instr of 24)", time_ms([&] { k_mix<8, 16, 0><<<grid, 256>>>((double*)buf, 1.0001, 0.5, 1.0001f, 0.5f); }), 96);
  rep("DFMA x8 + FFMA x8 (per instr of 16)", time_ms([&] { k_mix<8, 8, 0><<<grid, 256>>>((double*)buf, 1.0001, 0.5, 1.0001f, 0.5f); }), 64);
  rep("DFMA x4 + FFMA x16 (per instr of 20)", time_ms([&] { k_mix<4, 16, 0><<<grid, 256>>>((double*)buf, 1.0001, 0.5, 1.0001f, 0.5f); }), 80);
  rep("DFMA x8 + FFMA2 x8 (per instr of 16)", time_ms([&] { k_mix<8, 0, 8><<<grid, 256>>>((double*)buf, 1.0001, 0.5, 1.0001f, 0.5f); }), 64);
  rep("cvt f32->f64->f32 (per pair)", time_ms([&] { k_cvt<0><<<grid, 256>>>((double*)buf, 1.f); }), 8);
  rep("cvt f64->f32 + DADD (per pair)", time_ms([&] { k_cvt<1><<<grid, 256>>>((double*)buf, 1.f); }), 8);
  rep("cvt.rna.tf32 + FADD (per pair)", time_ms([&] { k_cvt<2><<<grid, 256>>>((double*)buf, 1.f); }), 8);
  return 0;
}
