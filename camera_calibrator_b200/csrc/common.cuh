// Shared host/device helpers of the ccg library (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstddef>

#include <string>
#include <vector>

#include "../../include/ccg.h"

namespace ccg {

// --------------------------------------------------------------------------
// context
// --------------------------------------------------------------------------
struct NcclApi;  // dlopen'ed entry points (ctx.cu)

}  // namespace ccg

struct ccg_ctx {
  int device = 0;
  int sm_count = 0;
  int cc_major = 0, cc_minor = 0;
  size_t global_mem = 0;
  cudaStream_t stream = nullptr;
  std::string last_error;
  int64_t launches = 0;
  // multi-GPU
  int rank = 0, world = 1;
  void* nccl_comm = nullptr;
  // peer-memory mailboxes for the fused reduce + allreduce + decide/solve kernels (ctx.cu)
  void* mbox = nullptr;             // this rank's mailbox (cudaMalloc, exported through CUDA IPC)
  void* mbox_peer[16] = {nullptr};  // every rank's mailbox mapped into this process ([rank] == mbox)
  unsigned long long* d_seq = nullptr;  // device counter of completed peer allreduces
  bool peer_ok = false;
  // pinned scratch for small D2H polls
  int* h_poll = nullptr;
  // pinned staging arena for the per-view arrays a problem build uploads (poses, offsets, work items): packed by
  // the host while the observation arrays are already on the wire, then copied with real asynchronous DMA (a copy from
  // pageable memory waits for the stream and bounces through the driver's own staging buffer)
  void* h_stage = nullptr;
  size_t h_stage_bytes = 0;
  // event pool for opts.profile
  std::vector<cudaEvent_t> events;
  // stream-ordered device memory pool: freed problem buffers stay cached for the next solve
  // (cudaMalloc/cudaFree of the ~1 GB of a large calibration cost more than its whole LM loop)
  cudaMemPool_t pool = nullptr;
  // per-device opt-in of the default sweep's dynamic shared memory ([RES variant])
  bool sweep_attr[4] = {false, false, false, false};
  bool lm_attr = false;              // the same for the persistent LM kernel
};

namespace ccg {

int set_error(ccg_ctx* ctx, int status, const char* what, cudaError_t e = cudaSuccess);

#define CCG_CUDA(ctx, call)                                              \
  do {                                                                   \
    cudaError_t e__ = (call);                                            \
    if (e__ != cudaSuccess) return ::ccg::set_error((ctx), CCG_ERR_CUDA, #call, e__); \
  } while (0)

int nccl_allreduce_sum(ccg_ctx* ctx, double* buf, size_t count);
int nccl_allreduce_max(ccg_ctx* ctx, double* buf, size_t count);
int peer_resync(ccg_ctx* ctx, double* scratch);  // scratch: one device double

// --------------------------------------------------------------------------
// Allreduce of 64 doubles over peer memory (NVLink / NVSwitch), fused into the kernel that needs
// the sum.  Every rank owns a mailbox with one slot per source rank and parity; a rank publishes
// its 64 values into the slot (parity, own rank) of EVERY mailbox with plain peer stores, fences,
// raises the slot's sequence number, then spins on the sequence numbers of its own mailbox and
// sums the slots in rank order - so all ranks obtain bit-identical sums and the LM state machines
// stay in lock step without a broadcast.  Two parities suffice: a rank can start allreduce k+2
// only after every peer published k+1, i.e. after it finished reading k.
// --------------------------------------------------------------------------
constexpr int kPeerMaxWorld = 16;
constexpr int kPeerVals = 64;
struct PeerSlot {
  double data[kPeerVals];
  unsigned long long seq;
  unsigned long long pad[7];
};
// The rig's reduced system is 6C x 6C (plus right-hand side and two scalars): a second set of slots of kPeerBigVals
// doubles behind the small ones in the same mailbox allocation (same protocol, same sequence counter).
constexpr int kPeerBigVals = 96 * 96 + 96 + 32;
struct PeerBigSlot {
  double data[kPeerBigVals];
  unsigned long long seq;
  unsigned long long pad[7];
};
constexpr size_t kPeerSmallBytes = sizeof(PeerSlot) * 2 * kPeerMaxWorld;
constexpr size_t kPeerMailboxBytes = kPeerSmallBytes + sizeof(PeerBigSlot) * 2 * kPeerMaxWorld;
struct PeerComm {
  PeerSlot* box[kPeerMaxWorld];  // box[p] = rank p's mailbox: PeerSlot[2][kPeerMaxWorld]
  unsigned long long* seq;       // this rank's counter of completed allreduces
  int rank, world;
  __host__ __device__ PeerBigSlot* big(int p) const {
    return reinterpret_cast<PeerBigSlot*>(reinterpret_cast<unsigned char*>(box[p]) + kPeerSmallBytes);
  }
};
PeerComm peer_comm(const ccg_ctx* ctx);

// device memory from the context's pool, ordered on ctx->stream
cudaError_t pool_alloc(ccg_ctx* ctx, void** out, size_t bytes);
void* host_stage(ccg_ctx* ctx, size_t bytes);  // pinned, >= bytes, valid until the next call; nullptr if unavailable
void pool_free(ccg_ctx* ctx, void* ptr);

// --------------------------------------------------------------------------
// LM state machine shared by the single-camera and the rig solver.  It lives
// in global memory, one per problem, and is advanced only by device code: the
// host enqueues a fixed kernel sequence per iteration and never reads it back
// until the loop ends (it polls one counter every few iterations).
// --------------------------------------------------------------------------
struct LmParams {
  int max_iterations;
  double initial_radius, max_radius, min_radius;
  double min_relative_decrease;
  double min_diagonal, max_diagonal;
  double function_tolerance, gradient_tolerance, parameter_tolerance;
  double huber_a;
  double noise_factor;
  int hessian_fp64;
};

inline LmParams to_params(const ccg_opts& o) {
  LmParams p;
  p.max_iterations = o.max_iterations;
  p.initial_radius = o.initial_radius;
  p.max_radius = o.max_radius;
  p.min_radius = o.min_radius;
  p.min_relative_decrease = o.min_relative_decrease;
  p.min_diagonal = o.min_diagonal;
  p.max_diagonal = o.max_diagonal;
  p.function_tolerance = o.function_tolerance;
  p.gradient_tolerance = o.gradient_tolerance;
  p.parameter_tolerance = o.parameter_tolerance;
  p.huber_a = o.huber_a;
  p.noise_factor = o.noise_factor;
  p.hessian_fp64 = o.hessian_fp64;
  return p;
}

constexpr int kHist = 48;
constexpr int kBlindSteps = 2;  // sub-resolution (unvalidatable) steps taken before stopping

struct LmState {
  double cost;       // at the current (accepted) point
  double cost0;      // at the initial point
  double radius, nu, lambda;
  double gTd_u, dDd_u;  // reduced-parameter share of g'd and d'Dd (written by the reduced solve)
  double grad_norm;     // sqrt(g' D^-1 g)/sqrt(2 cost) at the current point
  double n_obs;         // observations (global over ranks)
  double sumsq;         // raw sum ||r||^2 at the current point (RMS)
  double meas_scale;    // max |measurement| (sets the rounding noise of the cost)
  int par;              // which of the two per-view buffers holds the current point
  int status;           // 0 = running, else termination code
  int first;            // the next sweep evaluates the initial point
  int solve_ok;         // the last linear solve produced a candidate
  int finalize;         // finished this iteration: one more gradient-norm pass is pending
  int iterations, accepted, rejected, sweeps;
  int blind;            // sub-resolution steps taken so far (at most kBlindSteps)
  int n_hist;
  // per-evaluation trace (ring of the last kHist sweeps): cost, candidate cost,
  // gain ratio, radius used, relative step norm, accepted flag
  double hist[kHist][6];
};

__host__ __device__ inline void lm_init(LmState& s, const LmParams& P) {
  s.cost = 0.0;
  s.cost0 = 0.0;
  s.radius = P.initial_radius;
  s.nu = 2.0;
  s.lambda = 1.0 / P.initial_radius;
  s.gTd_u = 0.0;
  s.dDd_u = 0.0;
  s.grad_norm = 0.0;
  s.n_obs = 0.0;
  s.sumsq = 0.0;
  s.par = 0;
  s.status = 0;
  s.first = 1;
  s.solve_ok = 1;
  s.finalize = 0;
  s.iterations = s.accepted = s.rejected = s.sweeps = 0;
  s.blind = 0;
  s.n_hist = 0;
}

#ifdef __CUDACC__

// relative cost noise tolerated when a final sub-resolution step is taken
constexpr double kCostNoise = 1e-12;

// 1/x to within ~1 ulp without the IEEE slow path: hardware seed (~20 bits) + two Newton steps.
// Used for the perspective division of the sweeps (the same function evaluates the current and
// the candidate cost, so LM's gain ratio is unaffected).
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = fma(fma(-x, r, 1.0), r, r);
  r = fma(fma(-x, r, 1.0), r, r);
  return r;
}
__device__ __forceinline__ float fast_rcpf(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ double clampd(double x, double lo, double hi) {
  return fmin(fmax(x, lo), hi);
}

__device__ inline void lm_finish(LmState& s, int code, int* n_done, bool need_finalize) {
  s.status = code;
  s.finalize = need_finalize ? 1 : 0;
  atomicAdd(n_done, 1);
}

__device__ inline void lm_reject(LmState& s, const LmParams& P, int* n_done) {
  s.rejected++;
  s.radius /= s.nu;
  s.nu *= 2.0;
  if (s.radius < P.min_radius) {
    lm_finish(s, 5, n_done, true);
  } else {
    s.lambda = 1.0 / s.radius;
  }
}

// Step-acceptance logic after a sweep evaluated the candidate.  cost_c is the
// candidate cost, gTd/dDd the full-model terms.  Returns true when the
// candidate becomes the current point (the caller then flips `par` data).
// `hist` = the trace ring the row goes to: s.hist, or the one in global memory when s is a staged copy of the
// state's head (everything before hist) in shared memory.
__device__ inline bool lm_decide(LmState& s, const LmParams& P, double cost_c, double gTd,
                                 double dDd, int* n_done, double (*hist)[6]) {
  bool accept = false;
  s.sweeps++;
  double* h = hist[s.n_hist % kHist];
  s.n_hist++;
  h[0] = s.cost;
  h[1] = cost_c;
  h[2] = 0.0;
  h[3] = s.radius;
  h[4] = 0.0;
  if (s.first) {
    s.first = 0;
    s.cost0 = cost_c;
    accept = true;
    if (!isfinite(cost_c)) lm_finish(s, CCG_ERR_NOT_FINITE, n_done, false);
  } else {
    const double rnorm = sqrt(2.0 * s.cost);
    const double model = 0.5 * (s.lambda * dDd - gTd);
    // rounding noise of the FP64 cost ~ eps/2 * |measurement| * ||r|| (each residual is a
    // difference of numbers of size meas_scale): below it the gain ratio is meaningless
    const double noise = P.noise_factor * 1.1102230246251565e-16 * s.meas_scale * rnorm;
    const bool tiny = model <= noise || sqrt(dDd) <= P.parameter_tolerance * rnorm;
    const double rho = (s.cost - cost_c) / model;
    h[2] = rho;
    h[4] = rnorm > 0.0 ? sqrt(dDd) / rnorm : 0.0;
    if (tiny || s.blind > 0) {
      // below the resolution of the FP64 cost: take the step unless the cost rises; one
      // more such polishing step follows (the Gauss-Newton step is still accurate there,
      // only its validation is not), then stop
      if (isfinite(cost_c) && cost_c <= s.cost * (1.0 + kCostNoise) + 4.0 * noise) {
        accept = true;
        s.accepted++;
      } else {
        s.rejected++;
      }
      s.blind++;
      if (!accept || s.blind >= kBlindSteps) lm_finish(s, 3, n_done, true);
    } else if (isfinite(cost_c) && model > 0.0 && rho > P.min_relative_decrease) {
      accept = true;
      s.accepted++;
      const double f = 2.0 * rho - 1.0;
      s.radius = fmin(P.max_radius, s.radius / fmax(1.0 / 3.0, 1.0 - f * f * f));
      s.nu = 2.0;
      s.lambda = 1.0 / s.radius;
      if (s.cost - cost_c <= P.function_tolerance * s.cost) lm_finish(s, 2, n_done, true);
    } else {
      lm_reject(s, P, n_done);
    }
  }
  h[5] = accept ? 1.0 : 0.0;
  if (accept) {
    s.par ^= 1;
    s.cost = cost_c;
  }
  if (s.status == 0 && s.iterations >= P.max_iterations) lm_finish(s, 4, n_done, true);
  return accept;
}

__device__ inline bool lm_decide(LmState& s, const LmParams& P, double cost_c, double gTd, double dDd, int* n_done) {
  return lm_decide(s, P, cost_c, gTd, dDd, n_done, s.hist);
}

// 8-byte words of LmState before the trace ring: what a CTA stages in shared memory to run the state machine
constexpr int kLmHeadWords = (int)(offsetof(LmState, hist) / 8);
static_assert(offsetof(LmState, hist) % 8 == 0, "LmState head");

// Cholesky solve of an N x N SPD system by ONE thread, entirely in registers (fully unrolled; N small and known
// at compile time): S row-major in (shared) memory, lower triangle read; b in, solution out.  One rsqrt per pivot,
// no divisions.  Returns false when a pivot is not positive.
template <int N>
__device__ __forceinline__ bool chol_solve_reg(const double* S, double (&b)[N]) {
  double L[N][N], dinv[N];
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) L[i][j] = S[i * N + j];
  bool ok = true;
#pragma unroll
  for (int j = 0; j < N; ++j) {
    double dj = L[j][j];
#pragma unroll
    for (int k = 0; k < j; ++k) dj = fma(-L[j][k], L[j][k], dj);
    ok = ok && (dj > 0.0) && isfinite(dj);
    const double r = rsqrt(dj);
    dinv[j] = r;
#pragma unroll
    for (int i = j + 1; i < N; ++i) {
      double v = L[i][j];
#pragma unroll
      for (int k = 0; k < j; ++k) v = fma(-L[i][k], L[j][k], v);
      L[i][j] = v * r;
    }
  }
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double v = b[i];
#pragma unroll
    for (int k = 0; k < i; ++k) v = fma(-L[i][k], b[k], v);
    b[i] = v * dinv[i];
  }
#pragma unroll
  for (int i = N - 1; i >= 0; --i) {
    double v = b[i];
#pragma unroll
    for (int k = i + 1; k < N; ++k) v = fma(-L[k][i], b[k], v);
    b[i] = v * dinv[i];
  }
  return ok;
}

// Called by all threads of a CTA with at least kPeerVals threads; s_vals[0..64) in shared memory is
// replaced by the sum over ranks.  A peer that does not show up within ~30 s poisons the sum with
// NaN (the LM state machine then stops with CCG_ERR_NOT_FINITE) instead of hanging the GPU.
__device__ inline void peer_allreduce64(double* s_vals, const PeerComm& pc) {
  __shared__ unsigned long long s_seq;
  __shared__ int s_timeout;
  const int tid = threadIdx.x;
  if (tid == 0) {
    s_seq = *pc.seq + 1ull;
    s_timeout = 0;
  }
  __syncthreads();
  const unsigned long long seq = s_seq;
  const int par = (int)(seq & 1ull);
  if (tid < kPeerVals) {
    const double val = s_vals[tid];
    for (int p = 0; p < pc.world; ++p) pc.box[p][par * kPeerMaxWorld + pc.rank].data[tid] = val;
  }
  __syncthreads();
  if (tid < pc.world) {
    // release: the barrier orders the CTA's data stores before this thread, its system-scope fence (cumulative)
    // before the flag - one fence per destination instead of one per value
    __threadfence_system();
    unsigned long long* flag = &pc.box[tid][par * kPeerMaxWorld + pc.rank].seq;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(seq) : "memory");
    // acquire: wait for source rank `tid` in my own mailbox
    const unsigned long long* mine = &pc.box[pc.rank][par * kPeerMaxWorld + tid].seq;
    const long long t0 = clock64();
    unsigned long long got = 0ull;
    while (true) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(got) : "l"(mine) : "memory");
      if (got == seq) break;
      if (clock64() - t0 > 60000000000ll) {
        s_timeout = 1;
        break;
      }
    }
  }
  __syncthreads();
  if (tid < kPeerVals) {
    double acc = 0.0;
    for (int src = 0; src < pc.world; ++src) acc += __ldcg(&pc.box[pc.rank][par * kPeerMaxWorld + src].data[tid]);
    // a peer that never showed up: poison the sums and raise the failure flag (slot 63 is unused by both
    // reduced-system layouts); the consumer stops with CCG_ERR_NCCL instead of treating NaN as a bad step
    s_vals[tid] = s_timeout ? (tid == kPeerVals - 1 ? -1.0 : nan("")) : acc;
  }
  __syncthreads();
  if (tid == 0) *pc.seq = seq;
}

// The same for n <= kPeerBigVals values in shared OR global memory (the rig's reduced systems), called by all
// threads of ONE CTA of any size.  Returns false (and poisons vals with NaN) when a peer did not show up.
__device__ inline bool peer_allreduce_n(double* vals, int n, const PeerComm& pc) {
  __shared__ unsigned long long s_seq_n;
  __shared__ int s_timeout_n;
  const int tid = threadIdx.x, nt = blockDim.x;
  if (tid == 0) {
    s_seq_n = *pc.seq + 1ull;
    s_timeout_n = 0;
  }
  __syncthreads();
  const unsigned long long seq = s_seq_n;
  const int par = (int)(seq & 1ull);
  for (int i = tid; i < n; i += nt) {
    const double val = vals[i];
    for (int p = 0; p < pc.world; ++p) pc.big(p)[par * kPeerMaxWorld + pc.rank].data[i] = val;
  }
  __syncthreads();
  if (tid < pc.world) {
    __threadfence_system();
    unsigned long long* flag = &pc.big(tid)[par * kPeerMaxWorld + pc.rank].seq;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(seq) : "memory");
    const unsigned long long* mine = &pc.big(pc.rank)[par * kPeerMaxWorld + tid].seq;
    const long long t0 = clock64();
    unsigned long long got = 0ull;
    while (true) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(got) : "l"(mine) : "memory");
      if (got == seq) break;
      if (clock64() - t0 > 60000000000ll) {
        s_timeout_n = 1;
        break;
      }
    }
  }
  __syncthreads();
  const bool ok = s_timeout_n == 0;
  for (int i = tid; i < n; i += nt) {
    double acc = 0.0;
    for (int src = 0; src < pc.world; ++src) acc += __ldcg(&pc.big(pc.rank)[par * kPeerMaxWorld + src].data[i]);
    vals[i] = ok ? acc : nan("");
  }
  __syncthreads();
  if (tid == 0) *pc.seq = seq;
  return ok;
}

// --------------------------------------------------------------------------
// warp-level transposing reduction: every lane holds 32 values a[0..32); on
// return lane L holds (in a[0]) the sum over all lanes of a[L].  31 shuffles
// instead of 32 x 5, fixed summation order (deterministic).
// --------------------------------------------------------------------------
__device__ __forceinline__ double warp_reduce32(double (&a)[32], int lane) {
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1) {
    const bool up = (lane & half) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const double send = up ? a[i] : a[i + half];
      const double keep = up ? a[i + half] : a[i];
      a[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
    }
  }
  return a[0];
}

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

// Reduce N lane-private accumulators; element e ends up in lane (e & 31),
// returned slot e >> 5.  out[] has (N+31)/32 entries.
template <int N>
__device__ __forceinline__ void warp_reduce_array(const double (&acc)[N], int lane,
                                                  double (&out)[(N + 31) / 32]) {
#pragma unroll
  for (int c = 0; c < (N + 31) / 32; ++c) {
    double tmp[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) tmp[i] = (c * 32 + i < N) ? acc[c * 32 + i] : 0.0;
    out[c] = warp_reduce32(tmp, lane);
  }
}

// In-place Cholesky of a 6x6 SPD matrix held in registers (lower triangle).
__device__ __forceinline__ bool chol6(double (&A)[6][6]) {
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double d = A[j][j];
#pragma unroll
    for (int k = 0; k < j; ++k) d -= A[j][k] * A[j][k];
    ok = ok && (d > 0.0) && isfinite(d);
    d = sqrt(d);
    A[j][j] = d;
    const double inv = 1.0 / d;
#pragma unroll
    for (int i = j + 1; i < 6; ++i) {
      double s = A[i][j];
#pragma unroll
      for (int k = 0; k < j; ++k) s -= A[i][k] * A[j][k];
      A[i][j] = s * inv;
    }
  }
  return ok;
}

__device__ __forceinline__ void chol6_solve(const double (&L)[6][6], double (&b)[6]) {
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double s = b[i];
#pragma unroll
    for (int k = 0; k < i; ++k) s -= L[i][k] * b[k];
    b[i] = s / L[i][i];
  }
#pragma unroll
  for (int i = 5; i >= 0; --i) {
    double s = b[i];
#pragma unroll
    for (int k = i + 1; k < 6; ++k) s -= L[k][i] * b[k];
    b[i] = s / L[i][i];
  }
}

// sin and cos without the out-of-line argument-reduction routine of the CUDA math library (a function call inside
// a setmaxnreg region makes ptxas drop the register split - and crash): Cody-Waite reduction by pi/2 with three
// constants (exact for |x| < 1e5; a step angle that large is a rejected step anyway) and the fdlibm kernels on
// [-pi/4, pi/4], ~1 ulp.
__device__ __forceinline__ void sincos_inline(double x, double* sn_out, double* cs_out) {
  const double k = rint(x * 6.36619772367581382433e-01);
  double r = fma(-k, 1.57079632673412561417e+00, x);
  r = fma(-k, 6.07710050650619224932e-11, r);
  r = fma(-k, 2.02226624879595063154e-21, r);
  const int q = (int)k & 3;
  const double z = r * r;
  double sp = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
  sp = fma(z, sp, 2.75573137070700676789e-06);
  sp = fma(z, sp, -1.98412698298579493134e-04);
  sp = fma(z, sp, 8.33333333332248946124e-03);
  sp = fma(z, sp, -1.66666666666666324348e-01);
  const double sn = fma(r * z, sp, r);
  double cp = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
  cp = fma(z, cp, -2.75573143513906633035e-07);
  cp = fma(z, cp, 2.48015872894767294178e-05);
  cp = fma(z, cp, -1.38888888888741095749e-03);
  cp = fma(z, cp, 4.16666666666666019037e-02);
  const double cs = fma(z * z, cp, fma(-0.5, z, 1.0));
  *sn_out = (q & 1) ? ((q & 2) ? -cs : cs) : ((q & 2) ? -sn : sn);
  *cs_out = (q & 1) ? ((q & 2) ? sn : -sn) : ((q & 2) ? -cs : cs);
}

// [ext] ceres::QuaternionManifold::Plus: x_plus = exp_q(delta) * x, delta = half
// rotation vector (reference selects it at calibrator.cpp:298,
// extrinsics_calibrator.cpp:186,196).
__device__ __forceinline__ void quat_plus(const double (&x)[4], double d0, double d1, double d2,
                                          double (&out)[4]) {
  const double n2 = d0 * d0 + d1 * d1 + d2 * d2;
  double qw, s;
  if (n2 > 0.0) {
    const double n = sqrt(n2);
    double sn, cs;
    sincos_inline(n, &sn, &cs);
    qw = cs;
    s = sn / n;
  } else {
    qw = 1.0;
    s = 1.0;
  }
  const double qa = s * d0, qb = s * d1, qc = s * d2;
  out[0] = qw * x[0] - qa * x[1] - qb * x[2] - qc * x[3];
  out[1] = qw * x[1] + qa * x[0] + qb * x[3] - qc * x[2];
  out[2] = qw * x[2] - qa * x[3] + qb * x[0] + qc * x[1];
  out[3] = qw * x[3] + qa * x[2] - qb * x[1] + qc * x[0];
}

// rotation matrix (row-major) of q/|q|
__device__ __forceinline__ void quat_to_rot(const double (&q)[4], double (&R)[9]) {
  const double inv = 1.0 / sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const double w = q[0] * inv, a = q[1] * inv, b = q[2] * inv, c = q[3] * inv;
  R[0] = 1.0 - 2.0 * (b * b + c * c);
  R[1] = 2.0 * (a * b - w * c);
  R[2] = 2.0 * (a * c + w * b);
  R[3] = 2.0 * (a * b + w * c);
  R[4] = 1.0 - 2.0 * (a * a + c * c);
  R[5] = 2.0 * (b * c - w * a);
  R[6] = 2.0 * (a * c - w * b);
  R[7] = 2.0 * (b * c + w * a);
  R[8] = 1.0 - 2.0 * (a * a + b * b);
}

#endif  // __CUDACC__

// --------------------------------------------------------------------------
// host utilities
// --------------------------------------------------------------------------
// Event-pair profiler for opts.profile: tags kernel classes, sums elapsed time
// per tag after the loop (events are recorded on the library's own stream).
struct Profiler {
  ccg_ctx* ctx = nullptr;
  bool on = false;
  struct Rec {
    int tag;
    size_t e0, e1;
  };
  std::vector<Rec> recs;
  size_t used = 0;
  size_t grab();
  void begin(int tag);
  void end();
  void collect(double out_ms[4], int out_n[4]);
};

// split `n_views` views of each problem into work items of at most
// `views_per_item` consecutive views; returns items as (problem, v0, v1, 0)
struct WorkItem {
  int problem, v0, v1, pad;
};

}  // namespace ccg
