// The whole device-resident LM loop of the single-camera path as ONE launch (included by intr.cu).
//
// The multi-launch loop (elimination, reduced solve, sweep, accept/reject per iteration, intr.cu) spends
// 35-50 us per iteration outside the two big kernels: four dependent launches, each of which starts by
// pulling the LM state out of global memory.  At the sizes a partitioned calibration leaves per GPU, and for
// small problems altogether, that is most of the time.  Here a persistent CTA keeps the LM state of its
// problem in shared memory and runs the phases back to back: one problem, one CTA per work item (cooperative
// launch, all CTAs resident).  Per-item partials go through global memory; after a grid barrier EVERY CTA reduces
// them in the same fixed order and advances its own, bit-identical copy of the state machine - no broadcast, no
// second barrier.  A partitioned problem (shared = 1) lets CTA 0 run the peer-memory allreduce and hand the sums
// to the other CTAs through global memory and one more barrier.
// Used for small problems only (intr.cu, build_problem: measured cross-over at a few million observations).  A
// batch variant (a CTA per problem at a time, block barriers only) was measured too: 8.0 vs 6.0 ms on C4 - 512
// threads are too many for 50 views, the multi-launch loop keeps the whole GPU busy across problems - and dropped.
// The phases reuse the device functions of the multi-launch path (sw_sweep_item, schur_warp_pass and the
// logic of decide_tail / solve_tail), so both paths produce the same numbers: the same reductions in the
// same order.  The registers are split once, at the top (FP64 warpgroups 152, Jacobian warpgroups 104, as in the
// sweep), and the whole loop is instantiated once per role, so that no instruction is shared between the two
// budgets (ptxas allocates registers per setmaxnreg region; it crashed on a kernel that changed the split per
// phase).  The FP64 warpgroups also run the elimination (a lane per view) and the serial parts of the small phases.
#pragma once

namespace ccg {

struct LmCtl {                 // per problem object, zeroed before every launch
  unsigned long long bar;      // arrivals at the grid barrier (monotonic)
  int abort;                   // a barrier timed out: everybody leaves
  int pad_;
};

constexpr int kLmLeader = kSwPairs * 32;  // the thread that runs the serial parts: the first FP64-role thread

struct LmCta {  // shared memory: the LM state of the problem this CTA works on, and scratch of the small phases
  double head[kLmHeadWords];  // LmState up to the trace ring
  double intr[2][9], du[9], U[81], gu[9];
  double buf[kRedSeg * kSweepPart];
  double S[81], rhs[9], Du[9], x[9], k[9], step[2];
  double red[2][kSwWarps];
  double schur[kSwPairs][kSchurPart];
  unsigned long long epoch;
  int flag[2];
};

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// all CTAs of the (cooperative) launch; false when the barrier was abandoned.  IS_D (here and below): the
// instantiation for the FP64-role warps, which hold the leader thread - the serial parts exist only there.
template <bool IS_D>
__device__ __forceinline__ bool lm_grid_barrier(LmCtl* ctl, LmCta& s, int n_ctas) {
  __syncthreads();
  if (n_ctas == 1) return true;  // (a one-item problem: the block barrier is the grid barrier)
  if (IS_D && threadIdx.x == kLmLeader) {
    __threadfence();
    s.epoch += (unsigned long long)n_ctas;
    atomicAdd(&ctl->bar, 1ull);
    const long long t0 = clock64();
    int ok = 1;
    while (ld_acquire_u64(&ctl->bar) < s.epoch) {
      if (*reinterpret_cast<volatile int*>(&ctl->abort) || clock64() - t0 > 8000000000ll) {  // ~4 s
        ctl->abort = 1;
        ok = 0;
        break;
      }
    }
    __threadfence();
    s.flag[1] = ok;
  }
  __syncthreads();
  return s.flag[1] != 0;
}

// accept / reject on the CTA's copy of the state; r = the reduced sweep partial (shared memory)
template <bool IS_D>
__device__ __forceinline__ void lm_decide_cta(LmCta& s, const LmParams& P, bool evaluated, const double* r, IntrState& ps,
                                              int* n_done) {
  LmState& lm = *reinterpret_cast<LmState*>(s.head);
  if (IS_D && threadIdx.x == kLmLeader) {
    bool accept = false;
    if (!evaluated) {
      lm_reject(lm, P, n_done);
      lm.solve_ok = 1;
      if (lm.status == 0 && lm.iterations >= P.max_iterations) lm_finish(lm, 4, n_done, true);
    } else if (r[kSweepPart - 1] < 0.0) {
      lm_finish(lm, CCG_ERR_NCCL, n_done, false);
    } else {
      // the trace row goes to global memory (every CTA of the problem writes the same values)
      accept = lm_decide(lm, P, 0.5 * r[0], r[7] + lm.gTd_u, r[8] + lm.dDd_u, n_done, ps.lm.hist);
      if (accept) {
        lm.n_obs = r[3];
        lm.sumsq = r[0];
      }
    }
    s.flag[0] = accept ? 1 : 0;
  }
  __syncthreads();
  if (s.flag[0]) {
    for (int e = threadIdx.x; e < 90; e += blockDim.x) {
      if (e < 81) {
        const int a = e / 9, b = e - 9 * a;
        const int idx = u_partial_index(a < b ? a : b, a < b ? b : a);
        s.U[e] = idx >= 0 ? r[idx] : 0.0;
      } else {
        s.gu[e - 81] = r[16 + e - 81];
      }
    }
  }
  __syncthreads();
}

// reduced 9 x 9 solve on the CTA's copy of the state; r = the reduced elimination partial (shared memory)
template <bool IS_D>
__device__ __forceinline__ void lm_solve_cta(LmCta& s, const LmParams& P, const double* r, unsigned mask, int* n_done) {
  LmState& lm = *reinterpret_cast<LmState*>(s.head);
  if (IS_D && threadIdx.x == kLmLeader) {
    int go = 0;
    if (r[kSchurPart - 1] < 0.0) {
      lm.finalize = 0;
      if (lm.status == 0) lm_finish(lm, CCG_ERR_NCCL, n_done, false);
    } else {
      double gDg = r[54];
      for (int k = 0; k < 9; ++k)
        if (!((mask >> k) & 1u) && s.U[k * 9 + k] > 0.0) gDg += s.gu[k] * s.gu[k] / s.U[k * 9 + k];
      const double rnorm = sqrt(2.0 * lm.cost);
      lm.grad_norm = rnorm > 0.0 ? sqrt(gDg) / rnorm : 0.0;
      if (lm.finalize) {
        lm.finalize = 0;
      } else if (lm.cost == 0.0 || sqrt(gDg) <= P.gradient_tolerance * rnorm) {
        lm_finish(lm, 1, n_done, false);
      } else {
        lm.iterations++;
        if (r[55] > 0.0) lm.solve_ok = 0;
        else go = 1;
      }
    }
    s.flag[0] = go;
  }
  __syncthreads();
  if (!s.flag[0]) return;
  const double lambda = lm.lambda;
  for (int e = threadIdx.x; e < 90; e += blockDim.x) {
    if (e < 81) {
      const int a = e / 9, b = e - 9 * a;
      double v = s.U[e] - (a <= b ? r[up9(a, b)] : r[up9(b, a)]);
      if (a == b) v += lambda * clampd(s.U[e], P.min_diagonal, P.max_diagonal);
      if (((mask >> a) & 1u) || ((mask >> b) & 1u)) v = a == b ? 1.0 : 0.0;
      s.S[e] = v;
    } else {
      const int a = e - 81;
      s.Du[a] = clampd(s.U[a * 9 + a], P.min_diagonal, P.max_diagonal);
      s.rhs[a] = ((mask >> a) & 1u) ? 0.0 : -s.gu[a] + r[45 + a];
    }
  }
  __syncthreads();
  if (IS_D && threadIdx.x == kLmLeader) {
    double x[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) x[k] = s.rhs[k];
    const bool ok = chol_solve_reg<9>(s.S, x);
    if (ok) {
      const int par = lm.par;
      double gTd = 0.0, dDd = 0.0;
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const double dk = ((mask >> k) & 1u) ? 0.0 : x[k];
        s.du[k] = dk;
        s.intr[par ^ 1][k] = s.intr[par][k] + dk;
        gTd += s.gu[k] * dk;
        dDd += dk * dk * s.Du[k];
      }
      lm.gTd_u = gTd;
      lm.dDd_u = dDd;
      lm.solve_ok = 1;
    } else {
      lm.solve_ok = 0;
    }
  }
  __syncthreads();
}

// what a phase needs to know about the problem and the launch
struct LmJob {
  WorkItem it;
  int item, item0, item1, n_ctas, reduce_mode;
  unsigned mask;
  LmCtl* ctl;
  double* lm_schur_part;
};

// phase: sweep at the candidate point + accept / reject.  Returns false when a grid barrier was abandoned.
// IS_D: the calling warp's role (the loop is instantiated once per role).
template <bool IS_D>
__device__ __forceinline__ bool lm_sweep_decide(const IntrDev& d, const LmParams& P, const PeerComm& pc, const LmJob& job,
                                                LmCta& s, const SwSmem<kSwPairs>& sm, IntrState& ps) {
  LmState& lm = *reinterpret_cast<LmState*>(s.head);
  const int tid = threadIdx.x;
  const bool evaluate = lm.status == 0 && lm.solve_ok != 0;
  bool alive = true;
  if (evaluate) {
    const int par = lm.par;
    if (tid < 9) s.k[tid] = s.intr[par ^ 1][tid];
    __syncthreads();
    sw_sweep_item<false, IS_D ? 1 : 2, kSwPairs>(d, job.it, par, s.k, s.du, sm, s.red, s.step,
                                       d.sweep_part + (size_t)job.item * kSweepPart, nullptr);
    alive = lm_grid_barrier<IS_D>(job.ctl, s, job.n_ctas);
    if (alive) {
      if (job.reduce_mode == 2) {
        // CTA 0 sums the items and the ranks, the others pick the result up after one more barrier
        if (blockIdx.x == 0) {
          block_reduce_items<kSweepPart>(d.sweep_part, job.item0, job.item1, s.buf);
          peer_allreduce64(s.buf, pc);
          if (tid < kSweepPart) d.red_sweep[tid] = s.buf[tid];
        }
        alive = lm_grid_barrier<IS_D>(job.ctl, s, job.n_ctas);
        if (blockIdx.x != 0 && tid < kSweepPart) s.buf[tid] = __ldcg(d.red_sweep + tid);
        __syncthreads();
      } else {
        block_reduce_items<kSweepPart>(d.sweep_part, job.item0, job.item1, s.buf);
      }
    }
  }
  if (alive) lm_decide_cta<IS_D>(s, P, evaluate, s.buf, ps, d.n_done);
  return alive;
}

// phase: pose elimination (the FP64 warps, a lane per view) + reduced solve
template <bool IS_D>
__device__ __forceinline__ bool lm_schur_solve(const IntrDev& d, const LmParams& P, const PeerComm& pc, const LmJob& job,
                                               LmCta& s) {
  LmState& lm = *reinterpret_cast<LmState*>(s.head);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  {
    const int par = lm.par;
    const double lambda = lm.lambda;
    __syncthreads();
    if (IS_D) {
      double racc0 = 0.0, racc1 = 0.0;
      for (int64_t base = job.it.v0 + (warp - kSwPairs) * 32; base < job.it.v1; base += kSwPairs * 32)
        schur_warp_pass(d, P, par, lambda, base, (int)min((int64_t)32, job.it.v1 - base), lane, racc0, racc1);
      s.schur[warp - kSwPairs][lane] = racc0;
      s.schur[warp - kSwPairs][32 + lane] = racc1;
    }
    __syncthreads();
    if (tid < kSchurPart) {
      double acc = 0.0;
#pragma unroll
      for (int w = 0; w < kSwPairs; ++w) acc += s.schur[w][tid];
      job.lm_schur_part[(size_t)job.item * kSchurPart + tid] = acc;
    }
  }
  bool alive = lm_grid_barrier<IS_D>(job.ctl, s, job.n_ctas);
  if (alive) {
    if (job.reduce_mode == 2) {
      if (blockIdx.x == 0) {
        block_reduce_items<kSchurPart>(job.lm_schur_part, job.item0, job.item1, s.buf);
        peer_allreduce64(s.buf, pc);
        if (tid < kSchurPart) d.red_schur[tid] = s.buf[tid];
      }
      alive = lm_grid_barrier<IS_D>(job.ctl, s, job.n_ctas);
      if (blockIdx.x != 0 && tid < kSchurPart) s.buf[tid] = __ldcg(d.red_schur + tid);
      __syncthreads();
    } else {
      block_reduce_items<kSchurPart>(job.lm_schur_part, job.item0, job.item1, s.buf);
    }
  }
  if (alive) lm_solve_cta<IS_D>(s, P, s.buf, job.mask, d.n_done);
  return alive;
}

// the LM loop of the problem, as run by the warps of one role
template <bool IS_D>
__device__ __forceinline__ void lm_role_loop(const IntrDev& d, const LmParams& P, const PeerComm& pc, int reduce_mode,
                                             LmCtl* ctl, double* lm_schur_part, LmCta& s, const SwSmem<kSwPairs>& sm) {
  const int tid = threadIdx.x;
  IntrState& ps = d.ps[0];
  LmJob job;
  job.item = (int)blockIdx.x;  // blockIdx.x = work item of the problem
  job.it = d.items[job.item];
  job.item0 = ps.item0;
  job.item1 = ps.item1;
  job.n_ctas = gridDim.x;
  job.reduce_mode = reduce_mode;
  job.mask = ps.mask;
  job.ctl = ctl;
  job.lm_schur_part = lm_schur_part;
  LmState& lm = *reinterpret_cast<LmState*>(s.head);
  // ---- the problem's state (k_intr_init_state wrote it) into shared memory
  if (tid < kLmHeadWords) s.head[tid] = reinterpret_cast<const double*>(&ps.lm)[tid];
  for (int e = tid; e < 81 + 9 + 9 + 18; e += kSwThreads) {
    if (e < 81) s.U[e] = ps.U[e];
    else if (e < 90) s.gu[e - 81] = ps.gu[e - 81];
    else if (e < 99) s.du[e - 90] = ps.du[e - 90];
    else (&s.intr[0][0])[e - 99] = (&ps.intr[0][0])[e - 99];
  }
  __syncthreads();
  // ---- the LM loop (the sequence ccg_intrinsics_run enqueues in the multi-launch path)
  bool alive = true;
  for (int guard = 0; guard <= 2 * P.max_iterations + 64; ++guard) {
    alive = lm_sweep_decide<IS_D>(d, P, pc, job, s, sm, ps);  // (the first pass evaluates the initial point)
    if (!alive || !(lm.status == 0 || lm.finalize)) break;
    alive = lm_schur_solve<IS_D>(d, P, pc, job, s);
    if (!alive || lm.status != 0) break;  // finished: that was the final gradient-norm pass, or the gradient test fired
  }
  __syncthreads();
  if (IS_D && !alive && tid == kLmLeader && lm.status == 0) {
    lm.status = CCG_ERR_CUDA;  // a grid barrier was abandoned
    lm.finalize = 0;
  }
  __syncthreads();
  // ---- the state back to global memory (what ccg_intrinsics_download reads)
  if (blockIdx.x == 0) {
    if (tid < kLmHeadWords) reinterpret_cast<double*>(&ps.lm)[tid] = s.head[tid];
    for (int e = tid; e < 81 + 9 + 9 + 18; e += kSwThreads) {
      if (e < 81) ps.U[e] = s.U[e];
      else if (e < 90) ps.gu[e - 81] = s.gu[e - 81];
      else if (e < 99) ps.du[e - 90] = s.du[e - 90];
      else (&ps.intr[0][0])[e - 99] = (&s.intr[0][0])[e - 99];
    }
  }
}

// reduce_mode: 1 local sums, 2 the sums are all-reduced over peer memory (a partitioned problem)
__global__ void __launch_bounds__(kSwThreads, 1)
k_intr_lm(IntrDev d, LmParams P, PeerComm pc, int reduce_mode, LmCtl* ctl, double* lm_schur_part) {
  extern __shared__ __align__(16) unsigned char s_dyn[];
  __shared__ LmCta s;
  const SwSmem<kSwPairs> sm(s_dyn);
  if (threadIdx.x == kLmLeader) s.epoch = 0ull;
  if ((threadIdx.x >> 5) >= kSwPairs) {
    if (kSwRegD > kSwRegLaunch) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kSwRegD));
    lm_role_loop<true>(d, P, pc, reduce_mode, ctl, lm_schur_part, s, sm);
  } else {
    if (kSwRegJ < kSwRegLaunch) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kSwRegJ));
    lm_role_loop<false>(d, P, pc, reduce_mode, ctl, lm_schur_part, s, sm);
  }
}

}  // namespace ccg
