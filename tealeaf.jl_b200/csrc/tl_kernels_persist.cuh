// tl_kernels_persist.cuh -- the hot path, part 3: the WHOLE CG iteration loop as one persistent
// cooperative kernel (CG.solve! CG.jl:7-29 / CG.mainstep! CG.jl:32-44).
//
// Why: at 4096^2 the two kernels of an iteration sit on the DRAM roof while they stream, but every
// kernel boundary costs a ramp-up, a last-block reduction over up to 4096 partials, a drain and a
// launch gap -- about 13 us of the 247 us iteration, and nearly everything on small meshes (256^2:
// 12.7 us per iteration for 0.4 us of HBM time).  Here the grid is launched ONCE per solve with
// exactly as many CTAs as are co-resident (cudaLaunchCooperativeKernel); the phases of an
// iteration are separated by a grid barrier instead of a kernel boundary, and alpha, beta and the
// stop rule are evaluated redundantly by every CTA from the same partial sums, so nothing ever
// travels to the host or waits for a "last block".
//
// Work distribution: the work items are the CTAs of the multi-wave kernels (8 warp tasks of
// `rows_per_chunk` rows each, tl_march_setup_blk) -- the row bodies are the very same functions
// (tl_cg_a_item / tl_cg_b_item).  Items are handed out dynamically through a queue counter in L2
// (the next index is fetched while the current item streams), which keeps the memory-contiguous
// issue order and the load balancing of the hardware CTA scheduler.
//
// Determinism / parity: every item deposits its block sum in partials[item]; after the barrier
// every CTA adds all partials in the fixed order of tl_grid_sum's last block.  The sums -- and
// therefore alpha, beta, the iteration count and every field -- are BIT-IDENTICAL to the
// two-kernel path (tests/test_gpu_parity.py::test_persistent_cg_is_bit_identical).
//
// Coherence: r, p, u, w are rewritten by other CTAs inside the launch, so this kernel never uses
// the non-coherent load path (COH = true row bodies: ld.global.cg / cp.async through L2); the
// barrier is release (fence + atomic arrive) / acquire (ld.acquire spin + fence) at gpu scope, and
// bar.sync extends it to the whole CTA -- the pattern of cooperative groups' grid.sync().
#pragma once
#include "tl_kernels_ring.cuh"

// Grid-barrier and work-queue words (one 128-byte line each, in the slab, zeroed before a launch).
struct PersistSync {
  unsigned count;   unsigned pad0[31];   // arrivals at the current barrier
  unsigned gen;     unsigned pad1[31];   // number of completed barriers
  unsigned qa[2];   unsigned pad2[30];   // phase A item queue, by iteration parity
  unsigned qb[2];   unsigned pad3[30];   // phase B item queue
  int abort;        unsigned pad4[31];   // a CTA waited longer than TL_BARRIER_TIMEOUT_NS: everybody leaves
};
#define TL_BARRIER_TIMEOUT_NS 4000000000ull

struct CgPersistParams {
  CgAParams A;
  CgBParams B;
  PersistSync *sync;
  double *part_a, *part_b;   // one partial sum per work item of phase A / B
  int nitems_a, nitems_b;
};

__device__ __forceinline__ unsigned tl_ld_acquire_gpu(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void tl_st_release_gpu(unsigned *p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// All CTAs of the (co-resident) grid.  `gen` is this CTA's count of completed barriers.  Returns
// false when the barrier was abandoned (timeout somewhere): the caller must leave the kernel.
__device__ __forceinline__ bool tl_grid_barrier(PersistSync *s, unsigned &gen) {
  __shared__ int s_ok;
  __syncthreads();
  if (threadIdx.x == 0) {
    int ok = 1;
    gen++;
    __threadfence();                                    // release: this CTA's writes (all threads: bar.sync above)
    if (atomicAdd(&s->count, 1u) == gridDim.x - 1) {
      s->count = 0u;                                    // ordered before the release store below
      tl_st_release_gpu(&s->gen, gen);
    } else {
      // relaxed polling (an acquire load would invalidate this SM's L1 on every probe while the
      // other CTAs of the SM still stream); the fence below is the acquire
      unsigned spins = 0;
      unsigned long long t0 = 0;
      while (*(volatile unsigned *)&s->gen != gen) {
        __nanosleep(32);
        if ((++spins & 1023u) == 0u) {
          if (*(volatile int *)&s->abort) { ok = 0; break; }
          const unsigned long long t = tl_globaltimer();
          if (t0 == 0) t0 = t;
          else if (t - t0 > TL_BARRIER_TIMEOUT_NS) { *(volatile int *)&s->abort = 1; ok = 0; break; }
        }
      }
    }
    __threadfence();                                    // acquire side for the rest of the CTA
    s_ok = ok;
  }
  __syncthreads();
  return s_ok != 0;
}

// Sum of partials[0..n) in the order of tl_grid_sum's last block; the total is returned to every thread.
__device__ __forceinline__ double tl_sum_partials_all(const double *partials, int n, double *sm) {
  __shared__ double s_total;
  double t = 0.0;
  for (int b = threadIdx.x; b < n; b += blockDim.x) t += __ldcg(&partials[b]);
  t = tl_block_sum(t, sm);
  if (threadIdx.x == 0) s_total = t;
  __syncthreads();
  return s_total;
}

template <int S, int MINB>
__global__ void __launch_bounds__(TL_FUSED_THREADS, MINB) k_cg_persist(const CgPersistParams P) {
  extern __shared__ __align__(128) unsigned char ring_raw[];
  __shared__ double sm[32];
  __shared__ int s_item;
  SolveState *st = P.A.st;
  PersistSync *sy = P.sync;
  const StopCfg cfg = st->cfg;
  int it = st->iter;
  double rr_cur = st->red_rr;
  double rr_prev = 0.0, pw_prev = 0.0;
  if (it != cfg.first_it) { rr_prev = P.A.hist_rr[it - 1]; pw_prev = P.A.hist_pw[it]; }
  unsigned gen = 0;                     // the host zeroes the sync block before every launch
  double *part_a = P.part_a, *part_b = P.part_b;
  const bool lead = blockIdx.x == 0 && threadIdx.x == 0;

  for (int n = 0;; n++) {
    if (tl_should_stop(it, rr_cur, cfg)) break;
    unsigned *qa = &sy->qa[n & 1], *qb = &sy->qb[n & 1];
    if (lead) {
      // the queues of the next iteration: last used one iteration ago, every fetch from them
      // happened before the barrier that ended that iteration
      sy->qa[(n + 1) & 1] = 0u;
      sy->qb[(n + 1) & 1] = 0u;
      P.A.hist_rr[it] = rr_cur;
    }
    // ---- phase A: [p = r + beta p ; u += alpha_prev p_old] ; w = A p ; pw = sum(p.w) ----
    CgAIter I;
    I.it = it;
    I.first = (it == cfg.first_it);
    I.beta = I.first ? 0.0 : rr_cur / rr_prev;
    I.alpha_prev = I.first ? 0.0 : rr_prev / pw_prev;
    if (threadIdx.x == 0) s_item = (int)atomicAdd(qa, 1u);
    __syncthreads();
    int item = s_item;
    while (item < P.nitems_a) {
      int next = 0;
      if (threadIdx.x == 0) next = (int)atomicAdd(qa, 1u);    // in flight while this item streams
      double acc = 0.0;
      tl_cg_a_item<TL_U_EVERY, S, true>(P.A, I, item, ring_raw, acc);
      const double t = tl_block_sum(acc, sm);
      if (threadIdx.x == 0) { part_a[item] = t; s_item = next; }
      __syncthreads();
      item = s_item;
    }
    if (!tl_grid_barrier(sy, gen)) return;
    const double pw = tl_sum_partials_all(part_a, P.nitems_a, sm);
    const double alpha = rr_cur / pw;
    if (lead) {
      st->red_pw_local = pw;
      st->red_pw = pw;
      P.B.hist_pw[it + 1] = pw;
    }
    // ---- phase B: r -= alpha w ; rrn = sum(r.r) ----
    if (threadIdx.x == 0) s_item = (int)atomicAdd(qb, 1u);
    __syncthreads();
    item = s_item;
    while (item < P.nitems_b) {
      int next = 0;
      if (threadIdx.x == 0) next = (int)atomicAdd(qb, 1u);
      double acc = 0.0;
      tl_cg_b_item_ring<(S * TL_RING_STAGE_BYTES) / TL_BRING_ROW_BYTES, true>(P.B, alpha, item, ring_raw, acc);
      const double t = tl_block_sum(acc, sm);
      if (threadIdx.x == 0) { part_b[item] = t; s_item = next; }
      __syncthreads();
      item = s_item;
    }
    if (!tl_grid_barrier(sy, gen)) return;
    const double rrn = tl_sum_partials_all(part_b, P.nitems_b, sm);
    rr_prev = rr_cur;
    pw_prev = pw;
    rr_cur = rrn;
    it++;
    if (lead) {
      st->red_rr_local = rrn;
      st->red_rr = rrn;
      st->iter = it;
    }
  }
}
