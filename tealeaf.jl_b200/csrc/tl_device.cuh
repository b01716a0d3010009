// tl_device.cuh -- geometry, device-resident solve state and reduction helpers shared by
// every kernel of libtealeaf_b200 (sm_100a, FP64, HBM-bound: no tensor cores on this path).
//
// Device layout of one Chunk field (src/chunk.jl:25-38 holds them as Julia (x,y) matrices):
//   rows   j in [-hd, ny+hd)   (y, the strided dimension)
//   cols   i in [-hd, nx+hd)   (x, contiguous)
//   element (i,j) lives at f0[j*pitch + i], where f0 = slab + hd*pitch + TL_XPAD is the address
//   of interior cell (0,0).  TL_XPAD = 16 doubles, so f0 and every row start of the interior
//   are 128-byte aligned whatever halo_depth is, and pitch is a multiple of 16 doubles.
//   Reference index [kk,jj] (1-based, halos included) <-> (i,j) = (kk-1-hd, jj-1-hd).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define TL_XPAD 16
#define TL_PHYS_LEFT 1
#define TL_PHYS_RIGHT 2
#define TL_PHYS_BOTTOM 4
#define TL_PHYS_TOP 8

struct Geo {
  int nx, ny, hd;  // interior cells and halo depth of this tile
  int pitch;       // doubles per row
  int phys;        // TL_PHYS_* bits: sides that are physical (reflective) boundaries
};

// How the fused kernels cut the tile: one warp marches over `rows_per_chunk` rows of a
// 64-cell-wide strip (2 cells per lane, 16-byte vector accesses).
struct Tiling {
  int nstrips, nchunks, rows_per_chunk;
};

#define TL_CONV_SQRT 0  // CG.jl:24      sqrt(abs(error)) < eps
#define TL_CONV_ABS 1   // Cheby.jl:57, PPCG.jl:51   abs(error) < eps

// Stop rule of the CG-type iteration loops, evaluated on the device by every kernel at entry
// (so kernels launched ahead of the host's knowledge become no-ops once the solve stopped).
struct StopCfg {
  int max_iters;        // `for tt = 1:maxiters`
  int conv_mode;        // TL_CONV_*
  int switch_min;       // leave CG for Cheby/PPCG when iter >= switch_min && rr < switch_thresh
  int first_it;         // iteration count at which this phase started (0 for CG)
  double eps;
  double switch_thresh; // ERROR_SWITCH_MAX (kernels.jl:8) or epslim
};

// Device-resident scalars of a solve.  Written only by the last block of a kernel (all other
// blocks of that kernel have finished by then), read by every block of later kernels.
#define TL_MAX_RANKS 16
struct SolveState {
  StopCfg cfg;
  int iter;             // completed CG-type iterations (CG.jl:18 `tt`, PPCG outer included)
  int cheby_step;       // completed Chebyshev kernels (init counts as step 0 -> 1)
  int cheby_pairs;      // two-iteration Chebyshev kernels executed in this solve (buffer parity, tl_cheby_u_parity)
  int cheby_est;        // Cheby.calciter estimate (uploaded by the host after the first step)
  int cheby_tt0;        // outer iteration number `tt` of Chebyshev step 1
  int cheby_max_tt;     // maxiters
  int inner_steps;      // PPCG inner steps per outer iteration
  int inner_pp;         // PPCG: inner steps completed in the current outer iteration
  unsigned counter;     // last-block ticket
  unsigned pad0;
  double red_pw;        // sum(p.w) of the latest matvec (after the allreduce when tiled)
  double red_rr;        // sum(r.r) after `iter` iterations (rro when iter == first_it)
  double red_norm;      // Chebyshev: latest sum(r.r);   cheby_init: bb = sum(u0.u0)
  // this tile's parts: the kernels write *_local, the (out-of-place) allreduce publishes the
  // global value -- idempotent when a launched-ahead kernel was a no-op
  double red_pw_local, red_rr_local, red_norm_local;
  double red_aux[4];    // field summary / norm2 results
  double theta;
  double eps_cheby;
  // tile exchange (multi-GPU): number of mailbox exchanges this context has performed since it
  // was created (never reset: every tile performs the same sequence), and a sticky error flag
  // raised when a neighbour did not answer within TL_XCHG_TIMEOUT_NS
  unsigned long long xseq;
  int comm_error;
  int sched_error;      // internal assertion: a kernel of the lazy-u CG loop ran at an iteration of the wrong parity
  // tile_barrier(): an out-of-place all-tiles sum of `barrier_zero` (never written: stays 0.0) into
  // `barrier_out`, so that a rendezvous never accumulates into a slot somebody else uses
  double barrier_zero, barrier_out;
  // micro-profile of the kernel boundaries (option "prof"; tl_get_option "prof_*"): globaltimer stamps taken by
  // block 0 at kernel entry and by the LAST block around the grid sum, the fence.sys and the tile exchange
  // split tile exchange (CommDev::deferred): the exchange the previous kernel POSTED in its tail and that nobody has
  // collected yet -- number, which scalar it carries (TL_T_*), and the number of the last exchange block 0 of a later
  // kernel has published into red_pw / red_rr / red_norm
  unsigned long long pend_seq;
  unsigned long long pub_lo, pub_hi;   // the last COLLECTED exchange as an LL packet (number | half of the total): block 0 of the
                                       // collecting kernel publishes it, the other blocks read it with two loads and no fence
  int pend_target, pad3;
  // diagnostics of the last blocking exchange: where thread q sent its packet and what the slot it waited on held
  // when it gave up (reported with a time-out: tl_api.cu check_comm)
  unsigned long long dbg_dst[TL_MAX_RANKS], dbg_seen[TL_MAX_RANKS];
  int prof, pad2;
  unsigned long long prof_start, prof_prev_end;
  unsigned long long prof_acc[6];   // ns: [0] entry -> last block in the tail, [1] ticket + partial sums, [2] tile exchange,
                                    //     [3] previous kernel's end -> this entry (launch gap), [4] kernels counted, [5] fence.sys
};

// ---- multi-GPU: peer-mapped mailboxes and halo push targets --------------------------------
#define TL_XCHG_TIMEOUT_NS 10000000000ull
// One slot per (parity, sender), "LL" packets: each 8-byte word carries 32 bits of the double and
// the low 32 bits of the exchange number, so one (atomic) 8-byte store publishes data and flag
// together and no fence sits between them.
struct MailSlot { unsigned long long lo, hi; };
struct CommDev {
  int nranks, rank;
  MailSlot *mail[TL_MAX_RANKS];   // mail[r] = rank r's mailbox (2 x TL_MAX_RANKS slots), CUDA-IPC mapped
  int deferred, pad;              // 1: kernels post their packets in the tail and the NEXT kernel collects them at its entry
};
// which scalar of the SolveState an exchange carries
#define TL_T_NONE 0
#define TL_T_PW 1
#define TL_T_RR 2
#define TL_T_NORM 3
struct TlScal { double pw, rr, norm; };
// Where the edge cells of ONE buffer go: the same buffer of the neighbour tile on each
// tile-internal side (0 left, 1 right, 2 bottom, 3 top); f0 = interior origin, null on physical sides.
struct PushSide { double *f0; int pitch, nx, ny; };
struct Push { PushSide s[4]; };
// Depth-k exchanges (matrix-powers PPCG) also fill the corner blocks of the halo, so they address
// all EIGHT surrounding tiles: index (dy+1)*3 + (dx+1), dx/dy in {-1,0,1}; entry 4 (the tile
// itself) is unused; f0 is null where there is no tile.
struct Push8 { PushSide s[9]; };

__host__ __device__ inline bool tl_should_stop(int it, double rr, const StopCfg &c) {
  if (it >= c.max_iters) return true;
  if (it > c.first_it) {
    const double e = fabs(rr);
    if (c.conv_mode == TL_CONV_SQRT ? (sqrt(e) < c.eps) : (e < c.eps)) return true;
    if (it >= c.switch_min && rr < c.switch_thresh) return true;
  }
  return false;
}

#ifdef __CUDACC__

// Programmatic dependent launch (option use_pdl; both instructions are no-ops for a kernel launched without the
// programmatic attribute).  Mode 1 (round 1, measured slower): the dependents are released at kernel ENTRY, become
// resident at once and spin in griddepcontrol.wait next to the running CTAs.  Mode 2: the release is issued by every CTA
// when its rows are done, just before the kernel tail (tl_pdl_trigger) -- the next kernel's CTAs are scheduled into the
// slots that free up while the last block sums the partials and meets the other tiles, and the launch gap between two
// dependent kernels overlaps the tail.  Either way a dependent waits (griddepcontrol.wait) until the predecessor has
// completed and its writes are visible before it reads anything.
__device__ __forceinline__ void tl_pdl_entry() {
#ifdef TL_PDL_EARLY
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
__device__ __forceinline__ void tl_pdl_trigger() {
#ifndef TL_PDL_EARLY
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}

__device__ __forceinline__ unsigned long long tl_globaltimer();
// Kernel-entry stamp of the boundary micro-profile (SolveState::prof); free when profiling is off.
__device__ __forceinline__ void tl_prof_entry(SolveState *st) {
  if (blockIdx.x == 0 && threadIdx.x == 0 && st->prof) *(volatile unsigned long long *)&st->prof_start = tl_globaltimer();
}

__device__ __forceinline__ double tl_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Deterministic block sum (fixed shuffle tree, then warp 0 sums the warp totals in order).
// `sm` needs >= 32 doubles.  Result valid in thread 0.
__device__ __forceinline__ double tl_block_sum(double v, double *sm) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = tl_warp_sum(v);
  __syncthreads();
  if (lane == 0) sm[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (warp == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    t = lane < nw ? sm[lane] : 0.0;
    t = tl_warp_sum(t);
  }
  return t;
}

// Grid-wide deterministic sum: every block deposits its partial, the last block to arrive
// (ticket counter) adds all partials in a fixed order.  Returns true in ALL threads of the
// last block, with the total valid in thread 0 of that block.  No FP64 atomics, so results
// are run-to-run reproducible.  `NV` values are reduced at once.
template <int NV>
__device__ __forceinline__ bool tl_grid_sum(double (&v)[NV], double *partials, unsigned *counter, double *sm,
                                            bool sys_fence = false, unsigned long long *prof_fence_ns = nullptr) {
  __shared__ bool s_last;
#pragma unroll
  for (int q = 0; q < NV; q++) {
    double t = tl_block_sum(v[q], sm);
    if (threadIdx.x == 0) partials[(size_t)q * gridDim.x + blockIdx.x] = t;
  }
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned ticket = atomicAdd(counter, 1u);
    s_last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
  // tiles exchange after this sum: one thread makes everything this GPU wrote (the halo pushes
  // of all blocks, ordered before their tickets) visible system-wide while the others add up
  if (sys_fence && threadIdx.x == blockDim.x - 1) {
    if (prof_fence_ns) {
      unsigned long long t0;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
      __threadfence_system();
      unsigned long long t1;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
      *prof_fence_ns += t1 - t0;
    } else {
      __threadfence_system();
    }
  }
  // Partials in batches of 8 INDEPENDENT L2 loads per thread, added in the same fixed order as a plain strided loop
  // (a dependent load-add chain over 4096 partials cost ~5 us of every kernel: profiles/r02d_boundary_profile_n2.jsonl).
#pragma unroll
  for (int q = 0; q < NV; q++) {
    double t = 0.0;
    const double *pq = partials + (size_t)q * gridDim.x;
    for (unsigned base = 0; base < gridDim.x; base += 8u * blockDim.x) {
      double pv[8];
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const unsigned b = base + k * blockDim.x + threadIdx.x;
        pv[k] = b < gridDim.x ? __ldcg(pq + b) : 0.0;
      }
#pragma unroll
      for (int k = 0; k < 8; k++) t += pv[k];
    }
    v[q] = tl_block_sum(t, sm);
  }
  if (threadIdx.x == 0) *counter = 0u;
  return true;
}

// ---- tile exchange -------------------------------------------------------------------------
__device__ __forceinline__ void tl_st_relaxed_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long tl_ld_relaxed_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long tl_ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long tl_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// All-tiles sum of one double, executed by the LAST block of a kernel (all of its threads call
// this; `v` is valid in thread 0).  Thread q stores this tile's value into tile q's mailbox over
// NVLink (fire-and-forget peer stores), then waits for tile q's value to land in the local
// mailbox; thread 0 adds the values in rank order, so every tile gets the same bits.
//
// Ordering of the halo pushes: every block of this kernel ordered its peer stores before its
// ticket (bar.sync + fence.gpu + atomic), the last block observed all tickets and one of its
// threads executed fence.sys before the barrier that precedes the mailbox stores below -- the
// pattern cooperative groups' multi-grid sync uses.  A tile that has received our slot therefore
// also sees our pushed halo cells: the exchange is the halo-exchange completion barrier as well
// (and, being all-to-all, a full execution barrier: WAR hazards on the ping-pong buffers).
// Slots are double-buffered by the parity of the exchange number; exchange n+2 can only start
// after every tile finished reading exchange n.
__device__ __forceinline__ double tl_tile_exchange(const CommDev *cd, SolveState *st, double v, double *sm) {
  __shared__ unsigned long long s_seq;
  __syncthreads();
  if (threadIdx.x == 0) {
    sm[0] = v;
    s_seq = st->xseq + 1;
    st->xseq = s_seq;
  }
  __syncthreads();
  // The mailbox table is read with strong (L1-bypassing, system-scope) loads: it was written by a host copy, possibly
  // into memory that held an older context's table, and a stale cached pointer would send a packet astray.
  const int n = (int)tl_ld_relaxed_sys((const unsigned long long *)&cd->nranks) ;   // {nranks, rank} share one 8-byte word
  const int my_rank = (int)(tl_ld_relaxed_sys((const unsigned long long *)&cd->nranks) >> 32);
  const unsigned seq = (unsigned)s_seq;
  const int par = (int)(seq & 1u);
  if ((int)threadIdx.x < n) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(sm[0]);
    const unsigned long long w_lo = ((unsigned long long)seq << 32) | (bits & 0xffffffffull);
    const unsigned long long w_hi = ((unsigned long long)seq << 32) | (bits >> 32);
    MailSlot *dst = (MailSlot *)tl_ld_relaxed_sys((const unsigned long long *)&cd->mail[threadIdx.x]) + par * TL_MAX_RANKS + my_rank;
    tl_st_relaxed_sys(&dst->lo, w_lo);
    tl_st_relaxed_sys(&dst->hi, w_hi);
    st->dbg_dst[threadIdx.x] = (unsigned long long)(uintptr_t)dst;
    const MailSlot *src = (const MailSlot *)tl_ld_relaxed_sys((const unsigned long long *)&cd->mail[my_rank]) + par * TL_MAX_RANKS + threadIdx.x;
    unsigned long long lo, hi;
    unsigned spins = 0;
    unsigned long long t0 = 0;
    for (;;) {
      lo = tl_ld_relaxed_sys(&src->lo);
      hi = tl_ld_relaxed_sys(&src->hi);
      if ((unsigned)(lo >> 32) == seq && (unsigned)(hi >> 32) == seq) break;
      if ((++spins & 1023u) == 0u) {   // a neighbour that never arrives must not hang the GPU
        const unsigned long long t = tl_globaltimer();
        if (t0 == 0) t0 = t;
        else if (t - t0 > TL_XCHG_TIMEOUT_NS) {   // 1 + the rank not heard from
          st->comm_error = 1 + (int)threadIdx.x;
          st->dbg_seen[threadIdx.x] = (lo & 0xffffffff00000000ull) | (hi >> 32);   // exchange numbers found in the slot
          break;
        }
        // the packet is idempotent (same exchange number, same value): post it again while waiting
        dst = (MailSlot *)tl_ld_relaxed_sys((const unsigned long long *)&cd->mail[threadIdx.x]) + par * TL_MAX_RANKS + my_rank;
        tl_st_relaxed_sys(&dst->lo, w_lo);
        tl_st_relaxed_sys(&dst->hi, w_hi);
      }
    }
    sm[1 + threadIdx.x] = __longlong_as_double((long long)((lo & 0xffffffffull) | (hi << 32)));
    // acquire side, by the threads that observed the slots: later reads (the next kernel's halo loads) see the
    // neighbours' pushes that were ordered before those slots.  An acquire LOAD of the slot just observed, not a
    // fence.sys: the fence would also wait for this thread's own packet stores to be acknowledged across NVLink.
    (void)tl_ld_acquire_sys(&src->hi);
  }
  __syncthreads();
  double total = 0.0;
  if (threadIdx.x == 0)
    for (int r = 0; r < n; r++) total += sm[1 + r];
  return total;
}

// ---- split exchange: post in the tail, collect at the next kernel's entry ----------------------------------------
// The blocking exchange above makes the last block of a kernel wait for the slowest tile while every SM idles, and only
// then can the next kernel be launched.  Split form: the last block POSTS its packets and the kernel ends; every block
// of the NEXT kernel of the stream waits for the n packets at its entry and adds them in rank order (same bits on every
// tile and in every block), so the wait for the slowest tile overlaps the launch gap.  Block 0 also publishes the total
// into the SolveState for the host.  The guarantees are the blocking exchange's: a tile starts kernel k+1 only after
// every tile has finished kernel k (halo pushes complete, ping-pong buffers free); a slot of parity p is rewritten by
// exchange n+2 only after the receiver's kernel that collected exchange n has finished.
__device__ __forceinline__ bool tl_is_deferred(const CommDev *cd) { return cd != nullptr && cd->deferred != 0; }

// last block, all threads; v valid in thread 0.  Preceded by the fence.sys of the tail.
__device__ __forceinline__ void tl_tile_post(const CommDev *cd, SolveState *st, double v, int target, double *sm) {
  __shared__ unsigned long long s_pseq;
  __syncthreads();
  if (threadIdx.x == 0) {
    sm[0] = v;
    s_pseq = st->xseq + 1;
    st->xseq = s_pseq;
    st->pend_target = target;
    st->pend_seq = s_pseq;
  }
  __syncthreads();
  const int n = (int)tl_ld_relaxed_sys((const unsigned long long *)&cd->nranks);
  const int my_rank = (int)(tl_ld_relaxed_sys((const unsigned long long *)&cd->nranks) >> 32);
  const unsigned seq = (unsigned)s_pseq;
  if ((int)threadIdx.x < n) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(sm[0]);
    MailSlot *dst = (MailSlot *)tl_ld_relaxed_sys((const unsigned long long *)&cd->mail[threadIdx.x]) + (seq & 1u) * TL_MAX_RANKS + my_rank;
    tl_st_relaxed_sys(&dst->lo, ((unsigned long long)seq << 32) | (bits & 0xffffffffull));
    tl_st_relaxed_sys(&dst->hi, ((unsigned long long)seq << 32) | (bits >> 32));
  }
}

// Kernel entry, every thread: the scalars this kernel works with -- the SolveState's, with the total of the exchange
// the previous kernel posted patched in.  Fast path (all blocks but the earliest): block 0 has already collected the
// packets and published the total as an LL packet in the SolveState -- everything needed comes from ONE batch of
// independent loads, no fence, no polling.  Slow path (blocks that start before that): poll the n mailbox slots
// themselves (every block would have to wait for the slowest tile anyway) and add them in rank order -- the same bits.
__device__ __forceinline__ unsigned long long tl_ld_acquire_gpu(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ TlScal tl_entry_scalars(const CommDev *cd, SolveState *st) {
  TlScal S;
  if (!tl_is_deferred(cd)) {
    S.pw = st->red_pw; S.rr = st->red_rr; S.norm = st->red_norm;
    return S;
  }
  // one batch of independent loads (the tail of the previous kernel wrote pend_*; block 0 of THIS kernel may be writing pub_* / red_*)
  const unsigned long long pseq = *(volatile unsigned long long *)&st->pend_seq;
  const int target = *(volatile int *)&st->pend_target;
  const unsigned long long plo = tl_ld_relaxed_sys(&st->pub_lo);
  const unsigned long long phi = tl_ld_acquire_gpu(&st->pub_hi);   // acquire: the halo loads below are ordered after the publication
  S.pw = *(volatile double *)&st->red_pw; S.rr = *(volatile double *)&st->red_rr; S.norm = *(volatile double *)&st->red_norm;
  const unsigned seq = (unsigned)pseq;
  double total;
  if ((unsigned)(plo >> 32) == seq && (unsigned)(phi >> 32) == seq) {
    total = __longlong_as_double((long long)((plo & 0xffffffffull) | (phi << 32)));      // collected already: by block 0, or by an earlier kernel
  } else {
    const bool prof = st->prof != 0 && blockIdx.x == 0 && threadIdx.x == 0;
    const unsigned long long tp = prof ? tl_globaltimer() : 0;
    const int n = (int)tl_ld_relaxed_sys((const unsigned long long *)&cd->nranks);
    const int my_rank = (int)(tl_ld_relaxed_sys((const unsigned long long *)&cd->nranks) >> 32);
    const MailSlot *box = (const MailSlot *)tl_ld_relaxed_sys((const unsigned long long *)&cd->mail[my_rank]) + (seq & 1u) * TL_MAX_RANKS;
    total = 0.0;
    unsigned long long t0 = 0;
    for (int r = 0; r < n; r++) {          // warp-uniform addresses: one L2 transaction per warp and load
      unsigned long long lo, hi;
      unsigned spins = 0;
      for (;;) {
        lo = tl_ld_relaxed_sys(&box[r].lo);
        hi = tl_ld_relaxed_sys(&box[r].hi);
        if ((unsigned)(lo >> 32) == seq && (unsigned)(hi >> 32) == seq) break;
        if ((++spins & 1023u) == 0u) {
          const unsigned long long t = tl_globaltimer();
          if (t0 == 0) t0 = t;
          else if (t - t0 > TL_XCHG_TIMEOUT_NS) { st->comm_error = 1 + r; return S; }
        }
      }
      total += __longlong_as_double((long long)((lo & 0xffffffffull) | (hi << 32)));
    }
    (void)tl_ld_acquire_sys(&box[0].hi);   // acquire: the halo loads below see what the tiles pushed before posting
    if (blockIdx.x == 0 && threadIdx.x == 0) {     // publish: for the later blocks of this kernel, for later kernels, for the host
      if (target == TL_T_PW) st->red_pw = total;
      else if (target == TL_T_RR) st->red_rr = total;
      else if (target == TL_T_NORM) st->red_norm = total;
      __threadfence();
      const unsigned long long bits = (unsigned long long)__double_as_longlong(total);
      tl_st_relaxed_sys(&st->pub_lo, ((unsigned long long)seq << 32) | (bits & 0xffffffffull));
      tl_st_relaxed_sys(&st->pub_hi, ((unsigned long long)seq << 32) | (bits >> 32));
      if (prof) st->prof_acc[2] += tl_globaltimer() - tp;
    }
  }
  if (target == TL_T_PW) S.pw = total;
  else if (target == TL_T_RR) S.rr = total;
  else if (target == TL_T_NORM) S.norm = total;
  return S;
}

// Common end of a hot-loop kernel: grid-wide sum of acc[0] (do_sum) or just the last-block
// ticket, then -- when tiles exchange (cd != null) -- the all-tiles sum / barrier.  Returns true
// in thread 0 of the last block, with acc[0] = the (global) total.
// `target`: which SolveState scalar the sum is (TL_T_*).  With the split exchange (tl_is_deferred) acc[0] stays this
// tile's LOCAL sum: the caller must not publish it as the all-tiles value -- the next kernel's entry does.
__device__ __forceinline__ bool tl_kernel_tail(double (&acc)[1], bool do_sum, SolveState *st, double *partials,
                                               const CommDev *cd, double *sm, int target = TL_T_NONE) {
  const bool prof = st->prof != 0;
  unsigned long long t_in = 0;
  if (prof && threadIdx.x == 0) t_in = tl_globaltimer();
  tl_pdl_trigger();   // this CTA's rows are done: the next kernel of the stream may be scheduled (it waits for our completion)
  bool last;
  if (do_sum) {
    last = tl_grid_sum<1>(acc, partials, &st->counter, sm, cd != nullptr, prof ? &st->prof_acc[5] : nullptr);
  } else {
    __shared__ bool s_last_nosum;
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      s_last_nosum = (atomicAdd(&st->counter, 1u) == gridDim.x - 1);
      if (s_last_nosum) {
        const unsigned long long tf = prof ? tl_globaltimer() : 0;
        if (cd) __threadfence_system(); else __threadfence();
        if (prof) st->prof_acc[5] += tl_globaltimer() - tf;
        st->counter = 0u;
      }
    }
    __syncthreads();
    last = s_last_nosum;
    acc[0] = 0.0;
  }
  if (!last) return false;
  unsigned long long t_sum = 0;
  if (prof && threadIdx.x == 0) t_sum = tl_globaltimer();
  if (tl_is_deferred(cd)) tl_tile_post(cd, st, acc[0], do_sum ? target : TL_T_NONE, sm);
  else if (cd) acc[0] = tl_tile_exchange(cd, st, acc[0], sm);
  if (prof && threadIdx.x == 0) {
    const unsigned long long t_end = tl_globaltimer();
    const unsigned long long t_start = *(volatile unsigned long long *)&st->prof_start;
    if (t_start && t_in > t_start) st->prof_acc[0] += t_in - t_start;
    st->prof_acc[1] += t_sum - t_in;
    st->prof_acc[2] += t_end - t_sum;
    if (st->prof_prev_end && t_start > st->prof_prev_end) st->prof_acc[3] += t_start - st->prof_prev_end;
    st->prof_prev_end = t_end;
    st->prof_acc[4] += 1;
  }
  return threadIdx.x == 0;
}

__device__ __forceinline__ double2 tl_ld2(const double *p) { return __ldg(reinterpret_cast<const double2 *>(p)); }
// Coherent variants for data that other CTAs rewrite INSIDE the same launch (persistent kernels):
// ld.global.cg reads through L2 and never uses the non-coherent (read-only) path.
__device__ __forceinline__ double2 tl_ld2_cg(const double *p) {
  double2 v;
  asm volatile("ld.global.cg.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double tl_ld1_cg(const double *p) {
  double v;
  asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
template <bool COH>
__device__ __forceinline__ double tl_ld1(const double *p) { return COH ? tl_ld1_cg(p) : *p; }
template <bool COH>
__device__ __forceinline__ double tl_ldg1(const double *p) { return COH ? tl_ld1_cg(p) : __ldg(p); }
template <bool COH>
__device__ __forceinline__ double2 tl_ldg2(const double *p) { return COH ? tl_ld2_cg(p) : tl_ld2(p); }
__device__ __forceinline__ double2 tl_ld2_rw(const double *p) { return *reinterpret_cast<const double2 *>(p); }
__device__ __forceinline__ void tl_st2(double *p, double2 v) { *reinterpret_cast<double2 *>(p) = v; }

// L2 eviction-priority policies (createpolicy) and accesses that carry one.  Used by the CG
// kernels to keep r and w (the only operands re-touched by the NEXT kernel) resident in the
// 126 MB L2 while everything else streams through it.
#define TL_HINT_NORMAL 0
#define TL_HINT_FIRST 1
#define TL_HINT_LAST 2
__device__ __forceinline__ unsigned long long tl_policy(int kind) {
  unsigned long long p;
  if (kind == TL_HINT_FIRST) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  else if (kind == TL_HINT_LAST) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ double2 tl_ld2_hint(const double *p, unsigned long long pol) {
  double2 v;
  asm volatile("ld.global.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol) : "memory");
  return v;
}
__device__ __forceinline__ void tl_st2_hint(double *p, double2 v, unsigned long long pol) {
  asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(pol) : "memory");
}

#endif  // __CUDACC__
