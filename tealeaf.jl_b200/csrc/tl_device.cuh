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
struct SolveState {
  StopCfg cfg;
  int iter;             // completed CG-type iterations (CG.jl:18 `tt`, PPCG outer included)
  int cheby_step;       // completed Chebyshev kernels (init counts as step 0 -> 1)
  int cheby_done;       // (unused: the stop rule is a pure function of the state)
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
};

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
__device__ __forceinline__ bool tl_grid_sum(double (&v)[NV], double *partials, unsigned *counter, double *sm) {
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
#pragma unroll
  for (int q = 0; q < NV; q++) {
    double t = 0.0;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) t += __ldcg(&partials[(size_t)q * gridDim.x + b]);
    v[q] = tl_block_sum(t, sm);
  }
  if (threadIdx.x == 0) *counter = 0u;
  return true;
}

__device__ __forceinline__ double2 tl_ld2(const double *p) { return __ldg(reinterpret_cast<const double2 *>(p)); }
__device__ __forceinline__ double2 tl_ld2_rw(const double *p) { return *reinterpret_cast<const double2 *>(p); }
__device__ __forceinline__ void tl_st2(double *p, double2 v) { *reinterpret_cast<double2 *>(p) = v; }

#endif  // __CUDACC__
