// tl_kernels_fused.cuh -- the hot path, part 1: work decomposition, parameter blocks and the
// pointwise kernels of the fused CG / Chebyshev / PPCG iterations (FP64, HBM-bound).  The
// stencil kernels are in tl_kernels_ring.cuh.
//
// Work decomposition ("warp-strip marching"): the tile is cut into 64-cell-wide column strips
// and row chunks; ONE WARP owns one (strip, chunk) and marches up its rows.  Each lane owns two
// adjacent cells (one 16-byte vector access per field per row, 512 B per warp per field: fully
// coalesced).  Warps are fully independent until the final reduction.
//
// Boundaries: on physical sides the reflective halo (kernels.jl:191-210, depth 1) is applied
// by index clamping instead of a separate halo pass, and the reflected value is written
// through to the halo cell so that memory holds what the reference's haloupdate! leaves.
// On tile-internal sides (multi-GPU) the kernel that PRODUCES a stencil operand also stores its
// edge cells straight into the neighbour tile's halo cells over NVLink (tl_push_edges); the
// all-tiles sum in the kernel tail (tl_tile_exchange) doubles as the completion barrier, so a
// multi-GPU iteration is the same kernel sequence as a single-GPU one: no halo kernels, no
// NCCL calls, no host involvement inside the iteration loops.
//
// Arithmetic is expression-for-expression the oracle's (no FMA contraction: -fmad=false).
#pragma once
#include "tl_device.cuh"

#define TL_FUSED_THREADS 256
#define TL_STRIP 64

struct MarchCtx {
  int lane, i0, j0, j1, ecol;
  bool acta, actb, ld_ok, has_edge;
};

// `blk` = the work item: 8 consecutive warp tasks (a CTA of the multi-wave kernels, or an item a
// persistent CTA fetched from the queue, tl_kernels_persist.cuh)
__device__ __forceinline__ bool tl_march_setup_blk(const Geo &g, const Tiling &t, MarchCtx &m, int blk, int reverse) {
  m.lane = threadIdx.x & 31;
  int wt = blk * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wt >= t.nstrips * t.nchunks) return false;
  if (reverse) wt = t.nstrips * t.nchunks - 1 - wt;
  const int s = wt % t.nstrips, q = wt / t.nstrips;
  m.j0 = q * t.rows_per_chunk;
  m.j1 = min(g.ny, m.j0 + t.rows_per_chunk);
  m.i0 = s * TL_STRIP + 2 * m.lane;
  m.acta = m.i0 < g.nx;
  m.actb = m.i0 + 1 < g.nx;
  m.ld_ok = m.i0 <= g.nx;  // the pair (nx, nx+1) is still inside the padded row
  m.ecol = (m.lane == 0) ? s * TL_STRIP - 1 : s * TL_STRIP + TL_STRIP;
  m.has_edge = (m.lane == 0) || (m.lane == 31 && m.ecol <= g.nx);
  return m.j0 < m.j1;
}
__device__ __forceinline__ bool tl_march_setup(const Geo &g, const Tiling &t, MarchCtx &m, int reverse = 0) {
  return tl_march_setup_blk(g, t, m, blockIdx.x, reverse);
}

// Stores this lane's cells (i0, i0+1) of row j into the neighbours' depth-1 halo cells when they
// lie on a tile-internal edge (plain peer stores; tl_tile_exchange explains their ordering).
__device__ __forceinline__ void tl_push_edges(const Push &ps, const Geo &g, const MarchCtx &m, int j, double2 v) {
  if (!m.acta) return;
  if (ps.s[0].f0 && m.i0 == 0)                         // my column 0 -> left tile's column nx_left
    ps.s[0].f0[(long)j * ps.s[0].pitch + ps.s[0].nx] = v.x;
  if (ps.s[1].f0) {                                    // my column nx-1 -> right tile's column -1
    if (m.i0 == g.nx - 1) ps.s[1].f0[(long)j * ps.s[1].pitch - 1] = v.x;
    else if (m.i0 + 1 == g.nx - 1) ps.s[1].f0[(long)j * ps.s[1].pitch - 1] = v.y;
  }
  if (ps.s[2].f0 && j == 0) {                          // my row 0 -> bottom tile's row ny_bottom
    double *d = ps.s[2].f0 + (long)ps.s[2].ny * ps.s[2].pitch + m.i0;
    if (m.actb) tl_st2(d, v); else d[0] = v.x;
  }
  if (ps.s[3].f0 && j == g.ny - 1) {                   // my row ny-1 -> top tile's row -1
    double *d = ps.s[3].f0 - ps.s[3].pitch + m.i0;
    if (m.actb) tl_st2(d, v); else d[0] = v.x;
  }
}

// Depth-d variant for the matrix-powers PPCG groups: stores this lane's INTERIOR cells (i0, i0+1)
// of interior row j into the halo cells of every surrounding tile (corner tiles included) whose
// depth-d halo contains them.  ina / inb: cell i0 / i0+1 is an interior cell of this tile.
// Rows of the bottom/top bands go out as whole 16-byte row segments, columns of the left/right
// bands as 8-byte cells.  Same ordering argument as tl_push_edges (tl_tile_exchange).
__device__ __forceinline__ void tl_push_deep(const Push8 &ps, const Geo &g, int d, int i0, bool ina, bool inb, int j,
                                             double2 v) {
  const bool yb = j < d, yt = j >= g.ny - d;
  const bool xl = i0 < d, xr = i0 + 1 >= g.nx - d;
  if (!(yb || yt || xl || xr) || !(ina || inb)) return;
#pragma unroll
  for (int dyi = 0; dyi < 3; dyi++) {
    if ((dyi == 0 && !yb) || (dyi == 2 && !yt)) continue;
#pragma unroll
    for (int dxi = 0; dxi < 3; dxi++) {
      if (dyi == 1 && dxi == 1) continue;
      const PushSide &t = ps.s[dyi * 3 + dxi];
      if (!t.f0) continue;
      const int tj = (dyi == 1) ? j : (dyi == 0 ? t.ny + j : j - g.ny);
      double *row = t.f0 + (long)tj * t.pitch;
      if (dxi == 1) {
        if (ina && inb) tl_st2(row + i0, v);
        else if (ina) row[i0] = v.x;
        else row[i0 + 1] = v.y;
      } else if (dxi == 0) {
        if (ina && i0 < d) row[t.nx + i0] = v.x;
        if (inb && i0 + 1 < d) row[t.nx + i0 + 1] = v.y;
      } else {
        if (ina && i0 >= g.nx - d) row[i0 - g.nx] = v.x;
        if (inb && i0 + 1 >= g.nx - d) row[i0 + 1 - g.nx] = v.y;
      }
    }
  }
}

// Depth-1 reflective halo of one field as a write-through (haloupdate!, kernels.jl:191-210)
__device__ __forceinline__ void tl_reflect_edges(double *f, const Geo &g, const MarchCtx &m, int j, long oc, double2 v) {
  if (!m.acta) return;
  if ((g.phys & TL_PHYS_LEFT) && m.i0 == 0) f[oc - 1] = v.x;
  if (g.phys & TL_PHYS_RIGHT) {
    if (m.i0 == g.nx - 1) f[oc + 1] = v.x;
    if (m.i0 + 1 == g.nx - 1) f[oc + 2] = v.y;
  }
  if ((g.phys & TL_PHYS_BOTTOM) && j == 0) { f[oc - g.pitch] = v.x; if (m.actb) f[oc - g.pitch + 1] = v.y; }
  if ((g.phys & TL_PHYS_TOP) && j == g.ny - 1) { f[oc + g.pitch] = v.x; if (m.actb) f[oc + g.pitch + 1] = v.y; }
}

// ------------------------------------------------------------------------------------------
// CG kernel A (k_cg_fused_w_ring, tl_kernels_ring.cuh):  [p = r + beta p ; u += alpha_prev p_old]
// (deferred from the previous iteration: CG.p! CG.jl:101-104 and the u half of CG.ur! CG.jl:95)
// then  w = A p , pw = sum(p.w)  (CG.w! CG.jl:82-90).  p is ping-ponged between p0/p1 because
// neighbouring warps still read the old p.  UPDATE_U = false is the PPCG outer variant (u is
// advanced by k_ppcg_ur_sd instead).
// HBM traffic per cell: read r, p, u, kx, ky; write p, u, w  = 64 B (48 B without u; TL_U_LAZY, the default of the CG
// loop, touches u every second launch only: 72 / 48 B, see tl_kernels_ring.cuh).
// ------------------------------------------------------------------------------------------
// How a launch of kernel A advances u (template parameter UM):
//   TL_U_NONE   not at all (PPCG outer: k_ppcg_ur_sd does it; the odd iterations of a lazy CG loop)        48 B per cell
//   TL_U_EVERY  u += alpha(it-1) p(it-1)                                                                    64 B
//   TL_U_LAZY   u = (u + alpha(it-2) p(it-2)) + alpha(it-1) p(it-1): p(it-2) still sits in the ping-pong
//               buffer this launch is about to overwrite, so it costs one more read of 8 B                  72 B
// A lazy CG loop (option cg_lazy_u) launches TL_U_LAZY on the even iterations of the phase and TL_U_NONE on the odd
// ones: 60 B on average.  Same operations in the same order as an update per iteration (CG.jl:95), hence the same bits.
// The host alternates the two kernels (chunks of an even number of iterations), each checks the parity it was given.
#define TL_U_NONE 0
#define TL_U_EVERY 1
#define TL_U_LAZY 2
// pending u updates after `off` executed iterations of a lazy phase (off >= 1): what k_cg_flush has to apply
__host__ __device__ inline int tl_cg_lazy_pending(int off) { return (off & 1) ? 1 : 2; }

struct CgAParams {
  Geo g; Tiling t;
  SolveState *st;
  double *hist_rr; const double *hist_pw;
  const double *r; double *p0; double *p1; double *u; const double *kx; const double *ky; double *w;
  double *partials;
  int single;   // 1: one tile -- the kernel publishes its sums itself
  int hint_keep, hint_stream;   // TL_HINT_*: L2 priority of r, w / of everything else
  const CommDev *cd;            // tiles exchange in the kernel tail (null: single tile or NCCL mode)
  Push push_p0, push_p1;        // halo targets of the p buffer being written
  int lazy_role;                // 0: u is advanced by this kernel every launch or not at all; lazy CG loop (tl_kernels_ring.cuh):
                                // 1 = a launch between two u updates (odd iterations), 2 = a launch that applies both (even)
};

// ------------------------------------------------------------------------------------------
// CG kernel B:  alpha = rro/pw ; r -= alpha w ; rrn = sum(r.r)   (the r half of CG.ur!,
// CG.jl:93-98; alpha CG.jl:35).  Pointwise, 24 B/cell.  The last block publishes rrn and
// advances the iteration counter.
// ------------------------------------------------------------------------------------------
struct CgBParams {
  Geo g; Tiling t;
  SolveState *st;
  double *hist_pw;
  double *r; const double *w;
  double *partials;
  int single;
  int hint_keep, hint_stream;   // L2 priority of r (re-read by the next kernel A) / of w (dead after this kernel)
  int reverse;                  // 1: walk the tile top-down, i.e. start where kernel A just finished (its tail is in L2)
  const CommDev *cd;
  Push push_r;
};

// The rows of one work item of kernel B.  COH = true (persistent kernel: r and w change inside the
// launch) keeps every load coherent (no ld.global.nc).
template <bool COH>
__device__ __forceinline__ void tl_cg_b_item(const CgBParams &P, double alpha, int blk, double &acc0) {
  const Geo g = P.g;
  double *__restrict__ r = P.r;
  const double *w = P.w;
  const unsigned long long pol_r = tl_policy(P.hint_keep), pol_w = tl_policy(P.hint_stream);
  const bool tiled = P.cd != nullptr;
  MarchCtx m;
  if (!tl_march_setup_blk(g, P.t, m, blk, P.reverse)) return;
  int j = m.j0;
  if (m.actb) {
    for (; j + 4 <= m.j1; j += 4) {
      double2 rv[4], wv[4];
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const long o = (long)(j + q) * g.pitch + m.i0;
        rv[q] = COH ? tl_ld2_cg(r + o) : tl_ld2_hint(r + o, pol_r);
        wv[q] = COH ? tl_ld2_cg(w + o) : tl_ld2_hint(w + o, pol_w);
      }
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const long o = (long)(j + q) * g.pitch + m.i0;
        rv[q].x = rv[q].x - alpha * wv[q].x;
        rv[q].y = rv[q].y - alpha * wv[q].y;
        tl_st2_hint(r + o, rv[q], pol_r);
        acc0 += rv[q].x * rv[q].x;
        acc0 += rv[q].y * rv[q].y;
        if (tiled) tl_push_edges(P.push_r, g, m, j + q, rv[q]);
      }
    }
  }
  for (; j < m.j1; j++) {
    const long o = (long)j * g.pitch + m.i0;
    double2 v = make_double2(0.0, 0.0);
    if (m.acta) { v.x = tl_ld1<COH>(r + o) - alpha * tl_ld1<COH>(w + o); r[o] = v.x; acc0 += v.x * v.x; }
    if (m.actb) { v.y = tl_ld1<COH>(r + o + 1) - alpha * tl_ld1<COH>(w + o + 1); r[o + 1] = v.y; acc0 += v.y * v.y; }
    if (tiled) tl_push_edges(P.push_r, g, m, j, v);
  }
}

__global__ void __launch_bounds__(TL_FUSED_THREADS, 4) k_cg_fused_r(const CgBParams P) {
  tl_pdl_entry();
  __shared__ double sm[32];
  SolveState *st = P.st;
  tl_prof_entry(st);
  const TlScal sc = tl_entry_scalars(P.cd, st);
  const int it = st->iter;
  const double rr_cur = sc.rr;
  if (st->comm_error || tl_should_stop(it, rr_cur, st->cfg)) return;
  const double pw = sc.pw;
  const double alpha = rr_cur / pw;
  if (blockIdx.x == 0 && threadIdx.x == 0) P.hist_pw[it + 1] = pw;
  double acc[1] = {0.0};
  tl_cg_b_item<false>(P, alpha, blockIdx.x, acc[0]);
  if (tl_kernel_tail(acc, true, st, P.partials, P.cd, sm, TL_T_RR)) {
    st->red_rr_local = acc[0];
    if ((P.single || P.cd != nullptr) && !tl_is_deferred(P.cd)) st->red_rr = acc[0];
    st->iter = it + 1;
  }
}

// ------------------------------------------------------------------------------------------
// CG flush: applies the updates kernel A deferred for the last executed iteration
// (u += alpha p ; p = beta p + r) including the depth-1 halo write-through, so that memory
// holds the reference's post-iteration state.  Pointwise.
// ------------------------------------------------------------------------------------------
// UM: how the loop's kernel A advanced u (TL_U_NONE / TL_U_EVERY / TL_U_LAZY);
// after a TL_U_LAZY loop one or two updates are pending, the older p still sits in the other ping-pong buffer.
template <int UM>
__global__ void __launch_bounds__(TL_FUSED_THREADS, 4) k_cg_flush(const CgAParams P) {
  SolveState *st = P.st;
  const int it = st->iter;
  if (it == st->cfg.first_it) return;
  const double rr_cur = st->red_rr, rr_prev = P.hist_rr[it - 1];
  const double beta = rr_cur / rr_prev, alpha = rr_prev / P.hist_pw[it];
  const bool two = (UM == TL_U_LAZY) && tl_cg_lazy_pending(it - st->cfg.first_it) == 2;
  const double alpha2 = two ? P.hist_rr[it - 2] / P.hist_pw[it - 1] : 0.0;
  if (blockIdx.x == 0 && threadIdx.x == 0) P.hist_rr[it] = rr_cur;
  // pointwise, so done IN PLACE in the buffer kernel A of iteration `it` wrote: the current p
  // stays buffer (it & 1) and a following phase (Chebyshev / PPCG) starts from it.
  double *pin = (it & 1) ? P.p1 : P.p0;
  const double *pold = (it & 1) ? P.p0 : P.p1;
  double *pout = pin;
  const Geo g = P.g;
  const bool physL = g.phys & TL_PHYS_LEFT, physR = g.phys & TL_PHYS_RIGHT;
  const bool physB = g.phys & TL_PHYS_BOTTOM, physT = g.phys & TL_PHYS_TOP;
  MarchCtx m;
  if (!tl_march_setup(g, P.t, m)) return;
  for (int j = m.j0; j < m.j1; j++) {
#pragma unroll
    for (int c = 0; c < 2; c++) {
      const int i = m.i0 + c;
      if (i >= g.nx) continue;
      const long o = (long)j * g.pitch + i;
      const double pv = pin[o];
      const double pn = beta * pv + P.r[o];
      pout[o] = pn;
      double un = 0.0;
      if (UM != TL_U_NONE) {
        un = P.u[o];
        if (two) un = un + alpha2 * pold[o];
        un = un + alpha * pv;
        P.u[o] = un;
      }
      if (physL && i == 0) { pout[o - 1] = pn; if (UM != TL_U_NONE) P.u[o - 1] = un; }
      if (physR && i == g.nx - 1) { pout[o + 1] = pn; if (UM != TL_U_NONE) P.u[o + 1] = un; }
      if (physB && j == 0) { pout[o - g.pitch] = pn; if (UM != TL_U_NONE) P.u[o - g.pitch] = un; }
      if (physT && j == g.ny - 1) { pout[o + g.pitch] = pn; if (UM != TL_U_NONE) P.u[o + g.pitch] = un; }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Chebyshev iteration, one kernel (k_cheby_fused_ring):  w = A u ; r = u0 - w ;
// p = alpha p + beta r ; u += p  (Cheby.mainstep! Cheby.jl:84-106; FIRST: p = r/theta and
// bb = sum(u0.u0), Cheby.init! Cheby.jl:64-81).  u is ping-ponged (ua/ub) because neighbouring
// warps read the old u.  w and r are stored only on iterations whose state can be observed
// (norm iterations and the last permitted iteration); the norm sum(r.r) is formed from registers.
// HBM traffic per cell: read u, u0, p, kx, ky; write p, u = 56 B.
// ------------------------------------------------------------------------------------------
struct ChebyParams {
  Geo g; Tiling t;
  SolveState *st;
  const double *alphas; const double *betas;   // chebyα / chebyβ (0-based storage of the 1-based arrays)
  const double *u0; double *ua; double *ub; double *p; double *w; double *r;
  const double *kx; const double *ky;
  double *partials;
  int single;   // 1: one tile, the kernel publishes its norm itself
  const CommDev *cd;
  Push push_ua, push_ub;
  double *p1;   // ping-pong partner of p: the two-iteration kernel (k_cheby_pair_ring) writes p out of place
};
// Which buffers hold the current u / p: every kernel flips u, the two-iteration kernel also flips p
// and advances the step counter by two -- hence the parity of (steps + pair kernels) for u and of
// the pair kernels alone for p (SolveState::cheby_pairs; always 0 without the pair kernel).
__device__ __forceinline__ int tl_cheby_u_parity(int step, int pairs) { return (step + pairs) & 1; }

// Was Chebyshev step `chebyiters` (1-based) a norm iteration?  Cheby.jl:40-51
__host__ __device__ inline bool tl_cheby_is_norm_iter(int chebyiters, int tt0, int est) {
  if (chebyiters < 1) return false;
  if (chebyiters == 1) return true;
  const int tt = tt0 + chebyiters - 1;
  return (chebyiters >= est) && ((tt + 1) % 10 == 0);
}
// Stop rule of the Chebyshev loop after `step` kernels (init included), evaluated at the entry
// of every kernel and by the host: converged on the last norm (Cheby.jl:57, summed over the
// tiles) or out of iterations.
// `norm`: the latest all-tiles norm (s.red_norm, or the total a kernel collected at its entry: split exchange)
__host__ __device__ inline bool tl_cheby_should_stop_n(const SolveState &s, double norm) {
  const int done_iters = s.cheby_step - 1;   // completed main steps
  if (done_iters >= 1 && tl_cheby_is_norm_iter(done_iters, s.cheby_tt0, s.cheby_est) && fabs(norm) < s.eps_cheby)
    return true;
  return s.cheby_tt0 + s.cheby_step - 1 > s.cheby_max_tt;
}
__host__ __device__ inline bool tl_cheby_should_stop(const SolveState &s) { return tl_cheby_should_stop_n(s, s.red_norm); }

// ------------------------------------------------------------------------------------------
// PPCG.  Outer iteration = k_cg_fused_w_ring<false> (w = A p, pw)  ->  k_ppcg_ur_sd  ->
// inner_steps x k_ppcg_inner_ring  (the last one forms rrn and advances the iteration counter).
// ------------------------------------------------------------------------------------------
// alpha = rro/pw ; u += alpha p ; r -= alpha w ; sd = r/theta   (CG.ur! + PPCG.init!(chunk,hd),
// PPCG.jl:65-70, :97-100).  Pointwise, 56 B/cell.
struct PpcgUrParams {
  Geo g; Tiling t;
  SolveState *st;
  double *hist_pw;
  const double *p0; const double *p1; const double *w; double *u; double *r; double *sd0;
  double *partials;
  const CommDev *cd;
  Push push_sd0;
  // depth-k groups (tl_kernels_ring.cuh, k_ppcg_inner_dk): r goes to the buffer the first group
  // reads (r_out, may be r itself) and the first group's halos are pushed at depth d_sd / d_r
  int deep;
  double *r_out;
  int d_sd, d_r;
  Push8 push_sd8, push_r8;
};

__global__ void __launch_bounds__(TL_FUSED_THREADS, 4) k_ppcg_ur_sd(const PpcgUrParams P) {
  tl_pdl_entry();
  __shared__ double sm[32];
  SolveState *st = P.st;
  tl_prof_entry(st);
  const TlScal sc = tl_entry_scalars(P.cd, st);
  const int it = st->iter;
  const double rr_cur = sc.rr;
  if (st->comm_error || tl_should_stop(it, rr_cur, st->cfg)) return;
  const double pw = sc.pw, alpha = rr_cur / pw, theta = st->theta;
  if (blockIdx.x == 0 && threadIdx.x == 0) { P.hist_pw[it + 1] = pw; st->inner_pp = 0; }
  // kernel A of this iteration wrote p into the (it&1 ? p0 : p1) buffer
  const double *__restrict__ p = (it & 1) ? P.p0 : P.p1;
  const Geo g = P.g;
  const bool tiled = P.cd != nullptr;
  MarchCtx m;
  if (tl_march_setup(g, P.t, m)) {
    for (int j = m.j0; j < m.j1; j++) {
      double2 sv = make_double2(0.0, 0.0), rv2 = make_double2(0.0, 0.0);
#pragma unroll
      for (int c = 0; c < 2; c++) {
        const int i = m.i0 + c;
        if (i >= g.nx) continue;
        const long o = (long)j * g.pitch + i;
        P.u[o] += alpha * p[o];
        const double rv = P.r[o] - alpha * P.w[o];
        if (P.deep) P.r_out[o] = rv; else P.r[o] = rv;
        const double s = rv / theta;
        P.sd0[o] = s;
        if (c == 0) { sv.x = s; rv2.x = rv; } else { sv.y = s; rv2.y = rv; }
      }
      if (P.deep) {
        tl_push_deep(P.push_sd8, g, P.d_sd, m.i0, m.acta, m.actb, j, sv);
        if (P.d_r > 0) tl_push_deep(P.push_r8, g, P.d_r, m.i0, m.acta, m.actb, j, rv2);
      } else if (tiled) tl_push_edges(P.push_sd0, g, m, j, sv);
    }
  }
  if (tiled) {   // the first inner step reads the neighbours' sd: completion barrier
    double acc[1] = {0.0};
    tl_kernel_tail(acc, false, st, P.partials, P.cd, sm);
  }
}

// one inner step (k_ppcg_inner_ring):  r -= A sd ; u += sd ; sd' = alpha_pp sd + beta_pp r
// (PPCG.jl:75-84, two-phase semantics through the sd ping-pong, Appendix A #17).  64 B/cell.
struct PpcgInnerParams {
  Geo g; Tiling t;
  SolveState *st;
  const double *alphas; const double *betas;
  double *sda; double *sdb; double *r; double *u; const double *kx; const double *ky;
  double *partials;
  int single;
  const CommDev *cd;
  Push push_sda, push_sdb, push_r;
};

// TWO inner steps in one pass (k_ppcg_pair_ring, tl_kernels_ring.cuh; an odd inner_steps ends with one
// k_ppcg_inner_ring step):
//   step A (pp):    rA = r - A sd ;   uA = u + sd ;   sA = alpha_pp sd + beta_pp rA
//   step B (pp+1):  rB = rA - A sA ;  uB = uA + sA ;  sB = alpha_pp+1 sA + beta_pp+1 rB
// ~32 B per cell-step instead of 64.  sA is recomputed redundantly at warp-task borders (needs the
// OLD r there, so r is ping-ponged like sd); the host gives every pair of an outer iteration its
// own in/out buffers, phased so that the last pair leaves r in the Chunk's own buffer.
struct PpcgPairParams {
  Geo g; Tiling t;
  SolveState *st;
  const double *alphas; const double *betas;
  const double *sin; double *sout; const double *rin; double *rout;
  double *u; const double *kx; const double *ky;
  double *partials;
  const CommDev *cd;           // tiles (TILED instantiation): exchange in the tail
  Push8 push_s, push_r;        // halo targets of sout / rout on the eight surrounding tiles
};

// Matrix-powers variant of the inner steps for tiles (k_ppcg_inner_dk): the steps are grouped by
// k = the halo depth of the exchange.  A group starts with sd valid k cells deep in the
// tile-internal halos (r: k-1 cells; kx, ky pulled once per solve), step q of the group computes
// r, sd' on the tile EXTENDED by k-1-q cells towards its neighbour tiles (redundantly: the same
// arithmetic on the same inputs as the owner, so bit-identical), and only the group's last step
// pushes halos and meets the other tiles -- one exchange per k steps instead of one per step.
// Buffers rotate so that a tile that runs ahead never writes what a neighbour may still read:
//   sd: group g reads sin[g&1] in its first step and writes sin[(g+1)&1] (+ the neighbours'
//       halos of it) in its last; the steps in between ping-pong between wa/wb;
//   r : likewise rin[(G-g)&1] -> rw (in place in between) -> rin[(G-g-1)&1], G = number of
//       groups, so that the last group leaves r in rin[0] = the Chunk's r.
// The halos of wa/wb/rw are never written by another tile; those of sin/rin only at group ends,
// and only the copy the receiving tile is not using before the next rendezvous.
struct DkExt { int left, right, bottom, top, emax; };   // extra strips / chunks of the extension (0/1 each), k-1
struct PpcgDkParams {
  Geo g; Tiling t; DkExt ext;    // t: the plain kernels' decomposition of the interior; ext: tl_march_setup_ext
  SolveState *st;
  const double *alphas; const double *betas;
  double *sin0; double *sin1; double *wa; double *wb;
  double *rin0; double *rin1; double *rw;
  double *u; const double *kx; const double *ky;
  double *partials;
  const CommDev *cd;
  int k;                         // exchange depth (steps per group)
  Push8 push_sin0, push_sin1, push_rin0, push_rin1;
};

// Work decomposition of the extended tile.  The interior keeps EXACTLY the warps of the plain
// kernels (strip s, chunk q -> warp q*nstrips + s: same cells per warp, same warps per block, so
// the dot product is summed in the same order and depth-k results equal depth-1 results bit for
// bit); the extension is covered by extra warps appended after them: one chunk of emax rows
// below / above the tile over all strips, and one 64-column strip left / right of the regular
// ones over all regular chunks (the right one only if the last regular strip cannot hold the
// extension).  The active window [xlo,xhi) x [ylo,yhi) of the current step masks the rest.
__device__ __forceinline__ bool tl_march_setup_ext(const Geo &g, const Tiling &t, const DkExt &x, int xlo, int xhi, int ylo,
                                                   int yhi, MarchCtx &m) {
  m.lane = threadIdx.x & 31;
  int wt = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int nreg = t.nstrips * t.nchunks;
  int strip, r0, r1;
  if (wt < nreg) {
    strip = wt % t.nstrips;
    const int q = wt / t.nstrips;
    r0 = q * t.rows_per_chunk; r1 = min(g.ny, r0 + t.rows_per_chunk);
  } else {
    wt -= nreg;
    const int nall = t.nstrips + x.left + x.right;
    bool found = false;
    strip = 0; r0 = r1 = 0;
    if (x.bottom) { if (wt < nall) { strip = wt; r0 = -x.emax; r1 = 0; found = true; } else wt -= nall; }
    if (!found && x.top) { if (wt < nall) { strip = wt; r0 = g.ny; r1 = g.ny + x.emax; found = true; } else wt -= nall; }
    if (!found && x.left) { if (wt < t.nchunks) { strip = t.nstrips; r0 = wt * t.rows_per_chunk; r1 = min(g.ny, r0 + t.rows_per_chunk); found = true; } else wt -= t.nchunks; }
    if (!found && x.right) { if (wt < t.nchunks) { strip = t.nstrips + x.left; r0 = wt * t.rows_per_chunk; r1 = min(g.ny, r0 + t.rows_per_chunk); found = true; } }
    if (!found) return false;
  }
  // strip index -> first column: regular strips, then the left extra strip, then the right one
  const int sx = strip < t.nstrips ? strip * TL_STRIP : ((x.left && strip == t.nstrips) ? -TL_STRIP : t.nstrips * TL_STRIP);
  m.j0 = max(ylo, r0);
  m.j1 = min(yhi, r1);
  m.i0 = sx + 2 * m.lane;
  m.acta = m.i0 >= xlo && m.i0 < xhi;
  m.actb = m.i0 + 1 >= xlo && m.i0 + 1 < xhi;
  m.ld_ok = m.i0 >= -TL_XPAD && m.i0 + 1 < g.pitch - TL_XPAD;   // the pair lies inside the padded row
  m.ecol = (m.lane == 0) ? sx - 1 : sx + TL_STRIP;
  m.has_edge = (m.lane == 0) ? (m.ecol >= -TL_XPAD) : (m.lane == 31 && m.ecol < g.pitch - TL_XPAD);
  return m.j0 < m.j1 && sx < xhi && sx + TL_STRIP > xlo;
}

// ------------------------------------------------------------------------------------------
// Jacobi iteration, one kernel (k_jacobi_fused_ring):  r .= u ; u = (u0 + sum k*r_nbr) / diag ;
// error = sum |u - r|  (Jacobi.iterate!, Jacobi.jl:62-82).  u is ping-ponged, so the buffer read
// IS the reference's r.  HBM traffic per cell: read u, u0, kx, ky; write u' = 40 B
// (as written: 72 B).  Every 50th iteration k_jacobi_resid replaces the error by sum(r.r) of the
// true residual (Jacobi.jl:16-21).
// ------------------------------------------------------------------------------------------
struct JacobiParams {
  Geo g; Tiling t;
  SolveState *st;
  const double *u0; double *ua; double *ub; double *r;
  const double *kx; const double *ky;
  double *partials;
  int single;
  int force_resid;   // k_jacobi_resid: run whatever the iteration number is
  const CommDev *cd;
  Push push_ua, push_ub;
};

