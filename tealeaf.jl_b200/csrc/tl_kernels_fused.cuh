// tl_kernels_fused.cuh -- the hot path: fused whole-iteration kernels for CG, Chebyshev and
// PPCG on the 5-point stencil (FP64, HBM-bound).
//
// Work decomposition ("warp-strip marching"): the tile is cut into 64-cell-wide column strips
// and row chunks; ONE WARP owns one (strip, chunk) and marches up its rows.  Each lane owns two
// adjacent cells (one 16-byte vector access per field per row, 512 B per warp per field: fully
// coalesced).  The stencil operand of rows j-1, j, j+1 is carried in registers, so every field
// is read from HBM once; x-neighbours come from warp shuffles, and only lanes 0/31 fetch one
// extra scalar per row from the neighbouring strip.  Loads for row j+1 are issued before row j
// is computed (register double-buffering) to keep enough bytes in flight without needing
// shared memory or barriers; warps are fully independent until the final reduction.
//
// Boundaries: on physical sides the reflective halo (kernels.jl:191-210, depth 1) is applied
// by index clamping instead of a separate halo pass, and the reflected value is written
// through to the halo cell so that memory holds what the reference's haloupdate! leaves.
// On tile-internal sides the halo cells are read from memory (filled by k_pull_halo).
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

__device__ __forceinline__ bool tl_march_setup(const Geo &g, const Tiling &t, MarchCtx &m) {
  m.lane = threadIdx.x & 31;
  const int wt = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wt >= t.nstrips * t.nchunks) return false;
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

// ------------------------------------------------------------------------------------------
// CG kernel A:  [p = r + beta p ; u += alpha_prev p_old]  (deferred from the previous
// iteration: CG.p! CG.jl:101-104 and the u half of CG.ur! CG.jl:95)  then  w = A p ,
// pw = sum(p.w)  (CG.w! CG.jl:82-90).  p is ping-ponged between p0/p1 because neighbouring
// warps still read the old p.  UPDATE_U = false is the PPCG outer variant (u is advanced by
// k_ppcg_ur_sd instead).
// HBM traffic per cell: read r, p, u, kx, ky; write p, u, w  = 64 B (48 B without u).
// ------------------------------------------------------------------------------------------
struct CgAParams {
  Geo g; Tiling t;
  SolveState *st;
  double *hist_rr; const double *hist_pw;
  const double *r; double *p0; double *p1; double *u; const double *kx; const double *ky; double *w;
  double *partials;
  int single;   // 1: one tile -- the kernel publishes its sums itself (no allreduce follows)
};

template <bool UPDATE_U>
__global__ void __launch_bounds__(TL_FUSED_THREADS, 2) k_cg_fused_w(const CgAParams P) {
  __shared__ double sm[32];
  SolveState *st = P.st;
  const int it = st->iter;
  const double rr_cur = st->red_rr;
  if (tl_should_stop(it, rr_cur, st->cfg)) return;
  const bool first = (it == st->cfg.first_it);
  double beta = 0.0, alpha_prev = 0.0;
  if (!first) {
    const double rr_prev = P.hist_rr[it - 1];
    beta = rr_cur / rr_prev;
    alpha_prev = rr_prev / P.hist_pw[it];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) P.hist_rr[it] = rr_cur;
  const double *__restrict__ pin = (it & 1) ? P.p1 : P.p0;
  double *__restrict__ pout = (it & 1) ? P.p0 : P.p1;
  const double *__restrict__ r = P.r;
  const double *__restrict__ kx = P.kx;
  const double *__restrict__ ky = P.ky;
  double *__restrict__ u = P.u;
  double *__restrict__ w = P.w;
  const Geo g = P.g;
  const int pitch = g.pitch;
  const bool physL = g.phys & TL_PHYS_LEFT, physR = g.phys & TL_PHYS_RIGHT;
  const bool physB = g.phys & TL_PHYS_BOTTOM, physT = g.phys & TL_PHYS_TOP;

  double acc[1] = {0.0};
  MarchCtx m;
  if (tl_march_setup(g, P.t, m)) {
    const double2 z2 = make_double2(0.0, 0.0);
    // operand X = p_new = beta*p + r (first iteration: p itself)
    auto comb = [&](double rv, double pv) { return first ? pv : beta * pv + rv; };
    auto comb2 = [&](double2 rv, double2 pv) { return make_double2(comb(rv.x, pv.x), comb(rv.y, pv.y)); };
    struct Pref { double2 r, p, ky, kx, u; double re, pe, kxe; };
    auto issue = [&](int j) {
      Pref f;
      const int jn = (j + 1 >= g.ny && physT) ? g.ny - 1 : j + 1;
      const long on = (long)jn * pitch + m.i0, oc = (long)j * pitch + m.i0;
      f.r = m.ld_ok ? tl_ld2(r + on) : z2;
      f.p = m.ld_ok ? tl_ld2(pin + on) : z2;
      f.ky = m.ld_ok ? tl_ld2(ky + oc + pitch) : z2;
      f.kx = m.ld_ok ? tl_ld2(kx + oc) : z2;
      f.u = (UPDATE_U && m.acta) ? tl_ld2_rw(u + oc) : z2;
      const long oe = (long)jn * pitch + m.ecol;
      f.re = m.has_edge ? __ldg(r + oe) : 0.0;
      f.pe = m.has_edge ? __ldg(pin + oe) : 0.0;
      f.kxe = (m.lane == 31 && m.has_edge) ? __ldg(kx + (long)j * pitch + m.ecol) : 0.0;
      return f;
    };
    // prologue: rows j0-1 (clamped on a physical bottom) and j0
    double2 Xm, Xc, pc, kyc;
    double XcE;
    {
      const int jm = (m.j0 == 0 && physB) ? 0 : m.j0 - 1;
      const long om = (long)jm * pitch + m.i0, oc = (long)m.j0 * pitch + m.i0;
      const double2 rm = m.ld_ok ? tl_ld2(r + om) : z2, pm = m.ld_ok ? tl_ld2(pin + om) : z2;
      const double2 rc = m.ld_ok ? tl_ld2(r + oc) : z2;
      pc = m.ld_ok ? tl_ld2(pin + oc) : z2;
      kyc = m.ld_ok ? tl_ld2(ky + oc) : z2;
      const long oe = (long)m.j0 * pitch + m.ecol;
      const double re = m.has_edge ? __ldg(r + oe) : 0.0, pe = m.has_edge ? __ldg(pin + oe) : 0.0;
      Xm = comb2(rm, pm);
      Xc = comb2(rc, pc);
      XcE = comb(re, pe);
    }
    Pref pf = issue(m.j0);
    for (int j = m.j0; j < m.j1; j++) {
      const Pref cur = pf;
      if (j + 1 < m.j1) pf = issue(j + 1);
      const double2 Xn = comb2(cur.r, cur.p);
      const double XnE = comb(cur.re, cur.pe);
      double xl = __shfl_up_sync(0xffffffffu, Xc.y, 1);
      double xr = __shfl_down_sync(0xffffffffu, Xc.x, 1);
      double kxr = __shfl_down_sync(0xffffffffu, cur.kx.x, 1);
      if (m.lane == 0) xl = XcE;
      if (m.lane == 31) { xr = XcE; kxr = cur.kxe; }
      const double La = (physL && m.i0 == 0) ? Xc.x : xl;
      const double Ra = (physR && m.i0 == g.nx - 1) ? Xc.x : Xc.y;
      const double Lb = Xc.x;
      const double Rb = (physR && m.i0 + 1 == g.nx - 1) ? Xc.y : xr;
      const double wa = ((((1.0 + cur.kx.y) + cur.kx.x) + cur.ky.x) + kyc.x) * Xc.x -
                        (cur.kx.y * Ra + cur.kx.x * La) - (cur.ky.x * Xn.x + kyc.x * Xm.x);
      const double wb = ((((1.0 + kxr) + cur.kx.y) + cur.ky.y) + kyc.y) * Xc.y -
                        (kxr * Rb + cur.kx.y * Lb) - (cur.ky.y * Xn.y + kyc.y * Xm.y);
      const long oc = (long)j * pitch + m.i0;
      double2 un = z2;
      if (UPDATE_U) un = make_double2(cur.u.x + alpha_prev * pc.x, cur.u.y + alpha_prev * pc.y);
      if (m.actb) {
        tl_st2(w + oc, make_double2(wa, wb));
        tl_st2(pout + oc, Xc);
        if (UPDATE_U) tl_st2(u + oc, un);
        acc[0] += wa * Xc.x;
        acc[0] += wb * Xc.y;
      } else if (m.acta) {
        w[oc] = wa; pout[oc] = Xc.x;
        if (UPDATE_U) u[oc] = un.x;
        acc[0] += wa * Xc.x;
      }
      // write-through of the depth-1 reflective halo of p (and u): haloupdate!(.., [:u,:p]) CG.jl:22
      if (m.acta) {
        if (physL && m.i0 == 0) { pout[oc - 1] = Xc.x; if (UPDATE_U) u[oc - 1] = un.x; }
        if (physR && m.i0 == g.nx - 1) { pout[oc + 1] = Xc.x; if (UPDATE_U) u[oc + 1] = un.x; }
        if (physR && m.i0 + 1 == g.nx - 1) { pout[oc + 2] = Xc.y; if (UPDATE_U) u[oc + 2] = un.y; }
        if (physB && j == 0) {
          pout[oc - pitch] = Xc.x; if (UPDATE_U) u[oc - pitch] = un.x;
          if (m.actb) { pout[oc - pitch + 1] = Xc.y; if (UPDATE_U) u[oc - pitch + 1] = un.y; }
        }
        if (physT && j == g.ny - 1) {
          pout[oc + pitch] = Xc.x; if (UPDATE_U) u[oc + pitch] = un.x;
          if (m.actb) { pout[oc + pitch + 1] = Xc.y; if (UPDATE_U) u[oc + pitch + 1] = un.y; }
        }
      }
      Xm = Xc; Xc = Xn; XcE = XnE; pc = cur.p; kyc = cur.ky;
    }
  }
  if (tl_grid_sum<1>(acc, P.partials, &st->counter, sm) && threadIdx.x == 0) {
    st->red_pw_local = acc[0];
    if (P.single) st->red_pw = acc[0];
  }
}

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
};

__global__ void __launch_bounds__(TL_FUSED_THREADS, 4) k_cg_fused_r(const CgBParams P) {
  __shared__ double sm[32];
  SolveState *st = P.st;
  const int it = st->iter;
  const double rr_cur = st->red_rr;
  if (tl_should_stop(it, rr_cur, st->cfg)) return;
  const double pw = st->red_pw;
  const double alpha = rr_cur / pw;
  if (blockIdx.x == 0 && threadIdx.x == 0) P.hist_pw[it + 1] = pw;
  const Geo g = P.g;
  double *__restrict__ r = P.r;
  const double *__restrict__ w = P.w;
  double acc[1] = {0.0};
  MarchCtx m;
  if (tl_march_setup(g, P.t, m)) {
    int j = m.j0;
    if (m.actb) {
      for (; j + 4 <= m.j1; j += 4) {
        double2 rv[4], wv[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const long o = (long)(j + q) * g.pitch + m.i0;
          rv[q] = tl_ld2_rw(r + o);
          wv[q] = tl_ld2(w + o);
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const long o = (long)(j + q) * g.pitch + m.i0;
          rv[q].x = rv[q].x - alpha * wv[q].x;
          rv[q].y = rv[q].y - alpha * wv[q].y;
          tl_st2(r + o, rv[q]);
          acc[0] += rv[q].x * rv[q].x;
          acc[0] += rv[q].y * rv[q].y;
        }
      }
    }
    for (; j < m.j1; j++) {
      const long o = (long)j * g.pitch + m.i0;
      if (m.acta) { const double v = r[o] - alpha * w[o]; r[o] = v; acc[0] += v * v; }
      if (m.actb) { const double v = r[o + 1] - alpha * w[o + 1]; r[o + 1] = v; acc[0] += v * v; }
    }
  }
  if (tl_grid_sum<1>(acc, P.partials, &st->counter, sm) && threadIdx.x == 0) {
    st->red_rr_local = acc[0];
    if (P.single) st->red_rr = acc[0];
    st->iter = it + 1;
  }
}

// ------------------------------------------------------------------------------------------
// CG flush: applies the updates kernel A deferred for the last executed iteration
// (u += alpha p ; p = beta p + r) including the depth-1 halo write-through, so that memory
// holds the reference's post-iteration state.  Pointwise.
// ------------------------------------------------------------------------------------------
template <bool UPDATE_U>
__global__ void __launch_bounds__(TL_FUSED_THREADS, 4) k_cg_flush(const CgAParams P) {
  SolveState *st = P.st;
  const int it = st->iter;
  if (it == st->cfg.first_it) return;
  const double rr_cur = st->red_rr, rr_prev = P.hist_rr[it - 1];
  const double beta = rr_cur / rr_prev, alpha = rr_prev / P.hist_pw[it];
  if (blockIdx.x == 0 && threadIdx.x == 0) P.hist_rr[it] = rr_cur;
  // pointwise, so done IN PLACE in the buffer kernel A of iteration `it` wrote: the current p
  // stays buffer (it & 1) and a following phase (Chebyshev / PPCG) starts from it.
  double *pin = (it & 1) ? P.p1 : P.p0;
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
      if (UPDATE_U) { un = P.u[o] + alpha * pv; P.u[o] = un; }
      if (physL && i == 0) { pout[o - 1] = pn; if (UPDATE_U) P.u[o - 1] = un; }
      if (physR && i == g.nx - 1) { pout[o + 1] = pn; if (UPDATE_U) P.u[o + 1] = un; }
      if (physB && j == 0) { pout[o - g.pitch] = pn; if (UPDATE_U) P.u[o - g.pitch] = un; }
      if (physT && j == g.ny - 1) { pout[o + g.pitch] = pn; if (UPDATE_U) P.u[o + g.pitch] = un; }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Chebyshev iteration, one kernel:  w = A u ; r = u0 - w ; p = alpha p + beta r ; u += p
// (Cheby.mainstep! Cheby.jl:84-106; FIRST: p = r/theta and bb = sum(u0.u0), Cheby.init!
// Cheby.jl:64-81).  u is ping-ponged (ua/ub) because neighbouring warps read the old u.
// w and r are stored only on iterations whose state can be observed (norm iterations and the
// last permitted iteration); the norm sum(r.r) is formed from registers.
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
};

// Was Chebyshev step `chebyiters` (1-based) a norm iteration?  Cheby.jl:40-51
__host__ __device__ inline bool tl_cheby_is_norm_iter(int chebyiters, int tt0, int est) {
  if (chebyiters < 1) return false;
  if (chebyiters == 1) return true;
  const int tt = tt0 + chebyiters - 1;
  return (chebyiters >= est) && ((tt + 1) % 10 == 0);
}
// Stop rule of the Chebyshev loop after `step` kernels (init included), evaluated at the entry
// of every kernel and by the host: converged on the last norm (Cheby.jl:57, after the
// allreduce when tiled) or out of iterations.
__host__ __device__ inline bool tl_cheby_should_stop(const SolveState &s) {
  const int done_iters = s.cheby_step - 1;   // completed main steps
  if (done_iters >= 1 && tl_cheby_is_norm_iter(done_iters, s.cheby_tt0, s.cheby_est) && fabs(s.red_norm) < s.eps_cheby)
    return true;
  return s.cheby_tt0 + s.cheby_step - 1 > s.cheby_max_tt;
}

template <bool FIRST>
__global__ void __launch_bounds__(TL_FUSED_THREADS, 2) k_cheby_fused(const ChebyParams P) {
  __shared__ double sm[32];
  SolveState *st = P.st;
  const int step = st->cheby_step;  // completed Chebyshev kernels; FIRST runs at step 0
  double alpha = 0.0, beta = 0.0;
  bool calc_norm, store_wr;
  const double theta = st->theta;
  if (FIRST) {
    calc_norm = true;   // bb
    store_wr = true;
  } else {
    // `step` kernels done => this is chebyiters = step (init was step 0 -> 1), Cheby.jl:35-51
    if (tl_cheby_should_stop(*st)) return;
    const int chebyiters = step;
    const int tt = st->cheby_tt0 + chebyiters - 1;
    alpha = P.alphas[chebyiters];   // 1-based chebyα[chebyiters+1]
    beta = P.betas[chebyiters];
    calc_norm = tl_cheby_is_norm_iter(chebyiters, st->cheby_tt0, st->cheby_est);
    store_wr = calc_norm || (tt == st->cheby_max_tt);
  }
  const double *__restrict__ uin = (step & 1) ? P.ub : P.ua;
  double *__restrict__ uout = (step & 1) ? P.ua : P.ub;
  const double *__restrict__ u0 = P.u0;
  const double *__restrict__ kx = P.kx;
  const double *__restrict__ ky = P.ky;
  double *__restrict__ p = P.p;
  const Geo g = P.g;
  const int pitch = g.pitch;
  const bool physL = g.phys & TL_PHYS_LEFT, physR = g.phys & TL_PHYS_RIGHT;
  const bool physB = g.phys & TL_PHYS_BOTTOM, physT = g.phys & TL_PHYS_TOP;

  double acc[1] = {0.0};
  MarchCtx m;
  if (tl_march_setup(g, P.t, m)) {
    const double2 z2 = make_double2(0.0, 0.0);
    struct Pref { double2 x, ky, kx, u0, p; double xe, kxe; };
    auto issue = [&](int j) {
      Pref f;
      const int jn = (j + 1 >= g.ny && physT) ? g.ny - 1 : j + 1;
      const long on = (long)jn * pitch + m.i0, oc = (long)j * pitch + m.i0;
      f.x = m.ld_ok ? tl_ld2(uin + on) : z2;
      f.ky = m.ld_ok ? tl_ld2(ky + oc + pitch) : z2;
      f.kx = m.ld_ok ? tl_ld2(kx + oc) : z2;
      f.u0 = m.acta ? tl_ld2(u0 + oc) : z2;
      f.p = (!FIRST && m.acta) ? tl_ld2_rw(p + oc) : z2;
      f.xe = m.has_edge ? __ldg(uin + (long)jn * pitch + m.ecol) : 0.0;
      f.kxe = (m.lane == 31 && m.has_edge) ? __ldg(kx + (long)j * pitch + m.ecol) : 0.0;
      return f;
    };
    double2 Xm, Xc, kyc;
    double XcE;
    {
      const int jm = (m.j0 == 0 && physB) ? 0 : m.j0 - 1;
      const long om = (long)jm * pitch + m.i0, oc = (long)m.j0 * pitch + m.i0;
      Xm = m.ld_ok ? tl_ld2(uin + om) : z2;
      Xc = m.ld_ok ? tl_ld2(uin + oc) : z2;
      kyc = m.ld_ok ? tl_ld2(ky + oc) : z2;
      XcE = m.has_edge ? __ldg(uin + (long)m.j0 * pitch + m.ecol) : 0.0;
    }
    Pref pf = issue(m.j0);
    for (int j = m.j0; j < m.j1; j++) {
      const Pref cur = pf;
      if (j + 1 < m.j1) pf = issue(j + 1);
      const double2 Xn = cur.x;
      double xl = __shfl_up_sync(0xffffffffu, Xc.y, 1);
      double xr = __shfl_down_sync(0xffffffffu, Xc.x, 1);
      double kxr = __shfl_down_sync(0xffffffffu, cur.kx.x, 1);
      if (m.lane == 0) xl = XcE;
      if (m.lane == 31) { xr = XcE; kxr = cur.kxe; }
      const double La = (physL && m.i0 == 0) ? Xc.x : xl;
      const double Ra = (physR && m.i0 == g.nx - 1) ? Xc.x : Xc.y;
      const double Lb = Xc.x;
      const double Rb = (physR && m.i0 + 1 == g.nx - 1) ? Xc.y : xr;
      const double wa = ((((1.0 + cur.kx.y) + cur.kx.x) + cur.ky.x) + kyc.x) * Xc.x -
                        (cur.kx.y * Ra + cur.kx.x * La) - (cur.ky.x * Xn.x + kyc.x * Xm.x);
      const double wb = ((((1.0 + kxr) + cur.kx.y) + cur.ky.y) + kyc.y) * Xc.y -
                        (kxr * Rb + cur.kx.y * Lb) - (cur.ky.y * Xn.y + kyc.y * Xm.y);
      const double ra = cur.u0.x - wa, rb = cur.u0.y - wb;
      double2 pn;
      if (FIRST) { pn.x = ra / theta; pn.y = rb / theta; }
      else { pn.x = alpha * cur.p.x + beta * ra; pn.y = alpha * cur.p.y + beta * rb; }
      const double2 un = make_double2(Xc.x + pn.x, Xc.y + pn.y);
      const long oc = (long)j * pitch + m.i0;
      if (m.actb) {
        tl_st2(p + oc, pn);
        tl_st2(uout + oc, un);
        if (store_wr) { tl_st2(P.w + oc, make_double2(wa, wb)); tl_st2(P.r + oc, make_double2(ra, rb)); }
        if (FIRST) { acc[0] += cur.u0.x * cur.u0.x; acc[0] += cur.u0.y * cur.u0.y; }
        else { acc[0] += ra * ra; acc[0] += rb * rb; }
      } else if (m.acta) {
        p[oc] = pn.x; uout[oc] = un.x;
        if (store_wr) { P.w[oc] = wa; P.r[oc] = ra; }
        acc[0] += FIRST ? cur.u0.x * cur.u0.x : ra * ra;
      }
      // haloupdate!(.., [:u]) Cheby.jl:55/:78 as a write-through
      if (m.acta) {
        if (physL && m.i0 == 0) uout[oc - 1] = un.x;
        if (physR && m.i0 == g.nx - 1) uout[oc + 1] = un.x;
        if (physR && m.i0 + 1 == g.nx - 1) uout[oc + 2] = un.y;
        if (physB && j == 0) { uout[oc - pitch] = un.x; if (m.actb) uout[oc - pitch + 1] = un.y; }
        if (physT && j == g.ny - 1) { uout[oc + pitch] = un.x; if (m.actb) uout[oc + pitch + 1] = un.y; }
      }
      Xm = Xc; Xc = Xn; XcE = cur.xe; kyc = cur.ky;
    }
  }
  if (calc_norm) {
    if (tl_grid_sum<1>(acc, P.partials, &st->counter, sm) && threadIdx.x == 0) {
      st->red_norm_local = acc[0];
      if (P.single) st->red_norm = acc[0];   // tiled: k_stage_scalar + allreduce publish it
      st->cheby_step = step + 1;
    }
  } else {
    // no reduction needed: only the ticket, so that the last block can advance the step
    __shared__ bool s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      s_last = (atomicAdd(&st->counter, 1u) == gridDim.x - 1);
      if (s_last) { st->counter = 0u; st->cheby_step = step + 1; }
    }
  }
}

// ------------------------------------------------------------------------------------------
// PPCG.  Outer iteration = k_cg_fused_w<false> (w = A p, pw)  ->  k_ppcg_ur_sd  ->
// inner_steps x k_ppcg_inner  (the last one forms rrn and advances the iteration counter).
// ------------------------------------------------------------------------------------------
// alpha = rro/pw ; u += alpha p ; r -= alpha w ; sd = r/theta   (CG.ur! + PPCG.init!(chunk,hd),
// PPCG.jl:65-70, :97-100).  Pointwise, 56 B/cell.
struct PpcgUrParams {
  Geo g; Tiling t;
  SolveState *st;
  double *hist_pw;
  const double *p0; const double *p1; const double *w; double *u; double *r; double *sd0;
};

__global__ void __launch_bounds__(TL_FUSED_THREADS, 4) k_ppcg_ur_sd(const PpcgUrParams P) {
  SolveState *st = P.st;
  const int it = st->iter;
  const double rr_cur = st->red_rr;
  if (tl_should_stop(it, rr_cur, st->cfg)) return;
  const double pw = st->red_pw, alpha = rr_cur / pw, theta = st->theta;
  if (blockIdx.x == 0 && threadIdx.x == 0) { P.hist_pw[it + 1] = pw; st->inner_pp = 0; }
  // kernel A of this iteration wrote p into the (it&1 ? p0 : p1) buffer
  const double *__restrict__ p = (it & 1) ? P.p0 : P.p1;
  const Geo g = P.g;
  MarchCtx m;
  if (!tl_march_setup(g, P.t, m)) return;
  for (int j = m.j0; j < m.j1; j++) {
#pragma unroll
    for (int c = 0; c < 2; c++) {
      const int i = m.i0 + c;
      if (i >= g.nx) continue;
      const long o = (long)j * g.pitch + i;
      P.u[o] += alpha * p[o];
      const double rv = P.r[o] - alpha * P.w[o];
      P.r[o] = rv;
      P.sd0[o] = rv / theta;
    }
  }
}

// one inner step:  r -= A sd ; u += sd ; sd' = alpha_pp sd + beta_pp r   (PPCG.jl:75-84,
// two-phase semantics through the sd ping-pong, Appendix A #17).  64 B/cell.
struct PpcgInnerParams {
  Geo g; Tiling t;
  SolveState *st;
  const double *alphas; const double *betas;
  double *sda; double *sdb; double *r; double *u; const double *kx; const double *ky;
  double *partials;
  int single;
};

__global__ void __launch_bounds__(TL_FUSED_THREADS, 2) k_ppcg_inner(const PpcgInnerParams P) {
  __shared__ double sm[32];
  SolveState *st = P.st;
  const int it = st->iter;
  if (tl_should_stop(it, st->red_rr, st->cfg)) return;
  const int pp = st->inner_pp;           // completed inner steps of this outer iteration
  const bool last = (pp + 1 == st->inner_steps);
  const double alpha = P.alphas[pp], beta = P.betas[pp];   // chebyα[pp], 1-based pp
  const double *__restrict__ sin = (pp & 1) ? P.sdb : P.sda;
  double *__restrict__ sout = (pp & 1) ? P.sda : P.sdb;
  const double *__restrict__ kx = P.kx;
  const double *__restrict__ ky = P.ky;
  double *__restrict__ r = P.r;
  double *__restrict__ u = P.u;
  const Geo g = P.g;
  const int pitch = g.pitch;
  const bool physL = g.phys & TL_PHYS_LEFT, physR = g.phys & TL_PHYS_RIGHT;
  const bool physB = g.phys & TL_PHYS_BOTTOM, physT = g.phys & TL_PHYS_TOP;

  double acc[1] = {0.0};
  MarchCtx m;
  if (tl_march_setup(g, P.t, m)) {
    const double2 z2 = make_double2(0.0, 0.0);
    struct Pref { double2 x, ky, kx, r, u; double xe, kxe; };
    auto issue = [&](int j) {
      Pref f;
      const int jn = (j + 1 >= g.ny && physT) ? g.ny - 1 : j + 1;
      const long on = (long)jn * pitch + m.i0, oc = (long)j * pitch + m.i0;
      f.x = m.ld_ok ? tl_ld2(sin + on) : z2;
      f.ky = m.ld_ok ? tl_ld2(ky + oc + pitch) : z2;
      f.kx = m.ld_ok ? tl_ld2(kx + oc) : z2;
      f.r = m.acta ? tl_ld2_rw(r + oc) : z2;
      f.u = m.acta ? tl_ld2_rw(u + oc) : z2;
      f.xe = m.has_edge ? __ldg(sin + (long)jn * pitch + m.ecol) : 0.0;
      f.kxe = (m.lane == 31 && m.has_edge) ? __ldg(kx + (long)j * pitch + m.ecol) : 0.0;
      return f;
    };
    double2 Xm, Xc, kyc;
    double XcE;
    {
      const int jm = (m.j0 == 0 && physB) ? 0 : m.j0 - 1;
      const long om = (long)jm * pitch + m.i0, oc = (long)m.j0 * pitch + m.i0;
      Xm = m.ld_ok ? tl_ld2(sin + om) : z2;
      Xc = m.ld_ok ? tl_ld2(sin + oc) : z2;
      kyc = m.ld_ok ? tl_ld2(ky + oc) : z2;
      XcE = m.has_edge ? __ldg(sin + (long)m.j0 * pitch + m.ecol) : 0.0;
    }
    Pref pf = issue(m.j0);
    for (int j = m.j0; j < m.j1; j++) {
      const Pref cur = pf;
      if (j + 1 < m.j1) pf = issue(j + 1);
      const double2 Xn = cur.x;
      double xl = __shfl_up_sync(0xffffffffu, Xc.y, 1);
      double xr = __shfl_down_sync(0xffffffffu, Xc.x, 1);
      double kxr = __shfl_down_sync(0xffffffffu, cur.kx.x, 1);
      if (m.lane == 0) xl = XcE;
      if (m.lane == 31) { xr = XcE; kxr = cur.kxe; }
      const double La = (physL && m.i0 == 0) ? Xc.x : xl;
      const double Ra = (physR && m.i0 == g.nx - 1) ? Xc.x : Xc.y;
      const double Lb = Xc.x;
      const double Rb = (physR && m.i0 + 1 == g.nx - 1) ? Xc.y : xr;
      const double wa = ((((1.0 + cur.kx.y) + cur.kx.x) + cur.ky.x) + kyc.x) * Xc.x -
                        (cur.kx.y * Ra + cur.kx.x * La) - (cur.ky.x * Xn.x + kyc.x * Xm.x);
      const double wb = ((((1.0 + kxr) + cur.kx.y) + cur.ky.y) + kyc.y) * Xc.y -
                        (kxr * Rb + cur.kx.y * Lb) - (cur.ky.y * Xn.y + kyc.y * Xm.y);
      const double2 rn = make_double2(cur.r.x - wa, cur.r.y - wb);
      const double2 un = make_double2(cur.u.x + Xc.x, cur.u.y + Xc.y);
      const double2 sn = make_double2(alpha * Xc.x + beta * rn.x, alpha * Xc.y + beta * rn.y);
      const long oc = (long)j * pitch + m.i0;
      if (m.actb) {
        tl_st2(r + oc, rn); tl_st2(u + oc, un); tl_st2(sout + oc, sn);
        acc[0] += rn.x * rn.x;
        acc[0] += rn.y * rn.y;
      } else if (m.acta) {
        r[oc] = rn.x; u[oc] = un.x; sout[oc] = sn.x;
        acc[0] += rn.x * rn.x;
      }
      // halo(sd) of PPCG.jl:76 happens BEFORE each inner step, so after the last step memory
      // holds the reflection of the step's *input*; earlier steps leave the output's.
      if (m.acta) {
        const double ha = last ? Xc.x : sn.x, hb = last ? Xc.y : sn.y;
        if (physL && m.i0 == 0) sout[oc - 1] = ha;
        if (physR && m.i0 == g.nx - 1) sout[oc + 1] = ha;
        if (physR && m.i0 + 1 == g.nx - 1) sout[oc + 2] = hb;
        if (physB && j == 0) { sout[oc - pitch] = ha; if (m.actb) sout[oc - pitch + 1] = hb; }
        if (physT && j == g.ny - 1) { sout[oc + pitch] = ha; if (m.actb) sout[oc + pitch + 1] = hb; }
      }
      Xm = Xc; Xc = Xn; XcE = cur.xe; kyc = cur.ky;
    }
  }
  if (last) {
    if (tl_grid_sum<1>(acc, P.partials, &st->counter, sm) && threadIdx.x == 0) {
      st->red_rr_local = acc[0];      // PPCG.jl:88
      if (P.single) st->red_rr = acc[0];
      st->iter = it + 1;
      st->inner_pp = pp + 1;
    }
  } else {
    __shared__ bool s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      s_last = (atomicAdd(&st->counter, 1u) == gridDim.x - 1);
      if (s_last) { st->counter = 0u; st->inner_pp = pp + 1; }
    }
  }
}
