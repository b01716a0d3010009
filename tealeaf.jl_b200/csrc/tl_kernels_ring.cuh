// tl_kernels_ring.cuh -- the hot path, part 2: the stencil kernels (CG kernel A, the Chebyshev
// iteration, the PPCG inner step).
//
// Warp-strip marching (tl_kernels_fused.cuh): the stencil operand of rows j-1, j, j+1 is carried
// in registers, so every field is read from HBM once; x-neighbours come from warp shuffles, and
// only lanes 0/31 fetch one extra scalar per row from the neighbouring strip.  The loads for the
// next S-1 rows are in flight as asynchronous global->shared copies (cp.async / LDGSTS, 16 B per
// lane): every warp owns a private ring of S row-slots in shared memory, each lane
// reads back only the bytes it copied itself, so no barrier of any kind is needed -- completion
// is tracked per thread with cp.async commit/wait groups.  The memory system then sees
// (S-1) x 2.5 KB per warp in flight (vs 2.5 KB with register double-buffering), which is what an
// HBM3e stack at ~1 us loaded latency needs (Little: 6.5 TB/s x 1 us = 6.5 MB chip-wide).
#pragma once
#include "tl_kernels_basic.cuh"
#include "tl_kernels_fused.cuh"

#define TL_RING_FIELDS 5
#define TL_RING_STAGE_BYTES (TL_RING_FIELDS * 512 + 64)   // 5 x (32 lanes x 16 B) + edge scalars

__device__ __forceinline__ unsigned tl_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tl_cp16(unsigned dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void tl_cp16_hint(unsigned dst, const void *src, unsigned long long pol) {
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "l"(pol) : "memory");
}
__device__ __forceinline__ void tl_cp8(unsigned dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void tl_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tl_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ double2 tl_lds2(unsigned a) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ double tl_lds1(unsigned a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}

// CG kernel A (algorithm and citations: CgAParams in tl_kernels_fused.cuh).
// Scalars of the iteration a work item belongs to.
struct CgAIter {
  int it;
  bool first;
  double beta, alpha_prev;
  double alpha_prev2;   // TL_U_LAZY: alpha of iteration it-2
};

// Ring slot of kernel A: 4 / 5 / 6 fields of 512 B per row, by the way the launch advances u (TL_U_*, tl_kernels_fused.cuh)
#define TL_CGA_FIELDS(UM) ((UM) == TL_U_LAZY ? 6 : (UM) == TL_U_EVERY ? 5 : 4)
#define TL_CGA_STAGE_BYTES(UM) (TL_CGA_FIELDS(UM) * 512 + 64)
static_assert(TL_CGA_STAGE_BYTES(TL_U_EVERY) == TL_RING_STAGE_BYTES, "the persistent kernel sizes kernel A's ring with TL_RING_STAGE_BYTES");

// The rows of one work item (8 warp tasks) of kernel A.  COH = true (persistent kernel: r, p, u change
// inside the launch) keeps the prologue loads coherent; the ring loads are cp.async (L2) either way.
template <int UM, int S, bool COH>
__device__ __forceinline__ void tl_cg_a_item(const CgAParams &P, const CgAIter &I, int blk, unsigned char *ring_raw,
                                             double &acc0) {
  constexpr unsigned STAGE = TL_CGA_STAGE_BYTES(UM);
  constexpr unsigned NF = TL_CGA_FIELDS(UM);
  constexpr bool upd2 = UM == TL_U_LAZY;
  const bool upd = UM != TL_U_NONE && !(upd2 && I.first);   // the first iteration of a lazy phase has nothing pending
  const int it = I.it;
  const bool first = I.first;
  const double beta = I.beta, alpha_prev = I.alpha_prev;
  const double alpha_prev2 = I.alpha_prev2;
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

  const unsigned long long pol_keep = tl_policy(P.hint_keep), pol_stream = tl_policy(P.hint_stream);
  const bool tiled = P.cd != nullptr;
  const Push &push = (it & 1) ? P.push_p0 : P.push_p1;   // halo targets of pout
  MarchCtx m;
  if (!tl_march_setup_blk(g, P.t, m, blk, 0)) return;
  const double2 z2 = make_double2(0.0, 0.0);
  auto comb = [&](double rv, double pv) { return first ? pv : beta * pv + rv; };
  auto comb2 = [&](double2 rv, double2 pv) { return make_double2(comb(rv.x, pv.x), comb(rv.y, pv.y)); };
  // this warp's ring; slot layout: [field 0..4][lane] double2, then 8 edge doubles
  const unsigned ring = tl_smem_u32(ring_raw) + (threadIdx.x >> 5) * (S * STAGE);
  const unsigned lane_off = m.lane * 16;
  const unsigned edge_off = NF * 512 + (m.lane == 0 ? 0 : 24);   // lane 0: 3 doubles, lane 31: 3 doubles
  auto issue = [&](int j, int stage) {
    const unsigned base = ring + stage * STAGE;
    const int jn = (j + 1 >= g.ny && physT) ? g.ny - 1 : j + 1;
    const long on = (long)jn * pitch + m.i0, oc = (long)j * pitch + m.i0;
    if (m.ld_ok) {
      tl_cp16_hint(base + 0 * 512 + lane_off, r + on, pol_keep);
      tl_cp16_hint(base + 1 * 512 + lane_off, pin + on, pol_stream);
      tl_cp16_hint(base + 2 * 512 + lane_off, ky + oc + pitch, pol_stream);
      tl_cp16_hint(base + 3 * 512 + lane_off, kx + oc, pol_stream);
    }
    if (upd && m.acta) tl_cp16_hint(base + 4 * 512 + lane_off, u + oc, pol_stream);
    if (upd2 && upd && m.acta) tl_cp16_hint(base + 5 * 512 + lane_off, pout + oc, pol_stream);   // p(it-2), before it is overwritten
    if (m.has_edge) {
      const long oe = (long)jn * pitch + m.ecol;
      tl_cp8(base + edge_off + 0, r + oe);
      tl_cp8(base + edge_off + 8, pin + oe);
      if (m.lane == 31) tl_cp8(base + edge_off + 16, kx + (long)j * pitch + m.ecol);
    }
  };
  // prologue: rows j0-1 (clamped on a physical bottom) and j0, plain loads
  double2 Xm, Xc, pc, kyc;
  double XcE;
  {
    const int jm = (m.j0 == 0 && physB) ? 0 : m.j0 - 1;
    const long om = (long)jm * pitch + m.i0, oc = (long)m.j0 * pitch + m.i0;
    const double2 rm = m.ld_ok ? tl_ldg2<COH>(r + om) : z2, pm = m.ld_ok ? tl_ldg2<COH>(pin + om) : z2;
    const double2 rc = m.ld_ok ? tl_ldg2<COH>(r + oc) : z2;
    pc = m.ld_ok ? tl_ldg2<COH>(pin + oc) : z2;
    kyc = m.ld_ok ? tl_ld2(ky + oc) : z2;
    const long oe = (long)m.j0 * pitch + m.ecol;
    const double re = m.has_edge ? tl_ldg1<COH>(r + oe) : 0.0, pe = m.has_edge ? tl_ldg1<COH>(pin + oe) : 0.0;
    Xm = comb2(rm, pm);
    Xc = comb2(rc, pc);
    XcE = comb(re, pe);
  }
#pragma unroll
  for (int d = 0; d < S - 1; d++) {
    if (m.j0 + d < m.j1) issue(m.j0 + d, d);
    tl_cp_commit();
  }
  int stage = 0;            // slot holding row j
  int fill = S - 1;         // slot that receives row j + S - 1
  for (int j = m.j0; j < m.j1; j++) {
    if (j + S - 1 < m.j1) issue(j + S - 1, fill);
    tl_cp_commit();
    tl_cp_wait<S - 1>();
    const unsigned base = ring + stage * STAGE;
    const double2 c_r = m.ld_ok ? tl_lds2(base + 0 * 512 + lane_off) : z2;
    const double2 c_p = m.ld_ok ? tl_lds2(base + 1 * 512 + lane_off) : z2;
    const double2 c_ky = m.ld_ok ? tl_lds2(base + 2 * 512 + lane_off) : z2;
    const double2 c_kx = m.ld_ok ? tl_lds2(base + 3 * 512 + lane_off) : z2;
    double2 c_u = (upd && m.acta) ? tl_lds2(base + 4 * 512 + lane_off) : z2;
    if (upd2 && upd && m.acta) {
      const double2 c_po = tl_lds2(base + 5 * 512 + lane_off);
      c_u.x = c_u.x + alpha_prev2 * c_po.x;
      c_u.y = c_u.y + alpha_prev2 * c_po.y;
    }
    const double c_re = m.has_edge ? tl_lds1(base + edge_off + 0) : 0.0;
    const double c_pe = m.has_edge ? tl_lds1(base + edge_off + 8) : 0.0;
    const double c_kxe = (m.lane == 31 && m.has_edge) ? tl_lds1(base + edge_off + 16) : 0.0;
    stage = (stage + 1 == S) ? 0 : stage + 1;
    fill = (fill + 1 == S) ? 0 : fill + 1;

    const double2 Xn = comb2(c_r, c_p);
    const double XnE = comb(c_re, c_pe);
    double xl = __shfl_up_sync(0xffffffffu, Xc.y, 1);
    double xr = __shfl_down_sync(0xffffffffu, Xc.x, 1);
    double kxr = __shfl_down_sync(0xffffffffu, c_kx.x, 1);
    if (m.lane == 0) xl = XcE;
    if (m.lane == 31) { xr = XcE; kxr = c_kxe; }
    const double La = (physL && m.i0 == 0) ? Xc.x : xl;
    const double Ra = (physR && m.i0 == g.nx - 1) ? Xc.x : Xc.y;
    const double Lb = Xc.x;
    const double Rb = (physR && m.i0 + 1 == g.nx - 1) ? Xc.y : xr;
    const double wa = ((((1.0 + c_kx.y) + c_kx.x) + c_ky.x) + kyc.x) * Xc.x -
                      (c_kx.y * Ra + c_kx.x * La) - (c_ky.x * Xn.x + kyc.x * Xm.x);
    const double wb = ((((1.0 + kxr) + c_kx.y) + c_ky.y) + kyc.y) * Xc.y -
                      (kxr * Rb + c_kx.y * Lb) - (c_ky.y * Xn.y + kyc.y * Xm.y);
    const long oc = (long)j * pitch + m.i0;
    double2 un = z2;
    if (UM != TL_U_NONE) un = make_double2(c_u.x + alpha_prev * pc.x, c_u.y + alpha_prev * pc.y);
    if (m.actb) {
      tl_st2_hint(w + oc, make_double2(wa, wb), pol_keep);
      tl_st2_hint(pout + oc, Xc, pol_stream);
      if (upd) tl_st2_hint(u + oc, un, pol_stream);
      acc0 += wa * Xc.x;
      acc0 += wb * Xc.y;
    } else if (m.acta) {
      w[oc] = wa; pout[oc] = Xc.x;
      if (upd) u[oc] = un.x;
      acc0 += wa * Xc.x;
    }
    // haloupdate!(.., [:u,:p]) CG.jl:22: reflective sides as a write-through, tile-internal
    // sides as a push of p into the neighbour's halo (u's internal halos are filled after the loop)
    tl_reflect_edges(pout, g, m, j, oc, Xc);
    if (upd) tl_reflect_edges(u, g, m, j, oc, un);
    if (tiled) tl_push_edges(push, g, m, j, Xc);
    Xm = Xc; Xc = Xn; XcE = XnE; pc = c_p; kyc = c_ky;
  }
  tl_cp_wait<0>();
}

template <int UM, int S, int MINB>
__global__ void __launch_bounds__(TL_FUSED_THREADS, MINB) k_cg_fused_w_ring(const CgAParams P) {
  tl_pdl_entry();
  extern __shared__ __align__(128) unsigned char ring_raw[];
  __shared__ double sm[32];
  SolveState *st = P.st;
  tl_prof_entry(st);
  const TlScal sc = tl_entry_scalars(P.cd, st);
  const int it = st->iter;
  const double rr_cur = sc.rr;
  if (st->comm_error || tl_should_stop(it, rr_cur, st->cfg)) return;
  CgAIter I;
  I.it = it;
  I.first = (it == st->cfg.first_it);
  I.beta = 0.0; I.alpha_prev = 0.0;
  I.alpha_prev2 = 0.0;
  if (P.lazy_role) {   // the host alternates the kernels of a lazy loop: this one must sit on an iteration of its parity
    const int off = it - st->cfg.first_it;
    if ((off & 1) != (P.lazy_role == 1 ? 1 : 0)) {
      if (blockIdx.x == 0 && threadIdx.x == 0) st->sched_error = 1;
      return;
    }
  }
  if (!I.first) {
    const double rr_prev = P.hist_rr[it - 1];
    I.beta = rr_cur / rr_prev;
    I.alpha_prev = rr_prev / P.hist_pw[it];
    if (UM == TL_U_LAZY) I.alpha_prev2 = P.hist_rr[it - 2] / P.hist_pw[it - 1];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) P.hist_rr[it] = rr_cur;
  double acc[1] = {0.0};
  tl_cg_a_item<UM, S, false>(P, I, blockIdx.x, ring_raw, acc[0]);
  if (tl_kernel_tail(acc, true, st, P.partials, P.cd, sm, TL_T_PW)) {
    st->red_pw_local = acc[0];
    if ((P.single || P.cd != nullptr) && !tl_is_deferred(P.cd)) st->red_pw = acc[0];
  }
}

// ------------------------------------------------------------------------------------------
// CG kernel B with the cp.async ring (k_cg_fused_r_ring): the same rows, the same arithmetic and
// the same summation order as tl_cg_b_item, but the next D-1 rows of r and w are always in flight
// as asynchronous copies into this warp's private ring (1 KB per row: r at +0, w at +512), instead
// of batches of 4 rows loaded into registers -- the kernel is pure streaming, so bytes in flight
// are its only lever.  Lanes read back only the bytes they copied themselves: no barriers.
// ------------------------------------------------------------------------------------------
#define TL_BRING_ROW_BYTES 1024
template <int D, bool COH>
__device__ __forceinline__ void tl_cg_b_item_ring(const CgBParams &P, double alpha, int blk, unsigned char *ring_raw,
                                                  double &acc0) {
  const Geo g = P.g;
  double *r = P.r;
  const double *w = P.w;
  const unsigned long long pol_r = tl_policy(P.hint_keep), pol_w = tl_policy(P.hint_stream);
  const bool tiled = P.cd != nullptr;
  MarchCtx m;
  if (!tl_march_setup_blk(g, P.t, m, blk, P.reverse)) return;
  if (m.actb) {
    const unsigned ring = tl_smem_u32(ring_raw) + (threadIdx.x >> 5) * (D * TL_BRING_ROW_BYTES) + m.lane * 16;
    auto issue = [&](int j, int slot) {
      const long o = (long)j * g.pitch + m.i0;
      tl_cp16_hint(ring + slot * TL_BRING_ROW_BYTES, r + o, pol_r);
      tl_cp16_hint(ring + slot * TL_BRING_ROW_BYTES + 512, w + o, pol_w);
    };
#pragma unroll
    for (int d = 0; d < D - 1; d++) {
      if (m.j0 + d < m.j1) issue(m.j0 + d, d);
      tl_cp_commit();
    }
    int stage = 0, fill = D - 1;
    for (int j = m.j0; j < m.j1; j++) {
      if (j + D - 1 < m.j1) issue(j + D - 1, fill);
      tl_cp_commit();
      tl_cp_wait<D - 1>();
      double2 rv = tl_lds2(ring + stage * TL_BRING_ROW_BYTES);
      const double2 wv = tl_lds2(ring + stage * TL_BRING_ROW_BYTES + 512);
      stage = (stage + 1 == D) ? 0 : stage + 1;
      fill = (fill + 1 == D) ? 0 : fill + 1;
      rv.x = rv.x - alpha * wv.x;
      rv.y = rv.y - alpha * wv.y;
      tl_st2_hint(r + (long)j * g.pitch + m.i0, rv, pol_r);
      acc0 += rv.x * rv.x;
      acc0 += rv.y * rv.y;
      if (tiled) tl_push_edges(P.push_r, g, m, j, rv);
    }
    tl_cp_wait<0>();
  } else if (m.acta) {   // the single last cell of an odd-width tile
    for (int j = m.j0; j < m.j1; j++) {
      const long o = (long)j * g.pitch + m.i0;
      const double v = tl_ld1<COH>(r + o) - alpha * tl_ld1<COH>(w + o);
      r[o] = v;
      acc0 += v * v;
      if (tiled) tl_push_edges(P.push_r, g, m, j, make_double2(v, 0.0));
    }
  }
}

template <int D, int MINB>
__global__ void __launch_bounds__(TL_FUSED_THREADS, MINB) k_cg_fused_r_ring(const CgBParams P) {
  tl_pdl_entry();
  extern __shared__ __align__(128) unsigned char ring_raw[];
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
  tl_cg_b_item_ring<D, false>(P, alpha, blockIdx.x, ring_raw, acc[0]);
  if (tl_kernel_tail(acc, true, st, P.partials, P.cd, sm, TL_T_RR)) {
    st->red_rr_local = acc[0];
    if ((P.single || P.cd != nullptr) && !tl_is_deferred(P.cd)) st->red_rr = acc[0];
    st->iter = it + 1;
  }
}

// ------------------------------------------------------------------------------------------
// Shared pieces of the single-operand stencil kernels (Chebyshev: operand u; PPCG inner: sd).
// Ring slot: field 0 = operand at row jn, 1 = ky(j+1), 2 = kx(j), 3/4 = two pointwise fields
// at row j; edges: operand at (ecol, jn), kx at (ecol, j) for lane 31.
// ------------------------------------------------------------------------------------------
struct RingRow {
  double2 x, ky, kx, a, b;
  double xe, kxe;
};

template <int S>
struct RingMarch {
  unsigned ring, lane_off, edge_off;
  int stage, fill;
  __device__ __forceinline__ void init(const unsigned char *raw, const MarchCtx &m) {
    ring = tl_smem_u32(raw) + (threadIdx.x >> 5) * (S * TL_RING_STAGE_BYTES);
    lane_off = m.lane * 16;
    edge_off = TL_RING_FIELDS * 512 + (m.lane == 0 ? 0 : 24);
    stage = 0;
    fill = S - 1;
  }
  // fa / fb may be null (field not needed)
  __device__ __forceinline__ void issue(const Geo &g, const MarchCtx &m, bool physT, int j, int slot,
                                        const double *x, const double *ky, const double *kx, const double *fa,
                                        const double *fb) const {
    const unsigned base = ring + slot * TL_RING_STAGE_BYTES;
    const int jn = (j + 1 >= g.ny && physT) ? g.ny - 1 : j + 1;
    const long on = (long)jn * g.pitch + m.i0, oc = (long)j * g.pitch + m.i0;
    if (m.ld_ok) {
      tl_cp16(base + 0 * 512 + lane_off, x + on);
      tl_cp16(base + 1 * 512 + lane_off, ky + oc + g.pitch);
      tl_cp16(base + 2 * 512 + lane_off, kx + oc);
    }
    if (m.acta) {
      if (fa) tl_cp16(base + 3 * 512 + lane_off, fa + oc);
      if (fb) tl_cp16(base + 4 * 512 + lane_off, fb + oc);
    }
    if (m.has_edge) {
      tl_cp8(base + edge_off + 0, x + (long)jn * g.pitch + m.ecol);
      if (m.lane == 31) tl_cp8(base + edge_off + 8, kx + (long)j * g.pitch + m.ecol);
    }
  }
  __device__ __forceinline__ RingRow take(const MarchCtx &m, bool has_a, bool has_b) {
    const double2 z2 = make_double2(0.0, 0.0);
    const unsigned base = ring + stage * TL_RING_STAGE_BYTES;
    RingRow r;
    r.x = m.ld_ok ? tl_lds2(base + 0 * 512 + lane_off) : z2;
    r.ky = m.ld_ok ? tl_lds2(base + 1 * 512 + lane_off) : z2;
    r.kx = m.ld_ok ? tl_lds2(base + 2 * 512 + lane_off) : z2;
    r.a = (has_a && m.acta) ? tl_lds2(base + 3 * 512 + lane_off) : z2;
    r.b = (has_b && m.acta) ? tl_lds2(base + 4 * 512 + lane_off) : z2;
    r.xe = m.has_edge ? tl_lds1(base + edge_off + 0) : 0.0;
    r.kxe = (m.lane == 31 && m.has_edge) ? tl_lds1(base + edge_off + 8) : 0.0;
    stage = (stage + 1 == S) ? 0 : stage + 1;
    fill = (fill + 1 == S) ? 0 : fill + 1;
    return r;
  }
};

// w = A x at this lane's two cells (same expression order as tl_smvp / the oracle)
__device__ __forceinline__ void tl_stencil2(const Geo &g, const MarchCtx &m, bool physL, bool physR, double2 Xm,
                                            double2 Xc, double2 Xn, double XcE, double2 kxv, double kxe, double2 kyc,
                                            double2 kyn, double &wa, double &wb) {
  double xl = __shfl_up_sync(0xffffffffu, Xc.y, 1);
  double xr = __shfl_down_sync(0xffffffffu, Xc.x, 1);
  double kxr = __shfl_down_sync(0xffffffffu, kxv.x, 1);
  if (m.lane == 0) xl = XcE;
  if (m.lane == 31) { xr = XcE; kxr = kxe; }
  const double La = (physL && m.i0 == 0) ? Xc.x : xl;
  const double Ra = (physR && m.i0 == g.nx - 1) ? Xc.x : Xc.y;
  const double Lb = Xc.x;
  const double Rb = (physR && m.i0 + 1 == g.nx - 1) ? Xc.y : xr;
  wa = ((((1.0 + kxv.y) + kxv.x) + kyn.x) + kyc.x) * Xc.x - (kxv.y * Ra + kxv.x * La) - (kyn.x * Xn.x + kyc.x * Xm.x);
  wb = ((((1.0 + kxr) + kxv.y) + kyn.y) + kyc.y) * Xc.y - (kxr * Rb + kxv.y * Lb) - (kyn.y * Xn.y + kyc.y * Xm.y);
}

// Chebyshev iteration (algorithm and citations: ChebyParams in tl_kernels_fused.cuh).
template <bool FIRST, int S, int MINB>
__global__ void __launch_bounds__(TL_FUSED_THREADS, MINB) k_cheby_fused_ring(const ChebyParams P) {
  tl_pdl_entry();
  extern __shared__ __align__(128) unsigned char ring_raw[];
  __shared__ double sm[32];
  SolveState *st = P.st;
  tl_prof_entry(st);
  const TlScal sc = tl_entry_scalars(P.cd, st);
  const int step = st->cheby_step;
  double alpha = 0.0, beta = 0.0;
  bool calc_norm, store_wr;
  const double theta = st->theta;
  if (FIRST) {
    calc_norm = true;
    store_wr = true;
    if (st->comm_error) return;
  } else {
    if (st->comm_error || tl_cheby_should_stop_n(*st, sc.norm)) return;
    const int chebyiters = step;
    const int tt = st->cheby_tt0 + chebyiters - 1;
    alpha = P.alphas[chebyiters];
    beta = P.betas[chebyiters];
    calc_norm = tl_cheby_is_norm_iter(chebyiters, st->cheby_tt0, st->cheby_est);
    store_wr = calc_norm || (tt == st->cheby_max_tt);
  }
  const int pairs = FIRST ? 0 : st->cheby_pairs;
  const int upar = tl_cheby_u_parity(step, pairs);
  const double *__restrict__ uin = upar ? P.ub : P.ua;
  double *__restrict__ uout = upar ? P.ua : P.ub;
  const bool tiled = P.cd != nullptr;
  const Push &push = upar ? P.push_ua : P.push_ub;   // halo targets of uout
  const double *__restrict__ u0 = P.u0;
  const double *__restrict__ kx = P.kx;
  const double *__restrict__ ky = P.ky;
  double *__restrict__ p = (pairs & 1) ? P.p1 : P.p;
  const Geo g = P.g;
  const int pitch = g.pitch;
  const bool physL = g.phys & TL_PHYS_LEFT, physR = g.phys & TL_PHYS_RIGHT;
  const bool physB = g.phys & TL_PHYS_BOTTOM, physT = g.phys & TL_PHYS_TOP;

  double acc[1] = {0.0};
  MarchCtx m;
  if (tl_march_setup(g, P.t, m)) {
    const double2 z2 = make_double2(0.0, 0.0);
    RingMarch<S> rg;
    rg.init(ring_raw, m);
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
#pragma unroll
    for (int d = 0; d < S - 1; d++) {
      if (m.j0 + d < m.j1) rg.issue(g, m, physT, m.j0 + d, d, uin, ky, kx, u0, FIRST ? nullptr : p);
      tl_cp_commit();
    }
    for (int j = m.j0; j < m.j1; j++) {
      if (j + S - 1 < m.j1) rg.issue(g, m, physT, j + S - 1, rg.fill, uin, ky, kx, u0, FIRST ? nullptr : p);
      tl_cp_commit();
      tl_cp_wait<S - 1>();
      const RingRow cur = rg.take(m, true, !FIRST);
      const double2 Xn = cur.x;
      double wa, wb;
      tl_stencil2(g, m, physL, physR, Xm, Xc, Xn, XcE, cur.kx, cur.kxe, kyc, cur.ky, wa, wb);
      const double ra = cur.a.x - wa, rb = cur.a.y - wb;
      double2 pn;
      if (FIRST) { pn.x = ra / theta; pn.y = rb / theta; }
      else { pn.x = alpha * cur.b.x + beta * ra; pn.y = alpha * cur.b.y + beta * rb; }
      const double2 un = make_double2(Xc.x + pn.x, Xc.y + pn.y);
      const long oc = (long)j * pitch + m.i0;
      if (m.actb) {
        tl_st2(p + oc, pn);
        tl_st2(uout + oc, un);
        if (store_wr) { tl_st2(P.w + oc, make_double2(wa, wb)); tl_st2(P.r + oc, make_double2(ra, rb)); }
        if (FIRST) { acc[0] += cur.a.x * cur.a.x; acc[0] += cur.a.y * cur.a.y; }
        else { acc[0] += ra * ra; acc[0] += rb * rb; }
      } else if (m.acta) {
        p[oc] = pn.x; uout[oc] = un.x;
        if (store_wr) { P.w[oc] = wa; P.r[oc] = ra; }
        acc[0] += FIRST ? cur.a.x * cur.a.x : ra * ra;
      }
      // haloupdate!(.., [:u]) Cheby.jl:55/:78
      tl_reflect_edges(uout, g, m, j, oc, un);
      if (tiled) tl_push_edges(push, g, m, j, un);
      Xm = Xc; Xc = Xn; XcE = cur.xe; kyc = cur.ky;
    }
    tl_cp_wait<0>();
  }
  // no reduction on most iterations: only the ticket (and, tiled, the completion barrier)
  if (tl_kernel_tail(acc, calc_norm, st, P.partials, P.cd, sm, TL_T_NORM)) {
    if (calc_norm) {
      st->red_norm_local = acc[0];
      if ((P.single || tiled) && !tl_is_deferred(P.cd)) st->red_norm = acc[0];
    }
    st->cheby_step = step + 1;
  }
}

// ------------------------------------------------------------------------------------------
// TWO Chebyshev iterations in one pass (k_cheby_pair_ring) -- temporal blocking of the
// reduction-free iterations (Cheby.mainstep! Cheby.jl:84-106 applied twice):
//   step A (chebyiters = s):    wA = A u ;  rA = u0 - wA ; pA = alpha_s p + beta_s rA ;       uA = u + pA
//   step B (chebyiters = s+1):  wB = A uA ; rB = u0 - wB ; pB = alpha_s+1 pA + beta_s+1 rB ;  uB = uA + pB
// reads u, u0, p, kx, ky once and writes pB, uB once: ~30 B per cell-iteration instead of 56.
// uA and pA never reach memory.  A warp owns 60 columns x `rows_per_chunk` rows but marches over a
// window of 64 columns (two more on each side) and two more rows (one below, one above): the uA
// values step B needs from neighbouring warps are recomputed from the same inputs with the same
// expressions, hence bit-identical, and all x-neighbours come from warp shuffles (no edge loads:
// the outermost window columns only ever feed values nobody uses).  Physical boundaries: the
// reflective rule by clamping, for u and for uA alike (haloupdate!(u) between the iterations).
// Because neighbouring warps still need the OLD p, p is written out of place (p <-> p1) as u is.
// Step A is never a norm iteration (norm iterations have odd tt, Cheby.jl:40-51, and the host
// starts pairs on even tt); step B may be: its norm is summed from registers, and w / r are
// stored when its state is observable, exactly as in k_cheby_fused_ring.
//
// TILED = true is the same pass on a TILE of a decomposed mesh (multi-GPU): on tile-internal sides
// nothing is clamped -- the window columns -2, -1 / nx, nx+1 and the rows -2, -1 / ny, ny+1 are read
// from the halo, which holds the neighbours' u two cells deep (corner blocks included), p and u0 one
// cell deep and kx, ky two deep (tests/emulation/emulate_pair_tiled.py proves these depths sufficient
// and necessary).  u' and p' are pushed that deep into the EIGHT surrounding tiles from step B, and
// the mailbox exchange in the tail is the completion barrier -- ONE rendezvous per two iterations:
// the pair kernel is the depth-2 matrix-powers Chebyshev step.
// ------------------------------------------------------------------------------------------
#define TL_PAIR_OWN 60   // owned columns per warp (window: 64)
struct ChebyPairParams {
  ChebyParams c;
  Push8 push_ua, push_ub, push_p0, push_p1;   // TILED: halo targets of the u / p buffer being written
};
template <int S, int MINB, bool TILED>
__global__ void __launch_bounds__(TL_FUSED_THREADS, MINB) k_cheby_pair_ring(const ChebyPairParams PT) {
  const ChebyParams &P = PT.c;
  tl_pdl_entry();
  extern __shared__ __align__(128) unsigned char ring_raw[];
  __shared__ double sm[32];
  SolveState *st = P.st;
  tl_prof_entry(st);
  const CommDev *xcd = TILED ? P.cd : nullptr;
  const TlScal sc = tl_entry_scalars(xcd, st);
  const int step = st->cheby_step;
  if (st->comm_error || tl_cheby_should_stop_n(*st, sc.norm)) return;
  const int ttA = st->cheby_tt0 + step - 1;
  if (ttA + 1 > st->cheby_max_tt) return;                                            // step B not permitted
  if (tl_cheby_is_norm_iter(step, st->cheby_tt0, st->cheby_est)) return;            // never (see above): guard
  const double alphaA = P.alphas[step], betaA = P.betas[step];
  const double alphaB = P.alphas[step + 1], betaB = P.betas[step + 1];
  const bool calc_norm = tl_cheby_is_norm_iter(step + 1, st->cheby_tt0, st->cheby_est);
  const bool store_wr = calc_norm || (ttA + 1 == st->cheby_max_tt);
  const int pairs = st->cheby_pairs;
  const int upar = tl_cheby_u_parity(step, pairs);
  const double *__restrict__ uin = upar ? P.ub : P.ua;
  double *__restrict__ uout = upar ? P.ua : P.ub;
  const double *__restrict__ pin = (pairs & 1) ? P.p1 : P.p;
  double *__restrict__ pout = (pairs & 1) ? P.p : P.p1;
  const Push8 &push_u = upar ? PT.push_ua : PT.push_ub;            // halo targets of uout
  const Push8 &push_p = (pairs & 1) ? PT.push_p0 : PT.push_p1;     // ... of pout
  const double *__restrict__ u0 = P.u0;
  const double *__restrict__ kx = P.kx;
  const double *__restrict__ ky = P.ky;
  const Geo g = P.g;
  const int pitch = g.pitch;
  const bool physL = g.phys & TL_PHYS_LEFT, physR = g.phys & TL_PHYS_RIGHT;
  const bool physB = g.phys & TL_PHYS_BOTTOM, physT = g.phys & TL_PHYS_TOP;

  double acc[1] = {0.0};
  const int lane = threadIdx.x & 31;
  const int wt = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wt < P.t.nstrips * P.t.nchunks) {
    const int s = wt % P.t.nstrips, q = wt / P.t.nstrips;
    const int j0 = q * P.t.rows_per_chunk, j1 = min(g.ny, j0 + P.t.rows_per_chunk);
    const int own_lo = s * TL_PAIR_OWN, own_hi = min(g.nx, own_lo + TL_PAIR_OWN);
    MarchCtx m;                            // the WINDOW: what is loaded and computed
    m.lane = lane;
    m.i0 = own_lo - 2 + 2 * lane;
    m.j0 = j0; m.j1 = j1;
    // the pair (nx, nx+1) is still inside the padded row; on a tile-internal right side with odd nx, u(nx+1)
    // sits in the pair (nx+1, nx+2)
    m.ld_ok = m.i0 <= g.nx + ((TILED && !physR) ? 1 : 0);
    m.acta = m.ld_ok;                      // RingMarch loads the pointwise fields as pairs
    m.actb = m.ld_ok;
    m.ecol = 0; m.has_edge = false;        // no edge loads: neighbours come from the window
    MarchCtx mo = m;                       // the OWNED cells: what is stored
    mo.acta = m.i0 >= own_lo && m.i0 < own_hi;
    mo.actb = m.i0 + 1 >= own_lo && m.i0 + 1 < own_hi && mo.acta;
    if (j0 < j1) {
      const double2 z2 = make_double2(0.0, 0.0);
      // rows on which uA is needed: the chunk plus one row below / above (not beyond a physical side)
      const int ja_lo = (j0 == 0 && physB) ? 0 : j0 - 1;
      const int ja_hi = (j1 == g.ny && physT) ? g.ny - 1 : j1;
      RingMarch<S> rg;
      rg.init(ring_raw, m);
      double2 Um, Uc, kyc;
      {
        const int jm = (ja_lo == 0 && physB) ? 0 : ja_lo - 1;
        const long om = (long)jm * pitch + m.i0, oc = (long)ja_lo * pitch + m.i0;
        Um = m.ld_ok ? tl_ld2(uin + om) : z2;
        Uc = m.ld_ok ? tl_ld2(uin + oc) : z2;
        kyc = m.ld_ok ? tl_ld2(ky + oc) : z2;
      }
#pragma unroll
      for (int d = 0; d < S - 1; d++) {
        if (ja_lo + d <= ja_hi) rg.issue(g, m, physT, ja_lo + d, d, uin, ky, kx, u0, pin);
        tl_cp_commit();
      }
      // carried from the previous row (row j = jj - 1 of step B): uA(j-1), uA(j), pA(j), u0(j), kx(j), ky(j)
      double2 Am = z2, Ac = z2, pAc = z2, u0c = z2, kxc = z2, kyB = z2;
      auto step_b = [&](int j, double2 An, double2 kyn) {
        const double2 Bm = (j == 0 && physB) ? Ac : Am;          // uA(-1) := uA(0)
        double wa, wb;
        tl_stencil2(g, m, physL, physR, Bm, Ac, An, 0.0, kxc, 0.0, kyB, kyn, wa, wb);
        const double ra = u0c.x - wa, rb = u0c.y - wb;
        const double2 pn = make_double2(alphaB * pAc.x + betaB * ra, alphaB * pAc.y + betaB * rb);
        const double2 un = make_double2(Ac.x + pn.x, Ac.y + pn.y);
        const long oc = (long)j * pitch + m.i0;
        if (mo.actb) {
          tl_st2(pout + oc, pn);
          tl_st2(uout + oc, un);
          if (store_wr) { tl_st2(P.w + oc, make_double2(wa, wb)); tl_st2(P.r + oc, make_double2(ra, rb)); }
          acc[0] += ra * ra; acc[0] += rb * rb;
        } else if (mo.acta) {
          pout[oc] = pn.x; uout[oc] = un.x;
          if (store_wr) { P.w[oc] = wa; P.r[oc] = ra; }
          acc[0] += ra * ra;
        }
        tl_reflect_edges(uout, g, mo, j, oc, un);                // haloupdate!(.., [:u]) Cheby.jl:55/:78
        if (TILED) {   // tile-internal sides: the next pass reads u two cells deep (corner blocks included), p one cell deep
          tl_push_deep(push_u, g, 2, m.i0, mo.acta, mo.actb, j, un);
          tl_push_deep(push_p, g, 1, m.i0, mo.acta, mo.actb, j, pn);
        }
      };
      for (int jj = ja_lo; jj <= ja_hi; jj++) {
        if (jj + S - 1 <= ja_hi) rg.issue(g, m, physT, jj + S - 1, rg.fill, uin, ky, kx, u0, pin);
        tl_cp_commit();
        tl_cp_wait<S - 1>();
        const RingRow cur = rg.take(m, true, true);              // x = u(jj+1), ky = ky(jj+1), kx = kx(jj), a = u0(jj), b = p(jj)
        double wa, wb;
        tl_stencil2(g, m, physL, physR, Um, Uc, cur.x, 0.0, cur.kx, 0.0, kyc, cur.ky, wa, wb);
        const double ra = cur.a.x - wa, rb = cur.a.y - wb;
        const double2 pA = make_double2(alphaA * cur.b.x + betaA * ra, alphaA * cur.b.y + betaA * rb);
        const double2 uA = make_double2(Uc.x + pA.x, Uc.y + pA.y);
        if (jj - 1 >= j0) step_b(jj - 1, uA, kyc);               // kyc = ky(jj) = ky(j+1)
        Am = Ac; Ac = uA; pAc = pA; u0c = cur.a; kxc = cur.kx; kyB = kyc;
        Um = Uc; Uc = cur.x; kyc = cur.ky;
      }
      // top row of a physical top: uA(ny) := uA(ny-1)
      if (ja_hi == j1 - 1) step_b(j1 - 1, Ac, kyc);
      tl_cp_wait<0>();
    }
  }
  // tiles: all-tiles norm / completion barrier of the halo pushes
  if (tl_kernel_tail(acc, calc_norm, st, P.partials, xcd, sm, TL_T_NORM)) {
    if (calc_norm) {
      st->red_norm_local = acc[0];
      if (!tl_is_deferred(xcd)) st->red_norm = acc[0];
    }
    st->cheby_step = step + 2;
    st->cheby_pairs = pairs + 1;
  }
}

// Two PPCG inner steps in one pass (algorithm: PpcgPairParams in tl_kernels_fused.cuh; window,
// carry and clamp scheme: k_cheby_pair_ring above).  TILED: sd is read two cells deep (corner blocks
// included) and r one cell deep from the tile-internal halos, kx / ky two deep; step B pushes sd' and r'
// that deep into the eight surrounding tiles (after the last pair of an outer iteration only r, one cell
// deep: the next matvec reads it), and the tail's exchange is the completion barrier.
template <int S, int MINB, bool TILED>
__global__ void __launch_bounds__(TL_FUSED_THREADS, MINB) k_ppcg_pair_ring(const PpcgPairParams P) {
  tl_pdl_entry();
  extern __shared__ __align__(128) unsigned char ring_raw[];
  __shared__ double sm[32];
  SolveState *st = P.st;
  tl_prof_entry(st);
  const CommDev *xcd = TILED ? P.cd : nullptr;
  const TlScal sc = tl_entry_scalars(xcd, st);
  const int it = st->iter;
  if (st->comm_error || tl_should_stop(it, sc.rr, st->cfg)) return;
  const int pp = st->inner_pp;
  const bool last = (pp + 2 == st->inner_steps);
  const double alphaA = P.alphas[pp], betaA = P.betas[pp];
  const double alphaB = P.alphas[pp + 1], betaB = P.betas[pp + 1];
  const double *__restrict__ sin = P.sin;
  double *__restrict__ sout = P.sout;
  const double *__restrict__ rin = P.rin;
  double *__restrict__ rout = P.rout;
  const double *__restrict__ kx = P.kx;
  const double *__restrict__ ky = P.ky;
  double *u = P.u;
  const Geo g = P.g;
  const int pitch = g.pitch;
  const bool physL = g.phys & TL_PHYS_LEFT, physR = g.phys & TL_PHYS_RIGHT;
  const bool physB = g.phys & TL_PHYS_BOTTOM, physT = g.phys & TL_PHYS_TOP;

  double acc[1] = {0.0};
  const int lane = threadIdx.x & 31;
  const int wt = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wt < P.t.nstrips * P.t.nchunks) {
    const int s = wt % P.t.nstrips, q = wt / P.t.nstrips;
    const int j0 = q * P.t.rows_per_chunk, j1 = min(g.ny, j0 + P.t.rows_per_chunk);
    const int own_lo = s * TL_PAIR_OWN, own_hi = min(g.nx, own_lo + TL_PAIR_OWN);
    MarchCtx m;                            // the WINDOW
    m.lane = lane;
    m.i0 = own_lo - 2 + 2 * lane;
    m.j0 = j0; m.j1 = j1;
    m.ld_ok = m.i0 <= g.nx + ((TILED && !physR) ? 1 : 0);   // as in k_cheby_pair_ring
    m.acta = m.ld_ok;
    m.actb = m.ld_ok;
    m.ecol = 0; m.has_edge = false;
    MarchCtx mo = m;                       // the OWNED cells
    mo.acta = m.i0 >= own_lo && m.i0 < own_hi;
    mo.actb = m.i0 + 1 >= own_lo && m.i0 + 1 < own_hi && mo.acta;
    if (j0 < j1) {
      const double2 z2 = make_double2(0.0, 0.0);
      const int ja_lo = (j0 == 0 && physB) ? 0 : j0 - 1;
      const int ja_hi = (j1 == g.ny && physT) ? g.ny - 1 : j1;
      RingMarch<S> rg;
      rg.init(ring_raw, m);
      double2 Sm, Sc, kyc;
      {
        const int jm = (ja_lo == 0 && physB) ? 0 : ja_lo - 1;
        const long om = (long)jm * pitch + m.i0, oc = (long)ja_lo * pitch + m.i0;
        Sm = m.ld_ok ? tl_ld2(sin + om) : z2;
        Sc = m.ld_ok ? tl_ld2(sin + oc) : z2;
        kyc = m.ld_ok ? tl_ld2(ky + oc) : z2;
      }
#pragma unroll
      for (int d = 0; d < S - 1; d++) {
        if (ja_lo + d <= ja_hi) rg.issue(g, m, physT, ja_lo + d, d, sin, ky, kx, rin, u);
        tl_cp_commit();
      }
      // carried for step B on row j = jj - 1: sA(j-1), sA(j), rA(j), uA(j), kx(j), ky(j)
      double2 Am = z2, Ac = z2, rAc = z2, uAc = z2, kxc = z2, kyB = z2;
      auto step_b = [&](int j, double2 An, double2 kyn) {
        const double2 Bm = (j == 0 && physB) ? Ac : Am;
        double wa, wb;
        tl_stencil2(g, m, physL, physR, Bm, Ac, An, 0.0, kxc, 0.0, kyB, kyn, wa, wb);
        const double2 rn = make_double2(rAc.x - wa, rAc.y - wb);
        const double2 un = make_double2(uAc.x + Ac.x, uAc.y + Ac.y);
        const double2 sn = make_double2(alphaB * Ac.x + betaB * rn.x, alphaB * Ac.y + betaB * rn.y);
        const long oc = (long)j * pitch + m.i0;
        if (mo.actb) {
          tl_st2(rout + oc, rn); tl_st2(u + oc, un); tl_st2(sout + oc, sn);
          acc[0] += rn.x * rn.x;
          acc[0] += rn.y * rn.y;
        } else if (mo.acta) {
          rout[oc] = rn.x; u[oc] = un.x; sout[oc] = sn.x;
          acc[0] += rn.x * rn.x;
        }
        // halo(sd) of PPCG.jl:76 happens BEFORE each inner step (see k_ppcg_inner_ring)
        tl_reflect_edges(sout, g, mo, j, oc, last ? Ac : sn);
        if (TILED) {
          if (!last) tl_push_deep(P.push_s, g, 2, m.i0, mo.acta, mo.actb, j, sn);
          tl_push_deep(P.push_r, g, 1, m.i0, mo.acta, mo.actb, j, rn);
        }
      };
      for (int jj = ja_lo; jj <= ja_hi; jj++) {
        if (jj + S - 1 <= ja_hi) rg.issue(g, m, physT, jj + S - 1, rg.fill, sin, ky, kx, rin, u);
        tl_cp_commit();
        tl_cp_wait<S - 1>();
        const RingRow cur = rg.take(m, true, true);      // x = sd(jj+1), ky(jj+1), kx(jj), a = r(jj), b = u(jj)
        double wa, wb;
        tl_stencil2(g, m, physL, physR, Sm, Sc, cur.x, 0.0, cur.kx, 0.0, kyc, cur.ky, wa, wb);
        const double2 rA = make_double2(cur.a.x - wa, cur.a.y - wb);
        const double2 uA = make_double2(cur.b.x + Sc.x, cur.b.y + Sc.y);
        const double2 sA = make_double2(alphaA * Sc.x + betaA * rA.x, alphaA * Sc.y + betaA * rA.y);
        if (jj - 1 >= j0) step_b(jj - 1, sA, kyc);
        Am = Ac; Ac = sA; rAc = rA; uAc = uA; kxc = cur.kx; kyB = kyc;
        Sm = Sc; Sc = cur.x; kyc = cur.ky;
      }
      if (ja_hi == j1 - 1) step_b(j1 - 1, Ac, kyc);
      tl_cp_wait<0>();
    }
  }
  if (tl_kernel_tail(acc, last, st, P.partials, xcd, sm, TL_T_RR)) {
    if (last) {
      st->red_rr_local = acc[0];      // PPCG.jl:88
      if (!tl_is_deferred(xcd)) st->red_rr = acc[0];
      st->iter = it + 1;
    }
    st->inner_pp = pp + 2;
  }
}

// PPCG inner step (algorithm and citations: PpcgInnerParams in tl_kernels_fused.cuh).
template <int S, int MINB>
__global__ void __launch_bounds__(TL_FUSED_THREADS, MINB) k_ppcg_inner_ring(const PpcgInnerParams P) {
  tl_pdl_entry();
  extern __shared__ __align__(128) unsigned char ring_raw[];
  __shared__ double sm[32];
  SolveState *st = P.st;
  tl_prof_entry(st);
  const TlScal sc = tl_entry_scalars(P.cd, st);
  const int it = st->iter;
  if (st->comm_error || tl_should_stop(it, sc.rr, st->cfg)) return;
  const int pp = st->inner_pp;
  const bool last = (pp + 1 == st->inner_steps);
  const double alpha = P.alphas[pp], beta = P.betas[pp];
  const double *__restrict__ sin = (pp & 1) ? P.sdb : P.sda;
  double *__restrict__ sout = (pp & 1) ? P.sda : P.sdb;
  const bool tiled = P.cd != nullptr;
  const Push &push = (pp & 1) ? P.push_sda : P.push_sdb;   // halo targets of sout
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
    RingMarch<S> rg;
    rg.init(ring_raw, m);
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
#pragma unroll
    for (int d = 0; d < S - 1; d++) {
      if (m.j0 + d < m.j1) rg.issue(g, m, physT, m.j0 + d, d, sin, ky, kx, r, u);
      tl_cp_commit();
    }
    for (int j = m.j0; j < m.j1; j++) {
      if (j + S - 1 < m.j1) rg.issue(g, m, physT, j + S - 1, rg.fill, sin, ky, kx, r, u);
      tl_cp_commit();
      tl_cp_wait<S - 1>();
      const RingRow cur = rg.take(m, true, true);
      const double2 Xn = cur.x;
      double wa, wb;
      tl_stencil2(g, m, physL, physR, Xm, Xc, Xn, XcE, cur.kx, cur.kxe, kyc, cur.ky, wa, wb);
      const double2 rn = make_double2(cur.a.x - wa, cur.a.y - wb);
      const double2 un = make_double2(cur.b.x + Xc.x, cur.b.y + Xc.y);
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
      tl_reflect_edges(sout, g, m, j, oc, last ? Xc : sn);
      // tiled: the next inner step reads the neighbours' sd'; after the last step the next
      // outer iteration's matvec reads their r (p = r + beta p is recomputed at the neighbours)
      if (tiled) tl_push_edges(last ? P.push_r : push, g, m, j, last ? rn : sn);
      Xm = Xc; Xc = Xn; XcE = cur.xe; kyc = cur.ky;
    }
    tl_cp_wait<0>();
  }
  if (tl_kernel_tail(acc, last, st, P.partials, P.cd, sm, TL_T_RR)) {
    if (last) {
      st->red_rr_local = acc[0];      // PPCG.jl:88
      if ((P.single || tiled) && !tl_is_deferred(P.cd)) st->red_rr = acc[0];
      st->iter = it + 1;
    }
    st->inner_pp = pp + 1;
  }
}

// PPCG inner step, matrix-powers variant for tiles (algorithm, buffer rotation and citations:
// PpcgDkParams in tl_kernels_fused.cuh).  Same marching scheme and the same per-cell expressions
// as k_ppcg_inner_ring; what differs is the window it covers (the tile extended towards its
// neighbour tiles by the steps left in the group), where sd and r come from / go to, and that
// only a group's last step pushes halos and takes part in the tile exchange.
template <int S, int MINB>
__global__ void __launch_bounds__(TL_FUSED_THREADS, MINB) k_ppcg_inner_dk(const PpcgDkParams P) {
  tl_pdl_entry();
  extern __shared__ __align__(128) unsigned char ring_raw[];
  __shared__ double sm[32];
  SolveState *st = P.st;
  tl_prof_entry(st);
  const TlScal sc = tl_entry_scalars(P.cd, st);
  const int it = st->iter;
  if (st->comm_error || tl_should_stop(it, sc.rr, st->cfg)) return;
  const int pp = st->inner_pp, n = st->inner_steps, k = P.k;
  const int grp = pp / k, q = pp - grp * k, G = (n + k - 1) / k;
  const int L = min(k, n - grp * k);                 // steps in this group
  const int e = L - 1 - q;                           // extension of this step's window
  const bool group_end = (q == L - 1);
  const bool last = (pp + 1 == n);
  const int Lnext = last ? 0 : min(k, n - (grp + 1) * k);
  const double alpha = P.alphas[pp], beta = P.betas[pp];
  const double *__restrict__ sin = (q == 0) ? ((grp & 1) ? P.sin1 : P.sin0) : (((q - 1) & 1) ? P.wb : P.wa);
  double *__restrict__ sout = group_end ? (((grp + 1) & 1) ? P.sin1 : P.sin0) : ((q & 1) ? P.wb : P.wa);
  const double *rin = (q == 0) ? (((G - grp) & 1) ? P.rin1 : P.rin0) : P.rw;
  double *rout = group_end ? (((G - grp - 1) & 1) ? P.rin1 : P.rin0) : P.rw;
  const Push8 &push_s = ((grp + 1) & 1) ? P.push_sin1 : P.push_sin0;          // halo targets of sout at a group end
  const Push8 &push_r = ((G - grp - 1) & 1) ? P.push_rin1 : P.push_rin0;      // ... of rout
  const double *__restrict__ kx = P.kx;
  const double *__restrict__ ky = P.ky;
  double *__restrict__ u = P.u;
  const Geo g = P.g;
  const int pitch = g.pitch;
  const bool physL = g.phys & TL_PHYS_LEFT, physR = g.phys & TL_PHYS_RIGHT;
  const bool physB = g.phys & TL_PHYS_BOTTOM, physT = g.phys & TL_PHYS_TOP;
  const int xlo = physL ? 0 : -e, xhi = g.nx + (physR ? 0 : e);
  const int ylo = physB ? 0 : -e, yhi = g.ny + (physT ? 0 : e);

  double acc[1] = {0.0};
  MarchCtx m;
  if (tl_march_setup_ext(g, P.t, P.ext, xlo, xhi, ylo, yhi, m)) {
    const double2 z2 = make_double2(0.0, 0.0);
    const bool ina = m.i0 >= 0 && m.i0 < g.nx, inb = m.i0 + 1 >= 0 && m.i0 + 1 < g.nx;   // interior cells
    MarchCtx ml = m;                      // for the ring: the pointwise fields are loaded as pairs
    ml.acta = m.acta || m.actb;
    MarchCtx mi = m;                      // for the halo write-through: interior cells only
    mi.acta = ina; mi.actb = ina && inb;
    RingMarch<S> rg;
    rg.init(ring_raw, m);
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
#pragma unroll
    for (int d = 0; d < S - 1; d++) {
      if (m.j0 + d < m.j1) rg.issue(g, ml, physT, m.j0 + d, d, sin, ky, kx, rin, u);
      tl_cp_commit();
    }
    for (int j = m.j0; j < m.j1; j++) {
      if (j + S - 1 < m.j1) rg.issue(g, ml, physT, j + S - 1, rg.fill, sin, ky, kx, rin, u);
      tl_cp_commit();
      tl_cp_wait<S - 1>();
      const RingRow cur = rg.take(ml, true, true);
      const double2 Xn = cur.x;
      double wa, wb;
      tl_stencil2(g, m, physL, physR, Xm, Xc, Xn, XcE, cur.kx, cur.kxe, kyc, cur.ky, wa, wb);
      const double2 rn = make_double2(cur.a.x - wa, cur.a.y - wb);
      const double2 un = make_double2(cur.b.x + Xc.x, cur.b.y + Xc.y);
      const double2 sn = make_double2(alpha * Xc.x + beta * rn.x, alpha * Xc.y + beta * rn.y);
      const long oc = (long)j * pitch + m.i0;
      const bool jin = j >= 0 && j < g.ny;
      if (m.acta && m.actb) {
        tl_st2(rout + oc, rn); tl_st2(sout + oc, sn);
      } else if (m.acta) {
        rout[oc] = rn.x; sout[oc] = sn.x;
      } else if (m.actb) {
        rout[oc + 1] = rn.y; sout[oc + 1] = sn.y;
      }
      if (jin) {                                    // u and the dot product live on the interior only
        if (ina && inb) tl_st2(u + oc, un);
        else if (ina) u[oc] = un.x;
        else if (inb) u[oc + 1] = un.y;
        if (last) {
          if (ina) acc[0] += rn.x * rn.x;
          if (inb) acc[0] += rn.y * rn.y;
        }
        // reflective sides: as k_ppcg_inner_ring (PPCG.jl:76 reflects the step's input)
        tl_reflect_edges(sout, g, mi, j, oc, last ? Xc : sn);
        if (group_end) {
          if (last) tl_push_deep(P.push_rin0, g, 1, m.i0, ina, inb, j, rn);        // the next matvec reads r one cell deep
          else {
            tl_push_deep(push_s, g, Lnext, m.i0, ina, inb, j, sn);
            if (Lnext > 1) tl_push_deep(push_r, g, Lnext - 1, m.i0, ina, inb, j, rn);
          }
        }
      }
      Xm = Xc; Xc = Xn; XcE = cur.xe; kyc = cur.ky;
    }
    tl_cp_wait<0>();
  }
  if (tl_kernel_tail(acc, last, st, P.partials, group_end ? P.cd : nullptr, sm, TL_T_RR)) {
    if (last) {
      st->red_rr_local = acc[0];      // PPCG.jl:88
      if (!tl_is_deferred(P.cd)) st->red_rr = acc[0];
      st->iter = it + 1;
    }
    st->inner_pp = pp + 1;
  }
}

// Jacobi iteration (algorithm and citations: JacobiParams in tl_kernels_fused.cuh).
template <int S, int MINB>
__global__ void __launch_bounds__(TL_FUSED_THREADS, MINB) k_jacobi_fused_ring(const JacobiParams P) {
  tl_pdl_entry();
  extern __shared__ __align__(128) unsigned char ring_raw[];
  __shared__ double sm[32];
  SolveState *st = P.st;
  tl_prof_entry(st);
  const TlScal sc = tl_entry_scalars(P.cd, st);
  const int it = st->iter;
  if (st->comm_error || tl_should_stop(it, sc.rr, st->cfg)) return;
  const double *__restrict__ uin = (it & 1) ? P.ub : P.ua;
  double *__restrict__ uout = (it & 1) ? P.ua : P.ub;
  const bool tiled = P.cd != nullptr;
  const Push &push = (it & 1) ? P.push_ua : P.push_ub;   // halo targets of uout
  const double *__restrict__ u0 = P.u0;
  const double *__restrict__ kx = P.kx;
  const double *__restrict__ ky = P.ky;
  const Geo g = P.g;
  const int pitch = g.pitch;
  const bool physL = g.phys & TL_PHYS_LEFT, physR = g.phys & TL_PHYS_RIGHT;
  const bool physB = g.phys & TL_PHYS_BOTTOM, physT = g.phys & TL_PHYS_TOP;

  double acc[1] = {0.0};
  MarchCtx m;
  if (tl_march_setup(g, P.t, m)) {
    const double2 z2 = make_double2(0.0, 0.0);
    RingMarch<S> rg;
    rg.init(ring_raw, m);
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
#pragma unroll
    for (int d = 0; d < S - 1; d++) {
      if (m.j0 + d < m.j1) rg.issue(g, m, physT, m.j0 + d, d, uin, ky, kx, u0, nullptr);
      tl_cp_commit();
    }
    for (int j = m.j0; j < m.j1; j++) {
      if (j + S - 1 < m.j1) rg.issue(g, m, physT, j + S - 1, rg.fill, uin, ky, kx, u0, nullptr);
      tl_cp_commit();
      tl_cp_wait<S - 1>();
      const RingRow cur = rg.take(m, true, false);
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
      double2 un;
      un.x = tl_jacobi_cell(cur.a.x, cur.kx.x, cur.kx.y, kyc.x, cur.ky.x, La, Ra, Xm.x, Xn.x);
      un.y = tl_jacobi_cell(cur.a.y, cur.kx.y, kxr, kyc.y, cur.ky.y, Lb, Rb, Xm.y, Xn.y);
      const long oc = (long)j * pitch + m.i0;
      if (m.actb) {
        tl_st2(uout + oc, un);
        acc[0] += fabs(un.x - Xc.x);
        acc[0] += fabs(un.y - Xc.y);
      } else if (m.acta) {
        uout[oc] = un.x;
        acc[0] += fabs(un.x - Xc.x);
      }
      tl_reflect_edges(uout, g, m, j, oc, un);             // haloupdate!(.., [:u]) Jacobi.jl:23
      if (tiled) tl_push_edges(push, g, m, j, un);
      Xm = Xc; Xc = Xn; XcE = cur.xe; kyc = cur.ky;
    }
    tl_cp_wait<0>();
  }
  if (tl_kernel_tail(acc, true, st, P.partials, P.cd, sm, TL_T_RR)) {
    st->red_rr_local = acc[0];
    if (!tl_is_deferred(P.cd)) st->red_rr = acc[0];
    st->iter = it + 1;
  }
}

// Jacobi.jl:16-21: on every 50th iteration  r = u0 - A u ; error = sum(r.r).  Runs after the
// iteration kernel (whose tail exchange completed the halo pushes of u); idempotent, so the
// launched-ahead copies after the stop rule fired leave the state unchanged.
__global__ void __launch_bounds__(TL_BASIC_THREADS) k_jacobi_resid(const JacobiParams P) {
  tl_pdl_entry();
  __shared__ double sm[32];
  SolveState *st = P.st;
  tl_prof_entry(st);
  (void)tl_entry_scalars(P.cd, st);        // collect what the last iteration kernel posted (split exchange)
  const int it = st->iter;
  if (st->comm_error) return;
  if (!P.force_resid && (it == 0 || it % 50 != 0)) return;
  const double *__restrict__ u = (it & 1) ? P.ub : P.ua;
  const Geo g = P.g;
  double acc[1] = {0.0};
  TL_RECT_LOOP(0, g.nx, 0, g.ny) {
    const long o = (long)j * g.pitch + i;
    const double rv = P.u0[o] - tl_smvp(u, P.kx, P.ky, o, g.pitch);
    P.r[o] = rv;
    acc[0] += rv * rv;
  }
  if (tl_kernel_tail(acc, true, st, P.partials, P.cd, sm, P.force_resid ? TL_T_NONE : TL_T_RR)) {
    st->red_rr_local = acc[0];
    if (!P.force_resid && !tl_is_deferred(P.cd)) st->red_rr = acc[0];
  }
}

