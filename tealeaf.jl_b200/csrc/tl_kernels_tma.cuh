// tl_kernels_tma.cuh -- CG kernel A with the row ring filled by the Tensor Memory Accelerator (option a_tma).
//
// Same work decomposition, same per-cell arithmetic and the same summation order as k_cg_fused_w_ring
// (tl_kernels_ring.cuh): one warp marches up a 64-cell strip, rows j-1, j, j+1 of the stencil operand are carried in
// registers.  What changes is how the next S-1 rows reach shared memory: instead of five 16-byte cp.async per LANE and
// row (plus the 8-byte edge copies of lanes 0 / 31), ONE elected lane issues five cp.async.bulk.tensor.2d copies per row,
// each a box of 68 doubles x 1 row (the strip's 64 columns plus two on each side, so the neighbour strips' edge cells
// arrive in the same box and the separate edge loads disappear).  Completion is tracked by one mbarrier per (warp, ring
// slot) with expect_tx; the lanes wait on its phase parity.  No per-lane address arithmetic, no commit/wait groups;
// boxes that stick out of the padded row are zero-filled by the hardware.
//
// Shared memory: a box is 544 bytes, bulk-tensor destinations must be 128-byte aligned -> 640 bytes per field,
// 3200 bytes per row slot, S = 4 slots per warp = 100 KB per CTA, two CTAs per SM.
#pragma once
#include <cuda.h>   // CUtensorMap (types only: the encoder is fetched with cudaGetDriverEntryPoint)
#include "tl_kernels_ring.cuh"

#define TL_TMA_BOX 68                                   // doubles per box row: columns s*64-2 .. s*64+65
#define TL_TMA_FIELD_BYTES 640                          // 544 rounded up to the 128-byte destination alignment
#define TL_TMA_STAGE_BYTES (5 * TL_TMA_FIELD_BYTES)
enum { TMA_R = 0, TMA_P0, TMA_P1, TMA_KY, TMA_KX, TMA_U, TMA_NMAPS };

struct CgATmaParams {
  CgAParams a;
  alignas(64) CUtensorMap maps[TMA_NMAPS];   // one 2-D map (padded row x rows) per buffer of the slab
};

__device__ __forceinline__ void tl_mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void tl_mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tl_mbar_wait(unsigned bar, unsigned parity) {
  unsigned done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tl_tma_load_2d(unsigned dst, const CUtensorMap *map, int c0, int c1, unsigned bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}

template <bool UPDATE_U, int S, int MINB>
__global__ void __launch_bounds__(TL_FUSED_THREADS, MINB) k_cg_fused_w_tma(const __grid_constant__ CgATmaParams PT) {
  const CgAParams &P = PT.a;
  tl_pdl_entry();
  extern __shared__ __align__(128) unsigned char ring_raw[];
  __shared__ double sm[32];
  __shared__ __align__(8) unsigned long long bars[TL_FUSED_THREADS / 32][S];
  SolveState *st = P.st;
  tl_prof_entry(st);
  const TlScal sc = tl_entry_scalars(P.cd, st);
  const int it = st->iter;
  const double rr_cur = sc.rr;
  if (st->comm_error || tl_should_stop(it, rr_cur, st->cfg)) return;
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
  const CUtensorMap *map_p = &PT.maps[(it & 1) ? TMA_P1 : TMA_P0];
  const double *__restrict__ r = P.r;
  const double *__restrict__ ky = P.ky;
  double *__restrict__ u = P.u;
  double *__restrict__ w = P.w;
  const Geo g = P.g;
  const int pitch = g.pitch;
  const bool physL = g.phys & TL_PHYS_LEFT, physR = g.phys & TL_PHYS_RIGHT;
  const bool physB = g.phys & TL_PHYS_BOTTOM, physT = g.phys & TL_PHYS_TOP;
  const bool tiled = P.cd != nullptr;
  const Push &push = (it & 1) ? P.push_p0 : P.push_p1;

  double acc[1] = {0.0};
  MarchCtx m;
  if (tl_march_setup_blk(g, P.t, m, blockIdx.x, 0)) {
    const double2 z2 = make_double2(0.0, 0.0);
    auto comb = [&](double rv, double pv) { return first ? pv : beta * pv + rv; };
    auto comb2 = [&](double2 rv, double2 pv) { return make_double2(comb(rv.x, pv.x), comb(rv.y, pv.y)); };
    const int warp = threadIdx.x >> 5;
    const unsigned ring = tl_smem_u32(ring_raw) + warp * (S * TL_TMA_STAGE_BYTES);
    const unsigned bar0 = tl_smem_u32(&bars[warp][0]);
    if (m.lane == 0) {
#pragma unroll
      for (int q = 0; q < S; q++) tl_mbar_init(bar0 + 8 * q, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const int strip0 = m.i0 - 2 * m.lane;                 // first column of the strip
    const int c0 = TL_XPAD + strip0 - 2;                   // box origin in the padded row
    const unsigned lane_off = 16 + 16 * m.lane;            // this lane's pair inside a box
    const unsigned edge_off = (m.lane == 0) ? 8 : 8 * 66;  // column strip0-1 / strip0+64
    const unsigned stage_bytes = (UPDATE_U ? 5u : 4u) * (TL_TMA_BOX * 8u);
    auto issue = [&](int j, int slot) {                    // one elected lane; rows in tensor coordinates are j + hd
      const unsigned base = ring + slot * TL_TMA_STAGE_BYTES, bar = bar0 + 8 * slot;
      const int jn = (j + 1 >= g.ny && physT) ? g.ny - 1 : j + 1;
      tl_mbar_expect_tx(bar, stage_bytes);
      tl_tma_load_2d(base + 0 * TL_TMA_FIELD_BYTES, &PT.maps[TMA_R], c0, jn + g.hd, bar);
      tl_tma_load_2d(base + 1 * TL_TMA_FIELD_BYTES, map_p, c0, jn + g.hd, bar);
      tl_tma_load_2d(base + 2 * TL_TMA_FIELD_BYTES, &PT.maps[TMA_KY], c0, j + 1 + g.hd, bar);
      tl_tma_load_2d(base + 3 * TL_TMA_FIELD_BYTES, &PT.maps[TMA_KX], c0, j + g.hd, bar);
      if (UPDATE_U) tl_tma_load_2d(base + 4 * TL_TMA_FIELD_BYTES, &PT.maps[TMA_U], c0, j + g.hd, bar);
    };
    // prologue: rows j0-1 (clamped on a physical bottom) and j0, plain loads
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
    if (m.lane == 0) {
#pragma unroll
      for (int d = 0; d < S - 1; d++)
        if (m.j0 + d < m.j1) issue(m.j0 + d, d);
    }
    int stage = 0, fill = S - 1;
    unsigned parity = 0;                                   // phase of slot `stage`: flips every trip round the ring
    for (int j = m.j0; j < m.j1; j++) {
      __syncwarp();                                        // every lane has consumed the slot that is refilled now
      if (m.lane == 0 && j + S - 1 < m.j1) issue(j + S - 1, fill);
      tl_mbar_wait(bar0 + 8 * stage, parity);
      const unsigned base = ring + stage * TL_TMA_STAGE_BYTES;
      const double2 c_r = m.ld_ok ? tl_lds2(base + 0 * TL_TMA_FIELD_BYTES + lane_off) : z2;
      const double2 c_p = m.ld_ok ? tl_lds2(base + 1 * TL_TMA_FIELD_BYTES + lane_off) : z2;
      const double2 c_ky = m.ld_ok ? tl_lds2(base + 2 * TL_TMA_FIELD_BYTES + lane_off) : z2;
      const double2 c_kx = m.ld_ok ? tl_lds2(base + 3 * TL_TMA_FIELD_BYTES + lane_off) : z2;
      const double2 c_u = (UPDATE_U && m.acta) ? tl_lds2(base + 4 * TL_TMA_FIELD_BYTES + lane_off) : z2;
      const double c_re = m.has_edge ? tl_lds1(base + 0 * TL_TMA_FIELD_BYTES + edge_off) : 0.0;
      const double c_pe = m.has_edge ? tl_lds1(base + 1 * TL_TMA_FIELD_BYTES + edge_off) : 0.0;
      const double c_kxe = (m.lane == 31 && m.has_edge) ? tl_lds1(base + 3 * TL_TMA_FIELD_BYTES + edge_off) : 0.0;
      stage = (stage + 1 == S) ? 0 : stage + 1;
      if (stage == 0) parity ^= 1u;
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
      if (UPDATE_U) un = make_double2(c_u.x + alpha_prev * pc.x, c_u.y + alpha_prev * pc.y);
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
      tl_reflect_edges(pout, g, m, j, oc, Xc);
      if (UPDATE_U) tl_reflect_edges(u, g, m, j, oc, un);
      if (tiled) tl_push_edges(push, g, m, j, Xc);
      Xm = Xc; Xc = Xn; XcE = XnE; pc = c_p; kyc = c_ky;
    }
  }
  if (tl_kernel_tail(acc, true, st, P.partials, P.cd, sm, TL_T_PW)) {
    st->red_pw_local = acc[0];
    if ((P.single || P.cd != nullptr) && !tl_is_deferred(P.cd)) st->red_pw = acc[0];
  }
}
