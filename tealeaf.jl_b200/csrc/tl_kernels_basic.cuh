// tl_kernels_basic.cuh -- one CUDA kernel per reference function (the per-kernel drop-in
// layer behind tl_cg_calc_w & co).  These read and write exactly the cells the reference
// function does, halos included, so a host that drives the solve kernel by kernel (as
// src/solvers/*.jl do) gets the reference's state after every call.  The fused whole-solve
// kernels live in tl_kernels_fused.cuh.
#pragma once
#include "tl_device.cuh"

#define TL_BASIC_THREADS 256

// Iterate a 1-D grid over the rectangle [i_lo,i_hi) x [j_lo,j_hi): one block covers 256
// consecutive cells of one row; blocks stride over (row, x-block) pairs.
#define TL_RECT_LOOP(i_lo, i_hi, j_lo, j_hi)                                                        \
  const int _nbx = ((i_hi) - (i_lo) + TL_BASIC_THREADS - 1) / TL_BASIC_THREADS;                     \
  const long _nt = (long)_nbx * ((j_hi) - (j_lo));                                                  \
  for (long _t = blockIdx.x; _t < _nt; _t += gridDim.x)                                             \
    for (int j = (j_lo) + (int)(_t / _nbx), i = (i_lo) + (int)(_t % _nbx) * TL_BASIC_THREADS + threadIdx.x, \
             _once = 1;                                                                             \
         _once && i < (i_hi); _once = 0)

// src/kernels.jl:178-184 read as one expression, same operation order
// as the oracle: consum = (((1 + kx[k+1]) + kx[k]) + ky[j+1]) + ky[j].
__device__ __forceinline__ double tl_smvp(const double *__restrict__ a, const double *__restrict__ kx,
                                          const double *__restrict__ ky, long o, int pitch) {
  const double kxr = kx[o + 1], kxl = kx[o], kyu = ky[o + pitch], kyd = ky[o];
  const double consum = (((1.0 + kxr) + kxl) + kyu) + kyd;
  return consum * a[o] - (kxr * a[o + 1] + kxl * a[o - 1]) - (kyu * a[o + pitch] + kyd * a[o - pitch]);
}

// CG.jl:52-58: u = energy*density, p = r = 0 on the whole array; w = density^(+-1) on all
// but the outer ring.
__global__ void k_init_fields(Geo g, int coef, const double *__restrict__ energy, const double *__restrict__ density,
                              double *u, double *p, double *r, double *w) {
  TL_RECT_LOOP(-g.hd, g.nx + g.hd, -g.hd, g.ny + g.hd) {
    const long o = (long)j * g.pitch + i;
    const double d = density[o];
    u[o] = energy[o] * d;
    p[o] = 0.0;
    r[o] = 0.0;
    if (i > -g.hd && i < g.nx + g.hd - 1 && j > -g.hd && j < g.ny + g.hd - 1) w[o] = (coef == 1) ? d : 1.0 / d;
  }
}

// CG.jl:61-68: jj = hd+1:y-1, kk = hd+1:x-1 (1-based) <-> i in [0, nx+hd-1), j in [0, ny+hd-1)
__global__ void k_init_k(Geo g, double rx, double ry, const double *__restrict__ w, double *kx, double *ky) {
  TL_RECT_LOOP(0, g.nx + g.hd - 1, 0, g.ny + g.hd - 1) {
    const long o = (long)j * g.pitch + i;
    const double wc = w[o], wl = w[o - 1], wd = w[o - g.pitch];
    kx[o] = rx * (wl + wc) / (2.0 * wl * wc);
    ky[o] = ry * (wd + wc) / (2.0 * wd * wc);
  }
}

// CG.jl:72-77: w = A u; r = u - w; p = r; rro = sum r^2
__global__ void k_init_wrp(Geo g, const double *__restrict__ u, const double *__restrict__ kx,
                           const double *__restrict__ ky, double *w, double *r, double *p, double *partials,
                           unsigned *counter, double *out) {
  __shared__ double sm[32];
  double acc[1] = {0.0};
  TL_RECT_LOOP(0, g.nx, 0, g.ny) {
    const long o = (long)j * g.pitch + i;
    const double wv = tl_smvp(u, kx, ky, o, g.pitch);
    const double rv = u[o] - wv;
    w[o] = wv; r[o] = rv; p[o] = rv;
    acc[0] += rv * rv;
  }
  if (tl_grid_sum<1>(acc, partials, counter, sm) && threadIdx.x == 0) *out = acc[0];
}

// CG.jl:82-90
__global__ void k_calc_w(Geo g, const double *__restrict__ p, const double *__restrict__ kx,
                         const double *__restrict__ ky, double *w, double *partials, unsigned *counter, double *out) {
  __shared__ double sm[32];
  double acc[1] = {0.0};
  TL_RECT_LOOP(0, g.nx, 0, g.ny) {
    const long o = (long)j * g.pitch + i;
    const double wv = tl_smvp(p, kx, ky, o, g.pitch);
    w[o] = wv;
    acc[0] += wv * p[o];
  }
  if (tl_grid_sum<1>(acc, partials, counter, sm) && threadIdx.x == 0) *out = acc[0];
}

// CG.jl:93-98
__global__ void k_calc_ur(Geo g, double alpha, const double *__restrict__ p, const double *__restrict__ w,
                          double *u, double *r, double *partials, unsigned *counter, double *out) {
  __shared__ double sm[32];
  double acc[1] = {0.0};
  TL_RECT_LOOP(0, g.nx, 0, g.ny) {
    const long o = (long)j * g.pitch + i;
    u[o] += alpha * p[o];
    const double rv = r[o] - alpha * w[o];
    r[o] = rv;
    acc[0] += rv * rv;
  }
  if (tl_grid_sum<1>(acc, partials, counter, sm) && threadIdx.x == 0) *out = acc[0];
}

// CG.jl:101-104
__global__ void k_calc_p(Geo g, double beta, const double *__restrict__ r, double *p) {
  TL_RECT_LOOP(0, g.nx, 0, g.ny) {
    const long o = (long)j * g.pitch + i;
    p[o] = beta * p[o] + r[o];
  }
}

// kernels.jl:217-220 (and whole-buffer copies when full != 0)
__global__ void k_copy(Geo g, int full, const double *__restrict__ src, double *dst) {
  const int h = full ? g.hd : 0;
  TL_RECT_LOOP(-h, g.nx + h, -h, g.ny + h) {
    const long o = (long)j * g.pitch + i;
    dst[o] = src[o];
  }
}

// kernels.jl:227-232
__global__ void k_residual(Geo g, const double *__restrict__ u, const double *__restrict__ u0,
                           const double *__restrict__ kx, const double *__restrict__ ky, double *r) {
  TL_RECT_LOOP(0, g.nx, 0, g.ny) {
    const long o = (long)j * g.pitch + i;
    r[o] = u0[o] - tl_smvp(u, kx, ky, o, g.pitch);
  }
}

// kernels.jl:239-242
__global__ void k_finalise(Geo g, const double *__restrict__ u, const double *__restrict__ density, double *energy) {
  TL_RECT_LOOP(0, g.nx, 0, g.ny) {
    const long o = (long)j * g.pitch + i;
    energy[o] = u[o] / density[o];
  }
}

// sum(x->x^2, f[halo])  (Cheby.jl:68, :103; PPCG.jl:88)
__global__ void k_norm2(Geo g, const double *__restrict__ f, double *partials, unsigned *counter, double *out) {
  __shared__ double sm[32];
  double acc[1] = {0.0};
  TL_RECT_LOOP(0, g.nx, 0, g.ny) {
    const double v = f[(long)j * g.pitch + i];
    acc[0] += v * v;
  }
  if (tl_grid_sum<1>(acc, partials, counter, sm) && threadIdx.x == 0) *out = acc[0];
}

// Cheby.jl:71-75 (first != 0: p = r/theta) and Cheby.jl:92-98 (p = alpha p + beta r)
__global__ void k_cheby_wrp(Geo g, int first, double theta, double alpha, double beta, const double *__restrict__ u,
                            const double *__restrict__ u0, const double *__restrict__ kx,
                            const double *__restrict__ ky, double *w, double *r, double *p) {
  TL_RECT_LOOP(0, g.nx, 0, g.ny) {
    const long o = (long)j * g.pitch + i;
    const double wv = tl_smvp(u, kx, ky, o, g.pitch);
    const double rv = u0[o] - wv;
    w[o] = wv; r[o] = rv;
    p[o] = first ? rv / theta : alpha * p[o] + beta * rv;
  }
}

// Cheby.jl:76, :99-100 and PPCG.jl:80: dst += src on the interior
__global__ void k_add(Geo g, const double *__restrict__ src, double *dst) {
  TL_RECT_LOOP(0, g.nx, 0, g.ny) {
    const long o = (long)j * g.pitch + i;
    dst[o] += src[o];
  }
}

// PPCG.jl:97-100
__global__ void k_ppcg_init_sd(Geo g, double theta, const double *__restrict__ r, double *sd) {
  TL_RECT_LOOP(0, g.nx, 0, g.ny) {
    const long o = (long)j * g.pitch + i;
    sd[o] = r[o] / theta;
  }
}

// PPCG.jl:78-80, phase 1 of the two-phase inner step (Appendix A #17)
__global__ void k_ppcg_inner1(Geo g, const double *__restrict__ sd, const double *__restrict__ kx,
                              const double *__restrict__ ky, double *r, double *u) {
  TL_RECT_LOOP(0, g.nx, 0, g.ny) {
    const long o = (long)j * g.pitch + i;
    r[o] -= tl_smvp(sd, kx, ky, o, g.pitch);
    u[o] += sd[o];
  }
}
// PPCG.jl:81-82, phase 2
__global__ void k_ppcg_inner2(Geo g, double alpha, double beta, const double *__restrict__ r, double *sd) {
  TL_RECT_LOOP(0, g.nx, 0, g.ny) {
    const long o = (long)j * g.pitch + i;
    sd[o] = alpha * sd[o] + beta * r[o];
  }
}

// setchunkstate!, chunk.jl:122-151 ([k,j] indexing: Appendix A #3; celly from vertexy: A #4).
// Every state in turn, the first included, later states overwrite earlier ones; then
// u = energy0 * density on all but the outer ring (halo(ch, 1), chunk.jl:149-150).
#define TL_MAX_STATES 32
struct PaintState { double density, energy, xmin, ymin, xmax, ymax, radius; int geometry, pad; };
struct PaintParams {
  int n;
  double xmin, ymin, dx, dy;
  int x0, y0;
  PaintState s[TL_MAX_STATES];
};
__global__ void k_paint_states(Geo g, const PaintParams P, double *density, double *energy0, double *u) {
  TL_RECT_LOOP(-g.hd, g.nx + g.hd, -g.hd, g.ny + g.hd) {
    // chunk.jl:76-77: vertex k (0-based over the padded array) = min + d * (k - hd + offset)
    const double vx0 = P.xmin + P.dx * (double)(i + P.x0), vx1 = P.xmin + P.dx * (double)(i + 1 + P.x0);
    const double vy0 = P.ymin + P.dy * (double)(j + P.y0), vy1 = P.ymin + P.dy * (double)(j + 1 + P.y0);
    const double cx = 0.5 * (vx0 + vx1), cy = 0.5 * (vy0 + vy1);
    double d = P.s[0].density, e = P.s[0].energy;
    for (int q = 0; q < P.n; q++) {
      const PaintState &s = P.s[q];
      bool apply;
      if (s.geometry == 0) apply = vx1 >= s.xmin && vx0 < s.xmax && vy1 >= s.ymin && vy0 < s.ymax;
      else if (s.geometry == 1) apply = (cx - s.xmin) * (cx - s.xmin) + (cy - s.ymin) * (cy - s.ymin) <= s.radius * s.radius;
      else apply = vx0 == s.xmin && vy0 == s.ymin;
      if (apply) { d = s.density; e = s.energy; }
    }
    const long o = (long)j * g.pitch + i;
    density[o] = d;
    energy0[o] = e;
    const bool ring = i == -g.hd || i == g.nx + g.hd - 1 || j == -g.hd || j == g.ny + g.hd - 1;
    u[o] = ring ? 0.0 : e * d;
  }
}

// ---- Jacobi (src/solvers/Jacobi.jl) ----------------------------
// Jacobi.jl:39-41: u0 = u = energy .* density on the whole array
__global__ void k_jacobi_init_fields(Geo g, const double *__restrict__ energy, const double *__restrict__ density,
                                     double *u, double *u0) {
  TL_RECT_LOOP(-g.hd, g.nx + g.hd, -g.hd, g.ny + g.hd) {
    const long o = (long)j * g.pitch + i;
    const double t = energy[o] * density[o];
    u0[o] = t;
    u[o] = t;
  }
}
// Jacobi.jl:43-51: jj = hd+1:y-1, kk = hd+1:x-1; density^p, p = +-1
__global__ void k_jacobi_init_k(Geo g, int coef, double rx, double ry, const double *__restrict__ density, double *kx,
                                double *ky) {
  TL_RECT_LOOP(0, g.nx + g.hd - 1, 0, g.ny + g.hd - 1) {
    const long o = (long)j * g.pitch + i;
    double dc = density[o], dl = density[o - 1], dd = density[o - g.pitch];
    if (coef != 1) { dc = 1.0 / dc; dl = 1.0 / dl; dd = 1.0 / dd; }
    kx[o] = rx * (dl + dc) / (2.0 * dl * dc);
    ky[o] = ry * (dd + dc) / (2.0 * dd * dc);
  }
}
// the sweep of Jacobi.iterate!, Jacobi.jl:65-80 (r holds the previous u), -> sum |u - r|
__device__ __forceinline__ double tl_jacobi_cell(double u0, double kxl, double kxr, double kyd, double kyu, double xl,
                                                 double xr, double xd, double xu) {
  const double num = (((u0 + kxr * xr) + kxl * xl) + kyu * xu) + kyd * xd;
  const double den = (((1.0 + kxl) + kxr) + kyd) + kyu;
  return num / den;
}
__global__ void k_jacobi_sweep(Geo g, const double *__restrict__ u0, const double *__restrict__ kx,
                               const double *__restrict__ ky, const double *__restrict__ r, double *u, double *partials,
                               unsigned *counter, double *out) {
  __shared__ double sm[32];
  double acc[1] = {0.0};
  TL_RECT_LOOP(0, g.nx, 0, g.ny) {
    const long o = (long)j * g.pitch + i;
    const double un = tl_jacobi_cell(u0[o], kx[o], kx[o + 1], ky[o], ky[o + g.pitch], r[o - 1], r[o + 1], r[o - g.pitch],
                                     r[o + g.pitch]);
    u[o] = un;
    acc[0] += fabs(un - r[o]);
  }
  if (tl_grid_sum<1>(acc, partials, counter, sm) && threadIdx.x == 0) *out = acc[0];
}

// kernels.jl:191-210 with the 1-based reflection of Appendix A #2.  Only sides flagged
// physical in g.phys are reflected (tile-internal sides are filled by k_pull_halo).
// x faces run over interior rows, y faces over interior columns -- corners are not touched,
// exactly like the reference.
__global__ void k_halo_reflect(Geo g, int depth, double *f) {
  const int per_x = g.ny * depth, per_y = g.nx * depth;
  const int total = 2 * per_x + 2 * per_y;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    if (t < 2 * per_x) {
      const int side = t / per_x, q = t % per_x, j = q / depth, d = q % depth + 1;
      if (side == 0) { if (g.phys & TL_PHYS_LEFT)  f[(long)j * g.pitch - d] = f[(long)j * g.pitch + d - 1]; }
      else           { if (g.phys & TL_PHYS_RIGHT) f[(long)j * g.pitch + g.nx + d - 1] = f[(long)j * g.pitch + g.nx - d]; }
    } else {
      const int t2 = t - 2 * per_x, side = t2 / per_y, q = t2 % per_y, d = q / g.nx + 1, i = q % g.nx;
      if (side == 0) { if (g.phys & TL_PHYS_TOP)    f[(long)(g.ny + d - 1) * g.pitch + i] = f[(long)(g.ny - d) * g.pitch + i]; }
      else           { if (g.phys & TL_PHYS_BOTTOM) f[(long)(-d) * g.pitch + i] = f[(long)(d - 1) * g.pitch + i]; }
    }
  }
}

// Tile-internal faces: copy `depth` edge rows/columns of the neighbour tile's interior into
// this tile's halo, reading the neighbour's memory directly over NVLink (CUDA-IPC mapped peer
// pointers; nbr[s] == nullptr on physical sides).  nbr_* describe the neighbour tile.
struct PeerFace { const double *f0; int nx, ny, pitch; };
__global__ void k_pull_halo(Geo g, int depth, double *f, PeerFace left, PeerFace right, PeerFace bottom, PeerFace top) {
  const int per_x = g.ny * depth, per_y = g.nx * depth;
  const int total = 2 * per_x + 2 * per_y;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    if (t < 2 * per_x) {
      const int side = t / per_x, q = t % per_x, j = q / depth, d = q % depth + 1;
      if (side == 0) { if (left.f0)  f[(long)j * g.pitch - d] = __ldcv(&left.f0[(long)j * left.pitch + left.nx - d]); }
      else           { if (right.f0) f[(long)j * g.pitch + g.nx + d - 1] = __ldcv(&right.f0[(long)j * right.pitch + d - 1]); }
    } else {
      const int t2 = t - 2 * per_x, side = t2 / per_y, q = t2 % per_y, d = q / g.nx + 1, i = q % g.nx;
      if (side == 0) { if (top.f0)    f[(long)(g.ny + d - 1) * g.pitch + i] = __ldcv(&top.f0[(long)(d - 1) * top.pitch + i]); }
      else           { if (bottom.f0) f[(long)(-d) * g.pitch + i] = __ldcv(&bottom.f0[(long)(bottom.ny - d) * bottom.pitch + i]); }
    }
  }
}

// Depth-`depth` halo INCLUDING the corner blocks, in the classic two phases (a rendezvous of all
// tiles between them): phase 0 fills the left/right halo columns of the interior rows, phase 1 the
// bottom/top halo rows over the full width [-depth, nx+depth) -- the neighbour's own left/right
// halo columns (filled in its phase 0) carry the diagonal tiles' cells.  Used once per solve for
// kx, ky of the matrix-powers PPCG groups.
__global__ void k_pull_halo_wide(Geo g, int depth, int phase, double *f, PeerFace left, PeerFace right, PeerFace bottom,
                                 PeerFace top) {
  if (phase == 0) {
    const int total = 2 * g.ny * depth;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
      const int side = t / (g.ny * depth), q = t % (g.ny * depth), j = q / depth, d = q % depth + 1;
      if (side == 0) { if (left.f0)  f[(long)j * g.pitch - d] = __ldcv(&left.f0[(long)j * left.pitch + left.nx - d]); }
      else           { if (right.f0) f[(long)j * g.pitch + g.nx + d - 1] = __ldcv(&right.f0[(long)j * right.pitch + d - 1]); }
    }
  } else {
    const int wide = g.nx + 2 * depth, total = 2 * wide * depth;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
      const int side = t / (wide * depth), q = t % (wide * depth), d = q / wide + 1, i = q % wide - depth;
      if (side == 0) { if (top.f0)    f[(long)(g.ny + d - 1) * g.pitch + i] = __ldcv(&top.f0[(long)(d - 1) * top.pitch + i]); }
      else           { if (bottom.f0) f[(long)(-d) * g.pitch + i] = __ldcv(&bottom.f0[(long)(bottom.ny - d) * bottom.pitch + i]); }
    }
  }
}

// Phase 0 of k_pull_halo_wide extended upwards by one row on a PHYSICAL top: the neighbour's own halo row ny
// (ky there is what the stencil of row ny-1 reads) -- needed by kernels that recompute cells in the
// tile-internal halo columns next to a physical top (k_cheby_pair_tiled_ring; found by
// tests/emulation/emulate_pair_tiled.py).  EXPERIMENTAL, with that kernel.
__global__ void k_pull_halo_cols_top(Geo g, int depth, double *f, PeerFace left, PeerFace right) {
  if (!(g.phys & TL_PHYS_TOP)) return;
  const int total = 2 * depth;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const int side = t / depth, d = t % depth + 1, j = g.ny;
    if (side == 0) { if (left.f0)  f[(long)j * g.pitch - d] = __ldcv(&left.f0[(long)j * left.pitch + left.nx - d]); }
    else           { if (right.f0) f[(long)j * g.pitch + g.nx + d - 1] = __ldcv(&right.f0[(long)j * right.pitch + d - 1]); }
  }
}

// kernels.jl:119-133 (+ upstream vol/mass/ie): cell_mass = volume*density;
// vol += volume; mass += cell_mass; ie += cell_mass*energy0; temp += cell_mass*u.
__global__ void k_field_summary(Geo g, double cell_volume, const double *__restrict__ density,
                                const double *__restrict__ energy0, const double *__restrict__ u, double *partials,
                                unsigned *counter, double *out4) {
  __shared__ double sm[32];
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  TL_RECT_LOOP(0, g.nx, 0, g.ny) {
    const long o = (long)j * g.pitch + i;
    const double cm = cell_volume * density[o];
    acc[0] += cell_volume;
    acc[1] += cm;
    acc[2] += cm * energy0[o];
    acc[3] += cm * u[o];
  }
  if (tl_grid_sum<4>(acc, partials, counter, sm) && threadIdx.x == 0) {
    out4[0] = acc[0]; out4[1] = acc[1]; out4[2] = acc[2]; out4[3] = acc[3];
  }
}
