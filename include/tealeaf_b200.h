/*
 * tealeaf_b200.h -- C-ABI of libtealeaf_b200.so, the B200-native (sm_100a) drop-in for the
 * implicit heat-conduction solve of Laura7089/TeaLeaf.jl (CG / Chebyshev / PPCG on the
 * 5-point stencil).
 *
 * The reference has no FFI (it is pure Julia); the entry points below are what a
 * `ccall` binding placed under the reference's own function names would bind.  Each
 * declaration cites the reference function it replaces (file:line under the reference
 * tree).  julia/TeaLeafB200.jl holds that binding; INTEGRATION.md shows how it is wired
 * into src/TeaLeaf.jl.
 *
 * Conventions
 *  - plain C: pointers, ints, doubles; no C++/torch types cross the boundary.
 *  - every function returns TL_OK (0) or a negative tl_status; the message is available
 *    from tl_last_error(ctx) until the next call on that context.  Nothing throws.
 *  - the library owns all device memory.  Host pointers are borrowed for the duration of
 *    the call only.  Out-scalars are caller-owned and are valid when the call returns
 *    (every call is synchronous with respect to its outputs).
 *  - host field layout is the reference's: Julia column-major (x, y) Float64 including
 *    halos, x = xcells + 2*halo_depth contiguous (src/chunk.jl:68-70); element [kk,jj]
 *    (1-based) is host[(kk-1) + (jj-1)*ld].
 *  - one caller thread per context (the reference is single-threaded).
 *  - tl_create: one context == one GPU (the 1x1 case).  Several GPUs: either ONE context for the whole
 *    mesh in one process (tl_create_multi -- what a host with a single Chunk like the reference's
 *    needs), or one context per tile with one process per GPU (tl_create_tile + tl_comm_*).
 */
#ifndef TEALEAF_B200_H
#define TEALEAF_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define TL_ABI_VERSION 1

typedef struct tl_ctx tl_ctx;

typedef enum tl_status {
  TL_OK = 0,
  TL_ERR_ARG = -1,        /* invalid argument (e.g. coefficient not 1|2, src/solvers/CG.jl:48-50) */
  TL_ERR_CUDA = -2,       /* a CUDA runtime call failed */
  TL_ERR_NO_DEVICE = -3,  /* no usable sm_100 device: there is NO CPU fallback */
  TL_ERR_EIGEN = -4,      /* eigenvalue estimate failed / negative (src/kernels.jl:41-43, :80) */
  TL_ERR_COMM = -5,       /* NCCL / IPC failure */
  TL_ERR_STATE = -6       /* call sequence error (e.g. comm not connected) */
} tl_status;

/* Chunk fields, src/chunk.jl:25-38.  (density0, mi are never read on the path.) */
typedef enum tl_field {
  TL_DENSITY = 0, TL_ENERGY0 = 1, TL_ENERGY = 2, TL_U = 3, TL_U0 = 4, TL_P = 5,
  TL_R = 6, TL_W = 7, TL_KX = 8, TL_KY = 9, TL_SD = 10, TL_NUM_FIELDS = 11
} tl_field;
#define TL_MASK(f) (1u << (f))

#define TL_CONDUCTIVITY 1        /* src/settings.jl:13 */
#define TL_RECIP_CONDUCTIVITY 2  /* src/settings.jl:14 */

/* Result of a whole solve (what the reference logs: CG.jl:27, Cheby.jl:60, PPCG.jl:54). */
typedef struct tl_solve_info {
  int iters;         /* outer iterations executed (`tt` at loop exit) */
  int cg_iters;      /* CG iterations among them (all of them for CG; the presteps otherwise) */
  int cheby_iters;   /* Chebyshev iterations (Cheby) or PPCG outer iterations (PPCG) */
  int est_iters;     /* Cheby.calciter estimate (Cheby.jl:109-118) */
  int inner_total;   /* PPCG inner steps executed */
  int halo_depth_k;  /* PPCG: exchange depth the inner steps used (1 on a single tile) */
  double error;      /* final `error` (rr) */
  double eigmin, eigmax;
  double solve_ms;   /* device time of the solve, CUDA events on the solve stream */
  long long kernel_launches;  /* kernels launched by this solve */
} tl_solve_info;

/* ---- context ------------------------------------------------------------------ */

/* Chunk(settings), src/chunk.jl:68-89: allocates every field for an xcells x ycells tile
 * with `halo_depth` halo cells per side on CUDA device `device`. */
int tl_create(tl_ctx **out, int xcells, int ycells, int halo_depth, int max_iters, int device);
/* Same, for tile `rank` (= cx + cy*px) of a px x py decomposition; sides that touch
 * another tile are exchanged, the others are reflective (src/kernels.jl:191-210). */
int tl_create_tile(tl_ctx **out, int xcells, int ycells, int halo_depth, int max_iters, int device,
                   int rank, int px, int py);
/* ONE context spanning `ngpus` GPUs of THIS process: the reference is one process with one Chunk (run.jl:45-47,
 * src/TeaLeaf.jl:35-44), so its host gets one handle for the GLOBAL xcells x ycells mesh and every entry point of this
 * header works on it unchanged -- tl_set_field / tl_get_field scatter / gather the reference's global (x, y) matrices
 * over the px x py tiles (halos included), tl_paint_states paints every tile at its offset, every kernel / solve call
 * runs on all tiles at once (they exchange halos and dot products over NVLink inside the kernels) and returns the
 * all-tiles scalars.  devices: `ngpus` CUDA device indices (NULL: 0 .. ngpus-1; repeating an index puts several tiles
 * on one GPU -- a test mode that needs CUDA_MODULE_LOADING=EAGER); px, py: the decomposition (<= 0: 1x2, 2x2, 2x4, ...:
 * y, the strided dimension, is split first).  The caller stays single-threaded; the library runs one service thread
 * per tile.  tl_comm_* must not be called on such a context. */
int tl_create_multi(tl_ctx **out, int xcells, int ycells, int halo_depth, int max_iters, int ngpus,
                    const int *devices, int px, int py);
void tl_destroy(tl_ctx *ctx);
/* ctx == NULL: why the last tl_create / tl_create_tile / tl_create_multi of this process failed */
const char *tl_last_error(const tl_ctx *ctx);
int tl_abi_version(void);
/* Tuning / A-B knobs.  None of them changes per-cell arithmetic; those that change how the tile is cut into
 * warp tasks (chunk_rows, balanced_tiling, *_pair) change the order dot products are added up in
 * (tests/test_gpu_parity.py states what stays bit-identical).  Defaults are
 * the measured best.  The environment variable TEALEAF_B200_OPTS
 * ("name=value,name=value") applies options to every context the process creates.
 *   chunk_rows, pw_chunk_rows   rows per warp task of the stencil / pointwise kernels (-1 = auto: default_chunk_rows = 12 / 16 on
 *                               meshes that fill the machine)
 *   blocks_per_sm, pw_blocks_per_sm   CTAs per SM the tilings of small meshes are sized for (2 / 4)
 *   ring_stages                 cp.async ring depth of the stencil kernels: -1 auto, 3, 4, 6
 *   graph_iters, use_graph      iterations per CUDA graph launch (8) / plain launches instead
 *   b_reverse                   kernel B walks the tile top-down (1)
 *   b_ring                      kernel B flavour: 0 register batches (default), 6 / 8 cp.async ring
 *   a_tma                       kernel A's row ring filled by TMA (cp.async.bulk.tensor + mbarriers): 0 off (default: measured,
 *                               not faster), 3 / 4 ring slots
 *   cg_persist                  1: the CG loop of a single tile as ONE persistent cooperative kernel
 *   cg_lazy_u                   CG loop: u advanced every second iteration with both pending updates, in the same order as one
 *                               update per iteration (bit-identical; 84 instead of 88 bytes per cell-iteration; default 1; needs an
 *                               even graph_iters and the cp.async flavour of kernel A, else one update per iteration)
 *   cg_lazy_heavy_ctas          the u-updating launch of that loop at ring depth 3: 2 (default) or 3 CTAs per SM
 *   cheby_pair, ppcg_pair       two Chebyshev iterations / PPCG inner steps per pass (temporal blocking; default 1)
 *   pair_tiled                  the pair kernels also on tiles: depth-2 halos, one exchange per two iterations (default 1)
 *   pair_rows                   rows per warp task of the pair kernels (32)
 *   balanced_tiling             mid-size tiles: chunk length that puts exactly two CTAs on every SM (default 1)
 *   hint_keep, hint_stream, l2_persist_mb, l2_hit_scale, l2_persist_field    L2 policy experiments
 *   use_pdl                     programmatic dependent launch between the loop kernels, released before the kernel tails
 *   comm_fused                  tiles: 1 halo pushes + mailbox sums inside the kernels (default), 0 NCCL + pull kernels
 *   xchg_deferred               tiles: split exchange -- kernels post their packets in the tail, the next kernel collects them at
 *                               its entry (bit-identical; measured slower than the blocking exchange: default 0)
 *   pair_stages                 cp.async ring depth of the pair kernels: 4 (default) or 5
 *   ppcg_halo_depth             tiles, one kernel per inner step: exchange every k steps (0 = automatic: the pair kernels,
 *                               i.e. every 2 steps, or every halo_depth steps when they are off)
 *   prof                        kernel-boundary micro-profile: globaltimer stamps in the kernel tails; read the averages per
 *                               kernel (microseconds) with tl_get_option: prof_kernels, prof_body_us (entry -> last block in
 *                               the tail), prof_sum_us (ticket + partial sums), prof_fence_us (fence.sys), prof_xchg_us (tile
 *                               exchange), prof_gap_us (previous tail -> next entry; kernels without a tail count as gap)
 * Unknown names return TL_ERR_ARG. */
int tl_set_option(tl_ctx *ctx, const char *name, double value);
/* Read-back of any option above, and of derived quantities: ring_stages_effective (the ring depth in use),
 * rows_per_chunk / pw_rows_per_chunk / pair_rows_per_chunk and fused_grid / pw_grid / pair_grid (how the tile
 * is cut into warp tasks and CTAs), max_grid (the bound on a grid: the partials array), num_sms, default_chunk_rows, cg_u_mode (1: u advanced every
 * iteration, 2: every second one -- what cg_lazy_u resolves to), cg_a_blocks_per_sm (resident CTAs of kernel A), last_cg_phase_ms
 * (device time of the CG phase -- preamble, CG presteps, flush -- of the last Chebyshev / PPCG solve), prof_* (above).
 * On a tl_create_multi context: the maximum over the tiles. */
int tl_get_option(tl_ctx *ctx, const char *name, double *value);

/* ---- multi-GPU wiring (no counterpart in the reference: it has a single Chunk) ---- */
int tl_comm_blob_size(void);                       /* bytes of one exported blob */
int tl_comm_export(tl_ctx *ctx, void *blob);       /* CUDA-IPC handle + tile descriptor */
int tl_comm_unique_id(void *id128);                /* ncclGetUniqueId (rank 0), 128 bytes */
/* all_blobs: px*py blobs in rank order; id128: the broadcast NCCL id. */
int tl_comm_connect(tl_ctx *ctx, const void *all_blobs, const void *id128);

/* ---- field transfer (getfield/setfield of Chunk, src/chunk.jl:25-38) ---------- */
int tl_set_field(tl_ctx *ctx, int field, const double *host, long ld);
int tl_get_field(tl_ctx *ctx, int field, double *host, long ld);
/* `chunk.dst .= chunk.src` on the whole array, device to device (src/TeaLeaf.jl:41) */
int tl_copy_field(tl_ctx *ctx, int dst_field, int src_field);

/* One `state` line of tea.in, src/settings.jl:20-29 (bounds already nudged by +-dx/100, :159-162). */
#define TL_GEOM_RECTANGULAR 0
#define TL_GEOM_CIRCULAR 1
#define TL_GEOM_POINT 2
typedef struct tl_state {
  double density, energy, xmin, ymin, xmax, ymax, radius;
  int geometry;   /* TL_GEOM_* */
  int reserved;
} tl_state;
/* setchunkstate!, src/chunk.jl:122-151, on the device: paints density, energy0 and u of this
 * tile (halo cells included) from the state list; vertex coordinates as Chunk(settings) builds
 * them (src/chunk.jl:76-77): vertexx[k] = xmin + dx*(k - 1 - halo_depth + x0), 1-based k, where
 * (x0, y0) is the tile's offset in the global mesh (0, 0 for a single chunk). */
int tl_paint_states(tl_ctx *ctx, int nstates, const tl_state *states, double xmin, double ymin,
                    double dx, double dy, int x0, int y0);

/* ---- kernels, one per reference function -------------------------------------- */
/* haloupdate!/updateface!, src/kernels.jl:146-159, :191-210 (depth-`depth` faces of every
 * field in field_mask; the sticky `toexchange` set stays on the Julia side). */
int tl_halo_update(tl_ctx *ctx, unsigned field_mask, int depth);
/* CG.init!, src/solvers/CG.jl:47-79 -> rro */
int tl_cg_init(tl_ctx *ctx, int coefficient, double rx, double ry, double *rro);
/* CG.w!, src/solvers/CG.jl:82-90 -> pw */
int tl_cg_calc_w(tl_ctx *ctx, double *pw);
/* CG.ur!, src/solvers/CG.jl:93-98 -> rrn */
int tl_cg_calc_ur(tl_ctx *ctx, double alpha, double *rrn);
/* CG.p!, src/solvers/CG.jl:101-104 */
int tl_cg_calc_p(tl_ctx *ctx, double beta);
/* copyu!, src/kernels.jl:217-220 */
int tl_copy_u(tl_ctx *ctx);
/* residual!, src/kernels.jl:227-232 */
int tl_calc_residual(tl_ctx *ctx);
/* finalise!, src/kernels.jl:239-242 */
int tl_finalise(tl_ctx *ctx);
/* solvefinished!, src/kernels.jl:166-170 (residual if check_result, finalise, halo energy) */
int tl_solve_finished(tl_ctx *ctx, int check_result);
/* sum(x->x^2, field[halo]) as used at Cheby.jl:68, :103 and PPCG.jl:88 */
int tl_norm2(tl_ctx *ctx, int field, double *out);
/* field part of Cheby.init!, src/solvers/Cheby.jl:64-81 (theta from coef!) -> bb */
int tl_cheby_init(tl_ctx *ctx, double theta, double *bb);
/* Cheby.mainstep!, src/solvers/Cheby.jl:84-106; alpha/beta = chebyα/β[chebyiters+1];
 * *error is updated only when calc_2norm != 0. */
int tl_cheby_iterate(tl_ctx *ctx, double alpha, double beta, int calc_2norm, double *error);
/* PPCG.init!(chunk, hd), src/solvers/PPCG.jl:97-100: sd = r / theta */
int tl_ppcg_init_sd(tl_ctx *ctx, double theta);
/* the inner loop of PPCG.mainstep!, src/solvers/PPCG.jl:75-84: nsteps x
 * { halo(sd); r -= A sd; u += sd; sd = alphas[pp] sd + betas[pp] r } (two-phase). */
int tl_ppcg_inner(tl_ctx *ctx, const double *alphas, const double *betas, int nsteps);
/* Jacobi.init!, src/solvers/Jacobi.jl:33-60 */
int tl_jacobi_init(tl_ctx *ctx, int coefficient, double rx, double ry);
/* Jacobi.iterate!, src/solvers/Jacobi.jl:62-82 -> sum(|u - r|) */
int tl_jacobi_iterate(tl_ctx *ctx, double *error);
/* fieldsummary, src/kernels.jl:119-133: temp = sum(volume*density*u); vol, mass, ie are
 * the upstream companions named by the parity criterion. */
int tl_field_summary(tl_ctx *ctx, double cell_volume, double *vol, double *mass, double *ie, double *temp);

/* ---- whole-solve fast paths: `settings.solver.solve!(chunk, settings, rx, ry)`,
 *      src/TeaLeaf.jl:74.  Fused kernels, device-resident scalars, graph-launched. ---- */
/* CG.solve!, src/solvers/CG.jl:7-29.  cg_alphas/cg_betas (nullable) receive chunk.cgα/cgβ. */
int tl_cg_solve(tl_ctx *ctx, int coefficient, double rx, double ry, double eps, int max_iters,
                tl_solve_info *info, double *cg_alphas, double *cg_betas);
/* Cheby.solve!, src/solvers/Cheby.jl:10-61 */
int tl_cheby_solve(tl_ctx *ctx, int coefficient, double rx, double ry, double eps, int max_iters,
                   int presteps, double epslim, int errorswitch, tl_solve_info *info);
/* PPCG.solve!, src/solvers/PPCG.jl:9-55.  halo_depth_k (0..halo_depth; 0 = automatic, i.e. the
 * option "ppcg_halo_depth", which defaults to halo_depth) is the depth of the sd/r halo exchange
 * between tiles: the inner steps of PPCG.jl:75-84 run in groups of halo_depth_k with ONE exchange
 * per group (matrix-powers kernel: step q of a group also computes k-1-q cells into the
 * tile-internal halos, redundantly and bit-identically to the owning tile).  The result does not
 * depend on it bit for bit; a single chunk exchanges nothing and ignores it. */
int tl_ppcg_solve(tl_ctx *ctx, int coefficient, double rx, double ry, double eps, int max_iters,
                  int presteps, double epslim, int errorswitch, int inner_steps, int halo_depth_k,
                  tl_solve_info *info);

/* Jacobi.driver! -- the module's solve!, src/solvers/Jacobi.jl:7-31 */
int tl_jacobi_solve(tl_ctx *ctx, int coefficient, double rx, double ry, double eps, int max_iters,
                    tl_solve_info *info);

/* ---- measurement helpers (bench.py) -------------------------------------------- */
/* average device time (ms) of the named fused kernel over `reps` launches on the current
 * state, CUDA events on the launch stream; fields are left modified. */
int tl_time_kernel(tl_ctx *ctx, const char *kernel, int reps, double *avg_ms);
/* CUDA-event stopwatch on the context's stream (the stream every kernel is launched on). */
int tl_timer_start(tl_ctx *ctx);
int tl_timer_stop(tl_ctx *ctx, double *elapsed_ms);
/* kernels launched by this context so far (graph-launched kernels included) */
int tl_launch_count(tl_ctx *ctx, long long *count);

#ifdef __cplusplus
}
#endif
#endif /* TEALEAF_B200_H */
