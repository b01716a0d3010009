/*
 * tea_c_host.c -- a host written in plain C99 against include/tealeaf_b200.h: the timestep loop
 * of the reference (initialiseapp! src/TeaLeaf.jl:35-44, diffuse! src/TeaLeaf.jl:62-83) with
 * every kernel call going through the C-ABI, exactly as the Julia ccall glue of
 * julia/TeaLeafB200.jl does.  It exists to show that the boundary is usable from C with nothing
 * but the header and the shared library (no Python, no torch): tests/test_c_host.py compiles it
 * with `gcc -std=c99 -pedantic -Wall -Wextra -Werror`, runs it on the GPU box and compares its
 * output with the Python mirror (bit for bit) and with the CPU oracle.
 *
 * The deck is the classic 5-state TeaLeaf benchmark deck on [0,10]^2.
 *
 *   tea_c_host [--gpus N] [--share-gpu] <x_cells> <y_cells> <end_step> <cg|cheby|ppcg|jacobi> [max_iters] [ppcg_inner_steps]
 *
 * --gpus N: ONE context for the whole mesh spread over N GPUs of this process (tl_create_multi): the loop below is
 * unchanged -- the reference has one process and one Chunk (run.jl:45-47).  --share-gpu puts all N tiles on device 0
 * (a test mode for single-GPU boxes; needs CUDA_MODULE_LOADING=EAGER in the environment).  With --gpus the final u
 * field is downloaded through the global tl_get_field and its interior sum is printed as "usum <s>".
 *
 * Output: one line per timestep  "step <tt> iters <n> error <rr>"  and a final
 * "summary vol <v> mass <m> ie <e> temp <t>"  (%.17g, i.e. round-trip exact).
 * Exit code: 0 ok, 3 no sm_100 device (TL_ERR_NO_DEVICE: there is no CPU fallback), 1 any other error.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "tealeaf_b200.h"

#define CHECK(call)                                                                    \
  do {                                                                                 \
    int rc_ = (call);                                                                  \
    if (rc_ != TL_OK) {                                                                \
      fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, tl_last_error(ctx));        \
      tl_destroy(ctx);                                                                 \
      return 1;                                                                        \
    }                                                                                  \
  } while (0)

int main(int argc, char **argv) {
  int ngpus = 0, share_gpu = 0;
  while (argc > 1 && argv[1][0] == '-') {
    if (strcmp(argv[1], "--gpus") == 0 && argc > 2) { ngpus = atoi(argv[2]); argv += 2; argc -= 2; }
    else if (strcmp(argv[1], "--share-gpu") == 0) { share_gpu = 1; argv += 1; argc -= 1; }
    else break;
  }
  if (argc < 5) {
    fprintf(stderr, "usage: %s [--gpus N] [--share-gpu] x_cells y_cells end_step cg|cheby|ppcg|jacobi [max_iters] [ppcg_inner_steps]\n", argv[0]);
    return 2;
  }
  const int nx = atoi(argv[1]), ny = atoi(argv[2]), end_step = atoi(argv[3]);
  const char *solver = argv[4];
  const int max_iters = argc > 5 ? atoi(argv[5]) : 10000;     /* settings.jl:43 */
  const int inner_steps = argc > 6 ? atoi(argv[6]) : 10;      /* settings.jl:45 */
  const int halo_depth = 2;                                   /* settings.jl:47 */
  const double xmin = 0.0, ymin = 0.0, xmax = 10.0, ymax = 10.0;
  const double dt = 0.004, eps = 1.0e-15, epslim = 1.0e-5;    /* deck; settings.jl:56-57 */
  const int presteps = 30, errorswitch = 0;                   /* settings.jl:42, :53 */
  const double dx = (xmax - xmin) / nx, dy = (ymax - ymin) / ny;   /* settings.jl:131-132 */

  /* the five `state` lines; bounds of states 2.. nudged as readstate does (settings.jl:158-162) */
  tl_state st[5];
  memset(st, 0, sizeof st);
  st[0].density = 100.0; st[0].energy = 0.0001;
  {
    const double raw[4][6] = {/* density energy xmin xmax ymin ymax */
                              {0.1, 25.0, 0.0, 1.0, 1.0, 2.0},
                              {0.1, 0.1, 1.0, 6.0, 1.0, 2.0},
                              {0.1, 0.1, 5.0, 6.0, 1.0, 8.0},
                              {0.1, 0.1, 5.0, 10.0, 7.0, 8.0}};
    int q;
    for (q = 0; q < 4; q++) {
      st[q + 1].density = raw[q][0];
      st[q + 1].energy = raw[q][1];
      st[q + 1].xmin = raw[q][2] + dx / 100;
      st[q + 1].xmax = raw[q][3] - dx / 100;
      st[q + 1].ymin = raw[q][4] + dy / 100;
      st[q + 1].ymax = raw[q][5] - dy / 100;
      st[q + 1].geometry = TL_GEOM_RECTANGULAR;
    }
  }

  tl_ctx *ctx = NULL;
  int rc;
  if (ngpus > 0) {                                                /* one Chunk over N GPUs */
    int devices[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    rc = tl_create_multi(&ctx, nx, ny, halo_depth, max_iters, ngpus, share_gpu ? devices : NULL, 0, 0);
  } else {
    rc = tl_create(&ctx, nx, ny, halo_depth, max_iters, 0);       /* Chunk(settings), chunk.jl:68-89 */
  }
  if (rc == TL_ERR_NO_DEVICE) {
    fprintf(stderr, "tl_create: no sm_100 device (TL_ERR_NO_DEVICE); libtealeaf_b200 has no CPU fallback\n");
    return 3;
  }
  if (rc != TL_OK) {
    fprintf(stderr, "tl_create failed (%d): %s\n", rc, tl_last_error(NULL));
    return 1;
  }
  if (tl_abi_version() != TL_ABI_VERSION) {
    fprintf(stderr, "ABI mismatch\n");
    tl_destroy(ctx);
    return 1;
  }

  /* initialiseapp!, TeaLeaf.jl:35-44 */
  CHECK(tl_paint_states(ctx, 5, st, xmin, ymin, dx, dy, 0, 0));                                   /* :37 */
  CHECK(tl_halo_update(ctx, TL_MASK(TL_DENSITY) | TL_MASK(TL_ENERGY0) | TL_MASK(TL_ENERGY), 1)); /* :39 */
  CHECK(tl_copy_field(ctx, TL_ENERGY, TL_ENERGY0));                                               /* :41 */

  /* diffuse!, TeaLeaf.jl:62-83 */
  {
    int tt;
    for (tt = 1; tt <= end_step; tt++) {
      const double rx = dt / (dx * dx), ry = dt / (dy * dy);                                      /* :69-70 */
      tl_solve_info info;
      CHECK(tl_halo_update(ctx, TL_MASK(TL_ENERGY) | TL_MASK(TL_DENSITY), 1));                    /* :71 */
      if (strcmp(solver, "cg") == 0)                                                              /* :74 */
        CHECK(tl_cg_solve(ctx, TL_CONDUCTIVITY, rx, ry, eps, max_iters, &info, NULL, NULL));
      else if (strcmp(solver, "cheby") == 0)
        CHECK(tl_cheby_solve(ctx, TL_CONDUCTIVITY, rx, ry, eps, max_iters, presteps, epslim, errorswitch, &info));
      else if (strcmp(solver, "ppcg") == 0)
        CHECK(tl_ppcg_solve(ctx, TL_CONDUCTIVITY, rx, ry, eps, max_iters, presteps, epslim, errorswitch, inner_steps, 0,
                            &info));
      else if (strcmp(solver, "jacobi") == 0)
        CHECK(tl_jacobi_solve(ctx, TL_CONDUCTIVITY, rx, ry, eps, max_iters, &info));
      else {
        fprintf(stderr, "unknown solver %s\n", solver);
        tl_destroy(ctx);
        return 2;
      }
      CHECK(tl_solve_finished(ctx, 0));                                                           /* :77 */
      printf("step %d iters %d error %.17g\n", tt, info.iters, info.error);
    }
  }
  {
    double vol, mass, ie, temp;
    CHECK(tl_field_summary(ctx, dx * dy, &vol, &mass, &ie, &temp));                               /* :82 */
    printf("summary vol %.17g mass %.17g ie %.17g temp %.17g\n", vol, mass, ie, temp);
  }
  if (ngpus > 0) {   /* the global gather: chunk.u as the reference holds it, (x, y) column-major with halos */
    const long x = nx + 2 * halo_depth, y = ny + 2 * halo_depth;
    double *u = (double *)malloc((size_t)(x * y) * sizeof(double));
    double usum = 0.0;
    long i, j;
    if (!u) { tl_destroy(ctx); return 1; }
    CHECK(tl_get_field(ctx, TL_U, u, x));
    for (j = halo_depth; j < y - halo_depth; j++)
      for (i = halo_depth; i < x - halo_depth; i++) usum += u[j * x + i];
    printf("usum %.17g\n", usum);
    free(u);
  }
  tl_destroy(ctx);
  return 0;
}
