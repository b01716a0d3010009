/*
 * tealeaf_oracle.c -- CPU ORACLE for the TeaLeaf.jl implicit heat-conduction solve.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker, never the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may build, load or call it.  The product path (tealeaf.jl_b200/csrc) never
 * links or calls anything in oracle/.
 *
 * PARITY UNPINNED AGAINST THE REFERENCE ITSELF, PINNED EXTERNALLY.  The reference
 * (Laura7089/TeaLeaf.jl @ e696c54) ships no tests, no golden vectors and no
 * tea.in/tea.problems, and there is no Julia binary in this environment, so
 * this restatement cannot be checked against the reference's own outputs.  The
 * pin that stands in is external: the QA checking values upstream TeaLeaf
 * publishes for its benchmark decks -- the numbers the reference's own
 * fieldsummary gate (src/kernels.jl:119-133, src/settings.jl:180-196) compares
 * a run with -- which this oracle reproduces to <= 6e-14 relative at 10^2,
 * 250^2, 500^2 and 1000^2 cells (tests/test_upstream_pin.py,
 * tests/golden/upstream_qa.json: provenance and caveats are stated there).
 * Further checks: (1) an independent NumPy twin (oracle/np_twin.py), bit-equal
 * fields; (2) mathematical invariants (operator symmetry, energy conservation,
 * true-residual agreement); (3) solver-vs-solver agreement; see
 * tests/test_oracle.py and DESIGN.md section 4.
 *
 * What is restated: the algorithm the reference *states*, function by function,
 * with exactly the corrections of SURVEY.md Appendix A (cited as "A#n" below)
 * and nothing else.  All citations are file:line under /root/reference.
 *
 * Layout: Julia column-major (x,y) arrays, a[kk,jj] <-> a[(kk-1) + (jj-1)*x];
 * here 0-based: a[k + j*x], k in [0,x), j in [0,y); x = xcells + 2*hd.
 * Arithmetic: double, serial, loop order jj outer / kk inner as in the
 * reference; compile with -ffp-contract=off so no FMA contraction happens
 * (Julia never contracts implicitly).  OpenMP is used only when
 * tlo_set_threads(n>1) is called (CPU-baseline timing); with one thread every
 * sum is in the reference's serial order.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define TLO_ERROR_START 1e10      /* src/kernels.jl:7 */
#define TLO_ERROR_SWITCH_MAX 1.0  /* src/kernels.jl:8 */
#define TLO_CGEIGENITERS 20       /* src/solvers/Cheby.jl:7 */
#define TLO_CONDUCTIVITY 1        /* src/settings.jl:13 */
#define TLO_RECIP_CONDUCTIVITY 2  /* src/settings.jl:14 */

/* field ids: same numbering as include/tealeaf_b200.h */
enum { F_DENSITY = 0, F_ENERGY0, F_ENERGY, F_U, F_U0, F_P, F_R, F_W, F_KX, F_KY, F_SD, F_COUNT };

typedef struct {
  int x, y, hd, max_iters;
  double *f[F_COUNT];                 /* src/chunk.jl:25-38 (mi, density0 unused by the path) */
  double *cg_alphas, *cg_betas;       /* src/chunk.jl:56-57 */
  double *ch_alphas, *ch_betas;       /* src/chunk.jl:58-59 */
  double theta, eigmin, eigmax;       /* src/chunk.jl:52-54 */
  int literal_smvp;                   /* 1: reproduce the as-written diagonal operator (A#1) */
  int nthreads;
} tlo_chunk;

#define IDX(c, k, j) ((size_t)(k) + (size_t)(j) * (size_t)(c)->x)

tlo_chunk *tlo_create(int xcells, int ycells, int hd, int max_iters) {
  /* src/chunk.jl:68-89 */
  tlo_chunk *c = (tlo_chunk *)calloc(1, sizeof(tlo_chunk));
  c->x = xcells + 2 * hd;
  c->y = ycells + 2 * hd;
  c->hd = hd;
  c->max_iters = max_iters;
  size_t n = (size_t)c->x * (size_t)c->y;
  for (int i = 0; i < F_COUNT; i++) c->f[i] = (double *)calloc(n, sizeof(double));
  c->cg_alphas = (double *)calloc((size_t)max_iters + 2, sizeof(double));
  c->cg_betas = (double *)calloc((size_t)max_iters + 2, sizeof(double));
  c->ch_alphas = (double *)calloc((size_t)max_iters + 2, sizeof(double));
  c->ch_betas = (double *)calloc((size_t)max_iters + 2, sizeof(double));
  c->nthreads = 1;
  return c;
}

void tlo_destroy(tlo_chunk *c) {
  if (!c) return;
  for (int i = 0; i < F_COUNT; i++) free(c->f[i]);
  free(c->cg_alphas); free(c->cg_betas); free(c->ch_alphas); free(c->ch_betas);
  free(c);
}

double *tlo_field(tlo_chunk *c, int id) { return (id >= 0 && id < F_COUNT) ? c->f[id] : NULL; }
double *tlo_coef_array(tlo_chunk *c, int which) {
  switch (which) { case 0: return c->cg_alphas; case 1: return c->cg_betas;
                   case 2: return c->ch_alphas; case 3: return c->ch_betas; }
  return NULL;
}
void tlo_set_literal_smvp(tlo_chunk *c, int on) { c->literal_smvp = on; }
void tlo_set_threads(tlo_chunk *c, int n) {
  c->nthreads = n < 1 ? 1 : n;
#ifdef _OPENMP
  omp_set_num_threads(c->nthreads);
#endif
}
int tlo_max_threads(void) {
#ifdef _OPENMP
  return omp_get_num_procs();
#else
  return 1;
#endif
}
double tlo_get_scalar(tlo_chunk *c, int which) {
  return which == 0 ? c->theta : which == 1 ? c->eigmin : c->eigmax;
}
void tlo_set_eigs(tlo_chunk *c, double eigmin, double eigmax) { c->eigmin = eigmin; c->eigmax = eigmax; }

/* src/kernels.jl:178-184, read as ONE expression (A#1).  Julia's
 * sum((1, kx[x+1,y], kx[x,y], ky[x,y+1], ky[x,y])) folds left to right. */
static inline double smvp(const tlo_chunk *c, const double *a, int k, int j) {
  const double *kx = c->f[F_KX], *ky = c->f[F_KY];
  double consum = (((1.0 + kx[IDX(c, k + 1, j)]) + kx[IDX(c, k, j)]) + ky[IDX(c, k, j + 1)]) + ky[IDX(c, k, j)];
  if (c->literal_smvp) return consum * a[IDX(c, k, j)];      /* kernels.jl:181 as written */
  return consum * a[IDX(c, k, j)]
       - (kx[IDX(c, k + 1, j)] * a[IDX(c, k + 1, j)] + kx[IDX(c, k, j)] * a[IDX(c, k - 1, j)])
       - (ky[IDX(c, k, j + 1)] * a[IDX(c, k, j + 1)] + ky[IDX(c, k, j)] * a[IDX(c, k, j - 1)]);
}
double tlo_smvp(tlo_chunk *c, int field, int k, int j) { return smvp(c, c->f[field], k, j); }

/* src/kernels.jl:191-210 with the 1-based reflection of A#2:
 * b[hd+1-k] = b[hd+k], b[x-hd+k] = b[x-hd+1-k] (1-based), k = 1..depth;
 * x faces over the interior y range, y faces over the interior x range. */
static void update_face(tlo_chunk *c, int depth, double *b) {
  int x = c->x, y = c->y, hd = c->hd;
  for (int j = hd; j < y - hd; j++)
    for (int d = 1; d <= depth; d++) {
      b[IDX(c, hd - d, j)] = b[IDX(c, hd + d - 1, j)];          /* left  */
      b[IDX(c, x - hd + d - 1, j)] = b[IDX(c, x - hd - d, j)];  /* right */
    }
  for (int d = 1; d <= depth; d++)
    for (int k = hd; k < x - hd; k++) {
      b[IDX(c, k, y - hd + d - 1)] = b[IDX(c, k, y - hd - d)];  /* top    */
      b[IDX(c, k, hd - d)] = b[IDX(c, k, hd + d - 1)];          /* bottom */
    }
}

/* src/kernels.jl:146-159.  field_mask bit i <-> field id i.  The sticky
 * `toexchange` bookkeeping lives in the host driver (python), not here. */
void tlo_halo_update(tlo_chunk *c, unsigned field_mask, int depth) {
  for (int i = 0; i < F_COUNT; i++)
    if (field_mask & (1u << i)) update_face(c, depth, c->f[i]);
}

/* src/solvers/CG.jl:47-79.  Returns rro; *status = -1 on invalid coefficient (CG.jl:48-50). */
double tlo_cg_init(tlo_chunk *c, int coef, double rx, double ry, int *status) {
  int x = c->x, y = c->y, hd = c->hd;
  double *u = c->f[F_U], *p = c->f[F_P], *r = c->f[F_R], *w = c->f[F_W];
  double *kx = c->f[F_KX], *ky = c->f[F_KY];
  const double *energy = c->f[F_ENERGY], *density = c->f[F_DENSITY];
  if (status) *status = 0;
  if (coef != TLO_CONDUCTIVITY && coef != TLO_RECIP_CONDUCTIVITY) { if (status) *status = -1; return 0.0; }
  size_t n = (size_t)x * y;
  for (size_t i = 0; i < n; i++) { u[i] = energy[i] * density[i]; p[i] = 0.0; r[i] = 0.0; }  /* :52-54 */
  /* :56-58  w = density^(+1|-1) on all but the outer ring; Julia x^-1 == inv(x) */
  for (int j = 1; j < y - 1; j++)
    for (int k = 1; k < x - 1; k++)
      w[IDX(c, k, j)] = (coef == TLO_CONDUCTIVITY) ? density[IDX(c, k, j)] : 1.0 / density[IDX(c, k, j)];
  /* :61-68  jj = hd+1:y-1, kk = hd+1:x-1 (1-based) */
#pragma omp parallel for if (c->nthreads > 1)
  for (int j = hd; j < y - 1; j++)
    for (int k = hd; k < x - 1; k++) {
      kx[IDX(c, k, j)] = rx * (w[IDX(c, k - 1, j)] + w[IDX(c, k, j)]) / (2.0 * w[IDX(c, k - 1, j)] * w[IDX(c, k, j)]);
      ky[IDX(c, k, j)] = ry * (w[IDX(c, k, j - 1)] + w[IDX(c, k, j)]) / (2.0 * w[IDX(c, k, j - 1)] * w[IDX(c, k, j)]);
    }
  double temp = 0.0;  /* A#8: double accumulator */
#pragma omp parallel for reduction(+ : temp) if (c->nthreads > 1)
  for (int j = hd; j < y - hd; j++)
    for (int k = hd; k < x - hd; k++) {  /* :72-77 */
      size_t i = IDX(c, k, j);
      w[i] = smvp(c, u, k, j);
      r[i] = u[i] - w[i];
      p[i] = r[i];
      temp += r[i] * r[i];
    }
  return temp;
}

/* src/solvers/CG.jl:82-90 */
double tlo_cg_calc_w(tlo_chunk *c) {
  int x = c->x, y = c->y, hd = c->hd;
  double *w = c->f[F_W];
  const double *p = c->f[F_P];
  double temp = 0.0;
#pragma omp parallel for reduction(+ : temp) if (c->nthreads > 1)
  for (int j = hd; j < y - hd; j++)
    for (int k = hd; k < x - hd; k++) {
      size_t i = IDX(c, k, j);
      w[i] = smvp(c, p, k, j);
      temp += w[i] * p[i];
    }
  return temp;
}

/* src/solvers/CG.jl:93-98; the norm is over the interior (A#7: halos of r are 0). */
double tlo_cg_calc_ur(tlo_chunk *c, double alpha) {
  int x = c->x, y = c->y, hd = c->hd;
  double *u = c->f[F_U], *r = c->f[F_R];
  const double *p = c->f[F_P], *w = c->f[F_W];
  double temp = 0.0;
#pragma omp parallel for reduction(+ : temp) if (c->nthreads > 1)
  for (int j = hd; j < y - hd; j++)
    for (int k = hd; k < x - hd; k++) {
      size_t i = IDX(c, k, j);
      u[i] += alpha * p[i];
      r[i] -= alpha * w[i];
      temp += r[i] * r[i];
    }
  return temp;
}

/* src/solvers/CG.jl:101-104 */
void tlo_cg_calc_p(tlo_chunk *c, double beta) {
  int x = c->x, y = c->y, hd = c->hd;
  double *p = c->f[F_P];
  const double *r = c->f[F_R];
#pragma omp parallel for if (c->nthreads > 1)
  for (int j = hd; j < y - hd; j++)
    for (int k = hd; k < x - hd; k++) {
      size_t i = IDX(c, k, j);
      p[i] = beta * p[i] + r[i];
    }
}

/* src/kernels.jl:217-220 */
void tlo_copy_u(tlo_chunk *c) {
  for (int j = c->hd; j < c->y - c->hd; j++)
    for (int k = c->hd; k < c->x - c->hd; k++) c->f[F_U0][IDX(c, k, j)] = c->f[F_U][IDX(c, k, j)];
}

/* src/kernels.jl:227-232 */
void tlo_calc_residual(tlo_chunk *c) {
#pragma omp parallel for if (c->nthreads > 1)
  for (int j = c->hd; j < c->y - c->hd; j++)
    for (int k = c->hd; k < c->x - c->hd; k++)
      c->f[F_R][IDX(c, k, j)] = c->f[F_U0][IDX(c, k, j)] - smvp(c, c->f[F_U], k, j);
}

/* src/kernels.jl:239-242 */
void tlo_finalise(tlo_chunk *c) {
  for (int j = c->hd; j < c->y - c->hd; j++)
    for (int k = c->hd; k < c->x - c->hd; k++)
      c->f[F_ENERGY][IDX(c, k, j)] = c->f[F_U][IDX(c, k, j)] / c->f[F_DENSITY][IDX(c, k, j)];
}

/* sum of squares of a field over the interior (Cheby.jl:68, :103; PPCG.jl:88 with A#20) */
double tlo_norm2(tlo_chunk *c, int field) {
  const double *a = c->f[field];
  double temp = 0.0;
#pragma omp parallel for reduction(+ : temp) if (c->nthreads > 1)
  for (int j = c->hd; j < c->y - c->hd; j++)
    for (int k = c->hd; k < c->x - c->hd; k++) temp += a[IDX(c, k, j)] * a[IDX(c, k, j)];
  return temp;
}

/* src/kernels.jl:119-133: temp = sum(volume .* density .* u) over the interior.
 * vol/mass/ie are the upstream (UoB-HPC/TeaLeaf field_summary, recalled) companions
 * that BASELINE.json's parity criterion names; the reference computes only temp. */
void tlo_field_summary(tlo_chunk *c, double cell_volume, double *vol, double *mass, double *ie, double *temp) {
  double v = 0, m = 0, e = 0, t = 0;
  for (int j = c->hd; j < c->y - c->hd; j++)
    for (int k = c->hd; k < c->x - c->hd; k++) {
      size_t i = IDX(c, k, j);
      double cell_mass = cell_volume * c->f[F_DENSITY][i];
      v += cell_volume;
      m += cell_mass;
      e += cell_mass * c->f[F_ENERGY0][i];
      t += cell_mass * c->f[F_U][i];
    }
  *vol = v; *mass = m; *ie = e; *temp = t;
}

/* Eigenvalues of a symmetric tridiagonal matrix by implicit QL (the classic
 * "tqli" without eigenvectors), replacing the broken tqli!/minmax of
 * src/kernels.jl:60-112, :39 (A#9-11; the author's own TODO at :37).
 * d[0..n-1] diagonal, e[0..n-1] with e[i] = sub-diagonal between i-1 and i (e[0] unused). */
static int tqli_eigs(double *d, double *e, int n) {
  for (int i = 1; i < n; i++) e[i - 1] = e[i];
  e[n - 1] = 0.0;
  for (int l = 0; l < n; l++) {
    int iter = 0, m;
    do {
      for (m = l; m < n - 1; m++) {
        double dd = fabs(d[m]) + fabs(d[m + 1]);
        if (fabs(e[m]) + dd == dd) break;
      }
      if (m != l) {
        if (iter++ == 60) return -1;
        double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
        double r = hypot(g, 1.0);
        g = d[m] - d[l] + e[l] / (g + (g >= 0.0 ? fabs(r) : -fabs(r)));
        double s = 1.0, cc = 1.0, p = 0.0;
        int i;
        for (i = m - 1; i >= l; i--) {
          double f = s * e[i], b = cc * e[i];
          e[i + 1] = (r = hypot(f, g));
          if (r == 0.0) { d[i + 1] -= p; e[m] = 0.0; break; }
          s = f / r; cc = g / r;
          g = d[i + 1] - p;
          r = (d[i] - g) * s + 2.0 * cc * b;
          d[i + 1] = g + (p = s * r);
          g = cc * r - b;
        }
        if (r == 0.0 && i >= l) continue;
        d[l] -= p; e[l] = g; e[m] = 0.0;
      }
    } while (m != l);
  }
  return 0;
}

/* src/kernels.jl:19-51: Lanczos tridiagonal from cg_alphas/cg_betas, extreme
 * eigenvalues widened by 0.95/1.05.  Returns 0, -1 (QL failed), -2 (negative eigenvalue, :41-43). */
int tlo_eigenvalues(tlo_chunk *c, int cgiters) {
  if (cgiters < 1) return -3;
  double *diag = (double *)calloc((size_t)cgiters, sizeof(double));
  double *off = (double *)calloc((size_t)cgiters, sizeof(double));
  for (int i = 0; i < cgiters; i++) {
    diag[i] = 1.0 / c->cg_alphas[i];                                     /* :20 */
    if (i > 0) diag[i] += c->cg_betas[i - 1] / c->cg_alphas[i - 1];      /* :25-27 */
    if (i < cgiters - 1) off[i + 1] = sqrt(c->cg_betas[i]) / c->cg_alphas[i];  /* :28-30 */
  }
  int rc = tqli_eigs(diag, off, cgiters);
  double mn = diag[0], mx = diag[0];
  for (int i = 1; i < cgiters; i++) { if (diag[i] < mn) mn = diag[i]; if (diag[i] > mx) mx = diag[i]; }
  free(diag); free(off);
  if (rc) return -1;
  if (mn < 0.0 || mx < 0.0) { c->eigmin = mn; c->eigmax = mx; return -2; }
  c->eigmin = mn * 0.95;  /* :47 */
  c->eigmax = mx * 1.05;  /* :48 */
  return 0;
}

/* src/solvers/Cheby.jl:121-135 (1-based chebyα[ii] <-> ch_alphas[ii-1]) */
void tlo_cheby_coef(tlo_chunk *c, int n) {
  c->theta = (c->eigmax + c->eigmin) / 2.0;
  double delta = (c->eigmax - c->eigmin) / 2.0;
  double sigma = c->theta / delta;
  double rho_old = 1.0 / sigma;
  if (n > c->max_iters) n = c->max_iters;
  for (int i = 0; i < n; i++) {
    double rho_new = 1.0 / (2.0 * sigma - rho_old);
    c->ch_alphas[i] = rho_new * rho_old;
    c->ch_betas[i] = 2.0 * rho_new / delta;
    rho_old = rho_new;
  }
}

/* src/solvers/Cheby.jl:109-118 (A#25: the log base cancels).  Julia round = ties-to-even = rint. */
int tlo_cheby_calc_iter(tlo_chunk *c, double error, double bb) {
  double connum = c->eigmax / c->eigmin;
  double it_alpha = 2.220446049250313e-16 * bb / (4.0 * error);
  double gamma = (sqrt(connum) - 1.0) / (sqrt(connum) + 1.0);
  double v = rint(log(it_alpha) / (2.0 * log(gamma)));
  if (!(v == v) || v > 2.0e9 || v < -2.0e9) return 2000000000;
  return (int)v;
}

/* field part of src/solvers/Cheby.jl:64-81 (after eigenvalues!/coef!): returns bb = sum(u0^2). */
double tlo_cheby_init_fields(tlo_chunk *c) {
  int x = c->x, y = c->y, hd = c->hd;
  double *u = c->f[F_U], *p = c->f[F_P], *r = c->f[F_R], *w = c->f[F_W];
  const double *u0 = c->f[F_U0];
  double bb = tlo_norm2(c, F_U0);                 /* :68 */
#pragma omp parallel for if (c->nthreads > 1)
  for (int j = hd; j < y - hd; j++)
    for (int k = hd; k < x - hd; k++) {           /* :71-75 */
      size_t i = IDX(c, k, j);
      w[i] = smvp(c, u, k, j);
      r[i] = u0[i] - w[i];
      p[i] = r[i] / c->theta;
    }
#pragma omp parallel for if (c->nthreads > 1)
  for (int j = hd; j < y - hd; j++)
    for (int k = hd; k < x - hd; k++) u[IDX(c, k, j)] += p[IDX(c, k, j)];  /* :76 */
  tlo_halo_update(c, 1u << F_U, 1);               /* :78 */
  return bb;
}

/* src/solvers/Cheby.jl:84-106.  cheby_step = chebyiters (1-based); uses chebyα[chebyiters+1]. */
double tlo_cheby_iterate(tlo_chunk *c, int cheby_step, int calc_2norm, double error) {
  int x = c->x, y = c->y, hd = c->hd;
  double *u = c->f[F_U], *p = c->f[F_P], *r = c->f[F_R], *w = c->f[F_W];
  const double *u0 = c->f[F_U0];
  double alpha = c->ch_alphas[cheby_step], beta = c->ch_betas[cheby_step];  /* 1-based [chebyiters+1] */
#pragma omp parallel for if (c->nthreads > 1)
  for (int j = hd; j < y - hd; j++)
    for (int k = hd; k < x - hd; k++) {
      size_t i = IDX(c, k, j);
      w[i] = smvp(c, u, k, j);
      r[i] = u0[i] - w[i];
      p[i] = alpha * p[i] + beta * r[i];
    }
#pragma omp parallel for if (c->nthreads > 1)
  for (int j = hd; j < y - hd; j++)
    for (int k = hd; k < x - hd; k++) u[IDX(c, k, j)] += p[IDX(c, k, j)];
  if (calc_2norm) error = tlo_norm2(c, F_R);
  return error;
}

/* src/solvers/PPCG.jl:97-100 (the init! that mainstep! means to call, A#18) */
void tlo_ppcg_init_sd(tlo_chunk *c) {
  for (int j = c->hd; j < c->y - c->hd; j++)
    for (int k = c->hd; k < c->x - c->hd; k++)
      c->f[F_SD][IDX(c, k, j)] = c->f[F_R][IDX(c, k, j)] / c->theta;
}

/* one inner step of src/solvers/PPCG.jl:75-84, two-phase (A#17); pp is 1-based. */
void tlo_ppcg_inner(tlo_chunk *c, int pp) {
  int x = c->x, y = c->y, hd = c->hd;
  double *u = c->f[F_U], *r = c->f[F_R], *sd = c->f[F_SD];
  double alpha = c->ch_alphas[pp - 1], beta = c->ch_betas[pp - 1];
  tlo_halo_update(c, 1u << F_SD, 1);              /* :76 */
#pragma omp parallel for if (c->nthreads > 1)
  for (int j = hd; j < y - hd; j++)
    for (int k = hd; k < x - hd; k++) {
      size_t i = IDX(c, k, j);
      r[i] -= smvp(c, sd, k, j);
      u[i] += sd[i];
    }
#pragma omp parallel for if (c->nthreads > 1)
  for (int j = hd; j < y - hd; j++)
    for (int k = hd; k < x - hd; k++) {
      size_t i = IDX(c, k, j);
      sd[i] = alpha * sd[i] + beta * r[i];
    }
}

/* ------------------------------------------------------------------ */
/* Whole-solve drivers (the `solve!` of each solver module).           */
/* ------------------------------------------------------------------ */
typedef struct {
  int iters;        /* outer iterations executed (tt at exit) */
  int cg_iters;     /* CG iterations (presteps for Cheby/PPCG) */
  int cheby_iters;  /* Chebyshev iterations, or PPCG outer iterations */
  int est_iters;    /* Cheby estimate */
  int inner_total;  /* PPCG inner steps executed */
  int status;       /* 0 ok, <0 error */
  double error, eigmin, eigmax;
} tlo_result;

/* src/solvers/CG.jl:32-44 */
static double cg_mainstep(tlo_chunk *c, int tt, double rro) {
  double pw = tlo_cg_calc_w(c);
  double alpha = rro / pw;
  c->cg_alphas[tt - 1] = alpha;
  double rrn = tlo_cg_calc_ur(c, alpha);
  double beta = rrn / rro;
  c->cg_betas[tt - 1] = beta;
  tlo_cg_calc_p(c, beta);
  return rrn;
}

#define MASK_UP ((1u << F_U) | (1u << F_P))

/* src/solvers/CG.jl:7-29 */
void tlo_cg_solve(tlo_chunk *c, int coef, double rx, double ry, double eps, int max_iters, tlo_result *res) {
  memset(res, 0, sizeof(*res));
  double rro = tlo_cg_init(c, coef, rx, ry, &res->status);
  if (res->status) return;
  tlo_halo_update(c, MASK_UP, 1);   /* :11 */
  tlo_copy_u(c);                    /* :12 */
  double error = TLO_ERROR_START;
  int iters = 0;
  if (max_iters > c->max_iters) max_iters = c->max_iters;
  for (int tt = 1; tt <= max_iters; tt++) {
    iters = tt;
    rro = error = cg_mainstep(c, tt, rro);
    tlo_halo_update(c, MASK_UP, 1); /* :22 sticky {u,p} */
    if (sqrt(fabs(error)) < eps) break;   /* :24 */
  }
  res->iters = res->cg_iters = iters;
  res->error = error;
}

static int switch_now(int errorswitch, double error, double epslim, int tt, int presteps) {
  /* Cheby.jl:25-29 / PPCG.jl:25-30 */
  return errorswitch ? (error < epslim && tt > TLO_CGEIGENITERS) : (tt > presteps && error < TLO_ERROR_SWITCH_MAX);
}

/* src/solvers/Cheby.jl:10-61 with A#12 (CG preamble), A#13 (rro update), A#14 (tt-1 CG iterations). */
void tlo_cheby_solve(tlo_chunk *c, int coef, double rx, double ry, double eps, int max_iters, int presteps,
                     double epslim, int errorswitch, tlo_result *res) {
  memset(res, 0, sizeof(*res));
  double error = TLO_ERROR_START, bb = 0.0;
  int estiter = 0, chebyiters = 0, tt;
  double rro = tlo_cg_init(c, coef, rx, ry, &res->status);
  if (res->status) return;
  tlo_halo_update(c, MASK_UP, 1);
  tlo_copy_u(c);
  unsigned sticky = MASK_UP;
  if (max_iters > c->max_iters) max_iters = c->max_iters;
  for (tt = 1; tt <= max_iters; tt++) {
    res->iters = tt;
    int sw = chebyiters != 0 || switch_now(errorswitch, error, epslim, tt, presteps);
    if (!sw) {
      rro = error = cg_mainstep(c, tt, rro);
      res->cg_iters = tt;
    } else {
      chebyiters++;
      if (chebyiters == 1) {
        int rc = tlo_eigenvalues(c, tt - 1);
        if (rc) { res->status = rc; break; }
        tlo_cheby_coef(c, max_iters - (tt - 1));
        bb = tlo_cheby_init_fields(c);
        sticky = 1u << F_U;
        error = tlo_cheby_iterate(c, chebyiters, 1, error);
        estiter = tlo_cheby_calc_iter(c, error, bb);
      } else {
        int calc2norm = (chebyiters >= estiter) && ((tt + 1) % 10 == 0);
        error = tlo_cheby_iterate(c, chebyiters, calc2norm, error);
      }
    }
    tlo_halo_update(c, sticky, 1);
    if (fabs(error) < eps) break;   /* :57, no sqrt (A#24: keep as written) */
  }
  res->cheby_iters = chebyiters; res->est_iters = estiter; res->error = error;
  res->eigmin = c->eigmin; res->eigmax = c->eigmax;
}

/* src/solvers/PPCG.jl:64-94 with A#16, #17, #18, #20 */
static double ppcg_mainstep(tlo_chunk *c, int inner_steps, double rro) {
  double pw = tlo_cg_calc_w(c);
  double alpha = rro / pw;
  double rrn = tlo_cg_calc_ur(c, alpha);
  (void)rrn;
  tlo_ppcg_init_sd(c);
  for (int pp = 1; pp <= inner_steps; pp++) tlo_ppcg_inner(c, pp);
  rrn = tlo_norm2(c, F_R);
  double beta = rrn / rro;
  tlo_cg_calc_p(c, beta);
  return rrn;
}

/* src/solvers/PPCG.jl:9-55 with A#12, #14, #15, #16, #19 */
void tlo_ppcg_solve(tlo_chunk *c, int coef, double rx, double ry, double eps, int max_iters, int presteps,
                    double epslim, int errorswitch, int inner_steps, tlo_result *res) {
  memset(res, 0, sizeof(*res));
  int ppcgiters = 0, tt;
  double rro = tlo_cg_init(c, coef, rx, ry, &res->status);
  if (res->status) return;
  tlo_halo_update(c, MASK_UP, 1);
  tlo_copy_u(c);
  unsigned sticky = MASK_UP;
  double error = TLO_ERROR_START;
  if (max_iters > c->max_iters) max_iters = c->max_iters;
  for (tt = 1; tt <= max_iters; tt++) {
    res->iters = tt;
    int sw = ppcgiters != 0 || switch_now(errorswitch, error, epslim, tt, presteps);
    if (!sw) {
      rro = error = cg_mainstep(c, tt, rro);
      res->cg_iters = tt;
    } else {
      ppcgiters++;
      if (ppcgiters == 1) {
        int rc = tlo_eigenvalues(c, tt - 1);
        if (rc) { res->status = rc; break; }
        tlo_cheby_coef(c, inner_steps);
        tlo_calc_residual(c);                   /* PPCG.jl:59 */
        tlo_halo_update(c, 1u << F_P, 1);       /* PPCG.jl:60 */
        sticky = 1u << F_P;
        rro = tlo_norm2(c, F_R);                /* A#19 */
      }
      rro = error = ppcg_mainstep(c, inner_steps, rro);
      res->inner_total += inner_steps;
    }
    tlo_halo_update(c, sticky, 1);
    if (fabs(error) < eps) break;               /* PPCG.jl:51 */
  }
  res->cheby_iters = ppcgiters; res->error = error;
  res->eigmin = c->eigmin; res->eigmax = c->eigmax;
}

/* ------------------------------------------------------------------ */
/* Jacobi (src/solvers/Jacobi.jl) -- SURVEY.md section 8(f) item 1, A#21.     */
/* ------------------------------------------------------------------ */
/* src/solvers/Jacobi.jl:33-60.  *status = -1 on an invalid coefficient (:34-36, as written:
 * only values below min(CONDUCTIVITY, RECIP_CONDUCTIVITY) are rejected). */
void tlo_jacobi_init(tlo_chunk *c, int coef, double rx, double ry, int *status) {
  int x = c->x, y = c->y, hd = c->hd;
  double *u = c->f[F_U], *u0 = c->f[F_U0], *kx = c->f[F_KX], *ky = c->f[F_KY];
  const double *energy = c->f[F_ENERGY], *density = c->f[F_DENSITY];
  if (status) *status = 0;
  if (coef < TLO_CONDUCTIVITY) { if (status) *status = -1; return; }
  size_t n = (size_t)x * y;
  for (size_t i = 0; i < n; i++) { double t = energy[i] * density[i]; u0[i] = t; u[i] = t; }   /* :39-41 */
  /* :43-51  jj = hd+1:y-1, kk = hd+1:x-1 (1-based); density^p with p = +-1 (x^-1 == inv(x)) */
  for (int j = hd; j < y - 1; j++)
    for (int k = hd; k < x - 1; k++) {
      double dc = density[IDX(c, k, j)], dl = density[IDX(c, k - 1, j)], dd = density[IDX(c, k, j - 1)];
      if (coef != TLO_CONDUCTIVITY) { dc = 1.0 / dc; dl = 1.0 / dl; dd = 1.0 / dd; }
      kx[IDX(c, k, j)] = rx * (dl + dc) / (2.0 * dl * dc);
      ky[IDX(c, k, j)] = ry * (dd + dc) / (2.0 * dd * dc);
    }
  tlo_copy_u(c);   /* :53 */
}

/* src/solvers/Jacobi.jl:62-82: r .= u (whole array), Jacobi sweep on the interior, returns
 * sum(|u - r|) (halo cells contribute 0: r is a copy of u there). */
double tlo_jacobi_iterate(tlo_chunk *c) {
  int x = c->x, y = c->y, hd = c->hd;
  double *u = c->f[F_U], *r = c->f[F_R];
  const double *u0 = c->f[F_U0], *kx = c->f[F_KX], *ky = c->f[F_KY];
  memcpy(r, u, (size_t)x * y * sizeof(double));
  double err = 0.0;
#pragma omp parallel for reduction(+ : err) if (c->nthreads > 1)
  for (int j = hd; j < y - hd; j++)
    for (int k = hd; k < x - hd; k++) {
      size_t i = IDX(c, k, j);
      double num = (((u0[i] + kx[IDX(c, k + 1, j)] * r[IDX(c, k + 1, j)]) + kx[i] * r[IDX(c, k - 1, j)]) +
                    ky[IDX(c, k, j + 1)] * r[IDX(c, k, j + 1)]) + ky[i] * r[IDX(c, k, j - 1)];
      double den = (((1.0 + kx[i]) + kx[IDX(c, k + 1, j)]) + ky[i]) + ky[IDX(c, k, j + 1)];
      u[i] = num / den;
      err += fabs(u[i] - r[i]);
    }
  return err;
}

/* src/solvers/Jacobi.jl:7-31 with A#21 (entry point renamed solve!, `error +=` read as `=`). */
void tlo_jacobi_solve(tlo_chunk *c, int coef, double rx, double ry, double eps, int max_iters, tlo_result *res) {
  memset(res, 0, sizeof(*res));
  tlo_jacobi_init(c, coef, rx, ry, &res->status);
  if (res->status) return;
  double error = TLO_ERROR_START;
  if (max_iters > c->max_iters) max_iters = c->max_iters;
  for (int tt = 1; tt <= max_iters; tt++) {
    res->iters = tt;
    error = tlo_jacobi_iterate(c);             /* :14 */
    if (tt % 50 == 0) {                        /* :16-21 */
      tlo_halo_update(c, 1u << F_U, 1);
      tlo_calc_residual(c);
      error = tlo_norm2(c, F_R);
    }
    tlo_halo_update(c, 1u << F_U, 1);          /* :23 sticky {u} */
    if (fabs(error) < eps) break;              /* :26 */
  }
  res->error = error;
}

/* src/kernels.jl:166-170 */
void tlo_solve_finished(tlo_chunk *c, int check_result) {
  if (check_result) tlo_calc_residual(c);
  tlo_finalise(c);
  tlo_halo_update(c, 1u << F_ENERGY, 1);
}

/* CPU-baseline helper for bench.py: `iters` CG iterations (w!, ur!, p!, halo) without
 * a convergence test, on whatever state the chunk holds after tlo_cg_init. */
double tlo_cg_fixed_iters(tlo_chunk *c, double rro, int iters) {
  for (int tt = 1; tt <= iters; tt++) {
    rro = cg_mainstep(c, tt, rro);
    tlo_halo_update(c, MASK_UP, 1);
  }
  return rro;
}
