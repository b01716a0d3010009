// tl_eigen.h -- host-side scalar work of the Chebyshev/PPCG switch (tiny: <= max_iters
// scalars), kept on the host as in the reference:
//   eigenvalues!  src/kernels.jl:19-51   (Lanczos tridiagonal from cgα/cgβ, extreme
//                                         eigenvalues widened by 0.95 / 1.05)
//   Cheby.coef!   src/solvers/Cheby.jl:121-135
//   Cheby.calciter src/solvers/Cheby.jl:109-118
// The reference's tqli!/minmax are broken; its author's TODO
// (kernels.jl:37) asks for "any correct symmetric-tridiagonal eigen-solve".  Here the two
// extreme eigenvalues are found by Sturm-sequence bisection (independent of the QL
// iteration the oracle uses, so the two cross-check each other).
#pragma once
#include <cmath>
#include <vector>
#include <algorithm>

namespace tl {

// number of eigenvalues of the symmetric tridiagonal (d, e) that are < x
inline int sturm_count(const std::vector<double> &d, const std::vector<double> &e2, double x) {
  int count = 0;
  double q = 1.0;
  const int n = (int)d.size();
  for (int i = 0; i < n; i++) {
    const double off = (i == 0) ? 0.0 : e2[i];
    q = d[i] - x - (q != 0.0 ? off / q : off / 1e-300);
    if (q < 0.0) count++;
  }
  return count;
}

// k-th smallest eigenvalue (k = 0 .. n-1) by bisection to full double precision
inline double tridiag_eig_k(const std::vector<double> &d, const std::vector<double> &e2, int k, double lo, double hi) {
  for (int it = 0; it < 200; it++) {
    const double mid = 0.5 * (lo + hi);
    if (mid <= lo || mid >= hi) break;
    if (sturm_count(d, e2, mid) > k) hi = mid; else lo = mid;
  }
  return 0.5 * (lo + hi);
}

// src/kernels.jl:19-51.  Returns 0 ok, -2 negative eigenvalue, -3 no CG iterations.
inline int eigenvalues(const double *cg_alphas, const double *cg_betas, int cgiters, double *eigmin, double *eigmax) {
  if (cgiters < 1) return -3;
  std::vector<double> d(cgiters), e(cgiters, 0.0), e2(cgiters, 0.0);
  for (int i = 0; i < cgiters; i++) {
    d[i] = 1.0 / cg_alphas[i];
    if (i > 0) d[i] += cg_betas[i - 1] / cg_alphas[i - 1];
    if (i < cgiters - 1) e[i + 1] = std::sqrt(cg_betas[i]) / cg_alphas[i];
  }
  double lo = d[0], hi = d[0];
  for (int i = 0; i < cgiters; i++) {
    e2[i] = e[i] * e[i];
    const double rad = std::fabs(e[i]) + (i + 1 < cgiters ? std::fabs(e[i + 1]) : 0.0);
    lo = std::min(lo, d[i] - rad);
    hi = std::max(hi, d[i] + rad);
  }
  const double span = std::max(hi - lo, 1e-300);
  lo -= 1e-3 * span; hi += 1e-3 * span;
  const double mn = tridiag_eig_k(d, e2, 0, lo, hi);
  const double mx = tridiag_eig_k(d, e2, cgiters - 1, lo, hi);
  *eigmin = mn; *eigmax = mx;
  if (mn < 0.0 || mx < 0.0) return -2;
  *eigmin = mn * 0.95;
  *eigmax = mx * 1.05;
  return 0;
}

// src/solvers/Cheby.jl:121-135
inline double cheby_coef(double eigmin, double eigmax, int n, double *alphas, double *betas) {
  const double theta = (eigmax + eigmin) / 2.0;
  const double delta = (eigmax - eigmin) / 2.0;
  const double sigma = theta / delta;
  double rho_old = 1.0 / sigma;
  for (int i = 0; i < n; i++) {
    const double rho_new = 1.0 / (2.0 * sigma - rho_old);
    alphas[i] = rho_new * rho_old;
    betas[i] = 2.0 * rho_new / delta;
    rho_old = rho_new;
  }
  return theta;
}

// src/solvers/Cheby.jl:109-118
inline int cheby_calc_iter(double eigmin, double eigmax, double error, double bb) {
  const double connum = eigmax / eigmin;
  const double it_alpha = 2.220446049250313e-16 * bb / (4.0 * error);
  const double gamma = (std::sqrt(connum) - 1.0) / (std::sqrt(connum) + 1.0);
  const double v = std::rint(std::log(it_alpha) / (2.0 * std::log(gamma)));
  if (!(v == v) || v > 2.0e9 || v < -2.0e9) return 2000000000;
  return (int)v;
}

}  // namespace tl
