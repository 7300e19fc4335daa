// Trust-region-reflective least squares WITHOUT bounds, restated on the normal equations.
//
// The reference calls scipy.optimize.least_squares with defaults (analysis/fitter.py:300), i.e.
// method='trf', tr_solver='exact', x_scale=1, ftol=xtol=gtol=1e-8, max_nfev=100*n, no bounds.
// This header restates that algorithm (scipy/optimize/_lsq/trf.py:415-587 trf_no_bounds and
// _lsq/common.py:57-168 solve_lsq_trust_region, :222-245 update_tr_radius, :325-360
// evaluate_quadratic, :705-717 check_termination) as a state machine driven by evaluations of
//   cost = 0.5 r^T r,  g = J^T r,  A = J^T J
// -- everything scipy takes from svd(J) follows from the eigen-decomposition A = V diag(s^2) V^T:
//   s * (U^T f) = V^T g,   (U^T f) / s = V^T g / s^2,   ||J p||^2 = p^T A p.
// The same code runs on the host (single large fits) and inside the batched device solver.
#pragma once
#include <math.h>

#include "../../include/fitburst_b200.h"

#if defined(__CUDACC__)
#define FB_HD __host__ __device__
#else
#define FB_HD
#endif

namespace fb {

constexpr int kMaxN = FB_MAX_FREE;
constexpr double kEps = 2.220446049250313e-16;  // np.finfo(float).eps (common.py:14)

// Cyclic Jacobi eigen-decomposition of a symmetric n x n matrix (row-major, leading dim ld).
// On exit a's diagonal holds the eigenvalues and v (n x n, ld) the eigenvectors in columns;
// both are then sorted by DESCENDING eigenvalue to match the ordering of svd(J).
FB_HD inline void jacobi_eigh(double* a, double* v, double* lam, int n, int ld) {
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) v[i * ld + j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    int rotated = 0;
    for (int p = 0; p < n - 1; ++p) {
      for (int q = p + 1; q < n; ++q) {
        const double apq = a[p * ld + q];
        if (apq == 0.0) continue;
        const double app = a[p * ld + p], aqq = a[q * ld + q];
        // converged pair: |a_pq| negligible against sqrt(a_pp a_qq) (relative criterion keeps
        // the small eigenvalues of the ill-conditioned J^T J accurate).
        if (fabs(apq) <= 1e-17 * sqrt(fabs(app * aqq))) continue;
        rotated = 1;
        const double theta = (aqq - app) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; ++k) {  // A <- A P   (columns p, q)
          const double akp = a[k * ld + p], akq = a[k * ld + q];
          a[k * ld + p] = c * akp - s * akq;
          a[k * ld + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {  // A <- P^T A (rows p, q)
          const double apk = a[p * ld + k], aqk = a[q * ld + k];
          a[p * ld + k] = c * apk - s * aqk;
          a[q * ld + k] = s * apk + c * aqk;
        }
        a[p * ld + q] = 0.0;
        a[q * ld + p] = 0.0;
        for (int k = 0; k < n; ++k) {  // V <- V P
          const double vkp = v[k * ld + p], vkq = v[k * ld + q];
          v[k * ld + p] = c * vkp - s * vkq;
          v[k * ld + q] = s * vkp + c * vkq;
        }
      }
    }
    if (!rotated) break;
  }
  for (int i = 0; i < n; ++i) lam[i] = a[i * ld + i];
  // selection sort, descending, permuting eigenvector columns alongside.
  for (int i = 0; i < n - 1; ++i) {
    int best = i;
    for (int j = i + 1; j < n; ++j)
      if (lam[j] > lam[best]) best = j;
    if (best != i) {
      const double tmp = lam[i];
      lam[i] = lam[best];
      lam[best] = tmp;
      for (int k = 0; k < n; ++k) {
        const double t2 = v[k * ld + i];
        v[k * ld + i] = v[k * ld + best];
        v[k * ld + best] = t2;
      }
    }
  }
}

// State of one fit. Sized for FB_MAX_FREE parameters.
struct TrfState {
  int n;
  long long m;  // number of residuals (num_freq * num_time), used by the full-rank test
  double ftol, xtol, gtol;
  int max_nfev;
  // current accepted point
  double x[kMaxN], g[kMaxN];
  double A[kMaxN * kMaxN];  // J^T J at x (row-major, ld = n)
  double cost;
  // decomposition of A, valid for the current outer iteration
  double lam[kMaxN], V[kMaxN * kMaxN], suf[kMaxN];
  double work[kMaxN * kMaxN];  // scratch copy of A destroyed by the eigen-solver
  int decomposed;
  // trust region
  double Delta, alpha;
  // trial step
  double p[kMaxN], x_new[kMaxN];
  double predicted, step_norm;
  double actual;
  int nfev, njev;
  int status;    // -100 while running; scipy status once finished
  int finished;
  double g_norm;
};

FB_HD inline double vnorm(const double* v, int n) {
  // np.linalg.norm for 1-D: sqrt(dot(v, v))
  double s = 0.0;
  for (int i = 0; i < n; ++i) s += v[i] * v[i];
  return sqrt(s);
}

// common.py:57-168 on the eigen-system. Returns p and updates alpha.
FB_HD inline void solve_lsq_trust_region(TrfState& st) {
  const int n = st.n;
  const double Delta = st.Delta;
  double s[kMaxN];
  for (int i = 0; i < n; ++i) s[i] = sqrt(st.lam[i] > 0.0 ? st.lam[i] : 0.0);
  const double* suf = st.suf;
  bool full_rank = false;
  if (st.m >= n) {
    const double threshold = kEps * (double)st.m * s[0];
    full_rank = s[n - 1] > threshold;
  }
  double coef[kMaxN];
  if (full_rank) {
    for (int i = 0; i < n; ++i) coef[i] = suf[i] / (s[i] * s[i]);  // uf / s
    double pn = 0.0;
    for (int i = 0; i < n; ++i) {
      double acc = 0.0;
      for (int k = 0; k < n; ++k) acc += st.V[i * n + k] * coef[k];
      st.p[i] = -acc;
      pn += acc * acc;
    }
    if (sqrt(pn) <= Delta) {
      st.alpha = 0.0;
      return;
    }
  }
  double alpha_upper = vnorm(suf, n) / Delta;
  double alpha_lower = 0.0;
  auto phi_and_derivative = [&](double alpha, double& phi, double& phi_prime) {
    double pn2 = 0.0, acc = 0.0;
    for (int i = 0; i < n; ++i) {
      const double denom = s[i] * s[i] + alpha;
      const double q = suf[i] / denom;
      pn2 += q * q;
      acc += suf[i] * suf[i] / (denom * denom * denom);
    }
    const double p_norm = sqrt(pn2);
    phi = p_norm - Delta;
    phi_prime = -acc / p_norm;
  };
  if (full_rank) {
    double phi, phi_prime;
    phi_and_derivative(0.0, phi, phi_prime);
    alpha_lower = -phi / phi_prime;
  }
  double alpha = st.alpha;
  if (!full_rank && alpha == 0.0) {
    const double geo = sqrt(alpha_lower * alpha_upper);
    alpha = fmax(0.001 * alpha_upper, geo);
  }
  for (int it = 0; it < 10; ++it) {
    if (alpha < alpha_lower || alpha > alpha_upper) {
      const double geo = sqrt(alpha_lower * alpha_upper);
      alpha = fmax(0.001 * alpha_upper, geo);
    }
    double phi, phi_prime;
    phi_and_derivative(alpha, phi, phi_prime);
    if (phi < 0.0) alpha_upper = alpha;
    const double ratio = phi / phi_prime;
    alpha_lower = fmax(alpha_lower, alpha - ratio);
    alpha -= (phi + Delta) * ratio / Delta;
    if (fabs(phi) < 0.01 * Delta) break;
  }
  for (int i = 0; i < n; ++i) coef[i] = suf[i] / (s[i] * s[i] + alpha);
  double pn = 0.0;
  for (int i = 0; i < n; ++i) {
    double acc = 0.0;
    for (int k = 0; k < n; ++k) acc += st.V[i * n + k] * coef[k];
    st.p[i] = -acc;
    pn += acc * acc;
  }
  const double scale = Delta / sqrt(pn);
  for (int i = 0; i < n; ++i) st.p[i] *= scale;
  st.alpha = alpha;
}

// trf.py:418-456: state after the first evaluation at x0.
FB_HD inline void trf_init(TrfState& st, int n, long long m, const double* x0, double ftol, double xtol,
                           double gtol, int max_nfev) {
  st.n = n;
  st.m = m;
  st.ftol = ftol;
  st.xtol = xtol;
  st.gtol = gtol;
  st.max_nfev = max_nfev > 0 ? max_nfev : 100 * n;
  for (int i = 0; i < n; ++i) st.x[i] = x0[i];
  st.Delta = vnorm(x0, n);
  if (st.Delta == 0.0) st.Delta = 1.0;
  st.alpha = 0.0;
  st.nfev = 0;
  st.njev = 0;
  st.status = -100;
  st.finished = 0;
  st.decomposed = 0;
  st.actual = -1.0;
  st.g_norm = 0.0;
}

// Loads [J r]^T [J r] ((n+1) x (n+1), row-major) into the accepted-point slots.
FB_HD inline void trf_load_accepted(TrfState& st, const double* gram) {
  const int n = st.n, ld = n + 1;
  for (int i = 0; i < n; ++i) {
    st.g[i] = gram[i * ld + n];
    for (int j = 0; j < n; ++j) st.A[i * n + j] = gram[i * ld + j];
  }
  st.cost = 0.5 * gram[n * ld + n];
  st.decomposed = 0;
}

// Top of trf.py's outer loop (trf.py:465-482), scalar part: gtol / status / max_nfev checks.
// Returns 0 when the fit has finished, 1 when a new outer iteration starts (the caller must
// decompose J^T J), 2 when the current decomposition is still valid (the inner loop is retrying
// with a smaller radius).
FB_HD inline int trf_head_decide(TrfState& st) {
  const int n = st.n;
  if (st.decomposed) return 2;
  double gn = 0.0;
  for (int i = 0; i < n; ++i) gn = fmax(gn, fabs(st.g[i]));
  st.g_norm = gn;
  if (gn < st.gtol) st.status = 1;
  if (st.status != -100 || st.nfev == st.max_nfev) {
    if (st.status == -100) st.status = 0;
    st.finished = 1;
    return 0;
  }
  return 1;
}

// trf_head_decide + the scratch copy of J^T J the eigen-solver destroys.
FB_HD inline int trf_head(TrfState& st) {
  const int h = trf_head_decide(st);
  if (h == 1)
    for (int i = 0; i < st.n * st.n; ++i) st.work[i] = st.A[i];
  return h;
}

// After s^2, V = eigh(work): s * U^T f = V^T g.
FB_HD inline void trf_after_decomposition(TrfState& st) {
  const int n = st.n;
  for (int k = 0; k < n; ++k) {
    double acc = 0.0;
    for (int i = 0; i < n; ++i) acc += st.V[i * n + k] * st.g[i];
    st.suf[k] = acc;
  }
  st.decomposed = 1;
  st.actual = -1.0;
}

// Inner-loop body up to the function evaluation, trf.py:503-513: trial step and prediction.
FB_HD inline void trf_step(TrfState& st) {
  const int n = st.n;
  solve_lsq_trust_region(st);
  double q = 0.0, l = 0.0;
  for (int i = 0; i < n; ++i) {
    double acc = 0.0;
    for (int j = 0; j < n; ++j) acc += st.A[i * n + j] * st.p[j];
    q += st.p[i] * acc;
    l += st.p[i] * st.g[i];
  }
  st.predicted = -(0.5 * q + l);
  for (int i = 0; i < n; ++i) st.x_new[i] = st.x[i] + st.p[i];
  st.step_norm = vnorm(st.p, n);
}

// Returns 1 if a trial point st.x_new is ready to be evaluated, 0 if the fit has finished.
FB_HD inline int trf_propose(TrfState& st) {
  const int h = trf_head(st);
  if (h == 0) return 0;
  if (h == 1) {
    jacobi_eigh(st.work, st.V, st.lam, st.n, st.n);
    trf_after_decomposition(st);
  }
  trf_step(st);
  return 1;
}

// Remainder of the inner loop after evaluating the trial point (trf.py:514-567), scalar part:
// cost_new = 0.5 chi^2 at st.x_new. Returns 0 when the inner loop goes on (same point, new
// radius), 1 when it was left with the trial point ACCEPTED (the caller copies x_new -> x, loads
// the new J^T J / J^T r and counts a Jacobian evaluation), 2 when it was left without acceptance.
FB_HD inline int trf_update_decide(TrfState& st, double cost_new) {
  const int n = st.n;
  st.nfev += 1;
  bool inner_break = false;
  if (!isfinite(cost_new)) {
    st.Delta = 0.25 * st.step_norm;  // trf.py:519-521
  } else {
    st.actual = st.cost - cost_new;
    // update_tr_radius, common.py:222-245
    double ratio;
    if (st.predicted > 0.0) ratio = st.actual / st.predicted;
    else if (st.predicted == 0.0 && st.actual == 0.0) ratio = 1.0;
    else ratio = 0.0;
    double Delta_new = st.Delta;
    if (ratio < 0.25) Delta_new = 0.25 * st.step_norm;
    else if (ratio > 0.75 && st.step_norm > 0.95 * st.Delta) Delta_new = st.Delta * 2.0;
    // check_termination, common.py:705-717
    const double x_norm = vnorm(st.x, n);
    const bool ftol_ok = st.actual < st.ftol * st.cost && ratio > 0.25;
    const bool xtol_ok = st.step_norm < st.xtol * (st.xtol + x_norm);
    if (ftol_ok && xtol_ok) st.status = 4;
    else if (ftol_ok) st.status = 2;
    else if (xtol_ok) st.status = 3;
    if (st.status != -100) {
      inner_break = true;
    } else {
      st.alpha *= st.Delta / Delta_new;
      st.Delta = Delta_new;
    }
  }
  const bool loop_again = !inner_break && st.actual <= 0.0 && st.nfev < st.max_nfev;
  if (loop_again) return 0;  // same outer iteration: re-solve with the new radius
  st.decomposed = 0;  // next call runs the outer-loop head (gtol / status / max_nfev checks)
  return st.actual > 0.0 ? 1 : 2;
}

// trf_update_decide + the array part of an accepted step (trf.py:543-567): `gram` is
// [J r]^T [J r] at st.x_new. Returns 1 (another proposal / the final head check is needed).
FB_HD inline int trf_update(TrfState& st, const double* gram) {
  const int n = st.n, ld = n + 1;
  if (trf_update_decide(st, 0.5 * gram[n * ld + n]) == 1) {
    for (int i = 0; i < n; ++i) st.x[i] = st.x_new[i];
    trf_load_accepted(st, gram);  // f = f_new, J = jac(x), g = J^T f
    st.njev += 1;
  }
  return 1;
}

}  // namespace fb
