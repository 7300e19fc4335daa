// On-device batched trust-region solver: `count` independent bursts of identical shape are fitted
// concurrently. Per iteration the whole batch costs three launches -- k_tables_batched,
// k_eval_batched<MODE_GRAM> (blockIdx.y = burst) and k_trf_batched (one warp per burst) -- and the
// host only polls one counter; accept/reject, radius update, termination tests and the next trial
// point never leave the GPU. The state machine is the same fb_trf.h code the single-fit host
// driver runs (scipy trf_no_bounds restated); the n x n eigen-decomposition is done by the warp
// cooperatively in shared memory.
#pragma once
#include "fb_kernels.cuh"
#include "fb_trf.h"

namespace fb {

// Cyclic Jacobi of fb_trf.h with the k-loops of each rotation spread over the 32 lanes; the
// arithmetic per element is the same, so the result equals the serial routine.
__device__ inline void jacobi_eigh_warp(double* a, double* v, double* lam, int n, int lane) {
  for (int e = lane; e < n * n; e += 32) v[e] = (e / n == e % n) ? 1.0 : 0.0;
  __syncwarp();
  for (int sweep = 0; sweep < 60; ++sweep) {
    int rotated = 0;
    for (int p = 0; p < n - 1; ++p) {
      for (int q = p + 1; q < n; ++q) {
        const double apq = a[p * n + q];
        if (apq == 0.0) continue;
        const double app = a[p * n + p], aqq = a[q * n + q];
        if (fabs(apq) <= 1e-17 * sqrt(fabs(app * aqq))) continue;
        rotated = 1;
        const double theta = (aqq - app) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        __syncwarp();
        for (int k = lane; k < n; k += 32) {
          const double akp = a[k * n + p], akq = a[k * n + q];
          a[k * n + p] = c * akp - s * akq;
          a[k * n + q] = s * akp + c * akq;
        }
        __syncwarp();
        for (int k = lane; k < n; k += 32) {
          const double apk = a[p * n + k], aqk = a[q * n + k];
          a[p * n + k] = c * apk - s * aqk;
          a[q * n + k] = s * apk + c * aqk;
        }
        __syncwarp();
        if (lane == 0) {
          a[p * n + q] = 0.0;
          a[q * n + p] = 0.0;
        }
        for (int k = lane; k < n; k += 32) {
          const double vkp = v[k * n + p], vkq = v[k * n + q];
          v[k * n + p] = c * vkp - s * vkq;
          v[k * n + q] = s * vkp + c * vkq;
        }
        __syncwarp();
      }
    }
    if (!rotated) break;
  }
  __syncwarp();
  if (lane == 0) {
    for (int i = 0; i < n; ++i) lam[i] = a[i * n + i];
    for (int i = 0; i < n - 1; ++i) {  // descending, like svd(J)
      int best = i;
      for (int j = i + 1; j < n; ++j)
        if (lam[j] > lam[best]) best = j;
      if (best != i) {
        const double tmp = lam[i];
        lam[i] = lam[best];
        lam[best] = tmp;
        for (int k = 0; k < n; ++k) {
          const double t2 = v[k * n + i];
          v[k * n + i] = v[k * n + best];
          v[k * n + best] = t2;
        }
      }
    }
  }
  __syncwarp();
}

struct BatchedTrfArgs {
  int count, nc, n_free;
  int col_of[9];
  long long num_residuals;
  double ftol, xtol, gtol;
  int max_nfev;
  const double* partials;  // [count][blocks][ntiles*16]
  int blocks;              // eval blocks per burst
  double* gram;            // [count][(n+1)^2]
  double* params;          // [count][10*nc]
  TrfState* states;        // [count]
  int* started;            // [count]
  int* done;               // [count]
  int* num_done;
  fb_fit_result* results;  // [count]
};

__device__ inline void batched_pack(const BatchedTrfArgs& T, const double* params, double* x) {
  for (int k = 0; k < 9; ++k) {
    if (T.col_of[k] < 0) continue;
    if (is_fit_global(k)) x[T.col_of[k]] = params[k * T.nc];
    else
      for (int c = 0; c < T.nc; ++c) x[T.col_of[k] + c] = params[k * T.nc + c];
  }
}

__device__ inline void batched_unpack(const BatchedTrfArgs& T, double* params, const double* x) {
  for (int k = 0; k < 9; ++k) {
    if (T.col_of[k] < 0) continue;
    for (int c = 0; c < T.nc; ++c)
      params[k * T.nc + c] = is_fit_global(k) ? x[T.col_of[k]] : x[T.col_of[k] + c];
  }
}

// One warp per burst: finish the Gram reduction, advance the trust-region state machine,
// publish the next trial point (or the final result).
__global__ void __launch_bounds__(32) k_trf_batched(const BatchedTrfArgs T) {
  extern __shared__ double sm[];
  const int b = blockIdx.x, lane = threadIdx.x;
  if (T.done[b]) return;
  const int n = T.n_free, ncols = n + 1;
  double* gram = T.gram + (size_t)b * ncols * ncols;
  double* params = T.params + (size_t)b * FB_NUM_PARAMETERS * T.nc;
  TrfState& S = T.states[b];
  // 1. second stage of the deterministic Gram reduction (block order), symmetric fill
  {
    const int side = gram_side(ncols), ntiles = gram_ntiles(ncols);
    constexpr int kk = kTile * kTile;
    const double* part = T.partials + (size_t)b * T.blocks * ntiles * kk;
    for (int e = lane; e < ntiles * kk; e += 32) {
      double s = 0.0;
      for (int blk = 0; blk < T.blocks; ++blk) s += part[(size_t)blk * ntiles * kk + e];
      const int t = e / kk, r = (e % kk) / kTile, q = e % kTile;
      int ti, tj;
      tile_coords(t, side, ti, tj);
      const int gr = ti * kTile + r, gc = tj * kTile + q;
      if (gr < ncols && gc < ncols && (ti != tj || gr <= gc)) {
        gram[gr * ncols + gc] = s;
        gram[gc * ncols + gr] = s;
      }
    }
  }
  __syncwarp();
  __threadfence_block();
  // 2. state machine
  int head = 0;
  if (lane == 0) {
    if (!T.started[b]) {
      double x0[kMaxN];
      batched_pack(T, params, x0);
      trf_init(S, n, T.num_residuals, x0, T.ftol, T.xtol, T.gtol, T.max_nfev);
      if (!isfinite(gram[n * ncols + n])) {  // "Residuals are not finite in the initial point."
        S.status = -1;
        S.finished = 1;
        S.cost = NAN;  // nothing was accepted: no garbage from the cudaMalloc'ed state in the result
        S.g_norm = 0.0;
        for (int i = 0; i < n; ++i) S.g[i] = 0.0;
      } else {
        trf_load_accepted(S, gram);
        S.nfev = 1;
        S.njev = 1;
      }
      T.started[b] = 1;
    } else {
      trf_update(S, gram);
    }
    head = S.finished ? 0 : trf_head(S);
  }
  head = __shfl_sync(0xffffffffu, head, 0);
  if (head == 1) {
    double* a = sm;
    double* v = sm + n * n;
    for (int e = lane; e < n * n; e += 32) a[e] = S.work[e];
    __syncwarp();
    jacobi_eigh_warp(a, v, S.lam, n, lane);
    for (int e = lane; e < n * n; e += 32) S.V[e] = v[e];
    __syncwarp();
    __threadfence_block();
    if (lane == 0) trf_after_decomposition(S);
  }
  if (lane == 0) {
    if (head != 0) {
      trf_step(S);
      batched_unpack(T, params, S.x_new);
    } else {
      // finished: leave the best fit in the parameter block and publish the scalar results
      batched_unpack(T, params, S.x);
      fb_fit_result& R = T.results[b];
      R.status = S.status;
      R.nfev = S.nfev;
      R.njev = S.njev;
      R.num_free = n;
      R.cost = S.cost;
      R.optimality = S.g_norm;
      for (int i = 0; i < n; ++i) {
        R.x[i] = S.x[i];
        R.grad[i] = S.g[i];
      }
      T.done[b] = 1;
      atomicAdd(T.num_done, 1);
    }
  }
}

// ---- single-fit device driver --------------------------------------------------------------------
// One full-size fit keeps its trust-region loop on the GPU as well: after every evaluation
// (k_tables_group -> k_vals3 -> k_gram3 -> k_gram3_finish, or the general kernels) ONE warp runs
// the scipy state machine on the finished (n+1)^2 matrix and writes the next trial point straight
// into the plan's parameter block, so the host can enqueue evaluation after evaluation without
// ever waiting for one (fb_fit: chunks of launches, one flag read-back per chunk). Every kernel
// of an evaluation returns at once when flags[kFitStop] is set, so launches enqueued past the end
// of the fit cost a few microseconds.
enum FitFlag { kFitStop = 0, kFitFinished = 1, kFitBail = 2, kFitStarted = 3, kFitEvals = 4, kFitFlagCount = 8 };

struct SingleTrfArgs {
  int nc, n_free;
  int col_of[9];
  long long num_residuals;
  double ftol, xtol, gtol;
  int max_nfev;
  int guard_fast;          // trial points must stay inside the domain of the fast kernels
  const double* gram;      // [J r]^T [J r] of the evaluation that just finished, (n+1)^2
  double* gram_accepted;   // copy of the matrix at the accepted point (hessian_approx)
  double* params;          // the plan's (FB_NUM_PARAMETERS, nc) block: next trial point goes here
  TrfState* state;
  int* flags;              // FitFlag
  fb_fit_result* result;
};

// Two-sided Jacobi with the rotations of one round-robin step applied together: the n/2 pivot
// pairs of a step are disjoint, their 2x2 diagonal blocks are untouched by each other's rotations,
// so all angles come from the matrix at the start of the step and the column / row / eigenvector
// updates are spread over the lanes (the one-rotation-at-a-time order of jacobi_eigh_warp costs
// ~140 dependent rotations per sweep: 0.3 ms for n = 17, as long as a whole evaluation).
// Same rotation formulas and convergence criterion as jacobi_eigh (fb_trf.h); eigenvalues sorted
// descending like svd(J). a, v: n x n in shared memory; rot: 4 * 32 doubles of scratch.
__device__ inline void jacobi_eigh_warp_rr(double* a, double* v, double* lam, double* rot, int n, int lane) {
  for (int e = lane; e < n * n; e += 32) v[e] = (e / n == e % n) ? 1.0 : 0.0;
  __syncwarp();
  const int m = (n + 1) & ~1;          // players of the tournament (one dummy when n is odd)
  const int half = m / 2;
  double* rc = rot;                    // cos of pair k
  double* rs = rot + 32;               // sin
  int* rp = reinterpret_cast<int*>(rot + 64);  // p of pair k
  int* rq = rp + 32;                           // q
  for (int sweep = 0; sweep < 60; ++sweep) {
    int rotated = 0;
    for (int step = 0; step < m - 1; ++step) {
      for (int kbase = 0; kbase < half; kbase += 32) {  // n <= 67: at most two passes
        const int k = kbase + lane;
        const int slot = lane;
        bool act = false;
        double c = 1.0, s = 0.0;
        int pp = 0, qq = 0;
        if (k < half) {
          int x, y;
          if (k == 0) {
            x = m - 1;
            y = step;
          } else {
            x = (step + k) % (m - 1);
            y = (step - k + (m - 1)) % (m - 1);
          }
          pp = x < y ? x : y;
          qq = x < y ? y : x;
          if (qq < n) {
            const double apq = a[pp * n + qq];
            const double app = a[pp * n + pp], aqq = a[qq * n + qq];
            if (apq != 0.0 && !(fabs(apq) <= 1e-17 * sqrt(fabs(app * aqq)))) {
              act = true;
              const double theta = (aqq - app) / (2.0 * apq);
              const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
              c = 1.0 / sqrt(t * t + 1.0);
              s = t * c;
            }
          }
        }
        rc[slot] = c;
        rs[slot] = s;
        rp[slot] = act ? pp : -1;
        rq[slot] = qq;
        const unsigned any = __ballot_sync(0xffffffffu, act);
        __syncwarp();
        if (any == 0u) continue;
        rotated = 1;
        const int np = min(32, half - kbase);
        // A <- A P (columns p, q of every row)
        for (int e = lane; e < n * np; e += 32) {
          const int r = e / np, kk = e - r * np, p = rp[kk], q = rq[kk];
          if (p < 0) continue;
          const double cc = rc[kk], ss = rs[kk];
          const double akp = a[r * n + p], akq = a[r * n + q];
          a[r * n + p] = cc * akp - ss * akq;
          a[r * n + q] = ss * akp + cc * akq;
        }
        __syncwarp();
        // A <- P^T A (rows p, q of every column)
        for (int e = lane; e < n * np; e += 32) {
          const int col = e / np, kk = e - col * np, p = rp[kk], q = rq[kk];
          if (p < 0) continue;
          const double cc = rc[kk], ss = rs[kk];
          const double apk = a[p * n + col], aqk = a[q * n + col];
          a[p * n + col] = cc * apk - ss * aqk;
          a[q * n + col] = ss * apk + cc * aqk;
        }
        __syncwarp();
        if (lane < np && rp[lane] >= 0) {
          a[rp[lane] * n + rq[lane]] = 0.0;
          a[rq[lane] * n + rp[lane]] = 0.0;
        }
        // V <- V P
        for (int e = lane; e < n * np; e += 32) {
          const int r = e / np, kk = e - r * np, p = rp[kk], q = rq[kk];
          if (p < 0) continue;
          const double cc = rc[kk], ss = rs[kk];
          const double vkp = v[r * n + p], vkq = v[r * n + q];
          v[r * n + p] = cc * vkp - ss * vkq;
          v[r * n + q] = ss * vkp + cc * vkq;
        }
        __syncwarp();
      }
    }
    if (!rotated) break;
  }
  __syncwarp();
  if (lane == 0)
    for (int i = 0; i < n; ++i) lam[i] = a[i * n + i];
  __syncwarp();
  for (int i = 0; i < n - 1; ++i) {  // descending, like svd(J); column swaps spread over the lanes
    int best = i;
    for (int j = i + 1; j < n; ++j)
      if (lam[j] > lam[best]) best = j;
    __syncwarp();
    if (best != i) {
      if (lane == 0) {
        const double tmp = lam[i];
        lam[i] = lam[best];
        lam[best] = tmp;
      }
      for (int k = lane; k < n; k += 32) {
        const double t2 = v[k * n + i];
        v[k * n + i] = v[k * n + best];
        v[k * n + best] = t2;
      }
    }
    __syncwarp();
  }
}

__device__ inline void single_pack(const SingleTrfArgs& T, const double* params, double* x) {
  for (int k = 0; k < 9; ++k) {
    if (T.col_of[k] < 0) continue;
    if (is_fit_global(k)) x[T.col_of[k]] = params[k * T.nc];
    else
      for (int c = 0; c < T.nc; ++c) x[T.col_of[k] + c] = params[k * T.nc + c];
  }
}

__device__ inline void single_unpack(const SingleTrfArgs& T, double* params, const double* x) {
  for (int k = 0; k < 9; ++k) {
    if (T.col_of[k] < 0) continue;
    for (int c = 0; c < T.nc; ++c)
      params[k * T.nc + c] = is_fit_global(k) ? x[T.col_of[k]] : x[T.col_of[k] + c];
  }
}

// smem: 2 n^2 doubles (matrix + eigenvectors) + n eigenvalues + 128 doubles of rotation scratch
__host__ __device__ inline size_t trf_single_smem_bytes(int n) { return sizeof(double) * (2 * (size_t)n * n + n + 128); }

__global__ void __launch_bounds__(32) k_trf_single(const SingleTrfArgs T) {
  extern __shared__ double sm[];
  const int lane = threadIdx.x;
  int* flags = T.flags;
  if (flags[kFitStop]) return;
  const int n = T.n_free, ncols = n + 1;
  TrfState& S = *T.state;
  const double* gram = T.gram;
  int accepted = 0, head = 0;
  if (lane == 0) {
    if (!flags[kFitStarted]) {
      double x0[kMaxN];
      single_pack(T, T.params, x0);
      trf_init(S, n, T.num_residuals, x0, T.ftol, T.xtol, T.gtol, T.max_nfev);
      if (!isfinite(gram[n * ncols + n])) {  // "Residuals are not finite in the initial point."
        S.status = -1;
        S.finished = 1;
        S.cost = NAN;
        S.g_norm = 0.0;
        for (int i = 0; i < n; ++i) S.g[i] = 0.0;
      } else {
        trf_load_accepted(S, gram);
        S.nfev = 1;
        S.njev = 1;
        accepted = 1;
      }
      flags[kFitStarted] = 1;
    } else {
      const int njev_before = S.njev;
      trf_update(S, gram);
      accepted = S.njev != njev_before;
    }
    flags[kFitEvals] = S.nfev;
    head = S.finished ? 0 : trf_head(S);
  }
  __syncwarp();  // lane 0's writes to the state are visible to the warp
  accepted = __shfl_sync(0xffffffffu, accepted, 0);
  head = __shfl_sync(0xffffffffu, head, 0);
  if (accepted)
    for (int e = lane; e < ncols * ncols; e += 32) T.gram_accepted[e] = gram[e];
  if (head == 1) {
    double* a = sm;
    double* v = sm + n * n;
    double* lam = v + n * n;
    double* rot = lam + n;
    for (int e = lane; e < n * n; e += 32) a[e] = S.work[e];
    __syncwarp();
    jacobi_eigh_warp_rr(a, v, lam, rot, n, lane);
    for (int e = lane; e < n * n; e += 32) S.V[e] = v[e];
    for (int e = lane; e < n; e += 32) S.lam[e] = lam[e];
    __syncwarp();
    __threadfence_block();
    if (lane == 0) trf_after_decomposition(S);
  }
  if (lane == 0) {
    if (head != 0) {
      trf_step(S);
      bool ok = true;
      if (T.guard_fast) {
        // the fast kernels cover scattering_timescale > 0 and burst_width > 0 only; a trial point
        // outside (or a non-finite one) is handed back to the host-driven loop, which continues
        // from this state on the general kernels
        for (int i = 0; i < n; ++i) ok = ok && isfinite(S.x_new[i]);
        if (T.col_of[FB_SCATTERING_TIMESCALE] >= 0) ok = ok && S.x_new[T.col_of[FB_SCATTERING_TIMESCALE]] > 0.0;
        if (T.col_of[FB_BURST_WIDTH] >= 0)
          for (int c = 0; c < T.nc; ++c) ok = ok && S.x_new[T.col_of[FB_BURST_WIDTH] + c] > 0.0;
      }
      if (ok) {
        single_unpack(T, T.params, S.x_new);
      } else {
        flags[kFitBail] = 1;
        __threadfence();
        flags[kFitStop] = 1;
      }
    } else {
      // finished: the parameter block stays at the LAST EVALUATED point (analysis/fitter.py:258)
      fb_fit_result& R = *T.result;
      R.status = S.status;
      R.nfev = S.nfev;
      R.njev = S.njev;
      R.num_free = n;
      R.cost = S.cost;
      R.optimality = S.g_norm;
      for (int i = 0; i < n; ++i) {
        R.x[i] = S.x[i];
        R.grad[i] = S.g[i];
      }
      flags[kFitFinished] = 1;
      __threadfence();
      flags[kFitStop] = 1;
    }
  }
}

}  // namespace fb
