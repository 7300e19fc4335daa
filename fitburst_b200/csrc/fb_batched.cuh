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
enum FitFlag { kFitStop = 0, kFitFinished = 1, kFitBail = 2, kFitStarted = 3, kFitEvals = 4, kFitSeq = 7, kFitFlagCount = 8 };

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
  // when non-null: the flags are also written to this (page-locked, device-mapped) host slot at the
  // end of the step, so that the host can follow the fit without a device-to-host copy in the
  // stream (copies of several streams share one copy-engine queue and serialise the fits)
  int* host_flags;
  int seq;                 // written last into host_flags[kFitSeq]: the slot belongs to this step
};

__device__ __forceinline__ void publish_fit_flags(const SingleTrfArgs& T) {
  if (T.host_flags == nullptr) return;
  for (int k = 0; k < kFitSeq; ++k) T.host_flags[k] = T.flags[k];
  __threadfence_system();
  *reinterpret_cast<volatile int*>(T.host_flags + kFitSeq) = T.seq;
  __threadfence_system();
}

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
  if (flags[kFitStop]) {
    if (lane == 0) publish_fit_flags(T);
    return;
  }
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
    publish_fit_flags(T);
  }
}

// ---- block-parallel form of the single-fit step ---------------------------------------------------
// k_trf_single above spends ~0.3 ms per step in its one-warp Jacobi (ncu, n = 17): as long as a
// whole evaluation. Here the solver state lives in SHARED memory for the duration of the step
// (the scalar control flow of fb_trf.h runs on it unchanged, thread 0, at shared-memory instead of
// global-memory latency), and the eigen-decomposition is done by the whole block:
//  * all rotations of a round-robin step are applied in ONE pass: with per-index rotation data
//    (c_i, s_i, partner p_i; identity for idle indices) the two-sided update is
//      a'(i,j) = c_i [c_j a(i,j) + s_j a(i,p_j)] + s_i [c_j a(p_i,j) + s_j a(p_i,p_j)],
//    written into a second buffer (one block barrier per step instead of three warp-wide phases);
//  * warm start: consecutive outer iterations have nearly the same J^T J, so the matrix is first
//    rotated into the eigenbasis of the previous iteration (A' = V0^T A V0, almost diagonal) and
//    Jacobi needs 2-3 sweeps instead of ~10; V = V0 W afterwards.
constexpr int kTrfThreads = 128;

__host__ __device__ inline size_t trf_block_smem_bytes(int n) {
  // TrfState + second matrix / eigenvector buffers + per-index rotation data + flags
  return sizeof(TrfState) + sizeof(double) * (2 * (size_t)n * n + 3 * (size_t)kMaxN + 8) + 16;
}

__device__ inline void trf_state_copy(TrfState& dst, const TrfState& src, int n, int tid) {
  for (int e = tid; e < n * n; e += kTrfThreads) {
    dst.A[e] = src.A[e];
    dst.V[e] = src.V[e];
  }
  for (int e = tid; e < n; e += kTrfThreads) {
    dst.x[e] = src.x[e];
    dst.g[e] = src.g[e];
    dst.lam[e] = src.lam[e];
    dst.suf[e] = src.suf[e];
    dst.p[e] = src.p[e];
    dst.x_new[e] = src.x_new[e];
  }
  if (tid == 0) {
    dst.n = src.n;
    dst.m = src.m;
    dst.ftol = src.ftol;
    dst.xtol = src.xtol;
    dst.gtol = src.gtol;
    dst.max_nfev = src.max_nfev;
    dst.cost = src.cost;
    dst.decomposed = src.decomposed;
    dst.Delta = src.Delta;
    dst.alpha = src.alpha;
    dst.predicted = src.predicted;
    dst.step_norm = src.step_norm;
    dst.actual = src.actual;
    dst.nfev = src.nfev;
    dst.njev = src.njev;
    dst.status = src.status;
    dst.finished = src.finished;
    dst.g_norm = src.g_norm;
  }
}

// Eigen-decomposition of the symmetric n x n matrix in a0 (destroyed) by the whole block.
// v0 holds the start basis on entry when `warm` (the eigenvectors of the previous outer
// iteration), the eigenvectors on exit; a1 / v1 are scratch of the same size.
__device__ inline void jacobi_eigh_block(double* a0, double* v0, double* a1, double* v1, double* lam, double* rot,
                                         int* flag, int n, bool warm, int tid) {
  double* rc = rot;                                   // c_i
  double* rs = rot + kMaxN;                           // s_i (signed: see above)
  int* rp = reinterpret_cast<int*>(rot + 2 * kMaxN);  // partner p_i
  if (warm) {
    // a1 = A V0, then a0 = V0^T a1 (upper triangle computed, mirrored: exactly symmetric)
    for (int e = tid; e < n * n; e += kTrfThreads) {
      const int i = e / n, j = e - i * n;
      double acc = 0.0;
      for (int k = 0; k < n; ++k) acc = fma(a0[i * n + k], v0[k * n + j], acc);
      a1[e] = acc;
    }
    __syncthreads();
    for (int e = tid; e < n * n; e += kTrfThreads) {
      const int i = e / n, j = e - i * n;
      if (i > j) continue;
      double acc = 0.0;
      for (int k = 0; k < n; ++k) acc = fma(v0[k * n + i], a1[k * n + j], acc);
      a0[i * n + j] = acc;
      a0[j * n + i] = acc;
    }
  } else {
    for (int e = tid; e < n * n; e += kTrfThreads) v0[e] = (e / n == e % n) ? 1.0 : 0.0;
  }
  __syncthreads();
  const int m = (n + 1) & ~1;  // players of the tournament (one dummy when n is odd)
  const int half = m / 2;
  double* a = a0;
  double* an = a1;
  double* v = v0;
  double* vn = v1;
  for (int sweep = 0; sweep < 60; ++sweep) {
    int rotated = 0;
    for (int step = 0; step < m - 1; ++step) {
      if (tid == 0) *flag = 0;
      for (int e = tid; e < n; e += kTrfThreads) {  // identity for indices without an active pair
        rc[e] = 1.0;
        rs[e] = 0.0;
        rp[e] = e;
      }
      __syncthreads();
      if (tid < half) {
        int x, y;
        if (tid == 0) {
          x = m - 1;
          y = step;
        } else {
          x = (step + tid) % (m - 1);
          y = (step - tid + (m - 1)) % (m - 1);
        }
        const int p = x < y ? x : y, q = x < y ? y : x;
        if (q < n) {
          const double apq = a[p * n + q];
          const double app = a[p * n + p], aqq = a[q * n + q];
          if (apq != 0.0 && !(fabs(apq) <= 1e-17 * sqrt(fabs(app * aqq)))) {
            const double theta = (aqq - app) / (2.0 * apq);
            const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
            rc[p] = c;
            rs[p] = -s;  // column / row p <- c p - s q
            rp[p] = q;
            rc[q] = c;
            rs[q] = s;   // column / row q <- s p + c q
            rp[q] = p;
            *flag = 1;
          }
        }
      }
      __syncthreads();
      if (*flag == 0) continue;  // uniform: every thread reads the same word after the barrier
      rotated = 1;
      for (int e = tid; e < n * n; e += kTrfThreads) {
        const int i = e / n, j = e - i * n;
        const int pi = rp[i], pj = rp[j];
        const double ci = rc[i], si = rs[i], cj = rc[j], sj = rs[j];
        const double top = fma(cj, a[i * n + j], sj * a[i * n + pj]);
        const double bot = fma(cj, a[pi * n + j], sj * a[pi * n + pj]);
        double val = fma(ci, top, si * bot);
        if (pj == i && pi == j && i != j) val = 0.0;  // the annihilated pair of entries
        an[e] = val;
        vn[e] = fma(cj, v[i * n + j], sj * v[i * n + pj]);
      }
      __syncthreads();
      double* t1 = a;
      a = an;
      an = t1;
      double* t2 = v;
      v = vn;
      vn = t2;
    }
    if (!rotated) break;
  }
  // results back into the caller's buffers (a0 / v0) if the ping-pong ended on the other pair
  if (a != a0) {
    for (int e = tid; e < n * n; e += kTrfThreads) {
      a0[e] = a[e];
      v0[e] = v[e];
    }
  }
  __syncthreads();
  if (tid == 0)
    for (int i = 0; i < n; ++i) lam[i] = a0[i * n + i];
  __syncthreads();
  for (int i = 0; i < n - 1; ++i) {  // descending, like svd(J); column swaps spread over the threads
    int best = i;
    for (int j = i + 1; j < n; ++j)
      if (lam[j] > lam[best]) best = j;
    __syncthreads();
    if (best != i) {
      if (tid == 0) {
        const double tmp = lam[i];
        lam[i] = lam[best];
        lam[best] = tmp;
      }
      for (int k = tid; k < n; k += kTrfThreads) {
        const double t2 = v0[k * n + i];
        v0[k * n + i] = v0[k * n + best];
        v0[k * n + best] = t2;
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kTrfThreads) k_trf_block(const SingleTrfArgs T, const int warm_start) {
  extern __shared__ __align__(16) unsigned char trf_smem[];
  const int tid = threadIdx.x;
  int* flags = T.flags;
  if (*reinterpret_cast<volatile int*>(flags + kFitStop)) {
    if (tid == 0) publish_fit_flags(T);
    return;
  }
  const int n = T.n_free, ncols = n + 1;
  TrfState& S = *reinterpret_cast<TrfState*>(trf_smem);
  double* a1 = reinterpret_cast<double*>(trf_smem + sizeof(TrfState));
  double* v1 = a1 + (size_t)n * n;
  double* rot = v1 + (size_t)n * n;
  int* sflag = reinterpret_cast<int*>(rot + 3 * kMaxN);
  __shared__ int s_accepted, s_head, s_warm;
  const double* gram = T.gram;
  const bool started = flags[kFitStarted] != 0;
  if (started) trf_state_copy(S, *T.state, n, tid);
  __syncthreads();
  if (tid == 0) {
    int accepted = 0;
    if (!started) {
      double* x0 = rot;  // scratch
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
      s_warm = 0;  // no previous eigenbasis
    } else {
      const int njev_before = S.njev;
      trf_update(S, gram);
      accepted = S.njev != njev_before;
      s_warm = warm_start;
    }
    flags[kFitEvals] = S.nfev;
    s_accepted = accepted;
    s_head = S.finished ? 0 : trf_head(S);
  }
  __syncthreads();
  const int head = s_head;
  if (s_accepted)
    for (int e = tid; e < ncols * ncols; e += kTrfThreads) T.gram_accepted[e] = gram[e];
  if (head == 1) {
    // S.work holds the copy of J^T J (trf_head); S.V the eigenvectors of the previous iteration
    jacobi_eigh_block(S.work, S.V, a1, v1, S.lam, rot, sflag, n, s_warm != 0, tid);
    if (tid == 0) trf_after_decomposition(S);
    __syncthreads();
  }
  if (tid == 0) {
    if (head != 0) {
      trf_step(S);
      bool ok = true;
      if (T.guard_fast) {
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
  __syncthreads();
  trf_state_copy(*T.state, S, n, tid);  // persistent copy for the next step / a host hand-over
  if (tid == 0) publish_fit_flags(T);
}

// ---- Cholesky form of the single-fit step ---------------------------------------------------------
// Measured (ncu, n = 17): the eigen-decomposition costs 0.3 ms per step with one warp and
// 0.13-0.19 ms with the whole block and a warm start -- a chain of ~170 dependent rotation steps
// of double-precision sqrt / div; the host does it in ~15 us. scipy only needs the decomposition
// to solve (J^T J + alpha I) p = -g for a handful of alpha (More's iteration, common.py:57-168),
// and everything it computes from the singular values has a Cholesky equivalent:
//    p(alpha)    = -(A + alpha I)^-1 g,  A + alpha I = L L^T
//    phi(alpha)  = ||p|| - Delta
//    phi'(alpha) = -sum(suf^2 / (s^2 + alpha)^3) / ||p|| = -||L^-1 p||^2 / ||p||
//    alpha_upper = ||s uf|| / Delta = ||g|| / Delta
//    full rank  <=> A is numerically positive definite (s[-1] > eps m s[0] on the computed spectrum
//                   of J^T J is the same statement up to the noise floor of its smallest eigenvalue)
// i.e. More's original formulation (MINPACK), a 17-step chain per factorisation instead of 170.
// One warp; the state lives in shared memory during the step. A factorisation that breaks down
// at alpha > 0 (never seen; would need J^T J + alpha I indefinite by rounding) hands the fit to
// the host loop like a trial point outside the fast kernels' domain.
__host__ __device__ inline size_t trf_chol_smem_bytes(int n) {
  return sizeof(TrfState) + sizeof(double) * ((size_t)n * n + 4 * (size_t)kMaxN + 8);
}

// L (lower triangle of M, diagonal replaced by its RECIPROCAL) of the n x n matrix src + alpha I.
// Returns false when a pivot is not positive.
__device__ inline bool chol_warp(const double* __restrict__ src, double alpha, double* M, int n, int lane) {
  for (int e = lane; e < n * n; e += 32) M[e] = src[e] + ((e / n == e % n) ? alpha : 0.0);
  __syncwarp();
  bool ok = true;
  for (int k = 0; k < n; ++k) {
    const double piv = M[k * n + k];  // every lane reads the same word
    if (!(piv > 0.0)) {
      ok = false;
      break;
    }
    const double rd = 1.0 / sqrt(piv);
    __syncwarp();
    for (int i = k + lane; i < n; i += 32) M[i * n + k] = (i == k) ? rd : M[i * n + k] * rd;
    __syncwarp();
    for (int i = k + 1 + lane; i < n; i += 32) {
      const double lik = M[i * n + k];
      for (int j = k + 1; j <= i; ++j) M[i * n + j] = fma(-lik, M[j * n + k], M[i * n + j]);
    }
    __syncwarp();
  }
  return ok;
}

// y <- L^-1 y (in place, shared memory)
__device__ inline void chol_forward(const double* M, double* y, int n, int lane) {
  for (int k = 0; k < n; ++k) {
    const double yk = y[k] * M[k * n + k];
    __syncwarp();
    if (lane == 0) y[k] = yk;
    for (int i = k + 1 + lane; i < n; i += 32) y[i] = fma(-M[i * n + k], yk, y[i]);
    __syncwarp();
  }
}

// y <- L^-T y
__device__ inline void chol_backward(const double* M, double* y, int n, int lane) {
  for (int k = n - 1; k >= 0; --k) {
    const double yk = y[k] * M[k * n + k];
    __syncwarp();
    if (lane == 0) y[k] = yk;
    for (int i = lane; i < k; i += 32) y[i] = fma(-M[k * n + i], yk, y[i]);
    __syncwarp();
  }
}

__device__ inline double warp_norm(const double* v, int n, int lane) {
  double s = 0.0;
  for (int i = lane; i < n; i += 32) s = fma(v[i], v[i], s);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  return sqrt(s);
}

// solve_lsq_trust_region (common.py:57-168) on the normal equations: fills S.p, updates S.alpha.
// Returns false on a breakdown at alpha > 0.
__device__ inline bool trust_region_step_chol(TrfState& S, double* M, double* pv, double* wv, int lane) {
  const int n = S.n;
  const double Delta = S.Delta;
  bool full_rank = false;
  if (S.m >= n) full_rank = chol_warp(S.A, 0.0, M, n, lane);
  double p_norm = 0.0;
  if (full_rank) {
    for (int i = lane; i < n; i += 32) pv[i] = S.g[i];
    __syncwarp();
    chol_forward(M, pv, n, lane);
    for (int i = lane; i < n; i += 32) wv[i] = pv[i];  // L^-1 g
    __syncwarp();
    chol_backward(M, pv, n, lane);  // A^-1 g
    p_norm = warp_norm(pv, n, lane);
    if (p_norm <= Delta) {  // Gauss-Newton step inside the region
      for (int i = lane; i < n; i += 32) S.p[i] = -pv[i];
      if (lane == 0) S.alpha = 0.0;
      __syncwarp();
      return true;
    }
  }
  double alpha_upper = warp_norm(S.g, n, lane) / Delta;
  double alpha_lower = 0.0;
  if (full_rank) {
    // phi(0) = ||p|| - Delta, phi'(0) = -||L^-1 p||^2 / ||p||
    for (int i = lane; i < n; i += 32) wv[i] = pv[i];
    __syncwarp();
    chol_forward(M, wv, n, lane);
    const double wn = warp_norm(wv, n, lane);
    const double phi = p_norm - Delta, phi_prime = -(wn * wn) / p_norm;
    alpha_lower = -phi / phi_prime;
  }
  double alpha = S.alpha;
  if (!full_rank && alpha == 0.0) alpha = fmax(0.001 * alpha_upper, sqrt(alpha_lower * alpha_upper));
  for (int it = 0; it < 10; ++it) {
    if (alpha < alpha_lower || alpha > alpha_upper) alpha = fmax(0.001 * alpha_upper, sqrt(alpha_lower * alpha_upper));
    if (!chol_warp(S.A, alpha, M, n, lane)) return false;
    for (int i = lane; i < n; i += 32) pv[i] = S.g[i];
    __syncwarp();
    chol_forward(M, pv, n, lane);
    chol_backward(M, pv, n, lane);
    const double pn = warp_norm(pv, n, lane);
    for (int i = lane; i < n; i += 32) wv[i] = pv[i];
    __syncwarp();
    chol_forward(M, wv, n, lane);
    const double wn = warp_norm(wv, n, lane);
    const double phi = pn - Delta, phi_prime = -(wn * wn) / pn;
    if (phi < 0.0) alpha_upper = alpha;
    const double ratio = phi / phi_prime;
    alpha_lower = fmax(alpha_lower, alpha - ratio);
    alpha -= (phi + Delta) * ratio / Delta;
    if (fabs(phi) < 0.01 * Delta) break;
  }
  if (!chol_warp(S.A, alpha, M, n, lane)) return false;
  for (int i = lane; i < n; i += 32) pv[i] = S.g[i];
  __syncwarp();
  chol_forward(M, pv, n, lane);
  chol_backward(M, pv, n, lane);
  const double pn = warp_norm(pv, n, lane);
  const double scale = Delta / pn;  // "make the norm of p equal to Delta", common.py:163-166
  for (int i = lane; i < n; i += 32) S.p[i] = -pv[i] * scale;
  if (lane == 0) S.alpha = alpha;
  __syncwarp();
  return true;
}

__global__ void __launch_bounds__(32) k_trf_chol(const SingleTrfArgs T) {
  extern __shared__ __align__(16) unsigned char trf_smem[];
  const int lane = threadIdx.x;
  int* flags = T.flags;
  if (*reinterpret_cast<volatile int*>(flags + kFitStop)) {
    if (lane == 0) publish_fit_flags(T);
    return;
  }
  const int n = T.n_free, ncols = n + 1;
  TrfState& S = *reinterpret_cast<TrfState*>(trf_smem);
  double* M = reinterpret_cast<double*>(trf_smem + sizeof(TrfState));
  double* pv = M + (size_t)n * n;
  double* wv = pv + kMaxN;
  double* x0 = wv + kMaxN;
  const double* gram = T.gram;
  const TrfState& G = *T.state;
  const bool started = flags[kFitStarted] != 0;
  if (started) {  // the part of the persistent state a step needs
    for (int e = lane; e < n * n; e += 32) S.A[e] = G.A[e];
    for (int e = lane; e < n; e += 32) {
      S.x[e] = G.x[e];
      S.g[e] = G.g[e];
      S.p[e] = G.p[e];
      S.x_new[e] = G.x_new[e];
    }
    if (lane == 0) {
      S.n = G.n;
      S.m = G.m;
      S.ftol = G.ftol;
      S.xtol = G.xtol;
      S.gtol = G.gtol;
      S.max_nfev = G.max_nfev;
      S.cost = G.cost;
      S.decomposed = G.decomposed;
      S.Delta = G.Delta;
      S.alpha = G.alpha;
      S.predicted = G.predicted;
      S.step_norm = G.step_norm;
      S.actual = G.actual;
      S.nfev = G.nfev;
      S.njev = G.njev;
      S.status = G.status;
      S.finished = G.finished;
      S.g_norm = G.g_norm;
    }
  }
  __syncwarp();
  int action = 0;  // 1: the evaluated point becomes the accepted one
  if (lane == 0) {
    if (!started) {
      single_pack(T, T.params, x0);
      trf_init(S, n, T.num_residuals, x0, T.ftol, T.xtol, T.gtol, T.max_nfev);
      if (!isfinite(gram[n * ncols + n])) {  // "Residuals are not finite in the initial point."
        S.status = -1;
        S.finished = 1;
        S.cost = NAN;
        S.g_norm = 0.0;
        for (int i = 0; i < n; ++i) S.g[i] = 0.0;
      } else {
        S.nfev = 1;
        S.njev = 0;  // counted below with the load
        action = 1;
      }
      flags[kFitStarted] = 1;
    } else {
      action = trf_update_decide(S, 0.5 * gram[n * ncols + n]);
    }
  }
  __syncwarp();
  action = __shfl_sync(0xffffffffu, action, 0);
  if (action == 1) {  // trf_update's array part, spread over the lanes
    if (started)
      for (int i = lane; i < n; i += 32) S.x[i] = S.x_new[i];
    for (int e = lane; e < n * n; e += 32) S.A[e] = gram[(e / n) * ncols + (e % n)];
    for (int i = lane; i < n; i += 32) S.g[i] = gram[i * ncols + n];
    for (int e = lane; e < ncols * ncols; e += 32) T.gram_accepted[e] = gram[e];
    if (lane == 0) {
      S.cost = 0.5 * gram[n * ncols + n];
      S.decomposed = 0;
      S.njev += 1;
    }
  }
  __syncwarp();
  int head = 0;
  if (lane == 0) {
    flags[kFitEvals] = S.nfev;
    head = S.finished ? 0 : trf_head_decide(S);
    if (head == 1) {  // new outer iteration (trf_after_decomposition's bookkeeping)
      S.decomposed = 1;
      S.actual = -1.0;
    }
  }
  __syncwarp();
  head = __shfl_sync(0xffffffffu, head, 0);
  bool ok = true, chol_ok = true;
  if (head != 0) {
    ok = chol_ok = trust_region_step_chol(S, M, pv, wv, lane);
    if (ok) {
      // trf_step's tail: predicted reduction -(0.5 p^T A p + g^T p), trial point, step norm
      double q = 0.0, l = 0.0;
      for (int i = lane; i < n; i += 32) {
        double acc = 0.0;
        for (int j = 0; j < n; ++j) acc = fma(S.A[i * n + j], S.p[j], acc);
        q = fma(S.p[i], acc, q);
        l = fma(S.p[i], S.g[i], l);
        S.x_new[i] = S.x[i] + S.p[i];
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        q += __shfl_xor_sync(0xffffffffu, q, off);
        l += __shfl_xor_sync(0xffffffffu, l, off);
      }
      const double sn = warp_norm(S.p, n, lane);
      if (lane == 0) {
        S.predicted = -(0.5 * q + l);
        S.step_norm = sn;
      }
      __syncwarp();
    }
  }
  if (lane == 0) {
    if (head != 0) {
      if (ok && T.guard_fast) {
        for (int i = 0; i < n; ++i) ok = ok && isfinite(S.x_new[i]);
        if (T.col_of[FB_SCATTERING_TIMESCALE] >= 0) ok = ok && S.x_new[T.col_of[FB_SCATTERING_TIMESCALE]] > 0.0;
        if (T.col_of[FB_BURST_WIDTH] >= 0)
          for (int c = 0; c < T.nc; ++c) ok = ok && S.x_new[T.col_of[FB_BURST_WIDTH] + c] > 0.0;
      }
      if (ok) {
        single_unpack(T, T.params, S.x_new);
      } else {
        // hand-over to the host loop: 1 = x_new is proposed but outside the fast kernels'
        // domain, 2 = the factorisation broke down before a proposal was made
        flags[kFitBail] = chol_ok ? 1 : 2;
        __threadfence();
        flags[kFitStop] = 1;
      }
    } else {
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
  __syncwarp();
  // persistent copy for the next step / a host hand-over
  TrfState& O = *T.state;
  for (int e = lane; e < n * n; e += 32) O.A[e] = S.A[e];
  for (int e = lane; e < n; e += 32) {
    O.x[e] = S.x[e];
    O.g[e] = S.g[e];
    O.p[e] = S.p[e];
    O.x_new[e] = S.x_new[e];
  }
  if (lane == 0) {
    O.n = S.n;
    O.m = S.m;
    O.ftol = S.ftol;
    O.xtol = S.xtol;
    O.gtol = S.gtol;
    O.max_nfev = S.max_nfev;
    O.cost = S.cost;
    O.decomposed = S.decomposed;
    O.Delta = S.Delta;
    O.alpha = S.alpha;
    O.predicted = S.predicted;
    O.step_norm = S.step_norm;
    O.actual = S.actual;
    O.nfev = S.nfev;
    O.njev = S.njev;
    O.status = S.status;
    O.finished = S.finished;
    O.g_norm = S.g_norm;
    publish_fit_flags(T);
  }
}

}  // namespace fb
