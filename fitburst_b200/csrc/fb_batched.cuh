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

}  // namespace fb
