// Pre-fit data preparation on the device: the per-channel reductions,
// masking, block-mean down-sampling and de-dispersed windowing of
// utilities/bases.py:129-336 (ReaderBaseClass.preprocess_data / downsample / window_data).
// All of it is HBM-bound streaming over a (num_freq, num_time) f64 spectrum: one block per
// channel row (or per output row), coalesced 8-byte accesses, fixed-order tree reductions.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "fb_kernels.cuh"  // solve_small

namespace fb {

constexpr int kPreBlock = 256;

__device__ __forceinline__ double pre_block_sum(double v, double* red) {
  const int tid = threadIdx.x;
  red[tid] = v;
  __syncthreads();
  for (int off = kPreBlock / 2; off > 0; off >>= 1) {
    if (tid < off) red[tid] += red[tid + off];
    __syncthreads();
  }
  const double out = red[0];
  __syncthreads();
  return out;
}

// Per channel: sum(w), sum(d w), min(d), max(d)      (utilities/bases.py:251-264)
__global__ void __launch_bounds__(kPreBlock) k_pre_moments(const double* __restrict__ data,
                                                          const double* __restrict__ weights, int nt,
                                                          double* __restrict__ out) {
  __shared__ double red[kPreBlock];
  const int i = blockIdx.x, tid = threadIdx.x;
  const double* d = data + (size_t)i * nt;
  const double* w = weights + (size_t)i * nt;
  double sw = 0.0, sdw = 0.0, mn = INFINITY, mx = -INFINITY;
  for (int j = tid; j < nt; j += kPreBlock) {
    const double dv = d[j], wv = w[j];
    sw += wv;
    sdw = fma(dv, wv, sdw);
    mn = fmin(mn, dv);
    mx = fmax(mx, dv);
  }
  const double tsw = pre_block_sum(sw, red);
  const double tsdw = pre_block_sum(sdw, red);
  red[tid] = mn;
  __syncthreads();
  for (int off = kPreBlock / 2; off > 0; off >>= 1) {
    if (tid < off) red[tid] = fmin(red[tid], red[tid + off]);
    __syncthreads();
  }
  const double tmn = red[0];
  __syncthreads();
  red[tid] = mx;
  __syncthreads();
  for (int off = kPreBlock / 2; off > 0; off >>= 1) {
    if (tid < off) red[tid] = fmax(red[tid], red[tid + off]);
    __syncthreads();
  }
  if (tid == 0) {
    out[4 * i + 0] = tsw;
    out[4 * i + 1] = tsdw;
    out[4 * i + 2] = tmn;
    out[4 * i + 3] = red[0];
  }
}

// Baseline removal (utilities/bases.py:270-273): good channels d <- d / mean - 1, then every
// sample whose weight is zero is zeroed (all channels).
__global__ void __launch_bounds__(kPreBlock) k_pre_baseline(double* __restrict__ data,
                                                           const double* __restrict__ weights,
                                                           const uint8_t* __restrict__ good,
                                                           const double* __restrict__ mean, int nt) {
  const int i = blockIdx.x;
  double* d = data + (size_t)i * nt;
  const double* w = weights + (size_t)i * nt;
  const bool g = good[i] != 0;
  const double m = mean[i];
  for (int j = threadIdx.x; j < nt; j += kPreBlock) {
    double v = d[j];
    if (g) v = v / m - 1.0;
    if (w[j] == 0.0) v = 0.0;
    d[j] = v;
  }
}

// Per channel: sum(d^2), sum(d^3)                    (utilities/bases.py:276-279)
__global__ void __launch_bounds__(kPreBlock) k_pre_power_sums(const double* __restrict__ data, int nt,
                                                             double* __restrict__ out) {
  __shared__ double red[kPreBlock];
  const int i = blockIdx.x;
  const double* d = data + (size_t)i * nt;
  double s2 = 0.0, s3 = 0.0;
  for (int j = threadIdx.x; j < nt; j += kPreBlock) {
    const double v = d[j], v2 = v * v;
    s2 += v2;
    s3 = fma(v2, v, s3);
  }
  const double t2 = pre_block_sum(s2, red);
  const double t3 = pre_block_sum(s3, red);
  if (threadIdx.x == 0) {
    out[2 * i + 0] = t2;
    out[2 * i + 1] = t3;
  }
}

// data[i, :] = 0 where keep[i] == 0                  (utilities/bases.py:296, 168)
__global__ void __launch_bounds__(kPreBlock) k_pre_zero_channels(double* __restrict__ data,
                                                                const uint8_t* __restrict__ keep, int nt) {
  const int i = blockIdx.x;
  if (keep[i]) return;
  double* d = data + (size_t)i * nt;
  for (int j = threadIdx.x; j < nt; j += kPreBlock) d[j] = 0.0;
}

// downsample_2d (routines/manipulate.py:43-72): mean over factor_time (inner), then over
// factor_freq, of the (nf/ff, ff, nt/ft, ft) reshape. One thread per output sample.
__global__ void __launch_bounds__(kPreBlock) k_pre_downsample(const double* __restrict__ in, int nt_in, int ff,
                                                             int ft, int nt_out, double* __restrict__ out) {
  const int io = blockIdx.y;
  const int jo = blockIdx.x * kPreBlock + threadIdx.x;
  if (jo >= nt_out) return;
  double acc = 0.0;
  for (int a = 0; a < ff; ++a) {
    const double* row = in + (size_t)(io * ff + a) * nt_in + (size_t)jo * ft;
    double s = 0.0;
    for (int b = 0; b < ft; ++b) s += row[b];
    acc += s / (double)ft;
  }
  out[(size_t)io * nt_out + jo] = acc / (double)ff;
}

// window_data (utilities/bases.py:300-336): out[i, k] = data[i, start[i] + k], k < width.
__global__ void __launch_bounds__(kPreBlock) k_pre_window(const double* __restrict__ in, int nt,
                                                         const int* __restrict__ start, int width,
                                                         double* __restrict__ out) {
  const int i = blockIdx.x;
  const double* src = in + (size_t)i * nt + start[i];
  double* dst = out + (size_t)i * width;
  for (int k = threadIdx.x; k < width; k += kPreBlock) dst[k] = src[k];
}

// Band-averaged time series, FindPeak.find_peak: data.mean(axis=0) (analysis/peak_finder.py:80).
// Stage 1: block (column tile, row chunk) adds its rows in row order, threads along the columns
// (coalesced); stage 2 adds the chunk partials in chunk order and divides: fixed order, deterministic.
constexpr int kPreRowChunk = 64;

__global__ void __launch_bounds__(kPreBlock) k_pre_colsum_partial(const double* __restrict__ data, int nf, int nt,
                                                                 double* __restrict__ partial) {
  const int j = blockIdx.x * kPreBlock + threadIdx.x;
  if (j >= nt) return;
  const int r0 = blockIdx.y * kPreRowChunk, r1 = min(nf, r0 + kPreRowChunk);
  double s = 0.0;
  for (int i = r0; i < r1; ++i) s += data[(size_t)i * nt + j];
  partial[(size_t)blockIdx.y * nt + j] = s;
}

__global__ void __launch_bounds__(kPreBlock) k_pre_colmean_final(const double* __restrict__ partial, int chunks,
                                                                int nf, int nt, double* __restrict__ out) {
  const int j = blockIdx.x * kPreBlock + threadIdx.x;
  if (j >= nt) return;
  double s = 0.0;
  for (int c = 0; c < chunks; ++c) s += partial[(size_t)c * nt + j];
  out[j] = s / (double)nf;
}


// ---- upload of the usable channels of a host spectrum -----------------------------------------
// Copies the rows listed in rows_dev from a PAGE-LOCKED host array (read directly over PCIe:
// pinned memory is device-accessible under unified addressing) into the device spectrum and
// widens float32 input on the way. Flagged channels (analysis/fitter.py:506-542: weight 0) are
// never read by the fit, so their rows are not transferred: at 35 % flagged channels a third of
// the host-to-device traffic, which is what bounds the end-to-end rate of eight GPUs on one host.
template <typename T>
__global__ void __launch_bounds__(256) k_upload_rows(const T* __restrict__ host, const int* __restrict__ rows_dev,
                                                     int num_rows, int nt, double* __restrict__ dst) {
  for (int r = blockIdx.x; r < num_rows; r += gridDim.x) {
    const int i = rows_dev[r];
    const T* src = host + (size_t)i * nt;
    double* out = dst + (size_t)i * nt;
    if (sizeof(T) == 8) {
      const bool vec = (nt % 2 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
      if (vec) {
        const double2* s2 = reinterpret_cast<const double2*>(src);
        double2* o2 = reinterpret_cast<double2*>(out);
        for (int j = threadIdx.x; j < nt / 2; j += blockDim.x) o2[j] = s2[j];
      } else {
        for (int j = threadIdx.x; j < nt; j += blockDim.x) out[j] = (double)src[j];
      }
    } else {
      const bool vec = (nt % 4 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
      if (vec) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
        for (int j = threadIdx.x; j < nt / 4; j += blockDim.x) {
          const float4 v = s4[j];
          reinterpret_cast<double2*>(out)[2 * j] = make_double2((double)v.x, (double)v.y);
          reinterpret_cast<double2*>(out)[2 * j + 1] = make_double2((double)v.z, (double)v.w);
        }
      } else {
        for (int j = threadIdx.x; j < nt; j += blockDim.x) out[j] = (double)src[j];
      }
    }
  }
}

// ---- small public routines of the reference (routines/spectrum.py, routines/ism.py) ---------------
// Elementwise over an array of frequencies, same expression order as the reference.
// routines/spectrum.py:39-41: exp(sp_idx * log(f / ref) + sp_run * log(f / ref) ** 2)
__global__ void __launch_bounds__(256) k_spectrum_rpl(const double* __restrict__ freqs, long long count, double ref,
                                                      double sp_idx, double sp_run, double* __restrict__ out) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += (long long)gridDim.x * blockDim.x) {
    const double lf = log(freqs[e] / ref);
    out[e] = exp(sp_idx * lf + sp_run * (lf * lf));
  }
}

// kind 0: compute_time_dm_delay   a * b * (f ** c - d ** c)          (routines/ism.py:78)
// kind 1: compute_time_dm_smear   2 * a * b * d * f ** (c - 1)       (routines/ism.py:110)
// kind 2: compute_time_scattering b * (f / a) ** c                   (routines/ism.py:139)
__global__ void __launch_bounds__(256) k_ism_times(int kind, const double* __restrict__ freqs, long long count, double a,
                                                   double b, double c, double d, double* __restrict__ out) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += (long long)gridDim.x * blockDim.x) {
    const double f = freqs[e];
    double v;
    if (kind == 0) v = a * b * (pow(f, c) - pow(d, c));
    else if (kind == 1) v = 2.0 * a * b * d * pow(f, c - 1.0);
    else v = b * pow(f / a, c);
    out[e] = v;
  }
}

// float32 -> float64 widening of an uploaded spectrum (exact), 16-byte loads / 32-byte stores
__global__ void __launch_bounds__(256) k_widen_f32(const float* __restrict__ src, long long count, double* __restrict__ dst) {
  const long long quads = count >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += stride) {
    const float4 v = __ldcs(reinterpret_cast<const float4*>(src) + q);
    double2* o = reinterpret_cast<double2*>(dst) + 2 * q;
    o[0] = make_double2((double)v.x, (double)v.y);
    o[1] = make_double2((double)v.z, (double)v.w);
  }
  for (long long e = 4 * quads + (long long)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += stride) dst[e] = (double)src[e];
}

// compute_amplitude_per_channel (routines/ism.py:11-48) for ONE channel: G = P^T P, b = P^T d over
// the num_time samples (P is (num_time, num_comp) C order), then solve G a = b with partial
// pivoting like LAPACK gesv. One block; fixed-order tree sums. err_flag is set on a zero pivot.
__global__ void __launch_bounds__(256) k_amplitude_per_channel(const double* __restrict__ data, const double* __restrict__ model,
                                                               int num_time, int nc, double* __restrict__ amplitudes,
                                                               int* __restrict__ err_flag) {
  extern __shared__ double amp_sm[];  // [256] reduction scratch, then G (nc*nc) and b (nc)
  double* G = amp_sm + 256;
  double* b = G + nc * nc;
  const int tid = threadIdx.x;
  for (int e = 0; e < nc * (nc + 1); ++e) {
    const int r = e / (nc + 1), q = e % (nc + 1);
    if (q < nc && q < r) continue;  // symmetric: lower triangle copied below
    double acc = 0.0;
    for (int j = tid; j < num_time; j += 256) {
      const double pr = model[(size_t)j * nc + r];
      acc = fma(pr, q < nc ? model[(size_t)j * nc + q] : data[j], acc);
    }
    amp_sm[tid] = acc;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
      if (tid < off) amp_sm[tid] += amp_sm[tid + off];
      __syncthreads();
    }
    if (tid == 0) {
      if (q < nc) {
        G[r * nc + q] = amp_sm[0];
        G[q * nc + r] = amp_sm[0];
      } else {
        b[r] = amp_sm[0];
      }
    }
    __syncthreads();
  }
  if (tid == 0) {
    const bool ok = solve_small(G, b, nc, nc);
    for (int c = 0; c < nc; ++c) amplitudes[c] = ok ? b[c] : nan("");
    if (!ok && err_flag) atomicExch(err_flag, 1);
  }
}

}  // namespace fb
