// Exact chi^2 Hessian: the 45 mixed partials of routines/derivative.py:938-2836 evaluated per
// (sample, component) from the same three quantities the first derivatives use (spectrum,
// amplitude, time difference), contracted on the fly with the weighted residual so that no
// (Nf, Nt) derivative map is ever stored:   S_ab = sum (data - model) w^2 d2m/dadb,
//   H = 2 J^T J - 2 S   (LSFitter.compute_hessian, analysis/fitter.py:81-178).
// The reference's quirks are reproduced deliberately (they define bestfit_uncertainties); each
// is flagged "quirk" with its line.
#pragma once
#include "fb_kernels.cuh"
#include "fb_fast.cuh"

namespace fb {

struct D2Ctx {
  double M, T, w, tau, tau0, L, apbf, D, Dp, G, arg_erf;
  double it, iw, iw2, it0, ws, iap;  // 1/tau, 1/w, 1/w^2, 1/tau0, w/tau, 1/amp_pbf
  double d1[9];
  double fpow, ref_freq, q1, q2, k, dm;
  double dap_alpha, d2ap_alpha;
  double sum_dmidx;   // component-summed d model / d dm_index (quirk: derivative.py:2110, 2266)
  double apbf_outer;  // amp_pbf of the function's `component` argument (quirk: derivative.py:2807)
};

__host__ __device__ inline bool is_scalar_row(int p) {
  return p == FB_AMPLITUDE || p == FB_SPECTRAL_INDEX || p == FB_SPECTRAL_RUNNING;
}

// Functions of the reference that return the SUM over components whatever `component` is.
__host__ __device__ inline bool d2_is_sum(int pa, int pb) {
  if (pa > pb) {
    const int t = pa;
    pa = pb;
    pb = t;
  }
  if (is_scalar_row(pa) || is_scalar_row(pb)) return false;
  if (pa == FB_ARRIVAL_TIME || pa == FB_BURST_WIDTH) return false;
  if (pa == FB_DM && pb == FB_DM_INDEX) return false;
  return true;  // dm_dm, dm x scattering_*, dm_index x {dm_index, scattering_*}, scattering_* x scattering_*
}

__device__ __forceinline__ double d2_dae(int p, const D2Ctx& c) {  // deriv_argument_erf, :252-307
  switch (p) {
    case FB_ARRIVAL_TIME: return -c.iw * kSqrtHalf;
    case FB_BURST_WIDTH: return -(c.T * c.iw2 + c.it) * kSqrtHalf;
    case FB_DM: return c.D * c.iw * kSqrtHalf;
    case FB_DM_INDEX: return c.Dp * c.iw * kSqrtHalf;
    case FB_SCATTERING_TIMESCALE: return c.ws * c.it0 * kSqrtHalf;
    case FB_SCATTERING_INDEX: return c.L * c.ws * kSqrtHalf;
  }
  return 0.0;
}

__device__ __forceinline__ double d2_dax(int p, const D2Ctx& c) {  // deriv_argument_exp, :309-371
  double base = 0.0;
  const double ws = c.ws;
  switch (p) {
    case FB_ARRIVAL_TIME: base = c.it; break;
    case FB_BURST_WIDTH: base = ws * c.it; break;
    case FB_DM: base = -c.D * c.it; break;
    case FB_DM_INDEX: base = -c.Dp * c.it; break;
    case FB_SCATTERING_TIMESCALE: base = (c.T * c.it - ws * ws) * c.it0; break;
    case FB_SCATTERING_INDEX: base = c.L * (c.T * c.it - ws * ws); break;
  }
  return base - 2.0 * c.arg_erf * d2_dae(p, c);
}

// One component's term of deriv2_model_wrt_<pa>_<pb> (unordered pair). Divisions of the
// reference are multiplications by the reciprocals held in the context.
__device__ __forceinline__ double d2_term(int pa, int pb, const D2Ctx& c) {
  if (pa > pb) {
    const int t = pa;
    pa = pb;
    pb = t;
  }
  const double M = c.M, T = c.T, w = c.w, L = c.L, G = c.G;
  const double it = c.it, iw = c.iw, iw2 = c.iw2, it0 = c.it0, ws = c.ws;
  // rows that are scalar multiples of a first derivative, :938-1573
  if (is_scalar_row(pa) || is_scalar_row(pb)) {
    int s, o;  // the reference names amplitude first, then spectral_running, then spectral_index
    if (pa == FB_AMPLITUDE) { s = pa; o = pb; }
    else if (pb == FB_SPECTRAL_RUNNING) { s = pb; o = pa; }
    else if (pa == FB_SPECTRAL_RUNNING) { s = pa; o = pb; }
    else if (pb == FB_SPECTRAL_INDEX) { s = pb; o = pa; }
    else { s = pa; o = pb; }
    const double factor = (s == FB_AMPLITUDE) ? kLn10 : (s == FB_SPECTRAL_RUNNING ? L * L : L);
    return factor * c.d1[o];
  }
  const bool gauss = (c.tau == 0.0);
  const double it2 = it * it;
  const int key = pa * 16 + pb;
  switch (key) {
    case FB_ARRIVAL_TIME * 16 + FB_ARRIVAL_TIME:  // :1957-2014
      if (gauss) return T * T * M * (iw2 * iw2) - M * iw2;
      return c.d1[FB_ARRIVAL_TIME] * it + G * d2_dae(FB_ARRIVAL_TIME, c) * d2_dax(FB_ARRIVAL_TIME, c);
    case FB_ARRIVAL_TIME * 16 + FB_BURST_WIDTH:  // :1639-1700
      if (gauss) return -2.0 * T * M * (iw2 * iw) + T * T * T * M * (iw2 * iw2 * iw);
      return w * c.d1[FB_ARRIVAL_TIME] * it2 + G * (iw2 * kSqrtHalf) +
             G * d2_dae(FB_BURST_WIDTH, c) * d2_dax(FB_ARRIVAL_TIME, c);
    case FB_ARRIVAL_TIME * 16 + FB_DM:  // :2016-2080
      if (gauss) return T * c.d1[FB_DM] * iw2 + c.D * M * iw2;
      return c.d1[FB_DM] * it + G * d2_dae(FB_ARRIVAL_TIME, c) * d2_dax(FB_DM, c);
    case FB_ARRIVAL_TIME * 16 + FB_DM_INDEX:  // :2082-2115 quirk: Gaussian form always, summed first derivative
      return T * c.sum_dmidx * iw2 + (c.k * c.dm * c.q1) * M * iw2;
    case FB_ARRIVAL_TIME * 16 + FB_SCATTERING_TIMESCALE:  // :2117-2174
      return -M * it0 * it + c.d1[FB_SCATTERING_TIMESCALE] * it +
             G * d2_dae(FB_ARRIVAL_TIME, c) * d2_dax(FB_SCATTERING_TIMESCALE, c);
    case FB_ARRIVAL_TIME * 16 + FB_SCATTERING_INDEX:  // :2176-2234
      return -M * L * it + c.d1[FB_SCATTERING_INDEX] * it +
             (G * c.iap) * c.dap_alpha * d2_dae(FB_ARRIVAL_TIME, c) +
             G * d2_dae(FB_ARRIVAL_TIME, c) * d2_dax(FB_SCATTERING_INDEX, c);
    case FB_BURST_WIDTH * 16 + FB_BURST_WIDTH:  // :1575-1637
      if (gauss) return -3.0 * T * T * M * (iw2 * iw2) + T * T * T * T * M * (iw2 * iw2 * iw2);
      return M * it2 + w * c.d1[FB_BURST_WIDTH] * it2 + G * (kSqrt2 * T * (iw2 * iw)) +
             G * d2_dae(FB_BURST_WIDTH, c) * d2_dax(FB_BURST_WIDTH, c);
    case FB_BURST_WIDTH * 16 + FB_DM:  // :1822-1887
      if (gauss) return 2.0 * T * c.D * M * (iw2 * iw) + T * T * c.d1[FB_DM] * (iw2 * iw);
      return w * c.d1[FB_DM] * it2 + G * (-c.D * iw2 * kSqrtHalf) +
             G * d2_dae(FB_BURST_WIDTH, c) * d2_dax(FB_DM, c);
    case FB_BURST_WIDTH * 16 + FB_DM_INDEX:  // :1889-1955 quirk: -2 where the dm twin has +2
      if (gauss) return -2.0 * T * c.Dp * M * (iw2 * iw) + T * T * c.d1[FB_DM_INDEX] * (iw2 * iw);
      return w * c.d1[FB_DM_INDEX] * it2 + G * (-c.Dp * iw2 * kSqrtHalf) +
             G * d2_dae(FB_BURST_WIDTH, c) * d2_dax(FB_DM_INDEX, c);
    case FB_BURST_WIDTH * 16 + FB_SCATTERING_TIMESCALE:  // :1702-1759
      return -2.0 * w * M * it0 * it2 + w * c.d1[FB_SCATTERING_TIMESCALE] * it2 +
             G * (it0 * it * kSqrtHalf) + G * d2_dae(FB_BURST_WIDTH, c) * d2_dax(FB_SCATTERING_TIMESCALE, c);
    case FB_BURST_WIDTH * 16 + FB_SCATTERING_INDEX:  // :1761-1820
      return -2.0 * L * w * M * it2 + w * c.d1[FB_SCATTERING_INDEX] * it2 -
             G * d2_dae(FB_BURST_WIDTH, c) * L + G * (L * kSqrtHalf * it) +
             G * d2_dae(FB_BURST_WIDTH, c) * d2_dax(FB_SCATTERING_INDEX, c);
    case FB_DM * 16 + FB_DM:  // :2275-2341
      if (gauss) return (c.D * T) * (c.D * T) * M * (iw2 * iw2) - (c.D * iw) * (c.D * iw) * M;
      return -c.D * c.d1[FB_DM] * it + G * d2_dae(FB_DM, c) * d2_dax(FB_DM, c);
    case FB_DM * 16 + FB_DM_INDEX: {  // :2236-2273 quirks: Gaussian form, summed derivative, ref_freq ** 2
      double out = (-c.D) * T * c.sum_dmidx * iw2;
      out += c.k * c.k * c.dm * M * iw2 * (c.fpow - c.ref_freq * c.ref_freq) * c.q1;
      out -= c.k * c.q1 * T * M * iw2;
      return out;
    }
    case FB_DM * 16 + FB_SCATTERING_TIMESCALE:  // :2474-2535
      return -c.D * c.d1[FB_SCATTERING_TIMESCALE] * it + c.D * M * it0 * it +
             G * d2_dae(FB_DM, c) * d2_dax(FB_SCATTERING_TIMESCALE, c);
    case FB_DM * 16 + FB_SCATTERING_INDEX: {  // :2666-2720
      const double t0 = -(1.0 + ws * ws - T * it) * L;
      return (L * c.D * it) * M + t0 * c.d1[FB_DM] + G * d2_dae(FB_SCATTERING_INDEX, c) * d2_dax(FB_DM, c);
    }
    case FB_DM_INDEX * 16 + FB_DM_INDEX: {  // :2779-2836 quirk: amp_pbf of the outer component
      const double product = c.k * c.dm * c.q1, product_d = c.k * c.dm * c.q2;
      const double gs = G * (c.apbf_outer * c.iap) * kSqrtHalf;
      return product_d * M * it + product * c.d1[FB_DM_INDEX] * it - product_d * gs * iw * it -
             product * gs * iw * it * d2_dax(FB_DM_INDEX, c);
    }
    case FB_DM_INDEX * 16 + FB_SCATTERING_TIMESCALE: {  // :2537-2594
      const double t0 = -it0 - ws * ws * it0 + T * it0 * it;
      return (c.Dp * it * it0) * M + t0 * c.d1[FB_DM_INDEX] +
             G * d2_dae(FB_SCATTERING_TIMESCALE, c) * d2_dax(FB_DM_INDEX, c);
    }
    case FB_DM_INDEX * 16 + FB_SCATTERING_INDEX: {  // :2722-2777
      const double t0 = -(1.0 + ws * ws - T * it) * L;
      return (L * c.Dp * it) * M + t0 * c.d1[FB_DM_INDEX] +
             G * d2_dae(FB_SCATTERING_INDEX, c) * d2_dax(FB_DM_INDEX, c);
    }
    case FB_SCATTERING_TIMESCALE * 16 + FB_SCATTERING_TIMESCALE: {  // :2343-2404
      const double u = T * it - ws * ws;
      const double t0 = u * it0;
      const double t0d = -u * (it0 * it0) + (-T * it + 2.0 * ws * ws) * (it0 * it0);
      return t0d * M + t0 * c.d1[FB_SCATTERING_TIMESCALE] + G * (-kSqrt2 * ws * (it0 * it0)) +
             G * d2_dae(FB_SCATTERING_TIMESCALE, c) * d2_dax(FB_SCATTERING_TIMESCALE, c);
    }
    case FB_SCATTERING_TIMESCALE * 16 + FB_SCATTERING_INDEX: {  // :2406-2472
      const double e1 = d2_dae(FB_SCATTERING_TIMESCALE, c);
      const double t0 = -(ws * ws) * it0 + T * it0 * it;
      const double t0d2 = 2.0 * ws * ws * it0 * L - T * L * it * it0;
      return t0d2 * M + t0 * c.d1[FB_SCATTERING_INDEX] + (G * c.iap) * c.dap_alpha * e1 +
             G * e1 * d2_dax(FB_SCATTERING_INDEX, c) + G * (-L * e1);
    }
    case FB_SCATTERING_INDEX * 16 + FB_SCATTERING_INDEX: {  // :2596-2664
      const double ad = c.dap_alpha * c.iap, ed = d2_dae(FB_SCATTERING_INDEX, c);
      const double t0 = ad - (ws * ws - T * it) * L;
      const double t0d = c.d2ap_alpha * c.iap - ad * ad - L * (-2.0 * L * ws * ws + T * L * it);
      return t0d * M + t0 * c.d1[FB_SCATTERING_INDEX] + (G * c.iap) * c.dap_alpha * ed +
             G * ed * d2_dax(FB_SCATTERING_INDEX, c) + G * (-ws * L * L * kSqrtHalf);
    }
  }
  return 0.0;
}

__device__ __forceinline__ void make_d2ctx(const EvalArgs& A, const ChanTab& ch, const CompTab& ct, double amp,
                                           double m, double td, double sum_dmidx, double apbf_outer,
                                           D2Ctx& c, bool force_far = false) {
  const int nc = A.pc.num_comp;
  c.tau0 = global_param(A.params, nc, FB_SCATTERING_TIMESCALE);
  first_derivatives(ch, ch, ct, c.tau0, amp, m, td, c.d1, &c.G, &c.arg_erf, force_far);
  const double w = ct.w;
  c.M = m;
  c.T = td;
  c.w = w;
  c.tau = ch.sc_time;
  c.L = ch.logf;
  c.apbf = ch.amp_pbf;
  c.D = ch.d_dm;
  c.Dp = ch.d_dmidx;
  c.fpow = ch.fpow;
  c.ref_freq = ch.ref_freq;
  c.q1 = ch.q1;
  c.q2 = ch.q2;
  c.k = A.pc.dm_const;
  c.dm = global_param(A.params, nc, FB_DM);
  c.dap_alpha = ch.dap_alpha;
  c.d2ap_alpha = ch.d2ap_alpha;
  c.sum_dmidx = sum_dmidx;
  c.apbf_outer = apbf_outer;
  c.it = ch.inv_tau;
  c.iw = ct.inv_w;
  c.iw2 = ct.inv_w * ct.inv_w;
  c.it0 = 1.0 / c.tau0;
  c.ws = ct.w * ch.inv_tau;
  c.iap = ch.inv_apbf;
}

// Pair table entry (built on the host): unordered parameter pair, the `component` the reference
// would pass (analysis/fitter.py:163-168) and whether the function sums over components.
struct HessPair {
  int8_t pa, pb;
  int8_t component;
  int8_t is_sum;
};

constexpr int kHessChunk = 32;             // pairs staged per reduction round
constexpr int kHessStageStride = kBlock + 1;  // odd stride: conflict-free column sums

struct HessArgs {
  EvalArgs E;
  const HessPair* pairs;
  int num_pairs;
  const int* comp_list;  // pair indices applicable to each component, concatenated
  int comp_off[FB_MAX_COMPONENTS + 1];
  double* partials;  // [grid][num_pairs]
  // single-map mode (fb_derivative2_map)
  int map_pa, map_pb, map_component;
  double* map_out;
};

__host__ __device__ inline size_t hess_smem_bytes(int nc, int kf, int num_pairs) {
  size_t n = smem_bytes(nc, kf, 1, true);
  n += sizeof(double) * (size_t)kHessChunk * kHessStageStride;   // staged rho * d2 values
  n += sizeof(double) * (size_t)(kBlock / 32) * num_pairs;       // per-segment accumulators
  return n;
}

template <bool MAP>
__global__ void __launch_bounds__(kBlock, 2) k_hessian(const HessArgs H) {
  extern __shared__ double smem_raw[];
  const EvalArgs& A = H.E;
  const PlanConst& pc = A.pc;
  const int nc = pc.num_comp, kf = pc.kf, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  SmemLayout L = carve(smem_raw, nc, kf);
  const size_t base_doubles = smem_bytes(nc, kf, 1, true) / sizeof(double);
  double* stage = smem_raw + base_doubles;                        // [kHessChunk][kHessStageStride]
  double* wacc = stage + (size_t)kHessChunk * kHessStageStride;   // [segments][num_pairs]
  const int npairs = H.num_pairs;
  const double* vrow = L.vals + (size_t)tid * val_stride(nc);
  if (!MAP)
    for (int e = tid; e < (kBlock / 32) * npairs; e += kBlock) wacc[e] = 0.0;
  load_comp_tables(A, L);
  const int tiles = (pc.num_time + kBlock - 1) / kBlock;
  const double tau0 = global_param(A.params, nc, FB_SCATTERING_TIMESCALE);

  // scintillation needs whole channels per block (two passes); otherwise (channel, tile) items.
  // (amplitudes already solved by k_scint_solve: items like the non-scintillating model)
  const bool whole = pc.scintillation && A.scint_amps == nullptr;
  const long long total = whole ? (long long)A.num_chan : (long long)A.num_chan * tiles;
  const long long begin = total * blockIdx.x / gridDim.x, end = total * (blockIdx.x + 1) / gridDim.x;
  int cur = -1;
  for (long long item = begin; item < end; ++item) {
    const int chi = whole ? (int)item : (int)(item / tiles);
    const int i = A.chan_list ? A.chan_list[chi] : chi;
    const int tile_lo = whole ? 0 : (int)(item % tiles);
    const int tile_hi = whole ? tiles : tile_lo + 1;
    if (chi != cur) {
      load_channel_tables(A, L, i);
      cur = chi;
      if (whole) {
        // amplitudes from the per-channel least-squares solve (routines/ism.py:11-48)
        const int ncols2 = nc + 1, stride2 = gram_stride(ncols2), side2 = gram_side(ncols2);
        const int ntiles2 = gram_ntiles(ncols2), nslices2 = kBlock / ntiles2;
        const int s_tile = tid % ntiles2, s_slice = tid / ntiles2;
        const bool s_active = s_slice < nslices2;
        int s_ti, s_tj;
        tile_coords(s_tile, side2, s_ti, s_tj);
        double* row2 = L.X + (size_t)tid * stride2;
        double acc2[kTile * kTile];
#pragma unroll
        for (int k = 0; k < kTile * kTile; ++k) acc2[k] = 0.0;
        for (int tile = 0; tile < tiles; ++tile) {
          const int j = tile * kBlock + tid;
          eval_tile(A, L, i, tile);
          for (int k = 0; k < stride2; ++k) row2[k] = 0.0;
          if (j < pc.num_time) {
            const double* vrow = L.vals + (size_t)tid * val_stride(nc);
            for (int c = 0; c < nc; ++c) row2[c] = vrow[3 * c + 2];
            row2[nc] = A.data[(size_t)i * pc.num_time + j];
          }
          __syncthreads();
          gram_accumulate(L.X, stride2, kBlock, ntiles2, nslices2, s_ti, s_tj, s_slice, s_active, acc2);
          __syncthreads();
        }
        gram_block_reduce(L.X, ntiles2, nslices2, s_tile, s_slice, s_active, acc2, L.tilesum);
        if (tid == 0) {
          double* G = L.small;
          double* b = L.small + nc * nc;
          for (int t = 0; t < ntiles2; ++t) {
            int ti, tj;
            tile_coords(t, side2, ti, tj);
            for (int r = 0; r < kTile; ++r)
              for (int q = 0; q < kTile; ++q) {
                const int gr = ti * kTile + r, gc = tj * kTile + q;
                const double v = L.tilesum[t * kTile * kTile + r * kTile + q];
                if (gr < nc && gc < nc) {
                  G[gr * nc + gc] = v;
                  G[gc * nc + gr] = v;
                } else if (gr < nc && gc == nc) {
                  b[gr] = v;
                }
              }
          }
          const bool ok = solve_small(G, b, nc, nc);
          for (int c = 0; c < nc; ++c) L.scint_amp[c] = ok ? b[c] : nan("");
          if (!ok && A.err_flag) atomicExch(A.err_flag, 1);
        }
        __syncthreads();
      }
    }
    for (int tile = tile_lo; tile < tile_hi; ++tile) {
      const int j = tile * kBlock + tid;
      const bool valid = j < pc.num_time;
      const size_t s_idx = (size_t)i * pc.num_time + (valid ? j : 0);
      double msum = 0.0, sum_dmidx = 0.0;
      eval_tile(A, L, i, tile);
      if (valid) {
        for (int c = 0; c < nc; ++c) {
          SampleComp sc;
          sc.spectrum = vrow[3 * c];
          sc.timediff = vrow[3 * c + 1];
          sc.timeprof = vrow[3 * c + 2];
          double amp = L.chan[c].amp;
          if (pc.scintillation) {
            amp = L.scint_amp[c];
            sc.spectrum = amp * sc.timeprof;
          }
          msum += sc.spectrum;
          double d[9], g;
          first_derivatives(L.chan[c], L.chan[c], L.comp[c], tau0, amp, sc.spectrum, sc.timediff, d, &g);
          sum_dmidx += d[FB_DM_INDEX];
        }
      }
      if (MAP) {
        if (valid) {
          const bool sum = d2_is_sum(H.map_pa, H.map_pb);
          double out = 0.0;
          for (int c = 0; c < nc; ++c) {
            if (!sum && c != H.map_component) continue;
            D2Ctx ctx;
            const double amp = pc.scintillation ? L.scint_amp[c] : L.chan[c].amp;
            const double spec = pc.scintillation ? amp * vrow[3 * c + 2] : vrow[3 * c];
            make_d2ctx(A, L.chan[c], L.comp[c], amp, spec, vrow[3 * c + 1], sum_dmidx,
                       L.chan[H.map_component].amp_pbf, ctx);
            out += d2_term(H.map_pa, H.map_pb, ctx);
          }
          H.map_out[s_idx] = out;
        }
        continue;
      }
      double rho = 0.0;
      if (valid) {
        const double wi = A.weights[i];
        rho = (A.data[s_idx] - msum) * wi * wi;
      }
      for (int c = 0; c < nc; ++c) {
        D2Ctx ctx;
        if (valid) {
          const double amp = pc.scintillation ? L.scint_amp[c] : L.chan[c].amp;
          const double spec = pc.scintillation ? amp * vrow[3 * c + 2] : vrow[3 * c];
          make_d2ctx(A, L.chan[c], L.comp[c], amp, spec, vrow[3 * c + 1], sum_dmidx, 0.0, ctx);
        }
        // pairs this component contributes to, staged kHessChunk at a time: every thread writes
        // its rho * d2 values into a column of `stage`, then thread (k, seg) sums 32 rows of
        // staged pair k -- no shuffles, fixed order, one owner per (segment, pair) accumulator.
        const int lo = H.comp_off[c], cnt = H.comp_off[c + 1] - lo;
        for (int chunk = 0; chunk < cnt; chunk += kHessChunk) {
          const int m = min(kHessChunk, cnt - chunk);
          for (int k = 0; k < m; ++k) {
            double v = 0.0;
            if (valid) {
              const HessPair hp = H.pairs[H.comp_list[lo + chunk + k]];
              ctx.apbf_outer = L.chan[hp.component].amp_pbf;
              v = rho * d2_term(hp.pa, hp.pb, ctx);
            }
            stage[k * kHessStageStride + tid] = v;
          }
          __syncthreads();
          {
            const int k = tid & 31, seg = tid >> 5;
            if (k < m) {
              const double* col = stage + k * kHessStageStride + seg * 32;
              double sum = 0.0;
#pragma unroll 8
              for (int r = 0; r < 32; ++r) sum += col[r];
              wacc[seg * npairs + H.comp_list[lo + chunk + k]] += sum;
            }
          }
          __syncthreads();
        }
      }
    }
  }
  if (!MAP) {
    __syncthreads();
    for (int pi = tid; pi < npairs; pi += kBlock) {
      double s = 0.0;
      for (int wv = 0; wv < kBlock / 32; ++wv) s += wacc[wv * npairs + pi];
      H.partials[(size_t)blockIdx.x * npairs + pi] = s;
    }
  }
}

// ---- accumulate path -------------------------------------------------------------------------
// S_pair = sum_samples rho * d2 is linear in a small set of per-(sample, component) values: the
// nine first derivatives and the 21 non-trivial mixed partials ("term values", tv[0..29]); every
// one of the 45 functions is one of them times a per-CHANNEL factor (1, ln 10, L, L^2). So each
// thread accumulates rho * tv[k] in registers over a super-tile of samples, the block sums the
// columns once per (super-tile, component), and only then are the pair accumulators touched.
constexpr int kNumTerms = 30;
constexpr int kHessSuper = 2;      // tiles per super-tile
constexpr int kFlushChunk = 15;    // accumulators staged per flush round (fits the X region)

struct HessDesc {
  int8_t code;  // 0: 1, 1: ln 10, 2: L, 3: L^2
  int8_t idx;   // term-value index
};

__host__ __device__ inline void pair_descriptor(int pa, int pb, int& code, int& idx) {
  if (pa > pb) {
    const int t = pa;
    pa = pb;
    pb = t;
  }
  if (is_scalar_row(pa) || is_scalar_row(pb)) {
    int s, o;
    if (pa == FB_AMPLITUDE) { s = pa; o = pb; }
    else if (pb == FB_SPECTRAL_RUNNING) { s = pb; o = pa; }
    else if (pa == FB_SPECTRAL_RUNNING) { s = pa; o = pb; }
    else if (pb == FB_SPECTRAL_INDEX) { s = pb; o = pa; }
    else { s = pa; o = pb; }
    code = (s == FB_AMPLITUDE) ? 1 : (s == FB_SPECTRAL_RUNNING ? 3 : 2);
    idx = o;
    return;
  }
  // pa <= pb both in {ARRIVAL_TIME=1 .. SCATTERING_INDEX=6}: position in the upper triangle
  const int a = pa - 1, b = pb - 1;  // 0..5
  code = 0;
  idx = 9 + a * 6 - a * (a - 1) / 2 + (b - a);
}

// rho-weighted accumulation of every needed term value of one (sample, component).
__device__ __forceinline__ void accumulate_terms(const D2Ctx& c, double rho, unsigned need,
                                                 double (&acc)[kNumTerms]) {
#pragma unroll
  for (int k = 0; k < 9; ++k)
    if (need & (1u << k)) acc[k] = fma(rho, c.d1[k], acc[k]);
#define FB_TERM(IDX, PA, PB) \
  if (need & (1u << (IDX))) acc[IDX] = fma(rho, d2_term(PA, PB, c), acc[IDX]);
  FB_TERM(9, FB_ARRIVAL_TIME, FB_ARRIVAL_TIME)
  FB_TERM(10, FB_ARRIVAL_TIME, FB_BURST_WIDTH)
  FB_TERM(11, FB_ARRIVAL_TIME, FB_DM)
  FB_TERM(12, FB_ARRIVAL_TIME, FB_DM_INDEX)
  FB_TERM(13, FB_ARRIVAL_TIME, FB_SCATTERING_TIMESCALE)
  FB_TERM(14, FB_ARRIVAL_TIME, FB_SCATTERING_INDEX)
  FB_TERM(15, FB_BURST_WIDTH, FB_BURST_WIDTH)
  FB_TERM(16, FB_BURST_WIDTH, FB_DM)
  FB_TERM(17, FB_BURST_WIDTH, FB_DM_INDEX)
  FB_TERM(18, FB_BURST_WIDTH, FB_SCATTERING_TIMESCALE)
  FB_TERM(19, FB_BURST_WIDTH, FB_SCATTERING_INDEX)
  FB_TERM(20, FB_DM, FB_DM)
  FB_TERM(21, FB_DM, FB_DM_INDEX)
  FB_TERM(22, FB_DM, FB_SCATTERING_TIMESCALE)
  FB_TERM(23, FB_DM, FB_SCATTERING_INDEX)
  FB_TERM(24, FB_DM_INDEX, FB_DM_INDEX)
  FB_TERM(25, FB_DM_INDEX, FB_SCATTERING_TIMESCALE)
  FB_TERM(26, FB_DM_INDEX, FB_SCATTERING_INDEX)
  FB_TERM(27, FB_SCATTERING_TIMESCALE, FB_SCATTERING_TIMESCALE)
  FB_TERM(28, FB_SCATTERING_TIMESCALE, FB_SCATTERING_INDEX)
  FB_TERM(29, FB_SCATTERING_INDEX, FB_SCATTERING_INDEX)
#undef FB_TERM
}

struct HessAccArgs {
  EvalArgs E;
  const HessDesc* desc;   // per pair
  int num_pairs;
  const int* comp_list;   // pair indices applicable to each component, concatenated
  int comp_off[FB_MAX_COMPONENTS + 1];
  unsigned need[FB_MAX_COMPONENTS];  // term values needed per component
  int need_sum_dmidx;
  double* partials;       // [grid][num_pairs]
  // (spectrum, timediff) planes left in global memory by k_vals at the SAME parameters
  // (the last trust-region evaluation); nullptr = evaluate the model here
  const double* gvals;
  int gvals_layout;  // 2: k_vals tile planes (spectrum, timediff); 3: k_vals3 rows (spectrum only)
};

// Term value `idx` (0..8 first derivatives, 9..29 second derivatives in FB_TERM order).
__device__ inline double hess_term_value(int idx, const D2Ctx& c) {
  if (idx < 9) return c.d1[idx];
  constexpr int8_t pa[21] = {FB_ARRIVAL_TIME, FB_ARRIVAL_TIME, FB_ARRIVAL_TIME, FB_ARRIVAL_TIME, FB_ARRIVAL_TIME,
                             FB_ARRIVAL_TIME, FB_BURST_WIDTH, FB_BURST_WIDTH, FB_BURST_WIDTH, FB_BURST_WIDTH,
                             FB_BURST_WIDTH, FB_DM, FB_DM, FB_DM, FB_DM, FB_DM_INDEX, FB_DM_INDEX, FB_DM_INDEX,
                             FB_SCATTERING_TIMESCALE, FB_SCATTERING_TIMESCALE, FB_SCATTERING_INDEX};
  constexpr int8_t pb[21] = {FB_ARRIVAL_TIME, FB_BURST_WIDTH, FB_DM, FB_DM_INDEX, FB_SCATTERING_TIMESCALE,
                             FB_SCATTERING_INDEX, FB_BURST_WIDTH, FB_DM, FB_DM_INDEX, FB_SCATTERING_TIMESCALE,
                             FB_SCATTERING_INDEX, FB_DM, FB_DM_INDEX, FB_SCATTERING_TIMESCALE, FB_SCATTERING_INDEX,
                             FB_DM_INDEX, FB_SCATTERING_TIMESCALE, FB_SCATTERING_INDEX, FB_SCATTERING_TIMESCALE,
                             FB_SCATTERING_INDEX, FB_SCATTERING_INDEX};
  return d2_term(pa[idx - 9], pb[idx - 9], c);
}

__host__ __device__ inline size_t hess_acc_smem_bytes(int nc, int kf, int num_pairs) {
  size_t n = smem_bytes(nc, kf, 1, false, true);
  n += sizeof(double) * (size_t)kHessSuper * kBlock * val_stride(nc);  // term inputs of the super-tile
  n += sizeof(double) * (size_t)kHessSuper * kBlock * 2;               // rho, sum_dmidx
  n += sizeof(double) * ((size_t)(kBlock / 32) * 32 + 32);             // column partials, totals
  n += sizeof(double) * (size_t)num_pairs;
  n += sizeof(double) * ((size_t)3 * nc * kBlock + 3 * nc + 3 * kNumTerms);  // far-sample moments, node values
  return n;
}

// Far from the burst the shared factor G of the scattered-branch derivatives is exactly zero
// (first_derivatives: exp underflow cut at arg_exp < -746, |T| > 38.6 w), and every one of the 30
// term values is then the spectrum M times a polynomial of degree <= 2 in the time difference T
// with per-(channel, component) coefficients (see d2_term with G = 0). For those samples -- ~94 %
// at config 2 -- only the three moments sum(rho M x^p), x = T mapped to [-1, 1] over the channel,
// are accumulated, and at the end of the channel the term polynomials are evaluated at the three
// nodes x = -1, 0, 1 with the SAME device functions (M = 1, G = 0) and contracted with the
// Lagrange weights of the moments: exact for quadratics, no per-term algebra to keep in sync.
__device__ __forceinline__ void hessian_acc_body(const HessAccArgs& H, const int block_id, const int num_blocks) {
  extern __shared__ double smem_raw[];
  const EvalArgs& A = H.E;
  const PlanConst& pc = A.pc;
  const int nc = pc.num_comp, kf = pc.kf, tid = threadIdx.x;
  SmemLayout L = carve(smem_raw, nc, kf, true);
  const int vs = val_stride(nc);
  double* big = smem_raw + smem_bytes(nc, kf, 1, false, true) / sizeof(double);  // [super rows][vs]
  double* rho_s = big + (size_t)kHessSuper * kBlock * vs;
  double* sdm_s = rho_s + (size_t)kHessSuper * kBlock;
  double* part = sdm_s + (size_t)kHessSuper * kBlock;  // [8][32]
  double* total = part + (kBlock / 32) * 32;           // [32]
  double* pacc = total + 32;                           // [num_pairs]
  double* mu_s = pacc + H.num_pairs;                   // [nc * 3][kBlock] per-thread moment slots
  double* lam_s = mu_s + (size_t)3 * nc * kBlock;      // [nc * 3] Lagrange-weighted sums
  double* node_s = lam_s + 3 * nc;                     // [3][kNumTerms] term values at the nodes
  double* stage = L.X;                                 // [kFlushChunk][kBlock + 1] (X region >= 32 KB)
  const int npairs = H.num_pairs;
  for (int e = tid; e < npairs; e += kBlock) pacc[e] = 0.0;
  for (int e = tid; e < 3 * nc * kBlock; e += kBlock) mu_s[e] = 0.0;
  load_comp_tables(A, L);
  const int tiles = (pc.num_time + kBlock - 1) / kBlock;
  const int supers = (tiles + kHessSuper - 1) / kHessSuper;
  const double tau0 = global_param(A.params, nc, FB_SCATTERING_TIMESCALE);
  const double kt_mid = 0.5 * (double)(pc.kt - 1);
  const double tj_last = (double)((pc.num_time - 1) * pc.kt) + kt_mid;
  const bool shortcut = !H.need_sum_dmidx && pc.num_time > 2;
  const long long total_items = (long long)A.num_chan * supers;
  const long long begin = total_items * block_id / num_blocks, end = total_items * (block_id + 1) / num_blocks;

  // contracts the moments of the channel whose tables are in shared memory; block-uniform
  auto flush_moments = [&]() {
    __syncthreads();
    if (tid < 3 * nc) {
      double s = 0.0;
      for (int t = 0; t < kBlock; ++t) s += mu_s[(size_t)tid * kBlock + t];
      lam_s[tid] = s;
    }
    __syncthreads();
    for (int e = tid; e < 3 * nc * kBlock; e += kBlock) mu_s[e] = 0.0;
    for (int c = 0; c < nc; ++c) {
      const double m0 = lam_s[3 * c], m1 = lam_s[3 * c + 1], m2 = lam_s[3 * c + 2];
      if (m0 == 0.0 && m1 == 0.0 && m2 == 0.0) continue;  // block-uniform
      const unsigned need = H.need[c];
      if (tid < 3 * kNumTerms) {
        const int q = tid / kNumTerms, idx = tid % kNumTerms;
        double v = 0.0;
        if (need & (1u << idx)) {
          const CompTab& ct = L.comp[c];
          const ChanTab& ch = L.chan[c];
          const double t_lo = fma(ct.t_step, kt_mid, ct.t_start) - ch.mean_delay;
          const double t_hi = fma(ct.t_step, tj_last, ct.t_start) - ch.mean_delay;
          const double t_node = q == 0 ? t_lo : (q == 2 ? t_hi : 0.5 * (t_lo + t_hi));
          D2Ctx ctx;
          make_d2ctx(A, ch, ct, ch.amp, 1.0, t_node, 0.0, L.chan[0].amp_pbf, ctx, true);
          v = hess_term_value(idx, ctx);
        }
        node_s[tid] = v;
      }
      __syncthreads();
      if (tid < kNumTerms) {
        // Lagrange weights of the nodes -1, 0, 1 applied to the moments
        const double l_lo = 0.5 * (m2 - m1), l_mid = m0 - m2, l_hi = 0.5 * (m2 + m1);
        total[tid] = node_s[tid] * l_lo + node_s[kNumTerms + tid] * l_mid + node_s[2 * kNumTerms + tid] * l_hi;
      }
      __syncthreads();
      {
        const double lf = L.chan[c].logf;
        const int lo = H.comp_off[c], cnt = H.comp_off[c + 1] - lo;
        for (int p = tid; p < cnt; p += kBlock) {
          const int pair = H.comp_list[lo + p];
          const HessDesc d = H.desc[pair];
          const double factor = d.code == 0 ? 1.0 : (d.code == 1 ? kLn10 : (d.code == 2 ? lf : lf * lf));
          pacc[pair] += factor * total[d.idx];
        }
      }
      __syncthreads();
    }
  };

  int cur = -1;
  for (long long item = begin; item < end; ++item) {
    const int chi = (int)(item / supers), sup = (int)(item % supers);
    const int i = A.chan_list ? A.chan_list[chi] : chi;
    if (chi != cur) {
      if (cur >= 0 && shortcut) flush_moments();
      load_channel_tables(A, L, i);
      cur = chi;
    }
    const int t_lo = sup * kHessSuper, t_hi = min(tiles, t_lo + kHessSuper);
    const double wi = A.weights[i];
    // 1. term inputs (spectrum, timediff) of every component + rho for the super-tile
    for (int t = t_lo; t < t_hi; ++t) {
      double* dst = big + (size_t)(t - t_lo) * kBlock * vs;
      if (H.gvals && H.gvals_layout == 3) {
        const int jv = t * kBlock + tid;
        if (jv < pc.num_time) {
          const double tj = (double)(jv * pc.kt) + kt_mid;
          for (int c = 0; c < nc; ++c) {
            dst[(size_t)tid * vs + 3 * c] = H.gvals[((size_t)chi * nc + c) * pc.num_time + jv];
            // closed form of the mean time difference, fb_core.cuh: eval_sample_fast
            dst[(size_t)tid * vs + 3 * c + 1] = fma(L.comp[c].t_step, tj, L.comp[c].t_start) - L.chan[c].mean_delay;
          }
        }
      } else if (H.gvals) {
        const double* g = H.gvals + ((size_t)chi * tiles + t) * kBlock * 2 * nc + tid;
        for (int c = 0; c < nc; ++c) {
          dst[(size_t)tid * vs + 3 * c] = g[(size_t)(2 * c) * kBlock];
          dst[(size_t)tid * vs + 3 * c + 1] = g[(size_t)(2 * c + 1) * kBlock];
        }
      } else {
        eval_tile(A, L, i, t, dst);
      }
      const int j = t * kBlock + tid, r = (t - t_lo) * kBlock + tid;
      double rho = 0.0, sdm = 0.0;
      if (j < pc.num_time) {
        const double* vrow = dst + (size_t)tid * vs;
        double msum = 0.0;
        for (int c = 0; c < nc; ++c) msum += vrow[3 * c];
        rho = (A.data[(size_t)i * pc.num_time + j] - msum) * wi * wi;
        if (H.need_sum_dmidx) {
          for (int c = 0; c < nc; ++c) {
            double d[9], g;
            first_derivatives(L.chan[c], L.chan[c], L.comp[c], tau0, L.chan[c].amp, vrow[3 * c], vrow[3 * c + 1], d, &g);
            sdm += d[FB_DM_INDEX];
          }
        }
      }
      rho_s[r] = rho;
      sdm_s[r] = sdm;
    }
    __syncthreads();
    // 2. per component: register accumulation over the super-tile, then one block flush
    for (int c = 0; c < nc; ++c) {
      double acc[kNumTerms];
#pragma unroll
      for (int k = 0; k < kNumTerms; ++k) acc[k] = 0.0;
      const unsigned need = H.need[c];
      const CompTab& ct = L.comp[c];
      const ChanTab& ch = L.chan[c];
      const bool gate = shortcut && ch.sc_time != 0.0;
      // T -> x in [-1, 1] over the channel
      const double tc_lo = fma(ct.t_step, kt_mid, ct.t_start) - ch.mean_delay;
      const double tc_hi = fma(ct.t_step, tj_last, ct.t_start) - ch.mean_delay;
      const double tc_mid = 0.5 * (tc_lo + tc_hi), inv_half = 2.0 / (tc_hi - tc_lo);
      double mu0 = 0.0, mu1 = 0.0, mu2 = 0.0;
      int any_full = 0;
      for (int t = t_lo; t < t_hi; ++t) {
        const int j = t * kBlock + tid, r = (t - t_lo) * kBlock + tid;
        if (j < pc.num_time) {
          const double* vrow = big + (size_t)r * vs;
          const double m = vrow[3 * c], td = vrow[3 * c + 1];
          bool far = false;
          if (gate) {  // the very expressions of first_derivatives
            const double arg_erf = (td - ct.w * ct.w * ch.inv_tau) * ct.inv_w * kSqrtHalf;
            const double ws = ct.w * ch.inv_tau;
            const double arg_exp = 0.5 * ws * ws - td * ch.inv_tau - arg_erf * arg_erf;
            far = arg_exp < -746.0;
          }
          if (far) {
            const double x = (td - tc_mid) * inv_half, rm = rho_s[r] * m;
            mu0 += rm;
            mu1 = fma(rm, x, mu1);
            mu2 = fma(rm * x, x, mu2);
          } else {
            D2Ctx ctx;
            make_d2ctx(A, ch, ct, ch.amp, m, td, sdm_s[r], L.chan[0].amp_pbf, ctx);
            accumulate_terms(ctx, rho_s[r], need, acc);
            any_full = 1;
          }
        }
      }
      if (gate) {
        mu_s[(size_t)(3 * c) * kBlock + tid] += mu0;
        mu_s[(size_t)(3 * c + 1) * kBlock + tid] += mu1;
        mu_s[(size_t)(3 * c + 2) * kBlock + tid] += mu2;
      }
      if (!__syncthreads_or(any_full)) continue;  // nothing but far samples in this super-tile
      // flush: column sums of the needed accumulators (fixed order), then the pair updates
      for (int k0 = 0; k0 < kNumTerms; k0 += kFlushChunk) {
#pragma unroll
        for (int k = 0; k < kFlushChunk; ++k) stage[k * (kBlock + 1) + tid] = acc[k0 + k];
        __syncthreads();
        {
          const int k = tid & 31, seg = tid >> 5;
          if (k < kFlushChunk) {
            const double* col = stage + k * (kBlock + 1) + seg * 32;
            double sum = 0.0;
#pragma unroll 8
            for (int r = 0; r < 32; ++r) sum += col[r];
            part[seg * 32 + k] = sum;
          }
        }
        __syncthreads();
        if (tid < kFlushChunk) {
          double sum = 0.0;
          for (int seg = 0; seg < kBlock / 32; ++seg) sum += part[seg * 32 + tid];
          total[k0 + tid] = sum;
        }
        __syncthreads();
      }
      {
        const double lf = L.chan[c].logf;
        const int lo = H.comp_off[c], cnt = H.comp_off[c + 1] - lo;
        for (int p = tid; p < cnt; p += kBlock) {
          const int pair = H.comp_list[lo + p];
          const HessDesc d = H.desc[pair];
          const double factor = d.code == 0 ? 1.0 : (d.code == 1 ? kLn10 : (d.code == 2 ? lf : lf * lf));
          pacc[pair] += factor * total[d.idx];
        }
      }
      __syncthreads();
    }
  }
  if (cur >= 0 && shortcut) flush_moments();
  __syncthreads();
  for (int e = tid; e < npairs; e += kBlock) H.partials[(size_t)block_id * npairs + e] = pacc[e];
}

__global__ void __launch_bounds__(kBlock, 4) k_hessian_acc(const HessAccArgs H) {
  hessian_acc_body(H, blockIdx.x, gridDim.x);
}

// Batched variant (the on-device batched solver, fb_batched.cuh): blockIdx.y = burst, every
// per-burst buffer offset like k_eval_batched does; bursts whose fit failed are skipped.
__global__ void __launch_bounds__(kBlock, 4) k_hessian_acc_batched(const HessAccArgs H, const BatchArgs B,
                                                                   const int* __restrict__ skip) {
  const int y = blockIdx.y;
  if (skip[y]) return;
  HessAccArgs Hb = H;
  Hb.E.data = H.E.data + y * B.data_stride;
  Hb.E.weights = H.E.weights + y * B.weights_stride;
  Hb.E.params = H.E.params + y * B.params_stride;
  Hb.E.g_comp = H.E.g_comp + y * B.comp_stride;
  Hb.E.g_sub = H.E.g_sub + y * B.sub_stride;
  Hb.E.g_chan = H.E.g_chan + y * B.chan_stride;
  Hb.E.chan_list = H.E.chan_list + y * B.list_stride;
  Hb.E.num_chan = B.num_good[y];
  Hb.partials = H.partials + y * B.partials_stride;
  hessian_acc_body(Hb, blockIdx.x, gridDim.x);
}

// Sums the per-block partials of every (burst, pair): out[y * npairs + pair].
__global__ void __launch_bounds__(128) k_hess_finalize_batched(const double* __restrict__ partials, int nblocks,
                                                              int npairs, size_t burst_stride, double* __restrict__ out) {
  const int y = blockIdx.y;
  const int pi = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (pi >= npairs) return;
  const double* part = partials + (size_t)y * burst_stride;
  double s = 0.0;
  for (int b = lane; b < nblocks; b += 32) s += part[(size_t)b * npairs + pi];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) out[(size_t)y * npairs + pi] = s;
}

// ---- warp-granular Hessian contraction for the fast path (fb_fast.cuh conventions) ------------
// Same far-sample polynomial shortcut as k_hessian_acc, organised like k_gram3: a block owns one
// channel at a time and reads the spectrum rows k_vals3 left in global memory.
//   phase A  every warp streams 32-sample tiles; samples OUTSIDE the near range [j_lo, j_hi) of a
//            component (where the shared factor G is exactly zero) only feed the three moments
//            sum(rho M x^p), held in registers for the whole channel;
//   phase B  the near range of each component (a few dozen samples around the burst) is walked
//            by the whole block with the full 30-term evaluation, then one block flush;
//   end      the moments are contracted with the term polynomials at the nodes x = -1, 0, 1.
constexpr int kHess3Chunk = 15;

__host__ __device__ inline size_t hess3_smem_bytes(int nc, int num_pairs) {
  size_t n = sizeof(CompTab) * nc + 2 * sizeof(ChanTab) * nc;
  n += sizeof(double) * (size_t)kHess3Chunk * (kFastBlock + 1);         // flush stage
  n += sizeof(double) * ((size_t)nc * kFastWarps * 32 + (size_t)nc * 32);  // warp totals, block totals
  n += sizeof(double) * (size_t)num_pairs;
  n += sizeof(double) * ((size_t)kFastWarps * 3 * nc);                  // per-warp moments
  n += sizeof(double) * 2 * nc;                                         // x mapping per component
  n += sizeof(int) * (3 * nc + 4 + kFastBlock);                         // ranges, offsets, item component
  return n;
}

template <int NC>
__global__ void __launch_bounds__(kFastBlock, FB_H3_MINB) k_hess3(const HessAccArgs H) {
  extern __shared__ __align__(16) unsigned char fast_smem[];
  const EvalArgs& A = H.E;
  const PlanConst& pc = A.pc;
  const int nt = pc.num_time, kt = pc.kt;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int npairs = H.num_pairs;
  char* sp = reinterpret_cast<char*>(fast_smem);
  CompTab* comp = reinterpret_cast<CompTab*>(sp);
  sp += sizeof(CompTab) * NC;
  char* tabs = sp;
  sp += 2 * sizeof(ChanTab) * NC;
  double* stage = reinterpret_cast<double*>(sp);             // [kHess3Chunk][kFastBlock + 1]
  double* part = stage + (size_t)kHess3Chunk * (kFastBlock + 1);  // [NC][warps][32]
  double* total = part + NC * kFastWarps * 32;               // [NC][32]
  double* pacc = total + NC * 32;
  double* musum = pacc + npairs;                             // [warps][3 NC]
  double* xmap = musum + kFastWarps * 3 * NC;                // [NC][2]: x = tj * xmap[0] + xmap[1]
  int* range = reinterpret_cast<int*>(xmap + 2 * NC);        // [NC][2]
  int* offs = range + 2 * NC;                                // [NC + 1] item offsets of the near ranges
  int* compid = offs + NC + 4;                               // [kFastBlock] component of each thread's item

  {
    const double* src = reinterpret_cast<const double*>(A.g_comp);
    double* dst = reinterpret_cast<double*>(comp);
    for (int e = tid; e < NC * (int)(sizeof(CompTab) / sizeof(double)); e += kFastBlock) dst[e] = src[e];
  }
  for (int e = tid; e < npairs; e += kFastBlock) pacc[e] = 0.0;
  const int wtiles = (nt + 31) / 32;
  const int chan_chunks = (int)(sizeof(ChanTab) * NC / 16);
  auto prefetch_tables = [&](int chi, int buf) {
    if (chi < A.num_chan && tid < chan_chunks) {
      const int i = A.chan_list ? A.chan_list[chi] : chi;
      cp_async16(tabs + (size_t)buf * sizeof(ChanTab) * NC + 16 * tid,
                 reinterpret_cast<const char*>(A.g_chan + (size_t)i * NC) + 16 * tid);
    }
    cp_async_commit();
  };
  const double tau0 = global_param(A.params, NC, FB_SCATTERING_TIMESCALE);
  const double kt_mid = 0.5 * (double)(kt - 1);
  const double tj_last = (double)((nt - 1) * kt) + kt_mid;
  const bool shortcut = !H.need_sum_dmidx && nt > 2;

  prefetch_tables(blockIdx.x, 0);
  int buf = 0;
  for (int chi = blockIdx.x; chi < A.num_chan; chi += gridDim.x, buf ^= 1) {
    const int i = A.chan_list ? A.chan_list[chi] : chi;
    cp_async_wait<0>();
    __syncthreads();
    prefetch_tables(chi + gridDim.x, buf ^ 1);
    const ChanTab* chan = reinterpret_cast<const ChanTab*>(tabs + (size_t)buf * sizeof(ChanTab) * NC);
    if (tid < NC) {
      const CompTab& ct = comp[tid];
      const ChanTab& ch = chan[tid];
      int j_lo = 0, j_hi = nt;
      if (shortcut) {  // precomputed by the table kernel (finish_chan_tab)
        j_lo = ch.near_lo;
        j_hi = ch.near_hi;
      }
      range[2 * tid] = j_lo;
      range[2 * tid + 1] = j_hi;
      const double tc_lo = fma(ct.t_step, kt_mid, ct.t_start) - ch.mean_delay;
      const double tc_hi = fma(ct.t_step, tj_last, ct.t_start) - ch.mean_delay;
      const double tc_mid = 0.5 * (tc_lo + tc_hi), inv_half = 2.0 / (tc_hi - tc_lo);
      xmap[2 * tid] = ct.t_step * inv_half;
      xmap[2 * tid + 1] = (ct.t_start - ch.mean_delay - tc_mid) * inv_half;
    }
    __syncthreads();
    if (tid == 0) {
      int o = 0;
      for (int c = 0; c < NC; ++c) {
        offs[c] = o;
        o += range[2 * c + 1] - range[2 * c];
      }
      offs[NC] = o;
    }
    const double wi = A.weights[i];
    const double w2 = wi * wi;
    const double* srow = H.gvals + (size_t)chi * NC * nt;
    const double* drow = A.data + (size_t)i * nt;

    // phase A: far samples -> moments
    {
      double mu[NC][3];
#pragma unroll
      for (int c = 0; c < NC; ++c) mu[c][0] = mu[c][1] = mu[c][2] = 0.0;
      // eight tiles of the warp per round (a whole channel at 1024 samples): all their loads are issued before the first use (with
      // one tile at a time every tile paid a full DRAM latency)
      constexpr int kRound = 8;
      for (int wt = warp; wt < wtiles; wt += kRound * kFastWarps) {
        double m[kRound][NC], dv[kRound];
#pragma unroll
        for (int r = 0; r < kRound; ++r) {
          const int j = (wt + r * kFastWarps) * 32 + lane;
          const bool ok = j < nt;
#pragma unroll
          for (int c = 0; c < NC; ++c) m[r][c] = ok ? __ldg(srow + (size_t)c * nt + j) : 0.0;
          dv[r] = ok ? __ldg(drow + j) : 0.0;
        }
#pragma unroll
        for (int r = 0; r < kRound; ++r) {
          const int j = (wt + r * kFastWarps) * 32 + lane;
          if (j < nt) {
            double msum = 0.0;
#pragma unroll
            for (int c = 0; c < NC; ++c) msum += m[r][c];
            const double rho = (dv[r] - msum) * w2;
            const double tj = (double)(j * kt) + kt_mid;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
              if (j < range[2 * c] || j >= range[2 * c + 1]) {
                const double x = fma(tj, xmap[2 * c], xmap[2 * c + 1]), rm = rho * m[r][c];
                mu[c][0] += rm;
                mu[c][1] = fma(rm, x, mu[c][1]);
                mu[c][2] = fma(rm * x, x, mu[c][2]);
              }
            }
          }
        }
      }
#pragma unroll
      for (int c = 0; c < NC; ++c)
#pragma unroll
        for (int p = 0; p < 3; ++p) {
          double v = mu[c][p];
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(kFullMask, v, off);
          if (lane == 0) musum[warp * 3 * NC + 3 * c + p] = v;
        }
    }
    __syncthreads();

    // phase B: one item per thread -- the near samples of every component with the full term
    // evaluation, plus 3 node items per component that contract the far-sample moments. The 30
    // term values of the 32 items of a warp are summed per component by a halving exchange (31
    // double shuffles: afterwards lane k holds the warp's sum of term k) into registers that live
    // for the whole channel; the block meets once per channel, not once per pass and chunk.
    const int n_near = offs[NC];
    const int n_items = n_near + (shortcut ? 3 * NC : 0);
    double tot[NC];  // lane k: this warp's sum of term k for component c over the channel
#pragma unroll
    for (int c = 0; c < NC; ++c) tot[c] = 0.0;
    for (int base = 0; base < n_items; base += kFastBlock) {
      const int e = base + tid;
      double acc[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) acc[k] = 0.0;
      int my_c = -1;
      if (e < n_near) {
        int c = 0;
#pragma unroll
        for (int cc = 1; cc < NC; ++cc)
          if (e >= offs[cc]) c = cc;
        const int j = range[2 * c] + (e - offs[c]);
        double msum = 0.0, mc = 0.0, sdm = 0.0;
        const double tj = (double)(j * kt) + kt_mid;
#pragma unroll
        for (int cc = 0; cc < NC; ++cc) {
          const double mv = __ldg(srow + (size_t)cc * nt + j);
          msum += mv;
          if (cc == c) mc = mv;
          if (H.need_sum_dmidx) {
            double d[9], g;
            first_derivatives(chan[cc], chan[cc], comp[cc], tau0, chan[cc].amp, mv,
                              fma(comp[cc].t_step, tj, comp[cc].t_start) - chan[cc].mean_delay, d, &g);
            sdm += d[FB_DM_INDEX];
          }
        }
        const double rho = (__ldg(drow + j) - msum) * w2;
        const double td = fma(comp[c].t_step, tj, comp[c].t_start) - chan[c].mean_delay;
        D2Ctx ctx;
        make_d2ctx(A, chan[c], comp[c], chan[c].amp, mc, td, sdm, chan[0].amp_pbf, ctx);
        accumulate_terms(ctx, rho, H.need[c], reinterpret_cast<double (&)[kNumTerms]>(acc));
        my_c = c;
      } else if (e < n_items) {
        const int c = (e - n_near) / 3, q = (e - n_near) % 3;
        double m0 = 0.0, m1 = 0.0, m2 = 0.0;
        for (int wv = 0; wv < kFastWarps; ++wv) {
          m0 += musum[wv * 3 * NC + 3 * c];
          m1 += musum[wv * 3 * NC + 3 * c + 1];
          m2 += musum[wv * 3 * NC + 3 * c + 2];
        }
        // Lagrange weights of the nodes x = -1, 0, 1 applied to the moments
        const double lam = q == 0 ? 0.5 * (m2 - m1) : (q == 1 ? m0 - m2 : 0.5 * (m2 + m1));
        if (lam != 0.0) {
          const CompTab& ct = comp[c];
          const ChanTab& ch = chan[c];
          const double t_lo = fma(ct.t_step, kt_mid, ct.t_start) - ch.mean_delay;
          const double t_hi = fma(ct.t_step, tj_last, ct.t_start) - ch.mean_delay;
          const double t_node = q == 0 ? t_lo : (q == 2 ? t_hi : 0.5 * (t_lo + t_hi));
          D2Ctx ctx;
          make_d2ctx(A, ch, ct, ch.amp, 1.0, t_node, 0.0, chan[0].amp_pbf, ctx, true);
          accumulate_terms(ctx, lam, H.need[c], reinterpret_cast<double (&)[kNumTerms]>(acc));
          my_c = c;
        }
      }
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        if (__ballot_sync(kFullMask, my_c == c) == 0u) continue;  // warp-uniform
        const bool mine = my_c == c;
        double r16[16], r8[8], r4[4], r2[2];
        const bool h16 = (lane & 16) != 0, h8 = (lane & 8) != 0, h4 = (lane & 4) != 0, h2 = (lane & 2) != 0,
                   h1 = (lane & 1) != 0;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const double lo = mine ? acc[k] : 0.0, hi = mine ? acc[k + 16] : 0.0;
          r16[k] = (h16 ? hi : lo) + __shfl_xor_sync(kFullMask, h16 ? lo : hi, 16);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) r8[k] = (h8 ? r16[k + 8] : r16[k]) + __shfl_xor_sync(kFullMask, h8 ? r16[k] : r16[k + 8], 8);
#pragma unroll
        for (int k = 0; k < 4; ++k) r4[k] = (h4 ? r8[k + 4] : r8[k]) + __shfl_xor_sync(kFullMask, h4 ? r8[k] : r8[k + 4], 4);
#pragma unroll
        for (int k = 0; k < 2; ++k) r2[k] = (h2 ? r4[k + 2] : r4[k]) + __shfl_xor_sync(kFullMask, h2 ? r4[k] : r4[k + 2], 2);
        tot[c] += (h1 ? r2[1] : r2[0]) + __shfl_xor_sync(kFullMask, h1 ? r2[0] : r2[1], 1);
      }
    }
    // channel flush: warp totals -> block totals (warp order) -> pair accumulators
#pragma unroll
    for (int c = 0; c < NC; ++c) part[(c * kFastWarps + warp) * 32 + lane] = tot[c];
    __syncthreads();
    if (tid < 32 * NC) {
      const int c = tid >> 5, k = tid & 31;
      double sum = 0.0;
      for (int wv = 0; wv < kFastWarps; ++wv) sum += part[(c * kFastWarps + wv) * 32 + k];
      total[c * 32 + k] = sum;
    }
    __syncthreads();
    for (int c = 0; c < NC; ++c) {
      const double lf = chan[c].logf;
      const int lo = H.comp_off[c], cnt = H.comp_off[c + 1] - lo;
      for (int p = tid; p < cnt; p += kFastBlock) {
        const int pair = H.comp_list[lo + p];
        const HessDesc d = H.desc[pair];
        const double factor = d.code == 0 ? 1.0 : (d.code == 1 ? kLn10 : (d.code == 2 ? lf : lf * lf));
        pacc[pair] += factor * total[c * 32 + d.idx];
      }
      __syncthreads();
    }
  }
  cp_async_wait<0>();
  __syncthreads();
  for (int e = tid; e < npairs; e += kFastBlock) H.partials[(size_t)blockIdx.x * npairs + e] = pacc[e];
}

// One warp per pair: lanes stride over the blocks, fixed-shape shuffle tree (deterministic).
__global__ void __launch_bounds__(128) k_hess_finalize(const double* __restrict__ partials, int nblocks, int npairs,
                                                      double* __restrict__ out) {
  const int pi = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (pi >= npairs) return;
  double s = 0.0;
  for (int b = lane; b < nblocks; b += 32) s += partials[(size_t)b * npairs + pi];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) out[pi] = s;
}

}  // namespace fb
