// fitburst_b200 kernels: one thread per output sample, one block per (channel, 256-sample
// tile) work item, channel tables staged in shared memory, [J r]^T [J r] accumulated in
// registers through a shared-memory row stage and reduced deterministically in two stages.
#pragma once
#include "fb_core.cuh"

namespace fb {

constexpr int kBlock = 128;
constexpr int kTile = 4;  // register tile edge of the Gram accumulation

enum EvalMode { MODE_MODEL = 0, MODE_GRAM = 1, MODE_JAC = 2, MODE_DERIV = 3, MODE_CHI2 = 4 };

struct EvalArgs {
  PlanConst pc;
  const double* params;  // device, (FB_NUM_PARAMETERS, num_comp)
  // tables built by k_tables for the current parameters, indexed by channel number
  const CompTab* g_comp;  // (num_comp)
  const SubTab* g_sub;    // (num_freq, num_comp, kf)
  const ChanTab* g_chan;  // (num_freq, num_comp)
  const int* chan_list;  // channels to visit (nullptr = all)
  int num_chan;
  const double* data;     // (num_freq, num_time) or nullptr
  const double* weights;  // (num_freq) or nullptr
  // free-parameter layout (LSFitter.get_fit_parameters_list order, analysis/fitter.py:353-384)
  int n_free;
  int col_of[9];  // first packed column of parameter p, -1 if fixed
  // MODE_MODEL outputs
  double* model;
  double* c_amp;
  double* c_spec;
  double* c_td;
  double* c_tp;
  // MODE_GRAM output: per-block partial tiles
  double* partials;
  // MODE_JAC outputs
  double* jac;
  double* resid;
  // MODE_DERIV
  int d_param, d_comp, d_addall;
  double* d_out;
  int* err_flag;  // set to 1 on a singular scintillation system (routines/ism.py:46)
  // MODE_CHI2: per-block partial sums of ((data - model) w)^2
  double* chi2_partials;
  // values stored per (sample, component): 3 = spectrum, timediff, timeprof; 2 = without timeprof
  // (only the caches and the scintillation solve need the mean profile). 0 means 3.
  // Scintillation mode with 2: slot 0 holds the mean PROFILE instead (the spectrum there is
  // amplitude * profile with the amplitude of the per-channel solve: analysis/model.py:274-277).
  int vals_spc;
  // k_gram3: the spectrum-row / data tiles are fetched with cp.async.bulk (one elected lane, completion
  // on an mbarrier per ring slot) instead of one 8-byte cp.async per lane; needs an even num_time
  int bulk_rows;
  // scintillation: per-(channel, component) amplitudes solved beforehand (k_scint_partials +
  // k_scint_solve), (num_freq, num_comp). nullptr: a block owns whole channels and solves in place.
  const double* scint_amps;
};

__host__ __device__ inline bool is_fit_global(int p) {
  // LSFitter.global_parameters, analysis/fitter.py:72
  return p == FB_DM || p == FB_DM_INDEX || p == FB_SCATTERING_TIMESCALE || p == FB_SCATTERING_INDEX;
}

__host__ __device__ inline int gram_side(int ncols) { return (ncols + kTile - 1) / kTile; }
__host__ __device__ inline int gram_ntiles(int ncols) {
  const int s = gram_side(ncols);
  return s * (s + 1) / 2;
}
__host__ __device__ inline int gram_stride(int ncols) { return (gram_side(ncols) * kTile) | 1; }

// Maps a linear upper-triangle tile index to (ti <= tj).
__device__ inline void tile_coords(int t, int side, int& ti, int& tj) {
  int row = 0, rem = t;
  while (rem >= side - row) {
    rem -= side - row;
    ++row;
  }
  ti = row;
  tj = row + rem;
}

// Accumulates X^T X over `rows` staged rows. Thread layout: tile = tid % ntiles,
// slice = tid / ntiles; each thread owns a kTile x kTile register block.
__device__ __forceinline__ void gram_accumulate(const double* __restrict__ X, int stride, int rows,
                                                int ntiles, int nslices, int ti, int tj, int slice,
                                                bool active, double (&acc)[kTile * kTile]) {
  if (!active) return;
  const double* xa = X + ti * kTile;
  const double* xb = X + tj * kTile;
#pragma unroll 2
  for (int s = slice; s < rows; s += nslices) {
    const double* ra = xa + s * stride;
    const double* rb = xb + s * stride;
    double a[kTile], b[kTile];
#pragma unroll
    for (int k = 0; k < kTile; ++k) {
      a[k] = ra[k];
      b[k] = rb[k];
    }
#pragma unroll
    for (int r = 0; r < kTile; ++r)
#pragma unroll
      for (int q = 0; q < kTile; ++q) acc[r * kTile + q] = fma(a[r], b[q], acc[r * kTile + q]);
  }
}

struct SmemLayout;
// Sums the per-slice register tiles of a block in slice order (deterministic) and leaves
// tile-major sums out[tile * 16 + k] in `out` (shared or global).
__device__ inline void gram_block_reduce(double* scratch, int ntiles, int nslices, int tile,
                                         int slice, bool active, const double (&acc)[kTile * kTile],
                                         double* out) {
  constexpr int kk = kTile * kTile;
  __syncthreads();
  if (active) {
#pragma unroll
    for (int k = 0; k < kk; ++k) scratch[(slice * ntiles + tile) * kk + k] = acc[k];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < ntiles * kk; e += blockDim.x) {
    double s = 0.0;
    for (int sl = 0; sl < nslices; ++sl) s += scratch[sl * ntiles * kk + e];
    out[e] = s;
  }
  __syncthreads();
}

// Gaussian elimination with partial pivoting on an n x n system held in shared memory
// (n <= FB_MAX_COMPONENTS); one thread. Returns false on an exactly-zero pivot, the condition
// under which LAPACK gesv (np.linalg.solve, routines/ism.py:46) reports a singular matrix.
__device__ inline bool solve_small(double* a, double* b, int n, int ld) {
  for (int k = 0; k < n; ++k) {
    int piv = k;
    double best = fabs(a[k * ld + k]);
    for (int r = k + 1; r < n; ++r)
      if (fabs(a[r * ld + k]) > best) {
        best = fabs(a[r * ld + k]);
        piv = r;
      }
    if (best == 0.0 || best != best) return false;
    if (piv != k) {
      for (int c = 0; c < n; ++c) {
        const double t = a[k * ld + c];
        a[k * ld + c] = a[piv * ld + c];
        a[piv * ld + c] = t;
      }
      const double t = b[k];
      b[k] = b[piv];
      b[piv] = t;
    }
    for (int r = k + 1; r < n; ++r) {
      const double f = a[r * ld + k] / a[k * ld + k];
      for (int c = k + 1; c < n; ++c) a[r * ld + c] -= f * a[k * ld + c];
      b[r] -= f * b[k];
    }
  }
  for (int k = n - 1; k >= 0; --k) {
    double s = b[k];
    for (int c = k + 1; c < n; ++c) s -= a[k * ld + c] * b[c];
    b[k] = s / a[k * ld + k];
  }
  return true;
}

struct SmemLayout {
  CompTab* comp;
  ChanTab* chan;
  SubTab* sub;
  double* scint_amp;
  double* small;    // (nc+1)^2 + nc scratch for the scintillation solve
  double* tilesum;  // block-summed Gram tiles of the scintillation system
  double* park;     // Gram register tiles parked between tiles: [kTile*kTile][kBlock]
  double* vals;     // per-thread (spectrum, timediff, timeprof) per component, stride val_stride
  int* queue;       // (thread, component) items of the current tile that need the general case
  int* qcount;
  double* X;
};

__host__ __device__ inline int val_stride(int nc) { return (3 * nc) | 1; }

// shared-memory regions a kernel may ask for
enum SmemRegion { SM_VALS = 1, SM_PARK = 2, SM_X = 4, SM_QUEUE = 8, SM_ALL = 15 };

__host__ __device__ inline size_t smem_bytes(int nc, int kf, int ncols_main, bool scint, bool lean = false,
                                            int regions = SM_ALL) {
  if (lean) regions &= ~(SM_VALS | SM_PARK);
  size_t n = 0;
  n += sizeof(CompTab) * nc;
  n += sizeof(ChanTab) * nc;
  n += sizeof(SubTab) * nc * kf;
  n += sizeof(double) * nc;
  n += sizeof(double) * ((nc + 1) * (nc + 1) + nc + 2);
  n += sizeof(double) * gram_ntiles(nc + 1) * kTile * kTile;
  if (regions & SM_VALS) n += sizeof(double) * (size_t)kBlock * val_stride(nc);
  if (regions & SM_PARK) n += sizeof(double) * (size_t)kBlock * kTile * kTile;
  if (regions & SM_QUEUE) n += sizeof(int) * ((size_t)kBlock * nc + 2);
  if (!(regions & SM_X)) return n;
  int stride = gram_stride(ncols_main);
  if (scint) {
    const int s2 = gram_stride(nc + 1);
    stride = stride > s2 ? stride : s2;
  }
  size_t xr = (size_t)kBlock * stride;
  if (xr < (size_t)kBlock * kTile * kTile) xr = (size_t)kBlock * kTile * kTile;
  n += sizeof(double) * xr;
  return n;
}

__device__ inline SmemLayout carve(double* base, int nc, int kf, bool lean = false, int regions = SM_ALL) {
  if (lean) regions &= ~(SM_VALS | SM_PARK);
  SmemLayout L;
  char* p = reinterpret_cast<char*>(base);
  L.comp = reinterpret_cast<CompTab*>(p);
  p += sizeof(CompTab) * nc;
  L.chan = reinterpret_cast<ChanTab*>(p);
  p += sizeof(ChanTab) * nc;
  L.sub = reinterpret_cast<SubTab*>(p);
  p += sizeof(SubTab) * nc * kf;
  L.scint_amp = reinterpret_cast<double*>(p);
  p += sizeof(double) * nc;
  L.small = reinterpret_cast<double*>(p);
  p += sizeof(double) * ((nc + 1) * (nc + 1) + nc + 2);
  L.tilesum = reinterpret_cast<double*>(p);
  p += sizeof(double) * gram_ntiles(nc + 1) * kTile * kTile;
  L.vals = reinterpret_cast<double*>(p);
  if (regions & SM_VALS) p += sizeof(double) * (size_t)kBlock * val_stride(nc);
  L.park = reinterpret_cast<double*>(p);
  if (regions & SM_PARK) p += sizeof(double) * (size_t)kBlock * kTile * kTile;
  L.queue = reinterpret_cast<int*>(p);
  L.qcount = L.queue + (size_t)kBlock * nc;
  if (regions & SM_QUEUE) p += sizeof(int) * ((size_t)kBlock * nc + 2);
  L.X = reinterpret_cast<double*>(p);
  return L;
}

// Gram accumulation of one staged tile with the register tile parked in shared memory between
// tiles ([k][tid] layout, conflict-free), so that it is not live across the model evaluation.
__device__ __forceinline__ void gram_accumulate_parked(const SmemLayout& L, int stride, int ntiles, int nslices,
                                                       int ti, int tj, int slice, bool active) {
  if (!active) return;
  double acc[kTile * kTile];
  double* park = L.park + threadIdx.x;
#pragma unroll
  for (int k = 0; k < kTile * kTile; ++k) acc[k] = park[k * kBlock];
  gram_accumulate(L.X, stride, kBlock, ntiles, nslices, ti, tj, slice, active, acc);
#pragma unroll
  for (int k = 0; k < kTile * kTile; ++k) park[k * kBlock] = acc[k];
}

// Builds every table of the current parameter set: one thread per (channel, component).
__global__ void __launch_bounds__(128) k_tables(const PlanConst pc, const double* __restrict__ params,
                                                CompTab* __restrict__ g_comp, SubTab* __restrict__ g_sub,
                                                ChanTab* __restrict__ g_chan) {
  if (fit_stopped(pc)) return;
  const int nc = pc.num_comp;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= pc.num_freq * nc) return;
  const int i = e / nc, c = e % nc;
  CompTab ct;
  make_comp_tab(pc, params, c, ct);
  if (i == 0) g_comp[c] = ct;
  make_channel_tables(pc, params, ct, i, c, g_sub + (size_t)e * pc.kf, g_chan[e]);
}

// Same tables with one thread per (channel, component, SUB-FREQUENCY): k_tables above is a
// latency-bound chain of ~8 transcendental calls per sub-frequency per thread (49 k threads at
// config 2, 50 us); spreading the sub-frequencies over the lanes of a KF-wide group gives 8x the
// parallelism, and the group leader reads the per-lane results back in sub-frequency order, so
// every sum is formed exactly like in the serial loop. Visits only the listed channels.
template <int KF>
__global__ void __launch_bounds__(128) k_tables_group(const PlanConst pc, const double* __restrict__ params,
                                                      CompTab* __restrict__ g_comp, SubTab* __restrict__ g_sub,
                                                      ChanTab* __restrict__ g_chan, const int* __restrict__ chan_list,
                                                      int num_chan) {
  // ncu (round 2): 2.3e7 warp instructions, 23 of 32 lanes active -- every thread repeated the
  // per-component table (a pow), and the channel-level part ran on one lane of eight. Now the
  // component tables are made once per block and the channel-level part of the block's 128 / KF
  // (channel, component) items is finished by its first 128 / KF threads together: 1.6e7
  // instructions, 29 lanes active. The kernel stays latency-bound (42 us either way: chains of
  // pow / exp / erfcx calls at 20 warps per SM; finishing in a second, full-warp kernel measured
  // 34 + 13 us) -- the saving is issue slots for the kernels of the other fits in flight.
  constexpr int kGroups = 128 / KF;
  __shared__ CompTab s_ct[FB_MAX_COMPONENTS];
  __shared__ SubSummary s_sm[kGroups];
  __shared__ ChanPowers s_cp[kGroups];
  if (fit_stopped(pc)) return;
  const int nc = pc.num_comp;
  if ((int)threadIdx.x < nc) {
    CompTab t;
    make_comp_tab(pc, params, threadIdx.x, t);
    s_ct[threadIdx.x] = t;
    if (blockIdx.x == 0) g_comp[threadIdx.x] = t;
  }
  __syncthreads();
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int a = (int)(gid % KF), lane = threadIdx.x & 31, group = threadIdx.x / KF;
  long long item = gid / KF;
  const bool active = item < (long long)num_chan * nc;
  if (!active) item = 0;  // keeps the lane in the shuffles below
  const int chi = (int)(item / nc), c = (int)(item % nc);
  const int i = chan_list ? chan_list[chi] : chi;
  const CompTab& ct = s_ct[c];
  const int base = lane & ~(KF - 1);
  // the four channel-level powers, one per lane of the group (same pow calls as make_chan_powers)
  ChanPowers cp;
  {
    const double beta = global_param(params, nc, FB_DM_INDEX), alpha = global_param(params, nc, FB_SCATTERING_INDEX);
    const double fc = pc.freqs[i], fr = fc / ct.ref_freq;
    double vals[4];
#pragma unroll
    for (int round = 0; round < (4 + KF - 1) / KF; ++round) {
      const int q = round * KF + a;  // which power this lane computes in this round
      const double x = q == 0 ? ct.ref_freq : (q == 1 ? fc : fr);
      const double y = q <= 1 ? beta : (q == 2 ? alpha : -alpha);
      const double v = pow(x, y);
#pragma unroll
      for (int k = 0; k < KF; ++k)
        if (round * KF + k < 4) vals[round * KF + k] = __shfl_sync(0xffffffffu, v, base + k);
    }
    cp.fref_b = vals[0];
    cp.fc_b = vals[1];
    cp.fr_alpha = vals[2];
    cp.fr_nalpha = vals[3];
  }
  SubTab st;
  double p, sc_time;
  make_sub_tab(pc, params, ct, i, c, a, cp.fref_b, cp.fc_b, st, p, sc_time);
  if (active) g_sub[((size_t)i * nc + c) * KF + a] = st;
  SubSummary sm;
  sm.init();
#pragma unroll
  for (int k = 0; k < KF; ++k) {
    const double delay = __shfl_sync(0xffffffffu, st.delay, base + k);
    const double ratio = __shfl_sync(0xffffffffu, st.ratio, base + k);
    const double sed = __shfl_sync(0xffffffffu, st.sed, base + k);
    const double pk = __shfl_sync(0xffffffffu, p, base + k);
    const double sck = __shfl_sync(0xffffffffu, sc_time, base + k);
    if (sck > 0.0) sm.scattered = 1;
    sm.add(delay, ratio, sed, ct.w, pk);
  }
  if (a == 0) {
    s_sm[group] = sm;
    s_cp[group] = cp;
  }
  __syncthreads();
  if ((int)threadIdx.x < kGroups) {
    const long long item2 = (long long)blockIdx.x * kGroups + threadIdx.x;
    if (item2 < (long long)num_chan * nc) {
      const int chi2 = (int)(item2 / nc), c2 = (int)(item2 % nc);
      const int i2 = chan_list ? chan_list[chi2] : chi2;
      finish_chan_tab(pc, params, s_ct[c2], i2, c2, s_sm[threadIdx.x], s_cp[threadIdx.x], g_chan[(size_t)i2 * nc + c2]);
    }
  }
}

// Stages one channel's tables from global memory (L2-resident, written by k_tables).
__device__ inline void load_channel_tables(const EvalArgs& A, const SmemLayout& L, int i) {
  const PlanConst& pc = A.pc;
  const int nc = pc.num_comp, kf = pc.kf;
  __syncthreads();
  {
    const double* src = reinterpret_cast<const double*>(A.g_sub + (size_t)i * nc * kf);
    double* dst = reinterpret_cast<double*>(L.sub);
    const int nd = nc * kf * (int)(sizeof(SubTab) / sizeof(double));
    for (int e = threadIdx.x; e < nd; e += blockDim.x) dst[e] = src[e];
  }
  {
    const double* src = reinterpret_cast<const double*>(A.g_chan + (size_t)i * nc);
    double* dst = reinterpret_cast<double*>(L.chan);
    const int nd = nc * (int)(sizeof(ChanTab) / sizeof(double));
    for (int e = threadIdx.x; e < nd; e += blockDim.x) dst[e] = src[e];
  }
  if (pc.scintillation && A.scint_amps != nullptr && (int)threadIdx.x < nc)
    L.scint_amp[threadIdx.x] = A.scint_amps[(size_t)i * nc + threadIdx.x];
  __syncthreads();
}

__device__ inline void load_comp_tables(const EvalArgs& A, const SmemLayout& L) {
  const double* src = reinterpret_cast<const double*>(A.g_comp);
  double* dst = reinterpret_cast<double*>(L.comp);
  const int nd = A.pc.num_comp * (int)(sizeof(CompTab) / sizeof(double));
  for (int e = threadIdx.x; e < nd; e += blockDim.x) dst[e] = src[e];
  __syncthreads();
}

// Evaluates (spectrum, timediff, timeprof) of every component for the kBlock samples of one time
// tile of channel i into L.vals. Samples on the clamp plateau or in the exponential tail are
// closed-form and done by their own thread; the few samples per (channel, component) that need
// the general exp + erfcx sub-grid are pushed on a block-wide queue and evaluated one sub-sample
// per LANE, so that the expensive work is spread over the whole block instead of stalling
// every other warp at the next barrier.
__device__ inline void eval_tile(const EvalArgs& A, const SmemLayout& L, int i, int tile,
                                 double* vals_out = nullptr, int soa = 0) {
  const PlanConst& pc = A.pc;
  const int nc = pc.num_comp, kf = pc.kf, kt = pc.kt, tid = threadIdx.x;
  const int vs = val_stride(nc);
  const int npe = kf * kt;
  const bool use_queue = npe >= 8;
  const int j = tile * kBlock + tid;
  if (tid == 0) *L.qcount = 0;
  __syncthreads();
  // value (thread t, slot k) lives at vals[t * ts + k * ss]: per-thread rows in shared memory,
  // slot-major planes (coalesced) when the values go to global memory
  double* vals = vals_out ? vals_out : L.vals;
  const int ts = soa ? 1 : vs, ss = soa ? kBlock : 1;
  const int spc = A.vals_spc == 2 ? 2 : 3;
  const bool prof_first = spc == 2 && pc.scintillation != 0;  // slot 0 = mean profile
  double* v = vals + (size_t)tid * ts;
  if (j < pc.num_time) {
    for (int c = 0; c < nc; ++c) {
      SampleComp sc;
      if (eval_sample_fast(pc, L.comp[c], L.chan[c], L.sub + c * kf, j, sc)) {
        v[(spc * c) * ss] = prof_first ? sc.timeprof : sc.spectrum;
        v[(spc * c + 1) * ss] = sc.timediff;
        if (spc == 3) v[(spc * c + 2) * ss] = sc.timeprof;
      } else if (use_queue) {
        L.queue[atomicAdd(L.qcount, 1)] = (tid << 5) | c;
      } else {
        sc = eval_sample_general(pc, L.comp[c], L.chan[c], L.sub + c * kf, j);
        v[(spc * c) * ss] = prof_first ? sc.timeprof : sc.spectrum;
        v[(spc * c + 1) * ss] = sc.timediff;
        if (spc == 3) v[(spc * c + 2) * ss] = sc.timeprof;
      }
    }
  }
  __syncthreads();
  if (!use_queue) return;
  const int qn = *L.qcount;
  if (qn > 0) {
    // group of gs lanes per item (gs = power of two >= min(npe, 32)), 32/gs items per warp pass.
    int gs = 8;
    while (gs < npe && gs < 32) gs <<= 1;
    const int lane = tid & 31, warp = tid >> 5, nwarps = kBlock / 32;
    const int gpw = 32 / gs, g = lane / gs, gl = lane % gs;
    const double inv_n = 1.0 / (double)npe;
    for (int base = warp * gpw; base < qn; base += nwarps * gpw) {
      const int q = base + g;
      const bool active = q < qn;
      // only the SED-weighted profile sum needs the cross-lane reduction: the mean time
      // difference has the same closed form as in the fast regimes, and the plain profile mean
      // is wanted only when the caches / the scintillation solve ask for it (spc == 3)
      double s_p = 0.0, s_ps = 0.0;
      int t_idx = 0, c = 0;
      if (active) {
        const int item = L.queue[q];
        t_idx = item >> 5;
        c = item & 31;
        const CompTab& ct = L.comp[c];
        const SubTab* sub = L.sub + c * kf;
        const int scattered = L.chan[c].scattered;
        const int jj = tile * kBlock + t_idx;
        // sub-sample pe -> (sub-time b, sub-frequency a) without an integer division per step
        int b = gl / kf, a = gl - b * kf;
        const int db = gs / kf, da = gs - db * kf;
        for (int pe = gl; pe < npe; pe += gs) {
          double t, p;
          eval_subsample(pc, ct, scattered, sub[a], jj * kt + b, t, p);
          if (spc == 3 || prof_first) s_p += p;
          s_ps = fma(p, sub[a].sed, s_ps);
          b += db;
          a += da;
          if (a >= kf) {
            a -= kf;
            ++b;
          }
        }
      }
      for (int off = gs >> 1; off > 0; off >>= 1) {
        if (spc == 3 || prof_first) s_p += __shfl_xor_sync(0xffffffffu, s_p, off);
        s_ps += __shfl_xor_sync(0xffffffffu, s_ps, off);
      }
      if (active && gl == 0) {
        const CompTab& ct = L.comp[c];
        const int jj = tile * kBlock + t_idx;
        double* o = vals + (size_t)t_idx * ts + (size_t)(spc * c) * ss;
        o[0] = prof_first ? s_p * inv_n : ct.amp10 * (s_ps * inv_n);
        o[ss] = fma(ct.t_step, (double)(jj * kt) + 0.5 * (double)(kt - 1), ct.t_start) - L.chan[c].mean_delay;
        if (spc == 3) o[2 * ss] = s_p * inv_n;
      }
    }
  }
  __syncthreads();
}

// Everything a thread does for one sample of one channel, for every mode.
template <int MODE>
__device__ __forceinline__ void process_sample(const EvalArgs& A, const SmemLayout& L, int i, int j,
                                               bool valid, double* row, int stride,
                                               const double* vrow_in = nullptr, int vss = 1, int rs = 1,
                                               const double* data_staged = nullptr) {
  const PlanConst& pc = A.pc;
  const int nc = pc.num_comp, kf = pc.kf, n = A.n_free;
  const double wi = (A.weights != nullptr) ? A.weights[i] : 1.0;
  const double tau0 = global_param(A.params, nc, FB_SCATTERING_TIMESCALE);
  const double* vrow = vrow_in ? vrow_in : L.vals + (size_t)threadIdx.x * val_stride(nc);
  if (MODE == MODE_GRAM || MODE == MODE_JAC) {
    if (!valid) {
      for (int k = 0; k <= n; ++k) row[k * rs] = 0.0;
    } else {
      // per-component columns and the residual are assigned below; only the columns that
      // accumulate over components start from zero
#pragma unroll
      for (int p = 0; p < 9; ++p)
        if (is_fit_global(p) && A.col_of[p] >= 0) row[A.col_of[p] * rs] = 0.0;
    }
    if (MODE == MODE_GRAM)
      for (int k = n + 1; k < stride; ++k) row[k * rs] = 0.0;
  }
  if (MODE == MODE_CHI2) row[0] = 0.0;
  if (!valid) return;
  double msum = 0.0, dsum = 0.0;
  const size_t s_idx = (size_t)i * pc.num_time + j;
  for (int c = 0; c < nc; ++c) {
    const CompTab& ct = L.comp[c];
    const ChanTab& ch = L.chan[c];
    SampleComp sc;
    const int spc = A.vals_spc == 2 ? 2 : 3;
    sc.spectrum = vrow[(spc * c) * vss];
    sc.timediff = vrow[(spc * c + 1) * vss];
    sc.timeprof = spc == 3 ? vrow[(spc * c + 2) * vss] : (pc.scintillation ? sc.spectrum : 0.0);
    double amp = ch.amp;
    if (pc.scintillation) {  // analysis/model.py:274-277
      amp = L.scint_amp[c];
      sc.spectrum = amp * sc.timeprof;
    }
    msum += sc.spectrum;
    if (MODE == MODE_CHI2) continue;
    if (MODE == MODE_MODEL) {
      const size_t o = s_idx * nc + c;
      if (A.c_amp) A.c_amp[o] = amp;
      if (A.c_spec) A.c_spec[o] = sc.spectrum;
      if (A.c_td) A.c_td[o] = sc.timediff;
      if (A.c_tp) A.c_tp[o] = sc.timeprof;
    } else {
      double d[9], g;
      if (MODE == MODE_DERIV) {
        // the reference looks deriv_amplitude_pbf up with the requested component.
        first_derivatives(ch, L.chan[A.d_comp], ct, tau0, amp, sc.spectrum, sc.timediff, d, &g);
        const bool all = A.d_addall && is_fit_global(A.d_param);
        if (all || c == A.d_comp) dsum += d[A.d_param];
      } else {
        // LSFitter.compute_jacobian calls the global derivatives with component = 0
        // (analysis/fitter.py:213-223), per-component ones with their own index.
        first_derivatives(ch, ch, ct, tau0, amp, sc.spectrum, sc.timediff, d, &g);
        // the scattering_* columns look deriv_amplitude_pbf up with component 0: correct them
        // when that differs from the own-component value (only if ref_freq / width differ).
        const ChanTab& ch0 = L.chan[0];
        d[FB_SCATTERING_TIMESCALE] += sc.spectrum * (ch0.dap_tau - ch.dap_tau) * ch.inv_apbf;
        d[FB_SCATTERING_INDEX] += sc.spectrum * (ch0.dap_alpha - ch.dap_alpha) * ch.inv_apbf;
#pragma unroll
        for (int p = 0; p < 9; ++p) {
          const int col = A.col_of[p];
          if (col < 0) continue;
          if (is_fit_global(p)) row[col * rs] += -d[p] * wi;
          else row[(col + c) * rs] = -d[p] * wi;
        }
      }
    }
  }
  if (MODE == MODE_MODEL) {
    A.model[s_idx] = msum;
  } else if (MODE == MODE_CHI2) {
    const double r = (A.data[s_idx] - msum) * wi;
    row[0] = r * r;
  } else if (MODE == MODE_DERIV) {
    A.d_out[s_idx] = dsum;
  } else {
    const double dval = data_staged ? *data_staged : A.data[s_idx];
    const double r = (dval - msum) * wi;  // analysis/fitter.py:262-264
    row[n * rs] = r;
    if (MODE == MODE_JAC) {
      if (A.jac)
        for (int k = 0; k < n; ++k) A.jac[s_idx * n + k] = row[k * rs];
      if (A.resid) A.resid[s_idx] = r;
    }
  }
}

template <int MODE>
__device__ __forceinline__ void eval_body(const EvalArgs& A, const int block_id, const int num_blocks) {
  extern __shared__ double smem_raw[];
  const PlanConst& pc = A.pc;
  const int nc = pc.num_comp, kf = pc.kf, tid = threadIdx.x;
  SmemLayout L = carve(smem_raw, nc, kf);
  load_comp_tables(A, L);

  const int ncols = A.n_free + 1;
  const int stride = gram_stride(ncols);
  const int side = gram_side(ncols);
  const int ntiles = gram_ntiles(ncols);
  const int nslices = kBlock / ntiles;
  const int g_tile = tid % ntiles, g_slice = tid / ntiles;
  const bool g_active = (MODE == MODE_GRAM) && g_slice < nslices;
  int g_ti = 0, g_tj = 0;
  if (MODE == MODE_GRAM) tile_coords(g_tile, side, g_ti, g_tj);
#pragma unroll
  for (int k = 0; k < kTile * kTile; ++k) L.park[k * kBlock + tid] = 0.0;

  const int tiles = (pc.num_time + kBlock - 1) / kBlock;
  double* row = L.X + (size_t)tid * stride;
  double chi2_acc = 0.0;

  if (!pc.scintillation || A.scint_amps != nullptr) {
    const long long total = (long long)A.num_chan * tiles;
    const long long begin = total * block_id / num_blocks;
    const long long end = total * (block_id + 1) / num_blocks;
    int cur = -1;
    for (long long item = begin; item < end; ++item) {
      const int chi = (int)(item / tiles), tile = (int)(item % tiles);
      const int i = A.chan_list ? A.chan_list[chi] : chi;
      if (chi != cur) {
        load_channel_tables(A, L, i);
        cur = chi;
      }
      const int j = tile * kBlock + tid;
      eval_tile(A, L, i, tile);
      process_sample<MODE>(A, L, i, j, j < pc.num_time, row, stride);
      if (MODE == MODE_CHI2) chi2_acc += row[0];
      if (MODE == MODE_GRAM) {
        __syncthreads();
        gram_accumulate_parked(L, stride, ntiles, nslices, g_ti, g_tj, g_slice, g_active);
        __syncthreads();
      }
    }
  } else {
    // Scintillation: per-channel amplitudes come from a least-squares solve over ALL samples of
    // the channel (routines/ism.py:11-48), so a block owns whole channels and makes two passes
    // (batched small bursts; single spectra take the amplitudes from k_scint_solve instead).
    const int ncols2 = nc + 1;
    const int stride2 = gram_stride(ncols2);
    const int side2 = gram_side(ncols2);
    const int ntiles2 = gram_ntiles(ncols2);
    const int nslices2 = kBlock / ntiles2;
    const int s_tile = tid % ntiles2, s_slice = tid / ntiles2;
    const bool s_active = s_slice < nslices2;
    int s_ti, s_tj;
    tile_coords(s_tile, side2, s_ti, s_tj);
    double* row2 = L.X + (size_t)tid * stride2;
    for (int chi = block_id; chi < A.num_chan; chi += num_blocks) {
      const int i = A.chan_list ? A.chan_list[chi] : chi;
      load_channel_tables(A, L, i);
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
      // block sums land tile-major in X, then thread 0 assembles and solves.
      double* sums = L.tilesum;
      gram_block_reduce(L.X, ntiles2, nslices2, s_tile, s_slice, s_active, acc2, sums);
      if (tid == 0) {
        double* G = L.small;            // nc x nc
        double* b = L.small + nc * nc;  // nc
        for (int t = 0; t < ntiles2; ++t) {
          int ti, tj;
          tile_coords(t, side2, ti, tj);
          for (int r = 0; r < kTile; ++r)
            for (int q = 0; q < kTile; ++q) {
              const int gr = ti * kTile + r, gc = tj * kTile + q;
              const double v = sums[t * kTile * kTile + r * kTile + q];
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
      for (int tile = 0; tile < tiles; ++tile) {
        const int j = tile * kBlock + tid;
        eval_tile(A, L, i, tile);
        process_sample<MODE>(A, L, i, j, j < pc.num_time, row, stride);
        if (MODE == MODE_CHI2) chi2_acc += row[0];
        if (MODE == MODE_GRAM) {
          __syncthreads();
          gram_accumulate_parked(L, stride, ntiles, nslices, g_ti, g_tj, g_slice, g_active);
          __syncthreads();
        }
      }
    }
  }
  if (MODE == MODE_CHI2) {
    // deterministic block tree-sum of the per-thread partial chi^2
    __syncthreads();
    L.X[tid] = chi2_acc;
    __syncthreads();
    for (int off = kBlock / 2; off > 0; off >>= 1) {
      if (tid < off) L.X[tid] += L.X[tid + off];
      __syncthreads();
    }
    if (tid == 0) A.chi2_partials[block_id] = L.X[0];
  }

  if (MODE == MODE_GRAM) {
    double acc[kTile * kTile];
#pragma unroll
    for (int k = 0; k < kTile * kTile; ++k) acc[k] = L.park[k * kBlock + tid];
    gram_block_reduce(L.X, ntiles, nslices, g_tile, g_slice, g_active, acc,
                      A.partials + (size_t)block_id * ntiles * kTile * kTile);
  }
}

// Scintillation amplitudes ahead of the evaluation (routines/ism.py:11-48), parallel over
// (channel, chunk of tiles) instead of one block per channel: a long channel (config 5: 262144
// samples) no longer runs on a single block, a short channel list no longer leaves SMs idle.
// k_scint_partials: block-summed Gram tiles of [timeprof_1 .. timeprof_nc, data] of one chunk,
// partials[(chi * nchunks + k)][ntiles2 * 64] in the tile-major layout of gram_block_reduce.
__global__ void __launch_bounds__(kBlock, 4) k_scint_partials(const EvalArgs A, int nchunks,
                                                              double* __restrict__ partials) {
  extern __shared__ double smem_raw[];
  const PlanConst& pc = A.pc;
  if (fit_stopped(pc)) return;
  const int nc = pc.num_comp, kf = pc.kf, tid = threadIdx.x;
  SmemLayout L = carve(smem_raw, nc, kf);
  load_comp_tables(A, L);
  const int tiles = (pc.num_time + kBlock - 1) / kBlock;
  const int ncols2 = nc + 1;
  const int stride2 = gram_stride(ncols2), side2 = gram_side(ncols2), ntiles2 = gram_ntiles(ncols2);
  const int nslices2 = kBlock / ntiles2;
  const int s_tile = tid % ntiles2, s_slice = tid / ntiles2;
  const bool s_active = s_slice < nslices2;
  int s_ti, s_tj;
  tile_coords(s_tile, side2, s_ti, s_tj);
  double* row2 = L.X + (size_t)tid * stride2;
  const long long total = (long long)A.num_chan * nchunks;
  int cur = -1;
  for (long long item = blockIdx.x; item < total; item += gridDim.x) {
    const int chi = (int)(item / nchunks), k = (int)(item % nchunks);
    const int i = A.chan_list ? A.chan_list[chi] : chi;
    if (chi != cur) {
      load_channel_tables(A, L, i);
      cur = chi;
    }
    const int tile_lo = (int)((long long)tiles * k / nchunks), tile_hi = (int)((long long)tiles * (k + 1) / nchunks);
    double acc2[kTile * kTile];
#pragma unroll
    for (int q = 0; q < kTile * kTile; ++q) acc2[q] = 0.0;
    for (int tile = tile_lo; tile < tile_hi; ++tile) {
      const int j = tile * kBlock + tid;
      eval_tile(A, L, i, tile);
      for (int q = 0; q < stride2; ++q) row2[q] = 0.0;
      if (j < pc.num_time) {
        const double* vrow = L.vals + (size_t)tid * val_stride(nc);
        for (int c = 0; c < nc; ++c) row2[c] = vrow[3 * c + 2];
        row2[nc] = A.data[(size_t)i * pc.num_time + j];
      }
      __syncthreads();
      gram_accumulate(L.X, stride2, kBlock, ntiles2, nslices2, s_ti, s_tj, s_slice, s_active, acc2);
      __syncthreads();
    }
    gram_block_reduce(L.X, ntiles2, nslices2, s_tile, s_slice, s_active, acc2,
                      partials + (size_t)item * ntiles2 * kTile * kTile);
  }
}

// The same partial systems from the values k_vals stored (split evaluation, vals_spc = 2 in
// scintillation mode: the mean profile is slot 2 c of item chi * tiles + tile): no second
// evaluation of the profiles.
__global__ void __launch_bounds__(kBlock, 4) k_scint_partials_vals(const EvalArgs A, int nchunks,
                                                                   const double* __restrict__ gvals,
                                                                   double* __restrict__ partials) {
  extern __shared__ double smem_raw[];
  const PlanConst& pc = A.pc;
  if (fit_stopped(pc)) return;
  const int nc = pc.num_comp, tid = threadIdx.x;
  double* X = smem_raw;  // [kBlock][stride2] staged rows, then the scratch of the block reduction
  const int tiles = (pc.num_time + kBlock - 1) / kBlock;
  const int ncols2 = nc + 1;
  const int stride2 = gram_stride(ncols2), side2 = gram_side(ncols2), ntiles2 = gram_ntiles(ncols2);
  const int nslices2 = kBlock / ntiles2;
  const int s_tile = tid % ntiles2, s_slice = tid / ntiles2;
  const bool s_active = s_slice < nslices2;
  int s_ti, s_tj;
  tile_coords(s_tile, side2, s_ti, s_tj);
  double* row2 = X + (size_t)tid * stride2;
  const int vs = 2 * nc;
  const long long total = (long long)A.num_chan * nchunks;
  for (long long item = blockIdx.x; item < total; item += gridDim.x) {
    const int chi = (int)(item / nchunks), k = (int)(item % nchunks);
    const int i = A.chan_list ? A.chan_list[chi] : chi;
    const int tile_lo = (int)((long long)tiles * k / nchunks), tile_hi = (int)((long long)tiles * (k + 1) / nchunks);
    double acc2[kTile * kTile];
#pragma unroll
    for (int q = 0; q < kTile * kTile; ++q) acc2[q] = 0.0;
    for (int tile = tile_lo; tile < tile_hi; ++tile) {
      const int j = tile * kBlock + tid;
      const double* v = gvals + ((size_t)chi * tiles + tile) * kBlock * vs + tid;
      for (int q = 0; q < stride2; ++q) row2[q] = 0.0;
      if (j < pc.num_time) {
        for (int c = 0; c < nc; ++c) row2[c] = v[(size_t)(2 * c) * kBlock];
        row2[nc] = A.data[(size_t)i * pc.num_time + j];
      }
      __syncthreads();
      gram_accumulate(X, stride2, kBlock, ntiles2, nslices2, s_ti, s_tj, s_slice, s_active, acc2);
      __syncthreads();
    }
    gram_block_reduce(X, ntiles2, nslices2, s_tile, s_slice, s_active, acc2,
                      partials + (size_t)item * ntiles2 * kTile * kTile);
  }
}

__host__ __device__ inline size_t scint_vals_smem_bytes(int nc) {
  size_t rows = (size_t)kBlock * gram_stride(nc + 1);
  if (rows < (size_t)kBlock * kTile * kTile) rows = (size_t)kBlock * kTile * kTile;
  return sizeof(double) * rows;
}

// k_scint_solve: one block per channel; chunk partials summed in chunk order, then the nc x nc
// system of routines/ism.py:40-46 by Gaussian elimination with partial pivoting (solve_small).
__global__ void __launch_bounds__(64) k_scint_solve(const EvalArgs A, int nchunks, const double* __restrict__ partials,
                                                    double* __restrict__ amps) {
  __shared__ double sums[((FB_MAX_COMPONENTS + 1 + kTile - 1) / kTile) * ((FB_MAX_COMPONENTS + 1 + kTile - 1) / kTile + 1) / 2 *
                         kTile * kTile];
  __shared__ double G[FB_MAX_COMPONENTS * FB_MAX_COMPONENTS];
  __shared__ double b[FB_MAX_COMPONENTS];
  if (fit_stopped(A.pc)) return;
  const int nc = A.pc.num_comp, chi = blockIdx.x, tid = threadIdx.x;
  const int i = A.chan_list ? A.chan_list[chi] : chi;
  const int side2 = gram_side(nc + 1), ntiles2 = gram_ntiles(nc + 1);
  const int nent = ntiles2 * kTile * kTile;
  for (int e = tid; e < nent; e += blockDim.x) {
    double acc = 0.0;
    for (int k = 0; k < nchunks; ++k) acc += partials[((size_t)chi * nchunks + k) * nent + e];
    sums[e] = acc;
  }
  __syncthreads();
  if (tid == 0) {
    for (int t = 0; t < ntiles2; ++t) {
      int ti, tj;
      tile_coords(t, side2, ti, tj);
      for (int r = 0; r < kTile; ++r)
        for (int q = 0; q < kTile; ++q) {
          const int gr = ti * kTile + r, gc = tj * kTile + q;
          const double v = sums[t * kTile * kTile + r * kTile + q];
          if (gr < nc && gc < nc) {
            G[gr * nc + gc] = v;
            G[gc * nc + gr] = v;
          } else if (gr < nc && gc == nc) {
            b[gr] = v;
          }
        }
    }
    const bool ok = solve_small(G, b, nc, nc);
    for (int c = 0; c < nc; ++c) amps[(size_t)i * nc + c] = ok ? b[c] : nan("");
    if (!ok && A.err_flag) atomicExch(A.err_flag, 1);
  }
}

template <int MODE>
__global__ void __launch_bounds__(kBlock, 4) k_eval(const EvalArgs A) {
  if (fit_stopped(A.pc)) return;
  eval_body<MODE>(A, blockIdx.x, gridDim.x);
}

// Batched variant (many independent bursts of identical shape, blockIdx.y = burst): every
// per-burst buffer is offset by the burst index; bursts whose fit has finished are skipped.
struct BatchArgs {
  const int* done;       // (count) 1 once the burst's fit has terminated
  const int* num_good;   // (count) channels with non-zero weight
  size_t data_stride, weights_stride, params_stride, comp_stride, sub_stride, chan_stride, list_stride,
      partials_stride;
};

template <int MODE>
__global__ void __launch_bounds__(kBlock, 4) k_eval_batched(const EvalArgs A, const BatchArgs B) {
  const int y = blockIdx.y;
  if (B.done[y]) return;
  EvalArgs Ab = A;
  Ab.data = A.data + y * B.data_stride;
  Ab.weights = A.weights + y * B.weights_stride;
  Ab.params = A.params + y * B.params_stride;
  Ab.g_comp = A.g_comp + y * B.comp_stride;
  Ab.g_sub = A.g_sub + y * B.sub_stride;
  Ab.g_chan = A.g_chan + y * B.chan_stride;
  Ab.chan_list = A.chan_list + y * B.list_stride;
  Ab.num_chan = B.num_good[y];
  Ab.partials = A.partials + y * B.partials_stride;
  eval_body<MODE>(Ab, blockIdx.x, gridDim.x);
}

__global__ void __launch_bounds__(128) k_tables_batched(const PlanConst pc, const double* __restrict__ params,
                                                        CompTab* __restrict__ g_comp, SubTab* __restrict__ g_sub,
                                                        ChanTab* __restrict__ g_chan, const int* __restrict__ done) {
  const int y = blockIdx.y, nc = pc.num_comp;
  if (done[y]) return;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= pc.num_freq * nc) return;
  const int i = e / nc, c = e % nc;
  const double* pr = params + (size_t)y * FB_NUM_PARAMETERS * nc;
  CompTab ct;
  make_comp_tab(pc, pr, c, ct);
  if (i == 0) g_comp[(size_t)y * nc + c] = ct;
  const size_t ce = (size_t)y * pc.num_freq * nc + e;
  make_channel_tables(pc, pr, ct, i, c, g_sub + ce * pc.kf, g_chan[ce]);
}

// ---- split evaluation ---------------------------------------------------------------------
// The fused kernel is latency-bound at 16 warps/SM (122 registers, 52 KB shared memory per
// 128-thread block). Splitting it at the natural seam -- the (spectrum, timediff) pair of every
// (sample, component) -- gives two kernels that each need far fewer registers and far less
// shared memory, so twice as many warps hide the FP64 latency; the price is one round trip of
// 16 B per (sample, component) through HBM, which this path has to spare (the fused kernel
// uses 1 % of the DRAM bandwidth).
//   stage 1  k_vals   model side: regime split + general-case queue -> values in global memory
//   stage 2  k_gram2  derivative side: [J r] rows -> [J r]^T [J r]
__global__ void __launch_bounds__(kBlock, 6) k_vals(const EvalArgs A, double* __restrict__ gvals) {
  extern __shared__ double smem_raw[];
  const PlanConst& pc = A.pc;
  if (fit_stopped(pc)) return;
  const int nc = pc.num_comp, kf = pc.kf;
  SmemLayout L = carve(smem_raw, nc, kf, false, SM_QUEUE);
  load_comp_tables(A, L);
  const int tiles = (pc.num_time + kBlock - 1) / kBlock;
  const int vs = (A.vals_spc == 2 ? 2 : 3) * nc;  // slot planes per item in global memory
  const long long total = (long long)A.num_chan * tiles;
  const long long begin = total * blockIdx.x / gridDim.x, end = total * (blockIdx.x + 1) / gridDim.x;
  int cur = -1;
  for (long long item = begin; item < end; ++item) {
    const int chi = (int)(item / tiles), tile = (int)(item % tiles);
    const int i = A.chan_list ? A.chan_list[chi] : chi;
    if (chi != cur) {
      load_channel_tables(A, L, i);
      cur = chi;
    }
    eval_tile(A, L, i, tile, gvals + (size_t)item * kBlock * vs, 1);
  }
}

__global__ void __launch_bounds__(kBlock, 6) k_gram2(const EvalArgs A, const double* __restrict__ gvals) {
  extern __shared__ double smem_raw[];
  const PlanConst& pc = A.pc;
  if (fit_stopped(pc)) return;
  const int nc = pc.num_comp, kf = pc.kf, tid = threadIdx.x;
  SmemLayout L = carve(smem_raw, nc, kf, false, SM_PARK | SM_X);
  load_comp_tables(A, L);
  const int ncols = A.n_free + 1;
  const int stride = gram_stride(ncols), side = gram_side(ncols), ntiles = gram_ntiles(ncols);
  const int nslices = kBlock / ntiles;
  const int g_tile = tid % ntiles, g_slice = tid / ntiles;
  const bool g_active = g_slice < nslices;
  int g_ti = 0, g_tj = 0;
  tile_coords(g_tile, side, g_ti, g_tj);
#pragma unroll
  for (int k = 0; k < kTile * kTile; ++k) L.park[k * kBlock + tid] = 0.0;
  const int tiles = (pc.num_time + kBlock - 1) / kBlock;
  const int vs = (A.vals_spc == 2 ? 2 : 3) * nc;
  double* row = L.X + (size_t)tid * stride;
  const long long total = (long long)A.num_chan * tiles;
  const long long begin = total * blockIdx.x / gridDim.x, end = total * (blockIdx.x + 1) / gridDim.x;
  int cur = -1;
  for (long long item = begin; item < end; ++item) {
    const int chi = (int)(item / tiles), tile = (int)(item % tiles);
    const int i = A.chan_list ? A.chan_list[chi] : chi;
    if (chi != cur) {
      load_channel_tables(A, L, i);
      cur = chi;
    }
    const int j = tile * kBlock + tid;
    process_sample<MODE_GRAM>(A, L, i, j, j < pc.num_time, row, stride, gvals + (size_t)item * kBlock * vs + tid, kBlock);
    __syncthreads();
    gram_accumulate_parked(L, stride, ntiles, nslices, g_ti, g_tj, g_slice, g_active);
    __syncthreads();
  }
  double acc[kTile * kTile];
#pragma unroll
  for (int k = 0; k < kTile * kTile; ++k) acc[k] = L.park[k * kBlock + tid];
  gram_block_reduce(L.X, ntiles, nslices, g_tile, g_slice, g_active, acc,
                    A.partials + (size_t)blockIdx.x * ntiles * kTile * kTile);
}

// Stage 2 with FP64 tensor cores. ncu on k_gram2 showed the register-tile Gram limited by shared
// memory (short-scoreboard + MIO-throttle stalls, 8 LDS per 16 FMA, 24 % bank conflicts), not by
// the FP64 pipe. mma.sync.m8n8k4.f64 needs ONE shared-memory load per lane per 8-column block and
// 4-sample step -- the A fragment of block m and the B fragment of block n are the same element
// X[sample l%4][8 blk + l/4] -- i.e. 60x fewer loads. Each warp stages its own 32 rows
// (column-major, stride 36 = 4 mod 16: conflict-free for the lane-private writes AND the
// fragment reads), so a tile needs no block barrier at all; the C fragments of the NB(NB+1)/2
// upper block pairs stay in registers for the whole kernel.
constexpr int kMmaRowStride = 36;

__device__ __forceinline__ void dmma8x8x4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__host__ __device__ inline size_t gram_mma_smem_bytes(int nc, int kf, int ncols, int spc = 2) {
  const int nb = (ncols + 7) / 8;
  return smem_bytes(nc, kf, 1, false, false, 0) + sizeof(double) * (size_t)(kBlock / 32) * nb * 8 * kMmaRowStride +
         sizeof(double) * 2 * (size_t)(spc * nc + 1) * kBlock;  // double-buffered value + data prefetch
}

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gmem_src) {
  const unsigned saddr = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(saddr), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int NB>
__global__ void __launch_bounds__(kBlock, 6) k_gram2_mma(const EvalArgs A, const double* __restrict__ gvals) {
  extern __shared__ double smem_raw[];
  const PlanConst& pc = A.pc;
  if (fit_stopped(pc)) return;
  const int nc = pc.num_comp, kf = pc.kf, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  SmemLayout L = carve(smem_raw, nc, kf, false, 0);
  double* xt = smem_raw + smem_bytes(nc, kf, 1, false, false, 0) / sizeof(double) +
               (size_t)warp * NB * 8 * kMmaRowStride;  // this warp's [NB*8 columns][36] stage
  load_comp_tables(A, L);
  const int ncols = A.n_free + 1;
  constexpr int kPairs = NB * (NB + 1) / 2;
  double c0[kPairs], c1[kPairs];
#pragma unroll
  for (int k = 0; k < kPairs; ++k) c0[k] = c1[k] = 0.0;
  // padding columns are never written by process_sample: clear them once
  for (int e = lane; e < NB * 8 * kMmaRowStride; e += 32) xt[e] = 0.0;
  __syncwarp();
  const int tiles = (pc.num_time + kBlock - 1) / kBlock;
  const int vs = (A.vals_spc == 2 ? 2 : 3) * nc;
  const long long total = (long long)A.num_chan * tiles;
  const long long begin = total * blockIdx.x / gridDim.x, end = total * (blockIdx.x + 1) / gridDim.x;
  int cur = -1;
  const int fr = lane & 3, fc = lane >> 2;  // fragment element: sample fr of the step, column fc of the block
  // The values written by k_vals and the spectrum sample of the NEXT tile are fetched with
  // cp.async while this tile's rows are formed and multiplied (ncu: 35 % of the stall samples
  // of the unprefetched kernel sat on the first use of those global loads). Every lane copies
  // and later reads only its own elements, so cp.async.wait_group is the only synchronisation.
  double* pref = smem_raw + smem_bytes(nc, kf, 1, false, false, 0) / sizeof(double) +
                 (size_t)(kBlock / 32) * NB * 8 * kMmaRowStride;  // [2][vs + 1][kBlock]
  auto prefetch = [&](long long item, int buf) {
    const int chi = (int)(item / tiles), tile = (int)(item % tiles);
    const int j = tile * kBlock + tid;
    if (j < pc.num_time) {
      const int i = A.chan_list ? A.chan_list[chi] : chi;
      double* dst = pref + (size_t)buf * (vs + 1) * kBlock + tid;
      const double* src = gvals + (size_t)item * kBlock * vs + tid;
      for (int k = 0; k < vs; ++k) cp_async8(dst + (size_t)k * kBlock, src + (size_t)k * kBlock);
      cp_async8(dst + (size_t)vs * kBlock, A.data + (size_t)i * pc.num_time + j);
    }
    cp_async_commit();
  };
  if (begin < end) prefetch(begin, 0);
  for (long long item = begin; item < end; ++item) {
    const int chi = (int)(item / tiles), tile = (int)(item % tiles);
    const int i = A.chan_list ? A.chan_list[chi] : chi;
    if (chi != cur) {
      load_channel_tables(A, L, i);
      cur = chi;
    }
    const int buf = (int)((item - begin) & 1);
    if (item + 1 < end) {
      prefetch(item + 1, buf ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    const int j = tile * kBlock + tid;
    const double* staged = pref + (size_t)buf * (vs + 1) * kBlock + tid;
    // [J r] row of this lane's sample -> column `lane` of the warp's stage
    process_sample<MODE_GRAM>(A, L, i, j, j < pc.num_time, xt + lane, ncols, staged, kBlock, kMmaRowStride,
                              staged + (size_t)vs * kBlock);
    __syncwarp();
#pragma unroll
    for (int step = 0; step < 8; ++step) {
      double f[NB];
#pragma unroll
      for (int b = 0; b < NB; ++b) f[b] = xt[(8 * b + fc) * kMmaRowStride + 4 * step + fr];
      int k = 0;
#pragma unroll
      for (int bi = 0; bi < NB; ++bi)
#pragma unroll
        for (int bj = bi; bj < NB; ++bj) {
          dmma8x8x4(c0[k], c1[k], f[bi], f[bj]);
          ++k;
        }
    }
    __syncwarp();
  }
  // C fragment of pair k: rows 8 bi + lane/4, columns 8 bj + 2 (lane%4) + {0, 1}
  double* out = A.partials + ((size_t)blockIdx.x * (kBlock / 32) + warp) * kPairs * 64;
#pragma unroll
  for (int k = 0; k < kPairs; ++k) {
    out[k * 64 + (lane >> 2) * 8 + (lane & 3) * 2] = c0[k];
    out[k * 64 + (lane >> 2) * 8 + (lane & 3) * 2 + 1] = c1[k];
  }
}

// Sums the per-warp C fragments and writes the symmetric (ncols x ncols) matrix: one warp per
// matrix entry, lanes stride over the contributing warps, fixed-shape shuffle tree (deterministic).
__global__ void __launch_bounds__(256) k_gram_finalize_mma(const double* __restrict__ partials, int nwarps,
                                                          int ncols, int nb, double* __restrict__ gram,
                                                          const int* __restrict__ stop) {
  if (stop != nullptr && *stop != 0) return;
  const int pairs = nb * (nb + 1) / 2;
  const int e = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (e >= pairs * 64) return;
  double s = 0.0;
  for (int w = lane; w < nwarps; w += 32) s += partials[(size_t)w * pairs * 64 + e];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) {
    int k = e / 64, bi = 0;
    while (k >= nb - bi) {
      k -= nb - bi;
      ++bi;
    }
    const int bj = bi + k;
    const int gr = 8 * bi + (e % 64) / 8, gc = 8 * bj + (e % 8);
    if (gr < ncols && gc < ncols && (bi != bj || gr <= gc)) {
      gram[gr * ncols + gc] = s;
      gram[gc * ncols + gr] = s;
    }
  }
}

// Second stage: sums the per-block partial tiles (one warp per entry, lanes stride over the
// blocks, fixed-shape shuffle tree) and writes the symmetric (ncols x ncols) matrix.
__global__ void __launch_bounds__(256) k_gram_finalize(const double* __restrict__ partials, int nblocks,
                                                      int ncols, double* __restrict__ gram,
                                                      const int* __restrict__ stop) {
  if (stop != nullptr && *stop != 0) return;
  const int side = gram_side(ncols), ntiles = gram_ntiles(ncols);
  constexpr int kk = kTile * kTile;
  const int e = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (e >= ntiles * kk) return;
  double s = 0.0;
  for (int b = lane; b < nblocks; b += 32) s += partials[(size_t)b * ntiles * kk + e];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) {
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

// chi^2 only (grid sweeps, BASELINE.json config 4): the model side of the split evaluation with
// the residual accumulated in place -- no derivative state, so it runs at k_vals' occupancy.
__global__ void __launch_bounds__(kBlock, 6) k_chi2(const EvalArgs A) {
  extern __shared__ double smem_raw[];
  const PlanConst& pc = A.pc;
  const int nc = pc.num_comp, kf = pc.kf, tid = threadIdx.x;
  SmemLayout L = carve(smem_raw, nc, kf, false, SM_VALS | SM_QUEUE);
  load_comp_tables(A, L);
  const int tiles = (pc.num_time + kBlock - 1) / kBlock;
  const int vs = val_stride(nc);
  const long long total = (long long)A.num_chan * tiles;
  const long long begin = total * blockIdx.x / gridDim.x, end = total * (blockIdx.x + 1) / gridDim.x;
  int cur = -1;
  double acc = 0.0;
  for (long long item = begin; item < end; ++item) {
    const int chi = (int)(item / tiles), tile = (int)(item % tiles);
    const int i = A.chan_list ? A.chan_list[chi] : chi;
    if (chi != cur) {
      load_channel_tables(A, L, i);
      cur = chi;
    }
    eval_tile(A, L, i, tile);
    const int j = tile * kBlock + tid;
    if (j < pc.num_time) {
      const double* v = L.vals + (size_t)tid * vs;
      double msum = 0.0;
      for (int c = 0; c < nc; ++c) msum += v[3 * c];
      const double r = (A.data[(size_t)i * pc.num_time + j] - msum) * A.weights[i];
      acc = fma(r, r, acc);
    }
  }
  // deterministic block tree-sum (the value rows are free again)
  __syncthreads();
  double* red = L.vals;
  red[tid] = acc;
  __syncthreads();
  for (int off = kBlock / 2; off > 0; off >>= 1) {
    if (tid < off) red[tid] += red[tid + off];
    __syncthreads();
  }
  if (tid == 0) A.chi2_partials[blockIdx.x] = red[0];
}

// Sums per-block chi^2 partials into out[0]: strided per-thread sums, then a fixed-shape tree
// (deterministic). Launch with one block of 256 threads.
__global__ void __launch_bounds__(256) k_chi2_finalize(const double* __restrict__ partials, int nblocks,
                                                      double* __restrict__ out) {
  __shared__ double red[256];
  const int tid = threadIdx.x;
  double s = 0.0;
  for (int b = tid; b < nblocks; b += 256) s += partials[b];
  red[tid] = s;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (tid < off) red[tid] += red[tid + off];
    __syncthreads();
  }
  if (tid == 0) out[0] = red[0];
}

// Overrides the global dm / scattering_timescale rows of a parameter block (chi^2 grid sweep).
__global__ void k_set_grid_point(double* params, int nc, double dm, double sc_time) {
  const int c = threadIdx.x;
  if (c < nc) {
    params[FB_DM * nc + c] = dm;
    params[FB_SCATTERING_TIMESCALE * nc + c] = sc_time;
  }
}

// LSFitter._set_weights (analysis/fitter.py:506-542): one block per channel.
__global__ void __launch_bounds__(kBlock) k_weights(const double* __restrict__ data, int num_freq,
                                                   int num_time, const uint8_t* __restrict__ good,
                                                   int weighted, int lo, int hi,
                                                   double* __restrict__ weights,
                                                   double* __restrict__ chan_sumsq) {
  __shared__ double red[2][kBlock];
  const int i = blockIdx.x, tid = threadIdx.x;
  if (!good[i]) {  // weight 0 whatever the row holds (it may not even have been uploaded: fb_upload_rows)
    if (tid == 0) {
      weights[i] = 0.0;
      chan_sumsq[i] = 0.0;
    }
    return;
  }
  const double* rowp = data + (size_t)i * num_time;
  double s_range = 0.0, s_all = 0.0;
  for (int j = tid; j < num_time; j += kBlock) {
    const double v = rowp[j];
    const double v2 = v * v;
    s_all += v2;
    if (j >= lo && j < hi) s_range += v2;
  }
  red[0][tid] = s_range;
  red[1][tid] = s_all;
  __syncthreads();
  for (int off = kBlock / 2; off > 0; off >>= 1) {
    if (tid < off) {
      red[0][tid] += red[0][tid + off];
      red[1][tid] += red[1][tid + off];
    }
    __syncthreads();
  }
  if (tid == 0) {
    double w = 0.0;
    if (good[i]) {
      const int cnt = hi - lo;
      w = weighted ? 1.0 / sqrt(red[0][0] / (double)cnt) : 1.0;
    }
    weights[i] = w;
    chan_sumsq[i] = red[1][0] * w * w;
  }
}

// Stand-alone temporal profile on an arbitrary (rows, cols) array of times
// (SpectrumModeler.compute_profile, analysis/model.py:281-352; routines/profile.py:15-114).
__global__ void k_profile(const double* __restrict__ times, const double* __restrict__ freqs, int rows,
                          int cols, double toa, double sc_time_ref, double sc_index, double width,
                          double ref_freq, int scattered, int clamp, int folded, int normalize,
                          double* __restrict__ out) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)rows * cols) return;
  const int r = (int)(e / cols), m = (int)(e % cols);
  const double* row = times + (size_t)r * cols;
  const double sc_time = sc_time_ref * pow(freqs[r] / ref_freq, sc_index);
  const double ratio = width / sc_time;
  const double amp = normalize ? kSqrtHalfPi * ratio : sc_time_ref / sc_time;
  const double neg5w = -5 * width;
  auto shape = [&](double t) {
    if (scattered) {
      if (clamp && t < neg5w) t = neg5w;  // analysis/model.py:340
      return amp * emg_shape((t - toa) / width, ratio);
    }
    const double z = (t - toa) / width;
    return exp(-0.5 * z * z);
  };
  double p = shape(row[m]);
  if (folded) {  // analysis/model.py:327-332,349-350
    const double res = row[1] - row[0];
    const double start = row[cols - 1] + res, stop = row[cols - 1] + res * (double)cols;
    const double step = cols > 1 ? (stop - start) / (double)(cols - 1) : 0.0;
    p = 0.5 * (p + shape(linspace_at(start, step, stop, m, cols)));
  }
  out[e] = p;
}

// Dependent-free FP64 FMA chains: the measured roofline denominator for the model kernels.
__global__ void k_dfma_peak(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
         a6 = a0 + 6, a7 = a0 + 7;
  const double m = 0.999999, b = 1e-7;
  for (int it = 0; it < iters; ++it) {
    a0 = fma(a0, m, b);
    a1 = fma(a1, m, b);
    a2 = fma(a2, m, b);
    a3 = fma(a3, m, b);
    a4 = fma(a4, m, b);
    a5 = fma(a5, m, b);
    a6 = fma(a6, m, b);
    a7 = fma(a7, m, b);
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

}  // namespace fb
