// C ABI of fitburst_b200 (see include/fitburst_b200.h). Host-side plan management and the
// launch logic for the kernels in fb_kernels.cuh / fb_hessian.cuh / fb_batched.cuh.
#include <cuda_runtime.h>
#include <sched.h>
#include <stdarg.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <new>
#include <vector>

#include "fb_kernels.cuh"
#include "fb_fast.cuh"
#include "fb_fast4.cuh"
#include "fb_hessian.cuh"
#include "fb_batched.cuh"
#include "fb_preprocess.cuh"
#define FB_HAVE_HESSIAN 1
#include "fb_trf.h"

using namespace fb;

static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CUDA_TRY(expr)                                                                   \
  do {                                                                                   \
    cudaError_t e_ = (expr);                                                             \
    if (e_ != cudaSuccess)                                                               \
      return fail(FB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_),   \
                  __FILE__, __LINE__);                                                   \
  } while (0)

struct fb_plan_s {
  fb_plan_desc desc;
  PlanConst pc;
  double* d_freqs = nullptr;
  double* d_subfreqs = nullptr;
  double* d_params = nullptr;
  double* d_weights = nullptr;
  double* d_chan_sumsq = nullptr;
  double* d_partials = nullptr;
  double* d_scint_amps = nullptr;      // (num_freq, num_comp) scintillation amplitudes (k_scint_solve)
  double* d_scint_partials = nullptr;  // chunk partials of the per-channel systems
  size_t scint_partials_cap = 0;
  size_t partials_cap = 0;
  double* d_gram = nullptr;  // (FB_MAX_FREE+1)^2
  double* h_gram = nullptr;  // pinned
  double* d_vals = nullptr;  // split evaluation: (spectrum, timediff, timeprof) per (sample, component)
  size_t vals_cap = 0;
  int* d_bounds = nullptr;   // k_rise3: regime boundaries (jp, jt) per (good channel, component)
  size_t bounds_cap = 0;
  CompTab* d_comp = nullptr;
  SubTab* d_sub = nullptr;
  ChanTab* d_chan = nullptr;
  bool tables_dirty = true;
  bool tables_full = false;  // the current tables cover every channel (not only the good ones)
  bool last_eval_split = false;
  int vals_layout = 2;       // 2: k_vals planes (spectrum, timediff) per tile; 3: k_vals3 rows (spectrum only)
  double* d_racc = nullptr;  // k_gram3 accumulators summed over blocks
  double* d_slices = nullptr;  // k_gram3_finish: per-slice sums of every accumulator entry
  double* d_far = nullptr;   // k_gram3 far-tile matrix (upper triangle) summed over blocks
  unsigned* d_counter = nullptr;  // k_gram3_finish: blocks done (self-resetting)
  int occ_cache[64];         // occupancy of the fast kernels by instantiation (0 = not queried yet)
  double* d_hess_partials = nullptr;
  size_t hess_partials_cap = 0;
  HessPair* d_pairs = nullptr;
  size_t pairs_cap = 0;
  int* d_comp_list = nullptr;
  size_t comp_list_cap = 0;
  HessDesc* d_desc = nullptr;
  size_t desc_cap = 0;
  int* d_chan_list = nullptr;
  int num_good = 0;
  bool have_weights = false;
  int* d_err = nullptr;
  uint8_t* d_good = nullptr;
  double h_params[FB_NUM_PARAMETERS * FB_MAX_COMPONENTS];
  int free_mask[9];
  int col_of[9];
  int n_free = 0;
  int sm_count = 148;
  int64_t launches = 0;
  std::vector<double> h_weights;
  // last [J r]^T [J r] returned to the host, keyed by what it was computed from: lets
  // fb_hessian reuse the J^T J of the final trust-region evaluation instead of recomputing it.
  bool gram_cache_valid = false;
  std::vector<double> gram_cache, gram_cache_params;
  const double* gram_cache_data = nullptr;
  bool vals_match_cache = false;  // d_vals holds k_vals output of the cached evaluation
  // device-driven fit loop (k_trf_single): solver state, flags and results stay on the GPU
  TrfState* d_trf = nullptr;
  int* d_fit_flags = nullptr;        // FitFlag
  double* d_gram_accepted = nullptr; // (FB_MAX_FREE+1)^2
  fb_fit_result* d_fit_result = nullptr;
  int* h_fit_flags = nullptr;        // pinned + device-mapped, 8 slots of kFitFlagCount (slot 0 doubles as the final copy)
  int* h_fit_flags_dev = nullptr;    // device address of h_fit_flags
  fb_fit_result* h_fit_result = nullptr;  // pinned
  double* h_fit_grams = nullptr;     // pinned: last evaluated + accepted matrix
  double* h_fit_params = nullptr;    // pinned parameter block
  cudaEvent_t fit_events[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  SingleTrfArgs* fit_session = nullptr;  // fb_fit_device_* between _begin and _end
  // pair descriptors of the exact Hessian for the current free-parameter layout (fb_hessian_impl)
  bool hess_layout_valid = false;  // the device copies below belong to the current free-parameter layout
  std::vector<HessPair> hess_pairs;
  std::vector<int> hess_pair_a, hess_pair_b, hess_comp_list;
  int hess_comp_off[FB_MAX_COMPONENTS + 1];
  bool hess_desc_valid = false;
  std::vector<HessDesc> hess_desc;
  unsigned hess_need[FB_MAX_COMPONENTS];
  int hess_need_sum_dmidx = 0;
};

extern "C" const char* fb_last_error(void) { return g_err; }
extern "C" int fb_version(void) { return 100; }

static void layout_free(fb_plan_s* p) {
  int col = 0;
  for (int k = 0; k < 9; ++k) {
    if (p->free_mask[k]) {
      p->col_of[k] = col;
      col += is_fit_global(k) ? 1 : p->desc.num_components;
    } else {
      p->col_of[k] = -1;
    }
  }
  p->n_free = col;
}

extern "C" int fb_plan_create(const fb_plan_desc* desc, const double* freqs_host,
                              const double* times_host, fb_plan_t* plan_out) {
  if (!desc || !freqs_host || !times_host || !plan_out) return fail(FB_ERR_INVALID, "null argument");
  if (desc->num_freq < 2 || desc->num_time < 2)
    return fail(FB_ERR_INVALID, "need at least two channels and two samples (res_freq/res_time come "
                                "from the first two labels)");
  if (desc->num_components < 1 || desc->num_components > FB_MAX_COMPONENTS)
    return fail(FB_ERR_INVALID, "num_components must be in 1..%d", FB_MAX_COMPONENTS);
  if (desc->factor_freq_upsample < 1 || desc->factor_time_upsample < 1)
    return fail(FB_ERR_INVALID, "up-sampling factors must be >= 1");
  if ((long long)desc->num_time * desc->factor_time_upsample > 2000000000LL)
    return fail(FB_ERR_INVALID, "num_time * factor_time_upsample overflows int32");
  int dev_count = 0;
  if (cudaGetDeviceCount(&dev_count) != cudaSuccess || dev_count == 0)
    return fail(FB_ERR_CUDA, "no CUDA device: fitburst_b200 has no CPU fallback");
  fb_plan_s* p = new (std::nothrow) fb_plan_s();
  if (!p) return fail(FB_ERR_NOMEM, "out of host memory");
  p->desc = *desc;
  const int nf = desc->num_freq, nt = desc->num_time, kf = desc->factor_freq_upsample;
  PlanConst& pc = p->pc;
  pc.num_freq = nf;
  pc.num_time = nt;
  pc.num_comp = desc->num_components;
  pc.kf = kf;
  pc.kt = desc->factor_time_upsample;
  pc.is_folded = desc->is_folded;
  pc.scintillation = desc->scintillation;
  pc.normalize_pbf = desc->normalize_pbf;
  pc.ntk = nt * pc.kt;
  pc.dm_incoherent = desc->dm_incoherent;
  pc.dm_const = desc->dispersion_constant;
  pc.res_freq = fabs(freqs_host[1] - freqs_host[0]);  // analysis/model.py:87
  pc.res_time = fabs(times_host[1] - times_host[0]);  // analysis/model.py:88
  pc.time_first = times_host[0];
  pc.time_last = times_host[nt - 1];
  // sub-frequency labels, routines/manipulate.py:99-105 with np.linspace semantics:
  // arange(K) * step + lo (two roundings), last element forced to hi when K > 1.
  std::vector<double> sub((size_t)nf * kf);
  const double diff_new = pc.res_freq / (double)kf;
  for (int i = 0; i < nf; ++i) {
    volatile double lo = freqs_host[i] - (pc.res_freq / 2);
    lo = lo + (diff_new / 2);
    volatile double hi = freqs_host[i] + (pc.res_freq / 2);
    hi = hi - (diff_new / 2);
    const double step = kf > 1 ? (hi - lo) / (double)(kf - 1) : 0.0;
    for (int a = 0; a < kf; ++a) {
      volatile double v = (double)a * step;
      v = v + lo;
      sub[(size_t)i * kf + a] = (a == kf - 1 && kf > 1) ? (double)hi : (double)v;
    }
  }
  cudaDeviceProp prop;
  int dev = 0;
  cudaGetDevice(&dev);
  if (cudaGetDeviceProperties(&prop, dev) == cudaSuccess) p->sm_count = prop.multiProcessorCount;
  const size_t gsz = sizeof(double) * (FB_MAX_FREE + 1) * (FB_MAX_FREE + 1);
  cudaError_t e = cudaSuccess;
  auto chk = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
  chk(cudaMalloc(&p->d_freqs, sizeof(double) * nf));
  chk(cudaMalloc(&p->d_subfreqs, sizeof(double) * nf * kf));
  chk(cudaMalloc(&p->d_params, sizeof(double) * FB_NUM_PARAMETERS * FB_MAX_COMPONENTS));
  chk(cudaMalloc(&p->d_weights, sizeof(double) * nf));
  chk(cudaMalloc(&p->d_chan_sumsq, sizeof(double) * nf));
  chk(cudaMalloc(&p->d_chan_list, sizeof(int) * nf));
  chk(cudaMalloc(&p->d_good, nf));
  chk(cudaMalloc(&p->d_err, sizeof(int)));
  chk(cudaMalloc(&p->d_gram, gsz));
  chk(cudaMalloc(&p->d_racc, sizeof(double) * 64 * 16));
  chk(cudaMalloc(&p->d_slices, sizeof(double) * kFinishSlices * (64 * 16 + (FB_MAX_FREE + 1) * (FB_MAX_FREE + 2) / 2)));
  chk(cudaMalloc(&p->d_far, sizeof(double) * (FB_MAX_FREE + 1) * (FB_MAX_FREE + 2) / 2));
  chk(cudaMalloc(&p->d_comp, sizeof(CompTab) * FB_MAX_COMPONENTS));
  chk(cudaMalloc(&p->d_sub, sizeof(SubTab) * (size_t)nf * desc->num_components * kf));
  chk(cudaMalloc(&p->d_chan, sizeof(ChanTab) * (size_t)nf * desc->num_components));
  chk(cudaMallocHost(&p->h_gram, gsz));
  if (e == cudaSuccess) chk(cudaMemcpy(p->d_freqs, freqs_host, sizeof(double) * nf, cudaMemcpyHostToDevice));
  if (e == cudaSuccess) chk(cudaMemcpy(p->d_subfreqs, sub.data(), sizeof(double) * nf * kf, cudaMemcpyHostToDevice));
  chk(cudaMalloc(&p->d_counter, sizeof(unsigned)));
  if (e == cudaSuccess) chk(cudaMemset(p->d_err, 0, sizeof(int)));
  if (e == cudaSuccess) chk(cudaMemset(p->d_counter, 0, sizeof(unsigned)));
  if (e != cudaSuccess) {
    fb_plan_destroy(p);
    return fail(FB_ERR_CUDA, "plan allocation failed: %s", cudaGetErrorString(e));
  }
  pc.freqs = p->d_freqs;
  pc.subfreqs = p->d_subfreqs;
  pc.stop = nullptr;
  memset(p->h_params, 0, sizeof(p->h_params));
  memset(p->occ_cache, 0, sizeof(p->occ_cache));
  for (int k = 0; k < 9; ++k) p->free_mask[k] = 1;
  layout_free(p);
  *plan_out = p;
  return FB_OK;
}

extern "C" int fb_plan_destroy(fb_plan_t p) {
  if (!p) return FB_OK;
  cudaFree(p->d_freqs);
  cudaFree(p->d_subfreqs);
  cudaFree(p->d_params);
  cudaFree(p->d_weights);
  cudaFree(p->d_chan_sumsq);
  cudaFree(p->d_chan_list);
  cudaFree(p->d_good);
  cudaFree(p->d_err);
  cudaFree(p->d_gram);
  cudaFree(p->d_partials);
  cudaFree(p->d_scint_amps);
  cudaFree(p->d_scint_partials);
  cudaFree(p->d_hess_partials);
  cudaFree(p->d_pairs);
  cudaFree(p->d_comp_list);
  cudaFree(p->d_vals);
  cudaFree(p->d_bounds);
  cudaFree(p->d_racc);
  cudaFree(p->d_slices);
  cudaFree(p->d_far);
  cudaFree(p->d_counter);
  cudaFree(p->d_desc);
  cudaFree(p->d_comp);
  cudaFree(p->d_sub);
  cudaFree(p->d_chan);
  if (p->h_gram) cudaFreeHost(p->h_gram);
  cudaFree(p->d_trf);
  delete p->fit_session;
  cudaFree(p->d_fit_flags);
  cudaFree(p->d_gram_accepted);
  cudaFree(p->d_fit_result);
  if (p->h_fit_flags) cudaFreeHost(p->h_fit_flags);
  if (p->h_fit_result) cudaFreeHost(p->h_fit_result);
  if (p->h_fit_grams) cudaFreeHost(p->h_fit_grams);
  if (p->h_fit_params) cudaFreeHost(p->h_fit_params);
  for (int k = 0; k < 8; ++k)
    if (p->fit_events[k]) cudaEventDestroy(p->fit_events[k]);
  delete p;
  return FB_OK;
}

extern "C" int fb_set_parameters(fb_plan_t p, const double* params_host, void* stream) {
  if (!p || !params_host) return fail(FB_ERR_INVALID, "null argument");
  const size_t n = (size_t)FB_NUM_PARAMETERS * p->desc.num_components;
  memcpy(p->h_params, params_host, sizeof(double) * n);
  CUDA_TRY(cudaMemcpyAsync(p->d_params, p->h_params, sizeof(double) * n, cudaMemcpyHostToDevice,
                           (cudaStream_t)stream));
  p->tables_dirty = true;
  return FB_OK;
}

static bool env_off(const char* name) {
  const char* v = getenv(name);
  return v && !strcmp(v, "0");
}

// Rebuilds the per-channel tables when the parameters changed since the last build. The fit only
// ever reads the tables of the good channels (good_only); everything else needs all of them.
template <int KF>
static void launch_tables_group(fb_plan_s* p, const int* list, int count, cudaStream_t st) {
  const long long threads = (long long)count * p->desc.num_components * KF;
  k_tables_group<KF><<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(p->pc, p->d_params, p->d_comp, p->d_sub,
                                                                        p->d_chan, list, count);
}

static int ensure_tables(fb_plan_s* p, cudaStream_t st, bool good_only = false) {
  const bool list_ok = good_only && p->have_weights && p->num_good > 0;
  if (!p->tables_dirty && (p->tables_full || list_ok)) return FB_OK;
  const int* list = list_ok ? p->d_chan_list : nullptr;
  const int count = list_ok ? p->num_good : p->desc.num_freq;
  switch (p->pc.kf) {
    case 1: launch_tables_group<1>(p, list, count, st); break;
    case 2: launch_tables_group<2>(p, list, count, st); break;
    case 4: launch_tables_group<4>(p, list, count, st); break;
    case 8: launch_tables_group<8>(p, list, count, st); break;
    case 16: launch_tables_group<16>(p, list, count, st); break;
    case 32: launch_tables_group<32>(p, list, count, st); break;
    default: {
      const int total = p->desc.num_freq * p->desc.num_components;
      k_tables<<<(total + 127) / 128, 128, 0, st>>>(p->pc, p->d_params, p->d_comp, p->d_sub, p->d_chan);
      list = nullptr;
    }
  }
  CUDA_TRY(cudaGetLastError());
  p->launches += 1;
  p->tables_dirty = false;
  p->tables_full = list == nullptr;
  return FB_OK;
}

extern "C" int fb_get_parameters(fb_plan_t p, double* params_host, void* stream) {
  if (!p || !params_host) return fail(FB_ERR_INVALID, "null argument");
  const size_t n = (size_t)FB_NUM_PARAMETERS * p->desc.num_components;
  CUDA_TRY(cudaMemcpyAsync(params_host, p->d_params, sizeof(double) * n, cudaMemcpyDeviceToHost,
                           (cudaStream_t)stream));
  CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
  memcpy(p->h_params, params_host, sizeof(double) * n);
  return FB_OK;
}

static EvalArgs base_args(fb_plan_s* p) {
  EvalArgs A;
  memset(&A, 0, sizeof(A));
  A.pc = p->pc;
  A.params = p->d_params;
  A.chan_list = nullptr;
  A.num_chan = p->desc.num_freq;
  A.n_free = p->n_free;
  for (int k = 0; k < 9; ++k) A.col_of[k] = p->col_of[k];
  A.err_flag = p->d_err;
  A.g_comp = p->d_comp;
  A.g_sub = p->d_sub;
  A.g_chan = p->d_chan;
  return A;
}

static int ensure_capacity(double** buf, size_t* cap, size_t need);

// Scintillation amplitudes of the channels in A (routines/ism.py:11-48) ahead of an evaluation:
// (channel, chunk) partial systems on the whole GPU, then one small solve per channel. The tables
// must be current. On return A.scint_amps points at the plan's (num_freq, num_comp) block.
// FITBURST_B200_SCINT_PRESOLVE=0: one block per channel solves in place (round-1 behaviour).
static int presolve_scint_amps(fb_plan_s* p, EvalArgs& A, cudaStream_t st) {
  if (!p->pc.scintillation || A.scint_amps != nullptr || A.num_chan < 1 || !A.data) return FB_OK;
  if (env_off("FITBURST_B200_SCINT_PRESOLVE")) return FB_OK;
  const int nc = p->pc.num_comp, kf = p->pc.kf;
  if (!p->d_scint_amps) CUDA_TRY(cudaMalloc(&p->d_scint_amps, sizeof(double) * (size_t)p->desc.num_freq * nc));
  const size_t smem = smem_bytes(nc, kf, 1, true);
  if (smem > 227 * 1024) return fail(FB_ERR_INVALID, "shared-memory request %zu B exceeds 227 KB", smem);
  if (smem > 48 * 1024)
    CUDA_TRY(cudaFuncSetAttribute(k_scint_partials, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 1;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_scint_partials, kBlock, smem));
  if (occ < 1) occ = 1;
  const int tiles = (p->pc.num_time + kBlock - 1) / kBlock;
  const long long slots = (long long)p->sm_count * occ * 4;  // a few items per resident block: balance
  const int nchunks = (int)std::min<long long>(tiles, std::max<long long>(1, (slots + A.num_chan - 1) / A.num_chan));
  const size_t nent = (size_t)gram_ntiles(nc + 1) * kTile * kTile;
  const int rc = ensure_capacity(&p->d_scint_partials, &p->scint_partials_cap, (size_t)A.num_chan * nchunks * nent);
  if (rc != FB_OK) return rc;
  const long long items = (long long)A.num_chan * nchunks;
  const int grid = (int)std::min<long long>(items, (long long)p->sm_count * occ);
  k_scint_partials<<<grid, kBlock, smem, st>>>(A, nchunks, p->d_scint_partials);
  CUDA_TRY(cudaGetLastError());
  k_scint_solve<<<A.num_chan, 64, 0, st>>>(A, nchunks, p->d_scint_partials, p->d_scint_amps);
  CUDA_TRY(cudaGetLastError());
  p->launches += 2;
  A.scint_amps = p->d_scint_amps;
  return FB_OK;
}

template <int MODE>
static int launch_eval(fb_plan_s* p, const EvalArgs& A_in, int ncols_main, cudaStream_t st, int* grid_out) {
  {
    const int rc_t = ensure_tables(p, st, A_in.chan_list != nullptr && A_in.chan_list == p->d_chan_list);
    if (rc_t != FB_OK) return rc_t;
  }
  EvalArgs A = A_in;
  {
    const int rc_s = presolve_scint_amps(p, A, st);
    if (rc_s != FB_OK) return rc_s;
  }
  const size_t smem = smem_bytes(p->pc.num_comp, p->pc.kf, ncols_main, p->pc.scintillation != 0);
  if (smem > 227 * 1024) return fail(FB_ERR_INVALID, "shared-memory request %zu B exceeds 227 KB", smem);
  // the opt-in is per device: set it on every launch that needs it (a thread may switch devices)
  if (smem > 48 * 1024)
    CUDA_TRY(cudaFuncSetAttribute(k_eval<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 1;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_eval<MODE>, kBlock, smem));
  if (occ < 1) occ = 1;
  const int tiles = (p->pc.num_time + kBlock - 1) / kBlock;
  long long items = (p->pc.scintillation && !A.scint_amps) ? (long long)A.num_chan : (long long)A.num_chan * tiles;
  if (items < 1) items = 1;
  long long grid = (long long)p->sm_count * occ;
  if (MODE != MODE_GRAM) grid *= 8;  // short blocks balance better when nothing is kept in registers
  if (grid > items) grid = items;
  if (grid_out) *grid_out = (int)grid;
  EvalArgs B = A;
  if (MODE == MODE_GRAM) {
    const size_t need = (size_t)grid * gram_ntiles(ncols_main) * kTile * kTile;
    if (need > p->partials_cap) {
      cudaFree(p->d_partials);
      p->d_partials = nullptr;
      p->partials_cap = 0;
      CUDA_TRY(cudaMalloc(&p->d_partials, sizeof(double) * need));
      p->partials_cap = need;
    }
    B.partials = p->d_partials;
  }
  k_eval<MODE><<<(int)grid, kBlock, smem, st>>>(B);
  CUDA_TRY(cudaGetLastError());
  p->launches += 1;
  return FB_OK;
}

static int check_err_flag(fb_plan_s* p, cudaStream_t st) {
  if (!p->pc.scintillation) return FB_OK;
  int flag = 0;
  CUDA_TRY(cudaMemcpyAsync(&flag, p->d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  if (flag) {
    CUDA_TRY(cudaMemsetAsync(p->d_err, 0, sizeof(int), st));
    return fail(FB_ERR_SINGULAR, "Singular matrix");  // numpy.linalg.LinAlgError text, routines/ism.py:46
  }
  return FB_OK;
}

static bool fast_path_ok(const fb_plan_s* p, bool for_gram);
static int launch_vals3(fb_plan_s* p, const EvalArgs& A, double* out, cudaStream_t st, int* grid_out);

extern "C" int fb_model_eval(fb_plan_t p, const double* data_dev, double* model_dev,
                             double* amplitude_dev, double* spectrum_dev, double* timediff_dev,
                             double* timeprof_dev, void* stream) {
  if (!p || !model_dev) return fail(FB_ERR_INVALID, "null argument");
  if (p->pc.scintillation && !data_dev)
    return fail(FB_ERR_INVALID, "ERROR: scintillation modelling is desired by data are missing!");
  if (!amplitude_dev && !spectrum_dev && !timediff_dev && !timeprof_dev && fast_path_ok(p, false)) {
    // model only, scattered branch: the warp-granular k_vals3 over every channel + component sum
    cudaStream_t st = (cudaStream_t)stream;
    const int nf = p->desc.num_freq, nc = p->desc.num_components, nt = p->desc.num_time;
    int rc = ensure_tables(p, st);
    if (rc == FB_OK) rc = ensure_capacity(&p->d_vals, &p->vals_cap, (size_t)nf * nc * nt);
    if (rc != FB_OK) return rc;
    EvalArgs A = base_args(p);
    p->vals_match_cache = false;
    p->last_eval_split = false;
    rc = launch_vals3(p, A, p->d_vals, st, nullptr);
    if (rc != FB_OK) return rc;
    const long long total = (long long)nf * nt;
    k_model_from_vals<<<(unsigned)std::min<long long>((total + 255) / 256, 148LL * 32), 256, 0, st>>>(p->d_vals, nc, nt,
                                                                                                      total, model_dev);
    CUDA_TRY(cudaGetLastError());
    p->launches += 1;
    return FB_OK;
  }
  EvalArgs A = base_args(p);
  A.data = data_dev;
  A.model = model_dev;
  A.c_amp = amplitude_dev;
  A.c_spec = spectrum_dev;
  A.c_td = timediff_dev;
  A.c_tp = timeprof_dev;
  A.n_free = 0;
  int rc = launch_eval<MODE_MODEL>(p, A, 1, (cudaStream_t)stream, nullptr);
  if (rc != FB_OK) return rc;
  return check_err_flag(p, (cudaStream_t)stream);
}

static int rebuild_channel_list(fb_plan_s* p, cudaStream_t st) {
  std::vector<int> list;
  list.reserve(p->h_weights.size());
  for (size_t i = 0; i < p->h_weights.size(); ++i)
    if (p->h_weights[i] != 0.0) list.push_back((int)i);  // NaN weights are kept (they propagate)
  p->num_good = (int)list.size();
  if (!list.empty())
    CUDA_TRY(cudaMemcpyAsync(p->d_chan_list, list.data(), sizeof(int) * list.size(),
                             cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  p->have_weights = true;
  p->gram_cache_valid = false;
  if (!p->tables_full) p->tables_dirty = true;  // tables built for the previous channel list
  return FB_OK;
}

extern "C" int fb_set_weights(fb_plan_t p, const double* data_dev, const uint8_t* good_host,
                              int32_t weighted_fit, int32_t range_lo, int32_t range_hi,
                              double* weights_host, double* chisq_initial_host, void* stream) {
  if (!p || !data_dev || !good_host) return fail(FB_ERR_INVALID, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int nf = p->desc.num_freq, nt = p->desc.num_time;
  // python slice semantics of data[:, lo:hi] (analysis/fitter.py:528)
  int lo = range_lo, hi = range_hi;
  if (lo < 0) lo += nt;
  if (hi < 0) hi += nt;
  lo = std::max(0, std::min(lo, nt));
  hi = std::max(0, std::min(hi, nt));
  if (hi < lo) hi = lo;
  CUDA_TRY(cudaMemcpyAsync(p->d_good, good_host, nf, cudaMemcpyHostToDevice, st));
  k_weights<<<nf, kBlock, 0, st>>>(data_dev, nf, nt, p->d_good, weighted_fit, lo, hi, p->d_weights,
                                   p->d_chan_sumsq);
  CUDA_TRY(cudaGetLastError());
  p->launches += 1;
  p->h_weights.resize(nf);
  std::vector<double> sumsq(nf);
  CUDA_TRY(cudaMemcpyAsync(p->h_weights.data(), p->d_weights, sizeof(double) * nf, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(sumsq.data(), p->d_chan_sumsq, sizeof(double) * nf, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  if (weights_host) memcpy(weights_host, p->h_weights.data(), sizeof(double) * nf);
  if (chisq_initial_host) {
    double s = 0.0;
    for (int i = 0; i < nf; ++i) s += sumsq[i];
    *chisq_initial_host = s;
  }
  return rebuild_channel_list(p, st);
}

extern "C" int fb_load_weights(fb_plan_t p, const double* weights_host, void* stream) {
  if (!p || !weights_host) return fail(FB_ERR_INVALID, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int nf = p->desc.num_freq;
  p->h_weights.assign(weights_host, weights_host + nf);
  CUDA_TRY(cudaMemcpyAsync(p->d_weights, p->h_weights.data(), sizeof(double) * nf, cudaMemcpyHostToDevice, st));
  return rebuild_channel_list(p, st);
}

extern "C" int fb_set_free_parameters(fb_plan_t p, const int32_t* free_mask, int32_t* num_free_out) {
  if (!p || !free_mask) return fail(FB_ERR_INVALID, "null argument");
  int saved[9];
  for (int k = 0; k < 9; ++k) {
    saved[k] = p->free_mask[k];
    p->free_mask[k] = free_mask[k] != 0;
  }
  layout_free(p);
  for (int k = 0; k < 9; ++k)
    if (saved[k] != p->free_mask[k]) {
      p->gram_cache_valid = false;
      p->hess_layout_valid = false;
      p->hess_desc_valid = false;
    }
  if (p->n_free > FB_MAX_FREE) {
    const int n = p->n_free;
    for (int k = 0; k < 9; ++k) p->free_mask[k] = saved[k];
    layout_free(p);
    return fail(FB_ERR_INVALID, "%d free parameters exceed FB_MAX_FREE=%d", n, FB_MAX_FREE);
  }
  if (num_free_out) *num_free_out = p->n_free;
  return FB_OK;
}

// ---- fast evaluation (fb_fast.cuh) ---------------------------------------------------------------

static int ensure_capacity(double** buf, size_t* cap, size_t need) {
  if (need <= *cap) return FB_OK;
  cudaFree(*buf);
  *buf = nullptr;
  *cap = 0;
  CUDA_TRY(cudaMalloc(buf, sizeof(double) * need));
  *cap = need;
  return FB_OK;
}

// Reduced-column layout of k_gram3 for the plan's free-parameter set.
static GramLayout gram3_layout(const fb_plan_s* p) {
  const int nc = p->desc.num_components;
  GramLayout G;
  int r = 0;
  const bool has_s = p->col_of[FB_AMPLITUDE] >= 0 || p->col_of[FB_SPECTRAL_INDEX] >= 0 ||
                     p->col_of[FB_SPECTRAL_RUNNING] >= 0;
  G.rc_S = has_s ? r : -1;
  r += has_s ? nc : 0;
  G.rc_t = p->col_of[FB_ARRIVAL_TIME] >= 0 ? r : -1;
  r += G.rc_t >= 0 ? nc : 0;
  G.rc_w = p->col_of[FB_BURST_WIDTH] >= 0 ? r : -1;
  r += G.rc_w >= 0 ? nc : 0;
  G.rc_dm = p->col_of[FB_DM] >= 0 ? r++ : -1;
  G.rc_dmidx = p->col_of[FB_DM_INDEX] >= 0 ? r++ : -1;
  G.rc_tau = p->col_of[FB_SCATTERING_TIMESCALE] >= 0 ? r++ : -1;
  G.rc_alpha = p->col_of[FB_SCATTERING_INDEX] >= 0 ? r++ : -1;
  G.rc_res = r++;
  G.nr = r;
  return G;
}

// The fast kernels cover the scattered, unfolded, non-scintillating model with up to 4 components.
static bool fast_path_ok(const fb_plan_s* p, bool for_gram) {
  if (env_off("FITBURST_B200_FAST")) return false;
  const int nc = p->desc.num_components;
  if (p->pc.scintillation || p->pc.is_folded || nc > 4) return false;
  if (!(p->h_params[FB_SCATTERING_TIMESCALE * nc] > 0.0)) return false;
  for (int c = 0; c < nc; ++c) {
    if (!(p->h_params[FB_BURST_WIDTH * nc + c] > 0.0)) return false;
    if (!(p->h_params[FB_REF_FREQ * nc + c] > 0.0)) return false;
  }
  if (for_gram && gram3_layout(p).nr > 24) return false;
  if (vals3_smem_bytes(nc, p->pc.kf) > 200 * 1024) return false;  // tail tables of nc * kf rows do not fit
  return true;
}

template <int KF, int NC, bool CHI2 = false>
static int launch_vals3_tn(fb_plan_s* p, const EvalArgs& A, double* out, cudaStream_t st, int* grid_out) {
  const int nc = p->pc.num_comp, kf = p->pc.kf;
  const size_t smem = vals3_smem_bytes(nc, kf, CHI2);
  if (smem > 227 * 1024) return fail(FB_ERR_INVALID, "shared-memory request exceeds 227 KB");
  static int occ_table[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // per instantiation (function-local static of the template)
  int& occ = occ_table[0];
  if (occ == 0 || KF == 0 || NC == 0 || smem > 48 * 1024) {
    if (smem > 48 * 1024)
      CUDA_TRY(cudaFuncSetAttribute(k_vals3<KF, NC, CHI2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_vals3<KF, NC, CHI2>, kFastBlock, smem));
    if (occ < 1) occ = 1;
  }
  const int grid = std::min(A.num_chan, p->sm_count * occ);
  k_vals3<KF, NC, CHI2><<<grid, kFastBlock, smem, st>>>(A, out);
  CUDA_TRY(cudaGetLastError());
  p->launches += 1;
  if (grid_out) *grid_out = grid;
  return FB_OK;
}

template <int KF>
static int launch_vals3_t(fb_plan_s* p, const EvalArgs& A, double* out, cudaStream_t st, int* grid_out) {
  if (KF == 0) return launch_vals3_tn<0, 0>(p, A, out, st, grid_out);
  switch (p->pc.num_comp) {
    case 1: return launch_vals3_tn<KF, 1>(p, A, out, st, grid_out);
    case 2: return launch_vals3_tn<KF, 2>(p, A, out, st, grid_out);
    case 3: return launch_vals3_tn<KF, 3>(p, A, out, st, grid_out);
    default: return launch_vals3_tn<KF, 0>(p, A, out, st, grid_out);
  }
}

static int launch_vals3(fb_plan_s* p, const EvalArgs& A, double* out, cudaStream_t st, int* grid_out) {
  switch (p->pc.kf) {
    case 1: return launch_vals3_t<1>(p, A, out, st, grid_out);
    case 2: return launch_vals3_t<2>(p, A, out, st, grid_out);
    case 4: return launch_vals3_t<4>(p, A, out, st, grid_out);
    case 8: return launch_vals3_t<8>(p, A, out, st, grid_out);
    default: return launch_vals3_t<0>(p, A, out, st, grid_out);
  }
}

// chi^2 variant of k_vals3 (grid sweeps): A.chi2_partials receives kFastWarps partials per block
// (at most 16 resident blocks of 128 threads per SM)
static int launch_chisq3(fb_plan_s* p, const EvalArgs& A, cudaStream_t st, int* grid) {
  const int nc = p->pc.num_comp;
  switch (p->pc.kf) {
    case 1: return launch_vals3_tn<1, 0, true>(p, A, nullptr, st, grid);
    case 2: return launch_vals3_tn<2, 0, true>(p, A, nullptr, st, grid);
    case 4: return launch_vals3_tn<4, 0, true>(p, A, nullptr, st, grid);
    case 8:
      if (nc == 1) return launch_vals3_tn<8, 1, true>(p, A, nullptr, st, grid);
      if (nc == 2) return launch_vals3_tn<8, 2, true>(p, A, nullptr, st, grid);
      if (nc == 3) return launch_vals3_tn<8, 3, true>(p, A, nullptr, st, grid);
      return launch_vals3_tn<8, 0, true>(p, A, nullptr, st, grid);
    default: return launch_vals3_tn<0, 0, true>(p, A, nullptr, st, grid);
  }
}

template <int NC, int NB>
static int launch_gram3_t(fb_plan_s* p, const EvalArgs& A, const GramLayout& G, const FarLayout& F, cudaStream_t st,
                          int* grid_out) {
  const size_t smem = gram3_smem_bytes(NC, NB, F.ncols);
  if (smem > 227 * 1024) return fail(FB_ERR_INVALID, "shared-memory request exceeds 227 KB");
  // row tiles through the bulk-copy (TMA) unit: opt-in (slower, see k_gram3), and only when every
  // row segment is 16-byte aligned and sized
  static const bool bulk_env = getenv("FITBURST_B200_BULK") && atoi(getenv("FITBURST_B200_BULK")) != 0;
  const bool bulk = bulk_env && p->pc.num_time % 2 == 0 && (reinterpret_cast<uintptr_t>(A.data) & 15) == 0;
  int& occ = p->occ_cache[16 + (NC - 1) * 3 + (NB - 1)];
  if (occ == 0 || smem > 48 * 1024) {
    if (smem > 48 * 1024) {
      CUDA_TRY(cudaFuncSetAttribute(k_gram3<NC, NB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      CUDA_TRY(cudaFuncSetAttribute(k_gram3<NC, NB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_gram3<NC, NB, false>, kFastBlock, smem));
    if (occ < 1) occ = 1;
  }
  const int grid = std::min(A.num_chan, p->sm_count * occ);
  const int nent = F.ncols * (F.ncols + 1) / 2;
  const size_t near_sz = (size_t)grid * kFastWarps * g3_nacc(NB) * 64;  // one accumulator set per warp
  int rc = ensure_capacity(&p->d_partials, &p->partials_cap, near_sz + (size_t)grid * nent);
  if (rc != FB_OK) return rc;
  EvalArgs B = A;
  B.partials = p->d_partials;
  B.bulk_rows = bulk ? 1 : 0;
  if (bulk) k_gram3<NC, NB, true><<<grid, kFastBlock, smem, st>>>(B, G, F, p->d_vals, p->d_partials + near_sz);
  else k_gram3<NC, NB, false><<<grid, kFastBlock, smem, st>>>(B, G, F, p->d_vals, p->d_partials + near_sz);
  CUDA_TRY(cudaGetLastError());
  p->launches += 1;
  *grid_out = grid;
  return FB_OK;
}

template <int NC>
static int launch_gram3_nc(fb_plan_s* p, const EvalArgs& A, const GramLayout& G, const FarLayout& F, int nb,
                           cudaStream_t st, int* grid_out) {
  switch (nb) {
    case 1: return launch_gram3_t<NC, 1>(p, A, G, F, st, grid_out);
    case 2: return launch_gram3_t<NC, 2>(p, A, G, F, st, grid_out);
    default: return launch_gram3_t<NC, 3>(p, A, G, F, st, grid_out);
  }
}

template <int KF>
static int launch_rise3_t(fb_plan_s* p, const EvalArgs& A, double* out, int* bounds, cudaStream_t st) {
  const int nc = p->pc.num_comp, kf = p->pc.kf;
  const size_t smem = rise3_smem_bytes(nc, kf);
  if (smem > 227 * 1024) return fail(FB_ERR_INVALID, "shared-memory request exceeds 227 KB");
  int& occ = p->occ_cache[48 + (KF == 0 ? 0 : KF == 1 ? 1 : KF == 2 ? 2 : KF == 4 ? 3 : 4)];
  if (occ == 0 || KF == 0) {
    if (smem > 48 * 1024)
      CUDA_TRY(cudaFuncSetAttribute(k_rise3<KF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_rise3<KF>, kFastBlock, smem));
    if (occ < 1) occ = 1;
  }
  const int grid = std::min(A.num_chan, p->sm_count * occ);
  k_rise3<KF><<<grid, kFastBlock, smem, st>>>(A, out, bounds);
  CUDA_TRY(cudaGetLastError());
  p->launches += 1;
  return FB_OK;
}

static int launch_rise3(fb_plan_s* p, const EvalArgs& A, double* out, int* bounds, cudaStream_t st) {
  switch (p->pc.kf) {
    case 1: return launch_rise3_t<1>(p, A, out, bounds, st);
    case 2: return launch_rise3_t<2>(p, A, out, bounds, st);
    case 4: return launch_rise3_t<4>(p, A, out, bounds, st);
    case 8: return launch_rise3_t<8>(p, A, out, bounds, st);
    default: return launch_rise3_t<0>(p, A, out, bounds, st);
  }
}

template <int NC, int NB>
static int launch_gram4_t(fb_plan_s* p, const EvalArgs& A, const GramLayout& G, const FarLayout& F, cudaStream_t st,
                          int* grid_out) {
  const size_t smem = gram4_smem_bytes(NC, NB, p->pc.kf, F.ncols);
  if (smem > 227 * 1024) return fail(FB_ERR_INVALID, "shared-memory request exceeds 227 KB");
  int& occ = p->occ_cache[28 + (NC - 1) * 3 + (NB - 1)];
  if (occ == 0 || smem > 48 * 1024) {
    if (smem > 48 * 1024)
      CUDA_TRY(cudaFuncSetAttribute(k_gram4<NC, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_gram4<NC, NB>, kFastBlock, smem));
    if (occ < 1) occ = 1;
  }
  const int grid = std::min(A.num_chan, p->sm_count * occ);
  const int nent = F.ncols * (F.ncols + 1) / 2;
  const size_t near_sz = (size_t)grid * kFastWarps * g3_nacc(NB) * 64;
  int rc = ensure_capacity(&p->d_partials, &p->partials_cap, near_sz + (size_t)grid * nent);
  if (rc != FB_OK) return rc;
  EvalArgs B = A;
  B.partials = p->d_partials;
  k_gram4<NC, NB><<<grid, kFastBlock, smem, st>>>(B, G, F, p->d_vals, p->d_bounds, p->d_partials + near_sz);
  CUDA_TRY(cudaGetLastError());
  p->launches += 1;
  *grid_out = grid;
  return FB_OK;
}

template <int NC>
static int launch_gram4_nc(fb_plan_s* p, const EvalArgs& A, const GramLayout& G, const FarLayout& F, int nb,
                           cudaStream_t st, int* grid_out) {
  switch (nb) {
    case 1: return launch_gram4_t<NC, 1>(p, A, G, F, st, grid_out);
    case 2: return launch_gram4_t<NC, 2>(p, A, G, F, st, grid_out);
    default: return launch_gram4_t<NC, 3>(p, A, G, F, st, grid_out);
  }
}

// FITBURST_B200_EVAL=4 selects k_rise3 + k_gram4 (no spectrum-row round trip through HBM) instead
// of k_vals3 + k_gram3. Measured on B200 at config 2 it cuts the DRAM traffic of an evaluation
// but is SLOWER (0.54 ms against 0.42 ms: k_gram4 438 us + k_rise3 92 us against 206 + 206 us --
// the dot products cost more issue slots at 32-sample tiles than the round trip costs time at
// 1.4 of 6.5 TB/s), so it is opt-in; see profiles/README.md.
static bool eval_gen4(const fb_plan_s* p) {
  const char* v = getenv("FITBURST_B200_EVAL");
  if (!(v && !strcmp(v, "4"))) return false;
  return gram4_smem_bytes(p->pc.num_comp, 3, p->pc.kf, p->n_free + 1) <= 200 * 1024;
}

// k_vals3 + k_gram3 + reduction + expansion into the full (n+1)^2 matrix at gram_dev.
static int launch_fast_gram(fb_plan_s* p, EvalArgs& A, int ncols, double* gram_dev, cudaStream_t st) {
  int rc = ensure_tables(p, st, true);
  if (rc != FB_OK) return rc;
  const int nc = p->pc.num_comp, nt = p->pc.num_time;
  rc = ensure_capacity(&p->d_vals, &p->vals_cap, (size_t)A.num_chan * nc * nt);
  if (rc != FB_OK) return rc;
  const bool gen4 = eval_gen4(p);
  if (gen4) {
    const size_t need_b = (size_t)A.num_chan * nc * 2;
    if (need_b > p->bounds_cap) {
      cudaFree(p->d_bounds);
      p->d_bounds = nullptr;
      p->bounds_cap = 0;
      CUDA_TRY(cudaMalloc(&p->d_bounds, sizeof(int) * need_b));
      p->bounds_cap = need_b;
    }
    rc = launch_rise3(p, A, p->d_vals, p->d_bounds, st);
    p->last_eval_split = false;  // d_vals holds rise-region values only: nothing for k_hess3 to reuse
  } else {
    rc = launch_vals3(p, A, p->d_vals, st, nullptr);
  }
  if (rc != FB_OK) return rc;
  const GramLayout G = gram3_layout(p);
  const int nb = (G.nr + 7) / 8;
  // far-tile path: full-column descriptors (2 nc + 1 basis values must fit one 8-column block)
  FarLayout F;
  memset(&F, 0, sizeof(F));
  F.ncols = ncols;
  const int nent = ncols * (ncols + 1) / 2;
  F.enabled = (2 * nc + 1 <= 8 && nent <= 2 * kFastBlock && !env_off("FITBURST_B200_FAR")) ? 1 : 0;
  for (int k = 0; k < 9; ++k) {
    const int col = p->col_of[k];
    if (col < 0) continue;
    if (is_fit_global(k)) {
      F.param[col] = (int8_t)k;
      F.comp[col] = -1;
    } else {
      for (int c = 0; c < nc; ++c) {
        F.param[col + c] = (int8_t)k;
        F.comp[col + c] = (int8_t)c;
      }
    }
  }
  F.param[ncols - 1] = 9;
  F.comp[ncols - 1] = -1;
  int grid = 0;
  if (gen4) {
    switch (nc) {
      case 1: rc = launch_gram4_nc<1>(p, A, G, F, nb, st, &grid); break;
      case 2: rc = launch_gram4_nc<2>(p, A, G, F, nb, st, &grid); break;
      case 3: rc = launch_gram4_nc<3>(p, A, G, F, nb, st, &grid); break;
      default: rc = launch_gram4_nc<4>(p, A, G, F, nb, st, &grid); break;
    }
  } else {
    switch (nc) {
      case 1: rc = launch_gram3_nc<1>(p, A, G, F, nb, st, &grid); break;
      case 2: rc = launch_gram3_nc<2>(p, A, G, F, nb, st, &grid); break;
      case 3: rc = launch_gram3_nc<3>(p, A, G, F, nb, st, &grid); break;
      default: rc = launch_gram3_nc<4>(p, A, G, F, nb, st, &grid); break;
    }
  }
  if (rc != FB_OK) return rc;
  const int entries = g3_nacc(nb) * 64;
  ExpandArgs X;
  memset(&X, 0, sizeof(X));
  X.ncols = ncols;
  X.nb = nb;
  const double ref0 = p->h_params[FB_REF_FREQ * nc];
  for (int k = 0; k < 9; ++k) {
    const int col = p->col_of[k];
    if (col < 0) continue;
    if (is_fit_global(k)) {
      X.red[col] = k == FB_DM ? G.rc_dm : k == FB_DM_INDEX ? G.rc_dmidx : k == FB_SCATTERING_TIMESCALE ? G.rc_tau : G.rc_alpha;
      X.power[col] = 0;
      X.factor[col] = 1.0;
      X.shift[col] = 0.0;
      continue;
    }
    for (int c = 0; c < nc; ++c) {
      const bool spectral = k == FB_AMPLITUDE || k == FB_SPECTRAL_INDEX || k == FB_SPECTRAL_RUNNING;
      X.red[col + c] = (spectral ? G.rc_S : k == FB_ARRIVAL_TIME ? G.rc_t : G.rc_w) + c;
      X.power[col + c] = k == FB_SPECTRAL_INDEX ? 1 : k == FB_SPECTRAL_RUNNING ? 2 : 0;
      X.factor[col + c] = k == FB_AMPLITUDE ? kLn10 : 1.0;
      // ln(f / ref_c) = ln(f / ref_0) + ln(ref_0 / ref_c)
      X.shift[col + c] = spectral ? log(ref0 / p->h_params[FB_REF_FREQ * nc + c]) : 0.0;
    }
  }
  X.red[ncols - 1] = G.rc_res;
  X.power[ncols - 1] = 0;
  X.factor[ncols - 1] = 1.0;
  X.shift[ncols - 1] = 0.0;
  const int far_entries = F.enabled ? nent : 0;
  k_gram3_finish<<<dim3(gram3_finish_groups(entries, far_entries), kFinishSlices), 256, 0, st>>>(
      p->d_partials, grid, kFastWarps, gen4 ? 1 : kFastWarps, entries, far_entries, p->d_slices, p->d_racc, p->d_far, X,
      p->d_counter, gram_dev,
      p->pc.stop);
  CUDA_TRY(cudaGetLastError());
  p->launches += 1;
  return FB_OK;
}

// Two-kernel evaluation (k_vals + k_gram2 / k_gram2_mma), see fb_kernels.cuh "split evaluation".
// *mma_warps_out > 0 tells the caller to finalize with k_gram_finalize_mma.
template <int NB>
static int launch_gram2_mma(fb_plan_s* p, const EvalArgs& A, int ncols, long long items, cudaStream_t st, int* warps_out) {
  const int nc = p->pc.num_comp, kf = p->pc.kf;
  const size_t smem = gram_mma_smem_bytes(nc, kf, ncols, A.vals_spc == 2 ? 2 : 3);
  if (smem > 227 * 1024) return fail(FB_ERR_INVALID, "shared-memory request exceeds 227 KB");
  if (smem > 48 * 1024)
    CUDA_TRY(cudaFuncSetAttribute(k_gram2_mma<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 1;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_gram2_mma<NB>, kBlock, smem));
  const long long grid = std::min<long long>(items, (long long)p->sm_count * std::max(1, occ));
  const int warps = (int)grid * (kBlock / 32);
  const size_t need = (size_t)warps * (NB * (NB + 1) / 2) * 64;
  if (need > p->partials_cap) {
    cudaFree(p->d_partials);
    p->d_partials = nullptr;
    p->partials_cap = 0;
    CUDA_TRY(cudaMalloc(&p->d_partials, sizeof(double) * need));
    p->partials_cap = need;
  }
  EvalArgs B = A;
  B.partials = p->d_partials;
  k_gram2_mma<NB><<<(int)grid, kBlock, smem, st>>>(B, p->d_vals);
  CUDA_TRY(cudaGetLastError());
  *warps_out = warps;
  return FB_OK;
}

static int launch_split_gram(fb_plan_s* p, EvalArgs& A, int ncols, cudaStream_t st, int* grid_out, int* mma_warps_out) {
  int rc = ensure_tables(p, st, true);
  if (rc != FB_OK) return rc;
  const int nc = p->pc.num_comp, kf = p->pc.kf;
  const int tiles = (p->pc.num_time + kBlock - 1) / kBlock;
  const long long items = (long long)A.num_chan * tiles;
  const bool scint = p->pc.scintillation != 0;
  A.vals_spc = 2;  // spectrum (scintillation: mean profile, see EvalArgs.vals_spc) + time difference
  const size_t need_vals = (size_t)items * kBlock * 2 * nc;
  if (need_vals > p->vals_cap) {
    cudaFree(p->d_vals);
    p->d_vals = nullptr;
    p->vals_cap = 0;
    CUDA_TRY(cudaMalloc(&p->d_vals, sizeof(double) * need_vals));
    p->vals_cap = need_vals;
  }
  const size_t smem1 = smem_bytes(nc, kf, 1, false, false, SM_QUEUE);
  if (smem1 > 227 * 1024) return fail(FB_ERR_INVALID, "shared-memory request exceeds 227 KB");
  if (smem1 > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(k_vals, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
  int occ1 = 1;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ1, k_vals, kBlock, smem1));
  const long long grid1 = std::min<long long>(items, (long long)p->sm_count * std::max(1, occ1) * 4);
  k_vals<<<(int)grid1, kBlock, smem1, st>>>(A, p->d_vals);
  CUDA_TRY(cudaGetLastError());
  p->launches += 2;
  if (scint) {
    // per-channel amplitudes (routines/ism.py:11-48) from the stored mean profiles
    if (!p->d_scint_amps) CUDA_TRY(cudaMalloc(&p->d_scint_amps, sizeof(double) * (size_t)p->desc.num_freq * nc));
    const size_t smem_s = scint_vals_smem_bytes(nc);
    if (smem_s > 48 * 1024)
      CUDA_TRY(cudaFuncSetAttribute(k_scint_partials_vals, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s));
    int occ_s = 1;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_s, k_scint_partials_vals, kBlock, smem_s));
    if (occ_s < 1) occ_s = 1;
    const long long slots = (long long)p->sm_count * occ_s * 4;
    const int nchunks = (int)std::min<long long>(tiles, std::max<long long>(1, (slots + A.num_chan - 1) / A.num_chan));
    const size_t nent = (size_t)gram_ntiles(nc + 1) * kTile * kTile;
    rc = ensure_capacity(&p->d_scint_partials, &p->scint_partials_cap, (size_t)A.num_chan * nchunks * nent);
    if (rc != FB_OK) return rc;
    const int grid_s = (int)std::min<long long>((long long)A.num_chan * nchunks, (long long)p->sm_count * occ_s);
    k_scint_partials_vals<<<grid_s, kBlock, smem_s, st>>>(A, nchunks, p->d_vals, p->d_scint_partials);
    CUDA_TRY(cudaGetLastError());
    k_scint_solve<<<A.num_chan, 64, 0, st>>>(A, nchunks, p->d_scint_partials, p->d_scint_amps);
    CUDA_TRY(cudaGetLastError());
    p->launches += 2;
    A.scint_amps = p->d_scint_amps;
  }
  *mma_warps_out = 0;
  const int nb = (ncols + 7) / 8;
  const char* mma_env = getenv("FITBURST_B200_DMMA");
  if (nb <= 4 && !(mma_env && !strcmp(mma_env, "0"))) {
    switch (nb) {
      case 1: return launch_gram2_mma<1>(p, A, ncols, items, st, mma_warps_out);
      case 2: return launch_gram2_mma<2>(p, A, ncols, items, st, mma_warps_out);
      case 3: return launch_gram2_mma<3>(p, A, ncols, items, st, mma_warps_out);
      default: return launch_gram2_mma<4>(p, A, ncols, items, st, mma_warps_out);
    }
  }
  const size_t smem2 = smem_bytes(nc, kf, ncols, false, false, SM_PARK | SM_X);
  if (smem2 > 227 * 1024) return fail(FB_ERR_INVALID, "shared-memory request exceeds 227 KB");
  if (smem2 > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(k_gram2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
  int occ2 = 1;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, k_gram2, kBlock, smem2));
  const long long grid2 = std::min<long long>(items, (long long)p->sm_count * std::max(1, occ2));
  const size_t need = (size_t)grid2 * gram_ntiles(ncols) * kTile * kTile;
  if (need > p->partials_cap) {
    cudaFree(p->d_partials);
    p->d_partials = nullptr;
    p->partials_cap = 0;
    CUDA_TRY(cudaMalloc(&p->d_partials, sizeof(double) * need));
    p->partials_cap = need;
  }
  A.partials = p->d_partials;
  k_gram2<<<(int)grid2, kBlock, smem2, st>>>(A, p->d_vals);
  CUDA_TRY(cudaGetLastError());
  *grid_out = (int)grid2;
  return FB_OK;
}

static int normal_equations_async(fb_plan_s* p, const double* data_dev, double* gram_dev, cudaStream_t st) {
  if (!p->have_weights) return fail(FB_ERR_INVALID, "fb_set_weights / fb_load_weights must be called first");
  if (p->n_free < 1) return fail(FB_ERR_INVALID, "no free parameters");
  const int ncols = p->n_free + 1;
  if (p->num_good == 0) {
    CUDA_TRY(cudaMemsetAsync(gram_dev, 0, sizeof(double) * ncols * ncols, st));
    return FB_OK;
  }
  EvalArgs A = base_args(p);
  A.data = data_dev;
  A.weights = p->d_weights;
  A.chan_list = p->d_chan_list;
  A.num_chan = p->num_good;
  int grid = 0;
  int rc = FB_OK;
  const char* split_env = getenv("FITBURST_B200_SPLIT");
  // scintillation: split as well unless the amplitudes are to be solved in place
  // (FITBURST_B200_SCINT_PRESOLVE=0) or the stored values would not fit (2 doubles per (sample, component))
  const bool scint_split = !p->pc.scintillation ||
                           (!env_off("FITBURST_B200_SCINT_PRESOLVE") &&
                            (size_t)p->num_good * p->pc.num_time * p->pc.num_comp * 16 <= ((size_t)64 << 30));
  const bool split = scint_split && !(split_env && !strcmp(split_env, "0"));
  int mma_warps = 0;
  p->last_eval_split = split && !p->pc.scintillation;  // (scintillation stores 3 planes: never reused)
  p->vals_match_cache = false;
  if (split && fast_path_ok(p, true)) {
    p->vals_layout = 3;
    return launch_fast_gram(p, A, ncols, gram_dev, st);
  }
  p->vals_layout = 2;
  if (split) {
    rc = launch_split_gram(p, A, ncols, st, &grid, &mma_warps);
  } else {
    rc = launch_eval<MODE_GRAM>(p, A, ncols, st, &grid);
  }
  if (rc != FB_OK) return rc;
  if (mma_warps > 0) {
    const int nb = (ncols + 7) / 8, entries = nb * (nb + 1) / 2 * 64;
    k_gram_finalize_mma<<<(entries + 7) / 8, 256, 0, st>>>(p->d_partials, mma_warps, ncols, nb, gram_dev, p->pc.stop);
  } else {
    const int entries = gram_ntiles(ncols) * kTile * kTile;
    k_gram_finalize<<<(entries + 7) / 8, 256, 0, st>>>(p->d_partials, grid, ncols, gram_dev, p->pc.stop);
  }
  CUDA_TRY(cudaGetLastError());
  p->launches += 1;
  return FB_OK;
}

extern "C" int fb_normal_equations_dev(fb_plan_t p, const double* data_dev, double* gram_dev, void* stream) {
  if (!p || !data_dev || !gram_dev) return fail(FB_ERR_INVALID, "null argument");
  return normal_equations_async(p, data_dev, gram_dev, (cudaStream_t)stream);
}

extern "C" int fb_normal_equations(fb_plan_t p, const double* data_dev, double* gram_host, void* stream) {
  if (!p || !data_dev || !gram_host) return fail(FB_ERR_INVALID, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = normal_equations_async(p, data_dev, p->d_gram, st);
  if (rc != FB_OK) return rc;
  const int ncols = p->n_free + 1;
  CUDA_TRY(cudaMemcpyAsync(p->h_gram, p->d_gram, sizeof(double) * ncols * ncols, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  memcpy(gram_host, p->h_gram, sizeof(double) * ncols * ncols);
  p->gram_cache.assign(p->h_gram, p->h_gram + (size_t)ncols * ncols);
  p->gram_cache_params.assign(p->h_params, p->h_params + (size_t)FB_NUM_PARAMETERS * p->desc.num_components);
  p->gram_cache_data = data_dev;
  p->gram_cache_valid = true;
  p->vals_match_cache = p->last_eval_split;
  return check_err_flag(p, st);
}

static int dense_common(fb_plan_s* p, const double* data_dev, double* jac_dev, double* resid_dev, cudaStream_t st) {
  if (!p->have_weights) return fail(FB_ERR_INVALID, "fb_set_weights / fb_load_weights must be called first");
  EvalArgs A = base_args(p);
  A.data = data_dev;
  A.weights = p->d_weights;
  A.jac = jac_dev;
  A.resid = resid_dev;
  if (!jac_dev) {
    A.n_free = 0;
    for (int k = 0; k < 9; ++k) A.col_of[k] = -1;
  }
  int rc = launch_eval<MODE_JAC>(p, A, A.n_free + 1, st, nullptr);
  if (rc != FB_OK) return rc;
  return check_err_flag(p, st);
}

extern "C" int fb_residuals(fb_plan_t p, const double* data_dev, double* resid_dev, void* stream) {
  if (!p || !data_dev || !resid_dev) return fail(FB_ERR_INVALID, "null argument");
  return dense_common(p, data_dev, nullptr, resid_dev, (cudaStream_t)stream);
}

extern "C" int fb_jacobian_dense(fb_plan_t p, const double* data_dev, double* jac_dev, void* stream) {
  if (!p || !data_dev || !jac_dev) return fail(FB_ERR_INVALID, "null argument");
  if (p->n_free < 1) return fail(FB_ERR_INVALID, "no free parameters");
  return dense_common(p, data_dev, jac_dev, nullptr, (cudaStream_t)stream);
}

extern "C" int fb_derivative_map(fb_plan_t p, const double* data_dev, int32_t param, int32_t component,
                                 int32_t add_all, double* out_dev, void* stream) {
  if (!p || !out_dev) return fail(FB_ERR_INVALID, "null argument");
  if (param < 0 || param > 8) return fail(FB_ERR_INVALID, "parameter index %d out of range", param);
  if (component < 0 || component >= p->desc.num_components)
    return fail(FB_ERR_INVALID, "component %d out of range", component);
  if (p->pc.scintillation && !data_dev) return fail(FB_ERR_INVALID, "scintillation mode needs data");
  EvalArgs A = base_args(p);
  A.data = data_dev;
  A.d_param = param;
  A.d_comp = component;
  A.d_addall = add_all;
  A.d_out = out_dev;
  A.n_free = 0;
  int rc = launch_eval<MODE_DERIV>(p, A, 1, (cudaStream_t)stream, nullptr);
  if (rc != FB_OK) return rc;
  return check_err_flag(p, (cudaStream_t)stream);
}

// Packs / unpacks the free vector in LSFitter.get_fit_parameters_list order
// (analysis/fitter.py:353-430) against the plan's host copy of the parameter block.
static void pack_free(const fb_plan_s* p, double* x) {
  const int nc = p->desc.num_components;
  for (int k = 0; k < 9; ++k) {
    if (p->col_of[k] < 0) continue;
    if (is_fit_global(k)) x[p->col_of[k]] = p->h_params[k * nc];
    else
      for (int c = 0; c < nc; ++c) x[p->col_of[k] + c] = p->h_params[k * nc + c];
  }
}

static void unpack_free(fb_plan_s* p, const double* x) {
  const int nc = p->desc.num_components;
  for (int k = 0; k < 9; ++k) {
    if (p->col_of[k] < 0) continue;
    if (is_fit_global(k)) {
      // load_fit_parameters_list returns a length-1 list; update_parameters then stores that
      // list, and the model reads element 0 (analysis/model.py:155-159).
      for (int c = 0; c < nc; ++c) p->h_params[k * nc + c] = x[p->col_of[k]];
    } else {
      for (int c = 0; c < nc; ++c) p->h_params[k * nc + c] = x[p->col_of[k] + c];
    }
  }
}

int fb_hessian_impl(fb_plan_s* p, const double* data_dev, double* hessian_host, cudaStream_t st);

static fb_fit_options default_options(const fb_fit_options* options);

// Host-driven trust-region loop (one stream synchronisation per trial point): the scintillation
// mode, whose singular-system flag must be read after every evaluation, and the continuation of a
// device-driven fit whose trial point left the domain of the fast kernels (pending_trial: S.x_new
// is proposed but not evaluated yet). `gram` holds the last evaluated matrix on return.
static int fit_host_loop(fb_plan_s* p, const double* data_dev, TrfState& S, std::vector<double>& gram,
                         std::vector<double>& accepted, bool pending_trial, void* stream) {
  // FITBURST_B200_TIMING=1: host-side breakdown of the loop on stderr
  const bool timing = getenv("FITBURST_B200_TIMING") != nullptr;
  using clk = std::chrono::steady_clock;
  double t_host = 0.0, t_eval = 0.0;
  auto t_prev = clk::now();
  auto lap = [&](double& acc) {
    const auto now = clk::now();
    acc += std::chrono::duration<double, std::micro>(now - t_prev).count();
    t_prev = now;
  };
  for (;;) {
    if (!pending_trial) {
      const bool more = trf_propose(S);
      if (timing) lap(t_host);
      if (!more) break;
    }
    pending_trial = false;
    unpack_free(p, S.x_new);
    int rc = fb_set_parameters(p, p->h_params, stream);
    if (rc == FB_OK) rc = fb_normal_equations(p, data_dev, gram.data(), stream);
    if (rc != FB_OK) {
      if (getenv("FITBURST_B200_DEBUG")) {  // the trial point an evaluation failed at
        fprintf(stderr, "fb_fit: evaluation %d failed (%s) at parameters", S.nfev + 1, fb_last_error());
        for (int k = 0; k < FB_NUM_PARAMETERS * p->desc.num_components; ++k) fprintf(stderr, " %.17g", p->h_params[k]);
        fprintf(stderr, "\n");
      }
      return rc;
    }
    if (timing) lap(t_eval);
    const int njev_before = S.njev;
    trf_update(S, gram.data());
    if (S.njev != njev_before) accepted = gram;
    if (timing) lap(t_host);
  }
  if (timing)
    fprintf(stderr, "fb_fit: nfev %d, host TRF %.1f us, evaluations (launch + wait) %.1f us\n", S.nfev, t_host, t_eval);
  return FB_OK;
}

static int ensure_device_fit(fb_plan_s* p) {
  if (p->d_trf) return FB_OK;
  const size_t gsz = sizeof(double) * (FB_MAX_FREE + 1) * (FB_MAX_FREE + 1);
  CUDA_TRY(cudaMalloc(&p->d_trf, sizeof(TrfState)));
  CUDA_TRY(cudaMalloc(&p->d_fit_flags, sizeof(int) * kFitFlagCount));
  CUDA_TRY(cudaMalloc(&p->d_gram_accepted, gsz));
  CUDA_TRY(cudaMalloc(&p->d_fit_result, sizeof(fb_fit_result)));
  CUDA_TRY(cudaHostAlloc(&p->h_fit_flags, sizeof(int) * 8 * kFitFlagCount, cudaHostAllocMapped));
  CUDA_TRY(cudaHostGetDevicePointer(&p->h_fit_flags_dev, p->h_fit_flags, 0));
  CUDA_TRY(cudaMallocHost(&p->h_fit_result, sizeof(fb_fit_result)));
  CUDA_TRY(cudaMallocHost(&p->h_fit_grams, 2 * gsz));
  CUDA_TRY(cudaMallocHost(&p->h_fit_params, sizeof(double) * FB_NUM_PARAMETERS * FB_MAX_COMPONENTS));
  // (FITBURST_B200_POLL=0 only) blocking-sync events: a host thread that follows a fit sleeps instead
  // of spinning inside the driver
  for (int k = 0; k < 8; ++k)
    CUDA_TRY(cudaEventCreateWithFlags(&p->fit_events[k], cudaEventDisableTiming | cudaEventBlockingSync));
  return FB_OK;
}

// Arguments of k_trf_single for the plan's current free-parameter layout; resets the flags and
// arms the early-out of the evaluation kernels.
static int device_fit_setup(fb_plan_s* p, const fb_fit_options& opt, long long num_residuals, SingleTrfArgs& T,
                            cudaStream_t st) {
  int rc = ensure_device_fit(p);
  if (rc != FB_OK) return rc;
  const int n = p->n_free;
  CUDA_TRY(cudaMemsetAsync(p->d_fit_flags, 0, sizeof(int) * kFitFlagCount, st));
  memset(&T, 0, sizeof(T));
  T.nc = p->desc.num_components;
  T.n_free = n;
  for (int k = 0; k < 9; ++k) T.col_of[k] = p->col_of[k];
  T.num_residuals = num_residuals;
  T.ftol = opt.ftol;
  T.xtol = opt.xtol;
  T.gtol = opt.gtol;
  T.max_nfev = opt.max_nfev;
  const bool split = !p->pc.scintillation && !env_off("FITBURST_B200_SPLIT");
  T.guard_fast = (split && fast_path_ok(p, true)) ? 1 : 0;
  T.gram = p->d_gram;
  T.gram_accepted = p->d_gram_accepted;
  T.params = p->d_params;
  T.state = p->d_trf;
  T.flags = p->d_fit_flags;
  T.result = p->d_fit_result;
  const size_t trf_smem = trf_single_smem_bytes(n);
  if (trf_smem > 48 * 1024)
    CUDA_TRY(cudaFuncSetAttribute(k_trf_single, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)trf_smem));
  CUDA_TRY(cudaFuncSetAttribute(k_trf_block, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)trf_block_smem_bytes(n)));
  CUDA_TRY(cudaFuncSetAttribute(k_trf_chol, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)trf_chol_smem_bytes(n)));
  p->pc.stop = p->d_fit_flags;
  p->gram_cache_valid = false;
  return FB_OK;
}

// The trust-region step on the device. Default: k_trf_chol (More's iteration on Cholesky
// factorisations, ~10-30 us). FITBURST_B200_TRF=block / warp select the eigen-decomposition forms
// (block-parallel with warm-started Jacobi unless FITBURST_B200_TRF_WARM=0: 130-190 us; one warp:
// 340 us) that restate scipy's SVD route literally -- kept as cross-checks.
static int launch_trf_step(fb_plan_s* p, const SingleTrfArgs& T, cudaStream_t st) {
  const char* mode = getenv("FITBURST_B200_TRF");
  if (mode && !strcmp(mode, "warp"))
    k_trf_single<<<1, 32, trf_single_smem_bytes(T.n_free), st>>>(T);
  else if (mode && !strcmp(mode, "block"))
    k_trf_block<<<1, kTrfThreads, trf_block_smem_bytes(T.n_free), st>>>(T, env_off("FITBURST_B200_TRF_WARM") ? 0 : 1);
  else
    k_trf_chol<<<1, 32, trf_chol_smem_bytes(T.n_free), st>>>(T);
  CUDA_TRY(cudaGetLastError());
  p->launches += 1;
  return FB_OK;
}

// One evaluation at the plan's DEVICE parameter block followed by the trust-region step on the
// device, which also writes the flags to host slot `slot`. Nothing here waits for the GPU.
static int enqueue_eval_and_step(fb_plan_s* p, const double* data_dev, const SingleTrfArgs& T, int slot, int seq,
                                 cudaStream_t st) {
  p->tables_dirty = true;  // the parameter block changes on the device between evaluations
  int rc = normal_equations_async(p, data_dev, p->d_gram, st);
  if (rc != FB_OK) return rc;
  SingleTrfArgs T2 = T;
  T2.host_flags = p->h_fit_flags_dev + slot * kFitFlagCount;
  T2.seq = seq;
  return launch_trf_step(p, T2, st);
}

extern "C" int fb_fit(fb_plan_t p, const double* data_dev, const fb_fit_options* options,
                      fb_fit_result* result, double* gram_final_host, double* hessian_host,
                      void* stream) {
  if (!p || !data_dev || !result) return fail(FB_ERR_INVALID, "null argument");
  if (!p->have_weights) return fail(FB_ERR_INVALID, "fb_set_weights / fb_load_weights must be called first");
  const int n = p->n_free;
  if (n < 1) return fail(FB_ERR_INVALID, "no free parameters");
  cudaStream_t st = (cudaStream_t)stream;
  const fb_fit_options opt = default_options(options);
  const int ncols = n + 1;
  const size_t gcount = (size_t)ncols * ncols;
  const size_t nparam = (size_t)FB_NUM_PARAMETERS * p->desc.num_components;
  const long long m = (long long)p->desc.num_freq * p->desc.num_time;
  std::vector<double> gram(gcount), accepted(gcount);
  TrfState* stt = new (std::nothrow) TrfState();
  if (!stt) return fail(FB_ERR_NOMEM, "out of host memory");
  struct Guard {
    TrfState* s;
    fb_plan_s* p;
    ~Guard() {
      delete s;
      p->pc.stop = nullptr;
    }
  } guard{stt, p};
  TrfState& S = *stt;
  memset(result, 0, sizeof(*result));
  result->num_free = n;
  int rc = fb_set_parameters(p, p->h_params, stream);
  if (rc != FB_OK) return rc;
  // FITBURST_B200_TIMING=1: wall-clock breakdown of the call on stderr
  const bool fit_timing = getenv("FITBURST_B200_TIMING") != nullptr;
  const auto t_begin = std::chrono::steady_clock::now();
  auto since = [&](const std::chrono::steady_clock::time_point& t0) {
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
  };
  double us_loop = 0.0, us_readback = 0.0;

  // The loop stays on the device unless the mode needs a host check after every evaluation
  // (scintillation: singular per-channel systems) or FITBURST_B200_DEVICE_LOOP=0 asks for the
  // host-driven loop (kept as the cross-check of the device driver in the tests).
  const bool device_loop = !p->pc.scintillation && p->num_good > 0 && !env_off("FITBURST_B200_DEVICE_LOOP");
  bool pending_trial = false, need_host_loop = !device_loop;
  if (device_loop) {
    SingleTrfArgs T;
    rc = device_fit_setup(p, opt, m, T, st);
    if (rc != FB_OK) return rc;
    // Evaluations are enqueued AHEAD of what is known to be needed: the first kFirst at once
    // (fits rarely end earlier), then one more whenever the flags of evaluation e - kAhead have
    // come back without the stop bit, so the GPU always has the next evaluation queued and at
    // most kAhead evaluations (no-op launches of a few microseconds) are enqueued past the end.
    constexpr int kRing = 8;
    static const int kFirst = std::max(1, std::min(6, getenv("FITBURST_B200_FIRST") ? atoi(getenv("FITBURST_B200_FIRST")) : 5));
    static const int kAhead = std::max(1, std::min(6, getenv("FITBURST_B200_AHEAD") ? atoi(getenv("FITBURST_B200_AHEAD")) : 2));
    // The host follows the fit through the page-locked flag slots the step kernel writes: a short
    // spin on that memory (no CUDA call, microsecond reaction when the step has already run, which
    // is the rule with two evaluations of look-ahead), then a blocking wait on the evaluation's
    // event so that the threads of several fits in flight sleep instead of burning host cores.
    // FITBURST_B200_POLL=1: spin only; =0: events only.
    static const int kPollMode = getenv("FITBURST_B200_POLL") ? atoi(getenv("FITBURST_B200_POLL")) : 2;
    static const bool kPoll = kPollMode != 0;
    const int max_evals = (opt.max_nfev > 0 ? opt.max_nfev : 100 * n) + 2;
    int enqueued = 0;
    volatile int* hflags = reinterpret_cast<volatile int*>(p->h_fit_flags);
    for (int k = 0; k < kRing; ++k) hflags[k * kFitFlagCount + kFitSeq] = -1;
    auto enqueue_one = [&]() -> int {
      const int slot = enqueued % kRing;
      const int rc2 = enqueue_eval_and_step(p, data_dev, T, slot, enqueued, st);
      if (rc2 != FB_OK) return rc2;
      if (kPollMode != 1) CUDA_TRY(cudaEventRecord(p->fit_events[slot], st));
      ++enqueued;
      return FB_OK;
    };
    for (int k = 0; k < kFirst; ++k) {
      rc = enqueue_one();
      if (rc != FB_OK) return rc;
    }
    for (int checked = 0;; ++checked) {  // `checked`: the evaluation whose flags are inspected next
      if (checked >= enqueued) return fail(FB_ERR_CUDA, "device-driven fit did not terminate within %d evaluations", max_evals);
      const int slot = checked % kRing;
      if (kPoll) {
        // the step kernel writes the flags and then its sequence number into page-locked host memory
        long spins = 0;
        while (hflags[slot * kFitFlagCount + kFitSeq] != checked) {
          if (kPollMode == 2 && ++spins > 2000) {  // ~10-20 us of spinning: now sleep until the step is done
            CUDA_TRY(cudaEventSynchronize(p->fit_events[slot]));
            if (hflags[slot * kFitFlagCount + kFitSeq] != checked)
              return fail(FB_ERR_CUDA, "device-driven fit: step %d finished without reporting", checked);
            break;
          }
          if ((++spins & 0x3f) == 0) sched_yield();  // other fits' host threads may share the core
          if ((spins & 0xffff) == 0) {
            const cudaError_t q = cudaStreamQuery(st);
            if (q != cudaSuccess && q != cudaErrorNotReady)
              return fail(FB_ERR_CUDA, "device-driven fit: %s", cudaGetErrorString(q));
            if (q == cudaSuccess && hflags[slot * kFitFlagCount + kFitSeq] != checked)
              return fail(FB_ERR_CUDA, "device-driven fit: the stream drained without step %d reporting", checked);
          }
        }
      } else {
        CUDA_TRY(cudaEventSynchronize(p->fit_events[slot]));
      }
      if (hflags[slot * kFitFlagCount + kFitStop]) break;
      while (enqueued < max_evals && enqueued - (checked + 1) < kAhead) {
        rc = enqueue_one();
        if (rc != FB_OK) return rc;
      }
    }
    us_loop = since(t_begin);
    // everything the host needs, in one go: results, the last evaluated and the accepted matrix,
    // the parameter block (last evaluated point), the flags
    CUDA_TRY(cudaMemcpyAsync(p->h_fit_result, p->d_fit_result, sizeof(fb_fit_result), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(p->h_fit_grams, p->d_gram, sizeof(double) * gcount, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(p->h_fit_grams + gcount, p->d_gram_accepted, sizeof(double) * gcount, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(p->h_fit_params, p->d_params, sizeof(double) * nparam, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(p->h_fit_flags, p->d_fit_flags, sizeof(int) * kFitFlagCount, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    us_readback = since(t_begin) - us_loop;
    p->pc.stop = nullptr;
    gram.assign(p->h_fit_grams, p->h_fit_grams + gcount);
    accepted.assign(p->h_fit_grams + gcount, p->h_fit_grams + 2 * gcount);
    memcpy(p->h_params, p->h_fit_params, sizeof(double) * nparam);
    if (p->h_fit_flags[kFitBail]) {
      // a trial point left the fast kernels' domain: continue on the host from the device state
      CUDA_TRY(cudaMemcpy(stt, p->d_trf, sizeof(TrfState), cudaMemcpyDeviceToHost));
      // bail code 1: S.x_new is proposed but not evaluated; 2: the device solver broke down before
      // proposing (re-propose here). Either way the host needs its own decomposition of J^T J if
      // the hand-over happened inside an outer iteration (the Cholesky step keeps none).
      pending_trial = p->h_fit_flags[kFitBail] == 1;
      if (S.decomposed) {
        for (int i = 0; i < n * n; ++i) S.work[i] = S.A[i];
        jacobi_eigh(S.work, S.V, S.lam, n, n);
        trf_after_decomposition(S);
      }
      need_host_loop = true;
    } else {
      const fb_fit_result& R = *p->h_fit_result;
      if (R.status == -1) {
        result->status = -1;
        return fail(FB_ERR_INVALID, "Residuals are not finite in the initial point.");
      }
      result->status = R.status;
      result->nfev = R.nfev;
      result->njev = R.njev;
      result->cost = R.cost;
      result->optimality = R.optimality;
      for (int i = 0; i < n; ++i) {
        result->x[i] = R.x[i];
        result->grad[i] = R.grad[i];
      }
      // the last evaluation is what the exact Hessian starts from (analysis/fitter.py:105-108)
      p->gram_cache.assign(gram.begin(), gram.end());
      p->gram_cache_params.assign(p->h_params, p->h_params + nparam);
      p->gram_cache_data = data_dev;
      p->gram_cache_valid = true;
      p->vals_match_cache = p->last_eval_split;
      p->tables_dirty = false;  // tables, spectrum rows and parameters all belong to the last evaluation
    }
  } else {
    rc = fb_normal_equations(p, data_dev, gram.data(), stream);
    if (rc != FB_OK) return rc;
  }
  if (need_host_loop) {
    if (!pending_trial) {
      double x0[FB_MAX_FREE];
      pack_free(p, x0);
      trf_init(S, n, m, x0, opt.ftol, opt.xtol, opt.gtol, opt.max_nfev);
      if (!std::isfinite(gram[(size_t)n * ncols + n])) {
        // scipy raises "Residuals are not finite in the initial point." (least_squares.py:945-946)
        result->status = -1;
        return fail(FB_ERR_INVALID, "Residuals are not finite in the initial point.");
      }
      trf_load_accepted(S, gram.data());
      accepted = gram;
      S.nfev = 1;
      S.njev = 1;
    }
    rc = fit_host_loop(p, data_dev, S, gram, accepted, pending_trial, stream);
    if (rc != FB_OK) return rc;
    result->status = S.status;
    result->nfev = S.nfev;
    result->njev = S.njev;
    result->cost = S.cost;
    result->optimality = S.g_norm;
    for (int i = 0; i < n; ++i) {
      result->x[i] = S.x[i];
      result->grad[i] = S.g[i];
    }
  }
  if (gram_final_host) memcpy(gram_final_host, accepted.data(), sizeof(double) * gcount);
  // the model is left at the LAST EVALUATED point, like analysis/fitter.py:258 leaves it.
  const double us_before_hessian = since(t_begin);
  if (opt.compute_hessian || hessian_host) {
    std::vector<double> H((size_t)n * n);
    rc = fb_hessian_impl(p, data_dev, H.data(), st);
    if (rc != FB_OK) return rc;
    if (hessian_host) memcpy(hessian_host, H.data(), sizeof(double) * n * n);
  }
  if (fit_timing)
    fprintf(stderr, "fb_fit: nfev %d  loop %.0f us  readback %.0f us  host tail %.0f us  hessian %.0f us  total %.0f us\n",
            result->nfev, us_loop, us_readback, us_before_hessian - us_loop - us_readback,
            since(t_begin) - us_before_hessian, since(t_begin));
  return FB_OK;
}

// ---- the device-driven fit in steps (channel-sharded multi-GPU fits) ------------------------------
// Same kernels as fb_fit's loop, but the caller owns the matrix buffer and may reduce it over
// the ranks of a process group (NCCL all-reduce on the same stream) between fb_fit_device_evaluate
// and fb_fit_device_step. Nothing waits for the GPU except _poll and _end.
extern "C" int fb_fit_device_begin(fb_plan_t p, const fb_fit_options* options, int64_t num_residuals_total,
                                   void* stream) {
  if (!p) return fail(FB_ERR_INVALID, "null argument");
  if (!p->have_weights) return fail(FB_ERR_INVALID, "fb_set_weights / fb_load_weights must be called first");
  if (p->n_free < 1) return fail(FB_ERR_INVALID, "no free parameters");
  if (p->pc.scintillation)
    return fail(FB_ERR_INVALID, "the device-driven loop does not cover scintillation mode (singular-system check)");
  const fb_fit_options opt = default_options(options);
  int rc = fb_set_parameters(p, p->h_params, stream);
  if (rc != FB_OK) return rc;
  if (!p->fit_session) p->fit_session = new (std::nothrow) SingleTrfArgs();
  if (!p->fit_session) return fail(FB_ERR_NOMEM, "out of host memory");
  const long long m = num_residuals_total > 0 ? (long long)num_residuals_total
                                              : (long long)p->desc.num_freq * p->desc.num_time;
  return device_fit_setup(p, opt, m, *p->fit_session, (cudaStream_t)stream);
}

extern "C" int fb_fit_device_evaluate(fb_plan_t p, const double* data_dev, double* gram_dev, void* stream) {
  if (!p || !data_dev || !gram_dev) return fail(FB_ERR_INVALID, "null argument");
  if (!p->fit_session || p->pc.stop == nullptr) return fail(FB_ERR_INVALID, "fb_fit_device_begin must be called first");
  p->tables_dirty = true;
  return normal_equations_async(p, data_dev, gram_dev, (cudaStream_t)stream);
}

extern "C" int fb_fit_device_step(fb_plan_t p, const double* gram_dev, void* stream) {
  if (!p || !gram_dev) return fail(FB_ERR_INVALID, "null argument");
  if (!p->fit_session || p->pc.stop == nullptr) return fail(FB_ERR_INVALID, "fb_fit_device_begin must be called first");
  SingleTrfArgs T = *p->fit_session;
  T.gram = gram_dev;
  return launch_trf_step(p, T, (cudaStream_t)stream);
}

extern "C" int fb_fit_device_poll(fb_plan_t p, int32_t* stopped_out, int32_t* evaluations_out, void* stream) {
  if (!p || !stopped_out) return fail(FB_ERR_INVALID, "null argument");
  if (!p->fit_session || p->pc.stop == nullptr) return fail(FB_ERR_INVALID, "fb_fit_device_begin must be called first");
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY(cudaMemcpyAsync(p->h_fit_flags, p->d_fit_flags, sizeof(int) * kFitFlagCount, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  *stopped_out = p->h_fit_flags[kFitStop];
  if (evaluations_out) *evaluations_out = p->h_fit_flags[kFitEvals];
  return FB_OK;
}

extern "C" int fb_fit_device_end(fb_plan_t p, fb_fit_result* result, double* gram_accepted_host, int32_t* bailed_out,
                                 void* stream) {
  if (!p || !result) return fail(FB_ERR_INVALID, "null argument");
  if (!p->fit_session || p->pc.stop == nullptr) return fail(FB_ERR_INVALID, "fb_fit_device_begin must be called first");
  cudaStream_t st = (cudaStream_t)stream;
  const int n = p->n_free, ncols = n + 1;
  const size_t gcount = (size_t)ncols * ncols, nparam = (size_t)FB_NUM_PARAMETERS * p->desc.num_components;
  p->pc.stop = nullptr;
  CUDA_TRY(cudaMemcpyAsync(p->h_fit_result, p->d_fit_result, sizeof(fb_fit_result), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(p->h_fit_grams, p->d_gram_accepted, sizeof(double) * gcount, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(p->h_fit_params, p->d_params, sizeof(double) * nparam, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(p->h_fit_flags, p->d_fit_flags, sizeof(int) * kFitFlagCount, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  memcpy(p->h_params, p->h_fit_params, sizeof(double) * nparam);  // the last evaluated point
  p->tables_dirty = true;
  p->gram_cache_valid = false;  // the caller's matrix may be a sum over ranks: never reused for the local Hessian
  if (bailed_out) *bailed_out = p->h_fit_flags[kFitBail];
  memset(result, 0, sizeof(*result));
  result->num_free = n;
  if (p->h_fit_flags[kFitBail]) {
    result->status = -100;
    return FB_OK;
  }
  if (!p->h_fit_flags[kFitFinished]) return fail(FB_ERR_INVALID, "fb_fit_device_end before the fit has finished");
  *result = *p->h_fit_result;
  result->num_free = n;
  if (gram_accepted_host) memcpy(gram_accepted_host, p->h_fit_grams, sizeof(double) * gcount);
  if (result->status == -1) return fail(FB_ERR_INVALID, "Residuals are not finite in the initial point.");
  return FB_OK;
}

extern "C" int fb_profile_eval(const double* times_dev, const double* freqs_dev, int32_t num_rows,
                               int32_t num_cols, double arrival_time, double sc_time_ref, double sc_index,
                               double width, double ref_freq, int32_t scattered, int32_t clamp,
                               int32_t is_folded, int32_t normalize_pbf, double* profile_dev, void* stream) {
  if (!times_dev || !freqs_dev || !profile_dev) return fail(FB_ERR_INVALID, "null argument");
  if (num_rows < 1 || num_cols < 1) return fail(FB_ERR_INVALID, "empty array");
  if (is_folded && num_cols < 2) return fail(FB_ERR_INVALID, "folding needs at least two samples");
  const size_t total = (size_t)num_rows * num_cols;
  k_profile<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      times_dev, freqs_dev, num_rows, num_cols, arrival_time, sc_time_ref, sc_index, width, ref_freq,
      scattered, clamp, is_folded, normalize_pbf, profile_dev);
  CUDA_TRY(cudaGetLastError());
  return FB_OK;
}

extern "C" int fb_measure_fp64_peak(double* tflops_out, void* stream) {
  if (!tflops_out) return fail(FB_ERR_INVALID, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  cudaDeviceProp prop;
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 16;
  double* d_out = nullptr;
  CUDA_TRY(cudaMalloc(&d_out, sizeof(double) * blocks * threads));
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0));
  CUDA_TRY(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    CUDA_TRY(cudaEventRecord(e0, st));
    k_dfma_peak<<<blocks, threads, 0, st>>>(d_out, iters);
    CUDA_TRY(cudaEventRecord(e1, st));
    CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 2.0 * 8.0 * (double)iters * blocks * threads;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d_out);
  *tflops_out = best;
  return FB_OK;
}

// ---- pre-fit data preparation (fb_preprocess.cuh) ---------------------------------------------------
static int pre_check(const void* a, int nf, int nt) {
  if (!a) return fail(FB_ERR_INVALID, "null argument");
  if (nf < 1 || nt < 1) return fail(FB_ERR_INVALID, "empty spectrum");
  return FB_OK;
}

extern "C" int fb_pre_channel_moments(const double* data_dev, const double* weights_dev, int32_t nf, int32_t nt,
                                      double* moments_dev, void* stream) {
  if (!weights_dev || !moments_dev) return fail(FB_ERR_INVALID, "null argument");
  int rc = pre_check(data_dev, nf, nt);
  if (rc != FB_OK) return rc;
  k_pre_moments<<<nf, kPreBlock, 0, (cudaStream_t)stream>>>(data_dev, weights_dev, nt, moments_dev);
  CUDA_TRY(cudaGetLastError());
  return FB_OK;
}

extern "C" int fb_pre_remove_baseline(double* data_dev, const double* weights_dev, const uint8_t* good_dev,
                                      const double* mean_dev, int32_t nf, int32_t nt, void* stream) {
  if (!weights_dev || !good_dev || !mean_dev) return fail(FB_ERR_INVALID, "null argument");
  int rc = pre_check(data_dev, nf, nt);
  if (rc != FB_OK) return rc;
  k_pre_baseline<<<nf, kPreBlock, 0, (cudaStream_t)stream>>>(data_dev, weights_dev, good_dev, mean_dev, nt);
  CUDA_TRY(cudaGetLastError());
  return FB_OK;
}

extern "C" int fb_pre_power_sums(const double* data_dev, int32_t nf, int32_t nt, double* sums_dev, void* stream) {
  if (!sums_dev) return fail(FB_ERR_INVALID, "null argument");
  int rc = pre_check(data_dev, nf, nt);
  if (rc != FB_OK) return rc;
  k_pre_power_sums<<<nf, kPreBlock, 0, (cudaStream_t)stream>>>(data_dev, nt, sums_dev);
  CUDA_TRY(cudaGetLastError());
  return FB_OK;
}

extern "C" int fb_pre_zero_channels(double* data_dev, const uint8_t* keep_dev, int32_t nf, int32_t nt, void* stream) {
  if (!keep_dev) return fail(FB_ERR_INVALID, "null argument");
  int rc = pre_check(data_dev, nf, nt);
  if (rc != FB_OK) return rc;
  k_pre_zero_channels<<<nf, kPreBlock, 0, (cudaStream_t)stream>>>(data_dev, keep_dev, nt);
  CUDA_TRY(cudaGetLastError());
  return FB_OK;
}

extern "C" int fb_pre_downsample_2d(const double* in_dev, int32_t nf, int32_t nt, int32_t ff, int32_t ft,
                                    double* out_dev, void* stream) {
  if (!out_dev) return fail(FB_ERR_INVALID, "null argument");
  int rc = pre_check(in_dev, nf, nt);
  if (rc != FB_OK) return rc;
  if (ff < 1 || ft < 1 || nf % ff != 0 || nt % ft != 0)
    return fail(FB_ERR_INVALID, "cannot reshape array of size %d x %d by factors %d x %d", nf, nt, ff, ft);
  const int nfo = nf / ff, nto = nt / ft;
  const dim3 grid((nto + kPreBlock - 1) / kPreBlock, nfo);
  k_pre_downsample<<<grid, kPreBlock, 0, (cudaStream_t)stream>>>(in_dev, nt, ff, ft, nto, out_dev);
  CUDA_TRY(cudaGetLastError());
  return FB_OK;
}

extern "C" int fb_pre_window(const double* in_dev, int32_t nf, int32_t nt, const int32_t* start_dev, int32_t width,
                             double* out_dev, void* stream) {
  if (!start_dev || !out_dev) return fail(FB_ERR_INVALID, "null argument");
  int rc = pre_check(in_dev, nf, nt);
  if (rc != FB_OK) return rc;
  if (width < 1 || width > nt) return fail(FB_ERR_INVALID, "window width %d out of range", width);
  k_pre_window<<<nf, kPreBlock, 0, (cudaStream_t)stream>>>(in_dev, nt, start_dev, width, out_dev);
  CUDA_TRY(cudaGetLastError());
  return FB_OK;
}

extern "C" int fb_pre_time_series(const double* in_dev, int32_t nf, int32_t nt, double* scratch_dev, double* out_dev,
                                  void* stream) {
  if (!scratch_dev || !out_dev) return fail(FB_ERR_INVALID, "null argument");
  int rc = pre_check(in_dev, nf, nt);
  if (rc != FB_OK) return rc;
  const int chunks = (nf + kPreRowChunk - 1) / kPreRowChunk, tiles = (nt + kPreBlock - 1) / kPreBlock;
  k_pre_colsum_partial<<<dim3(tiles, chunks), kPreBlock, 0, (cudaStream_t)stream>>>(in_dev, nf, nt, scratch_dev);
  CUDA_TRY(cudaGetLastError());
  k_pre_colmean_final<<<tiles, kPreBlock, 0, (cudaStream_t)stream>>>(scratch_dev, chunks, nf, nt, out_dev);
  CUDA_TRY(cudaGetLastError());
  return FB_OK;
}

extern "C" int fb_upload_rows(const void* host_pinned, int32_t elem_size, const int32_t* rows_dev, int32_t num_rows,
                              int32_t nf, int32_t nt, double* dst_dev, void* stream) {
  if (!host_pinned || !dst_dev || (num_rows > 0 && !rows_dev)) return fail(FB_ERR_INVALID, "null argument");
  if (elem_size != 4 && elem_size != 8) return fail(FB_ERR_INVALID, "element size must be 4 (float32) or 8 (float64)");
  if (nf < 1 || nt < 1 || num_rows < 0 || num_rows > nf) return fail(FB_ERR_INVALID, "bad shape");
  if (num_rows == 0) return FB_OK;
  cudaPointerAttributes attr;
  CUDA_TRY(cudaPointerGetAttributes(&attr, host_pinned));
  if (attr.type != cudaMemoryTypeHost || !attr.devicePointer)
    return fail(FB_ERR_INVALID, "fb_upload_rows needs page-locked host memory");
  static int sm_count = 0;
  if (sm_count == 0) {
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
  }
  // a modest grid: the copy is bound by the host link, and the SMs are shared with the fits of
  // the other streams
  const int grid = std::min(num_rows, 2 * sm_count);
  if (elem_size == 8)
    k_upload_rows<double><<<grid, 256, 0, (cudaStream_t)stream>>>(static_cast<const double*>(attr.devicePointer),
                                                                  rows_dev, num_rows, nt, dst_dev);
  else
    k_upload_rows<float><<<grid, 256, 0, (cudaStream_t)stream>>>(static_cast<const float*>(attr.devicePointer),
                                                                 rows_dev, num_rows, nt, dst_dev);
  CUDA_TRY(cudaGetLastError());
  return FB_OK;
}

// ---- small public routines (routines/spectrum.py, routines/ism.py) ----------------------------------
static unsigned elementwise_grid(long long count) {
  return (unsigned)std::max<long long>(1, std::min<long long>((count + 255) / 256, 148LL * 16));
}

extern "C" int fb_spectrum_rpl(const double* freqs_dev, int64_t count, double freq_ref, double sp_idx, double sp_run,
                               double* out_dev, void* stream) {
  if (!freqs_dev || !out_dev) return fail(FB_ERR_INVALID, "null argument");
  if (count < 1) return fail(FB_ERR_INVALID, "empty array");
  k_spectrum_rpl<<<elementwise_grid(count), 256, 0, (cudaStream_t)stream>>>(freqs_dev, count, freq_ref, sp_idx, sp_run, out_dev);
  CUDA_TRY(cudaGetLastError());
  return FB_OK;
}

extern "C" int fb_ism_times(int32_t kind, const double* freqs_dev, int64_t count, double a, double b, double c, double d,
                            double* out_dev, void* stream) {
  if (!freqs_dev || !out_dev) return fail(FB_ERR_INVALID, "null argument");
  if (kind < 0 || kind > 2) return fail(FB_ERR_INVALID, "kind must be 0 (dm delay), 1 (dm smear) or 2 (scattering time)");
  if (count < 1) return fail(FB_ERR_INVALID, "empty array");
  k_ism_times<<<elementwise_grid(count), 256, 0, (cudaStream_t)stream>>>(kind, freqs_dev, count, a, b, c, d, out_dev);
  CUDA_TRY(cudaGetLastError());
  return FB_OK;
}

extern "C" int fb_amplitude_per_channel(const double* data_dev, const double* model_dev, int32_t num_time, int32_t num_comp,
                                        double* amplitudes_dev, void* stream) {
  if (!data_dev || !model_dev || !amplitudes_dev) return fail(FB_ERR_INVALID, "null argument");
  if (num_time < 1 || num_comp < 1 || num_comp > FB_MAX_COMPONENTS)
    return fail(FB_ERR_INVALID, "need num_time >= 1 and 1 <= num_components <= %d", FB_MAX_COMPONENTS);
  cudaStream_t st = (cudaStream_t)stream;
  int* d_flag = nullptr;
  CUDA_TRY(cudaMalloc(&d_flag, sizeof(int)));
  cudaMemsetAsync(d_flag, 0, sizeof(int), st);
  const size_t smem = sizeof(double) * (256 + (size_t)num_comp * (num_comp + 1));
  k_amplitude_per_channel<<<1, 256, smem, st>>>(data_dev, model_dev, num_time, num_comp, amplitudes_dev, d_flag);
  int flag = 0;
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d_flag);
  if (e != cudaSuccess) return fail(FB_ERR_CUDA, "fb_amplitude_per_channel: %s", cudaGetErrorString(e));
  if (flag) return fail(FB_ERR_SINGULAR, "Singular matrix");  // numpy.linalg.LinAlgError text, routines/ism.py:46
  return FB_OK;
}

extern "C" int fb_widen_f32(const float* src_dev, int64_t count, double* dst_dev, void* stream) {
  if (!src_dev || !dst_dev) return fail(FB_ERR_INVALID, "null argument");
  if (count < 0) return fail(FB_ERR_INVALID, "negative count");
  if (count == 0) return FB_OK;
  if ((reinterpret_cast<uintptr_t>(src_dev) & 15) || (reinterpret_cast<uintptr_t>(dst_dev) & 15))
    return fail(FB_ERR_INVALID, "fb_widen_f32 needs 16-byte aligned buffers");
  k_widen_f32<<<elementwise_grid((count + 3) / 4), 256, 0, (cudaStream_t)stream>>>(src_dev, count, dst_dev);
  CUDA_TRY(cudaGetLastError());
  return FB_OK;
}

extern "C" int fb_launch_count(fb_plan_t p, int64_t* count_out, int32_t reset) {
  if (!p || !count_out) return fail(FB_ERR_INVALID, "null argument");
  *count_out = p->launches;
  if (reset) p->launches = 0;
  return FB_OK;
}


// ---- exact Hessian (analysis/fitter.py:81-178) --------------------------------------------------
template <bool MAP>
static int launch_hessian(fb_plan_s* p, HessArgs& H, cudaStream_t st, int* grid_out) {
  {
    const int rc_t = ensure_tables(p, st, H.E.chan_list != nullptr && H.E.chan_list == p->d_chan_list);
    if (rc_t != FB_OK) return rc_t;
  }
  {
    const int rc_s = presolve_scint_amps(p, H.E, st);
    if (rc_s != FB_OK) return rc_s;
  }
  const size_t smem = hess_smem_bytes(p->pc.num_comp, p->pc.kf, MAP ? 0 : H.num_pairs);
  if (smem > 227 * 1024) return fail(FB_ERR_INVALID, "Hessian shared-memory request %zu B exceeds 227 KB", smem);
  if (smem > 48 * 1024)
    CUDA_TRY(cudaFuncSetAttribute(k_hessian<MAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 1;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_hessian<MAP>, kBlock, smem));
  if (occ < 1) occ = 1;
  const int tiles = (p->pc.num_time + kBlock - 1) / kBlock;
  long long items = (p->pc.scintillation && !H.E.scint_amps) ? (long long)H.E.num_chan : (long long)H.E.num_chan * tiles;
  if (items < 1) items = 1;
  long long grid = (long long)p->sm_count * occ;
  if (MAP) grid *= 8;
  if (grid > items) grid = items;
  if (!MAP) {
    const size_t need = (size_t)grid * H.num_pairs;
    if (need > p->hess_partials_cap) {
      cudaFree(p->d_hess_partials);
      p->d_hess_partials = nullptr;
      p->hess_partials_cap = 0;
      CUDA_TRY(cudaMalloc(&p->d_hess_partials, sizeof(double) * need));
      p->hess_partials_cap = need;
    }
    H.partials = p->d_hess_partials;
  }
  k_hessian<MAP><<<(int)grid, kBlock, smem, st>>>(H);
  CUDA_TRY(cudaGetLastError());
  p->launches += 1;
  if (grid_out) *grid_out = (int)grid;
  return FB_OK;
}

// Pair list (analysis/fitter.py:132-176) and term descriptors of the exact Hessian for the plan's
// current free-parameter layout: built and uploaded once per layout.
static int ensure_hessian_layout(fb_plan_s* p, cudaStream_t st) {
  const int n = p->n_free, nc = p->desc.num_components;
  // packed columns -> (parameter, component), analysis/fitter.py:116-127
  std::vector<int> cp(n), cc(n);
  for (int k = 0; k < 9; ++k) {
    if (p->col_of[k] < 0) continue;
    if (is_fit_global(k)) {
      cp[p->col_of[k]] = k;
      cc[p->col_of[k]] = -1;
    } else {
      for (int c = 0; c < nc; ++c) {
        cp[p->col_of[k] + c] = k;
        cc[p->col_of[k] + c] = c;
      }
    }
  }
  // pair list of the current free-parameter layout: built and uploaded once per layout
  std::vector<HessPair>& pairs = p->hess_pairs;
  std::vector<int>&pair_a = p->hess_pair_a, &pair_b = p->hess_pair_b, &comp_list = p->hess_comp_list;
  int* comp_off = p->hess_comp_off;
  if (!p->hess_layout_valid) {
    pairs.clear();
    pair_a.clear();
    pair_b.clear();
    comp_list.clear();
    for (int a = 0; a < n; ++a)
      for (int b = a; b < n; ++b) {
        const bool la = cc[a] >= 0, lb = cc[b] >= 0;
        if (la && lb && cc[a] != cc[b]) continue;  // mixed partial skipped, analysis/fitter.py:158-160
        HessPair hp;
        hp.pa = (int8_t)cp[a];
        hp.pb = (int8_t)cp[b];
        hp.component = (int8_t)(lb ? cc[b] : (la ? cc[a] : 0));  // analysis/fitter.py:163-168
        hp.is_sum = d2_is_sum(cp[a], cp[b]) ? 1 : 0;
        pairs.push_back(hp);
        pair_a.push_back(a);
        pair_b.push_back(b);
      }
    const int np0 = (int)pairs.size();
    if ((size_t)np0 > p->pairs_cap) {
      cudaFree(p->d_pairs);
      p->d_pairs = nullptr;
      p->pairs_cap = 0;
      CUDA_TRY(cudaMalloc(&p->d_pairs, sizeof(HessPair) * np0));
      p->pairs_cap = np0;
    }
    CUDA_TRY(cudaMemcpyAsync(p->d_pairs, pairs.data(), sizeof(HessPair) * np0, cudaMemcpyHostToDevice, st));
    // per component: the pairs it contributes to (summing functions: all components; others: the
    // one `component` the reference passes)
    for (int c = 0; c < nc; ++c) {
      comp_off[c] = (int)comp_list.size();
      for (int k = 0; k < np0; ++k)
        if (pairs[k].is_sum || pairs[k].component == c) comp_list.push_back(k);
    }
    comp_off[nc] = (int)comp_list.size();
    if (comp_list.size() > p->comp_list_cap) {
      cudaFree(p->d_comp_list);
      p->d_comp_list = nullptr;
      p->comp_list_cap = 0;
      CUDA_TRY(cudaMalloc(&p->d_comp_list, sizeof(int) * comp_list.size()));
      p->comp_list_cap = comp_list.size();
    }
    CUDA_TRY(cudaMemcpyAsync(p->d_comp_list, comp_list.data(), sizeof(int) * comp_list.size(),
                             cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    p->hess_layout_valid = true;
    p->hess_desc_valid = false;
  }
  const int npairs = (int)pairs.size();
  if (!p->hess_desc_valid) {
    std::vector<HessDesc>& desc = p->hess_desc;
    desc.assign(npairs, HessDesc());
    for (int c = 0; c < nc; ++c) p->hess_need[c] = 0u;
    p->hess_need_sum_dmidx = 0;
    for (int k = 0; k < npairs; ++k) {
      int code, idx;
      pair_descriptor(pairs[k].pa, pairs[k].pb, code, idx);
      desc[k].code = (int8_t)code;
      desc[k].idx = (int8_t)idx;
      for (int c = 0; c < nc; ++c)
        if (pairs[k].is_sum || pairs[k].component == c) p->hess_need[c] |= (1u << idx);
      if (idx == 12 || idx == 21) p->hess_need_sum_dmidx = 1;
    }
    if ((size_t)npairs > p->desc_cap) {
      cudaFree(p->d_desc);
      p->d_desc = nullptr;
      p->desc_cap = 0;
      CUDA_TRY(cudaMalloc(&p->d_desc, sizeof(HessDesc) * (npairs > 0 ? npairs : 1)));
      p->desc_cap = npairs;
    }
    CUDA_TRY(cudaMemcpyAsync(p->d_desc, desc.data(), sizeof(HessDesc) * npairs, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    p->hess_desc_valid = true;
  }
  return FB_OK;
}

int fb_hessian_impl(fb_plan_s* p, const double* data_dev, double* hessian_host, cudaStream_t st) {
  if (!p->have_weights) return fail(FB_ERR_INVALID, "fb_set_weights / fb_load_weights must be called first");
  const int n = p->n_free, nc = p->desc.num_components;
  if (n < 1) return fail(FB_ERR_INVALID, "no free parameters");
  int rc_layout = ensure_hessian_layout(p, st);
  if (rc_layout != FB_OK) return rc_layout;
  const std::vector<HessPair>& pairs = p->hess_pairs;
  const std::vector<int>&pair_a = p->hess_pair_a, &pair_b = p->hess_pair_b;
  const int* comp_off = p->hess_comp_off;
  const int npairs = (int)pairs.size();
  const int ncols = n + 1;
  std::vector<double> gram((size_t)ncols * ncols), S(npairs, 0.0);
  // first-derivative products = J^T J of the weighted Jacobian at the model's current state;
  // after a fit that is normally the last trust-region evaluation, still cached.
  int rc = FB_OK;
  const size_t nparam = (size_t)FB_NUM_PARAMETERS * nc;
  bool reuse_vals = false;
  if (p->gram_cache_valid && p->gram_cache_data == data_dev && p->gram_cache.size() == (size_t)ncols * ncols &&
      p->gram_cache_params.size() == nparam &&
      memcmp(p->gram_cache_params.data(), p->h_params, sizeof(double) * nparam) == 0) {
    gram = p->gram_cache;
    reuse_vals = p->vals_match_cache;
  } else {
    rc = normal_equations_async(p, data_dev, p->d_gram, st);
    if (rc != FB_OK) return rc;
    CUDA_TRY(cudaMemcpyAsync(gram.data(), p->d_gram, sizeof(double) * ncols * ncols, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    reuse_vals = p->last_eval_split;  // k_vals just ran at these parameters
  }
  if (p->num_good > 0 && !p->pc.scintillation) {
    // fast path: register accumulation of the 30 term values per (sample, component)
    HessAccArgs H;
    memset(&H, 0, sizeof(H));
    for (int c = 0; c < nc; ++c) H.need[c] = p->hess_need[c];
    H.need_sum_dmidx = p->hess_need_sum_dmidx;
    H.E = base_args(p);
    H.E.data = data_dev;
    H.E.weights = p->d_weights;
    H.E.chan_list = p->d_chan_list;
    H.E.num_chan = p->num_good;
    H.desc = p->d_desc;
    H.num_pairs = npairs;
    H.comp_list = p->d_comp_list;
    H.gvals = reuse_vals ? p->d_vals : nullptr;
    H.gvals_layout = p->vals_layout;
    for (int c = 0; c <= nc; ++c) H.comp_off[c] = comp_off[c];
    rc = ensure_tables(p, st, true);
    if (rc != FB_OK) return rc;
    if (fast_path_ok(p, false)) {
      // warp-granular contraction over the spectrum rows of k_vals3
      if (!(reuse_vals && p->vals_layout == 3)) {
        rc = ensure_capacity(&p->d_vals, &p->vals_cap, (size_t)p->num_good * nc * p->pc.num_time);
        if (rc == FB_OK) rc = launch_vals3(p, H.E, p->d_vals, st, nullptr);
        if (rc != FB_OK) return rc;
        p->vals_layout = 3;
        p->vals_match_cache = false;
      }
      H.gvals = p->d_vals;
      H.gvals_layout = 3;
      const size_t smem3 = hess3_smem_bytes(nc, npairs);
      if (smem3 > 227 * 1024) return fail(FB_ERR_INVALID, "Hessian shared-memory request %zu B exceeds 227 KB", smem3);
      auto kern = nc == 1 ? k_hess3<1> : nc == 2 ? k_hess3<2> : nc == 3 ? k_hess3<3> : k_hess3<4>;
      int& occ3 = p->occ_cache[40 + nc];
      if (occ3 == 0 || smem3 > 48 * 1024) {
        if (smem3 > 48 * 1024)
          CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ3, kern, kFastBlock, smem3));
        if (occ3 < 1) occ3 = 1;
      }
      const int grid3 = std::min(p->num_good, p->sm_count * occ3);
      const size_t need3 = (size_t)grid3 * npairs;
      if (need3 > p->hess_partials_cap) {
        cudaFree(p->d_hess_partials);
        p->d_hess_partials = nullptr;
        p->hess_partials_cap = 0;
        CUDA_TRY(cudaMalloc(&p->d_hess_partials, sizeof(double) * need3));
        p->hess_partials_cap = need3;
      }
      H.partials = p->d_hess_partials;
      kern<<<grid3, kFastBlock, smem3, st>>>(H);
      CUDA_TRY(cudaGetLastError());
      p->launches += 1;
      double* d_out3 = p->d_gram;  // reuse: npairs <= (n+1)^2
      k_hess_finalize<<<(npairs + 3) / 4, 128, 0, st>>>(p->d_hess_partials, grid3, npairs, d_out3);
      CUDA_TRY(cudaGetLastError());
      p->launches += 1;
      CUDA_TRY(cudaMemcpyAsync(S.data(), d_out3, sizeof(double) * npairs, cudaMemcpyDeviceToHost, st));
    } else {
    const size_t smem = hess_acc_smem_bytes(nc, p->pc.kf, npairs);
    if (smem > 227 * 1024) return fail(FB_ERR_INVALID, "Hessian shared-memory request %zu B exceeds 227 KB", smem);
    if (smem > 48 * 1024)
      CUDA_TRY(cudaFuncSetAttribute(k_hessian_acc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_hessian_acc, kBlock, smem));
    if (occ < 1) occ = 1;
    const int tiles = (p->pc.num_time + kBlock - 1) / kBlock;
    long long items = (long long)p->num_good * ((tiles + kHessSuper - 1) / kHessSuper);
    long long grid = (long long)p->sm_count * occ;
    if (grid > items) grid = items;
    const size_t need_cap = (size_t)grid * npairs;
    if (need_cap > p->hess_partials_cap) {
      cudaFree(p->d_hess_partials);
      p->d_hess_partials = nullptr;
      p->hess_partials_cap = 0;
      CUDA_TRY(cudaMalloc(&p->d_hess_partials, sizeof(double) * need_cap));
      p->hess_partials_cap = need_cap;
    }
    H.partials = p->d_hess_partials;
    k_hessian_acc<<<(int)grid, kBlock, smem, st>>>(H);
    CUDA_TRY(cudaGetLastError());
    p->launches += 1;
    double* d_out = p->d_gram;  // reuse: npairs <= (n+1)^2
    CUDA_TRY(cudaStreamSynchronize(st));
    k_hess_finalize<<<(npairs + 3) / 4, 128, 0, st>>>(p->d_hess_partials, (int)grid, npairs, d_out);
    CUDA_TRY(cudaGetLastError());
    p->launches += 1;
    CUDA_TRY(cudaMemcpyAsync(S.data(), d_out, sizeof(double) * npairs, cudaMemcpyDeviceToHost, st));
    }
  } else if (p->num_good > 0) {
    HessArgs H;
    memset(&H, 0, sizeof(H));
    H.E = base_args(p);
    H.E.data = data_dev;
    H.E.weights = p->d_weights;
    H.E.chan_list = p->d_chan_list;
    H.E.num_chan = p->num_good;
    H.pairs = p->d_pairs;
    H.num_pairs = npairs;
    H.comp_list = p->d_comp_list;
    for (int c = 0; c <= nc; ++c) H.comp_off[c] = comp_off[c];
    int grid = 0;
    rc = launch_hessian<false>(p, H, st, &grid);
    if (rc != FB_OK) return rc;
    double* d_out = p->d_gram;  // reuse: npairs <= (n+1)^2
    CUDA_TRY(cudaStreamSynchronize(st));
    k_hess_finalize<<<(npairs + 3) / 4, 128, 0, st>>>(p->d_hess_partials, grid, npairs, d_out);
    CUDA_TRY(cudaGetLastError());
    p->launches += 1;
    CUDA_TRY(cudaMemcpyAsync(S.data(), d_out, sizeof(double) * npairs, cudaMemcpyDeviceToHost, st));
  }
  CUDA_TRY(cudaStreamSynchronize(st));
  for (int a = 0; a < n; ++a)
    for (int b = 0; b < n; ++b) hessian_host[a * n + b] = 2.0 * gram[(size_t)a * ncols + b];
  for (int k = 0; k < npairs; ++k) {
    const int a = pair_a[k], b = pair_b[k];
    hessian_host[a * n + b] -= 2.0 * S[k];
    if (a != b) hessian_host[b * n + a] -= 2.0 * S[k];
  }
  return check_err_flag(p, st);
}

extern "C" int fb_hessian(fb_plan_t p, const double* data_dev, double* hessian_host, void* stream) {
  if (!p || !data_dev || !hessian_host) return fail(FB_ERR_INVALID, "null argument");
  return fb_hessian_impl(p, data_dev, hessian_host, (cudaStream_t)stream);
}

extern "C" int fb_derivative2_map(fb_plan_t p, const double* data_dev, int32_t param1, int32_t param2,
                                  int32_t component, double* out_dev, void* stream) {
  if (!p || !out_dev) return fail(FB_ERR_INVALID, "null argument");
  if (param1 < 0 || param1 > 8 || param2 < 0 || param2 > 8) return fail(FB_ERR_INVALID, "parameter index out of range");
  if (component < 0 || component >= p->desc.num_components) return fail(FB_ERR_INVALID, "component out of range");
  if (p->pc.scintillation && !data_dev) return fail(FB_ERR_INVALID, "scintillation mode needs data");
  HessArgs H;
  memset(&H, 0, sizeof(H));
  H.E = base_args(p);
  H.E.data = data_dev;
  H.map_pa = param1;
  H.map_pb = param2;
  H.map_component = component;
  H.map_out = out_dev;
  int rc = launch_hessian<true>(p, H, (cudaStream_t)stream, nullptr);
  if (rc != FB_OK) return rc;
  return check_err_flag(p, (cudaStream_t)stream);
}

// ---- entry points implemented in later translation units / sections ---------------------------
#ifndef FB_HAVE_HESSIAN
int fb_hessian_impl(fb_plan_s*, const double*, double*, cudaStream_t) {
  return fail(FB_ERR_INVALID, "fb_hessian: not built");
}
extern "C" int fb_hessian(fb_plan_t, const double*, double*, void*) {
  return fail(FB_ERR_INVALID, "fb_hessian: not built");
}
extern "C" int fb_derivative2_map(fb_plan_t, const double*, int32_t, int32_t, int32_t, double*, void*) {
  return fail(FB_ERR_INVALID, "fb_derivative2_map: not built");
}
#endif
// ---- trust-region state machine (see fb_trf.h) ---------------------------------------------------
struct fb_trf_s {
  TrfState S;
  std::vector<double> accepted;
  bool started = false;
};

static fb_fit_options default_options(const fb_fit_options* options) {
  fb_fit_options opt;
  opt.ftol = opt.xtol = opt.gtol = 1e-8;
  opt.max_nfev = 0;
  opt.compute_hessian = 0;
  if (options) opt = *options;
  return opt;
}

extern "C" int fb_trf_create(int32_t num_free, int64_t num_residuals, const double* x0_host,
                             const fb_fit_options* options, fb_trf_t* trf_out) {
  if (!x0_host || !trf_out) return fail(FB_ERR_INVALID, "null argument");
  if (num_free < 1 || num_free > FB_MAX_FREE) return fail(FB_ERR_INVALID, "num_free must be in 1..%d", FB_MAX_FREE);
  fb_trf_s* t = new (std::nothrow) fb_trf_s();
  if (!t) return fail(FB_ERR_NOMEM, "out of host memory");
  const fb_fit_options opt = default_options(options);
  trf_init(t->S, num_free, num_residuals, x0_host, opt.ftol, opt.xtol, opt.gtol, opt.max_nfev);
  t->accepted.assign((size_t)(num_free + 1) * (num_free + 1), 0.0);
  *trf_out = t;
  return FB_OK;
}

extern "C" int fb_trf_destroy(fb_trf_t t) {
  delete t;
  return FB_OK;
}

extern "C" int fb_trf_start(fb_trf_t t, const double* gram_host) {
  if (!t || !gram_host) return fail(FB_ERR_INVALID, "null argument");
  const int n = t->S.n, ncols = n + 1;
  if (!std::isfinite(gram_host[(size_t)n * ncols + n]))
    return fail(FB_ERR_INVALID, "Residuals are not finite in the initial point.");  // least_squares.py:945-946
  trf_load_accepted(t->S, gram_host);
  t->accepted.assign(gram_host, gram_host + (size_t)ncols * ncols);
  t->S.nfev = 1;
  t->S.njev = 1;
  t->started = true;
  return FB_OK;
}

extern "C" int fb_trf_propose(fb_trf_t t, double* x_trial_host) {
  if (!t || !x_trial_host || !t->started) {
    fail(FB_ERR_INVALID, "fb_trf_propose before fb_trf_start");
    return FB_ERR_INVALID;
  }
  if (!trf_propose(t->S)) return 0;
  for (int i = 0; i < t->S.n; ++i) x_trial_host[i] = t->S.x_new[i];
  return 1;
}

extern "C" int fb_trf_update(fb_trf_t t, const double* gram_host) {
  if (!t || !gram_host || !t->started) return fail(FB_ERR_INVALID, "fb_trf_update before fb_trf_start");
  const int njev_before = t->S.njev, ncols = t->S.n + 1;
  trf_update(t->S, gram_host);
  if (t->S.njev != njev_before) t->accepted.assign(gram_host, gram_host + (size_t)ncols * ncols);
  return FB_OK;
}

extern "C" int fb_trf_result(fb_trf_t t, fb_fit_result* result, double* gram_accepted_host) {
  if (!t || !result) return fail(FB_ERR_INVALID, "null argument");
  const TrfState& S = t->S;
  memset(result, 0, sizeof(*result));
  result->num_free = S.n;
  result->status = S.finished ? S.status : -100;
  result->nfev = S.nfev;
  result->njev = S.njev;
  result->cost = S.cost;
  result->optimality = S.g_norm;
  for (int i = 0; i < S.n; ++i) {
    result->x[i] = S.x[i];
    result->grad[i] = S.g[i];
  }
  if (gram_accepted_host) memcpy(gram_accepted_host, t->accepted.data(), sizeof(double) * t->accepted.size());
  return FB_OK;
}

// ---- chi^2 grid sweep (BASELINE.json config 4) ----------------------------------------------------
extern "C" int fb_chisq_grid(fb_plan_t p, const double* data_dev, const double* dm_values_host, int32_t num_dm,
                             const double* sc_values_host, int32_t num_sc, double* chisq_dev, void* stream) {
  if (!p || !data_dev || !dm_values_host || !sc_values_host || !chisq_dev) return fail(FB_ERR_INVALID, "null argument");
  if (!p->have_weights) return fail(FB_ERR_INVALID, "fb_set_weights / fb_load_weights must be called first");
  if (num_dm < 1 || num_sc < 1) return fail(FB_ERR_INVALID, "empty grid");
  cudaStream_t st = (cudaStream_t)stream;
  const int nc = p->desc.num_components;
  if (p->num_good == 0) {
    CUDA_TRY(cudaMemsetAsync(chisq_dev, 0, sizeof(double) * (size_t)num_dm * num_sc, st));
    return FB_OK;
  }
  // fast kernels (fb_fast.cuh) whenever the grid point is on the scattered branch
  bool fast_grid = !env_off("FITBURST_B200_FAST") && !p->pc.scintillation && !p->pc.is_folded && nc <= 4 &&
                   vals3_smem_bytes(nc, p->pc.kf) <= 200 * 1024;
  for (int c = 0; c < nc; ++c)
    fast_grid = fast_grid && p->h_params[FB_BURST_WIDTH * nc + c] > 0.0 && p->h_params[FB_REF_FREQ * nc + c] > 0.0;
  // FITBURST_B200_GRID_FUSED=0: spectrum values stored and reduced by a second kernel (the
  // round-1 sweep; kept for comparison)
  const bool fused_grid = fast_grid && !env_off("FITBURST_B200_GRID_FUSED") &&
                          vals3_smem_bytes(nc, p->pc.kf, true) <= 200 * 1024;
  for (int a = 0; a < num_dm; ++a)
    for (int b = 0; b < num_sc; ++b) {
      // the whole sweep is enqueued without a host round trip: the grid point is written into
      // the device parameter block by a kernel, tables and chi^2 follow on the same stream.
      k_set_grid_point<<<1, 32, 0, st>>>(p->d_params, nc, dm_values_host[a], sc_values_host[b]);
      p->launches += 1;
      p->tables_dirty = true;
      EvalArgs A = base_args(p);
      A.data = data_dev;
      A.weights = p->d_weights;
      A.chan_list = p->d_chan_list;
      A.num_chan = p->num_good;
      A.n_free = 0;
      int grid = 0;
      // per-block partials live in the Gram partial buffer
      const size_t need = (size_t)p->sm_count * 256;
      if (need > p->partials_cap) {
        cudaFree(p->d_partials);
        p->d_partials = nullptr;
        p->partials_cap = 0;
        CUDA_TRY(cudaMalloc(&p->d_partials, sizeof(double) * need));
        p->partials_cap = need;
      }
      A.chi2_partials = p->d_partials;
      int rc = FB_OK;
      if (p->pc.scintillation) {
        rc = launch_eval<MODE_CHI2>(p, A, 1, st, &grid);
      } else if (fast_grid && sc_values_host[b] > 0.0) {
        rc = ensure_tables(p, st, true);
        if (rc != FB_OK) return rc;
        if (fused_grid) {
          // one kernel: spectrum values and residual sums, nothing stored in between
          rc = ensure_capacity(&p->d_partials, &p->partials_cap, (size_t)p->sm_count * 16 * kFastWarps);
          if (rc != FB_OK) return rc;
          A.chi2_partials = p->d_partials;
          rc = launch_chisq3(p, A, st, &grid);
          if (rc != FB_OK) return rc;
          grid *= kFastWarps;
        } else {
          rc = ensure_capacity(&p->d_vals, &p->vals_cap, (size_t)p->num_good * nc * p->pc.num_time);
          if (rc == FB_OK) rc = ensure_capacity(&p->d_partials, &p->partials_cap, (size_t)p->num_good);
          if (rc != FB_OK) return rc;
          A.chi2_partials = p->d_partials;
          rc = launch_vals3(p, A, p->d_vals, st, nullptr);
          if (rc != FB_OK) return rc;
          k_chi2_vals<<<p->num_good, 256, 0, st>>>(A, p->d_vals);
          CUDA_TRY(cudaGetLastError());
          p->launches += 1;
          grid = p->num_good;
          p->vals_match_cache = false;
        }
      } else {
        rc = ensure_tables(p, st, true);
        if (rc != FB_OK) return rc;
        const size_t smem = smem_bytes(nc, p->pc.kf, 1, false, false, SM_VALS | SM_QUEUE);
        if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(k_chi2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int occ = 1;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_chi2, kBlock, smem));
        const long long items = (long long)p->num_good * ((p->pc.num_time + kBlock - 1) / kBlock);
        grid = (int)std::min<long long>(items, (long long)p->sm_count * std::max(1, occ) * 4);
        if ((size_t)grid > p->partials_cap) return fail(FB_ERR_INVALID, "chi^2 partial buffer too small");
        k_chi2<<<grid, kBlock, smem, st>>>(A);
        CUDA_TRY(cudaGetLastError());
        p->launches += 1;
      }
      if (rc != FB_OK) return rc;
      k_chi2_finalize<<<1, 256, 0, st>>>(p->d_partials, grid, chisq_dev + (size_t)a * num_sc + b);
      CUDA_TRY(cudaGetLastError());
      p->launches += 1;
    }
  // restore the plan's own parameters on the device
  const size_t n = (size_t)FB_NUM_PARAMETERS * nc;
  CUDA_TRY(cudaMemcpyAsync(p->d_params, p->h_params, sizeof(double) * n, cudaMemcpyHostToDevice, st));
  p->tables_dirty = true;
  return check_err_flag(p, st);
}

// uncertainties = sqrt(diag(inv(0.5 H))), analysis/fitter.py:484-485 (Gauss-Jordan on the small matrix)
static void invert_half_hessian(const std::vector<double>& H, int n, double* uncertainty) {
  std::vector<double> M((size_t)n * 2 * n, 0.0);
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < n; ++j) M[(size_t)i * 2 * n + j] = 0.5 * H[(size_t)i * n + j];
    M[(size_t)i * 2 * n + n + i] = 1.0;
  }
  bool ok = true;
  for (int k = 0; k < n && ok; ++k) {
    int piv = k;
    for (int r = k + 1; r < n; ++r)
      if (std::fabs(M[(size_t)r * 2 * n + k]) > std::fabs(M[(size_t)piv * 2 * n + k])) piv = r;
    if (M[(size_t)piv * 2 * n + k] == 0.0) { ok = false; break; }
    if (piv != k)
      for (int j = 0; j < 2 * n; ++j) std::swap(M[(size_t)k * 2 * n + j], M[(size_t)piv * 2 * n + j]);
    const double d = M[(size_t)k * 2 * n + k];
    for (int j = 0; j < 2 * n; ++j) M[(size_t)k * 2 * n + j] /= d;
    for (int r = 0; r < n; ++r) {
      if (r == k) continue;
      const double f = M[(size_t)r * 2 * n + k];
      if (f != 0.0)
        for (int j = 0; j < 2 * n; ++j) M[(size_t)r * 2 * n + j] -= f * M[(size_t)k * 2 * n + j];
    }
  }
  for (int i = 0; i < n; ++i) uncertainty[i] = ok ? std::sqrt(M[(size_t)i * 2 * n + n + i]) : NAN;
}

// ---- on-device batched solver (fb_batched.cuh) -----------------------------------------------------
struct DevBuf {
  void* ptr = nullptr;
  ~DevBuf() { cudaFree(ptr); }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&ptr, bytes ? bytes : 1); }
  template <typename T> T* as() { return reinterpret_cast<T*>(ptr); }
};

static int fit_batched_device(fb_plan_s* p, int count, const double* data_dev, const double* weights_dev,
                              double* params_dev, const fb_fit_options& opt, fb_fit_result* results_host,
                              cudaStream_t st) {
  const int nf = p->desc.num_freq, nt = p->desc.num_time, nc = p->desc.num_components, kf = p->pc.kf;
  const int n = p->n_free, ncols = n + 1;
  // per-burst lists of channels with non-zero weight (host-built: deterministic order)
  std::vector<double> w((size_t)count * nf);
  CUDA_TRY(cudaMemcpyAsync(w.data(), weights_dev, sizeof(double) * w.size(), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  std::vector<int> lists((size_t)count * nf, 0), num_good(count, 0);
  int max_good = 0;
  for (int b = 0; b < count; ++b) {
    int k = 0;
    for (int i = 0; i < nf; ++i)
      if (w[(size_t)b * nf + i] != 0.0) lists[(size_t)b * nf + k++] = i;
    num_good[b] = k;
    max_good = std::max(max_good, k);
  }
  const int tiles = (nt + kBlock - 1) / kBlock;
  const long long items = p->pc.scintillation ? (long long)max_good : (long long)max_good * tiles;
  int blocks = (int)std::max(1LL, std::min(64LL, (items + 3) / 4));
  const int ntiles = gram_ntiles(ncols);
  DevBuf d_lists, d_num_good, d_done, d_started, d_num_done, d_states, d_results, d_gram, d_comp, d_sub, d_chan, d_part;
  cudaError_t e = cudaSuccess;
  auto chk = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
  chk(d_lists.alloc(sizeof(int) * lists.size()));
  chk(d_num_good.alloc(sizeof(int) * count));
  chk(d_done.alloc(sizeof(int) * count));
  chk(d_started.alloc(sizeof(int) * count));
  chk(d_num_done.alloc(sizeof(int)));
  chk(d_states.alloc(sizeof(TrfState) * (size_t)count));
  chk(d_results.alloc(sizeof(fb_fit_result) * (size_t)count));
  chk(d_gram.alloc(sizeof(double) * (size_t)count * ncols * ncols));
  chk(d_comp.alloc(sizeof(CompTab) * (size_t)count * nc));
  chk(d_sub.alloc(sizeof(SubTab) * (size_t)count * nf * nc * kf));
  chk(d_chan.alloc(sizeof(ChanTab) * (size_t)count * nf * nc));
  chk(d_part.alloc(sizeof(double) * (size_t)count * blocks * ntiles * kTile * kTile));
  if (e != cudaSuccess) return fail(FB_ERR_NOMEM, "batched fit allocation failed: %s", cudaGetErrorString(e));
  CUDA_TRY(cudaMemcpyAsync(d_lists.ptr, lists.data(), sizeof(int) * lists.size(), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(d_num_good.ptr, num_good.data(), sizeof(int) * count, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemsetAsync(d_done.ptr, 0, sizeof(int) * count, st));
  CUDA_TRY(cudaMemsetAsync(d_started.ptr, 0, sizeof(int) * count, st));
  CUDA_TRY(cudaMemsetAsync(d_num_done.ptr, 0, sizeof(int), st));
  CUDA_TRY(cudaMemsetAsync(d_results.ptr, 0, sizeof(fb_fit_result) * (size_t)count, st));

  EvalArgs A = base_args(p);
  A.data = data_dev;
  A.weights = weights_dev;
  A.params = params_dev;
  A.g_comp = d_comp.as<CompTab>();
  A.g_sub = d_sub.as<SubTab>();
  A.g_chan = d_chan.as<ChanTab>();
  A.chan_list = d_lists.as<int>();
  A.partials = d_part.as<double>();
  BatchArgs B;
  B.done = d_done.as<int>();
  B.num_good = d_num_good.as<int>();
  B.data_stride = (size_t)nf * nt;
  B.weights_stride = nf;
  B.params_stride = (size_t)FB_NUM_PARAMETERS * nc;
  B.comp_stride = nc;
  B.sub_stride = (size_t)nf * nc * kf;
  B.chan_stride = (size_t)nf * nc;
  B.list_stride = nf;
  B.partials_stride = (size_t)blocks * ntiles * kTile * kTile;
  BatchedTrfArgs T;
  memset(&T, 0, sizeof(T));
  T.count = count;
  T.nc = nc;
  T.n_free = n;
  for (int k = 0; k < 9; ++k) T.col_of[k] = p->col_of[k];
  T.num_residuals = (long long)nf * nt;
  T.ftol = opt.ftol;
  T.xtol = opt.xtol;
  T.gtol = opt.gtol;
  T.max_nfev = opt.max_nfev;
  T.partials = d_part.as<double>();
  T.blocks = blocks;
  T.gram = d_gram.as<double>();
  T.params = params_dev;
  T.states = d_states.as<TrfState>();
  T.started = d_started.as<int>();
  T.done = d_done.as<int>();
  T.num_done = d_num_done.as<int>();
  T.results = d_results.as<fb_fit_result>();

  const size_t smem = smem_bytes(nc, kf, ncols, p->pc.scintillation != 0);
  if (smem > 227 * 1024) return fail(FB_ERR_INVALID, "shared-memory request %zu B exceeds 227 KB", smem);
  if (smem > 48 * 1024)
    CUDA_TRY(cudaFuncSetAttribute(k_eval_batched<MODE_GRAM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const size_t trf_smem = sizeof(double) * 2 * (size_t)n * n;
  if (trf_smem > 48 * 1024)
    CUDA_TRY(cudaFuncSetAttribute(k_trf_batched, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)trf_smem));
  const int max_iter = (opt.max_nfev > 0 ? opt.max_nfev : 100 * n) + 2;
  const dim3 tgrid((nf * nc + 127) / 128, count), egrid(blocks, count);
  int done_host = 0;
  for (int it = 0; it < max_iter && done_host < count; ++it) {
    k_tables_batched<<<tgrid, 128, 0, st>>>(p->pc, params_dev, d_comp.as<CompTab>(), d_sub.as<SubTab>(),
                                            d_chan.as<ChanTab>(), d_done.as<int>());
    k_eval_batched<MODE_GRAM><<<egrid, kBlock, smem, st>>>(A, B);
    k_trf_batched<<<count, 32, trf_smem, st>>>(T);
    CUDA_TRY(cudaGetLastError());
    p->launches += 3;
    CUDA_TRY(cudaMemcpyAsync(&done_host, d_num_done.ptr, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
  }
  CUDA_TRY(cudaMemcpyAsync(results_host, d_results.ptr, sizeof(fb_fit_result) * (size_t)count, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  p->tables_dirty = true;
  int rc_err = check_err_flag(p, st);
  if (rc_err != FB_OK || !opt.compute_hessian || p->pc.scintillation) return rc_err;

  // ---- exact Hessians of all bursts at once, at their best-fit points (the parameter blocks hold
  // them): tables, second-derivative sums S_ab (k_hessian_acc, blockIdx.y = burst), J^T J from the
  // solver states; H = 2 J^T J - 2 S and the 1-sigma uncertainties on the host (n x n per burst).
  int rc = ensure_hessian_layout(p, st);
  if (rc != FB_OK) return rc;
  const int npairs = (int)p->hess_pairs.size();
  std::vector<int> skip(count);
  for (int b = 0; b < count; ++b) skip[b] = results_host[b].status < 0 ? 1 : 0;
  CUDA_TRY(cudaMemcpyAsync(d_done.ptr, skip.data(), sizeof(int) * count, cudaMemcpyHostToDevice, st));
  k_tables_batched<<<tgrid, 128, 0, st>>>(p->pc, params_dev, d_comp.as<CompTab>(), d_sub.as<SubTab>(),
                                          d_chan.as<ChanTab>(), d_done.as<int>());
  CUDA_TRY(cudaGetLastError());
  const int supers = (tiles + kHessSuper - 1) / kHessSuper;
  const int hblocks = (int)std::max(1LL, std::min(64LL, ((long long)max_good * supers + 3) / 4));
  DevBuf d_hpart, d_hsum;
  if (d_hpart.alloc(sizeof(double) * (size_t)count * hblocks * npairs) != cudaSuccess ||
      d_hsum.alloc(sizeof(double) * (size_t)count * npairs) != cudaSuccess)
    return fail(FB_ERR_NOMEM, "batched Hessian allocation failed");
  HessAccArgs H;
  memset(&H, 0, sizeof(H));
  H.E = A;
  H.desc = p->d_desc;
  H.num_pairs = npairs;
  H.comp_list = p->d_comp_list;
  for (int c = 0; c <= nc; ++c) H.comp_off[c] = p->hess_comp_off[c];
  for (int c = 0; c < nc; ++c) H.need[c] = p->hess_need[c];
  H.need_sum_dmidx = p->hess_need_sum_dmidx;
  H.partials = d_hpart.as<double>();
  H.gvals = nullptr;
  H.gvals_layout = 2;
  BatchArgs BH = B;
  BH.partials_stride = (size_t)hblocks * npairs;
  const size_t hsmem = hess_acc_smem_bytes(nc, kf, npairs);
  if (hsmem > 227 * 1024) return fail(FB_ERR_INVALID, "Hessian shared-memory request %zu B exceeds 227 KB", hsmem);
  if (hsmem > 48 * 1024)
    CUDA_TRY(cudaFuncSetAttribute(k_hessian_acc_batched, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hsmem));
  k_hessian_acc_batched<<<dim3(hblocks, count), kBlock, hsmem, st>>>(H, BH, d_done.as<int>());
  CUDA_TRY(cudaGetLastError());
  k_hess_finalize_batched<<<dim3((npairs + 3) / 4, count), 128, 0, st>>>(d_hpart.as<double>(), hblocks, npairs,
                                                                         (size_t)hblocks * npairs, d_hsum.as<double>());
  CUDA_TRY(cudaGetLastError());
  p->launches += 3;
  std::vector<double> S((size_t)count * npairs), JtJ((size_t)count * n * n), Hm((size_t)n * n);
  CUDA_TRY(cudaMemcpyAsync(S.data(), d_hsum.ptr, sizeof(double) * S.size(), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpy2DAsync(JtJ.data(), sizeof(double) * n * n,
                             reinterpret_cast<const char*>(d_states.ptr) + offsetof(TrfState, A), sizeof(TrfState),
                             sizeof(double) * n * n, count, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  for (int b = 0; b < count; ++b) {
    if (skip[b]) {
      for (int k = 0; k < n; ++k) results_host[b].uncertainty[k] = NAN;
      continue;
    }
    const double* Ab = JtJ.data() + (size_t)b * n * n;  // J^T J at the accepted point (row-major, ld = n)
    for (int a = 0; a < n * n; ++a) Hm[a] = 2.0 * Ab[a];
    for (int k = 0; k < npairs; ++k) {
      const int a = p->hess_pair_a[k], c = p->hess_pair_b[k];
      const double v = 2.0 * S[(size_t)b * npairs + k];
      Hm[(size_t)a * n + c] -= v;
      if (a != c) Hm[(size_t)c * n + a] -= v;
    }
    invert_half_hessian(Hm, n, results_host[b].uncertainty);
  }
  p->tables_dirty = true;
  return FB_OK;
}

// ---- batched independent fits (BASELINE.json config 3) ---------------------------------------------
extern "C" int fb_fit_batched(fb_plan_t p, int32_t count, const double* data_dev, const double* weights_dev,
                              double* params_dev, const fb_fit_options* options, fb_fit_result* results_host,
                              void* stream) {
  if (!p || !data_dev || !weights_dev || !params_dev || !results_host) return fail(FB_ERR_INVALID, "null argument");
  if (count < 0) return fail(FB_ERR_INVALID, "negative count");
  cudaStream_t st = (cudaStream_t)stream;
  const int nf = p->desc.num_freq, nc = p->desc.num_components, n = p->n_free;
  const size_t spec = (size_t)nf * p->desc.num_time, nparam = (size_t)FB_NUM_PARAMETERS * nc;
  fb_fit_options opt = default_options(options);
  std::vector<double> w(nf), pr(nparam), H((size_t)n * n);
  // Many small bursts: the whole trust-region loop runs on the device for all bursts at once.
  // Large bursts already fill the GPU one at a time and their tables would not fit `count` times.
  const char* mode = getenv("FITBURST_B200_BATCHED");
  const size_t table_bytes = (size_t)count * nf * nc * (p->pc.kf * sizeof(SubTab) + sizeof(ChanTab));
  bool device_path = count >= 2 && table_bytes <= ((size_t)8 << 30) && spec <= ((size_t)1 << 22);
  if (mode && !strcmp(mode, "device")) device_path = true;
  if (mode && !strcmp(mode, "sequential")) device_path = false;
  bool fitted_on_device = false;
  if (device_path && n >= 1 && count >= 1) {
    int rc = fit_batched_device(p, count, data_dev, weights_dev, params_dev, opt, results_host, st);
    if (rc != FB_OK) return rc;
    fitted_on_device = true;
  }
  for (int b = 0; b < count; ++b) {
    if (fitted_on_device) {
      if (!opt.compute_hessian || results_host[b].status < 0 || !p->pc.scintillation) continue;  // done on the device
      // scintillation mode: exact Hessian at the best fit, burst by burst (general kernel)
      CUDA_TRY(cudaMemcpyAsync(w.data(), weights_dev + (size_t)b * nf, sizeof(double) * nf, cudaMemcpyDeviceToHost, st));
      CUDA_TRY(cudaMemcpyAsync(pr.data(), params_dev + (size_t)b * nparam, sizeof(double) * nparam, cudaMemcpyDeviceToHost, st));
      CUDA_TRY(cudaStreamSynchronize(st));
      int rc = fb_load_weights(p, w.data(), stream);
      if (rc == FB_OK) rc = fb_set_parameters(p, pr.data(), stream);
      if (rc == FB_OK) rc = fb_hessian_impl(p, data_dev + (size_t)b * spec, H.data(), st);
      if (rc != FB_OK) {
        for (int k = 0; k < n; ++k) results_host[b].uncertainty[k] = NAN;  // never a silent 0.0
        continue;
      }
      invert_half_hessian(H, n, results_host[b].uncertainty);
      continue;
    }
    CUDA_TRY(cudaMemcpyAsync(w.data(), weights_dev + (size_t)b * nf, sizeof(double) * nf, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(pr.data(), params_dev + (size_t)b * nparam, sizeof(double) * nparam, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    int rc = fb_load_weights(p, w.data(), stream);
    if (rc == FB_OK) rc = fb_set_parameters(p, pr.data(), stream);
    if (rc != FB_OK) return rc;
    fb_fit_options o1 = opt;
    o1.compute_hessian = 0;
    rc = fb_fit(p, data_dev + (size_t)b * spec, &o1, &results_host[b], nullptr, nullptr, stream);
    if (rc != FB_OK) {
      results_host[b].status = -1;
      continue;  // one bad burst must not abort the batch (fit() never raises, analysis/fitter.py:320-322)
    }
    // best fit back into the caller's parameter block
    unpack_free(p, results_host[b].x);
    CUDA_TRY(cudaMemcpyAsync(params_dev + (size_t)b * nparam, p->h_params, sizeof(double) * nparam,
                             cudaMemcpyHostToDevice, st));
    if (opt.compute_hessian) {
      rc = fb_set_parameters(p, p->h_params, stream);
      if (rc == FB_OK) rc = fb_hessian_impl(p, data_dev + (size_t)b * spec, H.data(), st);
      if (rc == FB_OK) {
        invert_half_hessian(H, n, results_host[b].uncertainty);
      } else {
        for (int k = 0; k < n; ++k) results_host[b].uncertainty[k] = NAN;
      }
    }
    CUDA_TRY(cudaStreamSynchronize(st));
  }
  return FB_OK;
}

