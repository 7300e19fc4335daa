/*
 * fitburst_b200 -- C ABI of the B200-native fitburst hot path (sm_100a).
 *
 * The reference (CHIMEFRB/fitburst) is pure Python and has no FFI: its "operator interface"
 * for this path is the SpectrumModeler / LSFitter class surface. The Python host mirror in
 * fitburst_b200/analysis/{model,fitter}.py binds exactly the entry points below through ctypes
 * (the stub a maintainer would add to the reference is shown in INTEGRATION.md). Each entry
 * point cites the reference code it replaces, as path:line under fitburst/ in the reference.
 *
 * Conventions
 *  - every function returns 0 on success, a negative fb_status otherwise; fb_last_error()
 *    returns a thread-local message for the last failure.
 *  - pointers named *_dev are BORROWED device pointers (owned by the caller, e.g. torch
 *    tensors); pointers named *_host are host pointers. No torch types cross this boundary.
 *  - all floating-point data are IEEE binary64. Spectra are C-order (num_freq, num_time);
 *    per-component caches are C-order (num_freq, num_time, num_components) like
 *    analysis/model.py:110-124.
 *  - `stream` is a cudaStream_t passed as void* (NULL = default stream). Calls that return
 *    results through host pointers synchronise that stream before returning; all others are
 *    asynchronous with respect to the host.
 *  - a plan is not thread-safe; use one plan per host thread / stream.
 */
#ifndef FITBURST_B200_H
#define FITBURST_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FB_MAX_COMPONENTS 16
#define FB_NUM_PARAMETERS 10 /* 9 fittable (analysis/model.py:92-102) + ref_freq */
#define FB_MAX_FREE 67       /* upper bound on packed free parameters */

/* Row index of each parameter in the packed (FB_NUM_PARAMETERS, num_components) block.
 * Order of the first nine = SpectrumModeler.parameters (analysis/model.py:92-102), which is also
 * the packing order of LSFitter.get_fit_parameters_list (analysis/fitter.py:353-384). */
enum fb_parameter {
  FB_AMPLITUDE = 0,
  FB_ARRIVAL_TIME = 1,
  FB_BURST_WIDTH = 2,
  FB_DM = 3,
  FB_DM_INDEX = 4,
  FB_SCATTERING_TIMESCALE = 5,
  FB_SCATTERING_INDEX = 6,
  FB_SPECTRAL_INDEX = 7,
  FB_SPECTRAL_RUNNING = 8,
  FB_REF_FREQ = 9
};

enum fb_status {
  FB_OK = 0,
  FB_ERR_INVALID = -1,
  FB_ERR_CUDA = -2,
  FB_ERR_NOMEM = -3,
  FB_ERR_SINGULAR = -4
};

/* Model configuration = the constructor arguments of SpectrumModeler (analysis/model.py:23-124)
 * plus the two values the reference reads from backend/general.yaml at call time. */
typedef struct fb_plan_desc {
  int32_t num_freq;
  int32_t num_time;
  int32_t num_components;
  int32_t factor_freq_upsample;
  int32_t factor_time_upsample;
  int32_t is_folded;     /* analysis/model.py:327-332,349-350 */
  int32_t scintillation; /* analysis/model.py:267-277 */
  int32_t normalize_pbf; /* backend/general.yaml:12, read at analysis/model.py:337 */
  double dm_incoherent;
  double dispersion_constant; /* backend/general.yaml:7 */
} fb_plan_desc;

typedef struct fb_plan_s* fb_plan_t;

const char* fb_last_error(void);
int fb_version(void);

/* SpectrumModeler.__init__ (analysis/model.py:23-124). freqs_host (num_freq), times_host
 * (num_time) are copied; res_freq / res_time are derived from the first two labels like
 * analysis/model.py:87-88 and the sub-frequency labels follow routines/manipulate.py:75-107. */
int fb_plan_create(const fb_plan_desc* desc, const double* freqs_host, const double* times_host,
                   fb_plan_t* plan_out);
int fb_plan_destroy(fb_plan_t plan);

/* SpectrumModeler.update_parameters (analysis/model.py:383-409): params_host is the packed
 * (FB_NUM_PARAMETERS, num_components) block; global rows are broadcast from element 0 by the
 * kernels exactly like analysis/model.py:155-159 reads them. */
int fb_set_parameters(fb_plan_t plan, const double* params_host, void* stream);
int fb_get_parameters(fb_plan_t plan, double* params_host, void* stream);

/* SpectrumModeler.compute_model (analysis/model.py:126-279). model_dev (num_freq, num_time) is
 * required. The four caches (analysis/model.py:110-124) are written only where a non-NULL
 * pointer is given. data_dev is required in scintillation mode (analysis/model.py:140-141). */
int fb_model_eval(fb_plan_t plan, const double* data_dev, double* model_dev,
                  double* amplitude_dev, double* spectrum_dev, double* timediff_dev,
                  double* timeprof_dev, void* stream);

/* SpectrumModeler.compute_profile (analysis/model.py:281-352) = routines/profile.py:15-114 on an
 * arbitrary (num_rows, num_cols) array of times: row r is evaluated at frequency freqs_dev[r]
 * (scattering time sc_time_ref * (f / ref_freq) ** sc_index). scattered != 0 selects the
 * pulse-broadening function with the -5 width clamp (analysis/model.py:339-341), else the
 * Gaussian; is_folded averages the second realisation (analysis/model.py:327-332,349-350);
 * arrival_time is subtracted like routines/profile.py:91. */
int fb_profile_eval(const double* times_dev, const double* freqs_dev, int32_t num_rows, int32_t num_cols,
                    double arrival_time, double sc_time_ref, double sc_index, double width, double ref_freq,
                    int32_t scattered, int32_t clamp, int32_t is_folded, int32_t normalize_pbf,
                    double* profile_dev, void* stream);

/* LSFitter._set_weights (analysis/fitter.py:506-542): w_i = 1/sqrt(mean(data[i, lo:hi]^2)) for
 * good channels (1 if not weighted_fit), 0 for flagged ones. good_host is a uint8 mask of length
 * num_freq. The plan keeps the weights on the device (and the list of channels with non-zero
 * weight, which is all the fit kernels visit); weights_host (optional, num_freq) receives a
 * copy, chisq_initial_host (optional) receives sum((data * w)^2) (analysis/fitter.py:465). */
int fb_set_weights(fb_plan_t plan, const double* data_dev, const uint8_t* good_host,
                   int32_t weighted_fit, int32_t range_lo, int32_t range_hi,
                   double* weights_host, double* chisq_initial_host, void* stream);
/* Caller-supplied weights (LSFitter.weights may be overwritten by user code). */
int fb_load_weights(fb_plan_t plan, const double* weights_host, void* stream);

/* Selects the free parameters: free_mask[p] != 0 for p in 0..8 means parameter p is fitted
 * (LSFitter.fit_parameters after fix_parameter, analysis/fitter.py:324-351). Returns the
 * packed length n through num_free_out (analysis/fitter.py:353-384: fitter-global names
 * dm, dm_index, scattering_timescale, scattering_index contribute one scalar, others
 * num_components). */
int fb_set_free_parameters(fb_plan_t plan, const int32_t* free_mask, int32_t* num_free_out);

/* One fused pass = LSFitter.compute_residuals + compute_jacobian + the J^T J / J^T r / chi^2
 * contractions that scipy's TRF needs (analysis/fitter.py:180-266, routines/derivative.py:454-936).
 * Evaluates at the plan's current parameters with the plan's weights. gram_host is the symmetric
 * (n+1, n+1) matrix [J r]^T [J r] in C order: J^T J in the leading n x n block, J^T r in the
 * last column, chi^2 = r^T r in the corner. Flagged channels (weight 0) contribute exactly 0
 * and are skipped. */
int fb_normal_equations(fb_plan_t plan, const double* data_dev, double* gram_host, void* stream);
/* Same, leaving the result on the device (gram_dev, (n+1)*(n+1) doubles), no host sync. */
int fb_normal_equations_dev(fb_plan_t plan, const double* data_dev, double* gram_dev, void* stream);

/* LSFitter.compute_residuals (analysis/fitter.py:231-266) at the current parameters:
 * resid_dev[i*num_time + j] = (data - model) * w_i. */
int fb_residuals(fb_plan_t plan, const double* data_dev, double* resid_dev, void* stream);

/* LSFitter.compute_jacobian (analysis/fitter.py:180-229) at the current parameters, dense:
 * jac_dev is (num_freq*num_time, n) C order, column k = -d model / d theta_k * w. Compatibility
 * path for results.jac; the fit itself never materialises J. */
int fb_jacobian_dense(fb_plan_t plan, const double* data_dev, double* jac_dev, void* stream);

/* routines/derivative.py:454-936: one first-derivative map d model / d theta (num_freq,
 * num_time) for parameter `param` and component `component`; add_all sums the per-component
 * maps of the four global-capable parameters like the reference's add_all=True. */
int fb_derivative_map(fb_plan_t plan, const double* data_dev, int32_t param, int32_t component,
                      int32_t add_all, double* out_dev, void* stream);

/* LSFitter.compute_hessian (analysis/fitter.py:81-178) with the 45 mixed partials of
 * routines/derivative.py:938-2836: hessian_host is (n, n) C order,
 * H_ab = sum 2 (d_a m d_b m - (data - m) d_ab m) w^2 over the free parameters. */
int fb_hessian(fb_plan_t plan, const double* data_dev, double* hessian_host, void* stream);
/* routines/derivative.py:938-2836: one mixed-partial map (num_freq, num_time) for the ordered
 * pair (param1, param2) as LSFitter.compute_hessian would request it. */
int fb_derivative2_map(fb_plan_t plan, const double* data_dev, int32_t param1, int32_t param2,
                       int32_t component, double* out_dev, void* stream);

/* Result of one least-squares fit = the scalar fields of scipy's OptimizeResult as produced by
 * least_squares(method='trf') without bounds (analysis/fitter.py:300). */
typedef struct fb_fit_result {
  int32_t status; /* scipy meanings: 0 max_nfev, 1 gtol, 2 ftol, 3 xtol, 4 ftol+xtol, -1 failure */
  int32_t nfev;
  int32_t njev;
  int32_t num_free;
  double cost;       /* 0.5 * chi^2 at x */
  double optimality; /* ||J^T r||_inf at x */
  double chisq_initial;
  double x[FB_MAX_FREE];
  double grad[FB_MAX_FREE];
  double uncertainty[FB_MAX_FREE]; /* sqrt(diag(inv(0.5 H))), analysis/fitter.py:484-489 */
} fb_fit_result;

typedef struct fb_fit_options {
  double ftol, xtol, gtol; /* scipy defaults 1e-8 */
  int32_t max_nfev;        /* <= 0: scipy default 100 * n */
  int32_t compute_hessian; /* != 0: exact Hessian + uncertainties after the fit */
} fb_fit_options;

/* LSFitter.fit(exact_jacobian=True) (analysis/fitter.py:268-322): trust-region-reflective
 * least squares without bounds, restated on the (n+1, n+1) normal-equation matrix, starting
 * from the plan's current parameters; leaves the plan at the last evaluated point like
 * analysis/fitter.py:258. gram_final_host (optional, (n+1)^2) receives [J r]^T [J r] at x
 * (hessian_approx of analysis/fitter.py:481); hessian_host (optional, n*n) the exact Hessian. */
int fb_fit(fb_plan_t plan, const double* data_dev, const fb_fit_options* options,
           fb_fit_result* result, double* gram_final_host, double* hessian_host, void* stream);

/* The trust-region state machine behind fb_fit, exposed so that a caller can insert its own
 * reduction between the evaluation and the step -- the channel-sharded multi-GPU fit
 * (BASELINE.json config 5) all-reduces the (n+1)^2 normal-equation matrix over NCCL and then every
 * rank advances an identical copy of this state (scipy/optimize/_lsq/trf.py:415-587 restated,
 * see fitburst_b200/csrc/fb_trf.h). Protocol: create(x0) -> start(gram at x0) ->
 * while (propose(x_trial) == 1) { evaluate gram at x_trial; update(gram); } -> result(). */
typedef struct fb_trf_s* fb_trf_t;
int fb_trf_create(int32_t num_free, int64_t num_residuals, const double* x0_host,
                  const fb_fit_options* options, fb_trf_t* trf_out);
int fb_trf_destroy(fb_trf_t trf);
int fb_trf_start(fb_trf_t trf, const double* gram_host);
/* returns 1 and fills x_trial_host when a trial point must be evaluated, 0 when finished */
int fb_trf_propose(fb_trf_t trf, double* x_trial_host);
int fb_trf_update(fb_trf_t trf, const double* gram_host);
int fb_trf_result(fb_trf_t trf, fb_fit_result* result, double* gram_accepted_host);

/* fb_fit keeps the whole trust-region loop on the device: after every evaluation one warp runs
 * the state machine above on the finished matrix and writes the next trial point into the plan's
 * device parameter block, evaluations are enqueued ahead in chunks and return at once after the
 * fit has stopped (one flag read-back per chunk instead of one synchronisation per trial point).
 * The same loop in steps, for callers that reduce the matrix over several GPUs in between (the
 * channel-sharded fit: NCCL all-reduce of gram_dev on the same stream between _evaluate and
 * _step; every rank then takes identical decisions):
 *   _begin(x0 = the plan's parameters) -> repeat { _evaluate(gram_dev); [all-reduce]; _step(gram_dev) }
 *   in chunks, _poll after each chunk until *stopped_out != 0 -> _end().
 * num_residuals_total: number of residuals of the WHOLE problem (scipy's full-rank threshold,
 * scipy/optimize/_lsq/common.py:94-101); <= 0 means the plan's own num_freq * num_time.
 * *bailed_out != 0: a trial point left the domain of the fast kernels (scattering_timescale or a
 * burst_width <= 0); result->status is -100 and the caller continues with the fb_trf_* protocol. */
int fb_fit_device_begin(fb_plan_t plan, const fb_fit_options* options, int64_t num_residuals_total, void* stream);
int fb_fit_device_evaluate(fb_plan_t plan, const double* data_dev, double* gram_dev, void* stream);
int fb_fit_device_step(fb_plan_t plan, const double* gram_dev, void* stream);
int fb_fit_device_poll(fb_plan_t plan, int32_t* stopped_out, int32_t* evaluations_out, void* stream);
int fb_fit_device_end(fb_plan_t plan, fb_fit_result* result, double* gram_accepted_host, int32_t* bailed_out,
                      void* stream);

/* Batched independent fits (BASELINE.json config 3): `count` bursts of identical shape, burst b
 * occupying data_dev + b*num_freq*num_time, weights_dev + b*num_freq, params_dev + b *
 * FB_NUM_PARAMETERS*num_components (start point in, best fit out; all on the device). Bursts
 * are fitted back to back, each using the whole GPU; results_host[b] receives the scalar fields
 * and, when options->compute_hessian is set, the 1-sigma uncertainties. */
/* With options->compute_hessian the exact Hessian is evaluated at the BEST FIT of every burst
 * (fb_fit / LSFitter evaluate it at the last evaluated point like the reference, analysis/fitter.py:
 * 105-108; the two coincide whenever the last trial step was accepted) -- for small bursts all
 * Hessians in one launch on the device. A burst whose Hessian cannot be computed gets NaN. */
int fb_fit_batched(fb_plan_t plan, int32_t count, const double* data_dev, const double* weights_dev,
                   double* params_dev, const fb_fit_options* options, fb_fit_result* results_host,
                   void* stream);

/* chi^2 surface over a (dm, scattering_timescale) grid (BASELINE.json config 4): reference
 * equivalent np.sum(LSFitter.compute_residuals(x)**2) per grid point. chisq_dev is (num_dm,
 * num_sc) C order; all other parameters are the plan's current ones. */
int fb_chisq_grid(fb_plan_t plan, const double* data_dev,
                  const double* dm_values_host, int32_t num_dm, const double* sc_values_host,
                  int32_t num_sc, double* chisq_dev, void* stream);

/* ---- pre-fit data preparation (utilities/bases.py:129-336), SURVEY.md section 8f.1 ----------
 * All arrays are device pointers; spectra are (num_freq, num_time) f64 C order. */
/* per channel [sum(w), sum(d*w), min(d), max(d)] -> moments_dev (num_freq, 4); bases.py:251-264 */
int fb_pre_channel_moments(const double* data_dev, const double* weights_dev, int32_t num_freq,
                           int32_t num_time, double* moments_dev, void* stream);
/* bases.py:270-273: good channels d <- d / mean - 1, then samples with zero weight are zeroed */
int fb_pre_remove_baseline(double* data_dev, const double* weights_dev, const uint8_t* good_dev,
                           const double* mean_dev, int32_t num_freq, int32_t num_time, void* stream);
/* per channel [sum(d^2), sum(d^3)] -> sums_dev (num_freq, 2); bases.py:276-279 */
int fb_pre_power_sums(const double* data_dev, int32_t num_freq, int32_t num_time, double* sums_dev,
                      void* stream);
/* data[i, :] = 0 where keep_dev[i] == 0; bases.py:296,168 */
int fb_pre_zero_channels(double* data_dev, const uint8_t* keep_dev, int32_t num_freq, int32_t num_time,
                         void* stream);
/* downsample_2d, routines/manipulate.py:43-72; out is (num_freq/factor_freq, num_time/factor_time) */
int fb_pre_downsample_2d(const double* in_dev, int32_t num_freq, int32_t num_time, int32_t factor_freq,
                         int32_t factor_time, double* out_dev, void* stream);
/* window_data, bases.py:300-336: out[i, k] = in[i, start_dev[i] + k] for k < width; the caller
 * guarantees 0 <= start[i] and start[i] + width <= num_time */
int fb_pre_window(const double* in_dev, int32_t num_freq, int32_t num_time, const int32_t* start_dev,
                  int32_t width, double* out_dev, void* stream);

/* band-averaged time series out_dev[j] = mean_i in_dev[i, j] (FindPeak.find_peak,
 * analysis/peak_finder.py:80 `self.data.mean(axis=0)`); scratch_dev holds
 * ceil(num_freq / 64) * num_time doubles */
int fb_pre_time_series(const double* in_dev, int32_t num_freq, int32_t num_time, double* scratch_dev,
                       double* out_dev, void* stream);

/* Host-to-device transfer of the usable channels only: copies rows rows_dev[0..num_rows) of a
 * PAGE-LOCKED host spectrum (num_freq x num_time, float32 or float64: elem_size 4 / 8) into the
 * float64 device spectrum dst_dev, reading the host memory directly over PCIe. Channels with
 * good_freq == False get weight 0 (LSFitter._set_weights, analysis/fitter.py:506-542) and are
 * never read by the fit, so their rows are left as they are on the device. */
int fb_upload_rows(const void* host_pinned, int32_t elem_size, const int32_t* rows_dev, int32_t num_rows,
                   int32_t num_freq, int32_t num_time, double* dst_dev, void* stream);

/* ---- small public routines of the reference, elementwise over device arrays ------------------ */
/* compute_spectrum_rpl, routines/spectrum.py:11-43: exp(sp_idx L + sp_run L^2), L = ln(f / freq_ref) */
int fb_spectrum_rpl(const double* freqs_dev, int64_t count, double freq_ref, double sp_idx, double sp_run,
                    double* out_dev, void* stream);
/* routines/ism.py:50-140 over an array of frequencies f:
 *   kind 0 compute_time_dm_delay   a = dm_value, b = dm_const, c = dm_idx, d = freq2
 *   kind 1 compute_time_dm_smear   a = dm_value, b = dm_const, c = dm_idx, d = width_chan
 *   kind 2 compute_time_scattering a = ref_freq, b = sc_time_ref, c = sc_idx (d unused) */
int fb_ism_times(int32_t kind, const double* freqs_dev, int64_t count, double a, double b, double c, double d,
                 double* out_dev, void* stream);
/* compute_amplitude_per_channel, routines/ism.py:11-48: data_dev (num_time), model_dev (num_time,
 * num_components) C order -> amplitudes_dev (num_components); FB_ERR_SINGULAR like np.linalg.solve */
int fb_amplitude_per_channel(const double* data_dev, const double* model_dev, int32_t num_time, int32_t num_comp,
                             double* amplitudes_dev, void* stream);
/* exact float32 -> float64 widening of an uploaded spectrum (telescope data are float32; the
 * upload then moves half the bytes); both buffers 16-byte aligned */
int fb_widen_f32(const float* src_dev, int64_t count, double* dst_dev, void* stream);

/* Measured FP64 FMA peak of the device (dependent-free DFMA chains on every SM), TFLOP/s.
 * MEASURED_PEAKS.json carries no FP64 figure; bench.py reports this one as the roofline peak. */
int fb_measure_fp64_peak(double* tflops_out, void* stream);

/* Device-time accounting for bench.py: number of kernel launches issued by this plan since
 * the last reset. */
int fb_launch_count(fb_plan_t plan, int64_t* count_out, int32_t reset);

#ifdef __cplusplus
}
#endif
#endif /* FITBURST_B200_H */
