// fitburst_b200 device core: per-channel tables, the scatter-broadened profile and the
// per-sample model / first-derivative evaluation shared by every kernel.
//
// Reference citations are path:line under fitburst/ in CHIMEFRB/fitburst (pure Python); this
// file restates WHAT those lines compute, laid out for one thread per output sample.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/fitburst_b200.h"

namespace fb {

constexpr double kSqrtHalf = 0.70710678118654752440;       // 1/sqrt(2)
constexpr double kSqrt2 = 1.41421356237309504880;
constexpr double kSqrtHalfPi = 1.25331413731550025121;     // sqrt(pi/2)
constexpr double kTwoOverSqrtPi = 1.12837916709551257390;  // 2/sqrt(pi)
constexpr double kLn10 = 2.30258509299404568402;
constexpr double kTailArg = -6.0;  // erfc(6)/2 = 1.1e-17 < ulp/2: below this the EMG is its pure exponential tail

// Static (per-plan) configuration, passed by value to kernels.
struct PlanConst {
  int num_freq, num_time, num_comp;
  int kf, kt;
  int is_folded, scintillation, normalize_pbf;
  int ntk;  // num_time * kt
  double dm_incoherent, dm_const;
  double res_freq, res_time;
  double time_first, time_last;  // times[0], times[num_time - 1]
  const double* freqs;           // (num_freq) channel centres
  const double* subfreqs;        // (num_freq, kf) routines/manipulate.py:99-105 labels
  // device-driven fit loop (fb_batched.cuh, k_trf_single): when non-null and *stop != 0 every
  // kernel of an evaluation returns at once -- launches enqueued past the end of a fit are no-ops
  const int* stop;
};

__device__ __forceinline__ bool fit_stopped(const PlanConst& pc) {
  return pc.stop != nullptr && *reinterpret_cast<const volatile int*>(pc.stop) != 0;
}

// Per-component constants (depend on parameters only).
struct CompTab {
  double t_start, t_step, t_stop;  // linspace of the up-sampled, arrival-corrected time labels
  double w, inv_w, neg5w;
  double amp10;                    // 10 ** amplitude (analysis/model.py:255-256)
  double ref_freq;
};

// Per (channel, sub-frequency, component) constants.
struct SubTab {
  double delay;     // analysis/model.py:184-199
  double ratio;     // width / sc_time(f_ia)             routines/profile.py:93
  double amp_term;  // sc_time_ref / sc_time, or sqrt(pi/2) * ratio   routines/profile.py:96
  double sed;       // running power law at f_ia         routines/spectrum.py:39-41
  // pure-exponential-tail fast path: sum_b P_ab / (kf kt) = tail_c * exp(tail_e - u0 * inv_tau)
  double inv_tau, tail_e, tail_c;
  double fold_start, fold_step, fold_stop;  // second realisation, analysis/model.py:327-332
};

// Per (channel, component) constants: regime boundaries of the model along the time axis and
// the channel-CENTRE quantities of the derivatives (routines/derivative.py evaluates everything
// at model.freqs, not at the sub-frequencies).
struct ChanTab {
  double amp;         // amplitude_per_component value: 10**A * S(f_i)   analysis/model.py:250-255
  // regimes (scattered branch, not folded): every sub-sample of a time bin is clamped at -5w
  // (analysis/model.py:340) when u_last - min_delay < -5w; every one is pure tail when u0 >= tail_lo.
  double min_delay, mean_delay, tail_lo;
  double plateau_p, plateau_ps;  // mean profile / 10**A * mean(profile * sed) on the clamp plateau
  double logf;        // ln(f_i / ref_freq)
  double sc_time;     // sc_time_ref * (f_i / ref_freq) ** sc_index
  double inv_tau;     // 1 / sc_time
  double amp_pbf;     // routines/derivative.py:71-103
  double inv_apbf;
  double d_dm;        // deriv_time_dm("dm")        routines/derivative.py:244-246
  double d_dmidx;     // deriv_time_dm("dm_index")  routines/derivative.py:248-250
  double dap_w, dap_tau, dap_alpha;  // deriv_amplitude_pbf with THIS component looked up (:105-152)
  double d2ap_alpha;  // deriv2_amplitude_pbf(scattering_index, scattering_index), :154-209
  double fpow, refpow;  // f_i ** dm_index, ref_freq ** dm_index
  double q1, q2;      // ln f f^b - ln fr fr^b  and  ln^2 f f^b - ln^2 fr fr^b (derivative.py:2267,2822)
  double ref_freq;
  int scattered;      // any(sc_time(f_ia) > 0) over the sub-frequencies, analysis/model.py:339
  int fast_ok;        // regime fast paths usable (scattered, not folded, finite tables)
  // integer sample boundaries the fast kernels need once per (channel, component); computed here,
  // by the table kernel (one thread per pair, all channels in parallel), because inside the fast
  // kernels they were a serial section behind a block barrier (ncu source view, round 2: a fifth
  // of k_gram3's stall samples sat on the two barriers around it)
  int near_lo, near_hi;  // near range of the derivatives' Gaussian factor (near_range)
  int jp, jt;            // first sample off the clamp plateau / first pure-tail sample (vals3_boundaries)
  int pad_[2];        // sizeof % 16 == 0: the fast kernels stage the tables with 16-byte cp.async
};
static_assert(sizeof(ChanTab) % 16 == 0 && sizeof(SubTab) % 16 == 0 && sizeof(CompTab) % 16 == 0,
              "tables are staged in 16-byte chunks");

__host__ __device__ inline double global_param(const double* params, int nc, int p) {
  return params[p * nc];  // element 0 is used for all components, analysis/model.py:155-159
}

// np.linspace(lo, hi, num)[m]: arange * step + start with two roundings, last element = stop.
__device__ __forceinline__ double linspace_at(double start, double step, double stop, int m, int num) {
  double v = __dadd_rn(__dmul_rn((double)m, step), start);
  return (m == num - 1 && num > 1) ? stop : v;
}

// Table-driven exp for the hot loops: x = (64 m + j) ln2/64 + r with |r| <= ln2/128,
// exp(x) = 2^m * 2^(j/64) * (1 + expm1(r)), expm1 by a degree-6 Taylor polynomial (truncation
// 1e-20). Relative error < 3e-16 (table entry rounding + one fma), i.e. as good as the library
// routine for our 1e-10 contract at about half its instruction count. Arguments whose result
// would be subnormal / overflow / NaN go through exp().
__device__ const double kExpTable[64] = {
    1.0, 1.0108892860517005, 1.0218971486541166, 1.0330248790212284,
    1.0442737824274138, 1.0556451783605572, 1.0671404006768237, 1.0787607977571199,
    1.0905077326652577, 1.102382583307841, 1.1143867425958924, 1.1265216186082418,
    1.1387886347566916, 1.1511892299529827, 1.1637248587775775, 1.1763969916502812,
    1.189207115002721, 1.202156731452703, 1.215247359980469, 1.22848053610687,
    1.241857812073484, 1.255380757024691, 1.2690509571917332, 1.2828700160787783,
    1.2968395546510096, 1.3109612115247644, 1.3252366431597413, 1.339667524053303,
    1.3542555469368927, 1.3690024229745905, 1.383909881963832, 1.3989796725383112,
    1.4142135623730951, 1.42961333839197, 1.4451808069770467, 1.460917794180647,
    1.4768261459394993, 1.4929077282912648, 1.5091644275934228, 1.5255981507445384,
    1.5422108254079407, 1.559004400237837, 1.5759808451078865, 1.593142151342267,
    1.6104903319492543, 1.6280274218573478, 1.645755478153965, 1.6636765803267364,
    1.681792830507429, 1.7001063537185235, 1.718619298122478, 1.7373338352737062,
    1.7562521603732995, 1.7753764925265212, 1.7947090750031072, 1.8142521755003989,
    1.8340080864093424, 1.8539791250833855, 1.8741676341103, 1.8945759815869656,
    1.9152065613971474, 1.9360617934922943, 1.9571441241754002, 1.978456026387951};

__device__ __forceinline__ double fb_exp(double x) {
  if (!(x > -708.0 && x < 709.0)) return exp(x);
  const double magic = 6755399441055744.0;  // 1.5 * 2^52: the low word of (v + magic) is rint(v)
  const double t = fma(x, 92.33248261689366, magic);
  const int k = __double2loint(t);
  const double kd = t - magic;
  double r = fma(kd, -0.01083042469326756, x);
  r = fma(kd, -2.981585464634229e-12, r);
  double p = fma(r, 0.001388888888888889, 0.008333333333333333);
  p = fma(p, r, 0.041666666666666664);
  p = fma(p, r, 0.16666666666666666);
  p = fma(p, r, 0.5);
  p = fma(p * r, r, r);
  const double tj = __ldg(&kExpTable[k & 63]);
  const double res = fma(tj, p, tj);
  return __hiloint2double(__double2hiint(res) + ((k >> 6) << 20), __double2loint(res));
}

// Exponentially-modified Gaussian without its amplitude term, routines/profile.py:90-112:
//   exp(-z^2/2) * erfcx((ratio - z)/sqrt(2))          == exp((ratio/2 - z) ratio) * erfc(.)
// evaluated so that neither factor over/underflows (the reference switches formula at
// arg <= -20; here the split is at arg = 0, exact to rounding):
//   arg >= 0       exp(-z^2/2) erfcx(arg)
//   -6 < arg < 0   2 exp(ratio (ratio/2 - z)) - exp(-z^2/2) erfcx(-arg)     [erfcx(-y) = 2e^{y^2} - erfcx(y)]
//   arg <= -6      2 exp(ratio (ratio/2 - z))                               [erfc(6)/2 = 1.1e-17 < ulp/2]
__device__ __forceinline__ double emg_shape(double z, double ratio) {
  // one exp + erfcx on |arg| for both signs (lanes of a warp straddle arg = 0 near the peak:
  // separate code per sign would execute both), the second exp only where arg < 0
  const double arg = (ratio - z) * kSqrtHalf;
  if (arg <= kTailArg) return 2.0 * fb_exp(ratio * (0.5 * ratio - z));
  const double e = fb_exp(-0.5 * z * z) * erfcx(fabs(arg));
  if (arg >= 0.0) return e;
  return 2.0 * fb_exp(ratio * (0.5 * ratio - z)) - e;
}

// First sample that is NOT on the clamp plateau (jp) and first sample in the pure exponential
// tail (jt) of one (channel, component): closed-form guess, then fixed up with the very
// predicates of eval_sample_fast so that the classification is identical to the per-sample one.
__device__ inline void vals3_boundaries(const PlanConst& pc, const CompTab& ct, const ChanTab& ch, int& jp, int& jt) {
  const int nt = pc.num_time, kt = pc.kt, n = pc.ntk;
  if (!ch.fast_ok) {
    jp = 0;
    jt = nt;
    return;
  }
  auto plateau = [&](int j) {
    return (linspace_at(ct.t_start, ct.t_step, ct.t_stop, j * kt + kt - 1, n) - ch.min_delay) < ct.neg5w;
  };
  auto tail = [&](int j) { return linspace_at(ct.t_start, ct.t_step, ct.t_stop, j * kt, n) >= ch.tail_lo; };
  const double span = ct.t_step * (double)kt;
  double g = ceil((ct.neg5w + ch.min_delay - ct.t_start) / span);
  jp = (int)fmin(fmax(g, 0.0), (double)nt);
  while (jp > 0 && !plateau(jp - 1)) --jp;
  while (jp < nt && plateau(jp)) ++jp;
  g = ceil((ch.tail_lo - ct.t_start) / span);
  jt = (int)fmin(fmax(g, 0.0), (double)nt);
  while (jt > 0 && tail(jt - 1)) --jt;
  while (jt < nt && !tail(jt)) ++jt;
}


// Near range of one (channel, component): samples whose arg_exp (first_derivatives) is >= -kNearCut,
// i.e. where the Gaussian factor G = g0 exp(arg_exp) of the derivatives is kept. The reference's
// own cut is the exp underflow (arg_exp < -746, |T| > 38.6 w: G exactly zero); outside
// |T| <= 10 w the factor is below e^-50 = 2e-22 of its peak, twelve orders of magnitude under
// the 1e-10 contract, and treating it as zero there makes three quarters of the former near
// samples far ones (a basis-block tile in k_gram3, three moments in k_hess3).
// arg_exp is a concave quadratic of the time difference, so the set is an interval; the guess
// from its roots is fixed up with the exact predicate.
constexpr double kNearCut = 50.0;
__device__ inline void near_range(const PlanConst& pc, const CompTab& ct, const ChanTab& ch, int& j_lo,
                                        int& j_hi) {
  const int nt = pc.num_time;
  if (ch.sc_time == 0.0) {  // Gaussian branch: no shortcut
    j_lo = 0;
    j_hi = nt;
    return;
  }
  const double kt_mid = 0.5 * (double)(pc.kt - 1);
  auto far = [&](int j) {
    const double td = fma(ct.t_step, (double)(j * pc.kt) + kt_mid, ct.t_start) - ch.mean_delay;
    const double arg_erf = (td - ct.w * ct.w * ch.inv_tau) * ct.inv_w * kSqrtHalf;
    const double ws = ct.w * ch.inv_tau;
    const double arg_exp = 0.5 * ws * ws - td * ch.inv_tau - arg_erf * arg_erf;
    return arg_exp < -kNearCut;
  };
  // arg_exp = -td^2 / (2 w^2): near where |td| <= w sqrt(2 kNearCut)
  const double half = ct.w * sqrt(2.0 * kNearCut);
  const double span = ct.t_step * (double)pc.kt;
  const double t0 = fma(ct.t_step, kt_mid, ct.t_start) - ch.mean_delay;  // td of sample 0
  double g = floor((-half - t0) / span);
  j_lo = (int)fmin(fmax(g, 0.0), (double)nt);
  g = ceil((half - t0) / span) + 1.0;
  j_hi = (int)fmin(fmax(g, 0.0), (double)nt);
  if (!(j_lo <= j_hi)) {  // NaN tables: no shortcut
    j_lo = 0;
    j_hi = nt;
    return;
  }
  if (j_lo == j_hi) {  // guess says "no near sample": make sure by probing the closest one
    const int jc = min(max((int)((0.0 - t0) / span), 0), nt - 1);
    if (far(jc)) return;
    j_lo = jc;
    j_hi = jc + 1;
  }
  while (j_lo > 0 && !far(j_lo - 1)) --j_lo;
  while (j_lo < j_hi && far(j_lo)) ++j_lo;
  while (j_hi < nt && !far(j_hi)) ++j_hi;
  while (j_hi > j_lo && far(j_hi - 1)) --j_hi;
}



__device__ inline void make_comp_tab(const PlanConst& pc, const double* params, int c, CompTab& ct) {
  const int nc = pc.num_comp;
  const double t_arr = params[FB_ARRIVAL_TIME * nc + c];
  const double w = params[FB_BURST_WIDTH * nc + c];
  // analysis/model.py:202-206 + routines/manipulate.py:99-105, same operation order.
  const double diff_new = pc.res_time / (double)pc.kt;
  const double first = pc.time_first - t_arr;
  const double last = pc.time_last - t_arr;
  const double lo = __dadd_rn(__dsub_rn(first, pc.res_time / 2), diff_new / 2);
  const double hi = __dsub_rn(__dadd_rn(last, pc.res_time / 2), diff_new / 2);
  ct.t_start = lo;
  ct.t_stop = hi;
  ct.t_step = (pc.ntk > 1) ? (hi - lo) / (double)(pc.ntk - 1) : 0.0;
  ct.w = w;
  ct.inv_w = 1.0 / w;
  ct.neg5w = -5 * w;
  ct.amp10 = pow(10.0, params[FB_AMPLITUDE * nc + c]);
  ct.ref_freq = params[FB_REF_FREQ * nc + c];
}

// Table entry of one (channel, component, sub-frequency a) and its plateau profile value.
__device__ inline void make_sub_tab(const PlanConst& pc, const double* params, const CompTab& ct, int i, int c, int a,
                                    double fref_b, double fc_b, SubTab& st, double& plateau_p,
                                    double& sc_time_out) {
  const int nc = pc.num_comp, kf = pc.kf, kt = pc.kt;
  const double dm = global_param(params, nc, FB_DM);
  const double beta = global_param(params, nc, FB_DM_INDEX);
  const double tau0 = global_param(params, nc, FB_SCATTERING_TIMESCALE);
  const double alpha = global_param(params, nc, FB_SCATTERING_INDEX);
  const double sp_idx = params[FB_SPECTRAL_INDEX * nc + c];
  const double sp_run = params[FB_SPECTRAL_RUNNING * nc + c];
  const double fref = ct.ref_freq;
  const double w = ct.w;
  const double f = pc.subfreqs[(size_t)i * kf + a];
  // routines/ism.py:78: dm_value * dm_const * (freq1 ** dm_idx - freq2 ** dm_idx)
  double delay = ((pc.dm_incoherent + dm) * pc.dm_const) * (pow(f, beta) - fref_b);
  delay -= (pc.dm_incoherent * pc.dm_const) * (fc_b - fref_b);
  st.delay = delay;
  const double fr = f / fref;
  const double sc_time = tau0 * pow(fr, alpha);  // analysis/model.py:336
  sc_time_out = sc_time;
  st.ratio = w / sc_time;
  st.amp_term = pc.normalize_pbf ? kSqrtHalfPi * st.ratio : tau0 / sc_time;
  const double lf = log(fr);
  st.sed = exp(sp_idx * lf + sp_run * (lf * lf));
  st.inv_tau = 1.0 / sc_time;
  st.tail_e = 0.5 * st.ratio * st.ratio + delay * st.inv_tau;
  // sum_b exp(-b step / tau) over the kt sub-times: one exp, the powers by multiplication
  // (kt <= a few dozen: the accumulated rounding stays below 1e-14 relative)
  double geom = 0.0;
  {
    const double r = exp(-ct.t_step * st.inv_tau);
    double rb = 1.0;
    for (int b = 0; b < kt; ++b) {
      geom += rb;
      rb *= r;
    }
  }
  st.tail_c = 2.0 * st.amp_term * geom / ((double)kt * (double)kf);
  if (pc.is_folded) {
    // analysis/model.py:327-331 on the delay-corrected labels of this sub-frequency row.
    const int n = pc.ntk;
    const double t0 = linspace_at(ct.t_start, ct.t_step, ct.t_stop, 0, n) - delay;
    const double t1 = linspace_at(ct.t_start, ct.t_step, ct.t_stop, 1, n) - delay;
    const double tl = linspace_at(ct.t_start, ct.t_step, ct.t_stop, n - 1, n) - delay;
    const double res = t1 - t0;
    st.fold_start = tl + res;
    st.fold_stop = tl + res * (double)n;
    st.fold_step = (n > 1) ? (st.fold_stop - st.fold_start) / (double)(n - 1) : 0.0;
  } else {
    st.fold_start = st.fold_step = st.fold_stop = 0.0;
  }
  plateau_p = st.amp_term * emg_shape(ct.neg5w * ct.inv_w, st.ratio);
}

// Sub-frequency summary of one (channel, component) accumulated in sub-frequency order.
struct SubSummary {
  double min_delay, max_delay, sum_delay, max_wr, plat_p, plat_ps;
  int scattered;
  __device__ void init() {
    min_delay = INFINITY;
    max_delay = -INFINITY;
    sum_delay = max_wr = plat_p = plat_ps = 0.0;
    scattered = 0;
  }
  __device__ void add(double delay, double ratio, double sed, double w, double p) {
    min_delay = fmin(min_delay, delay);
    max_delay = fmax(max_delay, delay);
    sum_delay += delay;
    max_wr = fmax(max_wr, w * ratio);
    plat_p += p;
    plat_ps += p * sed;
  }
};

// The ChanTab of one (channel, component) from the sub-frequency summary.
// The four powers every (channel, component) needs once: ref_freq ** dm_index, f ** dm_index,
// (f / ref_freq) ** sc_index and (f / ref_freq) ** -sc_index at the channel centre.
struct ChanPowers {
  double fref_b, fc_b, fr_alpha, fr_nalpha;
};

__device__ inline ChanPowers make_chan_powers(const PlanConst& pc, const double* params, const CompTab& ct, int i) {
  const int nc = pc.num_comp;
  const double beta = global_param(params, nc, FB_DM_INDEX), alpha = global_param(params, nc, FB_SCATTERING_INDEX);
  const double fc = pc.freqs[i], fr = fc / ct.ref_freq;
  ChanPowers cp;
  cp.fref_b = pow(ct.ref_freq, beta);
  cp.fc_b = pow(fc, beta);
  cp.fr_alpha = pow(fr, alpha);
  cp.fr_nalpha = pow(fr, -alpha);
  return cp;
}

__device__ inline void finish_chan_tab(const PlanConst& pc, const double* params, const CompTab& ct, int i, int c,
                                       const SubSummary& sm, const ChanPowers& cp, ChanTab& ch) {
  const int nc = pc.num_comp, kf = pc.kf;
  const double fc = pc.freqs[i];
  const double dm = global_param(params, nc, FB_DM);
  const double beta = global_param(params, nc, FB_DM_INDEX);
  const double tau0 = global_param(params, nc, FB_SCATTERING_TIMESCALE);
  const double sp_idx = params[FB_SPECTRAL_INDEX * nc + c];
  const double sp_run = params[FB_SPECTRAL_RUNNING * nc + c];
  const double fref = ct.ref_freq;
  const double fref_b = cp.fref_b, fc_b = cp.fc_b;
  const double w = ct.w;
  const double fr = fc / fref;
  const double lf = log(fr);
  ch.amp = exp(sp_idx * lf + sp_run * (lf * lf)) * ct.amp10;
  ch.min_delay = sm.min_delay;
  ch.mean_delay = sm.sum_delay / (double)kf;
  // pure tail for every sub-sample:  (T/w - ratio)/sqrt(2) >= 6  <=  T >= w*ratio + 6 sqrt(2) w
  ch.tail_lo = sm.max_delay + sm.max_wr + (6.0 * kSqrt2 + 1e-6) * w;
  ch.plateau_p = sm.plat_p / (double)kf;
  ch.plateau_ps = ct.amp10 * (sm.plat_ps / (double)kf);
  ch.logf = lf;
  ch.sc_time = tau0 * cp.fr_alpha;
  ch.inv_tau = 1.0 / ch.sc_time;
  double amp_pbf = cp.fr_nalpha;
  if (pc.normalize_pbf) amp_pbf *= kSqrtHalfPi * w / tau0;
  ch.amp_pbf = amp_pbf;
  ch.inv_apbf = 1.0 / amp_pbf;
  if (pc.normalize_pbf) {
    ch.dap_w = kSqrtHalfPi / ch.sc_time;
    ch.dap_alpha = -kSqrtHalfPi * lf * w / ch.sc_time;
    ch.dap_tau = -kSqrtHalfPi * w / (ch.sc_time * tau0);
    ch.d2ap_alpha = kSqrtHalfPi * lf * lf * w / ch.sc_time;
  } else {
    ch.dap_w = 0.0;
    ch.dap_tau = 0.0;
    ch.dap_alpha = -lf * cp.fr_nalpha;
    ch.d2ap_alpha = lf * lf * cp.fr_nalpha;
  }
  ch.d_dm = -pc.dm_const * (fc_b - fref_b);
  ch.q1 = log(fc) * fc_b - log(fref) * fref_b;
  ch.q2 = log(fc) * log(fc) * fc_b - log(fref) * log(fref) * fref_b;
  ch.d_dmidx = -pc.dm_const * dm * ch.q1;
  ch.fpow = fc_b;
  ch.refpow = fref_b;
  ch.ref_freq = fref;
  ch.scattered = sm.scattered;
  ch.fast_ok = sm.scattered && !pc.is_folded && isfinite(ch.tail_lo) && isfinite(ch.plateau_ps) &&
               isfinite(sm.min_delay) && w > 0.0 && tau0 > 0.0;
  ch.pad_[0] = ch.pad_[1] = 0;
  vals3_boundaries(pc, ct, ch, ch.jp, ch.jt);
  near_range(pc, ct, ch, ch.near_lo, ch.near_hi);
}

// All tables of one (channel, component): kf SubTab entries + the ChanTab. One thread.
__device__ inline void make_channel_tables(const PlanConst& pc, const double* params, const CompTab& ct,
                                           int i, int c, SubTab* sub, ChanTab& ch) {
  SubSummary sm;
  sm.init();
  const ChanPowers cp = make_chan_powers(pc, params, ct, i);
  for (int a = 0; a < pc.kf; ++a) {
    SubTab st;
    double p, sc_time;
    make_sub_tab(pc, params, ct, i, c, a, cp.fref_b, cp.fc_b, st, p, sc_time);
    sub[a] = st;
    if (sc_time > 0.0) sm.scattered = 1;  // analysis/model.py:339 any(sc_time > 0)
    sm.add(st.delay, st.ratio, st.sed, ct.w, p);
  }
  finish_chan_tab(pc, params, ct, i, c, sm, cp, ch);
}

// Per-sample, per-component result: the three cached maps of analysis/model.py:212-256.
struct SampleComp {
  double timediff;  // mean over sub-freq / sub-time of the UNCLAMPED time difference
  double timeprof;  // mean profile
  double spectrum;  // 10**A * mean(profile * sed)   (scintillation rescales later)
};

// Along the time axis a scattered component has three regimes and only the middle one needs the
// special functions (the split is exact to rounding, see emg_shape):
//   plateau  every sub-sample is clamped to -5w (analysis/model.py:340): the profile is a
//            per-(channel, component) constant, tabulated;
//   rise     general case: exp + erfcx per sub-sample;
//   tail     every sub-sample has erfc == 2 to rounding: the kt sub-samples of one sub-frequency
//            form a geometric series, one exp per (sample, sub-frequency).
// Returns true and fills `out` when sample j lies in the plateau or the tail.
__device__ __forceinline__ bool eval_sample_fast(const PlanConst& pc, const CompTab& ct, const ChanTab& ch,
                                                 const SubTab* __restrict__ sub, int j, SampleComp& out) {
  if (!ch.fast_ok) return false;
  const int kf = pc.kf, kt = pc.kt, n = pc.ntk;
  const int m0 = j * kt;
  const double u0 = linspace_at(ct.t_start, ct.t_step, ct.t_stop, m0, n);
  const double u_last = linspace_at(ct.t_start, ct.t_step, ct.t_stop, m0 + kt - 1, n);
  const bool plateau = (u_last - ch.min_delay) < ct.neg5w;
  const bool tail = u0 >= ch.tail_lo;
  if (!(plateau || tail)) return false;
  // mean over the kt sub-times of the arrival-corrected labels: start + step (m0 + (kt-1)/2)
  // (identical to summing them up to rounding; the forced last label differs from the
  // arithmetic one by an ulp of |stop|)
  out.timediff = fma(ct.t_step, (double)m0 + 0.5 * (double)(kt - 1), ct.t_start) - ch.mean_delay;
  if (plateau) {
    out.timeprof = ch.plateau_p;
    out.spectrum = ch.plateau_ps;
  } else {
    double s_p = 0.0, s_ps = 0.0;
#pragma unroll 4
    for (int a = 0; a < kf; ++a) {  // independent exps: unrolled for instruction-level parallelism
      const SubTab& st = sub[a];
      const double earg = fma(-u0, st.inv_tau, st.tail_e);
      const double p = (earg < -746.0) ? 0.0 : st.tail_c * fb_exp(earg);  // exp underflows to exactly 0
      s_p += p;
      s_ps = fma(p, st.sed, s_ps);
    }
    out.timeprof = s_p;
    out.spectrum = ct.amp10 * s_ps;
  }
  return true;
}

// One sub-sample (sub-frequency entry st, up-sampled time index m) of the general case:
// time difference t (analysis/model.py:208-209) and profile p (analysis/model.py:336-350).
__device__ __forceinline__ void eval_subsample(const PlanConst& pc, const CompTab& ct, int scattered,
                                               const SubTab& st, int m, double& t, double& p) {
  const int n = pc.ntk;
  const double u = linspace_at(ct.t_start, ct.t_step, ct.t_stop, m, n);
  t = u - st.delay;
  if (scattered) {
    const double tc = (t < ct.neg5w) ? ct.neg5w : t;  // analysis/model.py:340
    p = st.amp_term * emg_shape(tc * ct.inv_w, st.ratio);
    if (pc.is_folded) {
      double t2 = linspace_at(st.fold_start, st.fold_step, st.fold_stop, m, n);
      t2 = (t2 < ct.neg5w) ? ct.neg5w : t2;
      p = 0.5 * (p + st.amp_term * emg_shape(t2 * ct.inv_w, st.ratio));
    }
  } else {
    const double z = t * ct.inv_w;  // routines/profile.py:47
    p = fb_exp(-0.5 * z * z);
    if (pc.is_folded) {
      const double z2 = linspace_at(st.fold_start, st.fold_step, st.fold_stop, m, n) * ct.inv_w;
      p = 0.5 * (p + fb_exp(-0.5 * z2 * z2));
    }
  }
}

// General case of one (channel, sample, component) by a single thread: loops the kf x kt
// sub-grid in the reference's order (mean over sub-frequencies, then over sub-times).
__device__ __forceinline__ SampleComp eval_sample_general(const PlanConst& pc, const CompTab& ct,
                                                          const ChanTab& ch,
                                                          const SubTab* __restrict__ sub, int j) {
  const int kf = pc.kf, kt = pc.kt;
  const double inv_kf = 1.0 / (double)kf, inv_kt = 1.0 / (double)kt;
  double acc_t = 0.0, acc_p = 0.0, acc_ps = 0.0;
  for (int b = 0; b < kt; ++b) {
    double s_t = 0.0, s_p = 0.0, s_ps = 0.0;
    for (int a = 0; a < kf; ++a) {
      double t, p;
      eval_subsample(pc, ct, ch.scattered, sub[a], j * kt + b, t, p);
      s_t += t;
      s_p += p;
      s_ps += p * sub[a].sed;
    }
    acc_t += s_t * inv_kf;
    acc_p += s_p * inv_kf;
    acc_ps += s_ps * inv_kf;
  }
  SampleComp out;
  out.timediff = acc_t * inv_kt;
  out.timeprof = acc_p * inv_kt;
  out.spectrum = ct.amp10 * (acc_ps * inv_kt);
  return out;
}

__device__ __forceinline__ SampleComp eval_sample_component(const PlanConst& pc, const CompTab& ct,
                                                            const ChanTab& ch,
                                                            const SubTab* __restrict__ sub, int j) {
  SampleComp out;
  if (eval_sample_fast(pc, ct, ch, sub, j, out)) return out;
  return eval_sample_general(pc, ct, ch, sub, j);
}

// First derivatives of ONE component's spectrum at one sample, routines/derivative.py:454-936.
// m = spectrum_per_component value, td = timediff_per_component value, amp = the
// amplitude_per_component value, ch = this component's table, cho = the table of the component
// the reference passes to deriv_amplitude_pbf (the four "global" functions pass the OUTER
// component, routines/derivative.py:707,779,847,916). d[] is indexed by fb_parameter. gfac_out
// returns the shared factor amp * amp_pbf * exp(arg_exp) * 2/sqrt(pi) for the second derivatives.
__device__ __forceinline__ void first_derivatives(const ChanTab& ch, const ChanTab& cho, const CompTab& ct,
                                                  double tau0, double amp, double m, double td, double* d,
                                                  double* gfac_out, double* arg_erf_out = nullptr,
                                                  bool force_far = false) {
  const double lf = ch.logf, w = ct.w, inv_w = ct.inv_w;
  d[FB_AMPLITUDE] = kLn10 * m;
  d[FB_SPECTRAL_INDEX] = lf * m;
  d[FB_SPECTRAL_RUNNING] = (lf * lf) * m;
  const double inv_tau = ch.inv_tau;
  // shared scattered-branch factor, routines/derivative.py:14-69 (arg_exp evaluated as written).
  const double arg_erf = (td - w * w * inv_tau) * inv_w * kSqrtHalf;
  const double ws = w * inv_tau;
  const double arg_exp = 0.5 * ws * ws - td * inv_tau - arg_erf * arg_erf;
  // exp() is exactly 0 below -745.2: far from the burst (|T| > 38.6 w) skip the call
  // force_far: the G = 0 limit whatever td is (node evaluation of the far-sample polynomials)
  const double g = (force_far || arg_exp < -746.0) ? 0.0 : amp * ch.amp_pbf * fb_exp(arg_exp) * kTwoOverSqrtPi;
  *gfac_out = g;
  if (arg_erf_out) *arg_erf_out = arg_erf;
  if (ch.sc_time == 0.0) {  // per-channel Gaussian branch, routines/derivative.py:568,629,695,767
    const double inv_w2 = inv_w * inv_w;
    const double tm = td * m * inv_w2;
    d[FB_BURST_WIDTH] = td * tm * inv_w;
    d[FB_ARRIVAL_TIME] = tm;
    d[FB_DM] = -tm * ch.d_dm;
    d[FB_DM_INDEX] = -tm * ch.d_dmidx;
  } else {
    const double gs = g * inv_w * kSqrtHalf;
    d[FB_BURST_WIDTH] = m * (cho.dap_w * ch.inv_apbf + ws * inv_tau) - gs * (td * inv_w + ws);
    d[FB_ARRIVAL_TIME] = m * inv_tau - gs;
    d[FB_DM] = ch.d_dm * (gs - m * inv_tau);
    d[FB_DM_INDEX] = ch.d_dmidx * (gs - m * inv_tau);
  }
  // no Gaussian branch in the reference for these two (routines/derivative.py:800-936):
  // sc_time == 0 yields inf/nan there and here alike.
  const double inv_tau0 = 1.0 / tau0;
  const double gt = g * ws * kSqrtHalf;
  const double u = td * inv_tau - ws * ws;
  d[FB_SCATTERING_TIMESCALE] = m * (cho.dap_tau * ch.inv_apbf + u * inv_tau0) + gt * inv_tau0;
  d[FB_SCATTERING_INDEX] = m * (cho.dap_alpha * ch.inv_apbf + lf * u) + gt * lf;
}

}  // namespace fb
