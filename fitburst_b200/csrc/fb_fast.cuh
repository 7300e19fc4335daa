// fitburst_b200 fast evaluation: warp-granular kernels for the common case of the fit
// (scattered branch, not folded, no per-channel scintillation amplitudes).
//
// The block-granular kernels of fb_kernels.cuh treat every sample alike and pay for their
// generality in instruction issue (ncu, config 2: k_vals 5.4e8 warp instructions of which ~40 %
// are the exp of the exponential tail and ~35 % loop / queue / addressing overhead; k_gram2_mma
// 2.5e8 of which ~20 % is the run-time column scatter). Here
//
//   k_vals3   a block owns one channel at a time, its warps take 32-sample tiles from a shared
//             counter. Per (tile, component) the three regimes of the scattered profile
//             (fb_core.cuh: clamp plateau / rise / exponential tail) are decided WARP-wide:
//             * a tile wholly in the tail needs no exp per sample at all: with
//               u0(j0 + lane) = u0(j0) + lane kt step the kf geometric series factorise into
//               B[a] = exp(tail_e[a] - u0(j0) / tau[a])        (one exp per lane per tile, all
//                                                                components at once)
//               R[a][lane] = 10^A sed[a] tail_c[a] exp(-lane kt step / tau[a])   (per channel)
//               and the spectrum value is the kf-term dot product sum_a B[a] R[a][lane];
//             * a tile wholly on the plateau stores a table constant;
//             * in the one or two mixed tiles per (channel, component) the samples that need
//               the general exp + erfcx sub-grid are evaluated one after the other by the whole
//               warp (one sub-sample per lane, shuffle reduction).
//             Only the spectrum value goes to global memory: the mean time difference is a
//             closed form of the sample index (fb_core.cuh: eval_sample_fast) and is recomputed
//             by its consumers.
//   k_gram3   derivative side on the FP64 tensor cores over the REDUCED column set. The
//             amplitude / spectral_index / spectral_running columns of a component are the
//             spectrum times per-channel factors (ln 10, L, L^2 with L = ln(f / ref_freq),
//             routines/derivative.py:454-530,880-936), and the weight is per channel too
//             (analysis/fitter.py:506-542), so per sample only [S_c, dt_c, dw_c, globals, r] are
//             staged (14 instead of 18 columns at config 2: 2 instead of 3 column blocks, 24
//             instead of 48 DMMA per 32 samples); at the end of a channel the C fragments are
//             folded into the accumulators of w^2 L^p (p = 0..4) and cleared, and
//             k_gram3_finish rebuilds the full [J r]^T [J r].
#pragma once
#include "fb_kernels.cuh"

namespace fb {

#ifndef FB_FAST_BLOCK
#define FB_FAST_BLOCK 128  // measured: 128-thread blocks beat 256 (0.71 -> 0.60 ms per evaluation) and 64
#endif
#ifndef FB_V3_MINB
#define FB_V3_MINB 5
#endif
#ifndef FB_G3_MINB
#define FB_G3_MINB 4
#endif
#ifndef FB_H3_MINB
#define FB_H3_MINB 4
#endif
constexpr int kFastBlock = FB_FAST_BLOCK;
constexpr int kFastWarps = kFastBlock / 32;
constexpr unsigned kFullMask = 0xffffffffu;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned saddr = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(gmem_src) : "memory");
}

// ---- bulk asynchronous copies (TMA unit, 1-D) completing on an mbarrier ----------------------------
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
// orders this thread's earlier generic-proxy accesses to shared memory before later async-proxy ones
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}

// bytes: multiple of 16; dst / src 16-byte aligned
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)),
               "l"(src), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}

// Waits for the phase with the given parity; traps instead of hanging if it never completes.
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned addr = smem_addr(bar);
  unsigned done = 0;
  for (int spins = 0; !done; ++spins) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (!done && spins > (1 << 24)) __trap();
  }
}

// ---- model side ---------------------------------------------------------------------------
constexpr int kValsQueue = 256;  // rise-region (sample, component) items queued per warp and channel
#ifndef FB_V3_SPL
#define FB_V3_SPL 2  // measured: 2 (64-sample tiles) 0.454 ms per evaluation, 4 (128-sample tiles, two per warp) 0.510
#endif
constexpr int kSpl = FB_V3_SPL;    // consecutive samples per lane in k_vals3 (even)
constexpr int kVTile = 32 * kSpl;  // samples per warp tile
static_assert(kSpl % 2 == 0 && kSpl >= 2 && kSpl <= 8, "k_vals3 handles samples in 16-byte pairs");

#ifndef FB_TAY_BOUND
#define FB_TAY_BOUND 5e-13  // largest truncation-error bound at which the sub-frequency expansion is used
#endif
// Per (channel, component) constants of the sub-frequency expansion (see make_tay_tab).
struct TayTab {
  double d_bar, rho_bar, dz_max;
  double c0, c1z, c1r, c2zz, c2zr, c2rr, c3zzz, c3zzr, c3zrr, c3rrr;
  int ok, ser_ok;  // moment expansion usable / per-sub-frequency series usable (series_eval)
};


__host__ __device__ inline int vals3_bw_stride(int nck) { return (nck + 31) & ~31; }

// chi^2 variant of k_vals3: residuals of samples that wait for a queued rise-region value
constexpr int kChiSlots = 128;  // per warp and channel (= its queue length in that variant)

__host__ __device__ inline size_t vals3_smem_bytes(int nc, int kf, bool chi2 = false) {
  size_t n = 0;
  if (chi2)  // + 16: the slot block is re-aligned behind the expansion tables
    n += sizeof(double) * (size_t)kFastWarps * kChiSlots * (1 + nc) + sizeof(int) * kFastWarps * (kChiSlots + kVTile) + 16;
  n += sizeof(CompTab) * nc;
  n += 2 * (sizeof(ChanTab) * nc + sizeof(SubTab) * nc * kf);  // double-buffered channel tables
  n += sizeof(double) * (size_t)nc * kf * (kVTile + 1);       // R (sample offsets of a tile), D
  n += sizeof(double) * (size_t)kFastWarps * vals3_bw_stride(nc * kf);  // B per warp
  n += sizeof(int) * (kFastWarps * kValsQueue + 4 + 2 * FB_MAX_COMPONENTS);  // per-warp queues, regime boundaries
  n += sizeof(TayTab) * nc + 16;
  return n;
}

// General case of one (sample jj, component): the whole warp evaluates the kf x kt sub-grid, one
// sub-sample per lane (analysis/model.py:208-247); every lane returns the spectrum value.
template <int KF>
__device__ __forceinline__ double vals3_general_warp(const PlanConst& pc, const CompTab& ct, const ChanTab& ch,
                                                     const SubTab* __restrict__ sb, int jj, int lane) {
  const int kf = KF > 0 ? KF : pc.kf, kt = pc.kt;
  const int npe = kf * kt;
  double s_ps = 0.0;
  for (int pe = lane; pe < npe; pe += 32) {
    const int b = pe / kf, a = pe - b * kf;
    double t, p;
    eval_subsample(pc, ct, ch.scattered, sb[a], jj * kt + b, t, p);
    s_ps = fma(p, sb[a].sed, s_ps);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s_ps += __shfl_xor_sync(kFullMask, s_ps, off);
  return ct.amp10 * (s_ps * (1.0 / (double)npe));
}

// Rise-region samples, sub-frequency direction by expansion. Within one channel the kf
// sub-frequencies of a component differ only slightly in delay (d_a) and in ratio = w / tau_a
// (rho_a), so the sum over the sub-frequencies of one sub-time
//     sum_a c_a E((T - d_a) / w, rho_a),   E(z, r) = exp(-z^2 / 2) erfcx((r - z) / sqrt 2),
// c_a = amp_term_a sed_a, is expanded to third order about the channel means: ONE exp + erfcx per
// sub-time instead of kf, all derivatives of E follow from the same two values
// (erfcx' = 2 u erfcx - 2/sqrt(pi)). The expansion is used only where its truncation error is
// bounded below 5e-13 relative (fourth-order remainder with |d^n E / E| <= (rho + 8.5)^n on the
// rise region; measured worst case 1.3e-13) and where no sub-frequency of the sub-time is clamped
// (analysis/model.py:340); everything else takes the exact sub-grid.
__device__ inline void make_tay_tab(const PlanConst& pc, const CompTab& ct, const ChanTab& ch,
                                    const SubTab* __restrict__ sb, int kf, TayTab& t) {
  double d_bar = 0.0, rho_bar = 0.0;
  for (int a = 0; a < kf; ++a) {
    d_bar += sb[a].delay;
    rho_bar += sb[a].ratio;
  }
  d_bar /= (double)kf;
  rho_bar /= (double)kf;
  double c0 = 0.0, c1z = 0.0, c1r = 0.0, c2zz = 0.0, c2zr = 0.0, c2rr = 0.0, dz_max = 0.0, dr_max = 0.0;
  double c3zzz = 0.0, c3zzr = 0.0, c3zrr = 0.0, c3rrr = 0.0;
  for (int a = 0; a < kf; ++a) {
    const double c = sb[a].amp_term * sb[a].sed;
    const double dz = -(sb[a].delay - d_bar) * ct.inv_w, dr = sb[a].ratio - rho_bar;
    c0 += c;
    c1z = fma(c, dz, c1z);
    c1r = fma(c, dr, c1r);
    c2zz = fma(c * dz, dz, c2zz);
    c2zr = fma(c * dz, dr, c2zr);
    c2rr = fma(c * dr, dr, c2rr);
    c3zzz = fma(c * dz * dz, dz, c3zzz);
    c3zzr = fma(c * dz * dz, dr, c3zzr);
    c3zrr = fma(c * dz * dr, dr, c3zrr);
    c3rrr = fma(c * dr * dr, dr, c3rrr);
    dz_max = fmax(dz_max, fabs(dz));
    dr_max = fmax(dr_max, fabs(dr));
  }
  t.d_bar = d_bar;
  t.rho_bar = rho_bar;
  t.dz_max = dz_max;
  t.c0 = c0;
  t.c1z = c1z;
  t.c1r = c1r;
  t.c2zz = 0.5 * c2zz;
  t.c2zr = c2zr;
  t.c2rr = 0.5 * c2rr;
  t.c3zzz = c3zzz * (1.0 / 6.0);
  t.c3zzr = 0.5 * c3zzr;
  t.c3zrr = 0.5 * c3zrr;
  t.c3rrr = c3rrr * (1.0 / 6.0);
  const double s = fabs(rho_bar) + 8.5, d = dz_max + dr_max;
  const double bound = (s * s) * (s * s) * (1.0 / 24.0) * (d * d) * (d * d);
  t.ok = ch.fast_ok && !pc.is_folded && bound < FB_TAY_BOUND && rho_bar > 0.0;
  // series_eval: order-7 remainder of erfcx about the mean argument, |erfcx^(8) / erfcx| <= 12.75^8
  // for arguments >= -6.1 (the rise region ends at z = rho + 6 sqrt 2); 15 leaves a margin
  const double sd = 15.0 * kSqrtHalf * d, sd2 = sd * sd, sd4 = sd2 * sd2;
  t.ser_ok = ch.fast_ok && !pc.is_folded && rho_bar > 0.0 && sd4 * sd4 * (1.0 / 40320.0) < FB_TAY_BOUND;
#ifdef FB_NO_SERIES  // comparison builds (scripts/build_variant.sh)
  t.ser_ok = 0;
#endif
}

// sum_a c_a E(z - dz_a, rho_a) by the expansion about (z, rho_bar); z = (T - d_bar) / w unclamped.
__device__ __forceinline__ double tay_eval(const TayTab& t, double z) {
  const double h = kSqrtHalf;
  const double u = (t.rho_bar - z) * h;
  const double g = fb_exp(-0.5 * z * z);
  const double x0 = erfcx(u);
  const double x1 = fma(2.0 * u, x0, -kTwoOverSqrtPi);
  const double x2 = 2.0 * fma(u, x1, x0);
  const double x3 = fma(2.0 * u, x2, 4.0 * x1);
  const double e_z = -z * x0 - h * x1;
  const double e_r = h * x1;
  const double e_zz = fma(z * z - 1.0, x0, fma(2.0 * z * h, x1, 0.5 * x2));
  const double e_zr = -z * h * x1 - 0.5 * x2;
  const double e_rr = 0.5 * x2;
  const double e_rrr = 0.5 * h * x3;
  const double e_zzz = fma(3.0 * z - z * z * z, x0, fma(h * (3.0 - 3.0 * z * z), x1, -1.5 * z * x2 - 0.5 * h * x3));
  const double e_zzr = h * fma(z * z - 1.0, x1, fma(2.0 * z * h, x2, 0.5 * x3));
  const double e_zrr = -0.5 * z * x2 - 0.5 * h * x3;
  double s = t.c0 * x0;
  s = fma(t.c1z, e_z, s);
  s = fma(t.c1r, e_r, s);
  s = fma(t.c2zz, e_zz, s);
  s = fma(t.c2zr, e_zr, s);
  s = fma(t.c2rr, e_rr, s);
  s = fma(t.c3zzz, e_zzz, s);
  s = fma(t.c3zzr, e_zzr, s);
  s = fma(t.c3zrr, e_zrr, s);
  s = fma(t.c3rrr, e_rrr, s);
  return g * s;
}

// Spectrum values of component c for the kVTile samples of one warp tile, kSpl CONSECUTIVE samples
// per lane (j0 + kSpl lane + s: 16-byte table reads and stores, kSpl independent FMA chains per
// lane in the tail dot product). Samples in the rise region are pushed on the warp's queue
// (deferred[s] = true for that lane) unless the queue is full or the sub-grid is too small to
// share, in which case they are done here.
template <int KF>
__device__ __forceinline__ void vals3_component_tile(const PlanConst& pc, const CompTab& ct, const ChanTab& ch,
                                                     const SubTab* __restrict__ sb, const double* __restrict__ r_lane,
                                                     const double* __restrict__ bw, int c, int jp, int jt, int j0,
                                                     int lane, int nt, bool tail_dot_ok, int* __restrict__ queue,
                                                     int& qcount, double (&out)[kSpl], bool (&deferred)[kSpl],
                                                     int qcap = kValsQueue) {
  const int kf = KF > 0 ? KF : pc.kf, kt = pc.kt;
#pragma unroll
  for (int s = 0; s < kSpl; ++s) deferred[s] = false;
  if (j0 >= jt && j0 >= jp) {  // whole tile in the exponential tail
#pragma unroll
    for (int s = 0; s < kSpl; ++s) out[s] = 0.0;
#pragma unroll
    for (int a = 0; a < kf; ++a) {
      const double b = bw[a];
#pragma unroll
      for (int h = 0; h < kSpl / 2; ++h) {
        const double2 r = *reinterpret_cast<const double2*>(r_lane + a * kVTile + 2 * h);
        out[2 * h] = fma(b, r.x, out[2 * h]);
        out[2 * h + 1] = fma(b, r.y, out[2 * h + 1]);
      }
    }
    return;
  }
  if (j0 + kVTile <= jp) {  // whole tile on the clamp plateau
#pragma unroll
    for (int s = 0; s < kSpl; ++s) out[s] = ch.plateau_ps;
    return;
  }
#pragma unroll
  for (int s = 0; s < kSpl; ++s) {
    const int j = j0 + kSpl * lane + s;
    const bool plateau = j < jp;
    const bool tail = !plateau && j >= jt;
    const bool general = j < nt && !plateau && !tail;
    const unsigned gm = __ballot_sync(kFullMask, general);
    double val = 0.0;
    if (plateau) {
      val = ch.plateau_ps;
    } else if (tail && tail_dot_ok) {
      // tail samples of a mixed tile: the same dot product as in whole-tail tiles (every base
      // factor of the component at the tile start is finite: the tail begins inside this tile)
      double s0 = 0.0, s1 = 0.0;
      int a = 0;
#pragma unroll
      for (; a + 1 < kf; a += 2) {
        s0 = fma(bw[a], r_lane[a * kVTile + s], s0);
        s1 = fma(bw[a + 1], r_lane[(a + 1) * kVTile + s], s1);
      }
      for (; a < kf; ++a) s0 = fma(bw[a], r_lane[a * kVTile + s], s0);
      val = s0 + s1;
    } else if (tail) {
      const double u0 = linspace_at(ct.t_start, ct.t_step, ct.t_stop, j * kt, pc.ntk);
      double s_ps = 0.0;
#pragma unroll 4
      for (int a = 0; a < kf; ++a) {
        const SubTab& st = sb[a];
        const double earg = fma(-u0, st.inv_tau, st.tail_e);
        const double p = (earg < -746.0) ? 0.0 : st.tail_c * fb_exp(earg);
        s_ps = fma(p, st.sed, s_ps);
      }
      val = ct.amp10 * s_ps;
    }
    if (gm != 0u) {
      if (kf * kt < 16) {
        if (general) val = eval_sample_general(pc, ct, ch, sb, j).spectrum;
      } else {
        // the queue is private to the warp (its own tiles feed it, it alone drains it after
        // its tiles): the counter is a warp-uniform register, a request that does not fit is
        // refused whole and its samples are evaluated in place
        const int cnt = __popc(gm);
        if (ch.fast_ok && qcount + cnt <= qcap) {
          if (general) queue[qcount + __popc(gm & ((1u << lane) - 1u))] = (c << 27) | j;
          qcount += cnt;
          deferred[s] = general;
        } else {
          unsigned rem = gm;
          while (rem) {
            const int src = __ffs(rem) - 1;
            rem &= rem - 1;
            const double v = vals3_general_warp<KF>(pc, ct, ch, sb, j0 + kSpl * src + s, lane);
            if (lane == src) val = v;
          }
        }
      }
    }
    out[s] = val;
  }
}

// The same sum where the sub-frequencies of a channel are too far apart for the third-order
// moments (narrow scattering: rho = w / tau large; DM offsets of a grid sweep at the bottom of the
// band): E depends on the sub-frequency only through the Gaussian factor -- exact ratio
// exp(z dz_a - dz_a^2 / 2) to the mean one -- and through the erfcx argument u_a = u + du_a with
// du_a = (d rho_a + dz_a) / sqrt 2 independent of the sample, so ONE erfcx per sub-time and its
// derivatives (x_{n+1} = 2 u x_n + 2 n x_{n-1}) give every erfcx(u_a) by a Taylor polynomial of
// order 7; one short exp per sub-frequency remains. Measured worst relative error 1.5e-13 over
// rho in [0.01, 300] at the largest admitted spread (the conditioning of exp(-z^2/2) itself).
template <int KF>
__device__ __forceinline__ double series_eval(const PlanConst& pc, const TayTab& t, const CompTab& ct,
                                              const SubTab* __restrict__ sb, double z) {
  const int kf = KF > 0 ? KF : pc.kf;
  const double h = kSqrtHalf;
  const double u = (t.rho_bar - z) * h, u2 = 2.0 * u;
  const double x0 = erfcx(u);
  const double x1 = fma(u2, x0, -kTwoOverSqrtPi);
  const double x2 = fma(u2, x1, 2.0 * x0);
  const double x3 = fma(u2, x2, 4.0 * x1);
  const double x4 = fma(u2, x3, 6.0 * x2);
  const double x5 = fma(u2, x4, 8.0 * x3);
  const double x6 = fma(u2, x5, 10.0 * x4);
  const double x7 = fma(u2, x6, 12.0 * x5);
  const double q2 = 0.5 * x2, q3 = x3 * (1.0 / 6.0), q4 = x4 * (1.0 / 24.0), q5 = x5 * (1.0 / 120.0),
               q6 = x6 * (1.0 / 720.0), q7 = x7 * (1.0 / 5040.0);
  double s = 0.0;
#pragma unroll
  for (int a = 0; a < kf; ++a) {
    const double dz = (sb[a].delay - t.d_bar) * ct.inv_w;  // z_a = z - dz
    const double du = ((sb[a].ratio - t.rho_bar) + dz) * h;
    double q = fma(q7, du, q6);
    q = fma(q, du, q5);
    q = fma(q, du, q4);
    q = fma(q, du, q3);
    q = fma(q, du, q2);
    q = fma(q, du, x1);
    q = fma(q, du, x0);
    const double e = fb_exp(dz * fma(-0.5, dz, z));
    s = fma((sb[a].amp_term * sb[a].sed) * e, q, s);
  }
  return fb_exp(-0.5 * z * z) * s;
}

// Stores the spectrum value of every (good channel, component, sample):
// gvals[(chi * nc + c) * num_time + j].
// NC > 0: the number of components is a compile-time constant (the component loops unroll and
// their address arithmetic is hoisted: the source-line profile put a fifth of the stall samples on
// per-(tile, component) pointer and index arithmetic); NC = 0: run-time value.
//
// CHI2 = true (grid sweeps, BASELINE.json config 4): nothing is stored; the kernel sums the
// weighted squared residuals sum_j (w_i (d_ij - sum_c v_c))^2 (analysis/fitter.py:231-266) of its
// channels instead, one partial per (block, warp) in A.chi2_partials[blockIdx.x * kFastWarps + warp]
// (the channels of a block are fixed by the grid size: deterministic).
// A sample with a queued rise-region component parks its partial residual in a slot of the warp
// until phase B has the missing values (components subtracted in order: deterministic).
template <int KF, int NC = 0, bool CHI2 = false>
__global__ void __launch_bounds__(kFastBlock, FB_V3_MINB) k_vals3(const EvalArgs A, double* __restrict__ gvals) {
  extern __shared__ __align__(16) unsigned char fast_smem[];
  const PlanConst& pc = A.pc;
  if (fit_stopped(pc)) return;
  const int nc = NC > 0 ? NC : pc.num_comp, kf = KF > 0 ? KF : pc.kf, kt = pc.kt, nt = pc.num_time;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nck = nc * kf;
  char* sp = reinterpret_cast<char*>(fast_smem);
  CompTab* comp = reinterpret_cast<CompTab*>(sp);
  sp += sizeof(CompTab) * nc;
  const size_t tab_bytes = sizeof(ChanTab) * nc + sizeof(SubTab) * nck;
  char* tabs = sp;  // [2][ChanTab nc | SubTab nc*kf]
  sp += 2 * tab_bytes;
  double* R = reinterpret_cast<double*>(sp);  // [nck][kVTile]: tail factors at the sample offsets of a tile
  sp += sizeof(double) * (size_t)nck * kVTile;
  double* D = reinterpret_cast<double*>(sp);  // decay between two tiles of a warp
  sp += sizeof(double) * (size_t)nck;
  const int bws = vals3_bw_stride(nck);
  double* bw = reinterpret_cast<double*>(sp) + (size_t)warp * bws;
  sp += sizeof(double) * (size_t)kFastWarps * bws;
  int* queue = reinterpret_cast<int*>(sp) + warp * kValsQueue;  // this warp's rise-region items
  int* bounds = reinterpret_cast<int*>(sp) + kFastWarps * kValsQueue + 1;  // [nc][2]: jp, jt
  TayTab* tay = reinterpret_cast<TayTab*>((reinterpret_cast<uintptr_t>(bounds + 2 * FB_MAX_COMPONENTS + 2) + 15) & ~(uintptr_t)15);
  // CHI2: per-warp slots behind everything else
  double* slot_r = nullptr;  // [kChiSlots] partial residual d - sum of the values known in phase A
  double* slot_v = nullptr;  // [kChiSlots][nc] queued values (0 where the component was not queued)
  int* qslot = nullptr;      // [kChiSlots] slot of queue item q
  int* tile_slot = nullptr;  // [kVTile] slot of a sample of the current tile
  if constexpr (CHI2) {
    double* base = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(tay + nc) + 15) & ~(uintptr_t)15);
    slot_r = base + (size_t)warp * kChiSlots * (1 + nc);
    slot_v = slot_r + kChiSlots;
    int* ibase = reinterpret_cast<int*>(base + (size_t)kFastWarps * kChiSlots * (1 + nc));
    qslot = ibase + warp * (kChiSlots + kVTile);
    tile_slot = qslot + kChiSlots;
  }
  constexpr int kQueueCap = CHI2 ? kChiSlots : kValsQueue;

  {
    const double* src = reinterpret_cast<const double*>(A.g_comp);
    double* dst = reinterpret_cast<double*>(comp);
    for (int e = tid; e < nc * (int)(sizeof(CompTab) / sizeof(double)); e += kFastBlock) dst[e] = src[e];
  }
  const int wtiles = (nt + kVTile - 1) / kVTile;
  const bool pair_store = (nt & 1) == 0;  // rows start on 16-byte boundaries
  const unsigned kf_mask = kf >= 32 ? kFullMask : ((1u << kf) - 1u);
  // tiles (kVTile samples) of a warp: wt = warp + k * kFastWarps (interleaved: the few expensive tiles around the
  // burst spread over the warps; with contiguous ranges one or two warps held the block)
  const int chan_chunks = (int)(sizeof(ChanTab) * nc / 16), sub_chunks = (int)(sizeof(SubTab) * nck / 16);
  // channel numbers are read two iterations ahead: chan_list[] -> table address is a dependent
  // global-load chain that would otherwise sit in front of every table prefetch
  auto channel_of = [&](int chi) { return chi < A.num_chan ? (A.chan_list ? __ldg(A.chan_list + chi) : chi) : -1; };
  auto prefetch_tables = [&](int i, int buf) {
    if (i >= 0) {
      char* dst = tabs + (size_t)buf * tab_bytes;
      const char* src_c = reinterpret_cast<const char*>(A.g_chan + (size_t)i * nc);
      const char* src_s = reinterpret_cast<const char*>(A.g_sub + (size_t)i * nck);
      for (int e = tid; e < chan_chunks; e += kFastBlock) cp_async16(dst + 16 * e, src_c + 16 * e);
      for (int e = tid; e < sub_chunks; e += kFastBlock)
        cp_async16(dst + sizeof(ChanTab) * nc + 16 * e, src_s + 16 * e);
    }
    cp_async_commit();
  };
  int i_next = channel_of(blockIdx.x + gridDim.x);
  prefetch_tables(channel_of(blockIdx.x), 0);
  int buf = 0;
  double acc = 0.0;  // CHI2: this lane's sum of squared weighted residuals over the block's channels
  for (int chi = blockIdx.x; chi < A.num_chan; chi += gridDim.x, buf ^= 1) {
    const int i_after = channel_of(chi + 2 * gridDim.x);
    cp_async_wait<0>();
    __syncthreads();  // tables of this channel landed; every warp is done with the previous one
    prefetch_tables(i_next, buf ^ 1);
    i_next = i_after;
    const ChanTab* chan = reinterpret_cast<const ChanTab*>(tabs + (size_t)buf * tab_bytes);
    const SubTab* sub = reinterpret_cast<const SubTab*>(tabs + (size_t)buf * tab_bytes + sizeof(ChanTab) * nc);
    int qcount = 0;
    if (tid >= 32 && tid < 32 + nc) {
      bounds[2 * (tid - 32)] = chan[tid - 32].jp;  // precomputed by the table kernel (finish_chan_tab)
      bounds[2 * (tid - 32) + 1] = chan[tid - 32].jt;
    }
    if (tid >= 64 && tid < 64 + nc) make_tay_tab(pc, comp[tid - 64], chan[tid - 64], sub + (tid - 64) * kf, kf, tay[tid - 64]);
    for (int e = tid; e < nck * 32; e += kFastBlock) {  // a warp covers one (component, sub-frequency) row
      const int l = e & 31, ca = e >> 5, c = ca / kf;
      const SubTab& st = sub[ca];
      const double decay = fb_exp(-((double)(l * kt) * comp[c].t_step) * st.inv_tau);
      // offsets 32 m + l: decay(l + 32) = decay(l) decay(31) decay(1), no further exp
      const double d32 = __shfl_sync(kFullMask, decay, 31) * __shfl_sync(kFullMask, decay, 1);
      const double base = (comp[c].amp10 * st.tail_c) * st.sed;
      double dl = decay;
#pragma unroll
      for (int m = 0; m < kSpl; ++m) {
        R[ca * kVTile + 32 * m + l] = base * dl;
        dl *= d32;
      }
    }
    for (int e = tid; e < nck; e += kFastBlock)
      D[e] = fb_exp(-((double)(kVTile * kFastWarps * kt) * comp[e / kf].t_step) * sub[e].inv_tau);
    __syncthreads();
    double* out = CHI2 ? nullptr : gvals + (size_t)chi * nc * nt;
    const int i_chan = CHI2 ? (A.chan_list ? __ldg(A.chan_list + chi) : chi) : 0;
    const double* drow = CHI2 ? A.data + (size_t)i_chan * nt : nullptr;
    const double wi = CHI2 ? __ldg(A.weights + i_chan) : 0.0;
    int scount = 0;    // CHI2: slots in use (warp-uniform)
    // phase A: closed-form regimes tile by tile (uniform cost), rise-region samples queued
    for (int wt = warp, k = 0; wt < wtiles; wt += kFastWarps, ++k) {
      const int j0 = wt * kVTile, j = j0 + kSpl * lane;
      // tail base factors B of every (component, sub-frequency) at the first sample of the tile:
      // one exp per lane every 8th tile of the warp, the decay recurrence (over kFastWarps tiles)
      // in between (a factor that overflowed ahead of the burst is recomputed: inf * decay would
      // stay inf)
      const bool refresh = (k & 7) == 0;
      bool b_finite = true;
      for (int e = lane; e < nck; e += 32) {
        const double prev = bw[e];
        double next;
        if (refresh || !(prev < 1e290)) {
          const CompTab& ct = comp[e / kf];
          const double u0b = linspace_at(ct.t_start, ct.t_step, ct.t_stop, j0 * kt, pc.ntk);
          const double earg = fma(-u0b, sub[e].inv_tau, sub[e].tail_e);
          next = (earg < -746.0) ? 0.0 : fb_exp(earg);
        } else {
          next = prev * D[e];
        }
        bw[e] = next;
        b_finite = b_finite && next < 1e290;
      }
      // bit e set: base factor e is finite (usable for tail samples of a mixed tile); only
      // meaningful when the nc * kf factors fit one pass of the warp
      const unsigned bmask = nck <= 32 ? __ballot_sync(kFullMask, b_finite) : 0u;
      __syncwarp();
      double dval[kSpl], msum[kSpl];
      bool parked[kSpl];
      const int q_before = qcount;
      if constexpr (CHI2) {
#pragma unroll
        for (int h = 0; h < kSpl / 2; ++h) {
          if (pair_store && j + 2 * h + 1 < nt) {
            const double2 d2 = *reinterpret_cast<const double2*>(drow + j + 2 * h);
            dval[2 * h] = d2.x;
            dval[2 * h + 1] = d2.y;
          } else {
            dval[2 * h] = j + 2 * h < nt ? drow[j + 2 * h] : 0.0;
            dval[2 * h + 1] = j + 2 * h + 1 < nt ? drow[j + 2 * h + 1] : 0.0;
          }
        }
#pragma unroll
        for (int s = 0; s < kSpl; ++s) {
          msum[s] = 0.0;
          parked[s] = false;
        }
      }
#pragma unroll
      for (int c = 0; c < nc; ++c) {
        bool deferred[kSpl];
        double v[kSpl];
        vals3_component_tile<KF>(pc, comp[c], chan[c], sub + c * kf, R + (size_t)c * kf * kVTile + kSpl * lane,
                                 bw + c * kf, c, bounds[2 * c], bounds[2 * c + 1], j0, lane, nt,
                                 nck <= 32 && ((bmask >> (c * kf)) & kf_mask) == kf_mask, queue, qcount, v, deferred,
                                 kQueueCap);
        if constexpr (CHI2) {
#pragma unroll
          for (int s = 0; s < kSpl; ++s) {
            if (!deferred[s]) msum[s] += v[s];
            parked[s] = parked[s] || deferred[s];
          }
          continue;
        }
        double* dst = out + (size_t)c * nt + j;
#pragma unroll
        for (int h = 0; h < kSpl / 2; ++h) {
          if (pair_store && j + 2 * h + 1 < nt && !deferred[2 * h] && !deferred[2 * h + 1]) {
            *reinterpret_cast<double2*>(dst + 2 * h) = make_double2(v[2 * h], v[2 * h + 1]);
          } else {
            if (j + 2 * h < nt && !deferred[2 * h]) dst[2 * h] = v[2 * h];
            if (j + 2 * h + 1 < nt && !deferred[2 * h + 1]) dst[2 * h + 1] = v[2 * h + 1];
          }
        }
      }
      if constexpr (CHI2) {
        // samples without queued components are finished here; the others take a slot
#pragma unroll
        for (int s = 0; s < kSpl; ++s) {
          const double r = dval[s] - msum[s];
          const unsigned pm = __ballot_sync(kFullMask, parked[s]);
          if (pm != 0u) {
            if (parked[s]) {
              const int slot = scount + __popc(pm & ((1u << lane) - 1u));
              slot_r[slot] = r;
              for (int c = 0; c < nc; ++c) slot_v[slot * nc + c] = 0.0;
              tile_slot[kSpl * lane + s] = slot;
            }
            scount += __popc(pm);
          }
          if (!parked[s] && j + s < nt) {
            const double rw = r * wi;
            acc = fma(rw, rw, acc);
          }
        }
        if (qcount > q_before) {  // items queued by this tile learn their slots
          __syncwarp();
          for (int q = q_before + lane; q < qcount; q += 32) qslot[q] = tile_slot[(queue[q] & ((1 << 27) - 1)) - j0];
        }
      }
      __syncwarp();
    }
    __syncwarp();
    // phase B: the rise-region samples this warp queued. A group of lanes (one per sub-time)
    // evaluates an item by the sub-frequency expansion; items it does not cover (a clamped
    // sub-time, a component whose expansion bound is too large) take the exact sub-grid, one per
    // warp pass. No block barrier between the phases: the queue is the warp's own.
    const int qn = qcount;
    int lpi = 1;  // lanes per item: power of two >= kt
    while (lpi < kt && lpi < 32) lpi <<= 1;
    const int ipw = 32 / lpi;
    const double inv_n = 1.0 / (double)(kf * kt);
    for (int base = 0; base < qn; base += ipw) {
      const int q = base + lane / lpi, b = lane % lpi;
      bool have = q < qn, expanded = false;
      int c = 0, jj = 0;
      double v = 0.0;
      if (have) {
        const int item = queue[q];
        c = item >> 27;
        jj = item & ((1 << 27) - 1);
        const TayTab& t = tay[c];
        expanded = (t.ok || t.ser_ok) && kt <= 32;
        if (expanded && b < kt) {
          const CompTab& ct = comp[c];
          const double z = (linspace_at(ct.t_start, ct.t_step, ct.t_stop, jj * kt + b, pc.ntk) - t.d_bar) * ct.inv_w;
          if (z - t.dz_max >= -5.0)  // no sub-frequency clamped at -5 w
            v = t.ok ? tay_eval(t, z) : series_eval<KF>(pc, t, ct, sub + c * kf, z);
          else expanded = false;
        }
      }
      // an item is expanded only if every one of its sub-times is
      const unsigned fail = __ballot_sync(kFullMask, have && !expanded);
      for (int off = lpi >> 1; off > 0; off >>= 1) v += __shfl_xor_sync(kFullMask, v, off);
      const unsigned group = (lpi == 32 ? kFullMask : ((1u << lpi) - 1u)) << (lane / lpi * lpi);
      const bool item_fail = (fail & group) != 0u;
      if (have && !item_fail && b == 0) {
        const double value = comp[c].amp10 * (v * inv_n);
        if constexpr (CHI2) slot_v[qslot[q] * nc + c] = value;
        else out[(size_t)c * nt + jj] = value;
      }
      // exact sub-grid for the rest, whole warp per item
      unsigned todo = __ballot_sync(kFullMask, have && item_fail && b == 0);
      while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const int c2 = __shfl_sync(kFullMask, c, src), j2 = __shfl_sync(kFullMask, jj, src);
        const double v2 = vals3_general_warp<KF>(pc, comp[c2], chan[c2], sub + c2 * kf, j2, lane);
        if constexpr (CHI2) {
          const int q2 = __shfl_sync(kFullMask, q, src);
          if (lane == 0) slot_v[qslot[q2] * nc + c2] = v2;
        } else {
          if (lane == 0) out[(size_t)c2 * nt + j2] = v2;
        }
      }
    }
    if constexpr (CHI2) {
      // phase C: parked residuals, queued components subtracted in component order
      __syncwarp();
      for (int slot = lane; slot < scount; slot += 32) {
        double r = slot_r[slot];
        for (int c = 0; c < nc; ++c) r -= slot_v[slot * nc + c];
        const double rw = r * wi;
        acc = fma(rw, rw, acc);
      }
    }
  }
  cp_async_wait<0>();
  if constexpr (CHI2) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(kFullMask, acc, off);
    if (lane == 0) A.chi2_partials[(size_t)blockIdx.x * kFastWarps + warp] = acc;
  }
}

// Model spectrum from the stored rows: model[i, j] = sum_c gvals[(i * nc + c) * nt + j], components
// added in order like spectrum_per_component.sum(axis=2) (analysis/model.py:279).
__global__ void __launch_bounds__(256) k_model_from_vals(const double* __restrict__ gvals, int nc, int nt,
                                                         long long total, double* __restrict__ model) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long i = e / nt;
    const int j = (int)(e - i * nt);
    const double* v = gvals + (size_t)i * nc * nt + j;
    double m = 0.0;
    for (int c = 0; c < nc; ++c) m += v[(size_t)c * nt];
    model[e] = m;
  }
}

// chi^2 of the stored spectrum values against the data (grid sweeps, BASELINE.json config 4):
// one block per good channel row, fixed-order reduction.
__global__ void __launch_bounds__(256) k_chi2_vals(const EvalArgs A, const double* __restrict__ gvals) {
  __shared__ double red[8];
  const int chi = blockIdx.x, tid = threadIdx.x, nc = A.pc.num_comp, nt = A.pc.num_time;
  const int i = A.chan_list ? A.chan_list[chi] : chi;
  const double wi = A.weights[i];
  const double* v = gvals + (size_t)chi * nc * nt;
  const double* d = A.data + (size_t)i * nt;
  double acc = 0.0;
  for (int j = tid; j < nt; j += 256) {
    double m = 0.0;
    for (int c = 0; c < nc; ++c) m += v[(size_t)c * nt + j];
    const double r = (d[j] - m) * wi;
    acc = fma(r, r, acc);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(kFullMask, acc, off);
  if ((tid & 31) == 0) red[tid >> 5] = acc;
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += red[w];
    A.chi2_partials[chi] = s;
  }
}

// ---- derivative side ------------------------------------------------------------------------
// Reduced column layout (first column of every group, -1 = not fitted); built on the host.
struct GramLayout {
  int rc_S, rc_t, rc_w;                    // per-component groups: nc columns each
  int rc_dm, rc_dmidx, rc_tau, rc_alpha;   // global parameters: one column each (summed over components)
  int rc_res;                              // residual column = nr - 1
  int nr;                                  // number of reduced columns
};

// Per (channel, component) constants of the scattered-branch first derivatives
// (routines/derivative.py:14-152,454-936 evaluated at the channel centre).
struct GTab {
  double inv_tau, ws, k2, h, ws2, g0, a_w, a_tau, a_alpha, lf, d_dm, d_dmidx, mean_delay;
};

__host__ __device__ constexpr int g3_pairs(int nb) { return nb * (nb + 1) / 2; }
// powers of L kept for the block pair (bi, bj): the spectrum columns live in block 0
__host__ __device__ constexpr int g3_npow(int bi, int bj) { return bi == 0 ? (bj == 0 ? 5 : 3) : 1; }
__host__ __device__ constexpr int g3_acc_offset(int nb, int bi, int bj) {
  int off = 0;
  for (int i = 0; i < nb; ++i)
    for (int j = i; j < nb; ++j) {
      if (i == bi && j == bj) return off;
      off += g3_npow(i, j);
    }
  return off;
}
__host__ __device__ constexpr int g3_nacc(int nb) { return g3_acc_offset(nb, nb, nb); }

// Far tiles (every sample outside the near range of every component: the shared factor G of
// the first derivatives is exactly zero there): each column of [J r] is a per-channel linear
// combination of the 2 NC + 1 basis values b = [M_c, M_c xi ..., r], xi = sample position mapped
// to [-1, 1] (d/dtau and d/dalpha are M times a polynomial of degree 1 in the time difference).
// Such a tile contributes through ONE 8x8 block B = sum b b^T (8 DMMA per 32 samples instead of
// 24, a dozen flops per sample instead of ~100), and at the end of the channel the block adds
// T B T^T to its full-matrix accumulators, T = (n+1) x 8 per-channel coefficients.
struct FarLayout {
  int enabled;
  int ncols;                          // n + 1
  int8_t param[FB_MAX_FREE + 1];      // fb_parameter of full column k, 9 = residual
  int8_t comp[FB_MAX_FREE + 1];       // component of column k (-1: global / residual)
};

__host__ __device__ inline size_t gram3_far_smem_bytes(int ncols) {
  // T [2][ncols][8] (this channel's and the previous one's), U [8][ncols], per-warp B [warps][64], ranges
  return sizeof(double) * ((size_t)24 * ncols + (size_t)kFastWarps * 64 + 64) + sizeof(int) * 2 * FB_MAX_COMPONENTS;
}

// Tiles of spectrum rows + data in flight per warp: with one tile ahead the kernel was bound by
// memory latency (16 KB in flight per SM, 2.2 TB/s: an ablation with all arithmetic removed still
// took 165 of 230 us); three tiles ahead, across channel boundaries, lift that bound.
#ifndef FB_G3_DEPTH
#define FB_G3_DEPTH 3  // measured 2 / 3 / 5 tiles ahead: 0.424 / 0.427 / 0.420 ms per evaluation (insensitive)
#endif
constexpr int kG3Depth = FB_G3_DEPTH, kG3Ring = kG3Depth + 1;

__host__ __device__ inline size_t gram3_smem_bytes(int nc, int nb, int ncols) {
  size_t stage = sizeof(double) * (size_t)kFastWarps * nb * 8 * kMmaRowStride;
  const size_t red = sizeof(double) * (size_t)kFastWarps * g3_nacc(nb) * 64;
  if (red > stage) stage = red;  // the stage doubles as the block-reduction buffer
  const size_t pref = sizeof(double) * kG3Ring * (size_t)(nc + 1) * kFastBlock;  // ring of prefetched tiles
  return sizeof(CompTab) * nc + 2 * sizeof(ChanTab) * nc + sizeof(GTab) * nc + stage + pref +
         gram3_far_smem_bytes(ncols) + sizeof(unsigned long long) * kFastWarps * kG3Ring + 32;
}

// BULK: the spectrum-row / data tiles are fetched with cp.async.bulk (one elected lane per warp,
// completion on an mbarrier per ring slot: SASS UBLKCP + SYNCS) instead of one 8-byte cp.async
// per lane (LDGSTS). Measured on B200 at config 2 (ncu, profiles/README.md): 262 us against 233 us
// under ncu, 0.476 against 0.445 ms per evaluation -- the consumers' try_wait loops add 2.4e7 warp
// instructions to a kernel that is issue-bound -- so the per-lane ring stays the default
// (FITBURST_B200_BULK=1 selects this variant; it needs an even num_time).
template <int NC, int NB, bool BULK = false>
__global__ void __launch_bounds__(kFastBlock, FB_G3_MINB) k_gram3(const EvalArgs A, const GramLayout G, const FarLayout F,
                                                         const double* __restrict__ gvals,
                                                         double* __restrict__ far_partials) {
  extern __shared__ __align__(16) unsigned char fast_smem[];
  const PlanConst& pc = A.pc;
  if (fit_stopped(pc)) return;
  const int kt = pc.kt, nt = pc.num_time;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  char* sp = reinterpret_cast<char*>(fast_smem);
  CompTab* comp = reinterpret_cast<CompTab*>(sp);
  sp += sizeof(CompTab) * NC;
  char* tabs = sp;  // [2][ChanTab NC]
  sp += 2 * sizeof(ChanTab) * NC;
  GTab* gt = reinterpret_cast<GTab*>(sp);
  sp += sizeof(GTab) * NC;
  double* stage_all = reinterpret_cast<double*>(sp);
  double* xt = stage_all + (size_t)warp * NB * 8 * kMmaRowStride;  // this warp's [NB*8 columns][36]
  {
    size_t stage = sizeof(double) * (size_t)kFastWarps * NB * 8 * kMmaRowStride;
    const size_t red_bytes = sizeof(double) * (size_t)kFastWarps * g3_nacc(NB) * 64;
    if (red_bytes > stage) stage = red_bytes;
    sp += stage;
  }
  // [kG3Ring][NC + 1][32] per warp: every lane copies and later reads only its own elements, so
  // cp.async.wait_group is the only synchronisation the prefetch needs (16-byte aligned: the
  // bulk-copy mode writes whole row segments)
  if constexpr (BULK) sp = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(sp) + 15) & ~(uintptr_t)15);
  double* pref = reinterpret_cast<double*>(sp) + (size_t)warp * kG3Ring * (NC + 1) * 32 + lane;
  sp += sizeof(double) * kG3Ring * (size_t)(NC + 1) * kFastBlock;
  const int ncols = F.ncols;
  double* Tm2 = reinterpret_cast<double*>(sp);      // [2][ncols][8], indexed by the channel parity `buf`
  double* Um = Tm2 + (size_t)16 * ncols;            // [8][ncols]
  double* Bw = Um + (size_t)8 * ncols;              // [warps][64]
  double* Bs = Bw + kFastWarps * 64;                // [64]
  int* range = reinterpret_cast<int*>(Bs + 64);     // [NC][2] near range of every component
  // one mbarrier per ring slot of every warp (bulk-copy mode), after the ranges, 8-byte aligned
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(
                                 (reinterpret_cast<uintptr_t>(range + 2 * FB_MAX_COMPONENTS) + 7) & ~(uintptr_t)7) +
                             warp * kG3Ring;
  constexpr bool kFarFits = 2 * NC + 1 <= 8;
  const bool far_on = kFarFits && F.enabled && nt > 2;
  constexpr bool bulk = BULK;

  {
    const double* src = reinterpret_cast<const double*>(A.g_comp);
    double* dst = reinterpret_cast<double*>(comp);
    for (int e = tid; e < NC * (int)(sizeof(CompTab) / sizeof(double)); e += kFastBlock) dst[e] = src[e];
  }
  for (int e = lane; e < NB * 8 * kMmaRowStride; e += 32) xt[e] = 0.0;  // padding columns stay zero
  if (bulk) {
    if (lane == 0) {
      for (int s = 0; s < kG3Ring; ++s) mbar_init(bars + s, 1);
      fence_proxy_async();  // the barriers are visible to the copy unit before the first copy arrives
    }
    __syncwarp();
  }

  constexpr int kPairs = g3_pairs(NB), kAcc = g3_nacc(NB);
  double c0[kPairs], c1[kPairs];
#pragma unroll
  for (int k = 0; k < kPairs; ++k) c0[k] = c1[k] = 0.0;
  // power accumulators of this warp live in its slice of the partial buffer (L2-resident, touched
  // once per channel, every lane only its own two elements of each 8x8 block): keeping them in
  // registers cost 36 registers and a fifth block per SM
  double* accw = A.partials + ((size_t)blockIdx.x * kFastWarps + warp) * kAcc * 64 + (lane >> 2) * 8 + (lane & 3) * 2;
#pragma unroll
  for (int k = 0; k < kAcc; ++k) *reinterpret_cast<double2*>(accw + k * 64) = make_double2(0.0, 0.0);
  // far-tile block B of the current channel, even and odd k-steps in separate accumulators (a
  // tile is 8 DMMA on the same C fragment: one chain of 8 dependent tensor instructions)
  double cf0 = 0.0, cf1 = 0.0, cg0 = 0.0, cg1 = 0.0;
  // full-matrix accumulators of the far tiles: thread e owns upper-triangle entries e and e + 256
  const int nent = ncols * (ncols + 1) / 2;
  double gacc[2] = {0.0, 0.0};
  int ek[2] = {0, 0}, el[2] = {0, 0};
  if (far_on) {
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      int e = tid + s * kFastBlock, k = 0;
      if (e < nent) {
        while (e >= ncols - k) {
          e -= ncols - k;
          ++k;
        }
        ek[s] = k;
        el[s] = k + e;
      }
    }
  }

  const int wtiles = (nt + 31) / 32;
  const int chan_chunks = (int)(sizeof(ChanTab) * NC / 16);
  // channel number and weight are read ahead of their use (chan_list[] -> weights[] / table address
  // are dependent global-load chains: ncu showed them exposed once per channel)
  auto channel_of = [&](int chi) { return chi < A.num_chan ? (A.chan_list ? __ldg(A.chan_list + chi) : chi) : -1; };
  auto prefetch_tables = [&](int i, int buf) {
    if (i >= 0 && tid < chan_chunks) {
      cp_async16(tabs + (size_t)buf * sizeof(ChanTab) * NC + 16 * tid,
                 reinterpret_cast<const char*>(A.g_chan + (size_t)i * NC) + 16 * tid);
    }
    cp_async_commit();
  };
  const double tau0 = global_param(A.params, NC, FB_SCATTERING_TIMESCALE);
  const double inv_tau0 = 1.0 / tau0;
  const double kt_mid = 0.5 * (double)(kt - 1);
  // xi = (tj - tj_mid) / tj_half over the samples of a channel
  const double tj_half = 0.5 * (double)((nt - 1) * kt), tj_mid = kt_mid + tj_half;
  const double xi_s = 1.0 / tj_half, xi_o = -tj_mid / tj_half;
  const int fr = lane & 3, fc = lane >> 2;  // fragment element: sample fr of the step, column fc of the block
  int i_cur = channel_of(blockIdx.x), i_next = channel_of(blockIdx.x + gridDim.x);
  double w_cur = i_cur >= 0 ? __ldg(A.weights + i_cur) : 0.0;
  prefetch_tables(i_cur, 0);
  // tiles of this warp within a channel: wt = warp + k * kFastWarps, k < nk
  const int nk = wtiles > warp ? (wtiles - warp + kFastWarps - 1) / kFastWarps : 0;
  unsigned q_issue = 0, q_use = 0;  // ring positions (one cp.async group per issued tile, even if empty)
  // The prefetch runs kG3Depth tiles ahead of the consumer along the warp's own sequence of
  // (channel, tile) pairs -- across as many channel boundaries as that takes (a window of 256
  // samples has two tiles per warp and channel: three tiles ahead is the channel after the next).
  int is_chi = blockIdx.x, is_i = i_cur, is_k = 0;  // issue cursor
  auto issue_next = [&]() {
    if (is_i >= 0 && nk > 0) {
      const int j0i = (warp + is_k * kFastWarps) * 32, j = j0i + lane;
      const double* srow_t = gvals + (size_t)is_chi * NC * nt;
      if (bulk) {
        // one lane hands the NC + 1 row segments of the tile to the copy unit; the slot's mbarrier
        // completes when all their bytes have landed. The slot was last read by this warp one
        // iteration ago (a __syncwarp lies in between): the proxy fence orders those reads first.
        if (lane == 0) {
          const unsigned slot = q_issue % kG3Ring;
          const unsigned bytes = (unsigned)(min(32, nt - j0i) * (int)sizeof(double));
          double* dst = pref - lane + (size_t)slot * (NC + 1) * 32;
          fence_proxy_async();
          mbar_expect_tx(bars + slot, (NC + 1) * bytes);
#pragma unroll
          for (int c = 0; c < NC; ++c) bulk_copy_g2s(dst + c * 32, srow_t + (size_t)c * nt + j0i, bytes, bars + slot);
          bulk_copy_g2s(dst + NC * 32, A.data + (size_t)is_i * nt + j0i, bytes, bars + slot);
        }
      } else if (j < nt) {
        double* dst = pref + (size_t)(q_issue % kG3Ring) * (NC + 1) * 32;
#pragma unroll
        for (int c = 0; c < NC; ++c) cp_async8(dst + c * 32, srow_t + (size_t)c * nt + j);
        cp_async8(dst + NC * 32, A.data + (size_t)is_i * nt + j);
      }
      if (++is_k >= nk) {
        is_k = 0;
        is_chi += gridDim.x;
        is_i = channel_of(is_chi);
      }
    }
    if (!bulk) cp_async_commit();
    ++q_issue;
  };
  for (int k = 0; k < kG3Depth; ++k) issue_next();
  bool have_prev = false;  // a previous channel of this block left far-tile blocks to contract
  auto far_accumulate = [&](const double* Tm) {
#pragma unroll
    for (int s2 = 0; s2 < 2; ++s2) {
      if (tid + s2 * kFastBlock < nent) {
        double s = 0.0;
#pragma unroll
        for (int a = 0; a < 8; ++a) s = fma(Tm[ek[s2] * 8 + a], Um[a * ncols + el[s2]], s);
        gacc[s2] += s;
      }
    }
  };
  int buf = 0;
  for (int chi = blockIdx.x; chi < A.num_chan; chi += gridDim.x, buf ^= 1) {
    const int i = i_cur;
    const double wi = w_cur;
    const int i_after = channel_of(chi + 2 * gridDim.x);
    const double w_next = i_next >= 0 ? __ldg(A.weights + i_next) : 0.0;
    // the table group of this channel is older than the nk tile groups committed since (in
    // bulk-copy mode the tables are the only cp.async groups)
    if (nk >= kG3Depth && !bulk) cp_async_wait<kG3Depth>();
    else cp_async_wait<0>();
    __syncthreads();
    prefetch_tables(i_next, buf ^ 1);
    i_cur = i_next;
    i_next = i_after;
    w_cur = w_next;
    const ChanTab* chan = reinterpret_cast<const ChanTab*>(tabs + (size_t)buf * sizeof(ChanTab) * NC);
    double* Tm_cur = Tm2 + (size_t)buf * 8 * ncols;
    const double* Tm_prev = Tm2 + (size_t)(buf ^ 1) * 8 * ncols;
    bool std_ok = true;  // every component on the scattered branch at this channel
#pragma unroll
    for (int c = 0; c < NC; ++c) std_ok = std_ok && (chan[c].sc_time != 0.0);
    if (tid < NC) {
      const ChanTab& ch = chan[tid];
      const ChanTab& ch0 = chan[0];
      const CompTab& ct = comp[tid];
      GTab t;
      t.inv_tau = ch.inv_tau;
      t.ws = ct.w * ch.inv_tau;
      t.k2 = t.ws * kSqrtHalf;
      t.ws2 = t.ws * t.ws;
      t.h = 0.5 * t.ws2;
      t.g0 = ch.amp * ch.amp_pbf * kTwoOverSqrtPi;
      t.a_w = ch.dap_w * ch.inv_apbf + t.ws * ch.inv_tau;
      // LSFitter.compute_jacobian passes component 0 to the global derivative functions, which
      // look deriv_amplitude_pbf up with it (analysis/fitter.py:213-223, routines/derivative.py:847,916)
      t.a_tau = ch0.dap_tau * ch.inv_apbf;
      t.a_alpha = ch0.dap_alpha * ch.inv_apbf;
      t.lf = ch.logf;
      t.d_dm = ch.d_dm;
      t.d_dmidx = ch.d_dmidx;
      t.mean_delay = ch.mean_delay;
      gt[tid] = t;
      int j_lo = 0, j_hi = nt;
      if (far_on && std_ok) {  // precomputed by the table kernel (finish_chan_tab)
        j_lo = ch.near_lo;
        j_hi = ch.near_hi;
      }
      range[2 * tid] = j_lo;
      range[2 * tid + 1] = j_hi;
    } else if (far_on && have_prev) {
      // far tiles of the PREVIOUS channel (its B blocks are complete: every warp passed the
      // barrier above), done by the threads that would otherwise wait for the serial setup of
      // this one: B summed over the warps in warp order, U = B T^T
      for (int e = tid - NC; e < 8 * ncols; e += kFastBlock - NC) {
        const int a = e / ncols, l = e - a * ncols;
        double s = 0.0;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
          double bsum = 0.0;
#pragma unroll
          for (int w = 0; w < kFastWarps; ++w) bsum += Bw[w * 64 + a * 8 + b];
          s = fma(bsum, Tm_prev[l * 8 + b], s);
        }
        Um[e] = s;
      }
    }
    __syncthreads();
    if (far_on) {
      if (have_prev) far_accumulate(Tm_prev);  // G += T U of the previous channel
      // T[k][a]: coefficient of basis value a in full column k (weight included)
      for (int e = tid; e < ncols * 8; e += kFastBlock) {
        const int k = e >> 3, a = e & 7, prm = F.param[k], cc = F.comp[k];
        const int ca = a >> 1;            // component the basis value belongs to (a < 2 NC)
        const bool lin = (a & 1) != 0;    // M_c xi rather than M_c
        double v = 0.0;
        if (a < 2 * NC) {
          const GTab& t = gt[ca];
          const CompTab& ct = comp[ca];
          // time difference of component ca: td = p + q xi
          const double q = ct.t_step * tj_half;
          const double p = fma(ct.t_step, tj_mid, ct.t_start) - t.mean_delay;
          const double u0 = p * t.inv_tau - t.ws2, u1 = q * t.inv_tau;  // u = u0 + u1 xi
          if (cc == ca && !lin) {
            if (prm == FB_AMPLITUDE) v = -kLn10;
            else if (prm == FB_SPECTRAL_INDEX) v = -t.lf;
            else if (prm == FB_SPECTRAL_RUNNING) v = -t.lf * t.lf;
            else if (prm == FB_ARRIVAL_TIME) v = -t.inv_tau;
            else if (prm == FB_BURST_WIDTH) v = -t.a_w;
          }
          if (prm == FB_DM && !lin) v = t.d_dm * t.inv_tau;
          if (prm == FB_DM_INDEX && !lin) v = t.d_dmidx * t.inv_tau;
          if (prm == FB_SCATTERING_TIMESCALE) v = lin ? -u1 * inv_tau0 : -(t.a_tau + u0 * inv_tau0);
          if (prm == FB_SCATTERING_INDEX) v = lin ? -t.lf * u1 : -(t.a_alpha + t.lf * u0);
        } else if (a == 2 * NC && prm == 9) {
          v = 1.0;
        }
        Tm_cur[e] = v * wi;
      }
    }

    unsigned far_mask = 0u;  // bit k & 31: tile k of this warp is a far tile (refreshed every 32 tiles)
    for (int k = 0; k < nk; ++k) {
      const int wt = warp + k * kFastWarps;
      const int j0 = wt * 32, j = j0 + lane;
      const bool valid = j < nt;
      if ((k & 31) == 0) {
        // lane l classifies tile k + l of this warp: one ballot instead of 2 NC range loads and
        // compares per tile (the source view put 7 % of the stall samples on that test)
        const int jl = (warp + (k + lane) * kFastWarps) * 32;
        bool far_l = far_on;
#pragma unroll
        for (int c = 0; c < NC; ++c) far_l = far_l && (jl + 32 <= range[2 * c] || jl >= range[2 * c + 1]);
        far_mask = __ballot_sync(kFullMask, far_l);
      }
      issue_next();  // the tile kG3Depth positions ahead of this one
      if (bulk) mbar_wait(bars + q_use % kG3Ring, (q_use / kG3Ring) & 1u);
      else cp_async_wait<kG3Depth>();
      double m[NC], dval = 0.0;
      {
        const double* src = pref + (size_t)(q_use % kG3Ring) * (NC + 1) * 32;
        ++q_use;
#pragma unroll
        for (int c = 0; c < NC; ++c) m[c] = valid ? src[c * 32] : 0.0;
        if (valid) dval = src[NC * 32];
      }
      double* row = xt + lane;
      const double tj = (double)(j * kt) + kt_mid;
      const bool far_tile = ((far_mask >> (k & 31)) & 1u) != 0u;
      if (far_tile) {
        const double xi = fma(tj, xi_s, xi_o);
        double msum = 0.0;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          msum += m[c];
          row[(2 * c) * kMmaRowStride] = m[c];
          row[(2 * c + 1) * kMmaRowStride] = m[c] * xi;
        }
        row[(2 * NC) * kMmaRowStride] = valid ? dval - msum : 0.0;
#pragma unroll
        for (int a = 2 * NC + 1; a < 8; ++a) row[a * kMmaRowStride] = 0.0;
        __syncwarp();
        // (four accumulator pairs -- chains of two dependent DMMAs instead of four -- were measured:
        // 8 more registers spill at the 128-register cap, 0.372 -> 0.379 ms per evaluation)
#pragma unroll
        for (int step = 0; step < 8; step += 2) {
          const double f = xt[fc * kMmaRowStride + 4 * step + fr];
          const double g = xt[fc * kMmaRowStride + 4 * step + 4 + fr];
          dmma8x8x4(cf0, cf1, f, f);
          dmma8x8x4(cg0, cg1, g, g);
        }
        __syncwarp();
        continue;
      }
      double msum = 0.0, s_dm = 0.0, s_dmidx = 0.0, s_tau = 0.0, s_alpha = 0.0;
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const CompTab& ct = comp[c];
        const GTab& t = gt[c];
        const double td = fma(ct.t_step, tj, ct.t_start) - t.mean_delay;
        const double mc = m[c];
        double d_t, d_w, d_tau, d_alpha, d_dm, d_dmidx;
        if (std_ok) {
          const double k1 = ct.inv_w * kSqrtHalf;
          const double arg_erf = fma(td, k1, -t.k2);
          const double arg_exp = fma(-arg_erf, arg_erf, fma(-td, t.inv_tau, t.h));
          const double g = (arg_exp < -746.0) ? 0.0 : t.g0 * fb_exp(arg_exp);
          const double gs = g * k1;
          d_t = fma(mc, t.inv_tau, -gs);
          d_w = fma(mc, t.a_w, -gs * fma(td, ct.inv_w, t.ws));
          const double u = fma(td, t.inv_tau, -t.ws2);
          const double gtt = g * t.k2;
          d_tau = fma(mc, fma(u, inv_tau0, t.a_tau), gtt * inv_tau0);
          d_alpha = fma(mc, fma(t.lf, u, t.a_alpha), gtt * t.lf);
          d_dm = -t.d_dm * d_t;
          d_dmidx = -t.d_dmidx * d_t;
        } else {
          double d[9], gfac;
          first_derivatives(chan[c], chan[c], ct, tau0, chan[c].amp, mc, td, d, &gfac);
          d[FB_SCATTERING_TIMESCALE] += mc * (chan[0].dap_tau - chan[c].dap_tau) * chan[c].inv_apbf;
          d[FB_SCATTERING_INDEX] += mc * (chan[0].dap_alpha - chan[c].dap_alpha) * chan[c].inv_apbf;
          d_t = d[FB_ARRIVAL_TIME];
          d_w = d[FB_BURST_WIDTH];
          d_tau = d[FB_SCATTERING_TIMESCALE];
          d_alpha = d[FB_SCATTERING_INDEX];
          d_dm = d[FB_DM];
          d_dmidx = d[FB_DM_INDEX];
        }
        msum += mc;
        s_dm += d_dm;
        s_dmidx += d_dmidx;
        s_tau += d_tau;
        s_alpha += d_alpha;
        if (G.rc_S >= 0) row[(G.rc_S + c) * kMmaRowStride] = -mc;
        if (G.rc_t >= 0) row[(G.rc_t + c) * kMmaRowStride] = valid ? -d_t : 0.0;
        if (G.rc_w >= 0) row[(G.rc_w + c) * kMmaRowStride] = valid ? -d_w : 0.0;
      }
      if (G.rc_dm >= 0) row[G.rc_dm * kMmaRowStride] = valid ? -s_dm : 0.0;
      if (G.rc_dmidx >= 0) row[G.rc_dmidx * kMmaRowStride] = valid ? -s_dmidx : 0.0;
      if (G.rc_tau >= 0) row[G.rc_tau * kMmaRowStride] = valid ? -s_tau : 0.0;
      if (G.rc_alpha >= 0) row[G.rc_alpha * kMmaRowStride] = valid ? -s_alpha : 0.0;
      row[G.rc_res * kMmaRowStride] = valid ? dval - msum : 0.0;  // analysis/fitter.py:262-264 (weight at the flush)
      // columns the far path may have used beyond the reduced set
      for (int a = G.nr; a < 8; ++a) row[a * kMmaRowStride] = 0.0;
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
    // channel flush: acc[p] += w^2 L^p C, C = 0 (L of component 0; k_gram3_finish shifts it
    // for components with a different reference frequency)
    {
      const double lf0 = chan[0].logf;
      double sp_[5];
      sp_[0] = wi * wi;
#pragma unroll
      for (int p = 1; p < 5; ++p) sp_[p] = sp_[p - 1] * lf0;
      int k = 0;
#pragma unroll
      for (int bi = 0; bi < NB; ++bi)
#pragma unroll
        for (int bj = bi; bj < NB; ++bj) {
          const int off = g3_acc_offset(NB, bi, bj);
#pragma unroll
          for (int p = 0; p < g3_npow(bi, bj); ++p) {
            double2 v = *reinterpret_cast<double2*>(accw + (off + p) * 64);
            v.x = fma(sp_[p], c0[k], v.x);
            v.y = fma(sp_[p], c1[k], v.y);
            *reinterpret_cast<double2*>(accw + (off + p) * 64) = v;
          }
          c0[k] = c1[k] = 0.0;
          ++k;
        }
    }
    if (far_on) {
      // far tiles of the channel: every warp parks its block B; the contraction G += T (sum B) T^T
      // happens at the top of the block's next channel (or after the loop), inside barriers that
      // are there anyway -- three block barriers per channel less than flushing here
      Bw[warp * 64 + (lane >> 2) * 8 + (lane & 3) * 2] = cf0 + cg0;
      Bw[warp * 64 + (lane >> 2) * 8 + (lane & 3) * 2 + 1] = cf1 + cg1;
      cf0 = cf1 = cg0 = cg1 = 0.0;
      have_prev = true;
    }
  }
  if (far_on && have_prev) {  // far tiles of the block's last channel (buf was flipped by the loop)
    const double* Tm_last = Tm2 + (size_t)(buf ^ 1) * 8 * ncols;
    __syncthreads();
    for (int e = tid; e < 8 * ncols; e += kFastBlock) {
      const int a = e / ncols, l = e - a * ncols;
      double s = 0.0;
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        double bsum = 0.0;
#pragma unroll
        for (int w = 0; w < kFastWarps; ++w) bsum += Bw[w * 64 + a * 8 + b];
        s = fma(bsum, Tm_last[l * 8 + b], s);
      }
      Um[e] = s;
    }
    __syncthreads();
    far_accumulate(Tm_last);
  }
  cp_async_wait<0>();
  if (far_on) {
#pragma unroll
    for (int s2 = 0; s2 < 2; ++s2)
      if (tid + s2 * kFastBlock < nent) far_partials[(size_t)blockIdx.x * nent + tid + s2 * kFastBlock] = gacc[s2];
  }
  // the four per-warp accumulator sets of this block are summed into the first one (warp order):
  // k_gram3_finish then reads one set per block instead of one per warp
  __syncthreads();
  {
    double* base = A.partials + (size_t)blockIdx.x * kFastWarps * kAcc * 64;
    for (int e = tid; e < kAcc * 64; e += kFastBlock) {
      double t = base[e];
#pragma unroll
      for (int w = 1; w < kFastWarps; ++w) t += base[(size_t)w * kAcc * 64 + e];
      base[e] = t;
    }
  }
}

// Full column k of [J r] = factor * (L + shift)^power * reduced column `red` (per channel).
struct ExpandArgs {
  int ncols, nb;
  int red[FB_MAX_FREE + 1];
  int power[FB_MAX_FREE + 1];
  double factor[FB_MAX_FREE + 1];
  double shift[FB_MAX_FREE + 1];
};

__device__ inline double g3_binom(int n, int k) {
  return (k == 0 || k == n) ? 1.0 : (n == 2 ? 2.0 : (double)n);  // n <= 2 here (n = 3, 4 never occur)
}

// One launch finishes an evaluation. Block (g, y) sums 32 neighbouring accumulator entries over its
// share of the partial sets k_gram3 left (near accumulators: one set per warp of k_gram3; far
// matrix: one per block): a warp reads 32 consecutive entries of one set (coalesced), the sets are
// dealt to gridDim.y x 8 warps, eight independent partial sums per thread keep eight loads in
// flight (the one-warp-per-entry version spent 25 us on 74 dependent strided loads per lane).
// The last block to finish adds the gridDim.y slices of every entry, expands the reduced
// accumulators into the full symmetric (n+1) x (n+1) matrix [J r]^T [J r] and adds the far-tile
// matrix. Fixed summation shape: deterministic.
constexpr int kFinishSlices = 16;
__host__ __device__ inline int gram3_finish_groups(int near_entries, int far_entries) {
  return (near_entries + 31) / 32 + (far_entries + 31) / 32;
}

__global__ void __launch_bounds__(256) k_gram3_finish(const double* __restrict__ partials, int nblocks,
                                                      int warps_per_block, int near_stride, int near_entries, int far_entries,
                                                      double* __restrict__ slices, double* __restrict__ racc,
                                                      double* __restrict__ far, const ExpandArgs X,
                                                      unsigned* __restrict__ counter, double* __restrict__ gram,
                                                      const int* __restrict__ stop) {
  if (stop != nullptr && *stop != 0) return;
  __shared__ bool last;
  __shared__ double red[8][33];
  const int total_entries = near_entries + far_entries;
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int near_groups = (near_entries + 31) / 32;
    const bool is_far = (int)blockIdx.x >= near_groups;
    const int e = (is_far ? (int)blockIdx.x - near_groups : (int)blockIdx.x) * 32 + lane;
    const int entries = is_far ? far_entries : near_entries;
    // near accumulators: `nblocks * warps_per_block` sets were allocated; every near_stride-th one is
    // read (k_gram3 pre-sums the sets of a block into its first: near_stride = warps_per_block)
    const size_t near_sets = (size_t)nblocks * warps_per_block;
    const double* src = is_far ? partials + near_sets * near_entries : partials;
    const int count = is_far ? nblocks : (int)(near_sets / near_stride);
    const size_t set_pitch = is_far ? (size_t)far_entries : (size_t)near_stride * near_entries;
    const int stride = 8 * (int)gridDim.y;
    double a[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) a[u] = 0.0;
    if (e < entries) {
      int b = (int)blockIdx.y * 8 + warp;
      for (; b + 7 * stride < count; b += 8 * stride) {
#pragma unroll
        for (int u = 0; u < 8; ++u) a[u] += __ldcg(src + (size_t)(b + u * stride) * set_pitch + e);
      }
      for (; b < count; b += stride) a[0] += __ldcg(src + (size_t)b * set_pitch + e);
    }
    red[warp][lane] = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
    __syncthreads();
    if (warp == 0 && e < entries) {
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += red[w][lane];
      slices[(size_t)blockIdx.y * total_entries + (is_far ? near_entries : 0) + e] = t;
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(counter, 1u) == gridDim.x * gridDim.y - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  if (threadIdx.x == 0) *counter = 0u;  // ready for the next evaluation
  for (int e = threadIdx.x; e < total_entries; e += blockDim.x) {
    double t = 0.0;
    for (int y = 0; y < (int)gridDim.y; ++y) t += __ldcg(slices + (size_t)y * total_entries + e);
    if (e < near_entries) racc[e] = t;
    else far[e - near_entries] = t;
  }
  __syncthreads();
  const int n = X.ncols;
  for (int o = threadIdx.x; o < n * n; o += blockDim.x) {
    int k = o / n, l = o % n;
    if (k > l) {  // entries (k, l) and (l, k) go through the same arithmetic: exactly symmetric output
      const int t = k;
      k = l;
      l = t;
    }
    double far_part = 0.0;
    if (far_entries > 0) far_part = __ldcg(far + (k * n - k * (k - 1) / 2 + (l - k)));
    if (X.red[k] > X.red[l]) {
      const int t = k;
      k = l;
      l = t;
    }
    const int rk = X.red[k], rl = X.red[l];
    const int off = g3_acc_offset(X.nb, rk >> 3, rl >> 3);
    const double* base = racc + (size_t)off * 64 + (rk & 7) * 8 + (rl & 7);
    double s = 0.0;
    for (int a = 0; a <= X.power[k]; ++a) {
      double ca = g3_binom(X.power[k], a);
      for (int q = 0; q < X.power[k] - a; ++q) ca *= X.shift[k];
      for (int b = 0; b <= X.power[l]; ++b) {
        double cb = g3_binom(X.power[l], b);
        for (int q = 0; q < X.power[l] - b; ++q) cb *= X.shift[l];
        s += ca * cb * __ldcg(base + (size_t)(a + b) * 64);
      }
    }
    gram[o] = X.factor[k] * X.factor[l] * s + far_part;
  }
}

}  // namespace fb
