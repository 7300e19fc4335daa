// fitburst_b200 fast evaluation, second generation: no spectrum-row round trip through HBM.
//
// k_vals3 + k_gram3 (fb_fast.cuh) write every spectrum value of every (good channel, component,
// sample) to global memory and read it back: 2 x 263 MB per evaluation at config 2 against 87.6 MB
// of algorithmic traffic (ncu: 593 MB of DRAM traffic per evaluation, 6.8 x the spectrum). Of the
// 32.9 M (sample, component) items only 0.47 M lie in the rise region and need exp + erfcx; the
// rest are closed forms (fb_core.cuh): a table constant on the clamp plateau, a kf-term dot
// product sum_a B[a] R[a][lane] in the exponential tail. Here
//
//   k_rise3  visits ONLY the rise region [jp, jt) of every (good channel, component): the
//            sub-frequency expansion (one exp + erfcx per sub-time) or the exact sub-grid, exactly
//            as phase B of k_vals3 does; stores those few values and the two regime boundaries.
//   k_gram4  k_gram3 with the model side folded in: per channel it builds the tail tables R / D
//            from the SubTab entries, per tile the base factors B (one exp per lane every 8th
//            tile, decay recurrence in between), and forms the spectrum value of each lane's
//            sample in registers -- constant, dot product, or one load of a rise value -- right
//            before the derivative rows are staged for the DMMA contraction. Only the data
//            (8 B per sample) stream through the cp.async ring.
//
// k_vals3 stays for everything that needs whole spectrum rows (model evaluation, the chi^2 grid,
// the exact Hessian at the end of a fit).
#pragma once
#include "fb_fast.cuh"

namespace fb {

#ifndef FB_R3_MINB
#define FB_R3_MINB 5
#endif
#ifndef FB_G4_MINB
#define FB_G4_MINB 4
#endif

__host__ __device__ inline size_t rise3_smem_bytes(int nc, int kf) {
  size_t n = 0;
  n += sizeof(CompTab) * nc;
  n += 2 * (sizeof(ChanTab) * nc + sizeof(SubTab) * nc * kf);  // double-buffered channel tables
  n += sizeof(int) * (4 * FB_MAX_COMPONENTS + 8);              // regime boundaries, item offsets
  n += sizeof(TayTab) * nc + 16;
  return n;
}

// Rise-region spectrum values gvals[(chi * nc + c) * num_time + j], jp <= j < jt, and the boundaries
// gbounds[(chi * nc + c) * 2 + {0, 1}] = {jp, jt} of every (good channel, component).
template <int KF>
__global__ void __launch_bounds__(kFastBlock, FB_R3_MINB) k_rise3(const EvalArgs A, double* __restrict__ gvals,
                                                                  int* __restrict__ gbounds) {
  extern __shared__ __align__(16) unsigned char fast_smem[];
  const PlanConst& pc = A.pc;
  if (fit_stopped(pc)) return;
  const int nc = pc.num_comp, kf = KF > 0 ? KF : pc.kf, kt = pc.kt, nt = pc.num_time;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nck = nc * kf;
  char* sp = reinterpret_cast<char*>(fast_smem);
  CompTab* comp = reinterpret_cast<CompTab*>(sp);
  sp += sizeof(CompTab) * nc;
  const size_t tab_bytes = sizeof(ChanTab) * nc + sizeof(SubTab) * nck;
  char* tabs = sp;  // [2][ChanTab nc | SubTab nc*kf]
  sp += 2 * tab_bytes;
  int* bounds = reinterpret_cast<int*>(sp);      // [nc][2]: jp, jt
  int* offs = bounds + 2 * FB_MAX_COMPONENTS;    // [nc + 1]: item offsets of the components
  sp += sizeof(int) * (4 * FB_MAX_COMPONENTS + 8);
  TayTab* tay = reinterpret_cast<TayTab*>((reinterpret_cast<uintptr_t>(sp) + 15) & ~(uintptr_t)15);

  {
    const double* src = reinterpret_cast<const double*>(A.g_comp);
    double* dst = reinterpret_cast<double*>(comp);
    for (int e = tid; e < nc * (int)(sizeof(CompTab) / sizeof(double)); e += kFastBlock) dst[e] = src[e];
  }
  const int chan_chunks = (int)(sizeof(ChanTab) * nc / 16), sub_chunks = (int)(sizeof(SubTab) * nck / 16);
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
  int lpi = 1;  // lanes per item: power of two >= kt
  while (lpi < kt && lpi < 32) lpi <<= 1;
  const int ipw = 32 / lpi;
  const double inv_n = 1.0 / (double)(kf * kt);
  int buf = 0;
  for (int chi = blockIdx.x; chi < A.num_chan; chi += gridDim.x, buf ^= 1) {
    const int i_after = channel_of(chi + 2 * gridDim.x);
    cp_async_wait<0>();
    __syncthreads();  // tables of this channel landed; every warp is done with the previous one
    prefetch_tables(i_next, buf ^ 1);
    i_next = i_after;
    const ChanTab* chan = reinterpret_cast<const ChanTab*>(tabs + (size_t)buf * tab_bytes);
    const SubTab* sub = reinterpret_cast<const SubTab*>(tabs + (size_t)buf * tab_bytes + sizeof(ChanTab) * nc);
    if (tid < nc) {
      int jp = chan[tid].jp, jt = chan[tid].jt;  // precomputed by the table kernel
      if (jt < jp) jt = jp;  // a component that is clamped past its own tail start has no rise region
      bounds[2 * tid] = jp;
      bounds[2 * tid + 1] = jt;
      gbounds[((size_t)chi * nc + tid) * 2] = jp;
      gbounds[((size_t)chi * nc + tid) * 2 + 1] = jt;
    }
    if (tid >= 32 && tid < 32 + nc) make_tay_tab(pc, comp[tid - 32], chan[tid - 32], sub + (tid - 32) * kf, kf, tay[tid - 32]);
    __syncthreads();
    if (tid == 0) {
      int acc = 0;
      for (int c = 0; c < nc; ++c) {
        offs[c] = acc;
        acc += bounds[2 * c + 1] - bounds[2 * c];
      }
      offs[nc] = acc;
    }
    __syncthreads();
    const int total = offs[nc];
    double* out = gvals + (size_t)chi * nc * nt;
    auto item_of = [&](int e, int& c, int& jj) {
      c = 0;
      while (c + 1 < nc && e >= offs[c + 1]) ++c;
      jj = bounds[2 * c] + (e - offs[c]);
    };
    if (kf * kt < 16) {
      // small sub-grids are not worth sharing between lanes: one thread per item
      for (int e = tid; e < total; e += kFastBlock) {
        int c, jj;
        item_of(e, c, jj);
        out[(size_t)c * nt + jj] = eval_sample_general(pc, comp[c], chan[c], sub + c * kf, jj).spectrum;
      }
      continue;
    }
    // A group of lanes (one per sub-time) evaluates an item by the sub-frequency expansion; items
    // it does not cover (a clamped sub-time, a component whose expansion bound is too large) take
    // the exact sub-grid, one per warp pass -- phase B of k_vals3 on an enumerated item list.
    for (int base = warp * ipw; base < total; base += kFastWarps * ipw) {
      const int q = base + lane / lpi, b = lane % lpi;
      bool have = q < total, expanded = false;
      int c = 0, jj = 0;
      double v = 0.0;
      if (have) {
        item_of(q, c, jj);
        const TayTab& t = tay[c];
        expanded = t.ok && kt <= 32;
        if (expanded && b < kt) {
          const CompTab& ct = comp[c];
          const double z = (linspace_at(ct.t_start, ct.t_step, ct.t_stop, jj * kt + b, pc.ntk) - t.d_bar) * ct.inv_w;
          if (z - t.dz_max >= -5.0) v = tay_eval(t, z);  // no sub-frequency clamped at -5 w
          else expanded = false;
        }
      }
      const unsigned fail = __ballot_sync(kFullMask, have && !expanded);
      for (int off = lpi >> 1; off > 0; off >>= 1) v += __shfl_xor_sync(kFullMask, v, off);
      const unsigned group = (lpi == 32 ? kFullMask : ((1u << lpi) - 1u)) << (lane / lpi * lpi);
      const bool item_fail = (fail & group) != 0u;
      if (have && !item_fail && b == 0) out[(size_t)c * nt + jj] = comp[c].amp10 * (v * inv_n);
      unsigned todo = __ballot_sync(kFullMask, have && item_fail && b == 0);
      while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const int c2 = __shfl_sync(kFullMask, c, src), j2 = __shfl_sync(kFullMask, jj, src);
        const double v2 = vals3_general_warp<KF>(pc, comp[c2], chan[c2], sub + c2 * kf, j2, lane);
        if (lane == 0) out[(size_t)c2 * nt + j2] = v2;
      }
    }
  }
  cp_async_wait<0>();
}

__host__ __device__ inline int gram4_bw_stride(int nck) { return (nck + 7) & ~7; }

__host__ __device__ inline size_t gram4_smem_bytes(int nc, int nb, int kf, int ncols) {
  size_t stage = sizeof(double) * (size_t)kFastWarps * nb * 8 * kMmaRowStride;
  const size_t red = sizeof(double) * (size_t)kFastWarps * g3_nacc(nb) * 64;
  if (red > stage) stage = red;
  const size_t pref = sizeof(double) * kG3Ring * (size_t)kFastBlock;  // ring of prefetched data tiles
  const size_t tabs = 2 * (sizeof(ChanTab) * nc + sizeof(SubTab) * (size_t)nc * kf);
  const size_t tail = sizeof(double) * ((size_t)nc * kf * 33 + (size_t)kFastWarps * gram4_bw_stride(nc * kf));
  return sizeof(CompTab) * nc + tabs + sizeof(GTab) * nc + stage + pref + tail + gram3_far_smem_bytes(ncols) +
         sizeof(int) * 2 * FB_MAX_COMPONENTS;
}

template <int NC, int NB>
__global__ void __launch_bounds__(kFastBlock, FB_G4_MINB) k_gram4(const EvalArgs A, const GramLayout G, const FarLayout F,
                                                                  const double* __restrict__ gvals,
                                                                  const int* __restrict__ gbounds,
                                                                  double* __restrict__ far_partials) {
  extern __shared__ __align__(16) unsigned char fast_smem[];
  const PlanConst& pc = A.pc;
  if (fit_stopped(pc)) return;
  const int kt = pc.kt, nt = pc.num_time, kf = pc.kf;
  const int nck = NC * kf;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  char* sp = reinterpret_cast<char*>(fast_smem);
  CompTab* comp = reinterpret_cast<CompTab*>(sp);
  sp += sizeof(CompTab) * NC;
  const size_t tab_bytes = sizeof(ChanTab) * NC + sizeof(SubTab) * (size_t)nck;
  char* tabs = sp;  // [2][ChanTab NC | SubTab NC*kf]
  sp += 2 * tab_bytes;
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
  // [kG3Ring][32] per warp: every lane copies and later reads only its own element
  double* pref = reinterpret_cast<double*>(sp) + (size_t)warp * kG3Ring * 32 + lane;
  sp += sizeof(double) * kG3Ring * (size_t)kFastBlock;
  double* R = reinterpret_cast<double*>(sp);  // [nck][32]: tail factors at the sample offsets of a tile
  sp += sizeof(double) * (size_t)nck * 32;
  double* D = reinterpret_cast<double*>(sp);  // [nck]: decay between two tiles of a warp
  sp += sizeof(double) * (size_t)nck;
  const int bws = gram4_bw_stride(nck);
  double* bw = reinterpret_cast<double*>(sp) + (size_t)warp * bws;  // base factors of this warp's tile
  sp += sizeof(double) * (size_t)kFastWarps * bws;
  const int ncols = F.ncols;
  double* Tm2 = reinterpret_cast<double*>(sp);      // [2][ncols][8], indexed by the channel parity `buf`
  double* Um = Tm2 + (size_t)16 * ncols;            // [8][ncols]
  double* Bw = Um + (size_t)8 * ncols;              // [warps][64]
  double* Bs = Bw + kFastWarps * 64;                // [64]
  int* range = reinterpret_cast<int*>(Bs + 64);     // [NC][2] near range of every component
  int* bnd = range + 2 * FB_MAX_COMPONENTS;         // [NC][2] jp, jt (k_rise3)
  constexpr bool kFarFits = 2 * NC + 1 <= 8;
  const bool far_on = kFarFits && F.enabled && nt > 2;

  {
    const double* src = reinterpret_cast<const double*>(A.g_comp);
    double* dst = reinterpret_cast<double*>(comp);
    for (int e = tid; e < NC * (int)(sizeof(CompTab) / sizeof(double)); e += kFastBlock) dst[e] = src[e];
  }
  for (int e = lane; e < NB * 8 * kMmaRowStride; e += 32) xt[e] = 0.0;  // padding columns stay zero

  constexpr int kPairs = g3_pairs(NB), kAcc = g3_nacc(NB);
  double c0[kPairs], c1[kPairs];
#pragma unroll
  for (int k = 0; k < kPairs; ++k) c0[k] = c1[k] = 0.0;
  double* accw = A.partials + ((size_t)blockIdx.x * kFastWarps + warp) * kAcc * 64 + (lane >> 2) * 8 + (lane & 3) * 2;
#pragma unroll
  for (int k = 0; k < kAcc; ++k) *reinterpret_cast<double2*>(accw + k * 64) = make_double2(0.0, 0.0);
  double cf0 = 0.0, cf1 = 0.0, cg0 = 0.0, cg1 = 0.0;
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
  const int chan_chunks = (int)(sizeof(ChanTab) * NC / 16), sub_chunks = (int)(sizeof(SubTab) * nck / 16);
  auto channel_of = [&](int chi) { return chi < A.num_chan ? (A.chan_list ? __ldg(A.chan_list + chi) : chi) : -1; };
  auto prefetch_tables = [&](int i, int buf) {
    if (i >= 0) {
      char* dst = tabs + (size_t)buf * tab_bytes;
      const char* src_c = reinterpret_cast<const char*>(A.g_chan + (size_t)i * NC);
      const char* src_s = reinterpret_cast<const char*>(A.g_sub + (size_t)i * nck);
      for (int e = tid; e < chan_chunks; e += kFastBlock) cp_async16(dst + 16 * e, src_c + 16 * e);
      for (int e = tid; e < sub_chunks; e += kFastBlock)
        cp_async16(dst + sizeof(ChanTab) * NC + 16 * e, src_s + 16 * e);
    }
    cp_async_commit();
  };
  const double tau0 = global_param(A.params, NC, FB_SCATTERING_TIMESCALE);
  const double inv_tau0 = 1.0 / tau0;
  const double kt_mid = 0.5 * (double)(kt - 1);
  const double tj_half = 0.5 * (double)((nt - 1) * kt), tj_mid = kt_mid + tj_half;
  const double xi_s = 1.0 / tj_half, xi_o = -tj_mid / tj_half;
  const int fr = lane & 3, fc = lane >> 2;
  int i_cur = channel_of(blockIdx.x), i_next = channel_of(blockIdx.x + gridDim.x);
  double w_cur = i_cur >= 0 ? __ldg(A.weights + i_cur) : 0.0;
  prefetch_tables(i_cur, 0);
  const int nk = wtiles > warp ? (wtiles - warp + kFastWarps - 1) / kFastWarps : 0;
  unsigned q_issue = 0, q_use = 0;
  int is_chi = blockIdx.x, is_i = i_cur, is_k = 0;  // issue cursor of the data prefetch
  auto issue_next = [&]() {
    if (is_i >= 0 && nk > 0) {
      const int j = (warp + is_k * kFastWarps) * 32 + lane;
      if (j < nt) cp_async8(pref + (size_t)(q_issue % kG3Ring) * 32, A.data + (size_t)is_i * nt + j);
      if (++is_k >= nk) {
        is_k = 0;
        is_chi += gridDim.x;
        is_i = channel_of(is_chi);
      }
    }
    cp_async_commit();
    ++q_issue;
  };
  for (int k = 0; k < kG3Depth; ++k) issue_next();
  bool have_prev = false;
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
    const double wi = w_cur;
    const int i_after = channel_of(chi + 2 * gridDim.x);
    const double w_next = i_next >= 0 ? __ldg(A.weights + i_next) : 0.0;
    // the table group of this channel is older than the nk data groups committed since
    if (nk >= kG3Depth) cp_async_wait<kG3Depth>();
    else cp_async_wait<0>();
    __syncthreads();
    prefetch_tables(i_next, buf ^ 1);
    i_cur = i_next;
    i_next = i_after;
    w_cur = w_next;
    const ChanTab* chan = reinterpret_cast<const ChanTab*>(tabs + (size_t)buf * tab_bytes);
    const SubTab* sub = reinterpret_cast<const SubTab*>(tabs + (size_t)buf * tab_bytes + sizeof(ChanTab) * NC);
    double* Tm_cur = Tm2 + (size_t)buf * 8 * ncols;
    const double* Tm_prev = Tm2 + (size_t)(buf ^ 1) * 8 * ncols;
    bool std_ok = true;
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
      t.a_tau = ch0.dap_tau * ch.inv_apbf;
      t.a_alpha = ch0.dap_alpha * ch.inv_apbf;
      t.lf = ch.logf;
      t.d_dm = ch.d_dm;
      t.d_dmidx = ch.d_dmidx;
      t.mean_delay = ch.mean_delay;
      gt[tid] = t;
      int j_lo = 0, j_hi = nt;
      if (far_on && std_ok) {
        j_lo = ch.near_lo;
        j_hi = ch.near_hi;
      }
      range[2 * tid] = j_lo;
      range[2 * tid + 1] = j_hi;
      bnd[2 * tid] = __ldcg(gbounds + ((size_t)chi * NC + tid) * 2);
      bnd[2 * tid + 1] = __ldcg(gbounds + ((size_t)chi * NC + tid) * 2 + 1);
    } else if (far_on && have_prev && tid >= 32) {
      // far tiles of the PREVIOUS channel, by the warps that do not run the serial setup above
      for (int e = tid - 32; e < 8 * ncols; e += kFastBlock - 32) {
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
    // tail tables of this channel: R[ca][l] = 10^A sed tail_c exp(-l kt step / tau), l = 0..31,
    // and the decay over the kFastWarps tiles between two tiles of one warp
    for (int e = tid; e < nck * 32; e += kFastBlock) {
      const int l = e & 31, ca = e >> 5, c = ca / kf;
      const SubTab& st = sub[ca];
      const double decay = fb_exp(-((double)(l * kt) * comp[c].t_step) * st.inv_tau);
      R[ca * 32 + l] = ((comp[c].amp10 * st.tail_c) * st.sed) * decay;
    }
    for (int e = tid; e < nck; e += kFastBlock)
      D[e] = fb_exp(-((double)(32 * kFastWarps * kt) * comp[e / kf].t_step) * sub[e].inv_tau);
    __syncthreads();
    if (far_on) {
      if (have_prev) far_accumulate(Tm_prev);
      for (int e = tid; e < ncols * 8; e += kFastBlock) {
        const int k = e >> 3, a = e & 7, prm = F.param[k], cc = F.comp[k];
        const int ca = a >> 1;
        const bool lin = (a & 1) != 0;
        double v = 0.0;
        if (a < 2 * NC) {
          const GTab& t = gt[ca];
          const CompTab& ct = comp[ca];
          const double q = ct.t_step * tj_half;
          const double p = fma(ct.t_step, tj_mid, ct.t_start) - t.mean_delay;
          const double u0 = p * t.inv_tau - t.ws2, u1 = q * t.inv_tau;
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
    // first tile of this warp that holds a tail sample of any component: the base factors are
    // needed from there on
    int jt_min = nt;
#pragma unroll
    for (int c = 0; c < NC; ++c) jt_min = min(jt_min, max(bnd[2 * c + 1], bnd[2 * c]));
    const double* vrow = gvals + (size_t)chi * NC * nt;
    bool b_live = false;  // bw holds the factors of this warp's previous tile
    int b_age = 0;
    for (int k = 0; k < nk; ++k) {
      const int wt = warp + k * kFastWarps;
      const int j0 = wt * 32, j = j0 + lane;
      const bool valid = j < nt;
      issue_next();  // the data tile kG3Depth positions ahead of this one
      unsigned bmask = 0u;
      if (j0 + 32 > jt_min) {
        // base factors B of every (component, sub-frequency) at the first sample of the tile: one
        // exp per lane every 8th tile, the decay recurrence in between (a factor that overflowed
        // ahead of the burst is recomputed: inf * decay would stay inf)
        const bool refresh = !b_live || b_age >= 7;
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
        bmask = nck <= 32 ? __ballot_sync(kFullMask, b_finite) : 0u;
        b_age = refresh ? 0 : b_age + 1;
        b_live = true;
        __syncwarp();
      } else {
        b_live = false;
      }
      cp_async_wait<kG3Depth>();
      double m[NC], dval = 0.0;
      {
        const double* src = pref + (size_t)(q_use % kG3Ring) * 32;
        ++q_use;
        if (valid) dval = *src;
      }
      const unsigned kf_mask = kf >= 32 ? kFullMask : ((1u << kf) - 1u);
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const int jp = bnd[2 * c], jt = bnd[2 * c + 1];
        const double* rl = R + (size_t)c * kf * 32 + lane;
        const double* bc = bw + c * kf;
        double val = 0.0;
        if (j0 >= jt && j0 >= jp) {  // whole tile in the exponential tail
          double s0 = 0.0, s1 = 0.0;
          int a = 0;
          for (; a + 1 < kf; a += 2) {
            s0 = fma(bc[a], rl[a * 32], s0);
            s1 = fma(bc[a + 1], rl[(a + 1) * 32], s1);
          }
          if (a < kf) s0 = fma(bc[a], rl[a * 32], s0);
          val = s0 + s1;
        } else if (j0 + 32 <= jp) {  // whole tile on the clamp plateau
          val = chan[c].plateau_ps;
        } else {
          const bool dot_ok = nck <= 32 && ((bmask >> (c * kf)) & kf_mask) == kf_mask;
          if (j < jp) {
            val = chan[c].plateau_ps;
          } else if (j >= jt) {
            if (dot_ok) {
              double s0 = 0.0, s1 = 0.0;
              int a = 0;
              for (; a + 1 < kf; a += 2) {
                s0 = fma(bc[a], rl[a * 32], s0);
                s1 = fma(bc[a + 1], rl[(a + 1) * 32], s1);
              }
              if (a < kf) s0 = fma(bc[a], rl[a * 32], s0);
              val = s0 + s1;
            } else if (valid) {
              const CompTab& ct = comp[c];
              const double u0 = linspace_at(ct.t_start, ct.t_step, ct.t_stop, j * kt, pc.ntk);
              double s_ps = 0.0;
              for (int a = 0; a < kf; ++a) {
                const SubTab& st = sub[c * kf + a];
                const double earg = fma(-u0, st.inv_tau, st.tail_e);
                const double p = (earg < -746.0) ? 0.0 : st.tail_c * fb_exp(earg);
                s_ps = fma(p, st.sed, s_ps);
              }
              val = ct.amp10 * s_ps;
            }
          } else if (valid) {
            val = __ldcg(vrow + (size_t)c * nt + j);  // rise region: k_rise3
          }
        }
        m[c] = valid ? val : 0.0;
      }
      double* row = xt + lane;
      const double tj = (double)(j * kt) + kt_mid;
      bool far_tile = far_on;
#pragma unroll
      for (int c = 0; c < NC; ++c) far_tile = far_tile && (j0 + 32 <= range[2 * c] || j0 >= range[2 * c + 1]);
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
      row[G.rc_res * kMmaRowStride] = valid ? dval - msum : 0.0;
      for (int a = G.nr; a < 8; ++a) row[a * kMmaRowStride] = 0.0;
      __syncwarp();
#pragma unroll
      for (int step = 0; step < 8; ++step) {
        double f[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) f[b] = xt[(8 * b + fc) * kMmaRowStride + 4 * step + fr];
        int kk = 0;
#pragma unroll
        for (int bi = 0; bi < NB; ++bi)
#pragma unroll
          for (int bj = bi; bj < NB; ++bj) {
            dmma8x8x4(c0[kk], c1[kk], f[bi], f[bj]);
            ++kk;
          }
      }
      __syncwarp();
    }
    // channel flush: acc[p] += w^2 L^p C
    {
      const double lf0 = chan[0].logf;
      double sp_[5];
      sp_[0] = wi * wi;
#pragma unroll
      for (int p = 1; p < 5; ++p) sp_[p] = sp_[p - 1] * lf0;
      int kk = 0;
#pragma unroll
      for (int bi = 0; bi < NB; ++bi)
#pragma unroll
        for (int bj = bi; bj < NB; ++bj) {
          const int off = g3_acc_offset(NB, bi, bj);
#pragma unroll
          for (int p = 0; p < g3_npow(bi, bj); ++p) {
            double2 v = *reinterpret_cast<double2*>(accw + (off + p) * 64);
            v.x = fma(sp_[p], c0[kk], v.x);
            v.y = fma(sp_[p], c1[kk], v.y);
            *reinterpret_cast<double2*>(accw + (off + p) * 64) = v;
          }
          c0[kk] = c1[kk] = 0.0;
          ++kk;
        }
    }
    if (far_on) {
      Bw[warp * 64 + (lane >> 2) * 8 + (lane & 3) * 2] = cf0 + cg0;
      Bw[warp * 64 + (lane >> 2) * 8 + (lane & 3) * 2 + 1] = cf1 + cg1;
      cf0 = cf1 = cg0 = cg1 = 0.0;
      have_prev = true;
    }
  }
  if (far_on && have_prev) {
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
}

}  // namespace fb
