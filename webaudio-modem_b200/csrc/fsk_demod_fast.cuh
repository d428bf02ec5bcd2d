// fsk_demod_fast.cuh — fused FSK demodulator, mixed-precision fast path with certified decisions.
//
// Same job and same tile structure as fsk_demod_exact_kernel (fsk_demod.cuh): one thread walks one stream
// (FSKCore.demodulateData, src/modems/fsk.ts:190-375), a warp owns 32 streams and is its own CTA, samples arrive as
// 32 x 32 TMA tiles.  What changes is the arithmetic:
//   A1  the AGC gain recurrence (fsk.ts:52-76) stays in float64 — its branch `level > 0.5` is a discontinuity, and an
//       exact gain keeps the float32 store of fsk.ts:55 exact — but the band-pass pre-filter runs in float32;
//   A2  LO rotation, I/Q low-pass (packed f32x2: I and Q share their coefficients), /2 decimation, phase difference as
//       atan2(cross, dot) of consecutive phasors, post low-pass and slicer all run in float32.  The three biquads use
//       the NORMAL (coupled) form instead of the reference's direct form I: same transfer function, but in float32 its
//       round-off on filteredPhaseDiff is 1e-8 rms against 5e-7 (measured against the oracle, oracle/fastmodel.c);
//   B   the decimated-rate state machine (fsk.ts:278-375) is the exact kernel's event-driven one plus DOUBT TRACKING:
//       a hard bit whose |filteredPhaseDiff| is inside the float32 error band is doubtful (second bit ring), an
//       amplitude within a few ulps of the silence threshold is doubtful, and a DECISION — majority vote, sync
//       threshold, EOD — that the doubtful samples could turn FLAGS the stream.
// A flagged stream is demodulated again by the float64 kernels from the state it had at the start of the call (host:
// fast_demodulate in wam_api.cu), so the bytes and counters that leave the library are the float64 ones wherever the
// float32 arithmetic could have mattered.  Measured on the CPU model of this kernel (oracle/fastmodel.c, 16,384 noisy
// V.21 streams, -15..+30 dB): 0 streams differ from the oracle among the unflagged ones, 0.06 % are flagged.
//
// The per-stream state lives in the same arrays and the same (direct form) representation as the exact kernel's, so
// both kernels can run on a batch in any order: this kernel converts on the way in and out (float64, once per launch).
#pragma once

#include "fsk_demod.cuh"

#ifndef WAM_FAST_SEARCH_BATCH
#define WAM_FAST_SEARCH_BATCH 6  // groups of four sync-ring words requested up front by the frame search (0 / 1: off)
#endif

namespace wam {

struct FastA2 {    // everything resetState() zeroes on the DSP side, float32
  float lc, ls;    // LO rotation
  float2 w1, w2;   // I/Q low-pass, normal form, packed (I, Q)
  float ow1, ow2;  // post low-pass
  float psi, psq;  // previous decimated phasor (lastPhase as a vector)
  float2 acc;      // decimator
  float S, E, rsp; // doubt envelope: amplitude scale, error envelope of the post filter, 1 / (4 amp) of the last phasor
  uint32_t dsc;
};
struct FastB {     // BState + doubt tracking
  float sil_thr;
  uint32_t gsc, gmod, bsc, next_idx, bit_acc, bit_cnt, started, current, sil_cnt;
  int bitpos;
  uint32_t ring_pos, ring_len, amp_pos, amp_len, cur_word, dcur_word;
  int out_n;
  uint32_t dvote, silx, flag;
  uint32_t dlast, dcnt;  // ring position behind the newest doubtful hard bit (0: none yet); doubtful bits put since the
                         // last gap of total_bits + 32 positions without any (saturating)
};

constexpr int kFParkU = 23;  // pre-filter state (2 float words) + 21 state-machine words

__device__ __forceinline__ void fb_load(FastB& b, const uint32_t (*pu)[32], int lane) {
  b.sil_thr = __uint_as_float(pu[2][lane]);
  b.gsc = pu[3][lane]; b.gmod = pu[4][lane]; b.bsc = pu[5][lane]; b.next_idx = pu[6][lane];
  b.bit_acc = pu[7][lane]; b.bit_cnt = pu[8][lane];
  const uint32_t f = pu[9][lane];
  b.started = f & 1u; b.bitpos = (int)((f >> 8) & 0xffu) - 1; b.current = (f >> 16) & 0xffu;
  b.sil_cnt = pu[10][lane]; b.ring_pos = pu[11][lane]; b.ring_len = pu[12][lane];
  b.amp_pos = pu[13][lane]; b.amp_len = pu[14][lane]; b.cur_word = pu[15][lane]; b.out_n = (int)pu[16][lane];
  b.dcur_word = pu[17][lane]; b.dvote = pu[18][lane]; b.silx = pu[19][lane]; b.dlast = pu[20][lane];
  b.flag = pu[21][lane]; b.dcnt = pu[22][lane];
}
__device__ __forceinline__ void fb_store(const FastB& b, uint32_t (*pu)[32], int lane) {
  pu[2][lane] = __float_as_uint(b.sil_thr);
  pu[3][lane] = b.gsc; pu[4][lane] = b.gmod; pu[5][lane] = b.bsc; pu[6][lane] = b.next_idx;
  pu[7][lane] = b.bit_acc; pu[8][lane] = b.bit_cnt;
  pu[9][lane] = (b.started & 1u) | ((uint32_t)(b.bitpos + 1) << 8) | ((b.current & 0xffu) << 16);
  pu[10][lane] = b.sil_cnt; pu[11][lane] = b.ring_pos; pu[12][lane] = b.ring_len;
  pu[13][lane] = b.amp_pos; pu[14][lane] = b.amp_len; pu[15][lane] = b.cur_word; pu[16][lane] = (uint32_t)b.out_n;
  pu[17][lane] = b.dcur_word; pu[18][lane] = b.dvote; pu[19][lane] = b.silx; pu[20][lane] = b.dlast;
  pu[21][lane] = b.flag; pu[22][lane] = b.dcnt;
}

__device__ __forceinline__ void reset_state_fa2(FastA2& s) {
  s.lc = 1.0f; s.ls = 0.0f;
  s.w1 = make_float2(0.0f, 0.0f); s.w2 = make_float2(0.0f, 0.0f);
  s.ow1 = s.ow2 = 0.0f;
  s.psi = 1.0f; s.psq = 0.0f;  // lastPhase = 0
  s.acc = make_float2(0.0f, 0.0f);
  s.E = 0.0f; s.rsp = 0.0f;    // the error envelope belongs to the post filter's state
  s.dsc = 0;
}
__device__ __forceinline__ void reset_state_fb(FastB& b) {
  b.gsc = 0; b.gmod = 0; b.bsc = 0; b.bit_acc = 0; b.bit_cnt = 0; b.next_idx = 0;
  b.current = 0; b.bitpos = 0;
  b.started = 0;
  b.sil_cnt = 0;
  b.dvote = 0; b.silx = 0;
}

// ---- state conversion: direct form I history <-> normal-form state (float64, once per launch and stream) ----
// Both describe the filter just before its next input.  The future output depends on two numbers only:
//   s1 = b1 x1 + b2 x2 - a1 y1 - a2 y2      (zero-input part of the next output)
//   s2 = b2 x1 - a2 y1 - a1 s1              (zero-input part of the one after)
// and in normal form s1 = k . w, s2 = k . (R w).
__device__ __forceinline__ void df_to_normal(double b1, double b2, double a1, double a2, double k1, double k2, double sg,
                                             double om, double x1, double x2, double y1, double y2, float& w1, float& w2) {
  const double s1 = b1 * x1 + b2 * x2 - a1 * y1 - a2 * y2;
  const double s2 = b2 * x1 - a2 * y1 - a1 * s1;
  const double m21 = k1 * sg + k2 * om, m22 = k2 * sg - k1 * om;
  const double det = k1 * m22 - k2 * m21;
  w1 = (float)((s1 * m22 - k2 * s2) / det);
  w2 = (float)((k1 * s2 - m21 * s1) / det);
}
__device__ __forceinline__ void normal_to_df(double a1, double a2, double k1, double k2, double sg, double om, double w1,
                                             double w2, double& y1, double& y2) {
  const double s1 = k1 * w1 + k2 * w2;
  const double s2 = (k1 * sg + k2 * om) * w1 + (k2 * sg - k1 * om) * w2;
  y1 = -(s2 + a1 * s1) / a2;  // with x1 = x2 = 0
  y2 = -(s1 + a1 * y1) / a2;
}

// ---- float32 primitives ----
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rsqrt_approx(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// atan2(y, x), absolute error <= 1.5e-7 (degree-15 odd minimax on [0, 1] + float32 rounding); atan2(0, 0) = 0.
__device__ __forceinline__ float fast_atan2f(float y, float x) {
  const float ax = fabsf(x), ay = fabsf(y);
  const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  const float r = mx > 1e-37f ? mn * rcp_approx(mx) : 0.0f;  // (the reciprocal flushes denormals)
  const float z = r * r;
  float p = -0.004054499790072441f;
  p = fmaf(p, z, 0.021862739697098732f);
  p = fmaf(p, z, -0.05591205880045891f);
  p = fmaf(p, z, 0.09642183035612106f);
  p = fmaf(p, z, -0.1390862613916397f);
  p = fmaf(p, z, 0.19946566224098206f);
  p = fmaf(p, z, -0.33329862356185913f);
  p = fmaf(p, z, 0.9999993443489075f);
  float a = p * r;
  a = ay > ax ? 1.5707963267948966f - a : a;
  a = x < 0.0f ? 3.141592653589793f - a : a;
  return copysignf(a, y);
}

// ---- phase A1: float64 AGC (the exact kernel's arithmetic) + float32 normal-form pre-filter ----
__device__ __forceinline__ float fast_a1_sample(double& gain, float& w1, float& w2, float x, const FskDerived& d, bool agc,
                                                double att, double rel) {
  const float sg_agc = (float)((double)x * gain);
  const float sg = agc ? sg_agc : x;
  const float level = fabsf(sg);
  const float lv = fmaxf(level, 5e-31f);
  const double r = (double)rcp_approx(lv);
  const double inv = fma(r, fma(-(double)lv, r, 1.0), r);
  const double rate = (agc && level > 0.0f) ? (level > 0.5f ? att : rel) : 0.0;
  double g = fma(fma(inv, 0.5, -gain), rate, gain);
  {
    const bool over = g > 10.0, under = g < 0.1;
    g = over ? 10.0 : g;
    g = under ? 0.1 : g;
  }
  gain = g;
  const float y = fmaf(d.f_pre_k2, w2, fmaf(d.f_pre_k1, w1, d.f_pre_k0 * sg));
  const float n1 = fmaf(d.f_pre_sg, w1, fmaf(-d.f_pre_om, w2, sg));
  const float n2 = fmaf(d.f_pre_om, w1, d.f_pre_sg * w2);
  w1 = n1; w2 = n2;
  return y;
}

// ---- phase A2: one input sample through the LO and the packed I/Q low-pass ----
__device__ __forceinline__ float2 fast_a2_half(FastA2& s, float pf, const FskDerived& d) {
  const float2 x = __fmul2_rn(make_float2(pf, pf), make_float2(s.lc, s.ls));
  const float nc = fmaf(s.lc, d.f_cw, -(s.ls * d.f_sw));
  const float nsn = fmaf(s.ls, d.f_cw, s.lc * d.f_sw);
  s.lc = nc; s.ls = nsn;
  const float2 k0 = make_float2(d.f_lp_k0, d.f_lp_k0), k1 = make_float2(d.f_lp_k1, d.f_lp_k1);
  const float2 k2 = make_float2(d.f_lp_k2, d.f_lp_k2), sg = make_float2(d.f_lp_sg, d.f_lp_sg);
  const float2 om = make_float2(d.f_lp_om, d.f_lp_om), nom = make_float2(-d.f_lp_om, -d.f_lp_om);
  const float2 y = __ffma2_rn(k2, s.w2, __ffma2_rn(k1, s.w1, __fmul2_rn(k0, x)));
  const float2 n1 = __ffma2_rn(sg, s.w1, __ffma2_rn(nom, s.w2, x));
  const float2 n2 = __ffma2_rn(om, s.w1, __fmul2_rn(sg, s.w2));
  s.w1 = n1; s.w2 = n2;
  return y;
}

// decimated-rate discriminator on the summed pair (2 avgI, 2 avgQ): hard bit, doubt flag, amplitude (fsk.ts:246-264)
__device__ __forceinline__ void fast_a2_decim(FastA2& s, float2 sum, const FskDerived& d, uint32_t& bit, uint32_t& dbit,
                                              float& amp) {
  const float si = sum.x, sq = sum.y;
  // wrapped (phase - lastPhase) straight from the two phasors; the LO's float32 frequency offset is a known constant
  const float cross = fmaf(sq, s.psi, -(si * s.psq));
  const float dot = fmaf(si, s.psi, sq * s.psq);
  const float pd = fast_atan2f(cross, dot) - d.f_dphi_bias;
  const float pw = fmaf(si, si, sq * sq);
  const bool tiny = !(pw > 1e-30f);  // also catches NaN
  const float rs = tiny ? 0.0f : rsqrt_approx(pw);
  amp = 0.5f * pw * rs;              // fsk.ts:252 (amplitude of the averaged pair)
  s.psi = si; s.psq = sq;
  const float fpd = fmaf(d.f_lp_k2, s.ow2, fmaf(d.f_lp_k1, s.ow1, d.f_lp_k0 * pd));
  const float n1 = fmaf(d.f_lp_sg, s.ow1, fmaf(-d.f_lp_om, s.ow2, pd));
  const float n2 = fmaf(d.f_lp_om, s.ow1, d.f_lp_sg * s.ow2);
  s.ow1 = n1; s.ow2 = n2;
  bit = fpd > 0.0f ? 1u : 0u;
  // doubt band: the float32 error of a phasor's angle grows as (recent amplitude scale) / (its own length); the post
  // filter spreads it with |h(j)| <= gamma rho^j; a raw difference next to +-pi may have wrapped the other way (2 pi)
  s.S = fmaxf(amp, s.S * 0.9921875f);
  const float hrs = 0.5f * rs;
  const float e1 = d.f_kappa * s.S * (hrs + s.rsp);
  s.rsp = hrs;
  float et = tiny ? 10.0f : e1;  // a vanishing phasor has no usable angle
  if (fabsf(fabsf(pd) - 3.14159265f) < fmaf(4.0f, e1, d.f_bc_delta)) et += 6.3f;
  s.E = fmaf(d.f_rho_e, s.E, d.f_gamma * et);
  dbit = fabsf(fpd) < s.E + d.f_eps0 ? 1u : 0u;
}

// Upper bound of the doubtful hard bits among the newest total_bits ring samples, without touching the doubt ring:
// dcnt counts the doubtful bits put since the last gap of total_bits + 32 positions without any; dlast is the position
// behind the newest one.  Once ring_pos - dlast exceeds that gap the window is clean.
__device__ __forceinline__ void doubt_note(FastB& b, uint32_t p, uint32_t dchunk, const FskDerived& d) {
  if (b.dlast != 0u && p - b.dlast >= (uint32_t)d.total_bits + 32u) b.dcnt = 0u;
  b.dcnt = min(b.dcnt + (uint32_t)__popc(dchunk), 0xffffu);
  b.dlast = p + (32u - (uint32_t)__clz((int)dchunk));
}
__device__ __forceinline__ int doubt_bound(FastB& b, uint32_t pos, const FskDerived& d) {
  if (pos - b.dlast >= (uint32_t)d.total_bits + 32u) b.dcnt = 0u;
  return (int)b.dcnt;
}
// exact number of doubtful hard bits among the newest total_bits ring samples (rare: only when the bound matters)
__device__ __noinline__ int doubt_count(const uint32_t* __restrict__ dring, uint32_t pos, const FskDerived& d) {
  const uint32_t lo = pos - (uint32_t)d.total_bits;
  const uint32_t o = lo & 31u;
  const uint32_t wmask = (uint32_t)(d.ring_words - 1);
  uint32_t w = (lo >> 5) & wmask;
  uint32_t prev = dring[w];
  int n = 0;
  int left = d.total_bits;
  while (left > 0) {
    w = (w + 1u) & wmask;
    const uint32_t cur = dring[w];
    uint32_t v = __funnelshift_r(prev, cur, o);
    if (left < 32) v &= (1u << left) - 1u;
    n += __popc(v);
    prev = cur;
    left -= 32;
  }
  return n;
}

// sync_mismatches0v (fsk_demod.cuh) with a caller-supplied cut-off: with D doubtful bits in the window the search may
// only stop early once the threshold is out of reach even if all of them flipped.
template <bool CONST_SLOT>
__device__ __forceinline__ int sync_mismatches0v_cut(const uint32_t* __restrict__ ring, uint32_t pos, const FskDerived& d,
                                                     int cutoff) {
  const uint32_t lo = pos - (uint32_t)d.total_bits;
  const uint32_t o = lo & 31u;
  const uint32_t wmask = (uint32_t)(d.ring_words - 1);
  const uint32_t w = (lo >> 5) & wmask;
  const uint32_t s = w & 3u;
  uint32_t g = w & ~3u;
  const int n_end = (int)s + d.tmpl0_words;
  const uint32_t* __restrict__ ex = (CONST_SLOT ? c_tmpl[d.tmpl_slot][0] : d.tmpl0_expect) + 4 - (int)s;
  const uint32_t* __restrict__ mk = (CONST_SLOT ? c_tmpl[d.tmpl_slot][1] : d.tmpl0_mask) + 4 - (int)s;
  const int n_full = (int)s + d.tmpl0_full;
  uint4 v, nx;
  int mism, n;
  uint32_t prev;
#if WAM_FAST_SEARCH_BATCH > 1
  if (n_full >= 3 + 4 * (WAM_FAST_SEARCH_BATCH - 1)) {
    // first round: WAM_FAST_SEARCH_BATCH groups of four ring words requested at once (nobody is rejected before a dozen
    // words and the average search ends after ~20): one L2 round trip instead of one per group
    uint4 q[WAM_FAST_SEARCH_BATCH];
#pragma unroll
    for (int j = 0; j < WAM_FAST_SEARCH_BATCH; ++j) q[j] = ring_ld4(ring + ((g + 4u * (uint32_t)j) & wmask));
    g = (g + 4u * WAM_FAST_SEARCH_BATCH) & wmask;
    nx = ring_ld4(ring + g);
    g = (g + 4u) & wmask;
    mism = __popc((__funnelshift_r(q[0].x, q[0].y, o) ^ ex[0]) & mk[0]) + __popc((__funnelshift_r(q[0].y, q[0].z, o) ^ ex[1]) & mk[1]) +
           __popc((__funnelshift_r(q[0].z, q[0].w, o) ^ ex[2]) & mk[2]);
    prev = q[0].w;
    n = 3;
#pragma unroll
    for (int j = 1; j < WAM_FAST_SEARCH_BATCH; ++j) {
      mism += __popc(__funnelshift_r(prev, q[j].x, o) ^ ex[n]) + __popc(__funnelshift_r(q[j].x, q[j].y, o) ^ ex[n + 1]) +
              __popc(__funnelshift_r(q[j].y, q[j].z, o) ^ ex[n + 2]) + __popc(__funnelshift_r(q[j].z, q[j].w, o) ^ ex[n + 3]);
      prev = q[j].w;
      n += 4;
    }
    if (mism > cutoff) return mism;
  } else
#endif
  {
    v = ring_ld4(ring + g);
    g = (g + 4u) & wmask;
    nx = ring_ld4(ring + g);
    g = (g + 4u) & wmask;
    mism = __popc((__funnelshift_r(v.x, v.y, o) ^ ex[0]) & mk[0]) + __popc((__funnelshift_r(v.y, v.z, o) ^ ex[1]) & mk[1]) +
           __popc((__funnelshift_r(v.z, v.w, o) ^ ex[2]) & mk[2]);
    prev = v.w;
    n = 3;
  }
  for (; n + 4 <= n_full; n += 4) {
    v = nx;
    nx = ring_ld4(ring + g);
    g = (g + 4u) & wmask;
    mism += __popc(__funnelshift_r(prev, v.x, o) ^ ex[n]) + __popc(__funnelshift_r(v.x, v.y, o) ^ ex[n + 1]) +
            __popc(__funnelshift_r(v.y, v.z, o) ^ ex[n + 2]) + __popc(__funnelshift_r(v.z, v.w, o) ^ ex[n + 3]);
    prev = v.w;
    if (mism > cutoff) return mism;
  }
  for (; n < n_end; n += 4) {
    v = nx;
    nx = ring_ld4(ring + g);
    g = (g + 4u) & wmask;
    mism += __popc((__funnelshift_r(prev, v.x, o) ^ ex[n]) & mk[n]) +
            __popc((__funnelshift_r(v.x, v.y, o) ^ ex[n + 1]) & mk[n + 1]) +
            __popc((__funnelshift_r(v.y, v.z, o) ^ ex[n + 2]) & mk[n + 2]) +
            __popc((__funnelshift_r(v.z, v.w, o) ^ ex[n + 3]) & mk[n + 3]);
    prev = v.w;
  }
  return mism;
}
__device__ __noinline__ int sync_mismatches_fast_call(const uint32_t* __restrict__ ring, uint32_t pos, const FskDerived& d,
                                                      int cutoff) {
  if (d.tmpl_slot >= 0) return sync_mismatches0v_cut<true>(ring, pos, d, cutoff);
  return sync_mismatches0v_cut<false>(ring, pos, d, cutoff);
}

// FSKCore.processByte — fsk.ts:346-375.  Returns true when resetState() ran.
__device__ __forceinline__ bool process_byte_fast(FastB& b, int bit, const DemodArgs& a, int li, uint8_t* out_row) {
  const FskDerived& d = a.d;
  const int bp = b.bitpos;
  if (bp == 0) {
    if (bit != 0) { reset_state_fb(b); return true; }
  } else if (bp >= 1 && bp <= 8) {
    b.current |= (uint32_t)bit << (8 - bp);
  } else if (d.parity != 0 && bp == 9) {
  } else if (bp == d.stop_pos) {
    if (bit != 1) { b.started = 0; return false; }
    if (b.out_n < a.out_stride) out_row[b.out_n] = (uint8_t)b.current;
    else a.u32[(long)U_ERR * a.n_local + li] |= WAM_ERR_OUT_OVERFLOW;
    b.out_n++;
    b.current = 0;
    b.bitpos = -1;
  } else {
    b.started = 0;
    return false;
  }
  b.bitpos++;
  return false;
}

// One decimated sample of processDownsampledBit (fsk.ts:278-344) AFTER the ring puts, with doubt tracking
// (oracle/fastmodel.c: fm_decim is the same logic sample by sample).  sil / adoubt: this sample's amplitude is below
// the silence threshold / within the doubt band of it.  Returns true when resetState() ran.
__device__ __forceinline__ bool sm_step_fast(FastB& b, int bit, bool sil, bool adoubt, uint32_t ring_pos, bool ring_ready,
                                             uint32_t amp_next, uint32_t amp_len, const DemodArgs& a, int li,
                                             uint8_t* out_row, bool& thr_changed) {
  const FskDerived& d = a.d;
  const long ns = a.n_local;
  uint32_t* ring = ring_of(a, li);
  b.gsc++;
  b.gmod = (b.gmod + 1u == (uint32_t)d.check_period) ? 0u : b.gmod + 1u;
  // silence / EOD — fsk.ts:285-295.  silx: bit 31 = a doubtful compare is pending, low bits = the silent run the
  // float64 compare may have on top of ours (a doubtful sample we took for loud)
  if (adoubt) b.silx = (0x80000000u | b.silx) + (sil ? 0u : b.sil_cnt + 1u);
  else if (!sil) b.silx = 0u;
  if (sil) b.sil_cnt++;
  else b.sil_cnt = 0;
  if ((b.silx & 0x80000000u) && b.sil_cnt + (b.silx & 0x7fffffffu) >= (uint32_t)d.eod_count) {
    b.flag |= WAM_FLAG_EOD;
    b.silx = 0u;
  }
  if (sil && b.sil_cnt >= (uint32_t)d.eod_count) {
    a.u32[(long)U_EOD_EV * ns + li]++;  // emit('eod')
    reset_state_fb(b);
    return true;
  }
  if (!b.started) {
    // fsk.ts:297-328
    const bool due = d.check_period > 0 && b.gmod == 0u;
    if (due && ring_ready && d.total_bits > 0) {
      const uint32_t wmask = (uint32_t)(d.ring_words - 1);
      uint32_t* dring = a.doubt_ring + (size_t)li * (size_t)d.ring_words;
      if ((b.ring_pos & 31u) != 0u) {  // flush the register copies of the newest (partial) words
        ring_st<true>(ring + ((b.ring_pos >> 5) & wmask), b.cur_word);
        dring[(b.ring_pos >> 5) & wmask] = b.dcur_word;
      }
      // doubtful bits in the window: a cheap upper bound first, the exact count only when the bound could matter
      int D = b.dcnt != 0u ? doubt_bound(b, ring_pos, d) : 0;
      const int mism = sync_mismatches_fast_call(ring, ring_pos, d, d.max_mismatch + D);
      if (D > 0 && ((mism + D <= d.max_mismatch) != (mism - D <= d.max_mismatch))) {
        D = doubt_count(dring, ring_pos, d);
        if (D > 0 && ((mism + D <= d.max_mismatch) != (mism - D <= d.max_mismatch))) b.flag |= WAM_FLAG_SYNC;
      }
      if (mism <= d.max_mismatch) {
        b.started = 1;
        b.current = 0; b.bitpos = 0;
        b.bit_acc = 0; b.bit_cnt = 0; b.bsc = 0; b.next_idx = 0; b.dvote = 0;
        a.u32[(long)U_SYNC_DET * ns + li]++;
        b.sil_thr = (float)amp_ring_threshold(amp_of(a, li), amp_next, amp_len, (uint32_t)d.amp_phys);
        thr_changed = true;
      }
    }
    return false;
  }
  // fsk.ts:330-341 (the bulk advance has already counted this sample's doubt into dvote)
  b.bit_acc += (uint32_t)bit;
  b.bit_cnt++;
  b.bsc++;
  if (b.bsc >= b.next_idx) {
    const int decided = (2u * b.bit_acc > b.bit_cnt) ? 1 : 0;  // acc > count/2
    if (b.dvote) {
      const uint32_t d1 = b.dvote & 0xffffu, d0 = b.dvote >> 16;
      const bool lo = 2u * (b.bit_acc - d1) > b.bit_cnt, hi = 2u * (b.bit_acc + d0) > b.bit_cnt;
      if (lo != hi) {
        const int bp = b.bitpos;
        if (bp == 0) b.flag |= WAM_FLAG_VOTE_START;
        else if (bp == d.stop_pos) b.flag |= WAM_FLAG_VOTE_STOP;
        else if (bp >= 1 && bp <= 8) b.flag |= WAM_FLAG_VOTE_DATA;  // the parity bit is never looked at
      }
    }
    b.bit_acc = 0; b.bit_cnt = 0; b.dvote = 0;
    b.next_idx += (uint32_t)d.dspb;
    const bool rst = process_byte_fast(b, decided, a, li, out_row);
    if (!rst && !b.started && d.check_period > 0) b.gmod = b.gsc % (uint32_t)d.check_period;
    return rst;
  }
  return false;
}

// bulk put of `cnt` bits (chunk) at ring position p into a bit-packed ring with a register copy of the newest word
__device__ __forceinline__ void ring_put_bulk(uint32_t* ring, uint32_t wmask, uint32_t& cur, uint32_t p, uint32_t cnt,
                                              uint32_t chunk, bool hinted) {
  const uint32_t o = p & 31u;
  cur = (cur & ((1u << o) - 1u)) | (chunk << o);
  if (o + cnt >= 32u) {
    if (hinted) ring_st<true>(ring + ((p >> 5) & wmask), cur);
    else ring[(p >> 5) & wmask] = cur;
    cur = (o + cnt > 32u) ? (chunk >> (32u - o)) : 0u;
  }
}

// Event-driven state machine for one tile with doubt tracking.  bits / dmask: hard decisions and doubt flags of the
// decimated samples 0..nk-1, amp[k * 32]: their amplitudes (f32, smem).  Returns the decimated index at which
// resetState() ran, or -1.
__device__ __forceinline__ int sm_tile_events_fast(FastB& b, uint32_t bits, uint32_t dmask, const float* __restrict__ amp,
                                                   int b_from, int nk, uint32_t pos_t0, uint32_t len_t0, uint32_t slot_t0,
                                                   uint32_t alen_t0, const DemodArgs& a, int li, uint8_t* out_row) {
  const FskDerived& d = a.d;
  uint32_t* ring = ring_of(a, li);
  uint32_t* dring = a.doubt_ring + (size_t)li * (size_t)d.ring_words;
  float* aring = amp_of(a, li);
  const uint32_t wmask = (uint32_t)(d.ring_words - 1);

  uint32_t silent = 0u, adoubt = 0u;
  // ---- bulk ring puts for samples [b_from, nk) — fsk.ts:281-282
  {
    const uint32_t p = pos_t0 + (uint32_t)b_from;
    if (b_from > 0) {
      // replay pass: the words holding position p may already have been flushed
      ring_st<true>(ring + ((b.ring_pos >> 5) & wmask), b.cur_word);
      dring[(b.ring_pos >> 5) & wmask] = b.dcur_word;
      b.cur_word = ring[(p >> 5) & wmask];
      b.dcur_word = dring[(p >> 5) & wmask];
    }
    const uint32_t cnt = (uint32_t)(nk - b_from);
    const uint32_t keep = (1u << cnt) - 1u;
    ring_put_bulk(ring, wmask, b.cur_word, p, cnt, (bits >> b_from) & keep, true);
    const uint32_t dchunk = (dmask >> b_from) & keep;
    ring_put_bulk(dring, wmask, b.dcur_word, p, cnt, dchunk, false);
    if (dchunk) doubt_note(b, p, dchunk, d);
    b.ring_pos = pos_t0 + (uint32_t)nk;
    uint32_t slot = slot_t0 + (uint32_t)b_from;
    if (slot >= (uint32_t)d.amp_phys) slot -= (uint32_t)d.amp_phys;
    {
      // amplitude-ring puts (fsk.ts:282) and, from the same reads, the silence flags (fsk.ts:286) and their doubt
      // flags: positive floats order like their bit patterns, so both come from one integer difference
      const int it = __float_as_int(b.sil_thr);
      const uint32_t K = (uint32_t)d.f_amp_ulps;
      if (b_from == 0 && nk == kTile / 2 && (slot & 3u) == 0u && slot + (uint32_t)(kTile / 2) <= (uint32_t)d.amp_phys) {
#pragma unroll
        for (int q = 0; q < kTile / 8; ++q) {
          const float a0 = amp[(4 * q) * 32], a1 = amp[(4 * q + 1) * 32], a2 = amp[(4 * q + 2) * 32], a3 = amp[(4 * q + 3) * 32];
          amp_st4(aring + slot + 4 * q, make_float4(a0, a1, a2, a3));
          const int d0 = __float_as_int(a0) - it, d1 = __float_as_int(a1) - it, d2 = __float_as_int(a2) - it,
                    d3 = __float_as_int(a3) - it;
          silent |= ((d0 < 0 ? 1u : 0u) | (d1 < 0 ? 2u : 0u) | (d2 < 0 ? 4u : 0u) | (d3 < 0 ? 8u : 0u)) << (4 * q);
          adoubt |= (((uint32_t)d0 + K <= 2u * K ? 1u : 0u) | ((uint32_t)d1 + K <= 2u * K ? 2u : 0u) |
                     ((uint32_t)d2 + K <= 2u * K ? 4u : 0u) | ((uint32_t)d3 + K <= 2u * K ? 8u : 0u)) << (4 * q);
        }
      } else {
        float* p2 = aring + slot;
        int until_wrap = d.amp_phys - (int)slot;
#pragma unroll 4
        for (int k = b_from; k < nk; ++k) {
          const float av = amp[k * 32];
          amp_st(p2, av);
          ++p2;
          if (--until_wrap == 0) p2 = aring;
          const int dd = __float_as_int(av) - it;
          silent |= (dd < 0 ? 1u : 0u) << k;
          adoubt |= ((uint32_t)dd + K <= 2u * K ? 1u : 0u) << k;
        }
      }
    }
  }

  int k = b_from;
  while (k < nk) {
    // next sample at which an event can happen
    int k_evt = nk;
    {
      const uint32_t run = (uint32_t)__ffs((int)(~(silent >> k))) - 1u;  // leading silent run from k
      const uint32_t have = b.sil_cnt + ((b.silx & 0x80000000u) ? (b.silx & 0x7fffffffu) : 0u);
      const uint32_t need = (uint32_t)d.eod_count > have + 1u ? (uint32_t)d.eod_count - have - 1u : 0u;
      if (need < run) k_evt = min(k_evt, k + (int)need);
      if (adoubt >> k) k_evt = min(k_evt, k + __ffs((int)(adoubt >> k)) - 1);  // a doubtful compare is an event
    }
    if (!b.started) {
      if (d.check_period > 0) k_evt = min(k_evt, k + (int)((uint32_t)d.check_period - 1u - b.gmod));
    } else {
      const uint32_t nb = b.bsc + 1u;
      k_evt = min(k_evt, k + (int)(b.next_idx > nb ? b.next_idx - nb : 0u));
    }
    // ---- bulk advance over [k, k_evt): no doubtful amplitude in there
    const int len = k_evt - k;
    if (len > 0) {
      const uint32_t m = ((1u << len) - 1u) << k;
      b.gsc += (uint32_t)len;
      if (!b.started) b.gmod += (uint32_t)len;
      const uint32_t nz = ~silent & m;
      b.sil_cnt = nz ? (uint32_t)(k_evt - 1) - (31u - (uint32_t)__clz((int)nz)) : b.sil_cnt + (uint32_t)len;
      if (nz) b.silx = 0u;  // a certainly loud sample: every reading of the silent run restarts
      if (b.started) {
        b.bit_acc += (uint32_t)__popc(bits & m);
        b.bit_cnt += (uint32_t)len;
        b.bsc += (uint32_t)len;
      }
    }
    if (b.started && (dmask >> k)) {
      // doubtful samples of the running vote in [k, k_evt] (the event sample included)
      const uint32_t m2 = (k_evt >= 31 ? 0xffffffffu : ((2u << k_evt) - 1u)) & ~((1u << k) - 1u) & dmask;
      b.dvote += (uint32_t)__popc(bits & m2) + ((uint32_t)__popc(~bits & m2) << 16);
    }
    if (k_evt >= nk) break;
    // ---- the event sample itself
    const uint32_t pos_k = pos_t0 + (uint32_t)k_evt + 1u;
    const bool ready = len_t0 + (uint32_t)k_evt + 1u >= (uint32_t)d.total_bits;
    uint32_t slot_next = slot_t0 + (uint32_t)k_evt + 1u;
    if (slot_next >= (uint32_t)d.amp_phys) slot_next -= (uint32_t)d.amp_phys;
    const uint32_t alen = min(alen_t0 + (uint32_t)k_evt + 1u, (uint32_t)d.amp_cap);
    bool thr_changed = false;
    if (sm_step_fast(b, (int)((bits >> k_evt) & 1u), ((silent >> k_evt) & 1u) != 0u, ((adoubt >> k_evt) & 1u) != 0u, pos_k,
                     ready, slot_next, alen, a, li, out_row, thr_changed))
      return k_evt;
    if (thr_changed) {
      silent = 0u; adoubt = 0u;
      const int it = __float_as_int(b.sil_thr);
      const uint32_t K = (uint32_t)d.f_amp_ulps;
      for (int kk = k_evt + 1; kk < nk; ++kk) {
        const int dd = __float_as_int(amp[kk * 32]) - it;
        silent |= (dd < 0 ? 1u : 0u) << kk;
        adoubt |= ((uint32_t)dd + K <= 2u * K ? 1u : 0u) << kk;
      }
    }
    k = k_evt + 1;
  }
  return -1;
}

// Grid: one warp (32 streams) per CTA, TMA-staged tiles, time slabs as in fsk_demod_exact_kernel<.., STAGE_TMA>.
// Common case only (host: fast_eligible): rows contiguous and 16-byte aligned, integral sync ring, eod_count > 16,
// by-value sync template, no write-back / tap / ragged counts.
__global__ void __launch_bounds__(32, WAM_DEMOD_MIN_BLOCKS) fsk_demod_fast_kernel(const __grid_constant__ DemodLaunch L) {
  int gi = 0;
#pragma unroll
  for (int i = 1; i < kMaxGroupsPerLaunch; ++i)
    if (i < L.n_groups && (int)blockIdx.x >= L.block_begin[i]) gi = i;
  const DemodArgs& a = L.g[gi];
  __shared__ __align__(1024) float tiles[kStages][kTile * kTile];
  __shared__ __align__(8) uint64_t tma_bar[kStages];
  __shared__ __align__(128) float pfbuf[kTile * 32];  // pre-filtered samples [i][lane]
  __shared__ double park_g[32];                       // AGC gain
  __shared__ uint32_t park_u[kFParkU][32];

  const int lane = threadIdx.x;
  const int li = a.l_begin + ((int)blockIdx.x - L.block_begin[gi]) * 32 + lane;
  const bool active = li < a.l_end;
  const FskDerived& d = a.d;
  const long ns = a.n_local;
  bool slab_timeout = false;
  if (L.slab_done != nullptr && L.slab > 0) {
    // time-slab hand-over (launch_slabbed): bounded spin; on expiry the streams are flagged and left untouched
    const int* flag = L.slab_done + blockIdx.x;
    unsigned spins = 0;
    int v;
    do {
      asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
      if (v >= L.slab) break;
      __nanosleep(256);
    } while (++spins < (1u << 22));
    slab_timeout = v < L.slab;
    if (slab_timeout) {
      if (active) a.u32[(long)U_ERR * ns + li] |= WAM_ERR_SLAB_TIMEOUT;
      return;
    }
  }
  int row = -1;
  if (active) row = a.id0 + li - a.row_base;

  FastA2 s;
  reset_state_fa2(s);
  s.S = 0.0f;
  if (active) {
    const double* f = a.f64 + li;
    const uint32_t* u = a.u32 + li;
    // ---- direct form (the arrays) -> normal form
    s.lc = (float)f[F_LO_C * ns]; s.ls = (float)f[F_LO_S * ns];
    float iw1, iw2, qw1, qw2;
    df_to_normal(d.lp_b1, d.lp_b2, d.lp_a1, d.lp_a2, d.lp_nk1, d.lp_nk2, d.lp_nsg, d.lp_nom, f[F_IX1 * ns], f[F_IX2 * ns],
                 f[F_IY1 * ns], f[F_IY2 * ns], iw1, iw2);
    df_to_normal(d.lp_b1, d.lp_b2, d.lp_a1, d.lp_a2, d.lp_nk1, d.lp_nk2, d.lp_nsg, d.lp_nom, f[F_QX1 * ns], f[F_QX2 * ns],
                 f[F_QY1 * ns], f[F_QY2 * ns], qw1, qw2);
    s.w1 = make_float2(iw1, qw1); s.w2 = make_float2(iw2, qw2);
    df_to_normal(d.lp_b1, d.lp_b2, d.lp_a1, d.lp_a2, d.lp_nk1, d.lp_nk2, d.lp_nsg, d.lp_nom, f[F_OX1 * ns], f[F_OX2 * ns],
                 f[F_OY1 * ns], f[F_OY2 * ns], s.ow1, s.ow2);
    {
      double sn, cs;
      sincos(f[F_LAST_PHASE * ns], &sn, &cs);
      s.psi = (float)cs; s.psq = (float)sn;
    }
    s.acc = make_float2((float)f[F_IACC * ns], (float)f[F_QACC * ns]);
    s.S = (float)f[F_FAST_S * ns]; s.E = (float)f[F_FAST_E * ns]; s.rsp = (float)f[F_FAST_RSP * ns];
    s.dsc = u[U_DSC * ns];
    float pw1, pw2;
    df_to_normal(d.pre_b1, d.pre_b2, d.pre_a1, d.pre_a2, d.pre_nk1, d.pre_nk2, d.pre_nsg, d.pre_nom, f[F_PX1 * ns],
                 f[F_PX2 * ns], f[F_PY1 * ns], f[F_PY2 * ns], pw1, pw2);
    park_g[lane] = f[F_GAIN * ns];
    park_u[0][lane] = __float_as_uint(pw1); park_u[1][lane] = __float_as_uint(pw2);
    FastB b;
    b.sil_thr = (float)f[F_SIL_THR * ns];
    b.gsc = u[U_GSC * ns]; b.gmod = u[U_GMOD * ns]; b.bsc = u[U_BSC * ns]; b.next_idx = u[U_NEXT_IDX * ns];
    b.bit_acc = u[U_BIT_ACC * ns]; b.bit_cnt = u[U_BIT_CNT * ns]; b.started = u[U_STARTED * ns];
    b.bitpos = (int)u[U_BITPOS * ns]; b.current = u[U_CURRENT * ns]; b.sil_cnt = u[U_SIL_CNT * ns];
    b.ring_pos = u[U_RING_POS * ns]; b.ring_len = u[U_RING_LEN * ns];
    b.amp_pos = u[U_AMP_POS * ns]; b.amp_len = u[U_AMP_LEN * ns];
    b.out_n = a.append ? a.out_len[row] : 0;
    b.dvote = u[U_DVOTE * ns]; b.silx = u[U_SILX * ns]; b.dlast = u[U_LAST_DOUBT * ns]; b.flag = u[U_FLAG * ns];
    b.dcnt = u[U_DCNT * ns];
    b.cur_word = 0u; b.dcur_word = 0u;
    if ((b.ring_pos & 31u) != 0u) {
      const uint32_t wi = (b.ring_pos >> 5) & (uint32_t)(d.ring_words - 1);
      const uint32_t keep = (1u << (b.ring_pos & 31u)) - 1u;
      b.cur_word = ring_of(a, li)[wi] & keep;
      b.dcur_word = a.doubt_ring[(size_t)li * (size_t)d.ring_words + wi] & keep;
    }
    fb_store(b, park_u, lane);
  }
  __syncwarp();
  uint8_t* out_row = active ? a.out + (long)row * a.out_stride : nullptr;
  const uint32_t flag_in = active ? park_u[21][lane] : 0u;
  uint32_t n_doubt = 0u;

  const long n_tiles = (a.n + kTile - 1) / kTile;
  const int tma_row0 = a.id0 + a.l_begin + ((int)blockIdx.x - L.block_begin[gi]) * 32 - a.row_base;
  if (lane == 0) {
#pragma unroll
    for (int p = 0; p < kStages; ++p) tma_bar_init(&tma_bar[p]);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncwarp();
  for (int p = 0; p < kStages - 1; ++p)
    if (p < n_tiles && lane == 0) tma_load_tile(tiles[p], &L.tmap[gi], &tma_bar[p], p * kTile, tma_row0);
  for (long t = 0; t < n_tiles; ++t) {
    const long tn = t + kStages - 1;
    if (tn < n_tiles && lane == 0)
      tma_load_tile(tiles[tn % kStages], &L.tmap[gi], &tma_bar[tn % kStages], (int)(tn * kTile), tma_row0);
    tma_wait(&tma_bar[t % kStages], (uint32_t)((t / kStages) & 1));
    __syncwarp();
    float* tile = tiles[t % kStages];
    float* pbuf = tile;  // amplitudes [k][lane] after A1
    const long t0 = t * kTile;
    const int len = (int)min((long)kTile, a.n - t0);

    // ---------------- A1: AGC + pre-filter ----------------
    if (active) {
      double gain = park_g[lane];
      float pw1 = __uint_as_float(park_u[0][lane]), pw2 = __uint_as_float(park_u[1][lane]);
      const bool agc = d.agc_enabled != 0;
      const double att = d.agc_attack, rel = d.agc_release;
      if (len == kTile) {
#pragma unroll 1
        for (int ch = 0; ch < 8; ch += kA1Chunks) {
#pragma unroll
          for (int c = 0; c < kA1Chunks; ++c) {
            const float4 v = *reinterpret_cast<const float4*>(tile + tile_index(lane, (ch + c) * 4));
            const float p0 = fast_a1_sample(gain, pw1, pw2, v.x, d, agc, att, rel);
            const float p1 = fast_a1_sample(gain, pw1, pw2, v.y, d, agc, att, rel);
            const float p2 = fast_a1_sample(gain, pw1, pw2, v.z, d, agc, att, rel);
            const float p3 = fast_a1_sample(gain, pw1, pw2, v.w, d, agc, att, rel);
            float* pfp = pfbuf + ((ch + c) * 4) * 32 + lane;
            pfp[0] = p0; pfp[32] = p1; pfp[64] = p2; pfp[96] = p3;
          }
        }
      } else {
#pragma unroll 1
        for (int i = 0; i < len; ++i) pfbuf[i * 32 + lane] = fast_a1_sample(gain, pw1, pw2, tile[tile_index(lane, i)], d, agc, att, rel);
      }
      park_g[lane] = gain;
      park_u[0][lane] = __float_as_uint(pw1); park_u[1][lane] = __float_as_uint(pw2);
    }
    __syncwarp();  // every lane is done with the input tile; its storage becomes pbuf

    // ---------------- A2 + B with replay on resetState() ----------------
    const int dsc0 = active ? (int)s.dsc : 0;
    const int v_hi = dsc0 + len;
    const int nk = v_hi >> 1;
    int k_from = 0, b_from = 0, v_lo = dsc0;
    uint32_t bits = 0u, dmask = 0u;
    bool redo = active;
    const uint32_t pos_t0 = park_u[11][lane], len_t0 = park_u[12][lane];
    const uint32_t slot_t0 = park_u[13][lane], alen_t0 = park_u[14][lane];
    if (active) {
      // renormalise the LO rotation (one Newton step towards |(c, s)| = 1)
      const float m = fmaf(s.lc, s.lc, s.ls * s.ls);
      const float f = fmaf(-0.5f, m, 1.5f);
      s.lc *= f; s.ls *= f;
    }
    while (__any_sync(0xffffffffu, redo)) {
      if (redo) {
        const uint32_t keepm = (1u << k_from) - 1u;
        bits &= keepm; dmask &= keepm;
        if (dsc0 == 0 && (v_hi & 1) == 0) {
#pragma unroll 2
          for (int k = k_from; k < nk; ++k) {
            const float* pfp = pfbuf + (2 * k) * 32 + lane;
            const float2 y0 = fast_a2_half(s, pfp[0], d);
            const float2 y1 = fast_a2_half(s, pfp[32], d);
            uint32_t bit, dbit;
            float amp;
            fast_a2_decim(s, __fadd2_rn(y0, y1), d, bit, dbit, amp);
            bits |= bit << k; dmask |= dbit << k;
            pbuf[k * 32 + lane] = amp;
          }
        } else {
#pragma unroll 1
          for (int k = k_from; 2 * k < v_hi; ++k) {
            const int v0 = 2 * k, v1 = 2 * k + 1;
            if (v0 >= v_lo) s.acc = fast_a2_half(s, pfbuf[(v0 - dsc0) * 32 + lane], d);  // 0 + y
            if (v1 < v_hi) {
              const float2 y = fast_a2_half(s, pfbuf[(v1 - dsc0) * 32 + lane], d);
              uint32_t bit, dbit;
              float amp;
              fast_a2_decim(s, __fadd2_rn(s.acc, y), d, bit, dbit, amp);
              s.acc = make_float2(0.0f, 0.0f);
              bits |= bit << k; dmask |= dbit << k;
              pbuf[k * 32 + lane] = amp;
            }
          }
        }
        s.dsc = (uint32_t)(v_hi & 1);
        // ---------------- B ----------------
        FastB b;
        fb_load(b, park_u, lane);
        redo = false;
        const int k_reset = sm_tile_events_fast(b, bits, dmask, pbuf + lane, b_from, nk, pos_t0, len_t0, slot_t0, alen_t0,
                                                a, li, out_row);
        if (k_reset < 0 || 2 * (k_reset + 1) >= v_hi) {
          b.ring_len = min(len_t0 + (uint32_t)nk, (uint32_t)d.ring_cap_int);
          const uint32_t sl = slot_t0 + (uint32_t)nk;
          b.amp_pos = sl >= (uint32_t)d.amp_phys ? sl - (uint32_t)d.amp_phys : sl;
          b.amp_len = min(alen_t0 + (uint32_t)nk, (uint32_t)d.amp_cap);
        }
        n_doubt += (uint32_t)__popc((dmask >> b_from) & (k_reset >= 0 ? (2u << (k_reset - b_from)) - 1u : 0xffffffffu));
        if (k_reset >= 0) {
          reset_state_fa2(s);  // resetState(): A2 restarts from the zeroed state at the next pair (S is kept)
          k_from = k_reset + 1; b_from = k_reset + 1; v_lo = 2 * (k_reset + 1);
          redo = (v_lo < v_hi);
        }
        fb_store(b, park_u, lane);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    __syncwarp();
  }

  if (active) {
    double* f = a.f64 + li;
    uint32_t* u = a.u32 + li;
    // ---- normal form -> direct form (x history zero, y history carrying the state)
    f[F_LO_C * ns] = (double)s.lc; f[F_LO_S * ns] = (double)s.ls;
    double y1, y2;
    normal_to_df(d.lp_a1, d.lp_a2, d.lp_nk1, d.lp_nk2, d.lp_nsg, d.lp_nom, (double)s.w1.x, (double)s.w2.x, y1, y2);
    f[F_IX1 * ns] = 0.0; f[F_IX2 * ns] = 0.0; f[F_IY1 * ns] = y1; f[F_IY2 * ns] = y2;
    normal_to_df(d.lp_a1, d.lp_a2, d.lp_nk1, d.lp_nk2, d.lp_nsg, d.lp_nom, (double)s.w1.y, (double)s.w2.y, y1, y2);
    f[F_QX1 * ns] = 0.0; f[F_QX2 * ns] = 0.0; f[F_QY1 * ns] = y1; f[F_QY2 * ns] = y2;
    normal_to_df(d.lp_a1, d.lp_a2, d.lp_nk1, d.lp_nk2, d.lp_nsg, d.lp_nom, (double)s.ow1, (double)s.ow2, y1, y2);
    f[F_OX1 * ns] = 0.0; f[F_OX2 * ns] = 0.0; f[F_OY1 * ns] = y1; f[F_OY2 * ns] = y2;
    f[F_LAST_PHASE * ns] = atan2((double)s.psq, (double)s.psi);
    f[F_IACC * ns] = (double)s.acc.x; f[F_QACC * ns] = (double)s.acc.y;
    f[F_FAST_S * ns] = (double)s.S; f[F_FAST_E * ns] = (double)s.E; f[F_FAST_RSP * ns] = (double)s.rsp;
    u[U_DSC * ns] = s.dsc;
    normal_to_df(d.pre_a1, d.pre_a2, d.pre_nk1, d.pre_nk2, d.pre_nsg, d.pre_nom, (double)__uint_as_float(park_u[0][lane]),
                 (double)__uint_as_float(park_u[1][lane]), y1, y2);
    f[F_GAIN * ns] = park_g[lane];
    f[F_PX1 * ns] = 0.0; f[F_PX2 * ns] = 0.0; f[F_PY1 * ns] = y1; f[F_PY2 * ns] = y2;
    FastB b;
    fb_load(b, park_u, lane);
    f[F_SIL_THR * ns] = (double)b.sil_thr;
    u[U_GSC * ns] = b.gsc; u[U_GMOD * ns] = b.gmod; u[U_BSC * ns] = b.bsc; u[U_NEXT_IDX * ns] = b.next_idx;
    u[U_BIT_ACC * ns] = b.bit_acc; u[U_BIT_CNT * ns] = b.bit_cnt; u[U_STARTED * ns] = b.started;
    u[U_BITPOS * ns] = (uint32_t)b.bitpos; u[U_CURRENT * ns] = b.current; u[U_SIL_CNT * ns] = b.sil_cnt;
    u[U_RING_POS * ns] = b.ring_pos; u[U_RING_LEN * ns] = b.ring_len;
    u[U_AMP_POS * ns] = b.amp_pos; u[U_AMP_LEN * ns] = b.amp_len;
    u[U_DVOTE * ns] = b.dvote; u[U_SILX * ns] = b.silx; u[U_LAST_DOUBT * ns] = b.dlast; u[U_FLAG * ns] = b.flag;
    u[U_DCNT * ns] = b.dcnt;
    u[U_DOUBT_SAMPLES * ns] += n_doubt;
    if ((b.ring_pos & 31u) != 0u) {
      const uint32_t wi = (b.ring_pos >> 5) & (uint32_t)(d.ring_words - 1);
      ring_of(a, li)[wi] = b.cur_word;
      a.doubt_ring[(size_t)li * (size_t)d.ring_words + wi] = b.dcur_word;
    }
    a.out_len[row] = b.out_n < a.out_stride ? b.out_n : (int)a.out_stride;
    if (b.flag != 0u && flag_in == 0u) {  // first flag of this stream in this call: queue it for the float64 re-run
      u[U_FLAG_EVER * ns] |= b.flag;
      const int slot = atomicAdd(a.flag_count, 1);
      a.flag_list[slot] = li;
    }
  }
  if (L.slab_done != nullptr) slab_publish(L.slab_done + blockIdx.x, L.slab);
}

}  // namespace wam
