// wam_api.cu — C ABI of libwam.so (see include/wam.h).  Host logic + kernel launches.
// There is no CPU fallback in this file: every compute entry point launches sm_100a kernels.
#include "../../include/wam.h"

#include <algorithm>
#include <array>
#include <map>
#include <mutex>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>
#include <sched.h>
#include <cctype>
#include <cstdlib>

#include "crc_xmodem.cuh"
#include "filters.cuh"
#include "fsk_demod.cuh"
#include "fsk_demod_pipe.cuh"
#include "fsk_demod_fast.cuh"
#include "awgn.cuh"
#include "fsk_mod.cuh"
#include "wam_common.cuh"

using namespace wam;

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;

static int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}
#define CUDA_TRY(expr)                                                                                   \
  do {                                                                                                   \
    cudaError_t _e = (expr);                                                                             \
    if (_e != cudaSuccess)                                                                               \
      return fail(WAM_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));                      \
  } while (0)

extern "C" int wam_version(void) { return WAM_VERSION; }
extern "C" const char* wam_last_error(void) { return g_last_error.c_str(); }
extern "C" const char* wam_error_string(int code) {
  switch (code) {
    case WAM_OK: return "ok";
    case WAM_E_INVALID: return "invalid argument";
    case WAM_E_NOT_CONFIGURED: return "not configured";
    case WAM_E_CUDA: return "CUDA error";
    case WAM_E_NOMEM: return "out of memory";
    case WAM_E_CAPACITY: return "output buffer too small";
    case WAM_E_UNSUPPORTED: return "unsupported configuration";
    case WAM_E_FILTER_B_EMPTY: return "Feedforward coefficients (b) cannot be empty";
    case WAM_E_FILTER_A_EMPTY: return "Feedback coefficients (a) cannot be empty";
    case WAM_E_FILTER_A0_ZERO: return "First feedback coefficient (a[0]) cannot be zero";
    case WAM_E_PKT_SEQUENCE: return "Invalid sequence. Must be 1-255.";
    case WAM_E_PKT_PAYLOAD: return "Payload too large. Max 255 bytes.";
    default: return "unknown error";
  }
}
extern "C" int wam_device_count(int* count) {
  if (!count) return fail(WAM_E_INVALID, "count is NULL");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    *count = 0;
    return fail(WAM_E_CUDA, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
  }
  *count = n;
  return WAM_OK;
}

extern "C" void wam_fsk_default_config(wam_fsk_config* c) {  // fsk.ts:19-33
  static const uint8_t pre[2] = {0x55, 0x55};
  static const uint8_t sfd[1] = {0x7E};
  memset(c, 0, sizeof(*c));
  c->sampleRate = 48000; c->baudRate = 1200;
  c->markFrequency = 1650; c->spaceFrequency = 1850;
  c->preamblePattern = pre; c->preambleLength = 2;
  c->sfdPattern = sfd; c->sfdLength = 1;
  c->startBits = 1; c->stopBits = 1; c->parity = 0;
  c->syncThreshold = 0.85; c->agcEnabled = 1;
  c->preFilterBandwidth = 800; c->adaptiveThreshold = 1;
}

// ------------------------------------------------------------------------------------------
// filters.ts designs (host, float64, the reference's operation order)
// ------------------------------------------------------------------------------------------
extern "C" void wam_design_butterworth_lowpass(double fc, double fs, double b[3], double a[3]) {
  const double nyquist = fs / 2;
  const double nc = fc / nyquist;
  const double c = tan(M_PI * nc / 2);
  const double c2 = c * c;
  const double s2c = M_SQRT2 * c;
  const double den = 1 + s2c + c2;
  b[0] = c2 / den; b[1] = 2 * c2 / den; b[2] = c2 / den;
  a[0] = 1; a[1] = (2 * c2 - 2) / den; a[2] = (1 - s2c + c2) / den;
}
extern "C" void wam_design_butterworth_highpass(double fc, double fs, double b[3], double a[3]) {
  const double nyquist = fs / 2;
  const double nc = fc / nyquist;
  const double c = tan(M_PI * nc / 2);
  const double c2 = c * c;
  const double s2c = M_SQRT2 * c;
  const double den = 1 + s2c + c2;
  b[0] = 1 / den; b[1] = -2 / den; b[2] = 1 / den;
  a[0] = 1; a[1] = (2 * c2 - 2) / den; a[2] = (1 - s2c + c2) / den;
}
extern "C" void wam_design_butterworth_bandpass(double f0, double bandwidth, double fs, double b[3], double a[3]) {
  const double omega = 2 * M_PI * f0 / fs;
  const double bw = 2 * M_PI * bandwidth / fs;
  const double c = tan(bw / 2);
  const double dd = 2 * cos(omega);
  const double c2 = c * c;
  const double den = 1 + c + c2;
  b[0] = c / den; b[1] = 0; b[2] = -c / den;
  a[0] = 1; a[1] = (-dd * (1 + c2)) / den; a[2] = (1 - c + c2) / den;
}
extern "C" int wam_design_sinc_lowpass(double fc, double fs, int numTaps, double* out) {
  if (numTaps % 2 == 0) numTaps++;  // filters.ts:244-246
  const double nc = fc / fs;
  const double center = (numTaps - 1) / 2.0;
  for (int i = 0; i < numTaps; i++) {
    if ((double)i == center) {
      out[i] = 2 * nc;
    } else {
      const double x = M_PI * (i - center);
      out[i] = sin(2 * nc * x) / x;
    }
    out[i] *= 0.54 - 0.46 * cos(2 * M_PI * i / (numTaps - 1));
  }
  return numTaps;
}
extern "C" int wam_design_sinc_highpass(double fc, double fs, int numTaps, double* out) {
  // filters.ts:274-286.  With an even numTaps the low-pass has numTaps+1 entries, only the first
  // numTaps are negated and `lowpass[center] += 1` addresses a fractional index (no element).
  const int n = wam_design_sinc_lowpass(fc, fs, numTaps, out);
  const double center = (numTaps - 1) / 2.0;
  for (int i = 0; i < numTaps; i++) out[i] = -out[i];
  if (center == floor(center) && center >= 0 && center < n) out[(int)center] += 1;
  return n;
}
extern "C" int wam_design_sinc_bandpass(double f0, double bandwidth, double fs, int numTaps, double* out) {
  // filters.ts:296-314: truncated convolution of the high-pass and low-pass prototypes
  const double lo = f0 - bandwidth / 2;
  const double hi = f0 + bandwidth / 2;
  std::vector<double> hp((size_t)numTaps + 2), lp((size_t)numTaps + 2);
  wam_design_sinc_highpass(lo, fs, numTaps, hp.data());
  wam_design_sinc_lowpass(hi, fs, numTaps, lp.data());
  for (int i = 0; i < numTaps; i++) out[i] = 0;
  for (int i = 0; i < numTaps; i++)
    for (int j = 0; j < numTaps; j++)
      if (i + j < numTaps) out[i + j] += hp[i] * lp[j];
  return numTaps;
}

// ------------------------------------------------------------------------------------------
// FSKCore.configure(): derived parameters (fsk.ts:133-173, :426-462)
// ------------------------------------------------------------------------------------------
// Fast path (fsk_demod_fast.cuh): the biquads in normal form — poles sg +- j om, y = b0 x + k1 w1 + k2 w2 with
// k1 = b1 - b0 a1, k2 = (b2 - b0 a2 + k1 sg) / om — and the constants of the doubt band.  The band's constants were
// calibrated on the CPU model of the kernel (scripts/exp_fastmodel.py): over 16,384 noisy streams the float32 error of
// filteredPhaseDiff stayed below 0.23 of the band.
static bool normal_form(double b0, double b1, double b2, double a1, double a2, double& k1, double& k2, double& sg, double& om) {
  sg = -a1 / 2;
  const double om2 = a2 - sg * sg;
  if (!(om2 > 1e-12) || !(fabs(a2) > 1e-12) || !std::isfinite(om2)) return false;  // real or degenerate poles
  om = sqrt(om2);
  k1 = b1 - b0 * a1;
  k2 = (b2 - b0 * a2 + k1 * sg) / om;
  return std::isfinite(k1) && std::isfinite(k2);
}
// the doubt band's constants as the kernel uses them (fast_decim)
static void fold_band(FskDerived& d) {
  d.f_gk2 = d.f_gamma * d.f_kappa * 0.5f;  // gamma e1 = (gamma kappa / 2) S (1 / (2 amp) + 1 / (2 amp_prev))
  d.f_eps0r = (float)((double)d.f_eps0 * (1.0 - (double)d.f_rho_e));
  d.f_4og = 4.0f / d.f_gamma;
  d.f_bc_thr = d.f_bc_delta - d.f_4og * d.f_eps0r;
  d.f_g63 = 6.3f * d.f_gamma;
}
static void derive_fast(FskDerived& d) {
  bool ok = normal_form(d.pre_b0, d.pre_b1, d.pre_b2, d.pre_a1, d.pre_a2, d.pre_nk1, d.pre_nk2, d.pre_nsg, d.pre_nom);
  ok = normal_form(d.lp_b0, d.lp_b1, d.lp_b2, d.lp_a1, d.lp_a2, d.lp_nk1, d.lp_nk2, d.lp_nsg, d.lp_nom) && ok;
  if (!ok) { d.fast_ok = 0; return; }
  d.f_pre_k0 = (float)d.pre_b0; d.f_pre_k1 = (float)d.pre_nk1; d.f_pre_k2 = (float)d.pre_nk2;
  d.f_pre_sg = (float)d.pre_nsg; d.f_pre_om = (float)d.pre_nom;
  d.f_lp_k0 = (float)d.lp_b0; d.f_lp_k1 = (float)d.lp_nk1; d.f_lp_k2 = (float)d.lp_nk2;
  d.f_lp_sg = (float)d.lp_nsg; d.f_lp_om = (float)d.lp_nom;
  {  // two samples per step
    const double sg = d.pre_nsg, om = d.pre_nom, k1 = d.pre_nk1, k2 = d.pre_nk2;
    const float A = (float)(sg * sg - om * om), B = (float)(2 * sg * om);
    d.f2_pre_ks0 = make_float2((float)d.pre_b0, (float)k1); d.f2_pre_ks1 = make_float2(0.0f, (float)d.pre_b0);
    d.f2_pre_kw1 = make_float2((float)k1, (float)(k1 * sg + k2 * om));
    d.f2_pre_kw2 = make_float2((float)k2, (float)(k2 * sg - k1 * om));
    d.f2_pre_as0 = make_float2((float)sg, (float)om); d.f2_pre_as1 = make_float2(1.0f, 0.0f);
    d.f2_pre_aw1 = make_float2(A, B); d.f2_pre_aw2 = make_float2(-B, A);
  }
  {
    const double sg = d.lp_nsg, om = d.lp_nom, k0 = d.lp_b0, k1 = d.lp_nk1, k2 = d.lp_nk2;
    d.f_lp_A = (float)(sg * sg - om * om); d.f_lp_B = (float)(2 * sg * om);
    d.f_lp_kx0 = (float)(k0 + k1);
    d.f_lp_kw1 = (float)(k1 * (1 + sg) + k2 * om); d.f_lp_kw2 = (float)(k2 * (1 + sg) - k1 * om);
    d.f2_lp_pw1 = make_float2((float)sg, (float)om); d.f2_lp_pw2 = make_float2(-(float)om, (float)sg);
  }
  d.f_cw = (float)d.cos_omega; d.f_sw = (float)d.sin_omega;
  d.f_c2w = (float)cos(2 * d.omega); d.f_s2w = (float)sin(2 * d.omega);
  d.f_dphi_bias = (float)remainder(atan2((double)d.f_s2w, (double)d.f_c2w) - 2 * d.omega, 2 * M_PI);
  const double rho = sqrt(d.lp_a2);
  d.f_rho_e = (float)rho;
  d.f_gamma = (float)(sqrt(d.lp_nk1 * d.lp_nk1 + d.lp_nk2 * d.lp_nk2) / rho);
  d.f_kappa = 3e-7f;
  d.f_eps0 = 3e-7f;
  d.f_bc_delta = 2e-6f;
  d.f_amp_eps = 2e-6f;
  fold_band(d);
  // frame-search prefilter (fsk_demod_fast.cuh): sub-block i (i = 0 newest) covers the samples i * cp .. (i + 1) * cp - 1
  // back from the check instant, i.e. window j = i / 4 of fsk.ts:303-312 with expected bit pattern[nbits - j] (j = 0
  // is never compared)
  d.sub_ok = (d.check_period > 0 && d.dspb == 4 * d.check_period && 4 * d.nbits <= 128) ? 1 : 0;
  d.sub_half = d.check_period / 2;
  d.sub_blocks = 4 * d.nbits;
  if (d.sub_ok)
    for (int i = 4; i < 4 * d.nbits; i++) {
      const int j = i / 4, k = d.nbits - j;
      d.sub_mask[i >> 5] |= 1u << (i & 31);
      if ((d.pattern[k >> 5] >> (k & 31)) & 1u) d.sub_expect[i >> 5] |= 1u << (i & 31);
    }
  d.fast_ok = (!d.ring_fractional && d.eod_count > 16 && d.total_bits > 0 && d.check_period > 0 && rho < 1.0 &&
               d.amp_phys % 16 == 0) ? 1 : 0;
}

static int derive(const wam_fsk_config& c, FskDerived& d) {
  memset(&d, 0, sizeof(d));
  d.tmpl_slot = -1;
  if (!(c.sampleRate > 0) || !(c.baudRate > 0) || !std::isfinite(c.sampleRate) || !std::isfinite(c.baudRate) ||
      !std::isfinite(c.markFrequency) || !std::isfinite(c.spaceFrequency))
    return fail(WAM_E_UNSUPPORTED, "sampleRate/baudRate/frequencies must be finite and positive");
  if (c.preambleLength < 0 || c.sfdLength < 0 || (c.preambleLength > 0 && !c.preamblePattern) ||
      (c.sfdLength > 0 && !c.sfdPattern))
    return fail(WAM_E_INVALID, "bad preamble/sfd pattern");
  if (c.startBits < 0 || c.stopBits < 0 || c.startBits > 16 || c.stopBits > 16)
    return fail(WAM_E_UNSUPPORTED, "startBits/stopBits must be in 0..16");
  if (c.parity < 0 || c.parity > 2) return fail(WAM_E_INVALID, "parity must be 0 (none), 1 (even) or 2 (odd)");

  const double fs = c.sampleRate, baud = c.baudRate;
  const double downsampleRate = fs / 2;
  const double center = (c.markFrequency + c.spaceFrequency) / 2;
  const double spb = floor(fs / baud);
  const double dspb = floor(downsampleRate / baud);
  const int bpb = 8 + c.startBits + c.stopBits + (c.parity != 0 ? 1 : 0);
  if (spb < 1 || dspb < 1 || spb > 1e6) return fail(WAM_E_UNSUPPORTED, "samples per bit out of range (need fs/2 >= baud)");
  d.spb = (int)spb; d.dspb = (int)dspb; d.bpb = bpb;
  d.start_bits = c.startBits; d.stop_bits = c.stopBits; d.parity = c.parity;
  d.stop_pos = c.parity == 0 ? 9 : 10;

  d.agc_enabled = c.agcEnabled ? 1 : 0;
  d.agc_attack = 1.0 - exp(-1.0 / (fs * 0.001));
  d.agc_release = 1.0 - exp(-1.0 / (fs * 0.01));

  const double freqSpan = fabs(c.spaceFrequency - c.markFrequency);
  const double deviation = freqSpan / 2;
  const double carson = 2 * (deviation + baud);
  const double bw = (c.preFilterBandwidth != c.preFilterBandwidth || carson != carson)
                        ? NAN
                        : (c.preFilterBandwidth > carson ? c.preFilterBandwidth : carson);
  double b[3], a[3];
  wam_design_butterworth_bandpass(center, bw, fs, b, a);
  d.pre_b0 = b[0]; d.pre_b1 = b[1]; d.pre_b2 = b[2]; d.pre_a1 = a[1]; d.pre_a2 = a[2];
  wam_design_butterworth_lowpass(baud, fs, b, a);
  d.lp_b0 = b[0]; d.lp_b1 = b[1]; d.lp_b2 = b[2]; d.lp_a1 = a[1]; d.lp_a2 = a[2];

  d.omega = 2 * M_PI * center / fs;
  if (!(d.omega >= 0) || !(d.omega < 2 * M_PI)) return fail(WAM_E_UNSUPPORTED, "center frequency must be in [0, sampleRate)");
  d.cos_omega = cos(d.omega);
  d.sin_omega = sin(d.omega);

  // preambleSfdBits (fsk.ts:143-144, :159-173)
  const int nbytes = c.preambleLength + c.sfdLength;
  const long nbits = (long)nbytes * bpb;
  if (nbits > kMaxPatternWords * 32 || nbytes > (int)sizeof(d.preamble_sfd))
    return fail(WAM_E_UNSUPPORTED, "preamble+sfd longer than 256 line bits / 32 bytes");
  d.nbits = (int)nbits;
  d.n_preamble = c.preambleLength; d.n_sfd = c.sfdLength;
  int k = 0;
  for (int i = 0; i < nbytes; i++) {
    const int byte = i < c.preambleLength ? c.preamblePattern[i] : c.sfdPattern[i - c.preambleLength];
    d.preamble_sfd[i] = (uint8_t)byte;
    auto push = [&](int bit) {
      if (bit) d.pattern[k >> 5] |= 1u << (k & 31);
      k++;
    };
    for (int s = 0; s < c.startBits; s++) push(0);
    for (int s = 7; s >= 0; s--) push((byte >> s) & 1);
    if (c.parity != 0) {
      int p = 0;
      for (int s = 0; s < 8; s++) p ^= (byte >> s) & 1;
      push(c.parity == 1 ? p : 1 - p);
    }
    for (int s = 0; s < c.stopBits; s++) push(1);
  }
  d.total_bits = d.nbits * d.dspb;
  d.check_period = (int)floor(dspb / 4 + 0.5);  // Math.round
  const double samplesForEOD = (double)bpb * dspb * 0.7;
  d.eod_count = (int)ceil(samplesForEOD);
  // smallest matched with matched / total > syncThreshold (strict, float64 division as in JS)
  d.min_matched = INT_MAX;
  if (d.total_bits > 0 && c.syncThreshold == c.syncThreshold) {
    const double total = (double)d.total_bits;
    long lo = 0, hi = (long)d.total_bits + 1;  // first m in [0, total] with m/total > thr, else total+1
    while (lo < hi) {
      const long mid = (lo + hi) / 2;
      if ((double)mid / total > c.syncThreshold) hi = mid;
      else lo = mid + 1;
    }
    if (lo <= d.total_bits) d.min_matched = (int)lo;
  }

  const double maxSyncBits = (double)d.nbits + 32;
  d.ring_cap = maxSyncBits * dspb * 1.1;
  d.ring_cap_int = (int)trunc(d.ring_cap);
  d.ring_fractional = (d.ring_cap != floor(d.ring_cap)) ? 1 : 0;
  if (d.ring_cap > 5e7) return fail(WAM_E_UNSUPPORTED, "sync ring too large");
  if (!d.ring_fractional) {
    int words = 4;
    while ((long)words * 32 < (long)d.total_bits + 64) words <<= 1;
    d.ring_words = words;
  } else {
    d.ring_words = (d.ring_cap_int + 31) / 32 + 1;
  }
  d.amp_cap = d.dspb * 8;
  d.amp_phys = d.amp_cap + 32;
  d.mark = c.markFrequency; d.space = c.spaceFrequency; d.fs = fs;
  derive_fast(d);
  return WAM_OK;
}

// ------------------------------------------------------------------------------------------
// batch object
// ------------------------------------------------------------------------------------------
// Device buffers of the fast path, per configuration group (used by fast_host.inl).
constexpr int kVerifyClasses = 4;    // verification windows of 1..4 time slabs
constexpr int kFastSlabTiles = 60;   // time slab of the fast path = the verification windows' grain (halving it buys nothing: measured)
constexpr int kStageE1 = kVerifyClasses, kStageE2 = kVerifyClasses + 1, kAllClasses = kVerifyClasses + 2;  // two-stage end-of-data check
constexpr int kSlabStreams = 16;
constexpr int kVerifyCap = 1024;     // windows per class and call; the excess is re-run over the whole call
struct FastBuffers {
  // checkpoints 1..S of the per-stream state ([slab][field][stream]); checkpoint 0 is the live state
  double* ck_f64 = nullptr; size_t ck_f64_bytes = 0;
  uint32_t* ck_u32 = nullptr; size_t ck_u32_bytes = 0;
  // per-call linear histories: hard bits (one half word per tile) and amplitudes, each behind a prefix taken from the rings
  uint16_t* bit_hist = nullptr; size_t bit_hist_bytes = 0;
  float* amp_hist = nullptr; size_t amp_hist_bytes = 0;
  // streams flagged per time slab (li | cause << 24), their counts
  int32_t* slab_list = nullptr; size_t slab_list_bytes = 0;
  int32_t* slab_count = nullptr; size_t slab_count_bytes = 0;
  // streams to re-run over the whole call (hard list, no duplicates: hard_mark), statistics of the last call
  int32_t* hard_list = nullptr; size_t hard_list_bytes = 0;
  int32_t* hard_count = nullptr;   // [0] hard streams, [1] windows verified, [2] windows that failed, [3] windows dropped
  uint32_t* hard_mark = nullptr; size_t hard_mark_bytes = 0;
  // verification scratch, per class c (window of c + 1 slabs): items, state, rings, samples, output
  int32_t* item_li = nullptr; int32_t* item_slab = nullptr; int32_t* item_count = nullptr; int32_t* iota = nullptr;
  int32_t* item_res = nullptr;  // per window: 1 = the fast results stand, else 1 << 30 | bits naming what differed
  double* sv_f64 = nullptr; uint32_t* sv_u32 = nullptr; uint32_t* sv_ring = nullptr; float* sv_amp = nullptr;
  float* sv_samples = nullptr; size_t sv_samples_bytes = 0;
  uint8_t* sv_out = nullptr; size_t sv_out_bytes = 0;
  int32_t* sv_out_len = nullptr;
  bool scratch_ready = false;
  // carry: streams that ended the last fast call with float32 readings still undecided (a running vote, a silent run,
  // ring bits inside the doubt band) keep the call's last slabs — state, rings, samples — so that the next call can
  // begin by re-reading them in float64 (fast_carry_settle)
  int32_t* carry_li = nullptr; int32_t* carry_count = nullptr;  // carry_count[0] entries, [1] settled so far, [2] corrected so far
  int32_t* cv_iota = nullptr;
  double* cv_f64 = nullptr; uint32_t* cv_u32 = nullptr; uint32_t* cv_ring = nullptr; float* cv_amp = nullptr;
  float* cv_samples = nullptr; size_t cv_samples_bytes = 0;
  uint8_t* cv_out = nullptr; size_t cv_out_bytes = 0;
  int32_t* cv_out_len = nullptr;
  int carry_cap = 0;
  long carry_n = 0, carry_out_stride = 0;  // samples per entry, bytes of scratch output per entry
  bool carry_pending = false;
  void release() {
    for (void* q : {(void*)carry_li, (void*)carry_count, (void*)cv_iota, (void*)cv_f64, (void*)cv_u32, (void*)cv_ring, (void*)cv_amp,
                    (void*)cv_samples, (void*)cv_out, (void*)cv_out_len})
      if (q) cudaFree(q);
    for (void* q : {(void*)ck_f64, (void*)ck_u32, (void*)bit_hist, (void*)amp_hist, (void*)slab_list, (void*)slab_count,
                    (void*)hard_list, (void*)hard_count, (void*)hard_mark, (void*)item_li, (void*)item_slab,
                    (void*)item_count, (void*)iota, (void*)sv_f64, (void*)sv_u32, (void*)sv_ring, (void*)sv_amp,
                    (void*)sv_samples, (void*)sv_out, (void*)sv_out_len})
      if (q) cudaFree(q);
    *this = FastBuffers();
  }
};

struct Group {
  FskDerived d;
  wam_fsk_config cfg;
  std::vector<int32_t> ids;  // sorted global stream ids
  bool contiguous = true;
  int32_t* d_ids = nullptr;
  double* f64 = nullptr;
  uint32_t* u32 = nullptr;
  uint32_t* sync_ring = nullptr;
  float* amp_ring = nullptr;
  uint32_t* tmpl = nullptr;  // expect[32][W] then mask[32][W]
  FastBuffers fb;  // fast path: checkpoints, histories, flag lists, verification scratch (fast_host.inl)
  int doubt_state = 0;  // 0: clean (fresh streams), 1: maintained by the fast kernel, 2: stale (an exact kernel ran since)
  bool unaligned = false;  // some call since reset() was not a whole number of 32-sample tiles: fast path closed
};

struct wam_fsk_batch {
  int device = 0;
  int sm_count = 148;
  int32_t* stage_nvalid = nullptr;
  size_t stage_nvalid_bytes = 0;
  int fused_per_sm = -1;          // resident CTAs per SM of fsk_demod_exact_kernel<true, false, true>, -1 = not asked yet
  int fast_per_sm = -1;           // the same for fsk_demod_fast_kernel
  int pipe_per_sm_thin = -1;  // the same for the thin-compare variant of the aligned kernel (verification launches)
  int pipe_per_sm[2] = {-1, -1};  // resident CTAs per SM of fsk_demod_pipe_kernel<unaligned / aligned>, -1 = not asked yet
  size_t pipe_smem = 0;
  int pipe_ring_smem = 0;
  long n_streams = 0;
  std::vector<Group> groups;
  std::vector<int32_t> stream_group, stream_local;
  // host counters (identical for every stream of the batch: fsk.ts:195-196)
  double demodulation_calls = 0, total_samples = 0, configured_events = 1;
  long launches = 0;
  // staging for the HOST-buffer entry points
  cudaStream_t streams[2] = {nullptr, nullptr};
  cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_consumed[2] = {nullptr, nullptr};
  float* stage_samples[2] = {nullptr, nullptr};
  int16_t* stage_pcm[2] = {nullptr, nullptr};  // 16-bit PCM staging of wam_fsk_batch_demodulate_pcm16
  size_t stage_pcm_bytes = 0;
  uint8_t* stage_out[2] = {nullptr, nullptr};
  int32_t* stage_len[2] = {nullptr, nullptr};
  size_t stage_samples_bytes = 0, stage_out_bytes = 0, stage_len_bytes = 0;
  long fast_calls = 0, fast_launches = 0;
  unsigned long long* phase_cycles = nullptr;  // debug: [max CTAs][4]
  // time slabs of the fused kernel: two streams whose launches overlap, fork / join events, per-CTA progress flags
  cudaStream_t slab_streams[kSlabStreams] = {};
  cudaEvent_t slab_fork = nullptr, slab_join[kSlabStreams] = {};
  int* slab_done = nullptr;
  size_t slab_done_bytes = 0;
  long phase_ctas = 0;
  // modulator scratch
  uint32_t* mod_prefix = nullptr;
  size_t mod_prefix_bytes = 0;
  uint8_t* mod_data = nullptr;
  size_t mod_data_bytes = 0;
  float* mod_out = nullptr;
  size_t mod_out_bytes = 0;
  int32_t* mod_len = nullptr;
  size_t mod_len_bytes = 0;
};

// atan(k / 64) table shared by every kernel launch on a device
static int atan_table_device(int device, const double2** out) {
  static double2* tabs[64] = {nullptr};
  if (device < 0 || device >= 64) return fail(WAM_E_INVALID, "device index out of range");
  if (!tabs[device]) {
    double2 h[kAtanTableSize];
    for (int k = 0; k < kAtanTableSize; k++) { h[k].x = (double)k / 64.0; h[k].y = atan((double)k / 64.0); }
    double2* dptr = nullptr;
    CUDA_TRY(cudaMalloc(&dptr, sizeof(h)));
    CUDA_TRY(cudaMemcpy(dptr, h, sizeof(h), cudaMemcpyHostToDevice));
    tabs[device] = dptr;
  }
  *out = tabs[device];
  return WAM_OK;
}

// __constant__ template slots (c_tmpl in fsk_demod.cuh): per device, shared by content, reference counted.
namespace {
constexpr int kTmplWordsPerSlot = 2 * (4 + wam::kTmpl0Words + 4);
struct TmplSlot { std::vector<uint32_t> content; int refs = 0; };
std::mutex g_tmpl_mu;
std::map<int, std::array<TmplSlot, wam::kTmplSlots>> g_tmpl_slots;  // by device
}  // namespace
// Returns a slot holding `content` on the current device (uploading it if needed), or -1 when all slots are taken.
static int tmpl_slot_acquire(int device, const std::vector<uint32_t>& content) {
  std::lock_guard<std::mutex> lk(g_tmpl_mu);
  auto& slots = g_tmpl_slots[device];
  int free_slot = -1;
  for (int i = 0; i < wam::kTmplSlots; i++) {
    if (slots[i].refs > 0 && slots[i].content == content) { slots[i].refs++; return i; }
    if (slots[i].refs == 0 && free_slot < 0) free_slot = i;
  }
  if (free_slot < 0) return -1;
  if (cudaMemcpyToSymbol(wam::c_tmpl, content.data(), sizeof(uint32_t) * kTmplWordsPerSlot,
                         sizeof(uint32_t) * kTmplWordsPerSlot * (size_t)free_slot) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  slots[free_slot].content = content;
  slots[free_slot].refs = 1;
  return free_slot;
}
static void tmpl_slot_release(int device, int slot) {
  if (slot < 0) return;
  std::lock_guard<std::mutex> lk(g_tmpl_mu);
  auto it = g_tmpl_slots.find(device);
  if (it != g_tmpl_slots.end() && it->second[slot].refs > 0) it->second[slot].refs--;
}

// Word-aligned frame-sync templates (see sync_mismatches in fsk_demod.cuh).
static int build_sync_templates(Group& g, int device) {
  FskDerived& d = g.d;
  d.tmpl_expect = nullptr; d.tmpl_mask = nullptr; d.tmpl_words = 0; d.max_mismatch = -1; d.tmpl0_words = 0; d.tmpl0_full = 0; d.tmpl_slot = -1;
  if (d.ring_fractional || d.total_bits <= 0) return WAM_OK;
  const int care = d.total_bits - d.dspb;  // the newest dspb samples (j == 0) never match
  const int W = (31 + care + 31) / 32;
  std::vector<uint32_t> h((size_t)2 * 32 * W, 0u);
  for (int o = 0; o < 32; o++) {
    for (int idx = 0; idx < care; idx++) {
      const int back = d.total_bits - 1 - idx;  // samples back from the newest
      const int j = back / d.dspb;              // 1 .. nbits-1
      const int pb = d.nbits - j;
      const uint32_t expect = (d.pattern[pb >> 5] >> (pb & 31)) & 1u;
      const int bitpos = o + idx;
      h[(size_t)o * W + (bitpos >> 5)] |= expect << (bitpos & 31);
      h[(size_t)(32 + o) * W + (bitpos >> 5)] |= 1u << (bitpos & 31);
    }
  }
  CUDA_TRY(cudaMalloc(&g.tmpl, sizeof(uint32_t) * h.size()));
  CUDA_TRY(cudaMemcpy(g.tmpl, h.data(), sizeof(uint32_t) * h.size(), cudaMemcpyHostToDevice));
  d.tmpl_expect = g.tmpl;
  d.tmpl_mask = g.tmpl + (size_t)32 * W;
  d.tmpl_words = W;
  const int W0 = (care + 31) / 32;  // offset 0: bit idx of the window sits at bit idx
  d.tmpl0_words = 0;
  if (W0 <= kTmpl0Words) {
    d.tmpl0_words = W0;
    d.tmpl0_full = care / 32;
    for (int i = 0; i < 4 + kTmpl0Words + 4; i++) {
      const int wi = i - 4;
      d.tmpl0_expect[i] = (wi >= 0 && wi < W0) ? h[(size_t)wi] : 0u;
      d.tmpl0_mask[i] = (wi >= 0 && wi < W0) ? h[(size_t)32 * W + wi] : 0u;
    }
    std::vector<uint32_t> content(d.tmpl0_expect, d.tmpl0_expect + 4 + kTmpl0Words + 4);
    content.insert(content.end(), d.tmpl0_mask, d.tmpl0_mask + 4 + kTmpl0Words + 4);
    d.tmpl_slot = tmpl_slot_acquire(device, content);
  }
  d.max_mismatch = (d.min_matched == INT_MAX) ? -1 : care - d.min_matched;
  return WAM_OK;
}

__global__ void fill_f64_kernel(double* p, double v, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// State of freshly constructed + configured FSKCore instances, stream-ordered.
static int init_group_state(Group& g, cudaStream_t st) {
  const size_t n = g.ids.size();
  CUDA_TRY(cudaMemsetAsync(g.f64, 0, sizeof(double) * F64_COUNT * n, st));
  CUDA_TRY(cudaMemsetAsync(g.u32, 0, sizeof(uint32_t) * U32_COUNT * n, st));
  CUDA_TRY(cudaMemsetAsync(g.sync_ring, 0, sizeof(uint32_t) * (size_t)g.d.ring_words * n, st));
  CUDA_TRY(cudaMemsetAsync(g.amp_ring, 0, sizeof(float) * (size_t)g.d.amp_phys * n, st));
  g.doubt_state = 0;
  g.unaligned = false;
  g.fb.carry_pending = false;  // fresh streams carry nothing
  // AGC gain 1.0 (fsk.ts:46), silence threshold 0.01 (fsk.ts:128)
  const unsigned blocks = (unsigned)((n + 255) / 256);
  fill_f64_kernel<<<blocks, 256, 0, st>>>(g.f64 + (size_t)F_GAIN * n, 1.0, (long)n);
  fill_f64_kernel<<<blocks, 256, 0, st>>>(g.f64 + (size_t)F_SIL_THR * n, 0.01, (long)n);
  fill_f64_kernel<<<blocks, 256, 0, st>>>(g.f64 + (size_t)F_LO_C * n, 1.0, (long)n);  // cos(0)
  CUDA_TRY(cudaGetLastError());
  return WAM_OK;
}

static void free_batch(wam_fsk_batch* b) {
  if (!b) return;
  cudaSetDevice(b->device);
  for (auto& g : b->groups) {
    cudaFree(g.d_ids); cudaFree(g.f64); cudaFree(g.u32); cudaFree(g.sync_ring); cudaFree(g.amp_ring); cudaFree(g.tmpl);
    g.fb.release();
    tmpl_slot_release(b->device, g.d.tmpl_slot);
  }
  for (int i = 0; i < 2; i++) {
    if (b->streams[i]) cudaStreamDestroy(b->streams[i]);
    if (b->ev_copied[i]) cudaEventDestroy(b->ev_copied[i]);
    if (b->ev_consumed[i]) cudaEventDestroy(b->ev_consumed[i]);
    cudaFree(b->stage_samples[i]); cudaFree(b->stage_out[i]); cudaFree(b->stage_len[i]);
    cudaFree(b->stage_pcm[i]);
  }
  cudaFree(b->phase_cycles);
  cudaFree(b->slab_done);
  for (int i = 0; i < kSlabStreams; i++) {
    if (b->slab_streams[i]) cudaStreamDestroy(b->slab_streams[i]);
    if (b->slab_join[i]) cudaEventDestroy(b->slab_join[i]);
  }
  if (b->slab_fork) cudaEventDestroy(b->slab_fork);
  cudaFree(b->stage_nvalid);
  cudaFree(b->mod_prefix); cudaFree(b->mod_data); cudaFree(b->mod_out); cudaFree(b->mod_len);
  delete b;
}

extern "C" int wam_fsk_batch_create(int device, long n_streams, const wam_fsk_config* cfgs, int n_cfgs,
                                    const int32_t* cfg_index, wam_fsk_batch** out) {
  if (!out) return fail(WAM_E_INVALID, "out is NULL");
  *out = nullptr;
  if (n_streams <= 0 || n_streams > INT_MAX / 2) return fail(WAM_E_INVALID, "n_streams out of range");
  if (!cfgs || n_cfgs <= 0) return fail(WAM_E_INVALID, "no configuration given");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(WAM_E_CUDA, "no CUDA device available (libwam has no CPU fallback)");
  if (device < 0 || device >= ndev) return fail(WAM_E_INVALID, "device index out of range");
  CUDA_TRY(cudaSetDevice(device));

  wam_fsk_batch* b = new (std::nothrow) wam_fsk_batch();
  if (!b) return fail(WAM_E_NOMEM, "host allocation failed");
  b->device = device;
  if (cudaDeviceGetAttribute(&b->sm_count, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || b->sm_count < 1)
    b->sm_count = 148;
  b->n_streams = n_streams;
  b->groups.resize((size_t)n_cfgs);
  for (int i = 0; i < n_cfgs; i++) {
    b->groups[(size_t)i].cfg = cfgs[i];
    int rc = derive(cfgs[i], b->groups[(size_t)i].d);
    if (rc != WAM_OK) { free_batch(b); return rc; }
  }
  b->stream_group.resize((size_t)n_streams);
  b->stream_local.resize((size_t)n_streams);
  for (long s = 0; s < n_streams; s++) {
    const int gi = cfg_index ? cfg_index[s] : 0;
    if (gi < 0 || gi >= n_cfgs) { free_batch(b); return fail(WAM_E_INVALID, "cfg_index out of range"); }
    Group& g = b->groups[(size_t)gi];
    b->stream_group[(size_t)s] = gi;
    b->stream_local[(size_t)s] = (int32_t)g.ids.size();
    g.ids.push_back((int32_t)s);
  }
  for (auto& g : b->groups) {
    g.cfg.preamblePattern = nullptr; g.cfg.sfdPattern = nullptr;  // bytes live in d.preamble_sfd
    const size_t n = g.ids.size();
    if (n == 0) continue;
    g.contiguous = (size_t)(g.ids.back() - g.ids.front() + 1) == n;
    cudaError_t e = cudaSuccess;
    if (!g.contiguous) {
      e = cudaMalloc(&g.d_ids, sizeof(int32_t) * n);
      if (e == cudaSuccess) e = cudaMemcpy(g.d_ids, g.ids.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess) e = cudaMalloc(&g.f64, sizeof(double) * F64_COUNT * n);
    if (e == cudaSuccess) e = cudaMalloc(&g.u32, sizeof(uint32_t) * U32_COUNT * n);
    if (e == cudaSuccess) e = cudaMalloc(&g.sync_ring, sizeof(uint32_t) * (size_t)g.d.ring_words * n);
    if (e == cudaSuccess) e = cudaMalloc(&g.amp_ring, sizeof(float) * (size_t)g.d.amp_phys * n);
    if (e != cudaSuccess) {
      free_batch(b);
      return fail(e == cudaErrorMemoryAllocation ? WAM_E_NOMEM : WAM_E_CUDA, std::string("state allocation: ") + cudaGetErrorString(e));
    }
    int rc = build_sync_templates(g, device);
    if (rc == WAM_OK) rc = atan_table_device(device, &g.d.atan_tab);
    if (rc == WAM_OK) rc = init_group_state(g, nullptr);
    if (rc != WAM_OK) { free_batch(b); return rc; }
  }
  CUDA_TRY(cudaDeviceSynchronize());
  for (int i = 0; i < 2; i++) {
    cudaError_t e = cudaStreamCreateWithFlags(&b->streams[i], cudaStreamNonBlocking);
    if (e != cudaSuccess) { free_batch(b); return fail(WAM_E_CUDA, std::string("cudaStreamCreate: ") + cudaGetErrorString(e)); }
  }
  *out = b;
  return WAM_OK;
}

// new FSKCore() + configure(cfg) on every stream again: fresh AGC, filters, rings, counters.
extern "C" int wam_fsk_batch_renew(wam_fsk_batch* b, void* cuda_stream) {
  if (!b) return fail(WAM_E_INVALID, "batch is NULL");
  CUDA_TRY(cudaSetDevice(b->device));
  for (auto& g : b->groups) {
    if (g.ids.empty()) continue;
    int rc = init_group_state(g, (cudaStream_t)cuda_stream);
    if (rc != WAM_OK) return rc;
  }
  b->demodulation_calls = 0;
  b->total_samples = 0;
  return WAM_OK;
}

extern "C" int wam_fsk_batch_destroy(wam_fsk_batch* b) {
  free_batch(b);
  return WAM_OK;
}

// FSKCore.reset() on every stream — fsk.ts:464-469: resetState(), clear the sync ring, drop queued
// bytes, zero the debug counters.  AGC, pre-filter, amplitude ring and silence threshold survive.
// Channel model for the synthetic workloads (BASELINE config 5: modulate -> AWGN -> demodulate): Gaussian noise of
// standard deviation d_sigma[row] added in place to device rows; (seed, seq, row, column) fixes every value.
extern "C" int wam_awgn_add_device(float* d_samples, long stride, long n_rows, long n, const float* d_sigma,
                                   unsigned long long seed, unsigned int seq, void* cuda_stream) {
  if (!d_samples || !d_sigma || n_rows < 0 || n < 0 || stride < n) return fail(WAM_E_INVALID, "bad buffer description");
  if ((n & 3) || (stride & 3) || (reinterpret_cast<uintptr_t>(d_samples) & 15))
    return fail(WAM_E_INVALID, "wam_awgn_add_device needs 16-byte aligned rows and a multiple of 4 samples");
  if (n_rows == 0 || n == 0) return WAM_OK;
  const long total = n_rows * (n / 4);
  const unsigned blocks = (unsigned)std::min<long>((total + 255) / 256, 148L * 32);
  awgn_add_kernel<<<blocks, 256, 0, (cudaStream_t)cuda_stream>>>(d_samples, stride, n_rows, n, d_sigma, seed, seq);
  CUDA_TRY(cudaGetLastError());
  return WAM_OK;
}

// Test hook: widens (scale > 1) the fast kernel's doubt band — error floor, amplitude-relative term and the silence
// compare's band — so that many decisions are flagged and the float64 checks carry real load.  Results must not change.
extern "C" int wam_fsk_batch_debug_fast_band(wam_fsk_batch* b, double scale) {
  if (!b || !(scale > 0)) return fail(WAM_E_INVALID, "bad argument");
  for (auto& g : b->groups) {
    g.d.f_kappa = 3e-7f * (float)scale;
    g.d.f_eps0 = 3e-7f * (float)scale;
    g.d.f_amp_eps = std::min(2e-6f * (float)scale, 0.5f);
    fold_band(g.d);
  }
  return WAM_OK;
}

extern "C" int wam_fsk_batch_reset(wam_fsk_batch* b) {
  if (!b) return fail(WAM_E_INVALID, "batch is NULL");
  CUDA_TRY(cudaSetDevice(b->device));
  CUDA_TRY(cudaDeviceSynchronize());
  static const int f_zero[] = {F_LO_S, F_IX1, F_IX2, F_IY1, F_IY2, F_QX1, F_QX2, F_QY1, F_QY2, F_OX1, F_OX2,
                               F_OY1, F_OY2, F_LAST_PHASE, F_IACC, F_QACC, F_RING_WI, F_RING_RI, F_RING_LEN,
                               F_RAGGED_CALLS, F_RAGGED_TOTAL};
  static const int u_zero[] = {U_DSC, U_GSC, U_GMOD, U_BSC, U_NEXT_IDX, U_BIT_ACC, U_BIT_CNT, U_STARTED, U_BITPOS,
                               U_CURRENT, U_SIL_CNT, U_RING_POS, U_RING_LEN, U_SYNC_DET};
  for (auto& g : b->groups) {
    const size_t n = g.ids.size();
    if (n == 0) continue;
    for (int f : f_zero) CUDA_TRY(cudaMemset(g.f64 + (size_t)f * n, 0, sizeof(double) * n));
    for (int u : u_zero) CUDA_TRY(cudaMemset(g.u32 + (size_t)u * n, 0, sizeof(uint32_t) * n));
    fill_f64_kernel<<<(unsigned)((n + 255) / 256), 256>>>(g.f64 + (size_t)F_LO_C * n, 1.0, (long)n);
    CUDA_TRY(cudaGetLastError());
    g.doubt_state = 2;            // the doubt tracking restarts with the state machine
    g.fb.carry_pending = false;   // reset() empties the rings and counters the carried readings were about
  }
  b->demodulation_calls = 0;
  b->total_samples = 0;
  return WAM_OK;
}

extern "C" long wam_fsk_batch_out_capacity(wam_fsk_batch* b, long n_samples) {
  if (!b || n_samples < 0) return fail(WAM_E_INVALID, "bad argument");
  long cap = 0;
  for (auto& g : b->groups) {
    if (g.ids.empty()) continue;
    // processByte completes a byte every stop_pos + 1 decided bits whatever startBits / stopBits say (fsk.ts:346-375),
    // i.e. after at least (min(bpb, stop_pos + 1) - 1) * dspb + 1 decimated samples = 2x input samples
    const long bits = std::min<long>(g.d.bpb, g.d.stop_pos + 1);
    const long per = 2L * ((bits - 1) * g.d.dspb + 1);
    cap = std::max(cap, n_samples / per + 2);
  }
  return (cap + 15) / 16 * 16;
}

template <bool A, bool G>
static void launch_demod(const DemodLaunch& L, cudaStream_t st) {
  fsk_demod_exact_kernel<A, G><<<L.block_begin[L.n_groups], 32, 0, st>>>(L);
}

// cuTensorMapEncodeTiled through the runtime (libwam links only cudart)
typedef CUresult (*tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static tmap_encode_fn tmap_encoder() {
  static tmap_encode_fn fn = []() -> tmap_encode_fn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return (tmap_encode_fn)p;
  }();
  return fn;
}
// {n, rows} float32 view of a sample buffer, 32 x 32 boxes, 128-byte swizzle, zero fill outside
static bool make_sample_tmap(CUtensorMap* m, float* d_samples, long stride, long n, long rows) {
  tmap_encode_fn enc = tmap_encoder();
  if (!enc || n <= 0 || rows <= 0 || n > 0xffffffffL) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)n, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)stride * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)kTile, (cuuint32_t)kTile};
  const cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d_samples, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// launch the demodulator for all streams of `b` whose global id lies in [s0, s1); row 0 of the
// buffers is stream `row_base`.  All configuration groups go into one launch (up to
// kMaxGroupsPerLaunch per launch) so their one-warp CTAs share the SMs.
static int ensure(void** p, size_t* cur, size_t need);
static int fast_streams(wam_fsk_batch* b);

// The fused TMA kernel over a long call of many warps: the call is cut into time slabs of kSlabTiles tiles, one launch
// per slab, alternating between two streams so that slab j + 1 starts on the SM slots slab j frees; a CTA of slab
// j + 1 waits (slab_wait) until the CTA of slab j has published the state of the same 32 streams (slab_publish).
// Why: 2048 one-warp CTAs on 148 SMs x 4 schedulers leave schedulers with 4 and with 3 warps, and a warp on a 4-warp
// scheduler advances about 25 % slower (profiles/r01_notes.md); with one launch the kernel ends when those finish.
// With slabs the hardware CTA scheduler hands the next slab's CTAs to whichever slots come free first, so the streams
// migrate between slow and fast slots and all schedulers stay busy to the end.  At most two launches overlap (the
// third waits for the first on its stream), and the earlier one is fully resident by then: no CTA waits on a CTA that
// cannot run.
typedef void (*demod_kernel_fn)(const DemodLaunch);
// per_group: every configuration group of the call gets launches of its own (group 0 of a one-group DemodLaunch), on
// its own pair of streams — the fast kernel addresses its group's coefficients as compile-time constants.
static int launch_slabbed(wam_fsk_batch* b, const DemodLaunch& L0, const long* tmap_rows, long n, cudaStream_t st,
                          demod_kernel_fn kern, bool per_group = false) {
  const int W = L0.block_begin[L0.n_groups];
  const long slab_len = (long)kSlabTiles * kTile;
  const int n_str = per_group ? 2 * L0.n_groups : 2;
  if (n_str > 4) return fail(WAM_E_INVALID, "too many groups for per-group slab launches");
  int rc = fast_streams(b);
  if (rc != WAM_OK) return rc;
  rc = ensure((void**)&b->slab_done, &b->slab_done_bytes, sizeof(int) * (size_t)W);
  if (rc != WAM_OK) return rc;
  CUDA_TRY(cudaMemsetAsync(b->slab_done, 0, sizeof(int) * (size_t)W, st));
  CUDA_TRY(cudaEventRecord(b->slab_fork, st));
  for (int i = 0; i < n_str; i++) CUDA_TRY(cudaStreamWaitEvent(b->slab_streams[i], b->slab_fork, 0));
  int slab = 0;
  for (long t0 = 0; t0 < n; t0 += slab_len, ++slab) {
    // every slab is a call of its own on samples [t0, t0 + len): shifted sample pointers and TMA descriptors,
    // decoded bytes appended behind those of the earlier slabs
    DemodLaunch L = L0;
    const long len = std::min(slab_len, n - t0);
    for (int g = 0; g < L.n_groups; g++) {
      DemodArgs& a = L.g[g];
      a.samples = L0.g[g].samples + t0;
      a.n = len;
      if (slab > 0) a.append = 1;
      if (!make_sample_tmap(&L.tmap[g], a.samples, a.stride, len, tmap_rows[g]))
        return fail(WAM_E_CUDA, "cuTensorMapEncodeTiled failed for a time slab");
    }
    L.slab = slab;
    L.slab_done = b->slab_done;
    if (!per_group) {
      kern<<<W, 32, 0, b->slab_streams[slab % 2]>>>(L);
      b->launches++;
    } else {
      for (int g = 0; g < L.n_groups; g++) {
        DemodLaunch Lg;
        memset(&Lg, 0, sizeof(Lg));
        Lg.tmap[0] = L.tmap[g];
        Lg.n_groups = 1;
        Lg.block_begin[1] = L.block_begin[g + 1] - L.block_begin[g];
        Lg.slab = slab;
        Lg.slab_done = b->slab_done + L.block_begin[g];
        Lg.g[0] = L.g[g];
        kern<<<Lg.block_begin[1], 32, 0, b->slab_streams[2 * g + (slab % 2)]>>>(Lg);
        b->launches++;
      }
    }
  }
  CUDA_TRY(cudaGetLastError());
  for (int i = 0; i < n_str; i++) {
    CUDA_TRY(cudaEventRecord(b->slab_join[i], b->slab_streams[i]));
    CUDA_TRY(cudaStreamWaitEvent(st, b->slab_join[i], 0));
  }
  return WAM_OK;
}

#include "fast_host.inl"

static int launch_demod_range(wam_fsk_batch* b, long s0, long s1, long row_base, float* d_samples, long stride, long n,
                              uint8_t* d_out, long out_stride, int32_t* d_out_len, float* d_tap, uint32_t flags,
                              cudaStream_t st, bool append = false, const int32_t* d_n_valid = nullptr,
                              long n_valid_offset = 0, bool count_call = true) {
  const bool aligned = ((reinterpret_cast<uintptr_t>(d_samples) & 15) == 0) && (stride % 4 == 0);
  const bool wb = (flags & WAM_BATCH_WRITEBACK_AGC) != 0;
  const bool tap = (flags & WAM_BATCH_TAP_PREFILTER) != 0 && d_tap != nullptr;
  DemodLaunch L;
  memset(&L, 0, sizeof(L));
  const bool ragged = d_n_valid != nullptr;
  bool generic = wb || tap || (flags & WAM_BATCH_DEBUG_GENERIC_SM);
  bool tma_ok = aligned;  // every group of the launch has contiguous rows and a TMA descriptor
  long tmap_rows[kMaxGroupsPerLaunch] = {0, 0, 0, 0};
  Group* lg[kMaxGroupsPerLaunch] = {nullptr, nullptr, nullptr, nullptr};
  auto flush = [&]() -> int {
    if (L.n_groups == 0) return WAM_OK;
    // Fast path: float32 kernel with certified decisions, float64 re-run of the streams it flags.  Many streams and
    // long calls only (or WAM_BATCH_FORCE_FAST): short calls are latency-sized and stay with the float64 kernels.
    {
      bool fast = aligned && !generic && !ragged && tma_ok && n % kTile == 0 && L.n_groups <= kFastGroupsPerLaunch &&
                  !(flags & (WAM_BATCH_NO_TMA | WAM_BATCH_EXACT_ONLY));
      for (int g = 0; g < L.n_groups && fast; g++)
        fast = L.g[g].d.fast_ok != 0 && L.g[g].d.tmpl0_words > 0 && !lg[g]->unaligned;
      if (fast && !(flags & WAM_BATCH_FORCE_FAST))
        fast = L.block_begin[L.n_groups] >= 4 * b->sm_count && n >= 4 * (long)kSlabTiles * kTile;
      if (fast) {
        int rc = fast_demodulate(b, L, lg, tmap_rows, n, flags, st);
        memset(&L, 0, sizeof(L));
        generic = wb || tap || (flags & WAM_BATCH_DEBUG_GENERIC_SM);
        tma_ok = aligned;
        return rc;
      }
      for (int g = 0; g < L.n_groups; g++) {
        // a float64 kernel is about to run on these groups: first settle what the last fast call left undecided
        if (lg[g]->fb.carry_pending) { int rc = fast_carry_settle(b, *lg[g], st); if (rc != WAM_OK) return rc; }
        lg[g]->doubt_state = 2;
      }
    }
    // Few streams (<= 5 three-warp CTAs per SM): the warp-specialised pipeline (fsk_demod_pipe.cuh) advances a
    // stream at the longest of the three phase chains instead of their sum.  Many streams: the fused kernel,
    // whose one-warp CTAs already hide the chains across warps.
    bool pipe = !generic && !(flags & WAM_BATCH_NO_PIPELINE) && n >= 8 * kTile;
    if (pipe) {
      const int v = aligned ? 1 : 0;
      auto kern = aligned ? fsk_demod_pipe_kernel<true> : fsk_demod_pipe_kernel<false>;
      if (b->pipe_per_sm[v] < 0) {  // once per batch: shared-memory footprint and resident CTAs per SM
        int ring_words = 0;
        for (auto& g : b->groups) ring_words = std::max(ring_words, g.d.ring_words);
        b->pipe_smem = sizeof(PipeShared) + (size_t)ring_words * 32 * sizeof(uint32_t);
        b->pipe_ring_smem = 1;
        if (b->pipe_smem > 72 * 1024) { b->pipe_smem = sizeof(PipeShared); b->pipe_ring_smem = 0; }  // very low baud: rings stay in HBM
        int per_sm = 0;
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->pipe_smem));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kPipeThreads, b->pipe_smem));
        b->pipe_per_sm[v] = per_sm;
      }
      L.pipe_ring_smem = b->pipe_ring_smem;
      // one wave only: beyond it the fused kernel's one-warp CTAs hide the chains better
      pipe = b->pipe_per_sm[v] > 0 && L.block_begin[L.n_groups] <= b->pipe_per_sm[v] * b->sm_count;
      if (pipe) kern<<<L.block_begin[L.n_groups], kPipeThreads, b->pipe_smem, st>>>(L);
    }
    if (pipe) {
    } else if (aligned && !generic && !ragged && tma_ok && !(flags & WAM_BATCH_NO_TMA)) {
      const int W = L.block_begin[L.n_groups];
      const long n_tiles = (n + kTile - 1) / kTile;
      const bool force = (flags & WAM_BATCH_FORCE_SLABS) != 0;
      if (b->fused_per_sm < 0) {  // once per batch: resident CTAs per SM of the fused TMA kernel
        int per_sm = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fsk_demod_exact_kernel<true, false, true>, 32, 0));
        b->fused_per_sm = per_sm;
      }
      // slabs only when every CTA of a slab launch is resident at once (see launch_slabbed) and the launch is big
      // enough for the 4-warp / 3-warp scheduler imbalance to matter
      const bool one_wave = W <= b->fused_per_sm * b->sm_count;
      // warps per scheduler (4 per SM): slabs pay when a good part of the schedulers carries one warp more than the
      // rest (measured on B200: 2.0 and 3.0 per scheduler lose 2-7 % with slabs, 3.46 gains 11 %)
      const double per_sched = (double)W / (4.0 * b->sm_count);
      const double frac = per_sched - std::floor(per_sched);
      const bool uneven = frac > 0.03 && frac < 0.7;
      if (!(flags & WAM_BATCH_NO_SLABS) && one_wave &&
          ((W >= b->sm_count * 8 && n_tiles >= 4 * kSlabTiles && uneven) || (force && n_tiles > kSlabTiles))) {
        int rc = launch_slabbed(b, L, tmap_rows, n, st, fsk_demod_exact_kernel<true, false, true>);
        if (rc != WAM_OK) return rc;
        b->launches--;  // counted per slab inside
      } else {
        fsk_demod_exact_kernel<true, false, true><<<W, 32, 0, st>>>(L);
      }
    } else if (aligned) { if (generic || ragged) launch_demod<true, true>(L, st); else launch_demod<true, false>(L, st); }
    else         { if (generic || ragged) launch_demod<false, true>(L, st); else launch_demod<false, false>(L, st); }
    b->launches++;
    CUDA_TRY(cudaGetLastError());
    memset(&L, 0, sizeof(L));
    generic = wb || tap || (flags & WAM_BATCH_DEBUG_GENERIC_SM);
    tma_ok = aligned;
    return WAM_OK;
  };
  for (auto& g : b->groups) {
    if (g.ids.empty()) continue;
    const auto lo = std::lower_bound(g.ids.begin(), g.ids.end(), (int32_t)s0) - g.ids.begin();
    const auto hi = std::lower_bound(g.ids.begin(), g.ids.end(), (int32_t)s1) - g.ids.begin();
    if (hi <= lo) continue;
    DemodArgs& a = L.g[L.n_groups];
    lg[L.n_groups] = &g;
    a.d = g.d;
    a.ids = g.contiguous ? nullptr : g.d_ids;
    a.id0 = g.ids.front();
    a.n_local = (int)g.ids.size();
    a.l_begin = (int)lo; a.l_end = (int)hi;
    a.row_base = (int)row_base;
    a.f64 = g.f64; a.u32 = g.u32; a.sync_ring = g.sync_ring; a.amp_ring = g.amp_ring;
    a.samples = d_samples; a.stride = stride; a.n = n;
    a.out = d_out; a.out_stride = out_stride; a.out_len = d_out_len;
    a.tap = (tap || (flags & WAM_BATCH_TAP_FAST_DECISION)) ? d_tap : nullptr;
    if (n % kTile != 0 || ragged) g.unaligned = true;
    a.writeback = wb ? 1 : 0;
    if (g.d.ring_fractional || g.d.eod_count <= 16) generic = true;  // needs the per-sample state machine
    a.force_generic = (flags & WAM_BATCH_DEBUG_GENERIC_SM) ? 1 : 0;
    a.append = append ? 1 : 0;
    a.n_valid = d_n_valid; a.n_valid_offset = n_valid_offset; a.count_call = count_call ? 1 : 0;
    a.phase_cycles = b->phase_cycles;
    if (tma_ok && !generic && !ragged)
      tma_ok = g.contiguous && make_sample_tmap(&L.tmap[L.n_groups], d_samples, stride, n, s1 - row_base);
    tmap_rows[L.n_groups] = s1 - row_base;
    L.block_begin[L.n_groups + 1] = L.block_begin[L.n_groups] + (int)((hi - lo + 31) / 32);
    L.n_groups++;
    // a launch that may take the fast path carries at most kFastGroupsPerLaunch groups
    const bool fast_cand = aligned && !generic && !ragged && tma_ok && n % kTile == 0 &&
                           !(flags & (WAM_BATCH_NO_TMA | WAM_BATCH_EXACT_ONLY));
    if (L.n_groups == (fast_cand ? kFastGroupsPerLaunch : kMaxGroupsPerLaunch)) {
      int rc = flush();
      if (rc != WAM_OK) return rc;
    }
  }
  return flush();
}

static int demodulate_device_impl(wam_fsk_batch* b, float* d_samples, long stream_stride, long n_samples,
                                  const int32_t* d_n_valid, bool ragged, uint8_t* d_out, long out_stride,
                                  int32_t* d_out_len, float* d_tap, void* cuda_stream, uint32_t flags) {
  if (!b) return fail(WAM_E_INVALID, "batch is NULL");
  if (n_samples < 0 || stream_stride < n_samples || out_stride < 0 || !d_out_len || (!d_samples && n_samples > 0) ||
      (!d_out && out_stride > 0) || (ragged && !d_n_valid))
    return fail(WAM_E_INVALID, "bad buffer description");
  CUDA_TRY(cudaSetDevice(b->device));
  if (!ragged) {  // every stream receives the same call: counted once on the host (fsk.ts:195-196)
    b->demodulation_calls += 1;
    b->total_samples += (double)n_samples;
  }
  return launch_demod_range(b, 0, b->n_streams, 0, d_samples, stream_stride, n_samples, d_out, out_stride, d_out_len,
                            d_tap, flags, (cudaStream_t)cuda_stream, false, ragged ? d_n_valid : nullptr, 0, true);
}

extern "C" int wam_fsk_batch_demodulate_device(wam_fsk_batch* b, float* d_samples, long stream_stride, long n_samples,
                                               uint8_t* d_out, long out_stride, int32_t* d_out_len, float* d_tap,
                                               void* cuda_stream, uint32_t flags) {
  return demodulate_device_impl(b, d_samples, stream_stride, n_samples, nullptr, false, d_out, out_stride, d_out_len,
                                d_tap, cuda_stream, flags);
}

extern "C" int wam_fsk_batch_demodulate_ragged_device(wam_fsk_batch* b, float* d_samples, long stream_stride,
                                                      long n_samples, const int32_t* d_n_valid, uint8_t* d_out,
                                                      long out_stride, int32_t* d_out_len, void* cuda_stream,
                                                      uint32_t flags) {
  return demodulate_device_impl(b, d_samples, stream_stride, n_samples, d_n_valid, true, d_out, out_stride, d_out_len,
                                nullptr, cuda_stream, flags & ~(uint32_t)WAM_BATCH_TAP_PREFILTER);
}

static int ensure(void** p, size_t* cur, size_t need) {
  if (*cur >= need && *p) return WAM_OK;
  if (*p) cudaFree(*p);
  *p = nullptr;
  *cur = 0;
  cudaError_t e = cudaMalloc(p, need ? need : 16);
  if (e != cudaSuccess) return fail(e == cudaErrorMemoryAllocation ? WAM_E_NOMEM : WAM_E_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
  *cur = need;
  return WAM_OK;
}

// 16-bit PCM -> float32, sample / 32768 (exact: every int16 / 2^15 is a float32).  Rows of `len` samples; strides in
// elements; eight samples per thread where both rows allow 16-byte accesses.
__global__ void pcm16_to_f32_kernel(const int16_t* __restrict__ in, long in_stride, float* __restrict__ out, long out_stride,
                                    long n_rows, long len, int vec) {
  constexpr float k = 1.0f / 32768.0f;
  if (vec) {
    const long per_row = len / 8;
    const long total = n_rows * per_row;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
      const long r = i / per_row, c = (i - r * per_row) * 8;
      const int4 v = __ldcs(reinterpret_cast<const int4*>(in + r * in_stride + c));
      const int w[4] = {v.x, v.y, v.z, v.w};
      float f[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) { f[2 * j] = (float)(short)(w[j] & 0xffff) * k; f[2 * j + 1] = (float)(w[j] >> 16) * k; }
      float4* o = reinterpret_cast<float4*>(out + r * out_stride + c);
      o[0] = make_float4(f[0], f[1], f[2], f[3]);
      o[1] = make_float4(f[4], f[5], f[6], f[7]);
    }
    const long tail = len - per_row * 8;
    if (tail > 0)
      for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n_rows * tail; i += (long)gridDim.x * blockDim.x) {
        const long r = i / tail, c = per_row * 8 + (i - r * tail);
        out[r * out_stride + c] = (float)in[r * in_stride + c] * k;
      }
  } else {
    const long total = n_rows * len;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
      const long r = i / len, c = i - r * len;
      out[r * out_stride + c] = (float)in[r * in_stride + c] * k;
    }
  }
}

// `pcm` != nullptr: the samples arrive as 16-bit PCM (half the PCIe bytes) and are widened on the device.
static int demodulate_host_impl(wam_fsk_batch* b, float* samples, long stream_stride, long n_samples,
                                const int32_t* n_valid, bool ragged, uint8_t* out, long out_stride, int32_t* out_len,
                                uint32_t flags, const int16_t* pcm = nullptr) {
  if (!b) return fail(WAM_E_INVALID, "batch is NULL");
  if (pcm && (flags & WAM_BATCH_WRITEBACK_AGC)) return fail(WAM_E_INVALID, "WAM_BATCH_WRITEBACK_AGC needs float32 samples");
  if (n_samples < 0 || stream_stride < n_samples || out_stride < 0 || !out_len || (!samples && !pcm && n_samples > 0) ||
      (!out && out_stride > 0) || (ragged && !n_valid))
    return fail(WAM_E_INVALID, "bad buffer description");
  CUDA_TRY(cudaSetDevice(b->device));
  const int32_t* d_n_valid = nullptr;
  if (ragged) {
    for (long s = 0; s < b->n_streams; s++)
      if (n_valid[s] > n_samples) return fail(WAM_E_INVALID, "n_valid[s] must not exceed n_samples");
    int rc = ensure((void**)&b->stage_nvalid, &b->stage_nvalid_bytes, sizeof(int32_t) * (size_t)b->n_streams);
    if (rc != WAM_OK) return rc;
    CUDA_TRY(cudaMemcpyAsync(b->stage_nvalid, n_valid, sizeof(int32_t) * (size_t)b->n_streams, cudaMemcpyHostToDevice,
                             b->streams[1]));
    d_n_valid = b->stage_nvalid;
  } else {
    b->demodulation_calls += 1;
    b->total_samples += (double)n_samples;
  }
  flags &= (WAM_BATCH_WRITEBACK_AGC | WAM_BATCH_DEBUG_GENERIC_SM | WAM_BATCH_NO_PIPELINE | WAM_BATCH_NO_TMA | WAM_BATCH_NO_SLABS |
            WAM_BATCH_FORCE_SLABS);

  // The call is cut into TIME slabs (all streams, samples [t0, t1)): the H2D copy of slab k+1 overlaps
  // the kernel of slab k (copy stream + compute stream, two staging buffers), and every slab launch
  // keeps all of the batch's warps busy.  The per-stream state carries from slab to slab exactly as
  // it does between demodulateData() calls; decoded bytes are appended in device memory and read
  // back once.
  long slab = n_samples;
  {
    const long target = (long)(((size_t)768 << 20) / (sizeof(float) * (size_t)b->n_streams));
    if (target < n_samples) slab = std::max<long>(2048, target / 32 * 32);
    slab = std::min(slab, n_samples);
  }
  const long dstride = (std::max<long>(slab, 1) + 3) / 4 * 4;
  const size_t need_s = sizeof(float) * (size_t)dstride * (size_t)b->n_streams;
  const size_t need_o = (size_t)std::max<long>(out_stride, 1) * (size_t)b->n_streams;
  const size_t need_l = sizeof(int32_t) * (size_t)b->n_streams;
  {
    size_t cs = b->stage_samples_bytes, co = b->stage_out_bytes, cl = b->stage_len_bytes;
    for (int i = 0; i < 2; i++) {
      cs = b->stage_samples_bytes;
      int rc = ensure((void**)&b->stage_samples[i], &cs, need_s);
      if (rc != WAM_OK) { b->stage_samples_bytes = 0; return rc; }
    }
    b->stage_samples_bytes = cs;
    if (pcm) {
      size_t cp = b->stage_pcm_bytes;
      for (int i = 0; i < 2; i++) {
        cp = b->stage_pcm_bytes;
        int rc = ensure((void**)&b->stage_pcm[i], &cp, need_s / 2);
        if (rc != WAM_OK) { b->stage_pcm_bytes = 0; return rc; }
      }
      b->stage_pcm_bytes = cp;
    }
    int rc = ensure((void**)&b->stage_out[0], &co, need_o);
    if (rc == WAM_OK) rc = ensure((void**)&b->stage_len[0], &cl, need_l);
    if (rc != WAM_OK) { b->stage_out_bytes = b->stage_len_bytes = 0; return rc; }
    b->stage_out_bytes = co; b->stage_len_bytes = cl;
  }
  cudaStream_t copy_st = b->streams[0], comp_st = b->streams[1];
  if (!b->ev_copied[0]) {
    for (int i = 0; i < 2; i++) {
      CUDA_TRY(cudaEventCreateWithFlags(&b->ev_copied[i], cudaEventDisableTiming));
      CUDA_TRY(cudaEventCreateWithFlags(&b->ev_consumed[i], cudaEventDisableTiming));
    }
  }
  CUDA_TRY(cudaMemsetAsync(b->stage_len[0], 0, need_l, comp_st));
  const bool wb = (flags & WAM_BATCH_WRITEBACK_AGC) != 0;
  int k = 0;
  long nslab = 0;
  for (long t0 = 0; t0 < n_samples || (n_samples == 0 && nslab == 0); t0 += slab, k ^= 1, nslab++) {
    const long len = std::min(slab, n_samples - t0);
    if (nslab >= 2) CUDA_TRY(cudaStreamWaitEvent(copy_st, b->ev_consumed[k], 0));  // staging buffer k is free again
    if (len > 0 && pcm) {
      // half the bytes over PCIe; the widening pass runs on the copy stream right behind its slab's copy
      if (stream_stride == dstride && len == dstride)
        CUDA_TRY(cudaMemcpyAsync(b->stage_pcm[k], pcm, sizeof(int16_t) * (size_t)dstride * (size_t)b->n_streams,
                                 cudaMemcpyHostToDevice, copy_st));
      else
        CUDA_TRY(cudaMemcpy2DAsync(b->stage_pcm[k], sizeof(int16_t) * (size_t)dstride, pcm + t0,
                                   sizeof(int16_t) * (size_t)stream_stride, sizeof(int16_t) * (size_t)len,
                                   (size_t)b->n_streams, cudaMemcpyHostToDevice, copy_st));
      const int vec = (dstride % 8 == 0) ? 1 : 0;
      pcm16_to_f32_kernel<<<148 * 8, 256, 0, copy_st>>>(b->stage_pcm[k], dstride, b->stage_samples[k], dstride, b->n_streams, len, vec);
      CUDA_TRY(cudaGetLastError());
    } else if (len > 0) {
      if (stream_stride == dstride && len == dstride)
        CUDA_TRY(cudaMemcpyAsync(b->stage_samples[k], samples, sizeof(float) * (size_t)dstride * (size_t)b->n_streams,
                                 cudaMemcpyHostToDevice, copy_st));
      else
        CUDA_TRY(cudaMemcpy2DAsync(b->stage_samples[k], sizeof(float) * (size_t)dstride, samples + t0,
                                   sizeof(float) * (size_t)stream_stride, sizeof(float) * (size_t)len,
                                   (size_t)b->n_streams, cudaMemcpyHostToDevice, copy_st));
    }
    CUDA_TRY(cudaEventRecord(b->ev_copied[k], copy_st));
    CUDA_TRY(cudaStreamWaitEvent(comp_st, b->ev_copied[k], 0));
    int rc = launch_demod_range(b, 0, b->n_streams, 0, b->stage_samples[k], dstride, len, b->stage_out[0], out_stride,
                                b->stage_len[0], nullptr, flags, comp_st, /*append=*/true, d_n_valid, t0, nslab == 0);
    if (rc != WAM_OK) return rc;
    if (wb && len > 0)
      CUDA_TRY(cudaMemcpy2DAsync(samples + t0, sizeof(float) * (size_t)stream_stride, b->stage_samples[k],
                                 sizeof(float) * (size_t)dstride, sizeof(float) * (size_t)len, (size_t)b->n_streams,
                                 cudaMemcpyDeviceToHost, comp_st));
    CUDA_TRY(cudaEventRecord(b->ev_consumed[k], comp_st));
    if (n_samples == 0) break;
  }
  if (out_stride > 0)
    CUDA_TRY(cudaMemcpyAsync(out, b->stage_out[0], (size_t)out_stride * (size_t)b->n_streams, cudaMemcpyDeviceToHost, comp_st));
  CUDA_TRY(cudaMemcpyAsync(out_len, b->stage_len[0], need_l, cudaMemcpyDeviceToHost, comp_st));
  CUDA_TRY(cudaStreamSynchronize(copy_st));
  CUDA_TRY(cudaStreamSynchronize(comp_st));
  return WAM_OK;
}

extern "C" int wam_fsk_batch_demodulate(wam_fsk_batch* b, float* samples, long stream_stride, long n_samples,
                                        uint8_t* out, long out_stride, int32_t* out_len, uint32_t flags) {
  return demodulate_host_impl(b, samples, stream_stride, n_samples, nullptr, false, out, out_stride, out_len, flags);
}

extern "C" int wam_fsk_batch_demodulate_pcm16(wam_fsk_batch* b, const int16_t* samples, long stream_stride, long n_samples,
                                              uint8_t* out, long out_stride, int32_t* out_len, uint32_t flags) {
  if (!samples && n_samples > 0) return fail(WAM_E_INVALID, "samples is NULL");
  return demodulate_host_impl(b, nullptr, stream_stride, n_samples, nullptr, false, out, out_stride, out_len, flags, samples);
}

extern "C" int wam_fsk_batch_demodulate_ragged(wam_fsk_batch* b, float* samples, long stream_stride, long n_samples,
                                               const int32_t* n_valid, uint8_t* out, long out_stride, int32_t* out_len,
                                               uint32_t flags) {
  return demodulate_host_impl(b, samples, stream_stride, n_samples, n_valid, true, out, out_stride, out_len, flags);
}

extern "C" int wam_fsk_batch_status(wam_fsk_batch* b, wam_fsk_status* st) {
  if (!b || !st) return fail(WAM_E_INVALID, "bad argument");
  CUDA_TRY(cudaSetDevice(b->device));
  CUDA_TRY(cudaDeviceSynchronize());
  for (auto& g : b->groups) {
    const size_t n = g.ids.size();
    if (n == 0) continue;
    std::vector<double> f((size_t)F64_COUNT * n);
    std::vector<uint32_t> u((size_t)U32_COUNT * n);
    CUDA_TRY(cudaMemcpy(f.data(), g.f64, sizeof(double) * f.size(), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(u.data(), g.u32, sizeof(uint32_t) * u.size(), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < n; i++) {
      wam_fsk_status& s = st[g.ids[i]];
      s.ready = 1;
      s.frameStarted = (int32_t)u[(size_t)U_STARTED * n + i];
      s.globalSampleCounter = (double)u[(size_t)U_GSC * n + i];
      s.receivedBitsLength = g.d.ring_fractional ? f[(size_t)F_RING_LEN * n + i] : (double)u[(size_t)U_RING_LEN * n + i];
      s.byteBufferLength = 0;
      s.demodulationCalls = b->demodulation_calls + f[(size_t)F_RAGGED_CALLS * n + i];
      s.syncDetections = (double)u[(size_t)U_SYNC_DET * n + i];
      s.silenceThreshold = f[(size_t)F_SIL_THR * n + i];
      s.totalSamplesProcessed = b->total_samples + f[(size_t)F_RAGGED_TOTAL * n + i];
      s.eodEvents = (double)u[(size_t)U_EOD_EV * n + i];
      s.errorEvents = (double)u[(size_t)U_ERR * n + i];  // device-side error flags (WAM_ERR_*), 0 = none
      s.configuredEvents = b->configured_events;
    }
  }
  return WAM_OK;
}

extern "C" long wam_fsk_batch_launch_count(wam_fsk_batch* b) { return b ? b->launches : 0; }

extern "C" int wam_fsk_batch_fast_stats(wam_fsk_batch* b, wam_fast_stats* out) {
  if (!b || !out) return fail(WAM_E_INVALID, "bad argument");
  CUDA_TRY(cudaSetDevice(b->device));
  CUDA_TRY(cudaDeviceSynchronize());
  memset(out, 0, sizeof(*out));
  out->fast_calls = b->fast_calls;
  for (auto& g : b->groups) {
    const size_t n = g.ids.size();
    if (n == 0) continue;
    std::vector<uint32_t> ev(n), ds(n), er(n);
    CUDA_TRY(cudaMemcpy(ev.data(), g.u32 + (size_t)U_FLAG_EVER * n, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(ds.data(), g.u32 + (size_t)U_DOUBT_SAMPLES * n, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(er.data(), g.u32 + (size_t)U_ERR * n, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < n; i++) {
      if (ev[i]) out->flagged_streams++;
      out->flag_causes |= ev[i];
      out->doubtful_samples += ds[i];
      out->error_flags |= er[i];
    }
    if (g.fb.hard_count) {
      int32_t c[4] = {0, 0, 0, 0};
      CUDA_TRY(cudaMemcpy(c, g.fb.hard_count, sizeof(c), cudaMemcpyDeviceToHost));
      out->flagged_last_call += c[0];
      out->windows_confirmed += c[1];
      out->windows_refuted += c[2];
      out->windows_dropped += c[3];
    }
    if (g.fb.carry_count) {
      int32_t c[4] = {0, 0, 0, 0};
      CUDA_TRY(cudaMemcpy(c, g.fb.carry_count, sizeof(c), cudaMemcpyDeviceToHost));
      out->carried_settled += c[1];
      out->carried_corrected += c[2];
    }
  }
  return WAM_OK;
}

// Debug: the verification windows of the last fast call of configuration group `group`.  *n_classes receives the
// number of window classes C (classes 0..C-3: windows of c + 1 slabs; C-2 / C-1: the two stages of an end-of-data
// check); counts (nullable, C entries) the items per class; items (nullable, C * cap * 3 entries)
// items[(c * cap + i) * 3 + {0, 1, 2}] = stream (local index), slab, result (1 = confirmed, else 1 << 30 | what
// differed).  Returns cap (the per-class capacity), negative on error.
extern "C" int wam_fsk_batch_debug_fast_windows(wam_fsk_batch* b, int group, int* n_classes, int32_t* counts, int32_t* items,
                                                long items_cap) {
  if (!b || group < 0 || group >= (int)b->groups.size() || !n_classes) return fail(WAM_E_INVALID, "bad argument");
  *n_classes = kAllClasses;
  if (!counts) return kVerifyCap;
  CUDA_TRY(cudaSetDevice(b->device));
  CUDA_TRY(cudaDeviceSynchronize());
  FastBuffers& fb = b->groups[(size_t)group].fb;
  for (int c = 0; c < kAllClasses; c++) counts[c] = 0;
  if (!fb.scratch_ready) return kVerifyCap;
  CUDA_TRY(cudaMemcpy(counts, fb.item_count, sizeof(int32_t) * kAllClasses, cudaMemcpyDeviceToHost));
  counts[kStageE2] = counts[kStageE1];
  const size_t nv = (size_t)kAllClasses * kVerifyCap;
  if (items && items_cap >= (long)(nv * 3)) {
    std::vector<int32_t> li(nv), sl(nv), rs(nv);
    CUDA_TRY(cudaMemcpy(li.data(), fb.item_li, sizeof(int32_t) * nv, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(sl.data(), fb.item_slab, sizeof(int32_t) * nv, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(rs.data(), fb.item_res, sizeof(int32_t) * nv, cudaMemcpyDeviceToHost));
    for (size_t k = 0; k < nv; k++) { items[3 * k] = li[k]; items[3 * k + 1] = sl[k]; items[3 * k + 2] = rs[k]; }
  }
  return kVerifyCap;
}

// Debug: enable (enable != 0) / read-and-clear per-phase SM cycle counters of the demodulator kernel.
// out4 (nullable) receives the sums over CTAs of cycles spent in A1, A2, B and staging/other;
// per_cta (nullable, [n_ctas][4]) the individual counters (CTA order = launch order).
extern "C" int wam_fsk_batch_debug_phase_cycles(wam_fsk_batch* b, int enable, double* out4, double* per_cta,
                                                long n_ctas) {
  if (!b) return fail(WAM_E_INVALID, "batch is NULL");
  CUDA_TRY(cudaSetDevice(b->device));
  CUDA_TRY(cudaDeviceSynchronize());
  const long ctas = (b->n_streams + 31) / 32 + (long)b->groups.size();
  if (out4) {
    out4[0] = out4[1] = out4[2] = out4[3] = 0;
    if (b->phase_cycles) {
      std::vector<unsigned long long> h((size_t)b->phase_ctas * 4);
      CUDA_TRY(cudaMemcpy(h.data(), b->phase_cycles, sizeof(unsigned long long) * h.size(), cudaMemcpyDeviceToHost));
      for (size_t i = 0; i < h.size(); i++) out4[i & 3] += (double)h[i];
      if (per_cta)
        for (size_t i = 0; i < h.size() && i < (size_t)n_ctas * 4; i++) per_cta[i] = (double)h[i];
    }
  }
  if (enable && !b->phase_cycles) {
    CUDA_TRY(cudaMalloc(&b->phase_cycles, sizeof(unsigned long long) * (size_t)ctas * 4));
    b->phase_ctas = ctas;
  }
  if (!enable && b->phase_cycles) { cudaFree(b->phase_cycles); b->phase_cycles = nullptr; b->phase_ctas = 0; }
  if (b->phase_cycles) CUDA_TRY(cudaMemset(b->phase_cycles, 0, sizeof(unsigned long long) * (size_t)b->phase_ctas * 4));
  return WAM_OK;
}

// ------------------------------------------------------------------------------------------
// batched modulator
// ------------------------------------------------------------------------------------------
static long modulate_size(const FskDerived& d, long nbytes) {  // fsk.ts:391-394
  const long total_bytes = (long)d.n_preamble + d.n_sfd + nbytes;
  const long pad = total_bytes > 0 ? 2L * d.spb : 0;
  return total_bytes * d.bpb * d.spb + pad + (long)d.bpb * d.spb;
}

static int modulate_device_group(wam_fsk_batch* b, const Group& g, const uint8_t* d_data, long data_stride,
                                 const int32_t* d_data_len, long nbytes, long max_bytes, float* d_out, long out_stride,
                                 int32_t* d_out_len, long n_rows, cudaStream_t st) {
  const int tab_stride = std::max(1, (int)(g.d.n_preamble + g.d.n_sfd + max_bytes) * g.d.bpb);
  int rc = ensure((void**)&b->mod_prefix, &b->mod_prefix_bytes, sizeof(uint32_t) * (size_t)tab_stride * (size_t)n_rows);
  if (rc != WAM_OK) return rc;
  if (g.d.spb < 1) return fail(WAM_E_INVALID, "samplesPerBit must be >= 1");
  if (modulate_size(g.d, max_bytes) >= (1L << 31)) return fail(WAM_E_UNSUPPORTED, "frame longer than 2^31 samples");
  ModArgs a;
  a.d = g.d;
  a.data = d_data; a.data_stride = data_stride; a.data_len = d_data_len; a.nbytes = (int)nbytes;
  a.n_streams = (int)n_rows;
  a.out = d_out; a.out_stride = out_stride; a.out_len = d_out_len;
  a.bittab = b->mod_prefix; a.tab_stride = tab_stride;
  a.vec_ok = ((reinterpret_cast<uintptr_t>(d_out) & 15) == 0 && out_stride % 4 == 0) ? 1 : 0;
  // cycles per sample in 32-bit fixed point (the phase wraps modulo one cycle for free)
  a.step_fix[0] = (uint32_t)(unsigned long long)llround(fmod(g.d.space / g.d.fs, 1.0) * 4294967296.0);
  a.step_fix[1] = (uint32_t)(unsigned long long)llround(fmod(g.d.mark / g.d.fs, 1.0) * 4294967296.0);
  a.spb_magic = g.d.spb >= 2 ? (uint32_t)(4294967296ull / (unsigned long long)g.d.spb) : 0xffffffffu;
  const long max_total = std::min(modulate_size(g.d, max_bytes), out_stride);
  // blocks per row: whole rows per block when there are enough rows to fill the GPU, otherwise split rows
  const long per_pass = (long)kModThreads * 4;
  const long passes = (max_total + per_pass - 1) / per_pass;
  long bx = (8L * 148 + n_rows - 1) / n_rows;
  bx = std::max(1L, std::min(bx, passes));
  dim3 grid((unsigned)bx, (unsigned)n_rows);
  const size_t tab_bytes = sizeof(uint32_t) * (size_t)tab_stride;
  if (a.vec_ok && g.d.spb % 4 == 0 && tab_bytes <= 40 * 1024 && g.d.bpb < 64) {
    // one launch: every CTA builds its stream's phase table in shared memory
    fsk_modulate_fused_kernel<<<grid, kModThreads, tab_bytes, st>>>(a);
    CUDA_TRY(cudaGetLastError());
    b->launches += 1;
    return WAM_OK;
  }
  fsk_bit_phase_kernel<<<(unsigned)n_rows, kPhaseThreads, 0, st>>>(a);
  CUDA_TRY(cudaGetLastError());
  if (a.vec_ok && g.d.spb % 4 == 0) fsk_modulate_kernel<true><<<grid, kModThreads, 0, st>>>(a);
  else fsk_modulate_kernel<false><<<grid, kModThreads, 0, st>>>(a);
  CUDA_TRY(cudaGetLastError());
  b->launches += 2;
  return WAM_OK;
}

extern "C" int wam_fsk_batch_modulate_device(wam_fsk_batch* b, const uint8_t* d_data, long data_stride,
                                             const int32_t* d_data_len, long nbytes, float* d_out, long out_stride,
                                             int32_t* d_out_len, void* cuda_stream) {
  if (!b || !d_out || nbytes < 0 || (nbytes > 0 && !d_data)) return fail(WAM_E_INVALID, "bad argument");
  if (b->groups.size() != 1 || b->n_streams > 65535)
    return fail(WAM_E_UNSUPPORTED, "batched modulate supports one configuration and <= 65535 streams per call");
  CUDA_TRY(cudaSetDevice(b->device));
  return modulate_device_group(b, b->groups[0], d_data, data_stride, d_data_len, nbytes, nbytes, d_out, out_stride,
                               d_out_len, b->n_streams, (cudaStream_t)cuda_stream);
}

extern "C" int wam_fsk_batch_modulate(wam_fsk_batch* b, const uint8_t* data, long data_stride, const int32_t* data_len,
                                      long nbytes, float* out, long out_stride, int32_t* out_len) {
  if (!b || !out || nbytes < 0 || (nbytes > 0 && !data) || data_stride < nbytes) return fail(WAM_E_INVALID, "bad argument");
  if (b->groups.size() != 1) return fail(WAM_E_UNSUPPORTED, "batched modulate supports one configuration per batch");
  CUDA_TRY(cudaSetDevice(b->device));
  const Group& g = b->groups[0];
  const long n = b->n_streams;
  if (data_len)
    for (long s = 0; s < n; s++)
      if (data_len[s] < 0 || data_len[s] > nbytes) return fail(WAM_E_INVALID, "data_len[s] must be within 0..nbytes");
  cudaStream_t st = b->streams[0];
  const long dstride = std::max<long>(data_stride, 1);
  const long ostride = (out_stride + 3) / 4 * 4;
  const long chunk_max = 32768;
  for (long s0 = 0; s0 < n; s0 += chunk_max) {
    const long rows = std::min(chunk_max, n - s0);
    int rc = ensure((void**)&b->mod_data, &b->mod_data_bytes, (size_t)dstride * (size_t)rows);
    if (rc == WAM_OK) rc = ensure((void**)&b->mod_out, &b->mod_out_bytes, sizeof(float) * (size_t)ostride * (size_t)rows);
    if (rc == WAM_OK) rc = ensure((void**)&b->mod_len, &b->mod_len_bytes, sizeof(int32_t) * 2 * (size_t)rows);
    if (rc != WAM_OK) return rc;
    if (nbytes > 0)
      CUDA_TRY(cudaMemcpyAsync(b->mod_data, data + s0 * data_stride, (size_t)data_stride * (size_t)rows, cudaMemcpyHostToDevice, st));
    int32_t* d_len_in = nullptr;
    if (data_len) {
      d_len_in = b->mod_len + rows;
      CUDA_TRY(cudaMemcpyAsync(d_len_in, data_len + s0, sizeof(int32_t) * (size_t)rows, cudaMemcpyHostToDevice, st));
    }
    rc = modulate_device_group(b, g, b->mod_data, dstride, d_len_in, nbytes, nbytes, b->mod_out, ostride, b->mod_len, rows, st);
    if (rc != WAM_OK) return rc;
    CUDA_TRY(cudaMemcpy2DAsync(out + s0 * out_stride, sizeof(float) * (size_t)out_stride, b->mod_out,
                               sizeof(float) * (size_t)ostride, sizeof(float) * (size_t)out_stride, (size_t)rows,
                               cudaMemcpyDeviceToHost, st));
    if (out_len) CUDA_TRY(cudaMemcpyAsync(out_len + s0, b->mod_len, sizeof(int32_t) * (size_t)rows, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
  }
  return WAM_OK;
}

// ------------------------------------------------------------------------------------------
// single-stream modem (one FSKCore instance)
// ------------------------------------------------------------------------------------------
struct wam_fsk {
  int device = 0;
  wam_fsk_batch* b = nullptr;  // n_streams == 1
  bool ready = false;
  // instance fields that survive configure() in the reference
  bool has_agc = false;
  double agc_attack = 0, agc_release = 0;
  double demodulation_calls = 0, total_samples = 0, sync_detections_base = 0;
  double eod_events_base = 0, configured_events = 0;
  // pinned bounce buffers
  float* h_samples = nullptr; size_t h_samples_n = 0;
  uint8_t* h_out = nullptr; size_t h_out_n = 0;
  int32_t* h_len = nullptr;
};

extern "C" int wam_fsk_create(int device, wam_fsk** out) {
  if (!out) return fail(WAM_E_INVALID, "out is NULL");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(WAM_E_CUDA, "no CUDA device available (libwam has no CPU fallback)");
  if (device < 0 || device >= ndev) return fail(WAM_E_INVALID, "device index out of range");
  wam_fsk* m = new (std::nothrow) wam_fsk();
  if (!m) return fail(WAM_E_NOMEM, "host allocation failed");
  m->device = device;
  *out = m;
  return WAM_OK;
}

extern "C" int wam_fsk_destroy(wam_fsk* m) {
  if (!m) return WAM_OK;
  cudaSetDevice(m->device);
  free_batch(m->b);
  cudaFreeHost(m->h_samples); cudaFreeHost(m->h_out); cudaFreeHost(m->h_len);
  delete m;
  return WAM_OK;
}

static int read_f64(wam_fsk_batch* b, int field, double* v) {
  Group& g = b->groups[0];
  CUDA_TRY(cudaMemcpy(v, g.f64 + (size_t)field * g.ids.size(), sizeof(double), cudaMemcpyDeviceToHost));
  return WAM_OK;
}
static int write_f64(wam_fsk_batch* b, int field, double v) {
  Group& g = b->groups[0];
  CUDA_TRY(cudaMemcpy(g.f64 + (size_t)field * g.ids.size(), &v, sizeof(double), cudaMemcpyHostToDevice));
  return WAM_OK;
}
static int read_u32(wam_fsk_batch* b, int field, uint32_t* v) {
  Group& g = b->groups[0];
  CUDA_TRY(cudaMemcpy(v, g.u32 + (size_t)field * g.ids.size(), sizeof(uint32_t), cudaMemcpyDeviceToHost));
  return WAM_OK;
}

// configure() — fsk.ts:133-157.  A fresh AGC (when enabled), fresh filters and rings; what the
// reference keeps on the instance is carried over: silence threshold (fsk.ts:128 is only ever
// written at sync), debug counters, and a previously created AGC when agcEnabled is now false
// (fsk.ts:447-449 only assigns).
extern "C" int wam_fsk_configure(wam_fsk* m, const wam_fsk_config* c) {
  if (!m || !c) return fail(WAM_E_INVALID, "bad argument");
  CUDA_TRY(cudaSetDevice(m->device));
  double old_thr = 0.01, old_gain = 1.0;
  uint32_t old_sync = 0, old_eod = 0;
  if (m->b) {
    CUDA_TRY(cudaDeviceSynchronize());
    int rc = read_f64(m->b, F_SIL_THR, &old_thr);
    if (rc == WAM_OK) rc = read_f64(m->b, F_GAIN, &old_gain);
    if (rc == WAM_OK) rc = read_u32(m->b, U_SYNC_DET, &old_sync);
    if (rc == WAM_OK) rc = read_u32(m->b, U_EOD_EV, &old_eod);
    if (rc != WAM_OK) return rc;
  }
  wam_fsk_batch* nb = nullptr;
  int rc = wam_fsk_batch_create(m->device, 1, c, 1, nullptr, &nb);
  if (rc != WAM_OK) return rc;
  Group& g = nb->groups[0];
  if (c->agcEnabled) {
    m->has_agc = true;
    m->agc_attack = g.d.agc_attack;
    m->agc_release = g.d.agc_release;
  } else if (m->has_agc) {
    g.d.agc_enabled = 1;
    g.d.agc_attack = m->agc_attack;
    g.d.agc_release = m->agc_release;
    rc = write_f64(nb, F_GAIN, old_gain);
  }
  if (rc == WAM_OK) rc = write_f64(nb, F_SIL_THR, old_thr);
  if (rc != WAM_OK) { free_batch(nb); return rc; }
  if (m->b) {
    m->sync_detections_base += old_sync;
    m->eod_events_base += old_eod;
    free_batch(m->b);
  }
  m->b = nb;
  m->ready = true;
  m->configured_events += 1;
  return WAM_OK;
}

extern "C" int wam_fsk_is_ready(wam_fsk* m) { return (m && m->ready) ? 1 : 0; }

extern "C" long wam_fsk_modulate_size(wam_fsk* m, long nbytes) {
  if (!m || nbytes < 0) return fail(WAM_E_INVALID, "bad argument");
  if (!m->ready) return fail(WAM_E_NOT_CONFIGURED, "FSK modulator not configured");
  return modulate_size(m->b->groups[0].d, nbytes);
}

extern "C" int wam_fsk_modulate(wam_fsk* m, const uint8_t* data, long nbytes, float* out, long cap, long* n_out) {
  if (!m || nbytes < 0 || (nbytes > 0 && !data)) return fail(WAM_E_INVALID, "bad argument");
  if (!m->ready) return fail(WAM_E_NOT_CONFIGURED, "FSK modulator not configured");
  const long total = modulate_size(m->b->groups[0].d, nbytes);
  if (n_out) *n_out = total;
  if (!out || cap < total) return fail(WAM_E_CAPACITY, "output buffer too small for modulateData");
  int32_t len = 0;
  int rc = wam_fsk_batch_modulate(m->b, data, std::max<long>(nbytes, 1), nullptr, nbytes, out, total, &len);
  return rc;
}

extern "C" int wam_fsk_demodulate(wam_fsk* m, float* samples, long n, uint8_t* out, long cap, long* n_out) {
  if (!m || n < 0 || (n > 0 && !samples) || !n_out) return fail(WAM_E_INVALID, "bad argument");
  if (!m->ready) return fail(WAM_E_NOT_CONFIGURED, "FSK demodulator not configured");
  *n_out = 0;
  CUDA_TRY(cudaSetDevice(m->device));
  const long ocap = wam_fsk_batch_out_capacity(m->b, n);
  // refused before a sample is consumed: afterwards the bytes could not be handed over and would be lost
  if (cap < ocap || (!out && cap > 0)) return fail(WAM_E_CAPACITY, "output buffer smaller than wam_fsk_batch_out_capacity(n) bytes");
  m->demodulation_calls += 1;
  m->total_samples += (double)n;
  if (m->h_samples_n < (size_t)n || !m->h_samples) {
    cudaFreeHost(m->h_samples); m->h_samples = nullptr;
    const size_t want = std::max<size_t>((size_t)n, 4096);
    CUDA_TRY(cudaMallocHost((void**)&m->h_samples, sizeof(float) * want));
    m->h_samples_n = want;
  }
  if (m->h_out_n < (size_t)ocap || !m->h_out) {
    cudaFreeHost(m->h_out); m->h_out = nullptr;
    const size_t want = std::max<size_t>((size_t)ocap, 256);
    CUDA_TRY(cudaMallocHost((void**)&m->h_out, want));
    m->h_out_n = want;
  }
  if (!m->h_len) CUDA_TRY(cudaMallocHost((void**)&m->h_len, sizeof(int32_t) * 4));
  if (n > 0) memcpy(m->h_samples, samples, sizeof(float) * (size_t)n);
  const Group& g = m->b->groups[0];
  const uint32_t flags = g.d.agc_enabled ? WAM_BATCH_WRITEBACK_AGC : 0u;
  int rc = wam_fsk_batch_demodulate(m->b, m->h_samples, std::max<long>(n, 1), n, m->h_out, ocap, m->h_len, flags);
  if (rc != WAM_OK) return rc;
  if (flags && n > 0) memcpy(samples, m->h_samples, sizeof(float) * (size_t)n);  // fsk.ts:55 mutates the input
  const long nb = m->h_len[0];
  *n_out = nb;
  if (nb > cap) return fail(WAM_E_CAPACITY, "output buffer too small for demodulated bytes");
  if (nb > 0) memcpy(out, m->h_out, (size_t)nb);
  return WAM_OK;
}

extern "C" int wam_fsk_reset(wam_fsk* m) {  // fsk.ts:464-469
  if (!m) return fail(WAM_E_INVALID, "bad argument");
  m->demodulation_calls = 0;
  m->total_samples = 0;
  m->sync_detections_base = 0;
  if (!m->b) return WAM_OK;
  return wam_fsk_batch_reset(m->b);
}

extern "C" int wam_fsk_status_get(wam_fsk* m, wam_fsk_status* st) {
  if (!m || !st) return fail(WAM_E_INVALID, "bad argument");
  memset(st, 0, sizeof(*st));
  st->silenceThreshold = 0.01;
  st->configuredEvents = m->configured_events;
  if (!m->b) return WAM_OK;
  int rc = wam_fsk_batch_status(m->b, st);
  if (rc != WAM_OK) return rc;
  st->ready = m->ready ? 1 : 0;
  st->demodulationCalls = m->demodulation_calls;
  st->totalSamplesProcessed = m->total_samples;
  st->syncDetections += m->sync_detections_base;
  st->eodEvents += m->eod_events_base;
  st->configuredEvents = m->configured_events;
  return WAM_OK;
}

// ------------------------------------------------------------------------------------------
// CRC-16 / XModem
// ------------------------------------------------------------------------------------------
extern "C" uint16_t wam_crc16(const uint8_t* data, long n) {  // crc16.ts:21-38 (host scalar)
  uint32_t crc = 0xFFFF;
  for (long k = 0; k < n; k++) {
    crc ^= ((uint32_t)data[k] << 8);
    for (int i = 0; i < 8; i++) crc = (crc & 0x8000) ? ((crc << 1) ^ 0x1021) & 0xFFFF : (crc << 1) & 0xFFFF;
  }
  return (uint16_t)crc;
}

extern "C" long wam_xmodem_serialize(int sequence, const uint8_t* payload, long n, uint8_t* out, long cap) {
  if (sequence < 1 || sequence > 255) return fail(WAM_E_PKT_SEQUENCE, "Invalid sequence: " + std::to_string(sequence) + ". Must be 1-255.");
  if (n > 255) return fail(WAM_E_PKT_PAYLOAD, "Payload too large: " + std::to_string(n) + ". Max 255 bytes.");
  if (n < 0 || (n > 0 && !payload)) return fail(WAM_E_INVALID, "bad payload");
  const long total = 4 + n + 2;
  if (!out) return total;
  if (cap < total) return fail(WAM_E_CAPACITY, "output buffer too small");
  const uint16_t crc = wam_crc16(payload, n);
  out[0] = 0x01; out[1] = (uint8_t)sequence; out[2] = (uint8_t)((~sequence) & 0xFF); out[3] = (uint8_t)n;
  if (n > 0) memcpy(out + 4, payload, (size_t)n);
  out[4 + n] = (uint8_t)(crc >> 8); out[4 + n + 1] = (uint8_t)(crc & 0xFF);
  return total;
}

extern "C" int wam_xmodem_batch_check_device(const uint8_t* d_bytes, long stride, const int32_t* d_len,
                                             const int32_t* d_expected_seq, long n_streams, wam_pkt_result* d_results,
                                             void* cuda_stream) {
  if (n_streams < 0 || !d_len || !d_results || (!d_bytes && stride > 0)) return fail(WAM_E_INVALID, "bad argument");
  if (n_streams == 0) return WAM_OK;
  static_assert(sizeof(PktResultDev) == sizeof(wam_pkt_result), "layout");
  const long want_ctas = (n_streams + 3) / 4;
  xmodem_check_kernel<<<(unsigned)std::min<long>(want_ctas, 148L * 16), 128, 0, (cudaStream_t)cuda_stream>>>(
      d_bytes, stride, d_len, d_expected_seq, n_streams, reinterpret_cast<PktResultDev*>(d_results));
  CUDA_TRY(cudaGetLastError());
  return WAM_OK;
}

struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  int alloc(size_t bytes) {
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 16);
    if (e != cudaSuccess) return fail(e == cudaErrorMemoryAllocation ? WAM_E_NOMEM : WAM_E_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    return WAM_OK;
  }
};

static int select_device(int device) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(WAM_E_CUDA, "no CUDA device available (libwam has no CPU fallback)");
  if (device < 0 || device >= ndev) return fail(WAM_E_INVALID, "device index out of range");
  CUDA_TRY(cudaSetDevice(device));
  return WAM_OK;
}

extern "C" int wam_xmodem_batch_check(int device, const uint8_t* bytes, long stride, const int32_t* len,
                                      const int32_t* expected_seq, long n_streams, wam_pkt_result* results) {
  if (n_streams < 0 || !len || !results || stride < 0 || (!bytes && stride > 0)) return fail(WAM_E_INVALID, "bad argument");
  int rc = select_device(device);
  if (rc != WAM_OK) return rc;
  if (n_streams == 0) return WAM_OK;
  for (long s = 0; s < n_streams; s++)
    if (len[s] < 0 || len[s] > stride) return fail(WAM_E_INVALID, "len[s] must be within 0..stride");
  DevBuf db, dl, de, dr;
  const size_t nb = (size_t)stride * (size_t)n_streams;
  if ((rc = db.alloc(nb)) || (rc = dl.alloc(sizeof(int32_t) * (size_t)n_streams)) ||
      (rc = de.alloc(sizeof(int32_t) * (size_t)n_streams)) || (rc = dr.alloc(sizeof(wam_pkt_result) * (size_t)n_streams)))
    return rc;
  if (nb) CUDA_TRY(cudaMemcpy(db.p, bytes, nb, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(dl.p, len, sizeof(int32_t) * (size_t)n_streams, cudaMemcpyHostToDevice));
  if (expected_seq) CUDA_TRY(cudaMemcpy(de.p, expected_seq, sizeof(int32_t) * (size_t)n_streams, cudaMemcpyHostToDevice));
  rc = wam_xmodem_batch_check_device((const uint8_t*)db.p, stride, (const int32_t*)dl.p,
                                     expected_seq ? (const int32_t*)de.p : nullptr, n_streams, (wam_pkt_result*)dr.p, nullptr);
  if (rc != WAM_OK) return rc;
  CUDA_TRY(cudaMemcpy(results, dr.p, sizeof(wam_pkt_result) * (size_t)n_streams, cudaMemcpyDeviceToHost));
  return WAM_OK;
}

extern "C" int wam_xmodem_batch_receive_device(const uint8_t* d_bytes, long stride, const int32_t* d_len, long n_streams,
                                               int max_retries, wam_xmodem_rx_state* d_state, uint8_t* d_replies,
                                               int reply_cap, int32_t* d_n_replies, int32_t* d_consumed,
                                               uint8_t* d_data, long data_stride, void* cuda_stream) {
  if (n_streams < 0 || !d_len || !d_state || !d_n_replies || !d_consumed || (!d_bytes && stride > 0) ||
      reply_cap < 0 || (reply_cap > 0 && !d_replies) || data_stride < 0 || max_retries < 0)
    return fail(WAM_E_INVALID, "bad argument");
  if (n_streams == 0) return WAM_OK;
  static_assert(sizeof(XmodemRxStateDev) == sizeof(wam_xmodem_rx_state), "layout");
  const long want_ctas = (n_streams + 3) / 4;
  xmodem_receive_kernel<<<(unsigned)std::min<long>(want_ctas, 148L * 16), 128, 0, (cudaStream_t)cuda_stream>>>(
      d_bytes, stride, d_len, n_streams, max_retries, reinterpret_cast<XmodemRxStateDev*>(d_state), d_replies,
      reply_cap, d_n_replies, d_consumed, d_data, data_stride);
  CUDA_TRY(cudaGetLastError());
  return WAM_OK;
}

extern "C" int wam_xmodem_batch_receive(int device, const uint8_t* bytes, long stride, const int32_t* len, long n_streams,
                                        int max_retries, wam_xmodem_rx_state* state, uint8_t* replies, int reply_cap,
                                        int32_t* n_replies, int32_t* consumed, uint8_t* data, long data_stride) {
  if (n_streams < 0 || !len || !state || !n_replies || !consumed || stride < 0 || (!bytes && stride > 0) ||
      reply_cap < 0 || (reply_cap > 0 && !replies) || data_stride < 0 || max_retries < 0)
    return fail(WAM_E_INVALID, "bad argument");
  int rc = select_device(device);
  if (rc != WAM_OK) return rc;
  if (n_streams == 0) return WAM_OK;
  for (long s = 0; s < n_streams; s++)
    if (len[s] < 0 || len[s] > stride) return fail(WAM_E_INVALID, "len[s] must be within 0..stride");
  const size_t ns = (size_t)n_streams;
  const size_t nb = (size_t)stride * ns, nrep = (size_t)reply_cap * ns, nd = data ? (size_t)data_stride * ns : 0;
  DevBuf db, dl, ds, dr, dn, dc, dd;
  if ((rc = db.alloc(nb)) || (rc = dl.alloc(sizeof(int32_t) * ns)) || (rc = ds.alloc(sizeof(wam_xmodem_rx_state) * ns)) ||
      (rc = dr.alloc(nrep)) || (rc = dn.alloc(sizeof(int32_t) * ns)) || (rc = dc.alloc(sizeof(int32_t) * ns)) ||
      (rc = dd.alloc(nd)))
    return rc;
  if (nb) CUDA_TRY(cudaMemcpy(db.p, bytes, nb, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(dl.p, len, sizeof(int32_t) * ns, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(ds.p, state, sizeof(wam_xmodem_rx_state) * ns, cudaMemcpyHostToDevice));
  if (nd) CUDA_TRY(cudaMemcpy(dd.p, data, nd, cudaMemcpyHostToDevice));  // payloads are appended to what is there
  if (nrep) CUDA_TRY(cudaMemset(dr.p, 0, nrep));
  rc = wam_xmodem_batch_receive_device((const uint8_t*)db.p, stride, (const int32_t*)dl.p, n_streams, max_retries,
                                       (wam_xmodem_rx_state*)ds.p, (uint8_t*)dr.p, reply_cap, (int32_t*)dn.p,
                                       (int32_t*)dc.p, nd ? (uint8_t*)dd.p : nullptr, data_stride, nullptr);
  if (rc != WAM_OK) return rc;
  CUDA_TRY(cudaMemcpy(state, ds.p, sizeof(wam_xmodem_rx_state) * ns, cudaMemcpyDeviceToHost));
  if (nrep) CUDA_TRY(cudaMemcpy(replies, dr.p, nrep, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(n_replies, dn.p, sizeof(int32_t) * ns, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(consumed, dc.p, sizeof(int32_t) * ns, cudaMemcpyDeviceToHost));
  if (nd) CUDA_TRY(cudaMemcpy(data, dd.p, nd, cudaMemcpyDeviceToHost));
  return WAM_OK;
}

extern "C" int wam_crc16_batch(int device, const uint8_t* bytes, long stride, const int32_t* len, long n_blocks,
                               uint16_t* crc_out) {
  if (n_blocks < 0 || !len || !crc_out || stride < 0 || (!bytes && stride > 0)) return fail(WAM_E_INVALID, "bad argument");
  int rc = select_device(device);
  if (rc != WAM_OK) return rc;
  if (n_blocks == 0) return WAM_OK;
  for (long s = 0; s < n_blocks; s++)
    if (len[s] < 0 || len[s] > stride) return fail(WAM_E_INVALID, "len[s] must be within 0..stride");
  DevBuf db, dl, dc;
  const size_t nb = (size_t)stride * (size_t)n_blocks;
  if ((rc = db.alloc(nb)) || (rc = dl.alloc(sizeof(int32_t) * (size_t)n_blocks)) || (rc = dc.alloc(sizeof(uint16_t) * (size_t)n_blocks)))
    return rc;
  if (nb) CUDA_TRY(cudaMemcpy(db.p, bytes, nb, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(dl.p, len, sizeof(int32_t) * (size_t)n_blocks, cudaMemcpyHostToDevice));
  crc16_batch_kernel<<<(unsigned)((n_blocks + 3) / 4), 128>>>((const uint8_t*)db.p, stride, (const int32_t*)dl.p, n_blocks, (uint16_t*)dc.p);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpy(crc_out, dc.p, sizeof(uint16_t) * (size_t)n_blocks, cudaMemcpyDeviceToHost));
  return WAM_OK;
}

// ------------------------------------------------------------------------------------------
// filters.ts batched application
// ------------------------------------------------------------------------------------------
extern "C" long wam_iir_state_size(int nb, int na) {
  if (nb <= 0 || na <= 0) return 0;
  return (long)(nb - 1) + (long)(na - 1);
}

static int iir_fill_args(IirArgs& ia, const double* b, int nb, const double* a, int na) {
  if (!b || nb <= 0) return fail(WAM_E_FILTER_B_EMPTY, wam_error_string(WAM_E_FILTER_B_EMPTY));
  if (!a || na <= 0) return fail(WAM_E_FILTER_A_EMPTY, wam_error_string(WAM_E_FILTER_A_EMPTY));
  if (a[0] == 0) return fail(WAM_E_FILTER_A0_ZERO, wam_error_string(WAM_E_FILTER_A0_ZERO));
  if (nb > kMaxIirTaps || na > kMaxIirTaps) return fail(WAM_E_UNSUPPORTED, "IIR order above 7 not supported");
  memset(&ia, 0, sizeof(ia));
  // a0 normalisation — filters.ts:30-39
  ia.nb = nb; ia.na = na;
  for (int i = 0; i < nb; i++) ia.b[i] = (a[0] != 1) ? b[i] / a[0] : b[i];
  for (int i = 1; i < na; i++) ia.a[i] = (a[0] != 1) ? a[i] / a[0] : a[i];
  ia.a[0] = 1;
  return WAM_OK;
}

extern "C" int wam_iir_process_batch(int device, const double* b, int nb, const double* a, int na, const float* in,
                                     float* out, long stride, long n, long n_streams, double* state) {
  IirArgs ia;
  int rc = iir_fill_args(ia, b, nb, a, na);
  if (rc != WAM_OK) return rc;
  if (n < 0 || n_streams < 0 || stride < n || (n > 0 && n_streams > 0 && (!in || !out))) return fail(WAM_E_INVALID, "bad argument");
  rc = select_device(device);
  if (rc != WAM_OK) return rc;
  if (n == 0 || n_streams == 0) return WAM_OK;
  return iir_process_batch_host(ia, in, out, stride, n, n_streams, state);
}

extern "C" size_t wam_iir_scratch_bytes(long n, long n_streams) { return iir_scratch_bytes(n, n_streams); }

// DEVICE buffers on the current device, asynchronous on cuda_stream; d_state (nullable) is read and written in place.
extern "C" int wam_iir_process_batch_device(const double* b, int nb, const double* a, int na, const float* d_in, float* d_out,
                                            long stride, long n, long n_streams, double* d_state, void* d_scratch,
                                            size_t scratch_bytes, void* cuda_stream) {
  IirArgs ia;
  int rc = iir_fill_args(ia, b, nb, a, na);
  if (rc != WAM_OK) return rc;
  if (n < 0 || n_streams < 0 || stride < n || (n > 0 && n_streams > 0 && (!d_in || !d_out))) return fail(WAM_E_INVALID, "bad argument");
  return iir_process_batch_device(ia, d_in, d_out, stride, n, n_streams, d_state, d_scratch, scratch_bytes, (cudaStream_t)cuda_stream);
}

// DEVICE buffers (taps included); d_state (nullable) [n_streams][ntaps - 1] is updated in place through d_state_scratch.
extern "C" int wam_fir_process_batch_device(const double* d_taps, int ntaps, const float* d_in, float* d_out, long stride,
                                            long n, long n_streams, double* d_state, double* d_state_scratch, void* cuda_stream) {
  if (ntaps < 0 || (ntaps > 0 && !d_taps)) return fail(WAM_E_INVALID, "bad taps");
  if (ntaps > kMaxFirTaps) return fail(WAM_E_UNSUPPORTED, "FIR longer than 1024 taps not supported");
  if (n < 0 || n_streams < 0 || stride < n || (n > 0 && n_streams > 0 && (!d_in || !d_out))) return fail(WAM_E_INVALID, "bad argument");
  return fir_process_batch_device(d_taps, ntaps, d_in, d_out, stride, n, n_streams, d_state, d_state_scratch, (cudaStream_t)cuda_stream);
}

extern "C" int wam_fir_process_batch(int device, const double* taps, int ntaps, const float* in, float* out,
                                     long stride, long n, long n_streams, double* state) {
  if (ntaps < 0 || (ntaps > 0 && !taps)) return fail(WAM_E_INVALID, "bad taps");
  if (ntaps > kMaxFirTaps) return fail(WAM_E_UNSUPPORTED, "FIR longer than 1024 taps not supported");
  if (n < 0 || n_streams < 0 || stride < n || (n > 0 && n_streams > 0 && (!in || !out))) return fail(WAM_E_INVALID, "bad argument");
  int rc = select_device(device);
  if (rc != WAM_OK) return rc;
  if (n == 0 || n_streams == 0) return WAM_OK;
  return fir_process_batch_host(taps, ntaps, in, out, stride, n, n_streams, state);
}

// test hook for the device math primitives (tests/test_gpu_fastmath.py)
extern "C" int wam_debug_fastmath(int device, const double* y, const double* x, long n, double* out_atan2,
                                  double* out_sqrt, double* out_rcp) {
  if (n < 0 || (n > 0 && (!y || !x || !out_atan2 || !out_sqrt || !out_rcp))) return fail(WAM_E_INVALID, "bad argument");
  int rc = select_device(device);
  if (rc != WAM_OK) return rc;
  if (n == 0) return WAM_OK;
  const double2* tab = nullptr;
  if ((rc = atan_table_device(device, &tab)) != WAM_OK) return rc;
  DevBuf dy, dx, d1, d2, d3;
  const size_t nb = sizeof(double) * (size_t)n;
  if ((rc = dy.alloc(nb)) || (rc = dx.alloc(nb)) || (rc = d1.alloc(nb)) || (rc = d2.alloc(nb)) || (rc = d3.alloc(nb))) return rc;
  CUDA_TRY(cudaMemcpy(dy.p, y, nb, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(dx.p, x, nb, cudaMemcpyHostToDevice));
  fastmath_debug_kernel<<<(unsigned)((n + 255) / 256), 256>>>((const double*)dy.p, (const double*)dx.p, n, tab,
                                                             (double*)d1.p, (double*)d2.p, (double*)d3.p);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpy(out_atan2, d1.p, nb, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(out_sqrt, d2.p, nb, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(out_rcp, d3.p, nb, cudaMemcpyDeviceToHost));
  return WAM_OK;
}

// ------------------------------------------------------------------------------------------
// session multiplexer: many independent block-wise callers -> one ragged batch per flush
// ------------------------------------------------------------------------------------------
struct wam_fsk_mux {
  wam_fsk_batch* b = nullptr;
  long n_sessions = 0, max_block = 0, out_cap = 0;
  float* h_samples = nullptr;   // pinned [n_sessions][max_block]: blocks pushed since the last flush
  int32_t* h_pending = nullptr; // pinned [n_sessions]: samples pending per session (0 = nothing pushed)
  int32_t* h_nvalid = nullptr;  // pinned [n_sessions]: n_valid of the flush in flight
  uint8_t* h_out = nullptr;     // pinned [n_sessions][out_cap]
  int32_t* h_out_len = nullptr; // pinned [n_sessions]
  double flushes = 0, blocks = 0;
  // send half (ChunkedModulator per session, src/webaudio/chunked-modulator.ts): queued payloads, the modulated
  // signal of every session and how much of it has been handed out
  std::vector<std::vector<uint8_t>> tx_queued;
  std::vector<char> tx_has_queued;
  std::vector<std::vector<float>> tx_signal;
  std::vector<long> tx_pos;
};

extern "C" int wam_fsk_mux_destroy(wam_fsk_mux* m) {
  if (!m) return WAM_OK;
  if (m->b) cudaSetDevice(m->b->device);
  cudaFreeHost(m->h_samples); cudaFreeHost(m->h_pending); cudaFreeHost(m->h_nvalid);
  cudaFreeHost(m->h_out); cudaFreeHost(m->h_out_len);
  free_batch(m->b);
  delete m;
  return WAM_OK;
}

extern "C" int wam_fsk_mux_create(int device, long n_sessions, const wam_fsk_config* cfgs, int n_cfgs,
                                  const int32_t* cfg_index, long max_block, wam_fsk_mux** out) {
  if (!out) return fail(WAM_E_INVALID, "out is NULL");
  *out = nullptr;
  if (n_sessions <= 0 || max_block <= 0) return fail(WAM_E_INVALID, "n_sessions and max_block must be positive");
  wam_fsk_mux* m = new (std::nothrow) wam_fsk_mux();
  if (!m) return fail(WAM_E_NOMEM, "host allocation failed");
  int rc = wam_fsk_batch_create(device, n_sessions, cfgs, n_cfgs, cfg_index, &m->b);
  if (rc != WAM_OK) { delete m; return rc; }
  m->n_sessions = n_sessions;
  m->max_block = (max_block + 3) / 4 * 4;
  m->out_cap = wam_fsk_batch_out_capacity(m->b, m->max_block);
  const size_t ns = (size_t)n_sessions;
  cudaError_t e = cudaMallocHost(&m->h_samples, sizeof(float) * ns * (size_t)m->max_block);
  if (e == cudaSuccess) e = cudaMallocHost(&m->h_pending, sizeof(int32_t) * ns);
  if (e == cudaSuccess) e = cudaMallocHost(&m->h_nvalid, sizeof(int32_t) * ns);
  if (e == cudaSuccess) e = cudaMallocHost(&m->h_out, ns * (size_t)m->out_cap);
  if (e == cudaSuccess) e = cudaMallocHost(&m->h_out_len, sizeof(int32_t) * ns);
  if (e != cudaSuccess) {
    wam_fsk_mux_destroy(m);
    return fail(WAM_E_NOMEM, std::string("pinned allocation: ") + cudaGetErrorString(e));
  }
  memset(m->h_pending, 0, sizeof(int32_t) * ns);
  *out = m;
  return WAM_OK;
}

extern "C" int wam_fsk_mux_push(wam_fsk_mux* m, long session, const float* samples, long n) {
  if (!m || session < 0 || session >= m->n_sessions || n < 0 || (n > 0 && !samples)) return fail(WAM_E_INVALID, "bad argument");
  const long have = m->h_pending[session];
  if (have + n > m->max_block) return fail(WAM_E_CAPACITY, "session block buffer full: flush first (max_block samples per flush)");
  if (n > 0) memcpy(m->h_samples + (size_t)session * (size_t)m->max_block + have, samples, sizeof(float) * (size_t)n);
  m->h_pending[session] = (int32_t)(have + n);
  m->blocks += 1;
  return WAM_OK;
}

extern "C" long wam_fsk_mux_pending(wam_fsk_mux* m, long session) {
  if (!m || session < 0 || session >= m->n_sessions) return fail(WAM_E_INVALID, "bad argument");
  return m->h_pending[session];
}

extern "C" long wam_fsk_mux_out_capacity(wam_fsk_mux* m) { return m ? m->out_cap : fail(WAM_E_INVALID, "mux is NULL"); }
extern "C" wam_fsk_batch* wam_fsk_mux_batch(wam_fsk_mux* m) { return m ? m->b : nullptr; }

// One ragged batch over every session that pushed since the last flush (the others are not called).
// out [n_sessions][out_stride] (out_stride >= wam_fsk_mux_out_capacity), out_len[s] bytes completed for session s.
extern "C" int wam_fsk_mux_flush(wam_fsk_mux* m, uint8_t* out, long out_stride, int32_t* out_len) {
  if (!m || !out_len || out_stride < 0 || (!out && out_stride > 0)) return fail(WAM_E_INVALID, "bad argument");
  long n_max = 0;
  for (long s = 0; s < m->n_sessions; s++) {
    const int32_t p = m->h_pending[s];
    m->h_nvalid[s] = p > 0 ? p : -1;
    n_max = std::max<long>(n_max, p);
  }
  // refused before the state advances; the pushed samples stay queued until a flush succeeds
  if (n_max > 0 && out_stride < m->out_cap)
    return fail(WAM_E_CAPACITY, "out_stride smaller than wam_fsk_mux_out_capacity()");
  if (n_max == 0) {
    memset(out_len, 0, sizeof(int32_t) * (size_t)m->n_sessions);
    return WAM_OK;
  }
  n_max = (n_max + 3) / 4 * 4;
  int rc = demodulate_host_impl(m->b, m->h_samples, m->max_block, std::min(n_max, m->max_block), m->h_nvalid, true,
                                m->h_out, m->out_cap, m->h_out_len, 0);
  if (rc != WAM_OK) return rc;
  for (long s = 0; s < m->n_sessions; s++) {
    m->h_pending[s] = 0;
    const int32_t n = m->h_out_len[s];
    out_len[s] = n;
    if (n > 0) memcpy(out + (size_t)s * (size_t)out_stride, m->h_out + (size_t)s * (size_t)m->out_cap, (size_t)n);
  }
  m->flushes += 1;
  return WAM_OK;
}

// ---- send half of the session multiplexer: ChunkedModulator (src/webaudio/chunked-modulator.ts:31-87) per session,
// with the modulateData() calls of all sessions that queued a payload run as ONE batched modulate.
extern "C" int wam_fsk_mux_send(wam_fsk_mux* m, long session, const uint8_t* data, long n) {
  if (!m || session < 0 || session >= m->n_sessions || n < 0 || (n > 0 && !data)) return fail(WAM_E_INVALID, "bad argument");
  if (m->tx_queued.empty()) {
    m->tx_queued.resize((size_t)m->n_sessions); m->tx_has_queued.assign((size_t)m->n_sessions, 0);
    m->tx_signal.resize((size_t)m->n_sessions); m->tx_pos.assign((size_t)m->n_sessions, 0);
  }
  if (n == 0) {  // startModulation(empty) resets (chunked-modulator.ts:32-35)
    m->tx_has_queued[(size_t)session] = 0;
    m->tx_signal[(size_t)session].clear();
    m->tx_pos[(size_t)session] = 0;
    return WAM_OK;
  }
  m->tx_queued[(size_t)session].assign(data, data + n);
  m->tx_has_queued[(size_t)session] = 1;
  return WAM_OK;
}

// modulateData() for every session with a queued payload, one batched GPU call; their signals replace whatever those
// sessions were still sending (startModulation overwrites pendingSignal, chunked-modulator.ts:37-38).
extern "C" int wam_fsk_mux_modulate(wam_fsk_mux* m) {
  if (!m) return fail(WAM_E_INVALID, "mux is NULL");
  if (m->tx_queued.empty()) return WAM_OK;
  std::vector<long> who;
  long nbytes = 0;
  for (long s = 0; s < m->n_sessions; s++)
    if (m->tx_has_queued[(size_t)s]) { who.push_back(s); nbytes = std::max<long>(nbytes, (long)m->tx_queued[(size_t)s].size()); }
  if (who.empty()) return WAM_OK;
  if (m->b->groups.size() != 1) return fail(WAM_E_UNSUPPORTED, "the mux's send half supports one configuration per mux");
  // the batch modulates n_sessions rows: sessions without a payload get a zero-length row whose output is ignored
  const long n = m->n_sessions;
  std::vector<uint8_t> data((size_t)n * (size_t)nbytes, 0);
  std::vector<int32_t> len((size_t)n, 0), out_len((size_t)n, 0);
  for (long s : who) {
    const auto& q = m->tx_queued[(size_t)s];
    memcpy(data.data() + (size_t)s * (size_t)nbytes, q.data(), q.size());
    len[(size_t)s] = (int32_t)q.size();
  }
  const long total = modulate_size(m->b->groups[0].d, nbytes);
  std::vector<float> out((size_t)n * (size_t)total);
  int rc = wam_fsk_batch_modulate(m->b, data.data(), nbytes, len.data(), nbytes, out.data(), total, out_len.data());
  if (rc != WAM_OK) return rc;
  for (long s : who) {
    const float* row = out.data() + (size_t)s * (size_t)total;
    m->tx_signal[(size_t)s].assign(row, row + out_len[(size_t)s]);
    m->tx_pos[(size_t)s] = 0;
    m->tx_has_queued[(size_t)s] = 0;
  }
  return WAM_OK;
}

extern "C" int wam_fsk_mux_is_modulating(wam_fsk_mux* m, long session) {
  if (!m || session < 0 || session >= m->n_sessions) return fail(WAM_E_INVALID, "bad argument");
  return (!m->tx_signal.empty() && !m->tx_signal[(size_t)session].empty()) ? 1 : 0;
}

// getNextSamples(sampleCount) (chunked-modulator.ts:41-81).  Returns 1 and fills res / out when the session is
// sending, 0 when it is not (the reference returns null), negative on error.
extern "C" int wam_fsk_mux_pull(wam_fsk_mux* m, long session, float* out, long sample_count, wam_chunk_result* res) {
  if (!m || session < 0 || session >= m->n_sessions || sample_count < 0 || !res || (sample_count > 0 && !out))
    return fail(WAM_E_INVALID, "bad argument");
  memset(res, 0, sizeof(*res));
  if (m->tx_signal.empty()) return 0;
  auto& sig = m->tx_signal[(size_t)session];
  long& pos = m->tx_pos[(size_t)session];
  if (sig.empty()) return 0;
  const long remaining = (long)sig.size() - pos;
  if (remaining <= 0) return 0;
  const long k = std::min(sample_count, remaining);
  if (k > 0) memcpy(out, sig.data() + pos, sizeof(float) * (size_t)k);
  pos += k;
  res->samples = k;
  res->totalSamples = (long)sig.size();
  if (pos >= (long)sig.size()) {
    res->isComplete = 1;
    res->samplesConsumed = (long)sig.size();
    sig.clear();
    pos = 0;
  } else {
    res->isComplete = 0;
    res->samplesConsumed = pos;
  }
  return 1;
}

// Pin the calling thread to the CPUs next to `device` (its PCIe root's NUMA node, from sysfs), so that host staging
// buffers allocated and filled afterwards are node-local and the H2D copies do not cross the socket interconnect.
// Returns the number of CPUs in the new mask, 0 when the topology is not exposed (nothing changed).
extern "C" int wam_host_bind_near_device(int device) {
  char bus[32] = {0};
  CUDA_TRY(cudaDeviceGetPCIBusId(bus, (int)sizeof(bus), device));
  for (char* c = bus; *c; ++c) *c = (char)tolower((unsigned char)*c);
  const std::string path = std::string("/sys/bus/pci/devices/") + bus + "/local_cpulist";
  FILE* f = fopen(path.c_str(), "r");
  if (!f) return 0;
  char line[4096] = {0};
  const bool got = fgets(line, (int)sizeof(line), f) != nullptr;
  fclose(f);
  if (!got) return 0;
  cpu_set_t want, have;
  CPU_ZERO(&want);
  if (sched_getaffinity(0, sizeof(have), &have) != 0) return 0;
  int count = 0;
  for (char* p = line; *p && *p != '\n';) {  // "0-31,64-95"
    char* e = nullptr;
    const long lo = strtol(p, &e, 10);
    if (e == p) break;
    long hi = lo;
    p = e;
    if (*p == '-') { hi = strtol(p + 1, &e, 10); p = e; }
    for (long c = lo; c <= hi && c < CPU_SETSIZE; ++c)
      if (c >= 0 && CPU_ISSET((int)c, &have)) { CPU_SET((int)c, &want); ++count; }
    if (*p == ',') ++p;
  }
  if (count == 0) return 0;  // the process may not run there (cgroup mask): leave it alone
  if (sched_setaffinity(0, sizeof(want), &want) != 0) return 0;
  return count;
}

extern "C" int wam_host_alloc(void** p, size_t bytes) {
  if (!p) return fail(WAM_E_INVALID, "p is NULL");
  CUDA_TRY(cudaMallocHost(p, bytes ? bytes : 16));
  return WAM_OK;
}
extern "C" int wam_host_free(void* p) {
  if (p) CUDA_TRY(cudaFreeHost(p));
  return WAM_OK;
}

// definitions of the helpers declared in filters.cuh that need fail()/CUDA_TRY
#include "filters_host.inl"

#ifdef WAM_SEARCH_STATS
extern "C" int wam_debug_search_stats(unsigned long long* out52, int clear) {
  if (out52) cudaMemcpyFromSymbol(out52, wam::g_search_stats, sizeof(unsigned long long) * 52);
  if (clear) { unsigned long long z[52] = {0}; cudaMemcpyToSymbol(wam::g_search_stats, z, sizeof(z)); }
  return 0;
}
#endif
