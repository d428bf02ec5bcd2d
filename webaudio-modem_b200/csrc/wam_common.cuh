// wam_common.cuh — shared host/device definitions for libwam.so (B200 / sm_100a).
#pragma once

#include <cuda.h>  // CUtensorMap (type only: the encoder is fetched through cudaGetDriverEntryPoint)
#include <cuda_runtime.h>
#include <stdint.h>

namespace wam {

constexpr int kTmplSlots = 8;        // __constant__ copies of by-value sync templates (per device, shared by content)
constexpr int kTmpl0Words = 80;     // by-value sync template: up to 2560 compared samples ((nbits - 1) * dspb)
constexpr int kMaxPatternWords = 8;  // preamble+SFD template: up to 256 line bits
constexpr int kTile = 32;            // samples per stream per staged tile (one 128-byte row)
constexpr int kSlabTiles = 64;       // tiles per time slab of a long call (2048 samples per stream)
constexpr int kStages = 2;           // cp.async pipeline depth (per warp)

// Resident one-warp CTAs per SM the demodulator is compiled for.  BASELINE config 2 is 2048 warps
// over 148 SMs = 13.84 per SM: all of them must be resident at once (one wave), which needs
// <= 65536 / (14 * 32) = 146 registers per thread.
#ifndef WAM_DEMOD_MIN_BLOCKS
#define WAM_DEMOD_MIN_BLOCKS 14
#endif

// Everything FSKCore.configure() derives (src/modems/fsk.ts:133-157, :426-462), computed on the
// host in float64 with the reference's operation order, then passed to kernels by value so the
// coefficients sit in the constant bank.
struct FskDerived {
  // AGCProcessor (fsk.ts:44-50)
  int agc_enabled;
  double agc_attack, agc_release;
  // pre-filter: butterworthBandpass(center, max(preFilterBandwidth, carson), fs)  (fsk.ts:451-456)
  double pre_b0, pre_b1, pre_b2, pre_a1, pre_a2;
  // I/Q and post low-pass: butterworthLowpass(baud, fs)  (fsk.ts:457-461)
  double lp_b0, lp_b1, lp_b2, lp_a1, lp_a2;
  // local oscillator (fsk.ts:228)
  double omega, cos_omega, sin_omega;
  // framing (fsk.ts:426-444, :143-150)
  int spb, dspb, bpb, nbits, start_bits, stop_bits, parity;  // parity 0 none 1 even 2 odd
  int check_period;     // Math.round(dspb / 4)            (fsk.ts:299)
  int stop_pos;         // 9 or 10                         (fsk.ts:348)
  int eod_count;        // smallest integer >= bpb*dspb*0.7 (fsk.ts:148,288)
  int min_matched;      // smallest matched with matched/total > syncThreshold (fsk.ts:314-315)
  int total_bits;       // nbits * dspb
  uint32_t pattern[kMaxPatternWords];  // preambleSfdBits, bit k at [k>>5] bit (k&31)
  // sync ring (RingBuffer(Uint8Array, (nbits+32)*dspb*1.1), fsk.ts:149; utils.ts:14-18)
  int ring_fractional;  // capacity is not an integer: literal emulation path
  double ring_cap;      // maxLength as a JS number
  int ring_cap_int;     // trunc(ring_cap) = typed-array length
  int ring_words;       // integral path: power-of-two number of 32-bit words (>= total_bits+64 bits)
                        // fractional path: ceil(ring_cap_int/32)
  int amp_cap;          // dspb * 8 (fsk.ts:150): logical capacity of the amplitude ring
  int amp_phys;         // physical slots (amp_cap + 32): a tile's 16 amplitudes are stored ahead of use
  // word-aligned sync templates (integral ring): for every bit offset o = 0..31 of the window start,
  // tmpl_words words of expected bits and of compare masks (device memory, [32][tmpl_words])
  const uint32_t* tmpl_expect;
  const uint32_t* tmpl_mask;
  int tmpl_words;
  int max_mismatch;     // care_bits - min_matched (negative: can never sync)
  // the offset-0 template again, by value (constant bank): the search shifts the RING words into alignment
  // (funnel shift by the window's bit offset) and compares against these, so the template reads are the same for
  // every lane.  tmpl0_words == 0: template too long, use the per-offset tables above.
  int tmpl0_words;      // ceil(compared samples / 32)
  int tmpl0_full;       // floor(compared samples / 32): template words whose mask is all ones
  int tmpl_slot;        // >= 0: the same template also sits in __constant__ slot c_tmpl[tmpl_slot] (functions that are
                        // not inlined into the kernel cannot address the kernel parameters as constants)
  uint32_t tmpl0_expect[4 + kTmpl0Words + 4];  // word i at [4 + i]; the words around it are zero with a zero mask
  uint32_t tmpl0_mask[4 + kTmpl0Words + 4];
  const double2* atan_tab;  // device: {k / 64, atan(k / 64)}, k = 0..64
  // ---- fast path (fsk_demod_fast.cuh): float32 DSP behind a float64 AGC, decisions certified by a doubt band ----
  // The three biquads in NORMAL (coupled) form: w' = [[sg, -om], [om, sg]] w + (x, 0), y = k0 x + k1 w1 + k2 w2 (w before
  // the update).  Same transfer function as the reference's direct form I; in float32 its round-off is ~50x smaller
  // (measured on the CPU model of the kernel: 1e-8 rms on filteredPhaseDiff instead of 5e-7).
  double pre_nk1, pre_nk2, pre_nsg, pre_nom;  // float64 copies for the state conversion direct form <-> normal form
  double lp_nk1, lp_nk2, lp_nsg, lp_nom;
  float f_pre_k0, f_pre_k1, f_pre_k2, f_pre_sg, f_pre_om;
  float f_lp_k0, f_lp_k1, f_lp_k2, f_lp_sg, f_lp_om;
  // the same filters two input samples per step: state' = [[A, -B], [B, A]] state + (sg x0 + x1, om x0);
  // pre-filter second output y1 = k0 x1 + k1 x0 + c1 w1 + c2 w2; low-pass pair sum y0 + y1 = k0 x1 + kx0 x0 + kw1 w1 + kw2 w2
  // pre-filter, packed: (p0, p1) = ks1 s1 + ks0 s0 + kw1 w1 + kw2 w2, (w1', w2') = as1 s1 + as0 s0 + aw1 w1 + aw2 w2
  float2 f2_pre_ks0, f2_pre_ks1, f2_pre_kw1, f2_pre_kw2, f2_pre_as0, f2_pre_as1, f2_pre_aw1, f2_pre_aw2;
  float2 f2_lp_pw1, f2_lp_pw2;  // post filter state step: (w1', w2') = (sg, om) w1 + (-om, sg) w2 + (x, 0)
  float f_lp_A, f_lp_B, f_lp_kx0, f_lp_kw1, f_lp_kw2;
  float f_cw, f_sw;       // float32 LO: phasor of sample 1 after a reset
  float f_c2w, f_s2w;     // rotation by 2 omega (the even and the odd phasor both turn once per pair)
  float f_dphi_bias;      // atan2(f_s2w, f_c2w) - 2 omega: the float32 LO's constant offset on the phase difference
  float f_rho_e;          // envelope of the post filter's impulse response: |h(j)| <= f_gamma * f_rho_e^j
  float f_gamma;
  float f_kappa;          // relative float32 error of an I/Q output against the recent amplitude scale
  float f_eps0;           // floor of the doubt band on |filteredPhaseDiff|
  float f_bc_delta;       // a raw phase difference this close to +-pi may have wrapped the other way
  // the same constants folded for the kernel: gamma kappa / 2, eps0 (1 - rho), 4 / gamma, bc_delta - (4 / gamma) eps0 (1 - rho), 6.3 gamma
  float f_gk2, f_eps0r, f_4og, f_bc_thr, f_g63;
  float f_amp_eps;        // relative doubt band of the silence compare
  int fast_ok;            // the configuration qualifies for the fast kernel (complex poles, integral ring, by-value template)
  // frame-search prefilter: the window seen in sub-blocks of check_period samples (4 per line bit when dspb = 4 check_period):
  // expected majority bit and compare mask per sub-block, newest first; a sub-block whose majority differs from the
  // template holds at least sub_half mismatching samples, so sub_half * (differing sub-blocks) > max_mismatch rules a sync out
  uint32_t sub_expect[4], sub_mask[4];
  int sub_ok, sub_half, sub_blocks;
  // modulator (fsk.ts:389-424)
  double mark, space, fs;
  int n_preamble, n_sfd;
  uint8_t preamble_sfd[32];
};

// Per-stream streaming state, struct-of-arrays: f64[k * n + stream], u32[k * n + stream].
enum F64Field {
  F_GAIN = 0, F_PX1, F_PX2, F_PY1, F_PY2, F_LO_C, F_LO_S,
  F_IX1, F_IX2, F_IY1, F_IY2, F_QX1, F_QX2, F_QY1, F_QY2,
  F_OX1, F_OX2, F_OY1, F_OY2, F_LAST_PHASE, F_IACC, F_QACC, F_SIL_THR,
  F_RING_WI, F_RING_RI, F_RING_LEN,
  F_RAGGED_CALLS, F_RAGGED_TOTAL,  // demodulateData() calls / samples received through ragged launches (per stream)
  F_FAST_S, F_FAST_E, F_FAST_RSP,  // fast path: amplitude scale, error envelope, 1 / (2 amplitude) of the last phasor
  F64_COUNT
};
enum U32Field {
  U_DSC = 0, U_GSC, U_GMOD, U_BSC, U_NEXT_IDX, U_BIT_ACC, U_BIT_CNT, U_STARTED, U_BITPOS, U_CURRENT,
  U_SIL_CNT, U_RING_POS, U_RING_LEN, U_AMP_POS, U_AMP_LEN, U_SYNC_DET, U_EOD_EV, U_ERR,
  // fast path (doubt tracking): doubtful samples of the running vote (ones | zeros << 16); pending doubtful silence
  // compare (bit 31) + the silent run it would add; ring position behind the newest doubtful hard bit (0: none);
  // causes flagged in the current call (bit = WAM_FLAG_*), and over the batch's life (statistics)
  U_DVOTE, U_SILX, U_LAST_DOUBT, U_DCNT, U_FLAG, U_FLAG_EVER, U_DOUBT_SAMPLES,
  U_OUT_N,  // fast path: bytes this stream has produced so far in the current call (checkpointed per time slab)
  // fast path, frame-search prefilter: ones in the running sub-block (check_period samples), sub-blocks in the
  // sub-ring since the check phase was last broken (0xffffffff: wait for the next check instant), the sub-ring itself
  // (majority bit per sub-block, newest in bit 0 of word 0)
  U_SB_ONES, U_SB_VALID, U_SB0, U_SB1, U_SB2, U_SB3,
  U32_COUNT
};

struct DemodArgs {
  FskDerived d;
  // streams of this config group
  const int32_t* ids;   // global stream id per local index (nullptr: id = id0 + local)
  int id0;
  int n_local;          // streams in the group's state arrays (SoA stride)
  int l_begin, l_end;   // local index range processed by this launch
  int row_base;         // samples/out row of global stream id `row_base` is row 0
  // state
  double* f64;
  uint32_t* u32;
  uint32_t* sync_ring;  // [n_local][ring_words] (see ring_of)
  float* amp_ring;      // [n_local][amp_phys] (see amp_of)
  // data
  float* samples;       // [rows][stride]
  long stride;
  long n;               // samples per stream this call (ragged launches: the maximum)
  const int32_t* n_valid;  // nullable [rows]: stream's own sample count for the whole call, < 0 = stream not called
  long n_valid_offset;     // samples of the call consumed by earlier slabs: this launch sees n_valid[row] - offset
  int count_call;          // ragged launches: this launch is the first slab of a call (debug.demodulationCalls++)
  uint8_t* out;         // [rows][out_stride]
  long out_stride;
  int32_t* out_len;     // [rows]
  float* tap;           // optional [rows][stride]
  int force_generic;    // debug: per-sample state machine even where the event-driven one applies
  int writeback;        // write the AGC-scaled samples back into `samples` (fsk.ts:55)
  int append;           // out_len[row] holds the bytes already written for this stream: append after them
  unsigned long long* phase_cycles;  // debug (nullable): [CTA][4] SM cycles spent in A1, A2, B, staging/other
  // sub-selection (exact re-run of the streams the fast kernel flagged): local index j -> sel[j], and the number of
  // entries read from device memory (the host launches a grid for the worst case; surplus CTAs leave at once)
  double thin_margin;    // > 0: flag silence compares closer than this (relative) to the threshold (WAM_ERR_THIN_COMPARE)
  const int32_t* sel;
  const int32_t* sel_count;
  // fast path
  uint32_t* doubt_ring;  // (unused by the second fast kernel)
  int32_t* flag_list;    // local indices of the streams whose float64 re-run covers the whole call, appended at *flag_count
  int32_t* flag_count;
  // fast path, per launch (= one time slab): the state is read from f64 / u32 and written to f64_out / u32_out (the
  // slab's checkpoints); hard bits and amplitudes go to per-call LINEAR histories instead of the rings — stream li's
  // tile t of this launch is half word bit_hist[li * bh_stride + hist_t0 + t] and amplitudes
  // amp_hist[li * ah_stride + amp_t0 + 16 t ..]; both start with a prefix copied from the rings at the start of the
  // call.  Streams in which this launch flagged a decision are appended to slab_list as li | cause << 24.
  double* f64_out;
  uint32_t* u32_out;
  uint16_t* bit_hist;
  long bh_stride, hist_t0;
  float* amp_hist;
  long ah_stride, amp_t0;
  int32_t* slab_list;
  int32_t* slab_count;
};

// All configuration groups of a batch run in ONE launch (one-warp CTAs; blockIdx selects the group)
// so that their CTAs fill the SMs together.
constexpr int kMaxGroupsPerLaunch = 4;
struct DemodLaunch {
  // TMA descriptors of the groups' sample buffers ({n, rows} float32, 32 x 32 boxes, 128-byte swizzle), used by
  // the STAGE_TMA variant of fsk_demod_exact_kernel; first member so that each one is 64-byte aligned
  alignas(64) CUtensorMap tmap[kMaxGroupsPerLaunch];
  int n_groups;
  int block_begin[kMaxGroupsPerLaunch + 1];  // first CTA of every group, then the total
  int pipe_ring_smem;                        // fsk_demod_pipe_kernel: the sync rings are copied to shared memory
  // time slabs (fsk_demod_exact_kernel<.., STAGE_TMA>, host: launch_slabbed): this launch is slab number `slab` of a
  // call; slab_done[cta] = slabs of that warp-group published so far.  slab_done == nullptr: a call in one launch.
  int slab;
  int* slab_done;
  DemodArgs g[kMaxGroupsPerLaunch];
};

#define WAM_ERR_OUT_OVERFLOW 1u
#define WAM_ERR_PIPE_TIMEOUT 2u
#define WAM_ERR_SLAB_TIMEOUT 4u
#define WAM_ERR_CARRIED_DOUBT 16u  // a doubtful decision depended on float32 readings of the previous call (see fast_host.inl)
#define WAM_ERR_THIN_COMPARE 8u  // verification scratch only: an amplitude within thin_margin of the silence threshold

// causes of a flagged (doubtful) decision in the fast kernel
#define WAM_FLAG_VOTE_START 1u
#define WAM_FLAG_VOTE_DATA 2u
#define WAM_FLAG_VOTE_STOP 4u
#define WAM_FLAG_SYNC 8u
#define WAM_FLAG_EOD 16u
#define WAM_FLAG_RANGE 32u

}  // namespace wam
