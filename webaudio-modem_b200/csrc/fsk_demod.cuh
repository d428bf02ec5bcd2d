// fsk_demod.cuh — fused FSK demodulator, exact (float64) path.
//
// Restates FSKCore.demodulateData (src/modems/fsk.ts:190-222) for thousands of independent
// streams: one thread walks one stream in time and carries the whole reference state
// (AGC gain, four biquads, LO, decimator, both rings, bit-sync / byte / silence state);
// a warp owns 32 streams and is its own CTA.  Samples are [stream][time] float32 in
// HBM; each warp stages 32-stream x 32-sample tiles (one 128-byte line per stream) into shared
// memory with a double-buffered cp.async pipeline, XOR-swizzled so that the row-per-lane LDS.128
// reads are bank-conflict free.  Arithmetic is the reference's float64 pipeline with float32
// only at its Float32Array stores (fsk.ts:55, filters.ts:82-85, amplitude ring fsk.ts:150,282).
//
// Every tile is processed in three phases so that the floating-point work is branch-free:
//   A1  AGC (fsk.ts:52-76: non-linear gain recurrence, f64 gain, f32 store) and the band-pass
//       pre-filter (filters.ts:47-87, f32 output) over the 32 samples -> pf[i][lane] in smem;
//   A2  LO mix (fsk.ts:228-232), I/Q low-pass biquads, /2 boxcar decimation, atan2, wrapped phase
//       difference, post low-pass and slicer (fsk.ts:241-264) -> 16 hard bits (a register mask)
//       and 16 squared magnitudes in smem;
//   B   the decimated-rate state machine (fsk.ts:278-375): ring puts, silence/EOD, sync search,
//       majority-vote bit sampler, UART framing — event-driven per tile: ring puts, counters and
//       the vote accumulator advance in bulk with bit operations, and only the samples where
//       something can happen (EOD crossing, a due sync check, a bit decision) are stepped.
// resetState() (fsk.ts:175-188; on EOD or a bad start bit) zeroes the A2 state from inside B.  It is
// rare (about once per frame), so B reports the decimated index of the reset and A2 is REPLAYED
// for the rest of the tile from the zeroed state; A1 (AGC, pre-filter) is never reset and never
// replayed.  The result is the reference's exact causal order without a data-dependent branch in
// the per-sample arithmetic.
//
// Register budget: BASELINE config 2 needs 14 one-warp CTAs resident per SM (2048 warps / 148
// SMs), i.e. <= 128 registers per thread.  Only the A2 state lives in registers for the whole
// kernel; the A1 state and the state-machine state are parked in shared memory between phases.
#pragma once

#include "fastmath.cuh"
#include "wam_common.cuh"

// unroll factors of the two DSP loops (A/B-tuned on B200, profiles/r01_notes.md)
#ifndef WAM_A1_CHUNKS
#define WAM_A1_CHUNKS 2  // float4 chunks (4 samples each) per iteration of the AGC + pre-filter loop
#endif
#ifndef WAM_A2_UNROLL
#define WAM_A2_UNROLL 4  // pairs per iteration of the I/Q + discriminator loop
#endif

namespace wam {

constexpr int kA1Chunks = WAM_A1_CHUNKS;
constexpr int kA2Unroll = WAM_A2_UNROLL;

struct A1State {  // AGC + pre-filter (never reset)
  double gain, py1, py2;
  float px1, px2;  // pre-filter input history: float32 values (AGC output)
};
struct A2State {  // everything resetState() zeroes on the DSP side
  double lo_c, lo_s;
  double ix1, ix2, iy1, iy2, qx1, qx2, qy1, qy2;
  double ox1, ox2, oy1, oy2, last_phase, iacc, qacc;
  uint32_t dsc;
};
struct BState {  // decimated-rate state machine
  double sil_thr;
  uint32_t gsc, gmod, bsc, next_idx, bit_acc, bit_cnt, started, current, sil_cnt;
  int bitpos;
  uint32_t ring_pos, ring_len, amp_pos, amp_len, cur_word;
  int out_n;
};

// shared-memory parking slots (word-major, one column per lane)
constexpr int kParkD = 4;   // gain, py1, py2, sil_thr
constexpr int kParkU = 16;  // px1, px2 (float bits), 14 state-machine words

__device__ __forceinline__ void a1_load(A1State& s, const double (*pd)[32], const uint32_t (*pu)[32], int lane) {
  s.gain = pd[0][lane]; s.py1 = pd[1][lane]; s.py2 = pd[2][lane];
  s.px1 = __uint_as_float(pu[0][lane]); s.px2 = __uint_as_float(pu[1][lane]);
}
__device__ __forceinline__ void a1_store(const A1State& s, double (*pd)[32], uint32_t (*pu)[32], int lane) {
  pd[0][lane] = s.gain; pd[1][lane] = s.py1; pd[2][lane] = s.py2;
  pu[0][lane] = __float_as_uint(s.px1); pu[1][lane] = __float_as_uint(s.px2);
}
__device__ __forceinline__ void b_load(BState& b, const double (*pd)[32], const uint32_t (*pu)[32], int lane) {
  b.sil_thr = pd[3][lane];
  b.gsc = pu[2][lane]; b.gmod = pu[3][lane]; b.bsc = pu[4][lane]; b.next_idx = pu[5][lane];
  b.bit_acc = pu[6][lane]; b.bit_cnt = pu[7][lane];
  const uint32_t f = pu[8][lane];
  b.started = f & 1u; b.bitpos = (int)((f >> 8) & 0xffu) - 1; b.current = (f >> 16) & 0xffu;
  b.sil_cnt = pu[9][lane]; b.ring_pos = pu[10][lane]; b.ring_len = pu[11][lane];
  b.amp_pos = pu[12][lane]; b.amp_len = pu[13][lane]; b.cur_word = pu[14][lane]; b.out_n = (int)pu[15][lane];
}
__device__ __forceinline__ void b_store(const BState& b, double (*pd)[32], uint32_t (*pu)[32], int lane) {
  pd[3][lane] = b.sil_thr;
  pu[2][lane] = b.gsc; pu[3][lane] = b.gmod; pu[4][lane] = b.bsc; pu[5][lane] = b.next_idx;
  pu[6][lane] = b.bit_acc; pu[7][lane] = b.bit_cnt;
  pu[8][lane] = (b.started & 1u) | ((uint32_t)(b.bitpos + 1) << 8) | ((b.current & 0xffu) << 16);
  pu[9][lane] = b.sil_cnt; pu[10][lane] = b.ring_pos; pu[11][lane] = b.ring_len;
  pu[12][lane] = b.amp_pos; pu[13][lane] = b.amp_len; pu[14][lane] = b.cur_word; pu[15][lane] = (uint32_t)b.out_n;
}

// FSKCore.resetState — fsk.ts:175-188.  Not reset: AGC, pre-filter, rings, silence threshold.
// Split in two: the state machine (phase B) clears its own half where the reference calls resetState() and
// reports the reset; the DSP half is cleared by whoever owns the A2 state, outside the event loop (keeping the
// 17 doubles of A2State out of the loop-carried values of the state machine: 34 register moves per iteration).
__device__ __forceinline__ void reset_state_a2(A2State& s) {
  s.lo_c = 1.0; s.lo_s = 0.0; s.last_phase = 0.0;  // localOscPhase = 0
  s.ix1 = s.ix2 = s.iy1 = s.iy2 = 0.0;
  s.qx1 = s.qx2 = s.qy1 = s.qy2 = 0.0;
  s.ox1 = s.ox2 = s.oy1 = s.oy2 = 0.0;
  s.dsc = 0; s.iacc = 0.0; s.qacc = 0.0;
}
__device__ __forceinline__ void reset_state_b(BState& b) {
  b.gsc = 0; b.gmod = 0; b.bsc = 0; b.bit_acc = 0; b.bit_cnt = 0; b.next_idx = 0;
  b.current = 0; b.bitpos = 0;
  b.started = 0;
  b.sil_cnt = 0;
}

// ---- L2 residency hints.  The bit-packed sync rings (config 2: 33.5 MB) are re-read by every frame search and are
// the only global data with reuse; the sample stream (12.6 GB, read once) and the amplitude ring (168 MB, written
// once per slot and read only at a sync) would otherwise push them out of the 126 MB L2, so the ring words are
// loaded / stored evict_last and the streaming traffic evict_first.  (-DWAM_AB_NO_L2HINT: plain accesses.)
#ifndef WAM_AB_NO_L2HINT
__device__ __forceinline__ uint64_t l2_evict_last() {
  uint64_t p;
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_evict_first() {
  uint64_t p;
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
template <bool GLOBAL>
__device__ __forceinline__ uint32_t ring_ld(const uint32_t* p, uint64_t pol) {
  if (!GLOBAL) return *p;
  uint32_t v;
  // not volatile: only used inside the out-of-line search, where no store can intervene
  asm("ld.global.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return v;
}
template <bool GLOBAL>
__device__ __forceinline__ void ring_st(uint32_t* p, uint32_t v) {
  if (!GLOBAL) { *p = v; return; }
  asm volatile("st.global.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(l2_evict_last()) : "memory");
}
__device__ __forceinline__ void amp_st(float* p, float v) { __stcs(p, v); }
__device__ __forceinline__ void amp_st4(float* p, float4 v) { __stcs(reinterpret_cast<float4*>(p), v); }
#else
__device__ __forceinline__ uint64_t l2_evict_last() { return 0; }
template <bool GLOBAL>
__device__ __forceinline__ uint32_t ring_ld(const uint32_t* p, uint64_t) { return *p; }
template <bool GLOBAL>
__device__ __forceinline__ void ring_st(uint32_t* p, uint32_t v) { *p = v; }
__device__ __forceinline__ void amp_st(float* p, float v) { *p = v; }
__device__ __forceinline__ void amp_st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
#endif

// Sync rings in global memory: [stream][ring_words], a stream's words contiguous (the frame search reads four
// at a time); in shared memory (pipelined kernel) [word][lane].  Word w of a ring is ring[w * rstride].
__device__ __forceinline__ uint32_t* ring_of(const DemodArgs& a, int li) {
  return a.sync_ring + (size_t)li * (size_t)a.d.ring_words;
}
__device__ __forceinline__ uint4 ring_ld4(const uint32_t* p) {  // 16-byte aligned group of four ring words
#ifndef WAM_AB_NO_L2HINT
  uint4 v;
  asm("ld.global.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
      : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(l2_evict_last()) : "memory");
  return v;
#else
  return *reinterpret_cast<const uint4*>(p);
#endif
}

// Frame-sync template match, integral-capacity ring — fsk.ts:303-312.
// The reference compares the newest nbits*dspb ring samples with preambleSfdBits[nbits - j] for
// window j (j*dspb .. (j+1)*dspb-1 samples back); j == 0 compares against `undefined` and never
// matches.  Here the ring is bit-packed, and the expected bits / compare masks are precomputed on
// the host for each of the 32 possible bit offsets of the window start inside a ring word, so the
// match is one XOR + AND + POPC per 32 samples.  Returns the number of mismatches among the
// compared (j >= 1) samples, or a value > max_mismatch as soon as the threshold is out of reach.
__device__ __noinline__ int sync_mismatches(const uint32_t* __restrict__ ring, long ns, uint32_t pos,
                                            const FskDerived& d) {
  const uint32_t lo = pos - (uint32_t)d.total_bits;
  const uint32_t o = lo & 31u;
  const uint32_t wmask = (uint32_t)(d.ring_words - 1);
  uint32_t w = (lo >> 5) & wmask;
  const uint32_t* __restrict__ ex = d.tmpl_expect + (long)o * d.tmpl_words;
  const uint32_t* __restrict__ mk = d.tmpl_mask + (long)o * d.tmpl_words;
  int mism = 0;
  int i = 0;
  for (; i + 4 <= d.tmpl_words; i += 4) {
    const uint32_t r0 = ring[(long)((w + 0) & wmask) * ns], r1 = ring[(long)((w + 1) & wmask) * ns];
    const uint32_t r2 = ring[(long)((w + 2) & wmask) * ns], r3 = ring[(long)((w + 3) & wmask) * ns];
    mism += __popc((r0 ^ __ldg(ex + i)) & __ldg(mk + i)) + __popc((r1 ^ __ldg(ex + i + 1)) & __ldg(mk + i + 1)) +
            __popc((r2 ^ __ldg(ex + i + 2)) & __ldg(mk + i + 2)) + __popc((r3 ^ __ldg(ex + i + 3)) & __ldg(mk + i + 3));
    w += 4;
    if (mism > d.max_mismatch) return mism;  // cannot reach the threshold any more
  }
  for (; i < d.tmpl_words; ++i) {
    mism += __popc((ring[(long)(w & wmask) * ns] ^ __ldg(ex + i)) & __ldg(mk + i));
    ++w;
  }
  return mism;
}

// The same search with ONE template (offset 0, by value in the kernel parameters = constant bank: d.tmpl0_*):
// the window's ring words are shifted into alignment instead (funnel shift by the window's bit offset), so the
// template words are warp-uniform constant loads (no vector registers, no memory latency) and only the ring words
// come from memory, WAM_SYNC_ROUND of them in flight per round.  Word i of the aligned window holds window bits
// 32i..32i+31 = ring bits lo+32i..; the off-by-one template makes the newest dspb samples never match (masked out).
#ifndef WAM_SYNC_ROUND
#define WAM_SYNC_ROUND 4
#endif
template <bool GLOBAL>
__device__ __forceinline__ int sync_mismatches0(const uint32_t* __restrict__ ring, long ns, uint32_t pos,
                                                const FskDerived& d) {
  constexpr int R = WAM_SYNC_ROUND;
  const uint64_t keep = l2_evict_last();
  const uint32_t lo = pos - (uint32_t)d.total_bits;
  const uint32_t o = lo & 31u;
  const uint32_t wmask = (uint32_t)(d.ring_words - 1);
  const uint32_t st = (uint32_t)ns;  // ring_words * streams < 2^31 words: 32-bit element indices
  uint32_t w = (lo >> 5) & wmask;
  uint32_t prev = ring_ld<GLOBAL>(ring + w * st, keep);
  int mism = 0;
  int i = 0;
  for (; i + R <= d.tmpl0_words; i += R) {
    uint32_t r[R];
#pragma unroll
    for (int j = 0; j < R; ++j) r[j] = ring_ld<GLOBAL>(ring + ((w + 1u + (uint32_t)j) & wmask) * st, keep);
#pragma unroll
    for (int j = 0; j < R; ++j) {
      mism += __popc((__funnelshift_r(prev, r[j], o) ^ d.tmpl0_expect[4 + i + j]) & d.tmpl0_mask[4 + i + j]);
      prev = r[j];
    }
    w += R;
    if (mism > d.max_mismatch) return mism;  // cannot reach the threshold any more
  }
  for (; i < d.tmpl0_words; ++i) {
    const uint32_t r1 = ring_ld<GLOBAL>(ring + ((w + 1u) & wmask) * st, keep);
    mism += __popc((__funnelshift_r(prev, r1, o) ^ d.tmpl0_expect[4 + i]) & d.tmpl0_mask[4 + i]);
    prev = r1;
    ++w;
  }
  return mism;
}
// The same for a contiguous ring (global memory): aligned groups of four ring words per 16-byte load.  x_n = ring
// word (g + n) with g the aligned group holding the window's first word, s = first word's place in its group;
// window word i is the funnel shift of (x_{s+i}, x_{s+i+1}) and meets template word i, stored at [i + 4] behind four
// all-zero words so that the pairs in front of the window (i < 0) are masked out without a branch.
// __constant__ copies of the by-value templates, [slot][expect | mask][4 + kTmpl0Words + 4] (host: tmpl_slot_acquire)
__constant__ uint32_t c_tmpl[kTmplSlots][2][4 + kTmpl0Words + 4];

template <bool CONST_SLOT>
__device__ __forceinline__ int sync_mismatches0v(const uint32_t* __restrict__ ring, uint32_t pos, const FskDerived& d) {
  const uint32_t lo = pos - (uint32_t)d.total_bits;
  const uint32_t o = lo & 31u;
  const uint32_t wmask = (uint32_t)(d.ring_words - 1);
  const uint32_t w = (lo >> 5) & wmask;
  const uint32_t s = w & 3u;
  uint32_t g = w & ~3u;
  const int n_end = (int)s + d.tmpl0_words;  // pairs n = 0 .. n_end - 1 carry compared samples
  // pair n meets template word n - s; CONST_SLOT: read through the constant cache (LDC), otherwise generic loads
  const uint32_t* __restrict__ ex = (CONST_SLOT ? c_tmpl[d.tmpl_slot][0] : d.tmpl0_expect) + 4 - (int)s;
  const uint32_t* __restrict__ mk = (CONST_SLOT ? c_tmpl[d.tmpl_slot][1] : d.tmpl0_mask) + 4 - (int)s;
  // first group: pairs -1 .. 2, some of them in front of the window (zero mask)
  uint4 v = ring_ld4(ring + g);
  g = (g + 4u) & wmask;
  // the next group is always requested one round ahead (the ring is circular, so the last, unused request stays in
  // bounds): the early-exit test of a round no longer waits for that round's own memory round trip
  uint4 nx = ring_ld4(ring + g);
  g = (g + 4u) & wmask;
  int mism = __popc((__funnelshift_r(v.x, v.y, o) ^ ex[0]) & mk[0]) + __popc((__funnelshift_r(v.y, v.z, o) ^ ex[1]) & mk[1]) +
             __popc((__funnelshift_r(v.z, v.w, o) ^ ex[2]) & mk[2]);
  uint32_t prev = v.w;
  int n = 3;
  // interior groups: all four window words are compared in full (mask ~0): no mask words to fetch
  const int n_full = (int)s + d.tmpl0_full;  // pairs below n_full meet all-ones mask words
  for (; n + 4 <= n_full; n += 4) {
    v = nx;
    nx = ring_ld4(ring + g);
    g = (g + 4u) & wmask;
    mism += __popc(__funnelshift_r(prev, v.x, o) ^ ex[n]) + __popc(__funnelshift_r(v.x, v.y, o) ^ ex[n + 1]) +
            __popc(__funnelshift_r(v.y, v.z, o) ^ ex[n + 2]) + __popc(__funnelshift_r(v.z, v.w, o) ^ ex[n + 3]);
    prev = v.w;
    if (mism > d.max_mismatch) return mism;  // cannot reach the threshold any more
  }
  // last groups: the partial word at the end of the window and the words behind it (zero mask)
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
#ifdef WAM_SEARCH_STATS  // debug build (scripts/exp_search_stats.py)
__device__ unsigned long long g_search_stats[52];
#endif
// out-of-line copy for the fused kernel (no registers to spare for an inlined search)
__device__ __noinline__ int sync_mismatches0_call(const uint32_t* __restrict__ ring, uint32_t pos, const FskDerived& d) {
#ifdef WAM_SEARCH_STATS
  // [0] warp-level calls, [1] lane searches, [4 + lanes] calls by number of active lanes, [40 + pos / 2400] calls by time
  const unsigned am = __activemask();
  if ((int)(threadIdx.x & 31) == __ffs((int)am) - 1) {
    atomicAdd(&g_search_stats[0], 1ull);
    atomicAdd(&g_search_stats[4 + __popc(am)], 1ull);
    atomicAdd(&g_search_stats[40 + min(pos / 2400u, 11u)], 1ull);
  }
  atomicAdd(&g_search_stats[1], 1ull);
#endif
  if (d.tmpl_slot >= 0) return sync_mismatches0v<true>(ring, pos, d);
  return sync_mismatches0v<false>(ring, pos, d);
}

// ---- literal emulation of RingBuffer with a fractional capacity (utils.ts:14-47) ----
// writeIndex / readIndex / length live in the global f64 state (slow path, quirk configurations).
__device__ __forceinline__ double ring_fmod_cap(double x, double cap) {  // x in [0, 3*cap)
  const double cap2 = cap + cap;
  if (x >= cap2) return x - cap2;  // exact (Sterbenz)
  if (x >= cap) return x - cap;    // exact
  return x;
}
__device__ __forceinline__ bool ring_index_valid(double p, int buflen, int& ip) {
  ip = (int)p;
  return ((double)ip == p) && ip >= 0 && ip < buflen;
}
__device__ __noinline__ bool ring_put_fractional(double* __restrict__ f64, long ns, uint32_t* ring, int bit,
                                                 const FskDerived& d) {
  double wi = f64[F_RING_WI * ns], ri = f64[F_RING_RI * ns], flen = f64[F_RING_LEN * ns];
  int ip;
  if (ring_index_valid(wi, d.ring_cap_int, ip)) {
    uint32_t* w = ring + (ip >> 5);
    *w = (*w & ~(1u << (ip & 31))) | ((uint32_t)bit << (ip & 31));
  }
  wi = ring_fmod_cap(wi + 1.0, d.ring_cap);
  if (flen < d.ring_cap) flen += 1.0;
  else ri = ring_fmod_cap(ri + 1.0, d.ring_cap);
  f64[F_RING_WI * ns] = wi; f64[F_RING_RI * ns] = ri; f64[F_RING_LEN * ns] = flen;
  return flen >= (double)d.total_bits;
}
__device__ __noinline__ int sync_matched_fractional(const double* __restrict__ f64, long ns,
                                                    const uint32_t* __restrict__ ring, const FskDerived& d) {
  const double ring_ri = f64[F_RING_RI * ns], ring_flen = f64[F_RING_LEN * ns];
  int matched = 0;
  int remaining = d.nbits * d.dspb;
  for (int j = 0; j < d.nbits; ++j) {
    const int pb = d.nbits - j;
    const int expect = (j == 0) ? 0 : (d.pattern[pb >> 5] >> (pb & 31)) & 1;
    for (int k = 0; k < d.dspb; ++k) {
      const double idx = ring_flen - (double)(j * d.dspb + k) - 1.0;
      const double p = ring_fmod_cap(ring_ri + idx, d.ring_cap);
      int ip;
      const bool valid = ring_index_valid(p, d.ring_cap_int, ip);
      if (j == 0) {
        matched += valid ? 0 : 1;  // undefined === undefined (fsk.ts:306-307)
      } else if (valid) {
        const int bit = (ring[ip >> 5] >> (ip & 31)) & 1;
        matched += (bit == expect);
      }
    }
    remaining -= d.dspb;
    if (matched + remaining < d.min_matched) break;
  }
  return matched;
}

// FSKCore.processByte — fsk.ts:346-375.  Returns true when resetState() ran.
__device__ __forceinline__ bool process_byte(BState& b, int bit, const DemodArgs& a, int li,
                                             uint8_t* out_row) {
  const FskDerived& d = a.d;
  const int bp = b.bitpos;
  if (bp == 0) {
    if (bit != 0) { reset_state_b(b); return true; }
  } else if (bp >= 1 && bp <= 8) {
    b.current |= (uint32_t)bit << (8 - bp);
  } else if (d.parity != 0 && bp == 9) {
    // parity bit is skipped, never checked
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

// Amplitude rings in global memory: [stream][amp_phys] float32, a stream's slots contiguous (amp_phys is a multiple
// of 8, so rows are 16-byte aligned): a full tile's 16 amplitudes go out as four 16-byte stores, and the mean at a
// sync detection reads 16 bytes at a time.
__device__ __forceinline__ float* amp_of(const DemodArgs& a, int li) {
  return a.amp_ring + (size_t)li * (size_t)a.d.amp_phys;
}

// silence threshold = mean(amplitude ring) * 0.1, summed oldest -> newest in f64 — fsk.ts:321-326.
// amp_next = physical slot following the newest entry; the ring has amp_phys physical slots of
// which the newest amp_len (<= amp_cap) are the reference's ring contents.  `row` = this stream's slots.
__device__ __noinline__ double amp_ring_threshold(const float* __restrict__ row, uint32_t amp_next, uint32_t amp_len,
                                                  uint32_t amp_phys) {
  double sum = 0.0;
  uint32_t slot = (amp_next + amp_phys - amp_len) % amp_phys;
  uint32_t left = amp_len;
  while (left > 0u && (slot & 3u) != 0u) {
    sum += (double)row[slot];
    slot = (slot + 1u == amp_phys) ? 0u : slot + 1u;
    --left;
  }
  while (left >= 4u) {  // slot is a multiple of 4 and so is amp_phys: a group never straddles the end
    const float4 v = *reinterpret_cast<const float4*>(row + slot);
    sum += (double)v.x; sum += (double)v.y; sum += (double)v.z; sum += (double)v.w;  // same order as one by one
    slot = (slot + 4u == amp_phys) ? 0u : slot + 4u;
    left -= 4u;
  }
  while (left > 0u) {
    sum += (double)row[slot];
    slot = (slot + 1u == amp_phys) ? 0u : slot + 1u;
    --left;
  }
  return (sum / (double)amp_len) * 0.1;
}

// One decimated sample of FSKCore.processDownsampledBit (fsk.ts:278-344) AFTER the ring puts:
// silence/EOD, sync search or vote/bit decision.  ring_pos / amp_next describe the rings including
// this sample.  Returns true when resetState() ran.
template <bool GENERIC, bool RING_ARG>
__device__ __forceinline__ bool sm_step(BState& b, int bit, double amplitude, uint32_t ring_pos,
                                        bool ring_ready, uint32_t amp_next, uint32_t amp_len, const DemodArgs& a,
                                        int li, uint8_t* out_row, bool& thr_changed, uint32_t* ring_arg, long rstride_arg) {
  const FskDerived& d = a.d;
  const long ns = a.n_local;
  // RING_ARG: the caller keeps this stream's sync ring somewhere else (shared memory); otherwise the state array
  uint32_t* ring = RING_ARG ? ring_arg : ring_of(a, li);
  const long rstride = RING_ARG ? rstride_arg : 1;
  // silence / EOD — fsk.ts:285-295
  b.gsc++;
  b.gmod = (b.gmod + 1u == (uint32_t)d.check_period) ? 0u : b.gmod + 1u;
  if (amplitude < b.sil_thr) {
    b.sil_cnt++;
    if (b.sil_cnt >= (uint32_t)d.eod_count) {
      a.u32[(long)U_EOD_EV * ns + li]++;  // emit('eod')
      reset_state_b(b);
      return true;
    }
  } else {
    b.sil_cnt = 0;
  }
  if (!b.started) {
    // fsk.ts:297-328
    const bool due = d.check_period > 0 && b.gmod == 0u;
    if (due && ring_ready && d.total_bits > 0) {
      int matched;
      if (!GENERIC || !d.ring_fractional) {
        if ((b.ring_pos & 31u) != 0u)  // flush the register copy of the newest (partial) word
          ring_st<!RING_ARG>(ring + (long)((b.ring_pos >> 5) & (uint32_t)(d.ring_words - 1)) * rstride, b.cur_word);
        int mism;
        if (d.tmpl0_words > 0) {  // by-value template + funnel shift (templates up to kTmpl0Words words)
                    // RING_ARG (pipelined kernel): the ring may be in shared memory ([word][lane]) -> word-by-word walk
          mism = RING_ARG ? sync_mismatches0<false>(ring, rstride, ring_pos, d) : sync_mismatches0_call(ring, ring_pos, d);
        } else {
          mism = sync_mismatches(ring, rstride, ring_pos, d);
        }
        matched = (d.total_bits - d.dspb) - mism;
      } else {
        matched = sync_matched_fractional(a.f64 + li, ns, ring, d);
      }
      if (matched >= d.min_matched) {
        b.started = 1;
        b.current = 0; b.bitpos = 0;
        b.bit_acc = 0; b.bit_cnt = 0; b.bsc = 0; b.next_idx = 0;
        a.u32[(long)U_SYNC_DET * ns + li]++;
        b.sil_thr = amp_ring_threshold(amp_of(a, li), amp_next, amp_len, (uint32_t)d.amp_phys);
        thr_changed = true;
      }
    }
    return false;
  }
  // fsk.ts:330-341
  b.bit_acc += (uint32_t)bit;
  b.bit_cnt++;
  b.bsc++;
  if (b.bsc >= b.next_idx) {
    const int decided = (2u * b.bit_acc > b.bit_cnt) ? 1 : 0;  // acc > count/2
    b.bit_acc = 0; b.bit_cnt = 0;
    b.next_idx += (uint32_t)d.dspb;
    const bool rst = process_byte(b, decided, a, li, out_row);
    // gmod is only maintained while searching: resynchronise it when the frame ends
    if (!rst && !b.started && d.check_period > 0) b.gmod = b.gsc % (uint32_t)d.check_period;
    return rst;
  }
  return false;
}

// Generic per-sample state machine (any ring kind, any eod_count): puts + sm_step.
__device__ __forceinline__ bool sm_sample_generic(BState& b, int bit, double amplitude,
                                                  const DemodArgs& a, int li, uint8_t* out_row) {
  const FskDerived& d = a.d;
  const long ns = a.n_local;
  uint32_t* ring = ring_of(a, li);
  bool ready;
  // syncSamplesBuffer.put(bit) — fsk.ts:281
  if (!d.ring_fractional) {
    b.cur_word |= (uint32_t)bit << (b.ring_pos & 31u);
    b.ring_pos++;
    if ((b.ring_pos & 31u) == 0u) {
      ring[((b.ring_pos - 1u) >> 5) & (uint32_t)(d.ring_words - 1)] = b.cur_word;
      b.cur_word = 0u;
    }
    b.ring_len = min(b.ring_len + 1u, (uint32_t)d.ring_cap_int);
    ready = b.ring_len >= (uint32_t)d.total_bits;
  } else {
    ready = ring_put_fractional(a.f64 + li, ns, ring, bit, d);
  }
  // syncAmplitudeBuffer.put(amplitude) — fsk.ts:282 (Float32Array store)
  amp_of(a, li)[b.amp_pos] = (float)amplitude;
  b.amp_pos = (b.amp_pos + 1u == (uint32_t)d.amp_phys) ? 0u : b.amp_pos + 1u;
  b.amp_len = min(b.amp_len + 1u, (uint32_t)d.amp_cap);
  bool thr_changed = false;
  return sm_step<true, false>(b, bit, amplitude, b.ring_pos, ready, b.amp_pos, b.amp_len, a, li, out_row, thr_changed,
                              nullptr, 0);
}

// Event-driven state machine for one tile (integral ring, eod_count > 16).  `bits` holds the hard
// decisions of decimated samples 0..nk-1 of the tile, amp[k*32] their amplitudes (f64, smem).
// Returns the decimated index at which resetState() ran, or -1.
// THIN (verification launches only): flag silence compares closer than DemodArgs.thin_margin to the threshold.  A
// compile-time switch: as a run-time test the compiler hoists the sixteen compares of a tile above it, which costs the
// few-stream pipeline kernel a fifth of its speed.
template <bool RING_ARG, bool THIN = false>
__device__ __forceinline__ int sm_tile_events(BState& b, uint32_t bits, const double* __restrict__ amp,
                                              int b_from, int nk, uint32_t pos_t0, uint32_t len_t0, uint32_t slot_t0,
                                              uint32_t alen_t0, const DemodArgs& a, int li, uint8_t* out_row,
                                              uint32_t* ring_arg, long rstride_arg) {
  // this stream's bit-packed sync ring (word w at ring[w * rstride]): the global state array, or with RING_ARG
  // the caller's copy (shared memory in fsk_demod_pipe_kernel)
  const FskDerived& d = a.d;
  const long ns = a.n_local;
  uint32_t* ring = RING_ARG ? ring_arg : ring_of(a, li);
  const long rstride = RING_ARG ? rstride_arg : 1;
  float* aring = amp_of(a, li);
  const uint32_t wmask = (uint32_t)(d.ring_words - 1);

  uint32_t silent = 0u;
  // ---- bulk ring puts for samples [b_from, nk) — fsk.ts:281-282
  {
    const uint32_t p = pos_t0 + (uint32_t)b_from;
    if (b_from > 0) {
      // replay pass: the word holding position p may already have been flushed
      ring_st<!RING_ARG>(ring + (long)((b.ring_pos >> 5) & wmask) * rstride, b.cur_word);
      b.cur_word = ring[(long)((p >> 5) & wmask) * rstride];
    }
    const uint32_t cnt = (uint32_t)(nk - b_from);
    const uint32_t o = p & 31u;
    const uint32_t chunk = (bits >> b_from) & ((1u << cnt) - 1u);
    b.cur_word = (b.cur_word & ((1u << o) - 1u)) | (chunk << o);
    if (o + cnt >= 32u) {
      ring_st<!RING_ARG>(ring + (long)((p >> 5) & wmask) * rstride, b.cur_word);
      b.cur_word = (o + cnt > 32u) ? (chunk >> (32u - o)) : 0u;
    }
    b.ring_pos = pos_t0 + (uint32_t)nk;
    uint32_t slot = slot_t0 + (uint32_t)b_from;
    if (slot >= (uint32_t)d.amp_phys) slot -= (uint32_t)d.amp_phys;
    {
      // amplitude-ring puts (fsk.ts:282, Float32Array store) and the silence flags for the current threshold
      // (fsk.ts:286) from the same shared-memory reads; the ring wraps at most once inside a tile
      const double thr = b.sil_thr;
      if (b_from == 0 && nk == kTile / 2 && (slot & 3u) == 0u && slot + (uint32_t)(kTile / 2) <= (uint32_t)d.amp_phys) {
        // a whole tile into aligned slots: four 16-byte streaming stores
#pragma unroll
        for (int q = 0; q < kTile / 8; ++q) {
          const double a0 = amp[(4 * q) * 32], a1 = amp[(4 * q + 1) * 32], a2 = amp[(4 * q + 2) * 32], a3 = amp[(4 * q + 3) * 32];
          amp_st4(aring + slot + 4 * q, make_float4((float)a0, (float)a1, (float)a2, (float)a3));
          silent |= ((a0 < thr ? 1u : 0u) | (a1 < thr ? 2u : 0u) | (a2 < thr ? 4u : 0u) | (a3 < thr ? 8u : 0u)) << (4 * q);
        }
      } else {
        float* p = aring + slot;
        int until_wrap = d.amp_phys - (int)slot;
#pragma unroll 4
        for (int k = b_from; k < nk; ++k) {
          const double av = amp[k * 32];
          amp_st(p, (float)av);
          ++p;
          if (--until_wrap == 0) p = aring;
          silent |= (av < thr ? 1u : 0u) << k;
        }
      }
      if (THIN) {
        // verification runs whose threshold is exact to ~1e-12: a compare this close to it proves nothing
        const double m = thr * a.thin_margin;
        bool thin = false;
        for (int k = b_from; k < nk; ++k) thin = thin || fabs(amp[k * 32] - thr) <= m;
        if (thin) a.u32[(long)U_ERR * ns + li] |= WAM_ERR_THIN_COMPARE;
      }
    }
  }

  int k = b_from;
  while (k < nk) {
    // next sample at which an event can happen
    int k_evt = nk;
    {
      const uint32_t run = (uint32_t)__ffs((int)(~(silent >> k))) - 1u;  // leading silent run from k
      const uint32_t need = (uint32_t)d.eod_count - b.sil_cnt - 1u;      // silent samples before the EOD one
      if (need < run) k_evt = min(k_evt, k + (int)need);
    }
    if (!b.started) {
      if (d.check_period > 0) k_evt = min(k_evt, k + (int)((uint32_t)d.check_period - 1u - b.gmod));
    } else {
      const uint32_t nb = b.bsc + 1u;
      k_evt = min(k_evt, k + (int)(b.next_idx > nb ? b.next_idx - nb : 0u));
    }
    // ---- bulk advance over [k, k_evt)
    const int len = k_evt - k;
    if (len > 0) {
      const uint32_t m = ((1u << len) - 1u) << k;
      b.gsc += (uint32_t)len;
      if (!b.started) b.gmod += (uint32_t)len;  // stays below check_period by construction
      const uint32_t nz = ~silent & m;
      b.sil_cnt = nz ? (uint32_t)(k_evt - 1) - (31u - (uint32_t)__clz((int)nz)) : b.sil_cnt + (uint32_t)len;
      if (b.started) {
        b.bit_acc += (uint32_t)__popc(bits & m);
        b.bit_cnt += (uint32_t)len;
        b.bsc += (uint32_t)len;
      }
    }
    if (k_evt >= nk) break;
    // ---- the event sample itself
    const uint32_t pos_k = pos_t0 + (uint32_t)k_evt + 1u;
    const bool ready = len_t0 + (uint32_t)k_evt + 1u >= (uint32_t)d.total_bits;
    uint32_t slot_next = slot_t0 + (uint32_t)k_evt + 1u;
    if (slot_next >= (uint32_t)d.amp_phys) slot_next -= (uint32_t)d.amp_phys;
    const uint32_t alen = min(alen_t0 + (uint32_t)k_evt + 1u, (uint32_t)d.amp_cap);
    bool thr_changed = false;
    if (sm_step<false, RING_ARG>(b, (int)((bits >> k_evt) & 1u), amp[k_evt * 32], pos_k, ready, slot_next, alen, a, li,
                                 out_row, thr_changed, ring_arg, rstride_arg))
      return k_evt;
    if (thr_changed) {
      silent = 0u;
      for (int kk = k_evt + 1; kk < nk; ++kk) silent |= (amp[kk * 32] < b.sil_thr ? 1u : 0u) << kk;
    }
    k = k_evt + 1;
  }
  return -1;
}

// ---- phase A1: AGC + pre-filter for one sample -------------------------------------------------
// The AGC target 0.5 / level comes from 1 / level to ~1e-14 relative: MUFU.RCP seed + one Newton step in f64 (the
// reference divides in f64; the gain recurrence is a contraction, so an error this size is invisible next to the
// float32 store of fsk.ts:55 — see DESIGN.md "numerics").
__device__ __forceinline__ float phase_a1_sample(A1State& s, float x, const FskDerived& d, bool agc, double att,
                                                 double rel, float& agc_out) {
  // AGC, fsk.ts:52-76 (evaluated unconditionally, selected by `agc`, so the code stays branch-free)
  const float sg_agc = (float)((double)x * s.gain);
  const float sg = agc ? sg_agc : x;
  const float level = fabsf(sg);
  // 1 / level (MUFU.RCP seed + one Newton step); the 0.5 of "0.5 / level" rides on the update's first FMA instead of
  // an addition in front of the reciprocal
  const float lv = fmaxf(level, 5e-31f);
  float r0;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(lv));
  const double r = (double)r0;
  const double inv = fma(r, fma(-(double)lv, r, 1.0), r);
  // "gain unchanged" (AGC off, or level == 0: fsk.ts:60-66 has no branch for it) is rate 0: inv is finite
  // (lv >= 5e-31), so g == gain exactly and the clamps leave it alone — no select at the end of the recurrence chain
  const double rate = (agc && level > 0.0f) ? (level > 0.5f ? att : rel) : 0.0;
  double g = fma(fma(inv, 0.5, -s.gain), rate, s.gain);
  {
    // the two clamps exclude each other: both compares look at the unclamped g, so they run side by side and the
    // recurrence chain ends in compare -> select -> select instead of compare -> select -> compare -> select
    const bool over = g > 10.0, under = g < 0.1;
    g = over ? 10.0 : g;
    g = under ? 0.1 : g;
  }
  s.gain = g;
  agc_out = sg;
  // pre-filter: butterworthBandpass has b1 == 0 and b2 == -b0 exactly (filters.ts:230)
  double y = d.pre_b0 * ((double)sg - (double)s.px2);
  y = fma(-d.pre_a2, s.py2, y);
  y = fma(-d.pre_a1, s.py1, y);
  s.px2 = s.px1; s.px1 = sg; s.py2 = s.py1; s.py1 = y;
  return (float)y;  // Float32Array store of processBuffer (filters.ts:82-85)
}

// ---- phase A2: one input sample through LO mix and the I/Q low-pass filters --------------------
// butterworthLowpass has b1 == 2*b0 and b2 == b0 exactly (filters.ts:188).  The LO is a rotation
// recurrence started from (1, 0) at every reset (cos/sin of the pre-increment phase, fsk.ts:229-232),
// renormalised once per tile.
__device__ __forceinline__ void phase_a2_half(A2State& s, float pf, const FskDerived& d, double& yi, double& yq) {
  const double smp = (double)pf;
  const double xi = smp * s.lo_c;
  const double xq = smp * s.lo_s;
  const double nc = s.lo_c * d.cos_omega - s.lo_s * d.sin_omega;
  const double nsn = s.lo_s * d.cos_omega + s.lo_c * d.sin_omega;
  s.lo_c = nc; s.lo_s = nsn;
  double ui = xi + s.ix2;
  ui = fma(2.0, s.ix1, ui);
  yi = d.lp_b0 * ui;
  yi = fma(-d.lp_a2, s.iy2, yi);
  yi = fma(-d.lp_a1, s.iy1, yi);  // the newest output enters last: one DFMA on the recurrence chain
  s.ix2 = s.ix1; s.ix1 = xi; s.iy2 = s.iy1; s.iy1 = yi;
  double uq = xq + s.qx2;
  uq = fma(2.0, s.qx1, uq);
  yq = d.lp_b0 * uq;
  yq = fma(-d.lp_a2, s.qy2, yq);
  yq = fma(-d.lp_a1, s.qy1, yq);
  s.qx2 = s.qx1; s.qx1 = xq; s.qy2 = s.qy1; s.qy1 = yq;
}

// decimated-rate discriminator (fsk.ts:246-264) on the summed pair (2*avgI, 2*avgQ): returns the
// hard bit, and the squared magnitude whose root is twice the reference amplitude.
__device__ __forceinline__ int phase_a2_decim(A2State& s, double si, double sq, const FskDerived& d, double& p) {
  const double kPi = 3.141592653589793;
  const double phase = fast_atan2(sq, si, d.atan_tab);  // atan2(avgQ, avgI): scale invariant
  p = __dadd_rn(__dmul_rn(si, si), __dmul_rn(sq, sq));
  double pd = phase - s.last_phase;
  {
    // the same wrap (fsk.ts:255-256) with one compare: subtract 2*pi carrying pd's sign when |pd| > pi
    // (x - (-2pi) == x + 2pi exactly); 2*Math.PI = 0x401921FB54442D18
    const double adj = __hiloint2double((__double2hiint(pd) & (int)0x80000000) | 0x401921FB, 0x54442D18);
    const double wrapped = pd - adj;
    pd = fabs(pd) > kPi ? wrapped : pd;
  }
  s.last_phase = phase;
  double uo = pd + s.ox2;
  uo = fma(2.0, s.ox1, uo);
  double yo = d.lp_b0 * uo;
  yo = fma(-d.lp_a2, s.oy2, yo);
  yo = fma(-d.lp_a1, s.oy1, yo);
  s.ox2 = s.ox1; s.ox1 = pd; s.oy2 = s.oy1; s.oy1 = yo;
  return yo > 0.0 ? 1 : 0;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
  const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc, int src_bytes) {
  const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(dst), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// float index of (row, col) inside one swizzled 32x32 tile: 16-byte chunk index XOR (row & 7)
__device__ __forceinline__ int tile_index(int row, int col) {
  return row * kTile + ((((col >> 2) ^ (row & 7)) << 2) | (col & 3));
}

// Stage one 32-stream x 32-sample tile starting at sample t0 into `tile`.
template <bool ALIGNED>
__device__ __forceinline__ void stage_tile(float* tile, const DemodArgs& a, const int* rows, long t0, int lane) {
  if (ALIGNED) {
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int r = it * 4 + (lane >> 3);
      const int c = (lane & 7) * 4;
      const long row = rows[r];
      const long remain = a.n - (t0 + c);
      const int bytes = (row < 0 || remain <= 0) ? 0 : (remain >= 4 ? 16 : (int)remain * 4);
      const float* src = bytes ? a.samples + row * a.stride + t0 + c : a.samples;
      cp_async16(tile + tile_index(r, c), src, bytes);
    }
  } else {
#pragma unroll 4
    for (int r = 0; r < 32; ++r) {
      const long row = rows[r];
      const int bytes = (row >= 0 && t0 + lane < a.n) ? 4 : 0;
      const float* src = bytes ? a.samples + row * a.stride + t0 + lane : a.samples;
      cp_async4(tile + tile_index(r, lane), src, bytes);
    }
  }
}

// Ragged launches: n_valid[row] is the stream's sample count for the whole call (negative: the stream is not
// called at all and nothing of it is touched except out_len = 0); this launch covers [offset, offset + a.n).
__device__ __forceinline__ void ragged_count(const DemodArgs& a, int row, bool& active, long& n_l) {
  const long v = (long)a.n_valid[row];
  if (v < 0) {
    active = false;
    if (!a.append) a.out_len[row] = 0;
    return;
  }
  const long r = v - a.n_valid_offset;
  n_l = r < 0 ? 0 : (r < a.n ? r : a.n);
}
// per-stream debug counters of ragged calls (fsk.ts:195-196); uniform calls are counted on the host
__device__ __forceinline__ void ragged_account(const DemodArgs& a, int li, long n_l) {
  double* f = a.f64 + li;
  const long ns = a.n_local;
  if (a.count_call) f[F_RAGGED_CALLS * ns] += 1.0;
  f[F_RAGGED_TOTAL * ns] += (double)n_l;
}

// ---- TMA staging (cp.async.bulk.tensor): one instruction by one lane moves a whole 32-stream x 32-sample tile
// from the [rows][stride] sample buffer into shared memory, already in the 128-byte swizzle that tile_index()
// describes (16-byte chunk index XOR (row & 7)); completion is signalled on an mbarrier.  Out-of-range rows and
// samples are zero-filled by the hardware.
__device__ __forceinline__ void tma_bar_init(uint64_t* bar) {
  const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(b) : "memory");
}
__device__ __forceinline__ void tma_load_tile(float* dst, const CUtensorMap* tmap, uint64_t* bar, int x, int y) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
  // the buffer was last touched through the generic proxy (it doubles as the amplitude buffer of an earlier tile)
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(b), "r"(kTile * kTile * 4) : "memory");
#ifndef WAM_AB_NO_L2HINT
  // the sample stream is read exactly once: evict_first, so it does not displace the sync rings from L2
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;\n"
               ::"r"(d), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(x), "r"(y), "r"(b), "l"(l2_evict_first())
               : "memory");
#else
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n"
               ::"r"(d), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(x), "r"(y), "r"(b)
               : "memory");
#endif
}
__device__ __forceinline__ void tma_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(b), "r"(parity) : "memory");
  } while (!done);
}

// Time-slab hand-over between launches (see launch_slabbed): the CTA that ran a slab publishes its streams' state.
__device__ __noinline__ void slab_publish(int* flag, int slab) {
  __threadfence();  // this lane's state / ring / output stores before the publication
  __syncwarp();
  if ((threadIdx.x & 31) == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flag), "r"(slab + 1) : "memory");
}

// Grid: one warp (32 streams) per CTA, so that 2048 warps spread evenly over 148 SMs.
// GENERIC = false: the common case (no AGC write-back, no tap, integral sync ring, eod_count > 16) —
// only the event-driven state machine is compiled in, which keeps the kernel's code footprint small.
// GENERIC = true: write-back / tap by run-time flag and the per-sample state machine as fallback.
// STAGE_TMA: tiles arrive by TMA instead of per-lane cp.async (rows of the group contiguous, buffer 16-byte aligned).
template <bool ALIGNED, bool GENERIC, bool STAGE_TMA = false>
__global__ void __launch_bounds__(32, WAM_DEMOD_MIN_BLOCKS) fsk_demod_exact_kernel(const __grid_constant__ DemodLaunch L) {
  int gi = 0;
#pragma unroll
  for (int i = 1; i < kMaxGroupsPerLaunch; ++i)
    if (i < L.n_groups && (int)blockIdx.x >= L.block_begin[i]) gi = i;
  const DemodArgs& a = L.g[gi];
  // stage buffers: input tile (swizzled f32 [32][32]); after A1 the same 4 KiB hold the squared
  // magnitudes / amplitudes of the tile as f64 [16][32]
  __shared__ __align__(1024) float tiles[kStages][kTile * kTile];  // 1024: the 128-byte swizzle atom of TMA
  __shared__ __align__(8) uint64_t tma_bar[kStages];
  __shared__ __align__(128) float pfbuf[kTile * 32];  // pre-filtered samples [i][lane]
  __shared__ double park_d[kParkD][32];
  __shared__ uint32_t park_u[kParkU][32];
  __shared__ int rows[32];

  const int lane = threadIdx.x;
  // Time slabs (host: launch_slabbed): a long call arrives as launches of kSlabTiles tiles that overlap on two
  // streams; a CTA of slab j waits here until the CTA of slab j - 1 has published the state of the same 32 streams.
  // Every lane spins with an acquire load, at most 2^22 rounds of 256 ns (about a second): a predecessor that never
  // becomes resident (foreign kernels holding the SM slots) must not hang the GPU — on expiry the CTA's streams are
  // left untouched and WAM_ERR_SLAB_TIMEOUT is recorded (errorEvents).  One opaque asm block: a C++ loop here costs
  // the kernel 48 bytes of spills.
  if (STAGE_TMA && !GENERIC && L.slab_done != nullptr && L.slab > 0) {
    int expired;
    asm volatile("{\n\t.reg .pred p, q;\n\t.reg .s32 v;\n\t.reg .u32 c;\n\tmov.u32 c, 0;\n$L_slab_wait:\n\t"
                 "ld.acquire.gpu.global.s32 v, [%1];\n\t"
                 "setp.lt.s32 p, v, %2;\n\tadd.u32 c, c, 1;\n\tsetp.lt.u32 q, c, 4194304;\n\tand.pred q, p, q;\n\t"
                 "@q nanosleep.u32 256;\n\t@q bra $L_slab_wait;\n\tselp.s32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(expired) : "l"(L.slab_done + blockIdx.x), "r"(L.slab) : "memory");
    if (expired) {
      const int lt = a.l_begin + ((int)blockIdx.x - L.block_begin[gi]) * 32 + lane;
      if (lt < a.l_end) a.u32[(long)U_ERR * a.n_local + lt] |= WAM_ERR_SLAB_TIMEOUT;
      return;
    }
  }
  int li = a.l_begin + ((int)blockIdx.x - L.block_begin[gi]) * 32 + lane;
  bool active = li < a.l_end;
  if (!STAGE_TMA && a.sel != nullptr) {
    // sub-selection (float64 re-run of the streams the fast kernel flagged): entry j of the list names the stream;
    // the list length lives in device memory, the grid is sized for the worst case and surplus CTAs leave at once
    const int cnt = *a.sel_count;
    if (li - lane >= cnt) return;
    active = li < cnt;
    li = active ? a.sel[li] : a.l_begin;
  }
  int row = -1;
  if (active) row = (a.ids ? a.ids[li] : a.id0 + li) - a.row_base;
  long n_l = a.n;  // this stream's samples in this launch (ragged launches run the GENERIC variant)
  if (GENERIC && active && a.n_valid) ragged_count(a, row, active, n_l);
  rows[lane] = active ? row : -1;

  const FskDerived& d = a.d;
  const long ns = a.n_local;
  A2State s;
  s.lo_c = 1.0; s.lo_s = 0.0;
  s.ix1 = s.ix2 = s.iy1 = s.iy2 = s.qx1 = s.qx2 = s.qy1 = s.qy2 = 0.0;
  s.ox1 = s.ox2 = s.oy1 = s.oy2 = s.last_phase = s.iacc = s.qacc = 0.0;
  s.dsc = 0;
  if (active) {
    const double* f = a.f64 + li;
    const uint32_t* u = a.u32 + li;
    s.lo_c = f[F_LO_C * ns]; s.lo_s = f[F_LO_S * ns];
    s.ix1 = f[F_IX1 * ns]; s.ix2 = f[F_IX2 * ns]; s.iy1 = f[F_IY1 * ns]; s.iy2 = f[F_IY2 * ns];
    s.qx1 = f[F_QX1 * ns]; s.qx2 = f[F_QX2 * ns]; s.qy1 = f[F_QY1 * ns]; s.qy2 = f[F_QY2 * ns];
    s.ox1 = f[F_OX1 * ns]; s.ox2 = f[F_OX2 * ns]; s.oy1 = f[F_OY1 * ns]; s.oy2 = f[F_OY2 * ns];
    s.last_phase = f[F_LAST_PHASE * ns]; s.iacc = f[F_IACC * ns]; s.qacc = f[F_QACC * ns];
    s.dsc = u[U_DSC * ns];
    A1State a1;
    a1.gain = f[F_GAIN * ns]; a1.py1 = f[F_PY1 * ns]; a1.py2 = f[F_PY2 * ns];
    a1.px1 = (float)f[F_PX1 * ns]; a1.px2 = (float)f[F_PX2 * ns];
    a1_store(a1, park_d, park_u, lane);
    BState b;
    b.sil_thr = f[F_SIL_THR * ns];
    b.gsc = u[U_GSC * ns]; b.gmod = u[U_GMOD * ns]; b.bsc = u[U_BSC * ns]; b.next_idx = u[U_NEXT_IDX * ns];
    b.bit_acc = u[U_BIT_ACC * ns]; b.bit_cnt = u[U_BIT_CNT * ns]; b.started = u[U_STARTED * ns];
    b.bitpos = (int)u[U_BITPOS * ns]; b.current = u[U_CURRENT * ns]; b.sil_cnt = u[U_SIL_CNT * ns];
    b.ring_pos = u[U_RING_POS * ns]; b.ring_len = u[U_RING_LEN * ns];
    b.amp_pos = u[U_AMP_POS * ns]; b.amp_len = u[U_AMP_LEN * ns];
    b.out_n = a.append ? a.out_len[row] : 0;
    b.cur_word = 0u;
    if (!d.ring_fractional && (b.ring_pos & 31u) != 0u) {
      const uint32_t w = ring_of(a, li)[(b.ring_pos >> 5) & (uint32_t)(d.ring_words - 1)];
      b.cur_word = w & ((1u << (b.ring_pos & 31u)) - 1u);
    }
    b_store(b, park_d, park_u, lane);
  }
  __syncwarp();
  uint8_t* out_row = active ? a.out + (long)row * a.out_stride : nullptr;
  const bool WRITEBACK = GENERIC && a.writeback != 0;
  const bool TAP = GENERIC && a.tap != nullptr;
  float* tap_row = (TAP && active) ? a.tap + (long)row * a.stride : nullptr;
  const bool fast_sm = !GENERIC || (!d.ring_fractional && d.eod_count > 16 && !a.force_generic);

  unsigned long long cyc_a1 = 0, cyc_a2 = 0, cyc_b = 0;
#ifdef WAM_PHASE_TIMING
  const bool timing = a.phase_cycles != nullptr;
#else
  constexpr bool timing = false;  // build with -DWAM_PHASE_TIMING for per-phase SM cycle counters
#endif
  const long long clk_begin = timing ? clock64() : 0;
  const long n_tiles = (a.n + kTile - 1) / kTile;
  const int tma_row0 = a.id0 + a.l_begin + ((int)blockIdx.x - L.block_begin[gi]) * 32 - a.row_base;
  if (STAGE_TMA) {
    if (lane == 0) {
#pragma unroll
      for (int p = 0; p < kStages; ++p) tma_bar_init(&tma_bar[p]);
      asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncwarp();
  }
  for (int p = 0; p < kStages - 1; ++p) {
    if (STAGE_TMA) {
      if (p < n_tiles && lane == 0) tma_load_tile(tiles[p], &L.tmap[gi], &tma_bar[p], p * kTile, tma_row0);
    } else {
      if (p < n_tiles) stage_tile<ALIGNED>(tiles[p], a, rows, (long)p * kTile, lane);
      cp_async_commit();
    }
  }
  for (long t = 0; t < n_tiles; ++t) {
    const long tn = t + kStages - 1;
    if (STAGE_TMA) {
      if (tn < n_tiles && lane == 0)
        tma_load_tile(tiles[tn % kStages], &L.tmap[gi], &tma_bar[tn % kStages], (int)(tn * kTile), tma_row0);
      tma_wait(&tma_bar[t % kStages], (uint32_t)((t / kStages) & 1));
    } else {
      if (tn < n_tiles) stage_tile<ALIGNED>(tiles[tn % kStages], a, rows, tn * kTile, lane);
      cp_async_commit();
      cp_async_wait<kStages - 1>();
    }
    __syncwarp();
    float* tile = tiles[t % kStages];
    double* pbuf = reinterpret_cast<double*>(tile);  // [k][lane] after A1
    const long t0 = t * kTile;
    const int len = GENERIC ? (int)max(0L, min((long)kTile, n_l - t0))  // per lane in ragged launches
                            : (int)min((long)kTile, a.n - t0);

    // ---------------- A1: AGC + pre-filter ----------------
    long long clk0 = timing ? clock64() : 0;
    if (active) {
      A1State a1;
      a1_load(a1, park_d, park_u, lane);
      const bool agc = d.agc_enabled != 0;
      const double att = d.agc_attack, rel = d.agc_release;
      if (len == kTile && !WRITEBACK && !TAP) {
#pragma unroll 1
        for (int ch = 0; ch < 8; ch += kA1Chunks) {
#pragma unroll
          for (int c = 0; c < kA1Chunks; ++c) {
            const float4 v = *reinterpret_cast<const float4*>(tile + tile_index(lane, (ch + c) * 4));
            float sg;
            const float p0 = phase_a1_sample(a1, v.x, d, agc, att, rel, sg);
            const float p1 = phase_a1_sample(a1, v.y, d, agc, att, rel, sg);
            const float p2 = phase_a1_sample(a1, v.z, d, agc, att, rel, sg);
            const float p3 = phase_a1_sample(a1, v.w, d, agc, att, rel, sg);
            float* pfp = pfbuf + ((ch + c) * 4) * 32 + lane;
            pfp[0] = p0; pfp[32] = p1; pfp[64] = p2; pfp[96] = p3;
          }
        }
      } else {
#pragma unroll 1
        for (int i = 0; i < len; ++i) {
          float sg;
          const float pf = phase_a1_sample(a1, tile[tile_index(lane, i)], d, agc, att, rel, sg);
          pfbuf[i * 32 + lane] = pf;
          if (WRITEBACK) a.samples[(long)row * a.stride + t0 + i] = sg;  // fsk.ts:55 mutates the input
          if (TAP) tap_row[t0 + i] = pf;
        }
      }
      a1_store(a1, park_d, park_u, lane);
    }
    __syncwarp();  // every lane is done with the input tile; its storage becomes pbuf
    if (timing) { const long long c = clock64(); cyc_a1 += (unsigned long long)(c - clk0); clk0 = c; }

    // ---------------- A2 + B with replay on resetState() ----------------
    // virtual sample index v = i + dsc0; pair k = v >> 1; a pair's second half emits decimated k.
    const int dsc0 = active ? (int)s.dsc : 0;
    const int v_hi = dsc0 + len;
    const int nk = v_hi >> 1;       // decimated outputs completed inside this tile
    int k_from = 0;                 // first pair A2 has to (re)compute
    int b_from = 0;                 // first decimated output B has to consume
    int v_lo = dsc0;                // first virtual sample present
    uint32_t bits = 0u;
    bool redo = active && (!GENERIC || len > 0);
    // ring bookkeeping at the start of the tile (event-driven state machine)
    const uint32_t pos_t0 = park_u[10][lane], len_t0 = park_u[11][lane];
    const uint32_t slot_t0 = park_u[12][lane], alen_t0 = park_u[13][lane];
    if (active && (!GENERIC || len > 0)) {
      // renormalise the LO rotation (one Newton step towards |(c, s)| = 1)
      const double m = fma(s.lo_c, s.lo_c, s.lo_s * s.lo_s);
      const double f = fma(-0.5, m, 1.5);
      s.lo_c *= f; s.lo_s *= f;
    }
    while (__any_sync(0xffffffffu, redo)) {
      if (redo) {
        bits &= (1u << k_from) - 1u;
        if (dsc0 == 0 && (v_hi & 1) == 0) {
          // fast path: every pair is complete (two pairs per iteration: the biquad histories rotate in
          // place and two atan2 chains overlap)
#pragma unroll kA2Unroll
          for (int k = k_from; k < nk; ++k) {
            double yi0, yq0, yi1, yq1, pp;
            const float* pfp = pfbuf + (2 * k) * 32 + lane;
            phase_a2_half(s, pfp[0], d, yi0, yq0);
            phase_a2_half(s, pfp[32], d, yi1, yq1);
            const int bit = phase_a2_decim(s, yi0 + yi1, yq0 + yq1, d, pp);
            bits |= (uint32_t)bit << k;
            pbuf[k * 32 + lane] = 0.5 * fast_sqrt(pp);  // amplitude (fsk.ts:252)
          }
        } else {
#pragma unroll 1
          for (int k = k_from; 2 * k < v_hi; ++k) {
            const int v0 = 2 * k, v1 = 2 * k + 1;
            double yi, yq;
            if (v0 >= v_lo) {
              phase_a2_half(s, pfbuf[(v0 - dsc0) * 32 + lane], d, yi, yq);
              s.iacc = yi; s.qacc = yq;  // 0 + y
            }
            if (v1 < v_hi) {
              phase_a2_half(s, pfbuf[(v1 - dsc0) * 32 + lane], d, yi, yq);
              double pp;
              const int bit = phase_a2_decim(s, s.iacc + yi, s.qacc + yq, d, pp);
              s.iacc = 0.0; s.qacc = 0.0;
              bits |= (uint32_t)bit << k;
              pbuf[k * 32 + lane] = 0.5 * fast_sqrt(pp);
            }
          }
        }
        s.dsc = (uint32_t)(v_hi & 1);
        if (timing) { const long long c = clock64(); cyc_a2 += (unsigned long long)(c - clk0); clk0 = c; }
        // ---------------- B ----------------
        BState b;
        b_load(b, park_d, park_u, lane);
        redo = false;
        int k_reset = -1;
        if (fast_sm) {
          k_reset = sm_tile_events<false>(b, bits, pbuf + lane, b_from, nk, pos_t0, len_t0, slot_t0, alen_t0, a, li,
                                          out_row, nullptr, 0);
          if (k_reset < 0 || 2 * (k_reset + 1) >= v_hi) {
            b.ring_len = min(len_t0 + (uint32_t)nk, (uint32_t)d.ring_cap_int);
            const uint32_t sl = slot_t0 + (uint32_t)nk;
            b.amp_pos = sl >= (uint32_t)d.amp_phys ? sl - (uint32_t)d.amp_phys : sl;
            b.amp_len = min(alen_t0 + (uint32_t)nk, (uint32_t)d.amp_cap);
          }
        } else if (GENERIC) {
#pragma unroll 1
          for (int k = b_from; k < nk; ++k) {
            if (sm_sample_generic(b, (int)((bits >> k) & 1u), pbuf[k * 32 + lane], a, li, out_row)) {
              k_reset = k;
              break;
            }
          }
        }
        if (k_reset >= 0) {
          // resetState(): A2 restarts from the zeroed state at the next pair
          reset_state_a2(s);
          k_from = k_reset + 1; b_from = k_reset + 1; v_lo = 2 * (k_reset + 1);
          redo = (v_lo < v_hi);
        }
        b_store(b, park_d, park_u, lane);
        if (timing) { const long long c = clock64(); cyc_b += (unsigned long long)(c - clk0); clk0 = c; }
      }
    }
    // every lane's generic-proxy accesses to this tile's buffer are ordered before the TMA write that recycles it
    if (STAGE_TMA) asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    __syncwarp();
  }
  if (!STAGE_TMA) cp_async_wait<0>();
  if (timing && lane == 0) {
    unsigned long long* pc = a.phase_cycles + 4ull * blockIdx.x;
    pc[0] += cyc_a1; pc[1] += cyc_a2; pc[2] += cyc_b;
    pc[3] += (unsigned long long)(clock64() - clk_begin) - cyc_a1 - cyc_a2 - cyc_b;
  }

  if (active) {
    double* f = a.f64 + li;
    uint32_t* u = a.u32 + li;
    f[F_LO_C * ns] = s.lo_c; f[F_LO_S * ns] = s.lo_s;
    f[F_IX1 * ns] = s.ix1; f[F_IX2 * ns] = s.ix2; f[F_IY1 * ns] = s.iy1; f[F_IY2 * ns] = s.iy2;
    f[F_QX1 * ns] = s.qx1; f[F_QX2 * ns] = s.qx2; f[F_QY1 * ns] = s.qy1; f[F_QY2 * ns] = s.qy2;
    f[F_OX1 * ns] = s.ox1; f[F_OX2 * ns] = s.ox2; f[F_OY1 * ns] = s.oy1; f[F_OY2 * ns] = s.oy2;
    f[F_LAST_PHASE * ns] = s.last_phase; f[F_IACC * ns] = s.iacc; f[F_QACC * ns] = s.qacc;
    u[U_DSC * ns] = s.dsc;
    A1State a1;
    a1_load(a1, park_d, park_u, lane);
    f[F_GAIN * ns] = a1.gain; f[F_PY1 * ns] = a1.py1; f[F_PY2 * ns] = a1.py2;
    f[F_PX1 * ns] = (double)a1.px1; f[F_PX2 * ns] = (double)a1.px2;
    BState b;
    b_load(b, park_d, park_u, lane);
    f[F_SIL_THR * ns] = b.sil_thr;
    u[U_GSC * ns] = b.gsc; u[U_GMOD * ns] = b.gmod; u[U_BSC * ns] = b.bsc; u[U_NEXT_IDX * ns] = b.next_idx;
    u[U_BIT_ACC * ns] = b.bit_acc; u[U_BIT_CNT * ns] = b.bit_cnt; u[U_STARTED * ns] = b.started;
    u[U_BITPOS * ns] = (uint32_t)b.bitpos; u[U_CURRENT * ns] = b.current; u[U_SIL_CNT * ns] = b.sil_cnt;
    u[U_RING_POS * ns] = b.ring_pos; u[U_RING_LEN * ns] = b.ring_len;
    u[U_AMP_POS * ns] = b.amp_pos; u[U_AMP_LEN * ns] = b.amp_len;
    if (!d.ring_fractional && (b.ring_pos & 31u) != 0u)
      ring_of(a, li)[(b.ring_pos >> 5) & (uint32_t)(d.ring_words - 1)] = b.cur_word;
    a.out_len[row] = b.out_n < a.out_stride ? b.out_n : (int)a.out_stride;
    if (GENERIC && a.n_valid) ragged_account(a, li, n_l);
  }
  if (STAGE_TMA && !GENERIC && L.slab_done != nullptr) slab_publish(L.slab_done + blockIdx.x, L.slab);
}

}  // namespace wam
