// fsk_mod.cuh — phase-continuous FSK modulator (NCO), FSKCore.modulateData (src/modems/fsk.ts:377-424).
//
// The reference walks one phase accumulator through the whole frame (fsk.ts:398-405).  Here the
// phase of any sample is available in closed form, so time parallelises freely:
//   kernel 1 (bit_phase): one CTA per stream; a block scan gives the number of mark ('1') line bits
//       before every byte — the integer form of "phase carried across blocks" — and every line bit
//       gets one table word: the phase at its first sample, in cycles =
//       frac(spb * (marks_before*f_mark + spaces_before*f_space) / fs), evaluated in float64 (the
//       products are exact integers for integral tone frequencies), stored as a 31-bit binary
//       fraction with the bit's value in bit 0.  The table is 1/spb of the output in size.
//   (fsk_modulate_fused_kernel does both in one launch with the table in shared memory; the two-kernel form below
//   remains for frames whose table does not fit and for rows that are not float4-aligned)
//   kernel 2 (modulate): HBM-write bound.  One thread per float4 of output, a warp per 512
//       contiguous bytes; the phase inside a bit advances in 32-bit fixed point (wraps modulo one
//       cycle for free), so a sample costs IMAD + I2F + FMUL + MUFU.SIN.
// Layout: data [stream][data_stride] u8, out [stream][out_stride] f32 (float4 stores, coalesced).
#pragma once

#include "wam_common.cuh"

namespace wam {

struct ModArgs {
  FskDerived d;
  const uint8_t* data;   // [n_streams][data_stride]
  long data_stride;
  const int32_t* data_len;  // nullable
  int nbytes;               // used when data_len == nullptr
  int n_streams;
  float* out;               // [n_streams][out_stride]
  long out_stride;
  int32_t* out_len;         // nullable
  uint32_t* bittab;         // [n_streams][tab_stride]: per line bit, phase at its first sample | bit value
  int tab_stride;
  int vec_ok;               // out rows 16-byte aligned
  uint32_t step_fix[2];     // round(f / fs * 2^32) for space (0) and mark (1): cycles per sample, fixed point
  uint32_t spb_magic;       // floor(2^32 / spb): division by multiply-high plus one fix-up
};

__device__ __forceinline__ int frame_byte(const ModArgs& a, const uint8_t* row, int k) {
  const int nps = a.d.n_preamble + a.d.n_sfd;
  return k < nps ? a.d.preamble_sfd[k] : row[k - nps];
}

// line bit `b` (0..bpb-1) of a framed byte: start bits 0, 8 data bits MSB first, parity, stop bits 1
// (fsk.ts:408-420)
__device__ __forceinline__ int framed_bit(const FskDerived& d, int byte, int b) {
  if (b < d.start_bits) return 0;
  b -= d.start_bits;
  if (b < 8) return (byte >> (7 - b)) & 1;
  b -= 8;
  if (d.parity != 0) {
    if (b == 0) {
      const int p = __popc((unsigned)byte & 0xffu) & 1;
      return d.parity == 1 ? p : 1 - p;
    }
    b -= 1;
  }
  return 1;
}
__device__ __forceinline__ int framed_ones(const FskDerived& d, int byte) {
  int ones = __popc((unsigned)byte & 0xffu) + d.stop_bits;
  if (d.parity != 0) {
    const int p = __popc((unsigned)byte & 0xffu) & 1;
    ones += d.parity == 1 ? p : 1 - p;
  }
  return ones;
}

// number of mark bits among line bits 0..b-1 of a framed byte
__device__ __forceinline__ int framed_ones_before(const FskDerived& d, int byte, int b) {
  int rest = b - d.start_bits;
  if (rest <= 0) return 0;
  const int nb = rest < 8 ? rest : 8;
  int ones = __popc(((unsigned)byte & 0xffu) >> (8 - nb));
  rest -= 8;
  if (rest > 0 && d.parity != 0) {
    const int p = __popc((unsigned)byte & 0xffu) & 1;
    ones += d.parity == 1 ? p : 1 - p;
    rest -= 1;
  }
  if (rest > 0) ones += rest;  // stop bits
  return ones;
}

constexpr int kPhaseThreads = 128;

// Table words of one stream (see the header): NT threads of one CTA; bytes in chunks of NT — a block scan gives the
// mark bits before every byte of the chunk, then the threads fill the chunk's words one line bit each.
// tab: global (fsk_bit_phase_kernel) or shared memory (fsk_modulate_fused_kernel).
template <int NT>
__device__ __forceinline__ void bit_phase_fill(const ModArgs& a, int s, uint32_t* tab, uint32_t* s_pre, uint8_t* s_byte,
                                               uint32_t* s_wsum) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const FskDerived& d = a.d;
  const int nbytes = a.data_len ? a.data_len[s] : a.nbytes;
  const int total = d.n_preamble + d.n_sfd + nbytes;
  const uint8_t* row = a.data + (long)s * a.data_stride;
  const double spb_d = (double)d.spb, inv_fs = 1.0 / d.fs;
  // i / bpb as a multiply-high: magic = ceil(2^32 / bpb) is exact while i * (magic * bpb - 2^32) < 2^32, i.e. for every
  // i < 2^32 / bpb — the line bits of a chunk number NT * bpb at most
  const uint32_t bpb_magic = (uint32_t)((0x100000000ull + (uint32_t)d.bpb - 1u) / (uint32_t)d.bpb);
  uint32_t carry = 0;
  for (int base = 0; base < total; base += NT) {
    const int k = base + tid;
    const int byte = (k < total) ? frame_byte(a, row, k) : 0;
    const uint32_t v = (k < total) ? (uint32_t)framed_ones(d, byte) : 0u;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_wsum[wid] = incl;
    __syncthreads();
    uint32_t woff = 0, wtot = 0;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) {
      const uint32_t t = s_wsum[w];
      if (w < wid) woff += t;
      wtot += t;
    }
    s_pre[tid] = carry + woff + incl - v;
    s_byte[tid] = (uint8_t)byte;
    __syncthreads();
    const int nb = min(NT, total - base);
    for (int i = tid; i < nb * d.bpb; i += NT) {
      const int kk = (int)__umulhi((uint32_t)i, bpb_magic);
      const int b = i - kk * d.bpb;
      const int by = s_byte[kk];
      const int marks = (int)s_pre[kk] + framed_ones_before(d, by, b);
      const int bitidx = base * d.bpb + i;
      // cycles since the start of the frame at the first sample of this bit: an exact integer (for integral
      // tone frequencies) times 1/fs, i.e. ~1e-13 cycles of rounding at most
      const double spaces = (double)(bitidx - marks);
      const double cyc = (spb_d * fma((double)marks, d.mark, spaces * d.space)) * inv_fs;
      const double fr = cyc - floor(cyc);
      const uint32_t fix = __double2uint_rz(fr * 4294967296.0);
      tab[bitidx] = (fix & ~1u) | (uint32_t)framed_bit(d, by, b);
    }
    carry += wtot;
    __syncthreads();
  }
}

// one CTA per stream: the table in global memory (frames whose table does not fit in shared memory, or rows
// that are not float4-aligned)
__global__ void __launch_bounds__(kPhaseThreads) fsk_bit_phase_kernel(const __grid_constant__ ModArgs a) {
  __shared__ uint32_t s_pre[kPhaseThreads];
  __shared__ uint8_t s_byte[kPhaseThreads];
  __shared__ uint32_t s_wsum[kPhaseThreads / 32];
  const int s = blockIdx.x;
  bit_phase_fill<kPhaseThreads>(a, s, a.bittab + (long)s * a.tab_stride, s_pre, s_byte, s_wsum);
}

constexpr int kModThreads = 256;

// sin(2 pi * p / 2^32) for a 32-bit fixed-point phase: the signed reinterpretation is the angle in
// (-pi, pi], where MUFU.SIN is accurate to ~5e-7 absolute
__device__ __forceinline__ float sin_fix(uint32_t p) {
  return __sinf((float)(int32_t)p * 1.4629180792671596e-9f);  // 2 pi / 2^32
}

// generic float4: crosses a bit boundary or an end of the frame body (one sample at a time)
__device__ __noinline__ float4 mod_straddle(const ModArgs& a, const uint32_t* __restrict__ tab, uint32_t k0,
                                            uint32_t pad, uint32_t body_end) {
  const uint32_t spb = (uint32_t)a.d.spb;
  float w[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  const uint32_t kf = k0 > pad ? k0 : pad;  // first body sample of this float4
  const uint32_t m0 = kf - pad;
  uint32_t q = __umulhi(m0, a.spb_magic);
  uint32_t r = m0 - q * spb;
  if (r >= spb) { ++q; r -= spb; }
  uint32_t e = (kf < body_end) ? __ldg(tab + q) : 0u;
#pragma unroll
  for (uint32_t i = 0; i < 4u; ++i) {
    const uint32_t k = k0 + i;
    if (k >= kf && k < body_end) {
      w[i] = sin_fix((e & ~1u) + r * a.step_fix[e & 1u]);
      if (++r == spb) {
        r = 0;
        ++q;
        if (k + 1u < body_end) e = __ldg(tab + q);
      }
    }
  }
  return make_float4(w[0], w[1], w[2], w[3]);
}

__device__ __forceinline__ void mod_divmod(uint32_t m, uint32_t spb, uint32_t magic, uint32_t& q, uint32_t& r) {
  q = __umulhi(m, magic);  // floor(2^32 / spb) undershoots the quotient by at most one
  r = m - q * spb;
  if (r >= spb) { ++q; r -= spb; }
}

// grid: (blocks per row, n_streams); a block walks its row with stride gridDim.x * kModThreads float4s, so
// every warp store covers 512 contiguous bytes.  Sample indices inside a row are 32-bit (the host rejects
// frames of 2^31 samples or more).  The (bit, offset) pair of a thread's next float4 advances incrementally.
// ALIGNED (spb % 4 == 0, rows 16-byte aligned, out_stride % 4 == 0): every float4 lies inside one line bit
// or inside the zero padding, so the loop body is branch-light: table word, fixed-point phase, four MUFU.SIN,
// one streaming 16-byte store.
template <bool ALIGNED>
__global__ void __launch_bounds__(kModThreads) fsk_modulate_kernel(const __grid_constant__ ModArgs a) {
  const int s = blockIdx.y;
  const FskDerived& d = a.d;
  const int nbytes = a.data_len ? a.data_len[s] : a.nbytes;
  const uint32_t spb = (uint32_t)d.spb;
  const uint32_t total_bytes = (uint32_t)(d.n_preamble + d.n_sfd + nbytes);
  const uint32_t pad = total_bytes > 0 ? 2u * spb : 0u;  // fsk.ts:392
  const uint32_t body_end = pad + total_bytes * (uint32_t)d.bpb * spb;
  const uint32_t total = body_end + (uint32_t)d.bpb * spb;
  const uint32_t lim = (long)total < a.out_stride ? total : (uint32_t)a.out_stride;
  const uint32_t* __restrict__ tab = a.bittab + (long)s * a.tab_stride;
  float* __restrict__ out = a.out + (long)s * a.out_stride;
  if (blockIdx.x == 0 && threadIdx.x == 0 && a.out_len) a.out_len[s] = (int32_t)lim;
  const uint32_t st0 = a.step_fix[0], st1 = a.step_fix[1];
  const uint32_t stride = gridDim.x * (uint32_t)(4 * kModThreads);
  uint32_t k0 = (blockIdx.x * (uint32_t)kModThreads + threadIdx.x) * 4u;
  // (q, r) = divmod(k0 - pad, spb), kept valid while k0 >= pad; (dq, dr) = divmod(stride, spb)
  uint32_t q = 0, r = 0, dq, dr;
  mod_divmod(stride, spb, a.spb_magic, dq, dr);
  if (ALIGNED) {
    // pad is 0 or 2 * spb: divide k0 itself and shift the bit index by pad / spb
    mod_divmod(k0, spb, a.spb_magic, q, r);
    const uint32_t* __restrict__ tabq = tab - (pad ? 2 : 0);
    const uint32_t body_len = body_end - pad;
    const uint32_t lim4 = lim & ~3u;
    float* po = out + k0;
    // the table word of the NEXT float4 is requested before this one is evaluated (the table is read once,
    // from HBM: without the prefetch every iteration waits out a full DRAM latency)
    uint32_t e = (k0 < lim4 && k0 - pad < body_len) ? __ldg(tabq + q) : 0u;
    for (; k0 < lim4; k0 += stride, po += stride) {
      uint32_t qn = q + dq, rn = r + dr;
      if (rn >= spb) { rn -= spb; ++qn; }
      const uint32_t kn = k0 + stride;
      const uint32_t en = (kn < lim4 && kn - pad < body_len) ? __ldg(tabq + qn) : 0u;
      float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      if (k0 - pad < body_len) {  // unsigned: also false for k0 < pad
        const uint32_t st = (e & 1u) ? st1 : st0;
        const uint32_t p0 = (e & ~1u) + r * st;
        v.x = sin_fix(p0); v.y = sin_fix(p0 + st); v.z = sin_fix(p0 + 2u * st); v.w = sin_fix(p0 + 3u * st);
      }
      __stcs(reinterpret_cast<float4*>(po), v);
      q = qn; r = rn; e = en;
    }
    // a row cut short by out_stride can end inside a float4
    if (k0 < lim) {
      const float4 v = mod_straddle(a, tab, k0, pad, body_end);
      const float w[4] = {v.x, v.y, v.z, v.w};
      for (uint32_t i = 0; k0 + i < lim; ++i) out[k0 + i] = w[i];
    }
    return;
  }
  bool have_qr = false;
  const bool vec = a.vec_ok != 0;
  for (; k0 < lim; k0 += stride) {
    float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);  // lead padding / tail silence stay zero (fsk.ts:392-395)
    if (k0 >= pad) {
      if (!have_qr) { mod_divmod(k0 - pad, spb, a.spb_magic, q, r); have_qr = true; }
      if (k0 + 4u <= body_end) {
        if (r + 4u <= spb) {
          // the whole float4 lies inside one line bit
          const uint32_t e = __ldg(tab + q);
          const uint32_t st = (e & 1u) ? st1 : st0;
          const uint32_t p0 = (e & ~1u) + r * st;
          v.x = sin_fix(p0); v.y = sin_fix(p0 + st); v.z = sin_fix(p0 + 2u * st); v.w = sin_fix(p0 + 3u * st);
        } else {
          v = mod_straddle(a, tab, k0, pad, body_end);
        }
      } else if (k0 < body_end) {
        v = mod_straddle(a, tab, k0, pad, body_end);
      }
      q += dq; r += dr;
      if (r >= spb) { r -= spb; ++q; }
    } else if (k0 + 4u > pad && pad < body_end) {
      v = mod_straddle(a, tab, k0, pad, body_end);
    }
    if (vec && k0 + 4u <= lim) {
      __stcs(reinterpret_cast<float4*>(out + k0), v);
    } else {
      const float w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (k0 + i < lim) out[k0 + i] = w[i];
    }
  }
}

// Both steps in one launch (the common case: float4-aligned rows, spb % 4 == 0, table <= 40 KB): every CTA builds
// the table of ITS stream in shared memory, then walks its part of the row.  No table traffic through HBM, no
// second launch; with one CTA per row (many rows) the table is built exactly once per row.
__global__ void __launch_bounds__(kModThreads) fsk_modulate_fused_kernel(const __grid_constant__ ModArgs a) {
  extern __shared__ uint32_t s_tab[];
  __shared__ uint32_t s_pre[kModThreads];
  __shared__ uint8_t s_byte[kModThreads];
  __shared__ uint32_t s_wsum[kModThreads / 32];
  const int s = blockIdx.y;
  bit_phase_fill<kModThreads>(a, s, s_tab, s_pre, s_byte, s_wsum);  // ends with __syncthreads()
  const FskDerived& d = a.d;
  const int nbytes = a.data_len ? a.data_len[s] : a.nbytes;
  const uint32_t spb = (uint32_t)d.spb;
  const uint32_t total_bytes = (uint32_t)(d.n_preamble + d.n_sfd + nbytes);
  const uint32_t pad = total_bytes > 0 ? 2u * spb : 0u;  // fsk.ts:392
  const uint32_t body_end = pad + total_bytes * (uint32_t)d.bpb * spb;
  const uint32_t total = body_end + (uint32_t)d.bpb * spb;
  const uint32_t lim = (long)total < a.out_stride ? total : (uint32_t)a.out_stride;
  float* __restrict__ out = a.out + (long)s * a.out_stride;
  if (blockIdx.x == 0 && threadIdx.x == 0 && a.out_len) a.out_len[s] = (int32_t)lim;
  const uint32_t st0 = a.step_fix[0], st1 = a.step_fix[1];
  const uint32_t stride = gridDim.x * (uint32_t)(4 * kModThreads);
  uint32_t k0 = (blockIdx.x * (uint32_t)kModThreads + threadIdx.x) * 4u;
  uint32_t q, r, dq, dr;
  mod_divmod(stride, spb, a.spb_magic, dq, dr);
  mod_divmod(k0, spb, a.spb_magic, q, r);  // pad is 0 or 2 * spb: divide k0 itself, shift the bit index by pad / spb
  const uint32_t* tabq = s_tab - (pad ? 2 : 0);
  const uint32_t body_len = body_end - pad;
  const uint32_t lim4 = lim & ~3u;
  float* po = out + k0;
  for (; k0 < lim4; k0 += stride, po += stride) {
    float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);  // lead padding / tail silence stay zero (fsk.ts:392-395)
    if (k0 - pad < body_len) {  // unsigned: also false for k0 < pad
      const uint32_t e = tabq[q];
      const uint32_t st = (e & 1u) ? st1 : st0;
      const uint32_t p0 = (e & ~1u) + r * st;
      v.x = sin_fix(p0); v.y = sin_fix(p0 + st); v.z = sin_fix(p0 + 2u * st); v.w = sin_fix(p0 + 3u * st);
    }
    __stcs(reinterpret_cast<float4*>(po), v);
    q += dq; r += dr;
    if (r >= spb) { r -= spb; ++q; }
  }
  // a row cut short by out_stride can end inside a float4
  if (k0 < lim) {
    float w[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    uint32_t qq, rr;
    for (uint32_t i = 0; k0 + i < lim; ++i) {
      const uint32_t k = k0 + i;
      if (k - pad < body_len) {
        mod_divmod(k - pad, spb, a.spb_magic, qq, rr);
        const uint32_t e = s_tab[qq];
        w[i] = sin_fix((e & ~1u) + rr * ((e & 1u) ? st1 : st0));
      }
      out[k] = w[i];
    }
  }
}

}  // namespace wam
