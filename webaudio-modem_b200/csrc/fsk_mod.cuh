// fsk_mod.cuh — phase-continuous FSK modulator (NCO), FSKCore.modulateData (src/modems/fsk.ts:377-424).
//
// The reference walks one phase accumulator through the whole frame (fsk.ts:398-405).  Here the
// phase of any sample is available in closed form, so time parallelises freely:
//   kernel 1 (mark_prefix): per stream, the number of mark ('1') line bits before every byte —
//       the integer form of "phase carried across blocks";
//   kernel 2 (modulate):    one thread per 8 output samples; phase in cycles =
//       (spb * (marks_before*f_mark + spaces_before*f_space) + r * f_bit) / fs, reduced to its
//       fractional part in float64 (exact for integral tone frequencies), then sinpi in float32.
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
  uint32_t* prefix;         // [n_streams][prefix_stride]: mark bits before byte k (k = 0..totalBytes)
  int prefix_stride;
  int vec_ok;               // out rows 16-byte aligned
  float rot_mark_c, rot_mark_s, rot_space_c, rot_space_s;  // cos/sin(2 pi f / fs) of the two tones
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

// one warp per stream: exclusive prefix sum of mark bits per framed byte
__global__ void __launch_bounds__(128) fsk_mark_prefix_kernel(const __grid_constant__ ModArgs a) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= a.n_streams) return;
  const int nbytes = a.data_len ? a.data_len[warp] : a.nbytes;
  const int total = a.d.n_preamble + a.d.n_sfd + nbytes;
  const uint8_t* row = a.data + (long)warp * a.data_stride;
  uint32_t* pre = a.prefix + (long)warp * a.prefix_stride;
  uint32_t carry = 0;
  for (int base = 0; base <= total; base += 32) {
    const int k = base + lane;
    uint32_t v = (k < total) ? (uint32_t)framed_ones(a.d, frame_byte(a, row, k)) : 0u;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (k <= total) pre[k] = carry + incl - v;
    carry += __shfl_sync(0xffffffffu, incl, 31);
  }
}

// Per-thread bit cursor: everything that changes once per line bit.
struct ModCursor {
  int bitidx, bib, byteidx, byte, marks, cur;
  double base;  // fractional cycles at the first sample of the current bit
  double step;  // cycles per sample of the current bit
};

__device__ __forceinline__ void mod_cursor_phase(const FskDerived& d, ModCursor& c) {
  // cycles since the start of the frame at the start of this bit; the products are exact integers for
  // integral tone frequencies, so the only rounding is the division by fs
  const double spaces = (double)(c.bitidx - c.marks);
  double cyc = ((double)d.spb * ((double)c.marks * d.mark + spaces * d.space)) / d.fs;
  c.base = cyc - floor(cyc);
  c.step = (c.cur ? d.mark : d.space) / d.fs;
}

__device__ __forceinline__ void mod_cursor_seek(const ModArgs& a, const uint8_t* row, const uint32_t* pre, int bitidx,
                                                ModCursor& c) {
  const FskDerived& d = a.d;
  c.bitidx = bitidx;
  c.byteidx = bitidx / d.bpb;
  c.bib = bitidx - c.byteidx * d.bpb;
  c.byte = frame_byte(a, row, c.byteidx);
  c.marks = (int)pre[c.byteidx];
  for (int b = 0; b < c.bib; ++b) c.marks += framed_bit(d, c.byte, b);
  c.cur = framed_bit(d, c.byte, c.bib);
  mod_cursor_phase(d, c);
}

__device__ __forceinline__ void mod_cursor_next_bit(const ModArgs& a, const uint8_t* row, ModCursor& c) {
  const FskDerived& d = a.d;
  c.marks += c.cur;
  c.bitidx++;
  if (++c.bib == d.bpb) {
    c.bib = 0;
    c.byteidx++;
    c.byte = frame_byte(a, row, c.byteidx);
  }
  c.cur = framed_bit(d, c.byte, c.bib);
  mod_cursor_phase(d, c);
}

constexpr int kModPerThread = 8;   // consecutive samples per thread (two float4 stores)
constexpr int kModThreads = 128;

// grid: (ceil(max_total / (kModThreads * kModPerThread)), n_streams)
// One division per thread locates its first sample's bit; inside a bit the phase advances linearly, so a
// sample costs one FFMA + sinpif.  The run is re-based in float64 at its start and at every bit boundary,
// which keeps the float32 phase error below 1e-7 cycles whatever the baud rate.
__global__ void __launch_bounds__(kModThreads) fsk_modulate_kernel(const __grid_constant__ ModArgs a) {
  const int s = blockIdx.y;
  const FskDerived& d = a.d;
  const int nbytes = a.data_len ? a.data_len[s] : a.nbytes;
  const long total_bytes = (long)d.n_preamble + d.n_sfd + nbytes;
  const long pad = total_bytes > 0 ? 2L * d.spb : 0;
  const long body = total_bytes * d.bpb * d.spb;
  const long total = body + pad + (long)d.bpb * d.spb;
  const uint8_t* row = a.data + (long)s * a.data_stride;
  const uint32_t* pre = a.prefix + (long)s * a.prefix_stride;
  float* out = a.out + (long)s * a.out_stride;
  if (blockIdx.x == 0 && threadIdx.x == 0 && a.out_len) a.out_len[s] = (int32_t)(total < a.out_stride ? total : a.out_stride);
  const long k0 = ((long)blockIdx.x * blockDim.x + threadIdx.x) * kModPerThread;
  const long lim = total < a.out_stride ? total : a.out_stride;
  if (k0 >= lim) return;

  float v[kModPerThread];
#pragma unroll
  for (int i = 0; i < kModPerThread; ++i) v[i] = 0.0f;  // lead padding / tail silence stay zero (fsk.ts:392-395)
  const long first = k0 > pad ? k0 : pad;                // first body sample of this run
  const long last = (k0 + kModPerThread < pad + body) ? k0 + kModPerThread : pad + body;
  if (first < last) {
    const unsigned m0 = (unsigned)(first - pad);
    const int bitidx = (int)(m0 / (unsigned)d.spb);
    int r = (int)(m0 - (unsigned)bitidx * (unsigned)d.spb);
    ModCursor c;
    mod_cursor_seek(a, row, pre, bitidx, c);
    double ph = fma((double)r, c.step, c.base);
    ph -= floor(ph);
    if (first == k0 && last == k0 + kModPerThread && r + kModPerThread <= d.spb) {
      // whole run inside one line bit: a pure tone.  One sincospi seeds a rotation by the tone's
      // per-sample angle (cos/sin computed on the host in float64); 8 steps add < 1e-6 of error.
      float sn, cs;
      sincospif(2.0f * (float)ph, &sn, &cs);
      const float rc = c.cur ? a.rot_mark_c : a.rot_space_c;
      const float rs = c.cur ? a.rot_mark_s : a.rot_space_s;
#pragma unroll
      for (int i = 0; i < kModPerThread; ++i) {
        v[i] = sn;
        const float ns = fmaf(sn, rc, cs * rs);
        cs = fmaf(cs, rc, -sn * rs);
        sn = ns;
      }
    } else {
      float basef = (float)ph;
      float stepf = (float)c.step;
      int j = 0;  // samples since the last re-base
#pragma unroll
      for (int i = 0; i < kModPerThread; ++i) {
        const long k = k0 + i;
        if (k >= first && k < last) {
          v[i] = sinpif(2.0f * fmaf((float)j, stepf, basef));
          ++j;
          if (++r == d.spb && k + 1 < last) {  // next line bit: re-base in float64
            r = 0;
            mod_cursor_next_bit(a, row, c);
            basef = (float)c.base;
            stepf = (float)c.step;
            j = 0;
          }
        }
      }
    }
  }
  if (a.vec_ok && k0 + kModPerThread <= lim) {
    reinterpret_cast<float4*>(out + k0)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(out + k0)[1] = make_float4(v[4], v[5], v[6], v[7]);
  } else {
#pragma unroll
    for (int i = 0; i < kModPerThread; ++i)
      if (k0 + i < lim) out[k0 + i] = v[i];
  }
}

}  // namespace wam
