// fsk_mod.cuh — phase-continuous FSK modulator (NCO), FSKCore.modulateData (src/modems/fsk.ts:377-424).
//
// The reference walks one phase accumulator through the whole frame (fsk.ts:398-405).  Here the
// phase of any sample is available in closed form, so time parallelises freely:
//   kernel 1 (mark_prefix): per stream, the number of mark ('1') line bits before every byte —
//       the integer form of "phase carried across blocks";
//   kernel 2 (modulate):    one thread per 4 output samples; phase in cycles =
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

__device__ __forceinline__ float mod_sample(const ModArgs& a, const uint8_t* row, const uint32_t* pre, long k,
                                            long pad, long body) {
  if (k < pad || k >= pad + body) return 0.0f;  // lead padding / tail silence stay zero (fsk.ts:392-395)
  const FskDerived& d = a.d;
  const long m = k - pad;
  const int bitidx = (int)(m / d.spb);
  const int r = (int)(m - (long)bitidx * d.spb);
  const int byteidx = bitidx / d.bpb;
  const int bib = bitidx - byteidx * d.bpb;
  const int byte = frame_byte(a, row, byteidx);
  int marks = (int)pre[byteidx];
  for (int b = 0; b < bib; ++b) marks += framed_bit(d, byte, b);
  const int cur = framed_bit(d, byte, bib);
  const double spaces = (double)(bitidx - marks);
  const double fcur = cur ? d.mark : d.space;
  // cycles since the start of the frame; products are exact for integral tone frequencies
  double cyc = ((double)d.spb * ((double)marks * d.mark + spaces * d.space) + (double)r * fcur) / d.fs;
  cyc -= floor(cyc);
  return sinpif(2.0f * (float)cyc);
}

// grid: (ceil(max_total/(128*4)), n_streams); one thread = 4 consecutive samples
__global__ void __launch_bounds__(128) fsk_modulate_kernel(const __grid_constant__ ModArgs a) {
  const int s = blockIdx.y;
  const int nbytes = a.data_len ? a.data_len[s] : a.nbytes;
  const long total_bytes = (long)a.d.n_preamble + a.d.n_sfd + nbytes;
  const long pad = total_bytes > 0 ? 2L * a.d.spb : 0;
  const long body = total_bytes * a.d.bpb * a.d.spb;
  const long total = body + pad + (long)a.d.bpb * a.d.spb;
  const uint8_t* row = a.data + (long)s * a.data_stride;
  const uint32_t* pre = a.prefix + (long)s * a.prefix_stride;
  float* out = a.out + (long)s * a.out_stride;
  if (blockIdx.x == 0 && threadIdx.x == 0 && a.out_len) a.out_len[s] = (int32_t)(total < a.out_stride ? total : a.out_stride);
  const long k0 = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const long lim = total < a.out_stride ? total : a.out_stride;
  if (k0 >= lim) return;
  float v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = mod_sample(a, row, pre, k0 + i, pad, body);
  if (a.vec_ok && k0 + 4 <= lim) {
    *reinterpret_cast<float4*>(out + k0) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
    for (int i = 0; i < 4; ++i)
      if (k0 + i < lim) out[k0 + i] = v[i];
  }
}

}  // namespace wam
