// filters.cuh — batched IIRFilter / FIRFilter.processBuffer (src/dsp/filters.ts:8-167).
//
// IIR: time-chunked linear-recurrence scan, so one long stream parallelises in time as well as
// across streams.  y[n] = sum b_i x[n-i] - sum a_i y[n-i] is split into chunks of kIirChunk
// samples; a warp owns 32 consecutive chunks of one stream (a contiguous 16 KiB span staged
// through padded shared memory):
//   reduce: every lane runs its chunk from a ZERO output history (the input history is known)
//           and keeps the final output-history vector Z_c;
//   scan:   the true history at the start of chunk c obeys  Y_{c+1} = T * Y_c + Z_c  with
//           T = A^kIirChunk (A = companion matrix of the feedback taps, computed on the host in
//           float64).  Inside the warp this is a Kogge-Stone scan over lanes with warp shuffles
//           and the precomputed powers T^(2^k); between warps of the same stream the 32-chunk
//           carry is passed through a small global array (second kernel);
//   apply:  every lane re-runs its chunk from the true history and writes float32 outputs
//           (Float32Array store of filters.ts:82-85).
// FIR: shared-memory-staged direct convolution, float64 accumulation in tap order
// (filters.ts:129-136).
#pragma once

#include "wam_common.cuh"

namespace wam {

constexpr int kMaxIirTaps = 8;     // nb, na <= 8  (order <= 7)
constexpr int kIirChunk = 128;     // samples per lane chunk
constexpr int kIirWarpSpan = 32 * kIirChunk;
constexpr int kMaxFirTaps = 1024;

struct IirArgs {
  int nb, na;
  double b[kMaxIirTaps], a[kMaxIirTaps];
  // T^(2^k), k = 0..5, row-major M x M with M = na - 1; T = A^kIirChunk (k = 5: one warp span)
  double tpow[6][(kMaxIirTaps - 1) * (kMaxIirTaps - 1)];
  const float* in;     // [n_streams][stride]
  float* out;
  long stride, n;
  int n_streams;
  int spans;           // warp spans per stream = ceil(n / kIirWarpSpan)
  const double* state_in;   // nullable [n_streams][(nb-1)+(na-1)]: x[n-1..], y[n-1..]
  double* state_out;        // nullable
  double* chunk_z;     // [n_streams][spans][32][M]: zero-history end state of every chunk
  double* span_z;      // [n_streams][spans][M]: zero-start end state of every warp span
  double* span_start;  // [n_streams][spans][M]: true output history at the start of every span
};

int iir_process_batch_host(IirArgs& ia, const float* in, float* out, long stride, long n, long n_streams, double* state);
int fir_process_batch_host(const double* taps, int ntaps, const float* in, float* out, long stride, long n,
                           long n_streams, double* state);

// v <- Tm * u   (M x M)
__device__ __forceinline__ void mat_apply(const double* Tm, int M, const double* u, double* v) {
  for (int r = 0; r < kMaxIirTaps - 1; ++r) {
    if (r >= M) break;
    double acc = 0.0;
    for (int c = 0; c < kMaxIirTaps - 1; ++c) {
      if (c >= M) break;
      acc += Tm[r * M + c] * u[c];
    }
    v[r] = acc;
  }
}

// Run the recurrence over one chunk held in (padded) shared memory.
//   xs: this lane's chunk, len samples; xh: input history x[start-1], x[start-2], ...
//   yh: output history (most recent first), updated in place; outputs optionally written.
template <bool WRITE>
__device__ __forceinline__ void iir_run_chunk(const IirArgs& a, const float* xs, int len, double* xh, double* yh,
                                              float* out) {
  const int nb = a.nb, M = a.na - 1;
  for (int i = 0; i < len; ++i) {
    const double x0 = (double)xs[i];
    double y = a.b[0] * x0;  // accumulation order of filters.ts:56-66
    for (int k = 1; k < kMaxIirTaps; ++k) {
      if (k >= nb) break;
      y += a.b[k] * xh[k - 1];
    }
    for (int k = 1; k < kMaxIirTaps; ++k) {
      if (k > M) break;
      y -= a.a[k] * yh[k - 1];
    }
    for (int k = kMaxIirTaps - 2; k > 0; --k) {
      if (k < nb - 1) xh[k] = xh[k - 1];
      if (k < M) yh[k] = yh[k - 1];
    }
    if (nb > 1) xh[0] = x0;
    if (M > 0) yh[0] = y;
    if (WRITE) out[i] = (float)y;
  }
}

// PHASE: 0 = reduce+intra-warp scan (writes span_z), 2 = apply (reads span_start, writes out)
// grid: (spans, n_streams), block: 32 threads = one warp span of 32 chunks.
template <int PHASE>
__global__ void __launch_bounds__(32) iir_span_kernel(const __grid_constant__ IirArgs a) {
  extern __shared__ float smem[];  // [32][kIirChunk + 1]
  const int lane = threadIdx.x;
  const int span = blockIdx.x;
  const int s = blockIdx.y;
  const int M = a.na - 1;
  const int NX = a.nb - 1;
  const long span0 = (long)span * kIirWarpSpan;
  const long span_len = min((long)kIirWarpSpan, a.n - span0);
  const float* in = a.in + (long)s * a.stride;

  // coalesced stage of the contiguous span into padded rows
  for (long i = lane; i < span_len; i += 32) {
    const int c = (int)(i / kIirChunk), o = (int)(i % kIirChunk);
    smem[c * (kIirChunk + 1) + o] = in[span0 + i];
  }
  __syncwarp();

  const long start = span0 + (long)lane * kIirChunk;
  const int len = (int)max(0L, min((long)kIirChunk, a.n - start));
  const int stw = NX + M;

  // input history for this chunk: earlier samples of the stream, else the carried state
  double xh[kMaxIirTaps - 1], yh[kMaxIirTaps - 1];
  for (int k = 0; k < kMaxIirTaps - 1; ++k) {
    xh[k] = 0.0; yh[k] = 0.0;
    if (k < NX) {
      const long idx = start - 1 - k;
      if (idx >= 0) xh[k] = (idx < a.n) ? (double)in[idx] : 0.0;
      else if (a.state_in) xh[k] = a.state_in[(long)s * stw + (int)(-idx - 1)];
    }
  }

  double v[kMaxIirTaps - 1], u[kMaxIirTaps - 1], w[kMaxIirTaps - 1], s0[kMaxIirTaps - 1];
  double* zc = a.chunk_z + (((long)s * a.spans + span) * 32 + lane) * M;
  if (PHASE == 0) {
    // zero-history run of every chunk; Z_c = its final output history
    iir_run_chunk<false>(a, smem + lane * (kIirChunk + 1), len, xh, yh, nullptr);
    for (int k = 0; k < kMaxIirTaps - 1; ++k) {
      v[k] = (k < M && len > 0) ? yh[k] : 0.0;
      if (k < M) zc[k] = v[k];
      s0[k] = 0.0;
    }
  } else {
    for (int k = 0; k < kMaxIirTaps - 1; ++k) {
      v[k] = (k < M) ? zc[k] : 0.0;
      s0[k] = (k < M) ? a.span_start[((long)s * a.spans + span) * M + k] : 0.0;
    }
    if (lane == 0) {  // fold the span's starting history into chunk 0: Z_0' = Z_0 + T * S
      mat_apply(a.tpow[0], M, s0, w);
      for (int k = 0; k < kMaxIirTaps - 1; ++k)
        if (k < M) v[k] += w[k];
    }
  }
  // Kogge-Stone over lanes with warp shuffles: v_l = sum_{j<=l} T^(l-j) Z_j.  (Partial or empty
  // chunks only occur at the very end of a stream, where nothing downstream consumes the carry.)
#pragma unroll
  for (int st = 0; st < 5; ++st) {
    for (int k = 0; k < kMaxIirTaps - 1; ++k) u[k] = __shfl_up_sync(0xffffffffu, v[k], 1 << st);
    if (lane >= (1 << st)) {
      mat_apply(a.tpow[st], M, u, w);
      for (int k = 0; k < kMaxIirTaps - 1; ++k)
        if (k < M) v[k] += w[k];
    }
  }
  if (PHASE == 0) {
    if (lane == 31)
      for (int k = 0; k < M; ++k) a.span_z[((long)s * a.spans + span) * M + k] = v[k];
  } else {
    // history at the start of chunk `lane` = end state of chunk lane-1 (lane 0: the span start)
    for (int k = 0; k < kMaxIirTaps - 1; ++k) {
      const double prev = __shfl_up_sync(0xffffffffu, v[k], 1);
      yh[k] = (lane == 0) ? s0[k] : prev;
    }
    float* orow = smem + lane * (kIirChunk + 1);  // outputs overwrite the staged inputs in place
    iir_run_chunk<true>(a, orow, len, xh, yh, orow);
    if (a.state_out && len > 0 && start + len == a.n) {
      for (int k = 0; k < NX; ++k) a.state_out[(long)s * stw + k] = xh[k];
      for (int k = 0; k < M; ++k) a.state_out[(long)s * stw + NX + k] = yh[k];
    }
    __syncwarp();
    float* out = a.out + (long)s * a.stride;
    for (long i = lane; i < span_len; i += 32) {
      const int c = (int)(i / kIirChunk), o = (int)(i % kIirChunk);
      out[span0 + i] = smem[c * (kIirChunk + 1) + o];
    }
  }
}

// Between spans of one stream: Y_{span+1} = T^32 * Y_span + Z_span, one thread per stream walks
// its spans (spans = n / 4096: 704 for a 60 s stream at 48 kHz).  tp32 = T^32 = (T^16)^2.
__global__ void iir_span_carry_kernel(const __grid_constant__ IirArgs a) {
  const double* tp32 = a.tpow[5];
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= a.n_streams) return;
  const int M = a.na - 1;
  const int NX = a.nb - 1;
  double y[kMaxIirTaps - 1], w[kMaxIirTaps - 1];
  for (int k = 0; k < kMaxIirTaps - 1; ++k)
    y[k] = (k < M && a.state_in) ? a.state_in[(long)s * (NX + M) + NX + k] : 0.0;
  for (int sp = 0; sp < a.spans; ++sp) {
    double* dst = a.span_start + ((long)s * a.spans + sp) * M;
    for (int k = 0; k < M; ++k) dst[k] = y[k];
    mat_apply(tp32, M, y, w);
    const double* z = a.span_z + ((long)s * a.spans + sp) * M;
    for (int k = 0; k < M; ++k) y[k] = w[k] + z[k];
  }
}

// ---- FIR -------------------------------------------------------------------------------------
struct FirArgs {
  const double* taps;  // device [ntaps]
  int ntaps;
  const float* in;
  float* out;
  long stride, n;
  const double* state_in;  // nullable [n_streams][ntaps-1], most recent first
  double* state_out;
};

constexpr int kFirTile = 256;

// grid: (ceil(n / kFirTile), n_streams); block kFirTile threads; dynamic smem:
// taps (ntaps doubles) + input window (kFirTile + ntaps - 1 doubles)
__global__ void __launch_bounds__(kFirTile) fir_kernel(const __grid_constant__ FirArgs a) {
  extern __shared__ double dsm[];
  double* taps = dsm;
  double* win = dsm + a.ntaps;  // win[j] = x[t0 - (ntaps-1) + j]
  const int s = blockIdx.y;
  const long t0 = (long)blockIdx.x * kFirTile;
  const float* in = a.in + (long)s * a.stride;
  const int hist = a.ntaps - 1;
  for (int k = threadIdx.x; k < a.ntaps; k += blockDim.x) taps[k] = a.taps[k];
  for (int j = threadIdx.x; j < kFirTile + hist; j += blockDim.x) {
    const long idx = t0 - hist + j;
    double v = 0.0;
    if (idx >= 0) v = idx < a.n ? (double)in[idx] : 0.0;
    else if (a.state_in) v = a.state_in[(long)s * hist + (-idx - 1)];
    win[j] = v;
  }
  __syncthreads();
  const long t = t0 + threadIdx.x;
  if (t < a.n) {
    double acc = 0.0;  // output += c[i] * x[n-i], i ascending (filters.ts:133-136)
    for (int k = 0; k < a.ntaps; ++k) acc += taps[k] * win[hist + threadIdx.x - k];
    a.out[(long)s * a.stride + t] = (float)acc;
  }
  if (a.state_out && t0 + kFirTile >= a.n && threadIdx.x < hist) {
    // new history: x[n-1-k], falling back to the old history when the call was shorter than it
    const long idx = a.n - 1 - threadIdx.x;
    double v = 0.0;
    if (idx >= 0) v = (double)in[idx];
    else if (a.state_in) v = a.state_in[(long)s * hist + (-idx - 1)];
    a.state_out[(long)s * hist + threadIdx.x] = v;
  }
}

}  // namespace wam
