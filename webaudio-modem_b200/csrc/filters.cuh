// filters.cuh — batched IIRFilter / FIRFilter.processBuffer (src/dsp/filters.ts:8-167).
//
// IIR: ONE pass over the samples (4 B read + 4 B written per sample), parallel in time as well as across streams.
// y[n] = sum b_i x[n-i] - sum a_i y[n-i] is linear in the output history, so a stretch of samples run from a ZERO
// output history (the input history is known) plus a matrix power applied to the true history gives the true result:
//   chunk   kIirChunk samples, one lane: run from zero history -> end state Z_c;
//   span    32 chunks, one warp: Kogge-Stone scan over the lanes with warp shuffles and the precomputed powers
//           T^(2^k), T = A^kIirChunk (A = companion matrix of the feedback taps, float64, from the host);
//   tile    kIirWarps spans, one CTA (a contiguous 32 KiB stretch of one stream staged in padded shared memory with
//           16-byte loads): the span end states combine through shared memory;
//   stream  tiles are handed out by a ticket counter, tile index major (every stream's tile 0, then every tile 1, ...):
//           a tile publishes its aggregate (zero-start end state), looks back over its predecessors' aggregates /
//           inclusive states (decoupled look-back, the transition matrices multiply up on the way) and publishes its
//           inclusive state.  With many streams the predecessor finished long ago and the look-back is one load;
//           with few long streams the tiles of a stream run side by side and the chain is walked;
// then every lane re-runs its chunk from the true history, in place in shared memory, and the warp writes the span out.
// FIR: shared-memory window in float64 (converted once), eight outputs per thread with the window sliding through
// registers (one shared-memory load per eight multiply-adds), taps in the reference's order (filters.ts:129-136).
#pragma once

#include "wam_common.cuh"

namespace wam {

constexpr int kMaxIirTaps = 8;     // nb, na <= 8  (order <= 7)
constexpr int kMaxM = kMaxIirTaps - 1;
constexpr int kIirChunk = 32;      // samples per lane chunk
constexpr int kIirWarpSpan = 32 * kIirChunk;
constexpr int kIirWarps = 8;       // spans per tile
constexpr int kIirTile = kIirWarps * kIirWarpSpan;
constexpr int kIirPitch = kIirChunk + 1;
constexpr int kMaxFirTaps = 1024;

struct IirArgs {
  int nb, na;
  double b[kMaxIirTaps], a[kMaxIirTaps];
  double tpow[5][kMaxM * kMaxM];           // T^(2^k), k = 0..4, row-major M x M with M = na - 1
  double tspan[2][kMaxM * kMaxM];          // T^32 (one span) and (T^32)^kIirWarps (one tile)
  const float* in;     // [n_streams][stride]
  float* out;
  long stride, n;
  long n_streams;
  int tiles;           // tiles per stream = ceil(n / kIirTile)
  int vec;             // rows allow 16-byte accesses
  double* state;       // nullable [n_streams][(nb-1)+(na-1)]: x[n-1..], y[n-1..]; read at the start, written at the end
  // decoupled look-back, one record per (stream, tile)
  int* tile_flag;      // 0 nothing yet, 1 aggregate published, 2 inclusive state published
  double* tile_aggr;   // [..][M] end state of the tile run from a zero output history
  double* tile_incl;   // [..][M] true output history behind the tile
  unsigned int* ticket;
};

size_t iir_scratch_bytes(long n, long n_streams);
int iir_process_batch_device(IirArgs& ia, const float* d_in, float* d_out, long stride, long n, long n_streams, double* d_state,
                             void* d_scratch, size_t scratch_bytes, cudaStream_t st);
int iir_process_batch_host(IirArgs& ia, const float* in, float* out, long stride, long n, long n_streams, double* state);
int fir_process_batch_device(const double* d_taps, int ntaps, const float* d_in, float* d_out, long stride, long n,
                             long n_streams, double* d_state, double* d_state_new, cudaStream_t st);
int fir_process_batch_host(const double* taps, int ntaps, const float* in, float* out, long stride, long n,
                           long n_streams, double* state);

// v <- Tm * u   (M x M); MT >= 0 fixes M at compile time
template <int MT>
__device__ __forceinline__ void mat_apply(const double* Tm, int M, const double* u, double* v) {
#pragma unroll
  for (int r = 0; r < kMaxM; ++r) {  // (predicated, not broken off: the vectors stay in registers)
    if (r < (MT >= 0 ? MT : M)) {
      double acc = 0.0;
#pragma unroll
      for (int c = 0; c < kMaxM; ++c)
        if (c < (MT >= 0 ? MT : M)) acc += Tm[r * M + c] * u[c];
      v[r] = acc;
    }
  }
}

// Run the recurrence over one chunk held in (padded) shared memory.
//   xs: this lane's chunk, len samples; xh: input history x[start-1], x[start-2], ...
//   yh: output history (most recent first), updated in place; outputs optionally written over the inputs.
template <bool WRITE, int NXT, int MT>
__device__ __forceinline__ void iir_run_chunk(const IirArgs& a, float* xs, int len, double* xh, double* yh) {
  const int NX = NXT >= 0 ? NXT : a.nb - 1, M = MT >= 0 ? MT : a.na - 1;
#pragma unroll 8
  for (int i = 0; i < len; ++i) {
    const double x0 = (double)xs[i];
    double y = a.b[0] * x0;  // accumulation order of filters.ts:56-66
#pragma unroll
    for (int k = 1; k < kMaxIirTaps; ++k)
      if (k <= NX) y += a.b[k] * xh[k - 1];
#pragma unroll
    for (int k = 1; k < kMaxIirTaps; ++k)
      if (k <= M) y -= a.a[k] * yh[k - 1];
#pragma unroll
    for (int k = kMaxM - 1; k > 0; --k) {
      if (k < NX) xh[k] = xh[k - 1];
      if (k < M) yh[k] = yh[k - 1];
    }
    if (NX > 0) xh[0] = x0;
    if (M > 0) yh[0] = y;
    if (WRITE) xs[i] = (float)y;  // Float32Array store, filters.ts:82-85
  }
}

// Kogge-Stone over the lanes: v_l <- sum_{j<=l} T^(l-j) v_j
template <int MT>
__device__ __forceinline__ void iir_lane_scan(const IirArgs& a, int M, int lane, double* v) {
  double u[kMaxM], w[kMaxM];
#pragma unroll
  for (int st = 0; st < 5; ++st) {
#pragma unroll
    for (int k = 0; k < kMaxM; ++k)
      if (k < (MT >= 0 ? MT : M)) u[k] = __shfl_up_sync(0xffffffffu, v[k], 1 << st);
    if (lane >= (1 << st)) {
      mat_apply<MT>(a.tpow[st], M, u, w);
#pragma unroll
      for (int k = 0; k < kMaxM; ++k)
        if (k < (MT >= 0 ? MT : M)) v[k] += w[k];
    }
  }
}

// grid: n_streams * tiles CTAs of kIirWarps warps; tiles are taken by ticket, a stream's tiles in order, so that a tile
// only ever waits for tiles whose CTAs are already running.  NXT / MT: nb - 1 / na - 1 at compile time (the biquads of
// FilterFactory), or -1: read from the arguments.
template <int NXT, int MT>
__global__ void __launch_bounds__(kIirWarps * 32) iir_scan_kernel(const __grid_constant__ IirArgs a) {
  extern __shared__ float smem[];  // [kIirWarps][32][kIirPitch]
  __shared__ unsigned s_ticket;
  __shared__ double s_z[kIirWarps][kMaxM];  // span end states from a zero output history
  __shared__ double s_y[kIirWarps][kMaxM];  // true output history at the start of every span
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int M = MT >= 0 ? MT : a.na - 1;
  const int NX = NXT >= 0 ? NXT : a.nb - 1;
  const int stw = NX + M;
  if (threadIdx.x == 0) s_ticket = atomicAdd(a.ticket, 1u);
  __syncthreads();
  const long ticket = (long)s_ticket;
  const int ti = (int)(ticket / a.n_streams);
  const long s = ticket - (long)ti * a.n_streams;
  const long tile = s * a.tiles + ti;  // record index: the predecessor tile of the same stream is tile - 1
  const long t0 = (long)ti * kIirTile;
  const long tile_len = min((long)kIirTile, a.n - t0);
  const float* in = a.in + s * a.stride;
  float* out = a.out + s * a.stride;

  // ---- stage the tile: coalesced 16-byte loads into padded chunk rows
  if (a.vec) {
    const float4* src = reinterpret_cast<const float4*>(in + t0);
#pragma unroll 4
    for (long q = threadIdx.x; q < tile_len / 4; q += kIirWarps * 32) {
      const float4 v = __ldcs(src + q);
      const int i = (int)q * 4, c = i / kIirChunk, o = i % kIirChunk;
      float* row = smem + c * kIirPitch + o;
      row[0] = v.x; row[1] = v.y; row[2] = v.z; row[3] = v.w;
    }
    for (long i = tile_len / 4 * 4 + threadIdx.x; i < tile_len; i += kIirWarps * 32)
      smem[(int)(i / kIirChunk) * kIirPitch + (int)(i % kIirChunk)] = in[t0 + i];
  } else {
    for (long i = threadIdx.x; i < tile_len; i += kIirWarps * 32)
      smem[(int)(i / kIirChunk) * kIirPitch + (int)(i % kIirChunk)] = in[t0 + i];
  }
  __syncthreads();

  const int chunk = w * 32 + lane;
  const long start = t0 + (long)chunk * kIirChunk;
  const int len = (int)max(0L, min((long)kIirChunk, a.n - start));
  float* xs = smem + chunk * kIirPitch;

  // input history of this chunk: earlier samples of the stream, else the carried state
  double xh0[kMaxM], xh[kMaxM], yh[kMaxM], zc[kMaxM], v[kMaxM];
#pragma unroll
  for (int k = 0; k < kMaxM; ++k) {
    xh0[k] = 0.0;
    if (k < NX) {
      const long idx = start - 1 - k;
      const int rel = chunk * kIirChunk - 1 - k;  // position inside the staged tile
      if (rel >= 0) xh0[k] = (idx < a.n) ? (double)smem[(rel / kIirChunk) * kIirPitch + rel % kIirChunk] : 0.0;
      else if (idx >= 0) xh0[k] = (idx < a.n) ? (double)in[idx] : 0.0;
      else if (a.state) xh0[k] = a.state[s * stw + (int)(-idx - 1)];
    }
    xh[k] = xh0[k]; yh[k] = 0.0;
  }
  // ---- zero-history run of every chunk, scan inside the warp
  iir_run_chunk<false, NXT, MT>(a, xs, len, xh, yh);
#pragma unroll
  for (int k = 0; k < kMaxM; ++k) { zc[k] = (k < M && len > 0) ? yh[k] : 0.0; v[k] = zc[k]; }
  // (partial or empty chunks only occur at the very end of a stream, where nothing downstream uses the carry)
  iir_lane_scan<MT>(a, M, lane, v);
  if (lane == 31)
    for (int k = 0; k < M; ++k) s_z[w][k] = v[k];
  __syncthreads();

  // ---- between spans and tiles: one thread
  if (threadIdx.x == 0) {
    const long rec = tile * (long)max(M, 1);
    double zt[kMaxM], y0[kMaxM], t[kMaxM];
    for (int k = 0; k < kMaxM; ++k) { zt[k] = (k < M) ? s_z[0][k] : 0.0; y0[k] = 0.0; }
    for (int j = 1; j < kIirWarps; ++j) {  // tile aggregate, Horner over the spans
      mat_apply<MT>(a.tspan[0], M, zt, t);
      for (int k = 0; k < M; ++k) zt[k] = t[k] + s_z[j][k];
    }
    if (ti == 0) {
      for (int k = 0; k < M; ++k) y0[k] = a.state ? a.state[s * stw + NX + k] : 0.0;
    } else {
      for (int k = 0; k < M; ++k) a.tile_aggr[rec + k] = zt[k];
      __threadfence();
      asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(a.tile_flag + tile), "r"(1) : "memory");
      // look back: y0 = sum over predecessors of (transition so far) * (their aggregate), closed by an inclusive state
      double P[kMaxM * kMaxM];
      for (int r = 0; r < M; ++r)
        for (int c = 0; c < M; ++c) P[r * M + c] = r == c ? 1.0 : 0.0;
      const double* TT = a.tspan[1];
      for (long j = tile - 1;; --j) {
        int f;
        do {
          asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(f) : "l"(a.tile_flag + j) : "memory");
          if (f == 0) __nanosleep(64);
        } while (f == 0);
        const double* src = (f == 2 ? a.tile_incl : a.tile_aggr) + j * (long)max(M, 1);
        double u[kMaxM];
        for (int k = 0; k < M; ++k) u[k] = src[k];
        mat_apply<MT>(P, M, u, t);
        for (int k = 0; k < M; ++k) y0[k] += t[k];
        if (f == 2) break;  // (the stream's first tile always publishes an inclusive state)
        double Q[kMaxM * kMaxM];
        for (int r = 0; r < M; ++r)
          for (int c = 0; c < M; ++c) {
            double acc = 0.0;
            for (int k = 0; k < M; ++k) acc += P[r * M + k] * TT[k * M + c];
            Q[r * M + c] = acc;
          }
        for (int k = 0; k < M * M; ++k) P[k] = Q[k];
      }
    }
    if (ti + 1 < a.tiles) {  // inclusive state = transition of the whole tile applied to y0, plus the aggregate
      mat_apply<MT>(a.tspan[1], M, y0, t);
      for (int k = 0; k < M; ++k) a.tile_incl[rec + k] = t[k] + zt[k];
      __threadfence();
      asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(a.tile_flag + tile), "r"(2) : "memory");
    }
    for (int k = 0; k < M; ++k) s_y[0][k] = y0[k];
    for (int j = 1; j < kIirWarps; ++j) {
      double u[kMaxM];
      for (int k = 0; k < M; ++k) u[k] = s_y[j - 1][k];
      mat_apply<MT>(a.tspan[0], M, u, t);
      for (int k = 0; k < M; ++k) s_y[j][k] = t[k] + s_z[j - 1][k];
    }
  }
  __syncthreads();

  // ---- true history per chunk: fold the span's start into chunk 0, scan again, shift by one lane
  double s0[kMaxM], t[kMaxM];
#pragma unroll
  for (int k = 0; k < kMaxM; ++k) { s0[k] = (k < M) ? s_y[w][k] : 0.0; v[k] = zc[k]; }
  if (lane == 0) {
    mat_apply<MT>(a.tpow[0], M, s0, t);
#pragma unroll
    for (int k = 0; k < kMaxM; ++k)
      if (k < M) v[k] += t[k];
  }
  iir_lane_scan<MT>(a, M, lane, v);
#pragma unroll
  for (int k = 0; k < kMaxM; ++k) {
    const double prev = (k < M) ? __shfl_up_sync(0xffffffffu, v[k], 1) : 0.0;
    yh[k] = (lane == 0) ? s0[k] : prev;
    xh[k] = xh0[k];
  }
  iir_run_chunk<true, NXT, MT>(a, xs, len, xh, yh);
  if (a.state && len > 0 && start + len == a.n) {
    for (int k = 0; k < NX; ++k) a.state[s * stw + k] = xh[k];
    for (int k = 0; k < M; ++k) a.state[s * stw + NX + k] = yh[k];
  }
  __syncwarp();
  // ---- the warp writes its span
  const long span0 = (long)w * kIirWarpSpan;
  const long span_len = max(0L, min((long)kIirWarpSpan, tile_len - span0));
  const float* rows = smem + w * 32 * kIirPitch;
  if (a.vec) {
    float4* dst = reinterpret_cast<float4*>(out + t0 + span0);
    for (long q = lane; q < span_len / 4; q += 32) {
      const int i = (int)q * 4;
      const float* row = rows + (i / kIirChunk) * kIirPitch + (i % kIirChunk);
      __stcs(dst + q, make_float4(row[0], row[1], row[2], row[3]));
    }
    for (long i = span_len / 4 * 4 + lane; i < span_len; i += 32)
      out[t0 + span0 + i] = rows[(int)(i / kIirChunk) * kIirPitch + (int)(i % kIirChunk)];
  } else {
    for (long i = lane; i < span_len; i += 32) out[t0 + span0 + i] = rows[(int)(i / kIirChunk) * kIirPitch + (int)(i % kIirChunk)];
  }
}

// ---- FIR -------------------------------------------------------------------------------------
struct FirArgs {
  const double* taps;  // device [ntaps]
  int ntaps;
  const float* in;
  float* out;
  long stride, n;
  long n_streams;
  int tiles;           // tiles per stream
  int vec;
  double* state;       // nullable [n_streams][ntaps-1], most recent first; read at the start, written by the last tile
  double* state_new;   // scratch of the same shape (the last tile writes here, a copy kernel moves it over)
};

constexpr int kFirThreads = 128;
constexpr int kFirPer = 8;                          // outputs per thread
constexpr int kFirTile = kFirThreads * kFirPer;     // outputs per CTA
// window element e lives at e + e / 8: a warp's loads of elements 8 apart fall into 16 different 8-byte bank pairs
// (two wavefronts, the minimum for 256 bytes)
__device__ __forceinline__ int fir_slot(int e) { return e + (e >> 3); }
__host__ __device__ __forceinline__ int fir_taps8(int ntaps) { return (ntaps + 7) & ~7; }
// dynamic shared memory of fir_kernel, in doubles
__host__ __device__ __forceinline__ int fir_smem_doubles(int ntaps) {
  const int wn = kFirTile + fir_taps8(ntaps);
  return fir_taps8(ntaps) + wn + (wn >> 3) + 2;
}

// grid: n_streams * tiles CTAs.  Shared memory: the taps padded with zeros to a multiple of eight, then the window
// x[t0 - H .. t0 + kFirTile), H = the padded tap count (element e = x[t0 - H + e] at fir_slot(e)).
__global__ void __launch_bounds__(kFirThreads) fir_kernel(const __grid_constant__ FirArgs a) {
  extern __shared__ double dsm[];
  const int H = fir_taps8(a.ntaps);
  double* taps = dsm;
  double* win = dsm + H;
  const long s = blockIdx.x / a.tiles;
  const long t0 = (long)(blockIdx.x - s * a.tiles) * kFirTile;
  const float* in = a.in + s * a.stride;
  const int hist = a.ntaps - 1;
  for (int k = threadIdx.x; k < H; k += kFirThreads) taps[k] = k < a.ntaps ? a.taps[k] : 0.0;
  // history part of the window: earlier samples of the stream, else the carried state, else zeros
  for (int e = threadIdx.x; e < H; e += kFirThreads) {
    const long idx = t0 - H + e;
    double v = 0.0;
    if (idx >= 0) v = (double)in[idx];
    else if (a.state && -idx - 1 < hist) v = a.state[s * hist + (-idx - 1)];
    win[fir_slot(e)] = v;
  }
  // the tile's own samples: 16-byte loads where the tile is whole
  if (a.vec && t0 + kFirTile <= a.n) {
    const float4* src = reinterpret_cast<const float4*>(in + t0);
#pragma unroll
    for (int q = threadIdx.x; q < kFirTile / 4; q += kFirThreads) {
      const float4 v = __ldcs(src + q);
      double* d = win + fir_slot(H + 4 * q);  // H + 4 q is a multiple of four: the four slots are contiguous
      d[0] = (double)v.x; d[1] = (double)v.y; d[2] = (double)v.z; d[3] = (double)v.w;
    }
  } else {
    for (int e = threadIdx.x; e < kFirTile; e += kFirThreads) win[fir_slot(H + e)] = (t0 + e < a.n) ? (double)in[t0 + e] : 0.0;
  }
  __syncthreads();
  // outputs o = 8 t + r: acc[r] = sum_k taps[k] * x[t0 + o - k], k ascending (filters.ts:133-136).  R[(r - k) & 7]
  // holds x[t0 + 8 t + (r - k)]: each step one register leaves the window and is refilled with the next older sample.
  // Element of x[t0 + 8 t] is H + 8 t, a multiple of eight like k0, so the eight refills of one round sit at fixed
  // offsets from one pointer that moves back nine doubles per round.
  double R[kFirPer], acc[kFirPer];
  const double* p = win + fir_slot(H + kFirPer * (int)threadIdx.x);
#pragma unroll
  for (int r = 0; r < kFirPer; ++r) { R[r] = p[r]; acc[r] = 0.0; }
  p -= 1;  // slot of element (base - 8) + 8, i.e. refill u reads p[-(u + 1)] ... see below
  for (int k0 = 0; k0 < H; k0 += kFirPer) {
    // elements base - k0 - 1 .. base - k0 - 8 share the quotient (base - k0) / 8 - 1: slots q - 1 .. q - 8 with
    // q = (base - k0) + (base - k0) / 8 - 1 = p
#pragma unroll
    for (int u = 0; u < kFirPer; ++u) {
      const double c = taps[k0 + u];
#pragma unroll
      for (int r = 0; r < kFirPer; ++r) acc[r] += c * R[(r - u) & (kFirPer - 1)];
      R[(kFirPer - 1 - u) & (kFirPer - 1)] = p[-(u + 1)];
    }
    p -= kFirPer + 1;
  }
  const long t = t0 + kFirPer * threadIdx.x;
  float* out = a.out + s * a.stride + t;
  if (a.vec && t + kFirPer <= a.n) {
    __stcs(reinterpret_cast<float4*>(out), make_float4((float)acc[0], (float)acc[1], (float)acc[2], (float)acc[3]));
    __stcs(reinterpret_cast<float4*>(out) + 1, make_float4((float)acc[4], (float)acc[5], (float)acc[6], (float)acc[7]));
  } else {
#pragma unroll
    for (int r = 0; r < kFirPer; ++r)
      if (t + r < a.n) out[r] = (float)acc[r];
  }
  if (a.state_new && t0 + kFirTile >= a.n) {
    // new history: x[n-1-k], falling back to the old history when the call was shorter than it
    for (int k = threadIdx.x; k < hist; k += kFirThreads) {
      const long idx = a.n - 1 - k;
      double v = 0.0;
      if (idx >= 0) v = (double)in[idx];
      else if (a.state) v = a.state[s * hist + (-idx - 1)];
      a.state_new[s * hist + k] = v;
    }
  }
}

}  // namespace wam
