// fsk_demod_pipe.cuh — warp-specialised FSK demodulator for FEW streams (long-stream configurations).
//
// fsk_demod_exact_kernel (fsk_demod.cuh) gives every stream one thread that runs the three phases of a tile
// back to back, so a stream advances at the sum of the three dependency chains (about 670 SM cycles per input
// sample).  With tens of thousands of streams the SMs hide that latency across warps; with a few hundred or a
// few thousand streams (BASELINE configs 3 and 4) they cannot, and the kernel runs at the latency floor of a
// single warp whatever the GPU could issue.
//
// Here a CTA is THREE warps that own the same 32 streams and form a software pipeline in time:
//   warp 0 (A1)  stages input tiles (cp.async, double-buffered) and runs AGC + pre-filter      -> pf ring  (smem)
//   warp 1 (A2)  LO mix, I/Q low-pass, /2 decimation, atan2, post filter, slicer, amplitude     -> dec ring (smem)
//   warp 2 (B)   decimated-rate state machine (event-driven, sm_tile_events): rings, EOD, sync, bits, bytes
// so a stream advances at the LONGEST chain instead of their sum, and each role keeps only its own state in
// registers (nothing is parked in shared memory between phases).  The arithmetic is the same device code as the
// fused kernel (phase_a1_sample, phase_a2_half, phase_a2_decim, sm_tile_events): results are identical.
//
// resetState() (fsk.ts:175-188) feeds back from B into A2, which by then has run ahead.  A1 is never reset.
// B publishes a roll-back request {tile, decimated index per lane}; A2 zeroes the state of the lanes concerned
// and re-computes THEIR tiles from the request point (the pre-filtered samples are still in the pf ring: a pf
// slot is only recycled once B has finished with its tile), the other lanes keep what they already produced.
// Lanes therefore carry their own "next tile" position and the warp always works on the minimum.
//
// Synchronisation is by monotonic counters in shared memory (volatile accesses + __threadfence_block):
//   a1_done            tiles A1 has published                    (A2 waits for tile < a1_done)
//   a2_pub             {epoch acknowledged, min over lanes of next tile} in one 64-bit word (B waits on it)
//   b_done             tiles B has finished with                 (A1 and A2 wait for ring space)
//   epoch_req          roll-back requests issued by B            (A2 polls it)
// Every wait loop has a spin limit: on expiry the CTA flags WAM_ERR_PIPE_TIMEOUT for its streams and drains.
#pragma once

#include "fsk_demod.cuh"

namespace wam {

constexpr int kPipePf = 4;    // pf ring depth (tiles)
constexpr int kPipeDec = 4;   // dec ring depth (tiles)
constexpr int kPipeThreads = 96;
#ifndef WAM_PIPE_A1_CHUNKS
#define WAM_PIPE_A1_CHUNKS 1
#endif
#ifndef WAM_PIPE_A2_UNROLL
#define WAM_PIPE_A2_UNROLL 2
#endif
constexpr int kPipeA1Chunks = WAM_PIPE_A1_CHUNKS;  // float4 chunks per iteration of the A1 loop
constexpr int kPipeA2Unroll = WAM_PIPE_A2_UNROLL;  // pairs per iteration of the A2 loop
constexpr unsigned kPipeSpinLimit = 1u << 24;

struct PipeShared {
  float tiles[kStages][kTile * kTile];     // input staging (swizzled), A1 only
  float pf[kPipePf][kTile * 32];           // pre-filtered samples [i][lane]
  double amp[kPipeDec][16 * 32];           // amplitudes [k][lane]
  uint32_t bits[kPipeDec][32];             // hard decisions of the tile, bit k = decimated sample k
  int reset_k[32];                         // roll-back request: decimated index of the reset per lane, -1 none
  int rows[32];
  volatile int a1_done;
  volatile int b_done;
  volatile int epoch_req;
  volatile int reset_tile;
  volatile int abort_flag;
  volatile unsigned long long a2_pub;      // (epoch_ack << 32) | min next tile
};

__device__ __forceinline__ bool pipe_spin(unsigned& spins, PipeShared& sh) {
  if (sh.abort_flag) return false;
  if (++spins > kPipeSpinLimit) { sh.abort_flag = 1; return false; }
  __nanosleep(32);
  return true;
}

template <bool ALIGNED, bool THIN = false>
__global__ void __launch_bounds__(kPipeThreads, 5) fsk_demod_pipe_kernel(const __grid_constant__ DemodLaunch L) {
  int gi = 0;
#pragma unroll
  for (int i = 1; i < kMaxGroupsPerLaunch; ++i)
    if (i < L.n_groups && (int)blockIdx.x >= L.block_begin[i]) gi = i;
  const DemodArgs& a = L.g[gi];
  extern __shared__ __align__(128) unsigned char pipe_smem[];
  PipeShared& sh = *reinterpret_cast<PipeShared*>(pipe_smem);
  // B's copy of the bit-packed sync rings of the CTA's 32 streams, [ring_words][32] (when it fits): the frame
  // search reads it 4 words at a time per check, and from L2 each of those rounds costs a full memory latency
  uint32_t* ring_s = reinterpret_cast<uint32_t*>(pipe_smem + sizeof(PipeShared));
  const bool ring_in_smem = L.pipe_ring_smem != 0;

  const int lane = threadIdx.x & 31;
  const int role = threadIdx.x >> 5;
  int li = a.l_begin + ((int)blockIdx.x - L.block_begin[gi]) * 32 + lane;
  bool active = li < a.l_end;
  if (a.sel != nullptr) {  // sub-selection, see fsk_demod_exact_kernel; the whole CTA takes the same way out
    const int cnt = *a.sel_count;
    if (li - lane >= cnt) return;
    active = li < cnt;
    li = active ? a.sel[li] : a.l_begin;
  }
  int row = -1;
  if (active) row = (a.ids ? a.ids[li] : a.id0 + li) - a.row_base;
  long n_l = a.n;  // this stream's samples in this launch (ragged launches: its own count)
  if (active && a.n_valid) {
    const long v = (long)a.n_valid[row];
    if (v < 0) {  // demodulateData() is not called on this stream
      active = false;
      if (role == 2 && !a.append) a.out_len[row] = 0;
    } else {
      const long r = v - a.n_valid_offset;
      n_l = r < 0 ? 0 : (r < a.n ? r : a.n);
    }
  }
  if (role == 0) {
    sh.rows[lane] = active ? row : -1;
    sh.reset_k[lane] = -1;
    if (lane == 0) { sh.a1_done = 0; sh.b_done = 0; sh.epoch_req = 0; sh.reset_tile = 0; sh.abort_flag = 0; sh.a2_pub = 0ull; }
  }
  __syncthreads();

  const FskDerived& d = a.d;
  const long ns = a.n_local;
  const int n_tiles = (int)((a.n + kTile - 1) / kTile);
#ifdef WAM_PHASE_TIMING
  const bool timing = a.phase_cycles != nullptr;  // busy SM cycles of each role (waits excluded), per CTA
#else
  constexpr bool timing = false;
#endif
  unsigned long long busy = 0;
  const long long clk_begin = timing ? clock64() : 0;
  const int dsc0 = active ? (int)a.u32[(long)U_DSC * ns + li] : 0;  // decimator phase: the same at every tile start

  if (role == 0) {
    // =========================== A1: staging + AGC + pre-filter ===========================
    A1State a1;
    a1.gain = 1.0; a1.py1 = a1.py2 = 0.0; a1.px1 = a1.px2 = 0.0f;
    if (active) {
      const double* f = a.f64 + li;
      a1.gain = f[F_GAIN * ns]; a1.py1 = f[F_PY1 * ns]; a1.py2 = f[F_PY2 * ns];
      a1.px1 = (float)f[F_PX1 * ns]; a1.px2 = (float)f[F_PX2 * ns];
    }
    const bool agc = d.agc_enabled != 0;
    const double att = d.agc_attack, rel = d.agc_release;
    for (int p = 0; p < kStages - 1; ++p) {
      if (p < n_tiles) stage_tile<ALIGNED>(sh.tiles[p], a, sh.rows, (long)p * kTile, lane);
      cp_async_commit();
    }
    unsigned spins = 0;
    for (int t = 0; t < n_tiles; ++t) {
      const int tn = t + kStages - 1;
      if (tn < n_tiles) stage_tile<ALIGNED>(sh.tiles[tn % kStages], a, sh.rows, (long)tn * kTile, lane);
      cp_async_commit();
      // pf slot t % kPipePf is free once B has finished with tile t - kPipePf
      bool ok = true;
      while (t - sh.b_done >= kPipePf) { if (!pipe_spin(spins, sh)) { ok = false; break; } }
      if (!ok) break;
      spins = 0;
      cp_async_wait<kStages - 1>();
      __syncwarp();
      const long long c0 = timing ? clock64() : 0;
      const float* tile = sh.tiles[t % kStages];
      float* pfb = sh.pf[t % kPipePf];
      const int len = (int)max(0L, min((long)kTile, n_l - (long)t * kTile));
      if (active) {
        if (len == kTile) {
#pragma unroll 1
          for (int ch = 0; ch < 8; ch += kPipeA1Chunks) {
#pragma unroll
            for (int c = 0; c < kPipeA1Chunks; ++c) {
              const float4 v = *reinterpret_cast<const float4*>(tile + tile_index(lane, (ch + c) * 4));
              float sg;
              const float p0 = phase_a1_sample(a1, v.x, d, agc, att, rel, sg);
              const float p1 = phase_a1_sample(a1, v.y, d, agc, att, rel, sg);
              const float p2 = phase_a1_sample(a1, v.z, d, agc, att, rel, sg);
              const float p3 = phase_a1_sample(a1, v.w, d, agc, att, rel, sg);
              float* pfp = pfb + ((ch + c) * 4) * 32 + lane;
              pfp[0] = p0; pfp[32] = p1; pfp[64] = p2; pfp[96] = p3;
            }
          }
        } else {
#pragma unroll 1
          for (int i = 0; i < len; ++i) {
            float sg;
            pfb[i * 32 + lane] = phase_a1_sample(a1, tile[tile_index(lane, i)], d, agc, att, rel, sg);
          }
        }
      }
      __syncwarp();
      __threadfence_block();
      if (lane == 0) sh.a1_done = t + 1;
      if (timing) busy += (unsigned long long)(clock64() - c0);
    }
    cp_async_wait<0>();
    if (timing && lane == 0) a.phase_cycles[4ull * blockIdx.x + 0] += busy;
    if (active) {
      double* f = a.f64 + li;
      f[F_GAIN * ns] = a1.gain; f[F_PY1 * ns] = a1.py1; f[F_PY2 * ns] = a1.py2;
      f[F_PX1 * ns] = (double)a1.px1; f[F_PX2 * ns] = (double)a1.px2;
    }
  } else if (role == 1) {
    // =========================== A2: mix, I/Q filters, discriminator ===========================
    A2State s;
    s.lo_c = 1.0; s.lo_s = 0.0;
    s.ix1 = s.ix2 = s.iy1 = s.iy2 = s.qx1 = s.qx2 = s.qy1 = s.qy2 = 0.0;
    s.ox1 = s.ox2 = s.oy1 = s.oy2 = s.last_phase = s.iacc = s.qacc = 0.0;
    s.dsc = 0;
    if (active) {
      const double* f = a.f64 + li;
      s.lo_c = f[F_LO_C * ns]; s.lo_s = f[F_LO_S * ns];
      s.ix1 = f[F_IX1 * ns]; s.ix2 = f[F_IX2 * ns]; s.iy1 = f[F_IY1 * ns]; s.iy2 = f[F_IY2 * ns];
      s.qx1 = f[F_QX1 * ns]; s.qx2 = f[F_QX2 * ns]; s.qy1 = f[F_QY1 * ns]; s.qy2 = f[F_QY2 * ns];
      s.ox1 = f[F_OX1 * ns]; s.ox2 = f[F_OX2 * ns]; s.oy1 = f[F_OY1 * ns]; s.oy2 = f[F_OY2 * ns];
      s.last_phase = f[F_LAST_PHASE * ns]; s.iacc = f[F_IACC * ns]; s.qacc = f[F_QACC * ns];
    }
    constexpr int kNever = 0x7fffffff;
    int lane_pos = active ? 0 : kNever;   // next tile this lane has to compute
    int lane_kfrom = 0;                   // first pair to compute in that tile (after a roll-back)
    int epoch_seen = 0;
    unsigned spins = 0;
    bool alive = true;
    while (alive) {
      // ---- roll-back request from B?
      const int er = sh.epoch_req;
      if (er != epoch_seen) {
        __threadfence_block();
        const int rk = sh.reset_k[lane];
        if (rk >= 0) {
          // FSKCore.resetState(), DSP side (fsk.ts:175-188): the lane restarts right after decimated sample rk
          reset_state_a2(s);
          lane_pos = sh.reset_tile;
          lane_kfrom = rk + 1;
        }
        epoch_seen = er;
      }
      int cur = lane_pos;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) cur = min(cur, __shfl_xor_sync(0xffffffffu, cur, o));
      if (lane == 0) sh.a2_pub = ((unsigned long long)(unsigned)epoch_seen << 32) | (unsigned)min(cur, n_tiles);
      if (cur >= n_tiles) {
        // everything computed: stay around until B has finished (it may still ask for a roll-back)
        if (sh.b_done >= n_tiles) break;
        if (!pipe_spin(spins, sh)) break;
        continue;
      }
      // ---- wait for the pre-filtered tile and for a free dec slot
      if (cur >= sh.a1_done || cur - sh.b_done >= kPipeDec) {
        if (!pipe_spin(spins, sh)) break;
        continue;
      }
      spins = 0;
      __threadfence_block();
      const long long c0 = timing ? clock64() : 0;
      if (lane_pos == cur) {
        const float* pfbuf = sh.pf[cur % kPipePf];
        double* pbuf = sh.amp[cur % kPipeDec];
        const int len = (int)max(0L, min((long)kTile, n_l - (long)cur * kTile));
        const int v_hi = dsc0 + len;
        const int nk = v_hi >> 1;
        const int k_from = lane_kfrom;
        const int v_lo = k_from > 0 ? 2 * k_from : dsc0;
        uint32_t bits = 0u;
        if (k_from > 0) {
          bits = sh.bits[cur % kPipeDec][lane] & ((1u << k_from) - 1u);
        } else if (len > 0) {
          // renormalise the LO rotation (one Newton step towards |(c, s)| = 1)
          const double m = fma(s.lo_c, s.lo_c, s.lo_s * s.lo_s);
          const double f = fma(-0.5, m, 1.5);
          s.lo_c *= f; s.lo_s *= f;
        }
        if (dsc0 == 0 && (v_hi & 1) == 0) {
          // whole pairs only (two pairs per iteration: the biquad histories rotate in place and two atan2
          // chains overlap).  Splitting the tile into passes (all I/Q sums, then 16 independent atan2, then the
          // post-filter chain) was tried and is slower: 300 vs 222 cycles per sample (profiles/r01_notes.md).
#pragma unroll kPipeA2Unroll
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
        sh.bits[cur % kPipeDec][lane] = bits;
        lane_pos = cur + 1;
        lane_kfrom = 0;
      }
      __syncwarp();
      __threadfence_block();
      if (timing) busy += (unsigned long long)(clock64() - c0);
    }
    if (timing && lane == 0) a.phase_cycles[4ull * blockIdx.x + 1] += busy;
    if (active) {
      double* f = a.f64 + li;
      f[F_LO_C * ns] = s.lo_c; f[F_LO_S * ns] = s.lo_s;
      f[F_IX1 * ns] = s.ix1; f[F_IX2 * ns] = s.ix2; f[F_IY1 * ns] = s.iy1; f[F_IY2 * ns] = s.iy2;
      f[F_QX1 * ns] = s.qx1; f[F_QX2 * ns] = s.qx2; f[F_QY1 * ns] = s.qy1; f[F_QY2 * ns] = s.qy2;
      f[F_OX1 * ns] = s.ox1; f[F_OX2 * ns] = s.ox2; f[F_OY1 * ns] = s.oy1; f[F_OY2 * ns] = s.oy2;
      f[F_LAST_PHASE * ns] = s.last_phase; f[F_IACC * ns] = s.iacc; f[F_QACC * ns] = s.qacc;
      a.u32[(long)U_DSC * ns + li] = (uint32_t)((dsc0 + n_l) & 1);
    }
  } else {
    // =========================== B: decimated-rate state machine ===========================
    BState b;
    b.sil_thr = 0.01;
    b.gsc = b.gmod = b.bsc = b.next_idx = b.bit_acc = b.bit_cnt = b.started = b.current = b.sil_cnt = 0u;
    b.bitpos = 0; b.ring_pos = b.ring_len = b.amp_pos = b.amp_len = b.cur_word = 0u; b.out_n = 0;
    if (active) {
      const double* f = a.f64 + li;
      const uint32_t* u = a.u32 + li;
      b.sil_thr = f[F_SIL_THR * ns];
      b.gsc = u[U_GSC * ns]; b.gmod = u[U_GMOD * ns]; b.bsc = u[U_BSC * ns]; b.next_idx = u[U_NEXT_IDX * ns];
      b.bit_acc = u[U_BIT_ACC * ns]; b.bit_cnt = u[U_BIT_CNT * ns]; b.started = u[U_STARTED * ns];
      b.bitpos = (int)u[U_BITPOS * ns]; b.current = u[U_CURRENT * ns]; b.sil_cnt = u[U_SIL_CNT * ns];
      b.ring_pos = u[U_RING_POS * ns]; b.ring_len = u[U_RING_LEN * ns];
      b.amp_pos = u[U_AMP_POS * ns]; b.amp_len = u[U_AMP_LEN * ns];
      b.out_n = a.append ? a.out_len[row] : 0;
      if ((b.ring_pos & 31u) != 0u) {
        const uint32_t w = ring_of(a, li)[(b.ring_pos >> 5) & (uint32_t)(d.ring_words - 1)];
        b.cur_word = w & ((1u << (b.ring_pos & 31u)) - 1u);
      }
      if (ring_in_smem)
        for (int w = 0; w < d.ring_words; ++w) ring_s[w * 32 + lane] = ring_of(a, li)[w];
    }
    __syncwarp();
    uint32_t* ring = ring_in_smem ? ring_s + lane : ring_of(a, li);
    const long rstride = ring_in_smem ? 32 : 1;
    uint8_t* out_row = active ? a.out + (long)row * a.out_stride : nullptr;
    int epoch = 0;
    unsigned spins = 0;
    bool alive = true;
    for (int t = 0; t < n_tiles && alive; ++t) {
      const int len = (int)max(0L, min((long)kTile, n_l - (long)t * kTile));
      const int v_hi = dsc0 + len;
      const int nk = v_hi >> 1;
      const uint32_t pos_t0 = b.ring_pos, len_t0 = b.ring_len, slot_t0 = b.amp_pos, alen_t0 = b.amp_len;
      int b_from = 0;
      bool lane_busy = active && len > 0;
      for (;;) {
        // tile t published by A2 for the current epoch?
        for (;;) {
          const unsigned long long pub = sh.a2_pub;
          if ((int)(pub >> 32) == epoch && (int)(pub & 0xffffffffu) > t) break;
          if (!pipe_spin(spins, sh)) { alive = false; break; }
        }
        if (!alive) break;
        spins = 0;
        __threadfence_block();
        const long long c0 = timing ? clock64() : 0;
        int k_reset = -1;
        if (lane_busy) {
          const uint32_t bits = sh.bits[t % kPipeDec][lane];
          k_reset = sm_tile_events<true, THIN>(b, bits, sh.amp[t % kPipeDec] + lane, b_from, nk, pos_t0, len_t0, slot_t0,
                                         alen_t0, a, li, out_row, ring, rstride);
          if (k_reset < 0 || 2 * (k_reset + 1) >= v_hi) {
            // this lane is done with the tile: end-of-tile ring bookkeeping
            b.ring_len = min(len_t0 + (uint32_t)nk, (uint32_t)d.ring_cap_int);
            const uint32_t sl = slot_t0 + (uint32_t)nk;
            b.amp_pos = sl >= (uint32_t)d.amp_phys ? sl - (uint32_t)d.amp_phys : sl;
            b.amp_len = min(alen_t0 + (uint32_t)nk, (uint32_t)d.amp_cap);
            lane_busy = false;
          } else {
            b_from = k_reset + 1;
          }
        }
        const unsigned any_reset = __ballot_sync(0xffffffffu, k_reset >= 0);
        if (any_reset) {
          // resetState() ran in some lanes: A2 must restart them from zero right after k_reset
          sh.reset_k[lane] = k_reset;
          __syncwarp();
          if (lane == 0) sh.reset_tile = t;
          __threadfence_block();
          ++epoch;
          if (lane == 0) sh.epoch_req = epoch;
          __syncwarp();
        }
        if (timing) busy += (unsigned long long)(clock64() - c0);
        if (!__any_sync(0xffffffffu, lane_busy)) {
          if (any_reset) {
            // wait for the acknowledgement before moving on, so that the request slots can be reused
            for (;;) {
              if ((int)(sh.a2_pub >> 32) == epoch) break;
              if (!pipe_spin(spins, sh)) { alive = false; break; }
            }
          }
          break;
        }
      }
      if (!alive) break;
      __syncwarp();
      __threadfence_block();
      if (lane == 0) sh.b_done = t + 1;
    }
    if (active) {
      double* f = a.f64 + li;
      uint32_t* u = a.u32 + li;
      f[F_SIL_THR * ns] = b.sil_thr;
      u[U_GSC * ns] = b.gsc; u[U_GMOD * ns] = b.gmod; u[U_BSC * ns] = b.bsc; u[U_NEXT_IDX * ns] = b.next_idx;
      u[U_BIT_ACC * ns] = b.bit_acc; u[U_BIT_CNT * ns] = b.bit_cnt; u[U_STARTED * ns] = b.started;
      u[U_BITPOS * ns] = (uint32_t)b.bitpos; u[U_CURRENT * ns] = b.current; u[U_SIL_CNT * ns] = b.sil_cnt;
      u[U_RING_POS * ns] = b.ring_pos; u[U_RING_LEN * ns] = b.ring_len;
      u[U_AMP_POS * ns] = b.amp_pos; u[U_AMP_LEN * ns] = b.amp_len;
      if (ring_in_smem)
        for (int w = 0; w < d.ring_words; ++w) ring_of(a, li)[w] = ring_s[w * 32 + lane];
      if ((b.ring_pos & 31u) != 0u)
        ring_of(a, li)[(b.ring_pos >> 5) & (uint32_t)(d.ring_words - 1)] = b.cur_word;
      a.out_len[row] = b.out_n < a.out_stride ? b.out_n : (int)a.out_stride;
      if (a.n_valid) ragged_account(a, li, n_l);
      if (sh.abort_flag) u[(long)U_ERR * ns] |= WAM_ERR_PIPE_TIMEOUT;
    }
    if (timing && lane == 0) {
      a.phase_cycles[4ull * blockIdx.x + 2] += busy;
      a.phase_cycles[4ull * blockIdx.x + 3] += (unsigned long long)(clock64() - clk_begin);  // B's whole life
    }
  }
}

}  // namespace wam
