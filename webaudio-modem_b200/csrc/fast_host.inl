// fast_host.inl — host side of the mixed-precision fast path (included by wam_api.cu).
//
// One fast call over a configuration group:
//   1. prologue   the rings' newest contents become the prefix of the call's linear histories (hard bits, amplitudes);
//   2. fast pass  fsk_demod_fast_kernel, one launch per time slab: reads checkpoint j of the per-stream state, writes
//                 checkpoint j + 1, appends bytes, writes the histories, and lists the streams in which a decision of
//                 the slab was doubtful (float32 error band, see fsk_demod_fast.cuh);
//   3. check      every listed (stream, slab) becomes a WINDOW of 2 slabs (more when the doubtful decision is a frame
//                 sync, whose template looks further back): state of the checkpoint at the window's start, rings
//                 rebuilt from the histories, the window's samples — demodulated by the float64 kernel in a scratch
//                 batch.  The first slab of a window re-converges the filters (the float32 state it starts from is
//                 1e-7 off; the filters forget that within a few hundred samples), the last one decides the doubtful
//                 decision in float64.  State machine and bytes at the window's end equal to the fast pass'
//                 checkpoint: the fast results stand (the usual outcome: a doubtful sample is wrong 1 time in 300).
//                 Otherwise, or when a window cannot be formed, the stream joins the hard list;
//                 A doubtful END-OF-DATA decision (an amplitude inside the doubt band of the silence threshold) is
//                 checked in two stages, because the threshold itself is a float32 mean in the fast pass: a window
//                 around the sync detection that set the threshold yields the float64 threshold, a window around the
//                 doubtful decision then runs with that threshold and refuses compares closer than 1e-9 to it;
//   4. hard list  those streams run through the float64 kernel over the whole call from the live state (checkpoint
//                 0, untouched so far) and the rings (untouched too);
//   5. epilogue   every other stream's last checkpoint becomes its live state, the histories' tails its rings.
// So a doubtful decision costs the float64 kernel a few thousand samples instead of the whole call, and what leaves the
// library is what the float64 kernels alone would have produced.

struct FastGeom {
  int ns;            // streams of the group
  long n;            // samples per stream in this call
  long slab_len;     // samples per time slab (a multiple of kTile); only the last slab may be shorter
  int n_slabs;
  int ph;            // prefix of the bit history, half words (= 2 ring_words)
  int pa;            // prefix of the amplitude history, entries (amp_cap rounded up to 16)
  long bh_stride, ah_stride;
  int sync_slabs;    // window length, in slabs, that checks a doubtful sync decision
  int e1_len, e2_len;  // two-stage check of a doubtful end-of-data decision: window lengths in slabs (0: not available)
  long out_stride;
  long sv_stride;    // samples per row of the verification scratch (kVerifyClasses slabs)
};

struct FastCtx {  // what the small kernels below need, by value
  FastGeom q;
  int ring_words, amp_phys, amp_cap;
  double* f64; uint32_t* u32;              // live state (= checkpoint 0)
  double* ck_f64; uint32_t* ck_u32;        // checkpoints 1..S
  uint32_t* sync_ring; float* amp_ring;    // live rings
  uint16_t* bit_hist; float* amp_hist;
  int32_t* slab_list; int32_t* slab_count;
  int32_t* hard_list; int32_t* hard_count; uint32_t* hard_mark;
  int32_t* item_li; int32_t* item_slab; int32_t* item_count; int32_t* item_res;
  double* sv_f64; uint32_t* sv_u32; uint32_t* sv_ring; float* sv_amp; float* sv_samples; uint8_t* sv_out; int32_t* sv_out_len;
  const float* samples; long stride;       // the call's sample buffer
  const int32_t* ids; int id0, row_base;
  uint8_t* out; int32_t* out_len;
  int append;
};

__device__ __forceinline__ long fc_row(const FastCtx& c, int li) { return (long)(c.ids ? c.ids[li] : c.id0 + li) - c.row_base; }
__device__ __forceinline__ const double* fc_ck_f64(const FastCtx& c, int k) {  // checkpoint k (0 = live)
  return k == 0 ? c.f64 : c.ck_f64 + (size_t)(k - 1) * F64_COUNT * c.q.ns;
}
__device__ __forceinline__ const uint32_t* fc_ck_u32(const FastCtx& c, int k) {
  return k == 0 ? c.u32 : c.ck_u32 + (size_t)(k - 1) * U32_COUNT * c.q.ns;
}
__device__ __forceinline__ void fc_hard(const FastCtx& c, int li) {  // onto the hard list, once
  if (atomicExch(c.hard_mark + li, 1u) == 0u) c.hard_list[atomicAdd(c.hard_count, 1)] = li;
}

// Window of an item of class `cls` (v = its item_slab entry): first slab, length in slabs, checkpoint behind it.
// Classes 0..kVerifyClasses-1: cls + 1 slabs ending with slab v.  kStageE1: e1_len slabs ending with slab v (the sync
// detection that set the silence threshold; v = -1: none in this call, the entry is a placeholder), pushed forward
// when it would start before the call.  kStageE2: e2_len slabs ending with slab v, likewise.
__device__ __forceinline__ void fc_window(const FastCtx& c, int cls, int v, int& j0, int& wl, int& end) {
  if (cls < kVerifyClasses) { j0 = v - cls; wl = cls + 1; end = v + 1; }
  else if (cls == kStageE1) { wl = c.q.e1_len; j0 = max(v - wl + 1, 0); end = j0 + wl; }
  else { wl = c.q.e2_len; end = max(v + 1, wl); j0 = end - wl; }
}
__device__ __forceinline__ int fc_count(const FastCtx& c, int cls) {  // both stages share one item count
  return min(c.item_count[cls == kStageE2 ? kStageE1 : cls], kVerifyCap);
}

// (1) prologue: a block per stream
__global__ void fast_prologue_kernel(const FastCtx c) {
  const int li = blockIdx.x;
  const int ns = c.q.ns;
  const uint32_t ring_pos = c.u32[(size_t)U_RING_POS * ns + li];
  const uint32_t amp_pos = c.u32[(size_t)U_AMP_POS * ns + li];
  const uint16_t* ring16 = reinterpret_cast<const uint16_t*>(c.sync_ring + (size_t)li * c.ring_words);
  uint16_t* bh = c.bit_hist + (size_t)li * c.q.bh_stride;
  const uint32_t hmask = (uint32_t)(2 * c.ring_words - 1);
  const uint32_t p16 = ring_pos >> 4;  // aligned calls: ring_pos is a multiple of 16
  for (int m = threadIdx.x; m < c.q.ph; m += blockDim.x) bh[c.q.ph - 1 - m] = ring16[(p16 - 1u - (uint32_t)m) & hmask];
  const float* ar = c.amp_ring + (size_t)li * c.amp_phys;
  float* ah = c.amp_hist + (size_t)li * c.q.ah_stride;
  for (int k = threadIdx.x; k < c.q.pa; k += blockDim.x) {
    float v = 0.0f;
    if (k < c.amp_cap) {
      int slot = (int)amp_pos - 1 - k;
      if (slot < 0) slot += c.amp_phys;
      v = ar[slot];
    }
    ah[c.q.pa - 1 - k] = v;
  }
  if (threadIdx.x == 0) {
    c.u32[(size_t)U_OUT_N * ns + li] = c.append ? (uint32_t)c.out_len[fc_row(c, li)] : 0u;
    c.u32[(size_t)U_FLAG * ns + li] = 0u;
    c.hard_mark[li] = 0u;
  }
}

// (3a) collect: the slab lists become windows, sorted into classes by their length in slabs
__global__ void fast_collect_kernel(const FastCtx c) {
  const int ns = c.q.ns;
  for (int j = blockIdx.y; j < c.q.n_slabs; j += gridDim.y) {
    const int cnt = min(c.slab_count[j], ns);
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < cnt; e += gridDim.x * blockDim.x) {
      const int v = c.slab_list[(size_t)j * ns + e];
      const int li = v & 0xffffff;
      const uint32_t cause = (uint32_t)v >> 24;
      bool hard = (cause & WAM_FLAG_RANGE) != 0;
      // a short last slab has no window class of its own
      if (j == c.q.n_slabs - 1 && c.q.n != (long)c.q.n_slabs * c.q.slab_len) hard = true;
      const int want = (cause & WAM_FLAG_SYNC) ? c.q.sync_slabs : 2;
      if (want > kVerifyClasses) hard = true;
      if (!hard && (cause & WAM_FLAG_EOD)) {
        // two stages: the threshold's origin (last slab before the window in which a sync was detected), then the window
        int j0, wl, e2;
        fc_window(c, kStageE2, j, j0, wl, e2);
        bool ok = c.q.e2_len >= want;
        const uint32_t sd0 = fc_ck_u32(c, j0)[(size_t)U_SYNC_DET * ns + li];
        ok = ok && fc_ck_u32(c, e2)[(size_t)U_SYNC_DET * ns + li] == sd0;  // no new threshold inside the window
        int js = -1;
        if (ok) {
          uint32_t later = sd0;
          for (int k = j0 - 1; k >= 0; --k) {
            const uint32_t sd = fc_ck_u32(c, k)[(size_t)U_SYNC_DET * ns + li];
            if (sd != later) { js = k; break; }
            later = sd;
          }
          if (js >= 0) {
            int s1, wl1, e1;
            fc_window(c, kStageE1, js, s1, wl1, e1);
            // the stage-1 window must end under the threshold that is in force in the stage-2 window
            ok = fc_ck_u32(c, e1)[(size_t)U_SYNC_DET * ns + li] == sd0;
          }
        }
        if (ok) {
          const int slot = atomicAdd(c.item_count + kStageE1, 1);
          if (slot < kVerifyCap) {
            c.item_li[kStageE1 * kVerifyCap + slot] = li; c.item_slab[kStageE1 * kVerifyCap + slot] = js;
            c.item_li[kStageE2 * kVerifyCap + slot] = li; c.item_slab[kStageE2 * kVerifyCap + slot] = j;
          } else {
            ok = false;
            atomicAdd(c.hard_count + 3, 1);
          }
        }
        if (!ok) hard = true;
      } else if (!hard) {
        const int len = min(want, j + 1);  // windows are cut at the start of the call (checkpoint 0 is the exact state)
        const int slot = atomicAdd(c.item_count + (len - 1), 1);
        if (slot < kVerifyCap) {
          c.item_li[(len - 1) * kVerifyCap + slot] = li;
          c.item_slab[(len - 1) * kVerifyCap + slot] = j;
        } else {
          hard = true;
          atomicAdd(c.hard_count + 3, 1);
        }
      }
      {
        // A window cut at the start of the call trusts checkpoint 0.  Doubt carried over from the previous call (a
        // running vote, a silent run or ring bits read in float32 whose samples are gone) cannot be checked any more:
        // the stream is re-run in float64 from checkpoint 0 and the condition is put on record.
        const int need = (cause & WAM_FLAG_EOD) ? c.q.e2_len : want;
        if (j + 1 < need) {
          const uint32_t* u0 = c.u32;
          bool carried = false;
          if (cause & (WAM_FLAG_VOTE_START | WAM_FLAG_VOTE_DATA | WAM_FLAG_VOTE_STOP)) carried = carried || u0[(size_t)U_DVOTE * ns + li] != 0u;
          if (cause & WAM_FLAG_SYNC) carried = carried || u0[(size_t)U_DCNT * ns + li] != 0u;
          if (cause & WAM_FLAG_EOD) carried = carried || (u0[(size_t)U_SILX * ns + li] >> 31) != 0u;
          if (carried) { hard = true; atomicOr(c.u32 + (size_t)U_ERR * ns + li, WAM_ERR_CARRIED_DOUBT); }
        }
      }
      if (hard) atomicOr(c.u32 + (size_t)U_FLAG_EVER * ns + li, cause);  // the statistics see the hard streams' causes too
      if (hard) fc_hard(c, li);
    }
  }
}

// Scratch entry `si` of a scratch batch of `nv` streams <- stream li as of checkpoint j0: state (doubt tracking cleared),
// rings rebuilt from the histories, and wslabs slabs of samples.  One block.
struct FastScratch {
  double* f64; uint32_t* u32; uint32_t* ring; float* amp; float* samples; long stride; int32_t* out_len;
};
__device__ __forceinline__ void fc_gather_body(const FastCtx& c, const FastScratch& d, int li, int j0, int wslabs, size_t si, size_t nv) {
  const int ns = c.q.ns;
  const double* f = fc_ck_f64(c, j0);
  const uint32_t* u = fc_ck_u32(c, j0);
  for (int k = threadIdx.x; k < F64_COUNT; k += blockDim.x) {
    double v = f[(size_t)k * ns + li];
    if (k == F_FAST_E || k == F_FAST_RSP) v = 0.0;
    d.f64[(size_t)k * nv + si] = v;
  }
  for (int k = threadIdx.x; k < U32_COUNT; k += blockDim.x) {
    uint32_t v = u[(size_t)k * ns + li];
    if (k == U_DVOTE || k == U_SILX || k == U_LAST_DOUBT || k == U_DCNT || k == U_FLAG || k == U_ERR) v = 0u;
    d.u32[(size_t)k * nv + si] = v;
  }
  // rings as of the window's start, from the histories
  const uint32_t ring_pos = u[(size_t)U_RING_POS * ns + li];
  const uint32_t amp_pos = u[(size_t)U_AMP_POS * ns + li];
  const long tiles_per_slab = c.q.slab_len / kTile;
  const uint16_t* bh = c.bit_hist + (size_t)li * c.q.bh_stride + c.q.ph + (long)j0 * tiles_per_slab;  // behind the newest tile
  uint16_t* ring16 = reinterpret_cast<uint16_t*>(d.ring + si * c.ring_words);
  const uint32_t hmask = (uint32_t)(2 * c.ring_words - 1);
  const uint32_t p16 = ring_pos >> 4;
  for (int m = threadIdx.x; m < 2 * c.ring_words; m += blockDim.x) ring16[(p16 - 1u - (uint32_t)m) & hmask] = bh[-1 - m];
  const float* ah = c.amp_hist + (size_t)li * c.q.ah_stride + c.q.pa + (long)j0 * (c.q.slab_len / 2);
  float* ar = d.amp + si * c.amp_phys;
  for (int k = threadIdx.x; k < c.amp_cap; k += blockDim.x) {
    int slot = (int)amp_pos - 1 - k;
    if (slot < 0) slot += c.amp_phys;
    ar[slot] = ah[-1 - k];
  }
  // the window's samples
  const long wlen = (long)wslabs * c.q.slab_len;
  const float4* src = reinterpret_cast<const float4*>(c.samples + fc_row(c, li) * c.stride + (long)j0 * c.q.slab_len);
  float4* dst = reinterpret_cast<float4*>(d.samples + si * d.stride);
  for (long k = threadIdx.x; k < wlen / 4; k += blockDim.x) dst[k] = src[k];
  if (threadIdx.x == 0) d.out_len[si] = 0;
}

// (3b) gather: a block per window of class `cls`.  Scratch index of window i of class cls: cls * kVerifyCap + i.
__global__ void fast_gather_kernel(const FastCtx c, int cls) {
  const int i = blockIdx.x;
  if (i >= fc_count(c, cls)) return;
  const int li = c.item_li[cls * kVerifyCap + i];
  int j0, wslabs, wend;  // first slab of the window = the checkpoint it starts from
  fc_window(c, cls, c.item_slab[cls * kVerifyCap + i], j0, wslabs, wend);
  const size_t si = (size_t)cls * kVerifyCap + i;
  const size_t nv = (size_t)kAllClasses * kVerifyCap;  // streams of the scratch batch
  const FastScratch d = {c.sv_f64, c.sv_u32, c.sv_ring, c.sv_amp, c.sv_samples, c.q.sv_stride, c.sv_out_len};
  fc_gather_body(c, d, li, j0, wslabs, si, nv);
  if (cls == kStageE2) {
    // the float64 silence threshold from stage 1 (its run and its check are ahead of this kernel on the same stream)
    __syncthreads();
    if (threadIdx.x == 0 && c.item_slab[kStageE1 * kVerifyCap + i] >= 0)
      c.sv_f64[(size_t)F_SIL_THR * nv + si] = c.sv_f64[(size_t)F_SIL_THR * nv + (size_t)kStageE1 * kVerifyCap + i];
  }
}

// ---- carry: what a call leaves undecided is re-read in float64 at the start of the next call ----
// A stream that ends a fast call with float32 readings still open — a running vote with a doubtful sample, a silent
// run with a doubtful compare, ring bits inside the doubt band — cannot have them checked later: the samples belong to
// the caller and are gone.  So the call's last slabs of every such stream are kept (state of the checkpoint at their
// start, rings, samples), and the next call — fast or not — begins by running them through the float64 kernel and
// comparing the result with the live state.  Equal (the usual case): the open readings were right, the doubt records are
// cleared.  Different: the float64 result replaces the live state.  Either way the call starts from certified readings.
struct CarryCtx {
  int ns, cap, ring_words, amp_phys, amp_cap;
  double* f64; uint32_t* u32; uint32_t* sync_ring; float* amp_ring;   // live
  const int32_t* li; int32_t* count;
  const uint32_t* hard_mark;  // of the call that kept the slabs: those streams ended it in the float64 kernel's own state
  double* cv_f64; uint32_t* cv_u32; uint32_t* cv_ring; float* cv_amp;
};

// (6a) which streams carry: a thread per stream, final checkpoint
__global__ void fast_carry_collect_kernel(const FastCtx c, int32_t* carry_li, int32_t* carry_count, int cap) {
  const int li = blockIdx.x * blockDim.x + threadIdx.x;
  const int ns = c.q.ns;
  if (li >= ns) return;  // (runs beside the checks: streams that end up on the hard list are skipped when settling)
  const uint32_t* u = fc_ck_u32(c, c.q.n_slabs);
  const bool open = u[(size_t)U_DVOTE * ns + li] != 0u || (u[(size_t)U_SILX * ns + li] >> 31) != 0u || u[(size_t)U_DCNT * ns + li] != 0u;
  if (!open) return;
  const int slot = atomicAdd(carry_count, 1);
  if (slot < cap) carry_li[slot] = li;
}
// (6b) keep their last `wslabs` slabs: a block per entry
__global__ void fast_carry_gather_kernel(const FastCtx c, const int32_t* carry_li, const int32_t* carry_count, int cap, int wslabs,
                                         FastScratch d) {
  const int i = blockIdx.x;
  if (i >= min(carry_count[0], cap)) return;
  fc_gather_body(c, d, carry_li[i], c.q.n_slabs - wslabs, wslabs, (size_t)i, (size_t)cap);
}
// (6c) after the float64 run of the kept slabs: a warp per entry compares with the live state and, where they differ,
// puts the float64 result in its place
__global__ void fast_carry_apply_kernel(const CarryCtx c) {
  const int i = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  const int lane = threadIdx.x & 31;
  if (i >= min(c.count[0], c.cap)) return;
  const int li = c.li[i];
  if (c.hard_mark[li]) return;
  const size_t ns = (size_t)c.ns, nv = (size_t)c.cap;
  const int fields[] = {U_GSC, U_BSC, U_NEXT_IDX, U_BIT_ACC, U_BIT_CNT, U_STARTED, U_BITPOS, U_CURRENT, U_SIL_CNT,
                        U_SYNC_DET, U_EOD_EV, U_RING_POS, U_RING_LEN, U_AMP_POS, U_AMP_LEN, U_DSC};
  bool same = true;
  if (lane == 0) {
    for (int k : fields) same = same && c.cv_u32[(size_t)k * nv + i] == c.u32[(size_t)k * ns + li];
    if (!c.u32[(size_t)U_STARTED * ns + li]) same = same && c.cv_u32[(size_t)U_GMOD * nv + i] == c.u32[(size_t)U_GMOD * ns + li];
    const double ta = c.cv_f64[(size_t)F_SIL_THR * nv + i], tb = c.f64[(size_t)F_SIL_THR * ns + li];
    same = same && fabs(ta - tb) <= 1e-5 * fabs(tb);
    same = same && c.cv_u32[(size_t)U_ERR * nv + i] == 0u;
  }
  // ring contents, half word by half word (aligned calls put 16 bits at a time)
  const uint16_t* ra = reinterpret_cast<const uint16_t*>(c.cv_ring + (size_t)i * c.ring_words);
  const uint16_t* rb = reinterpret_cast<const uint16_t*>(c.sync_ring + (size_t)li * c.ring_words);
  const uint32_t hmask = (uint32_t)(2 * c.ring_words - 1);
  const uint32_t p16 = c.u32[(size_t)U_RING_POS * ns + li] >> 4;
  const uint32_t halves = min(c.u32[(size_t)U_RING_LEN * ns + li] >> 4, (uint32_t)(2 * c.ring_words - 2));
  for (uint32_t m = lane; m < halves; m += 32) same = same && ra[(p16 - 1u - m) & hmask] == rb[(p16 - 1u - m) & hmask];
  same = __all_sync(0xffffffffu, same);
  if (!same) {
    // the float64 reading replaces the live state (the fast-path fields keep their values: they are bounds, not readings)
    for (int k = lane; k < F64_COUNT; k += 32)
      if (k != F_FAST_S && k != F_FAST_E && k != F_FAST_RSP) c.f64[(size_t)k * ns + li] = c.cv_f64[(size_t)k * nv + i];
    for (int k = lane; k < U32_COUNT; k += 32)
      if (k != U_ERR && k != U_FLAG_EVER && k != U_DOUBT_SAMPLES && k != U_OUT_N && k != U_FLAG && k != U_SB_ONES && k != U_SB_VALID &&
          k != U_SB0 && k != U_SB1 && k != U_SB2 && k != U_SB3)
        c.u32[(size_t)k * ns + li] = c.cv_u32[(size_t)k * nv + i];
    uint32_t* dr = c.sync_ring + (size_t)li * c.ring_words;
    const uint32_t* sr = c.cv_ring + (size_t)i * c.ring_words;
    for (int k = lane; k < c.ring_words; k += 32) dr[k] = sr[k];
    float* da = c.amp_ring + (size_t)li * c.amp_phys;
    const float* sa = c.cv_amp + (size_t)i * c.amp_phys;
    for (int k = lane; k < c.amp_phys; k += 32) da[k] = sa[k];
    if (lane == 0) c.u32[(size_t)U_SB_VALID * ns + li] = 0xffffffffu;  // the search prefilter starts over
  }
  if (lane == 0) {
    // certified either way: nothing is open any more
    for (int k : {(int)U_DVOTE, (int)U_SILX, (int)U_LAST_DOUBT, (int)U_DCNT}) c.u32[(size_t)k * ns + li] = 0u;
    atomicAdd(c.count + 1, 1);
    if (!same) atomicAdd(c.count + 2, 1);
  }
}

// (3c) compare: a warp per window.  The float64 run of the window must end in the state machine state of the fast
// pass' checkpoint behind the window's last slab, having produced the same bytes.
__global__ void fast_compare_kernel(const FastCtx c, int cls) {
  const int i = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  const int lane = threadIdx.x & 31;
  if (i >= fc_count(c, cls)) return;
  const int ns = c.q.ns;
  const int li = c.item_li[cls * kVerifyCap + i];
  if (cls == kStageE1 && c.item_slab[cls * kVerifyCap + i] < 0) return;  // placeholder: the threshold predates the call
  int j0, wslabs, wend;
  fc_window(c, cls, c.item_slab[cls * kVerifyCap + i], j0, wslabs, wend);
  const size_t si = (size_t)cls * kVerifyCap + i;
  const size_t nv = (size_t)kAllClasses * kVerifyCap;
  const uint32_t* ue = fc_ck_u32(c, wend);  // fast pass, behind the window
  const uint32_t* us = fc_ck_u32(c, j0);    // fast pass, at the window's start
  const double* fe = fc_ck_f64(c, wend);
  bool same = true;
  int why = 0;  // debug record: bit k = field k of the list differs, 16 threshold, 17 error flag, 18 byte count, 19 bytes
  if (lane == 0) {
    const int fields[] = {U_GSC, U_BSC, U_NEXT_IDX, U_BIT_ACC, U_BIT_CNT, U_STARTED, U_BITPOS, U_CURRENT, U_SIL_CNT,
                          U_SYNC_DET, U_EOD_EV, U_RING_POS, U_RING_LEN, U_AMP_POS, U_AMP_LEN, U_DSC};
    int fi = 0;
    for (int k : fields) { if (c.sv_u32[(size_t)k * nv + si] != ue[(size_t)k * ns + li]) why |= 1 << fi; ++fi; }
    if (!ue[(size_t)U_STARTED * ns + li] && c.sv_u32[(size_t)U_GMOD * nv + si] != ue[(size_t)U_GMOD * ns + li])
      why |= 1 << 20;  // (the check phase only exists while searching)
    // the silence threshold is a float32 mean in one run and a float64 mean in the other
    const double ta = c.sv_f64[(size_t)F_SIL_THR * nv + si], tb = fe[(size_t)F_SIL_THR * ns + li];
    if (!(fabs(ta - tb) <= 1e-5 * fabs(tb))) why |= 1 << 16;
    if (c.sv_u32[(size_t)U_ERR * nv + si] != 0u) why |= 1 << 17;
  }
  const int o0 = (int)us[(size_t)U_OUT_N * ns + li], o1 = (int)ue[(size_t)U_OUT_N * ns + li];
  const int nb = c.sv_out_len[si];
  if (lane == 0 && nb != o1 - o0) why |= 1 << 18;
  why = __shfl_sync(0xffffffffu, why, 0);
  same = why == 0;
  if (same) {
    const uint8_t* pa = c.sv_out + si * c.q.out_stride;
    const uint8_t* pb = c.out + fc_row(c, li) * c.q.out_stride + o0;
    bool eq = true;
    for (int k = lane; k < nb && o0 + k < c.q.out_stride; k += 32) eq = eq && pa[k] == pb[k];
    same = __all_sync(0xffffffffu, eq);
    if (!same) why |= 1 << 19;
  }
  // ---- local repair.  When the float64 run differs from the fast pass ONLY in what a byte holds — bytes of the window,
  // or the byte under construction at its end (a data bit voted the other way) — while every counter, position and
  // timing field agrees, the two trajectories are the same from here on: the data bits of a byte feed nothing but
  // the byte.  Then the float64 bytes replace the fast ones in place, and a byte still under construction is
  // corrected where it lands (or in the last checkpoint, if the call ends first).  No re-run.
  const bool cls_plain = cls < kVerifyClasses;
  const int only_bytes = 1 << 19, only_current = 1 << 7;
  if (!same && cls_plain && (why & ~(only_bytes | only_current)) == 0 && o1 <= c.q.out_stride) {
    // (bytes are only compared when the state agreed; with a differing byte under construction they may differ too)
    const uint8_t* pa = c.sv_out + si * c.q.out_stride;
    uint8_t* pb = c.out + fc_row(c, li) * c.q.out_stride + o0;
    for (int k = lane; k < nb; k += 32) pb[k] = pa[k];
    if (lane == 0 && (why & only_current)) {
      // the data bits decided so far (bit positions 1 .. bitpos - 1 of the frame sit in bits 7 .. 9 - bitpos of the byte)
      // take the float64 run's values; SET, not toggled: another window of the same stream may repair the same byte
      const int bp = (int)ue[(size_t)U_BITPOS * ns + li];
      const uint32_t mask = (bp >= 1 && bp <= 9) ? (0xffu & ~((1u << (9 - bp)) - 1u)) : 0u;
      const uint32_t good = c.sv_u32[(size_t)U_CURRENT * nv + si] & mask;
      const uint32_t sd = ue[(size_t)U_SYNC_DET * ns + li];
      // follow the byte: later checkpoints carry it until it is written (byte count grows) or its frame ends
      for (int k = wend; k <= c.q.n_slabs && mask != 0u; ++k) {
        uint32_t* uk = (k == 0 ? c.u32 : c.ck_u32 + (size_t)(k - 1) * U32_COUNT * ns);
        if ((int)uk[(size_t)U_OUT_N * ns + li] > o1) {  // written at index o1 (no other frame can have synced meanwhile)
          if (o1 < c.q.out_stride) {
            uint8_t* q8 = c.out + fc_row(c, li) * c.q.out_stride + o1;
            *q8 = (uint8_t)((*q8 & ~mask) | good);
          }
          break;
        }
        if (!uk[(size_t)U_STARTED * ns + li] || uk[(size_t)U_SYNC_DET * ns + li] != sd) break;  // the frame ended without it
        uint32_t* cur = uk + (size_t)U_CURRENT * ns + li;  // still under construction at this checkpoint
        *cur = (*cur & ~mask) | good;
      }
    }
    same = true;
    why |= 1 << 29;  // (debug record: repaired in place)
  }
  if (lane == 0) c.item_res[si] = (same && !(why & (1 << 29))) ? 1 : (why | (1 << 30));
  if (lane == 0) {
    atomicAdd(c.hard_count + (same ? 1 : 2), 1);
    if (!same) fc_hard(c, li);
  }
}

// (4) hard list: the doubt tracking of these streams restarts clean behind the float64 run, and their byte count goes
// back to the start of the call.
__global__ void fast_hard_prepare_kernel(const FastCtx c) {
  const int ns = c.q.ns;
  const int cnt = min(c.hard_count[0], ns);
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < cnt; e += gridDim.x * blockDim.x) {
    const int li = c.hard_list[e];
    c.f64[(size_t)F_FAST_S * ns + li] = 16.0;
    c.f64[(size_t)F_FAST_E * ns + li] = 0.0;
    c.f64[(size_t)F_FAST_RSP * ns + li] = 0.0;
    for (int k : {(int)U_DVOTE, (int)U_SILX, (int)U_LAST_DOUBT, (int)U_DCNT}) c.u32[(size_t)k * ns + li] = 0u;
    c.u32[(size_t)U_SB_VALID * ns + li] = 0xffffffffu;
    c.out_len[fc_row(c, li)] = (int)c.u32[(size_t)U_OUT_N * ns + li];
  }
}

// (5) epilogue: a block per stream that is not on the hard list
__global__ void fast_epilogue_kernel(const FastCtx c) {
  const int li = blockIdx.x;
  const int ns = c.q.ns;
  if (c.hard_mark[li]) return;
  const double* f = fc_ck_f64(c, c.q.n_slabs);
  const uint32_t* u = fc_ck_u32(c, c.q.n_slabs);
  for (int k = threadIdx.x; k < F64_COUNT; k += blockDim.x) c.f64[(size_t)k * ns + li] = f[(size_t)k * ns + li];
  for (int k = threadIdx.x; k < U32_COUNT; k += blockDim.x) c.u32[(size_t)k * ns + li] = u[(size_t)k * ns + li];
  const uint32_t ring_pos = u[(size_t)U_RING_POS * ns + li];
  const uint32_t amp_pos = u[(size_t)U_AMP_POS * ns + li];
  const long n_tiles = c.q.n / kTile;
  const uint16_t* bh = c.bit_hist + (size_t)li * c.q.bh_stride + c.q.ph + n_tiles;
  uint16_t* ring16 = reinterpret_cast<uint16_t*>(c.sync_ring + (size_t)li * c.ring_words);
  const uint32_t hmask = (uint32_t)(2 * c.ring_words - 1);
  const uint32_t p16 = ring_pos >> 4;
  for (int m = threadIdx.x; m < 2 * c.ring_words; m += blockDim.x) ring16[(p16 - 1u - (uint32_t)m) & hmask] = bh[-1 - m];
  const float* ah = c.amp_hist + (size_t)li * c.q.ah_stride + c.q.pa + c.q.n / 2;
  float* ar = c.amp_ring + (size_t)li * c.amp_phys;
  for (int k = threadIdx.x; k < c.amp_cap; k += blockDim.x) {
    int slot = (int)amp_pos - 1 - k;
    if (slot < 0) slot += c.amp_phys;
    ar[slot] = ah[-1 - k];
  }
}

// After a float64 kernel ran on a group the doubt-tracking state is stale: its bits are certain (no doubtful samples),
// the error envelope restarts, and the amplitude scale — which the float64 kernels do not track — is set to a safe
// maximum (it decays to the true scale within a few hundred decimated samples).
static int fast_clean_doubt(Group& g, cudaStream_t st) {
  const size_t n = g.ids.size();
  for (int u : {(int)U_DVOTE, (int)U_SILX, (int)U_LAST_DOUBT, (int)U_DCNT})
    CUDA_TRY(cudaMemsetAsync(g.u32 + (size_t)u * n, 0, sizeof(uint32_t) * n, st));
  CUDA_TRY(cudaMemsetAsync(g.u32 + (size_t)U_SB_VALID * n, 0xff, sizeof(uint32_t) * n, st));
  for (int f : {(int)F_FAST_E, (int)F_FAST_RSP}) CUDA_TRY(cudaMemsetAsync(g.f64 + (size_t)f * n, 0, sizeof(double) * n, st));
  fill_f64_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(g.f64 + (size_t)F_FAST_S * n, 16.0, (long)n);
  CUDA_TRY(cudaGetLastError());
  return WAM_OK;
}

// The float64 kernels over the streams named by a list in device memory (sub-selection launches: the list length is
// read on the device, the grid is sized for `max_streams` and surplus CTAs leave at once).
static int launch_exact_selected(wam_fsk_batch* b, DemodLaunch& L, const int32_t* sel, const int32_t* sel_count,
                                 int max_streams, long n, cudaStream_t st) {
  L.slab = 0; L.slab_done = nullptr;
  L.n_groups = 1;
  DemodArgs& a = L.g[0];
  a.sel = sel; a.sel_count = sel_count;
  a.flag_list = nullptr; a.flag_count = nullptr;
  a.l_begin = 0; a.l_end = a.n_local;
  L.block_begin[0] = 0;
  L.block_begin[1] = (max_streams + 31) / 32;
  const int W = L.block_begin[1];
  const bool thin = a.thin_margin > 0.0;
  bool pipe = n >= 8 * kTile;
  if (pipe) {
    auto kern = thin ? fsk_demod_pipe_kernel<true, true> : fsk_demod_pipe_kernel<true, false>;
    int& per_sm_cached = thin ? b->pipe_per_sm_thin : b->pipe_per_sm[1];
    if (per_sm_cached < 0) {
      int ring_words = 0;
      for (auto& g : b->groups) ring_words = std::max(ring_words, g.d.ring_words);
      b->pipe_smem = sizeof(PipeShared) + (size_t)ring_words * 32 * sizeof(uint32_t);
      b->pipe_ring_smem = 1;
      if (b->pipe_smem > 72 * 1024) { b->pipe_smem = sizeof(PipeShared); b->pipe_ring_smem = 0; }
      int per_sm = 0;
      CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->pipe_smem));
      CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kPipeThreads, b->pipe_smem));
      per_sm_cached = per_sm;
    }
    L.pipe_ring_smem = b->pipe_ring_smem;
    pipe = per_sm_cached > 0;
    if (pipe) kern<<<W, kPipeThreads, b->pipe_smem, st>>>(L);
  }
  if (!pipe && thin) return fail(WAM_E_UNSUPPORTED, "thin-compare verification needs the pipeline kernel");
  if (!pipe) fsk_demod_exact_kernel<true, false><<<W, 32, 0, st>>>(L);
  b->launches++;
  CUDA_TRY(cudaGetLastError());
  return WAM_OK;
}

static size_t round_up(size_t v, size_t m) { return (v + m - 1) / m * m; }

// time slab length for a call of n samples (n a multiple of kTile): kFastSlabTiles tiles, or the divisor of the call's
// tile count nearest to it (so that no slab is short and every doubtful decision has a window).  A slab is also the
// warm-up of a verification window: ~1000 samples are dozens of time constants of the filters at the usual baud rates.
static long fast_slab_len(long n) {
  const long tiles = n / kTile;
  if (tiles <= kFastSlabTiles) return tiles * kTile;
  for (int delta = 0; delta <= kFastSlabTiles / 2; delta++)
    for (int sgn : {-1, 1}) {
      const long t = kFastSlabTiles + sgn * delta;
      if (t > 0 && tiles % t == 0) return t * kTile;
    }
  return (long)kFastSlabTiles * kTile;
}

static int fast_prepare_buffers(wam_fsk_batch* b, Group& g, const FastGeom& q, cudaStream_t st) {
  FastBuffers& fb = g.fb;
  const size_t ns = (size_t)q.ns;
  int rc = WAM_OK;
  if ((rc = ensure((void**)&fb.ck_f64, &fb.ck_f64_bytes, sizeof(double) * F64_COUNT * ns * q.n_slabs)) != WAM_OK) return rc;
  if ((rc = ensure((void**)&fb.ck_u32, &fb.ck_u32_bytes, sizeof(uint32_t) * U32_COUNT * ns * q.n_slabs)) != WAM_OK) return rc;
  if ((rc = ensure((void**)&fb.bit_hist, &fb.bit_hist_bytes, sizeof(uint16_t) * ns * q.bh_stride)) != WAM_OK) return rc;
  if ((rc = ensure((void**)&fb.amp_hist, &fb.amp_hist_bytes, sizeof(float) * ns * q.ah_stride)) != WAM_OK) return rc;
  if ((rc = ensure((void**)&fb.slab_list, &fb.slab_list_bytes, sizeof(int32_t) * ns * q.n_slabs)) != WAM_OK) return rc;
  if ((rc = ensure((void**)&fb.slab_count, &fb.slab_count_bytes, sizeof(int32_t) * q.n_slabs)) != WAM_OK) return rc;
  if ((rc = ensure((void**)&fb.hard_list, &fb.hard_list_bytes, sizeof(int32_t) * ns)) != WAM_OK) return rc;
  if ((rc = ensure((void**)&fb.hard_mark, &fb.hard_mark_bytes, sizeof(uint32_t) * ns)) != WAM_OK) return rc;
  const size_t nv = (size_t)kAllClasses * kVerifyCap;
  if (!fb.scratch_ready) {
    CUDA_TRY(cudaMalloc(&fb.hard_count, sizeof(int32_t) * 4));
    CUDA_TRY(cudaMalloc(&fb.item_li, sizeof(int32_t) * nv));
    CUDA_TRY(cudaMalloc(&fb.item_slab, sizeof(int32_t) * nv));
    CUDA_TRY(cudaMalloc(&fb.item_res, sizeof(int32_t) * nv));
    CUDA_TRY(cudaMalloc(&fb.item_count, sizeof(int32_t) * kAllClasses));
    CUDA_TRY(cudaMalloc(&fb.iota, sizeof(int32_t) * nv));
    std::vector<int32_t> h(nv);
    for (size_t i = 0; i < nv; i++) h[i] = (int32_t)i;
    CUDA_TRY(cudaMemcpy(fb.iota, h.data(), sizeof(int32_t) * nv, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&fb.sv_f64, sizeof(double) * F64_COUNT * nv));
    CUDA_TRY(cudaMalloc(&fb.sv_u32, sizeof(uint32_t) * U32_COUNT * nv));
    CUDA_TRY(cudaMalloc(&fb.sv_ring, sizeof(uint32_t) * (size_t)g.d.ring_words * nv));
    CUDA_TRY(cudaMalloc(&fb.sv_amp, sizeof(float) * (size_t)g.d.amp_phys * nv));
    CUDA_TRY(cudaMalloc(&fb.sv_out_len, sizeof(int32_t) * nv));
    fb.scratch_ready = true;
  }
  if ((rc = ensure((void**)&fb.sv_samples, &fb.sv_samples_bytes, sizeof(float) * nv * q.sv_stride)) != WAM_OK) return rc;
  if ((rc = ensure((void**)&fb.sv_out, &fb.sv_out_bytes, nv * (size_t)std::max<long>(q.out_stride, 1))) != WAM_OK) return rc;
  CUDA_TRY(cudaMemsetAsync(fb.slab_count, 0, sizeof(int32_t) * q.n_slabs, st));
  CUDA_TRY(cudaMemsetAsync(fb.hard_count, 0, sizeof(int32_t) * 4, st));
  CUDA_TRY(cudaMemsetAsync(fb.item_count, 0, sizeof(int32_t) * kAllClasses, st));
  (void)b;
  return WAM_OK;
}

static int fast_streams(wam_fsk_batch* b) {
  if (!b->slab_streams[0]) {
    // the carry save (last stream) runs beside the verification windows and must not get in their way: the windows are
    // a few latency-bound CTAs on the critical path, the carry save is thousands of copy blocks off it
    int lo = 0, hi = 0;
    CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    for (int i = 0; i < kSlabStreams; i++) {
      CUDA_TRY(cudaStreamCreateWithPriority(&b->slab_streams[i], cudaStreamNonBlocking, i == kSlabStreams - 1 ? lo : hi));
      CUDA_TRY(cudaEventCreateWithFlags(&b->slab_join[i], cudaEventDisableTiming));
    }
    CUDA_TRY(cudaEventCreateWithFlags(&b->slab_fork, cudaEventDisableTiming));
  }
  return WAM_OK;
}

// The kept slabs of the last fast call through the float64 kernel, then fast_carry_apply_kernel.  Runs on `st` ahead of
// whatever the call does next.
static int fast_carry_settle(wam_fsk_batch* b, Group& g, cudaStream_t st) {
  FastBuffers& fb = g.fb;
  if (!fb.carry_pending) return WAM_OK;
  fb.carry_pending = false;
  DemodLaunch Lv;
  memset(&Lv, 0, sizeof(Lv));
  DemodArgs& a = Lv.g[0];
  a.d = g.d;
  a.ids = nullptr; a.id0 = 0; a.row_base = 0;
  a.n_local = fb.carry_cap;
  a.f64 = fb.cv_f64; a.u32 = fb.cv_u32; a.sync_ring = fb.cv_ring; a.amp_ring = fb.cv_amp;
  a.samples = fb.cv_samples; a.stride = fb.carry_n; a.n = fb.carry_n;
  a.out = fb.cv_out; a.out_stride = fb.carry_out_stride; a.out_len = fb.cv_out_len;
  int rc = launch_exact_selected(b, Lv, fb.cv_iota, fb.carry_count, fb.carry_cap, a.n, st);
  if (rc != WAM_OK) return rc;
  CarryCtx cc;
  cc.ns = (int)g.ids.size(); cc.cap = fb.carry_cap; cc.ring_words = g.d.ring_words; cc.amp_phys = g.d.amp_phys; cc.amp_cap = g.d.amp_cap;
  cc.f64 = g.f64; cc.u32 = g.u32; cc.sync_ring = g.sync_ring; cc.amp_ring = g.amp_ring;
  cc.li = fb.carry_li; cc.count = fb.carry_count; cc.hard_mark = fb.hard_mark;
  cc.cv_f64 = fb.cv_f64; cc.cv_u32 = fb.cv_u32; cc.cv_ring = fb.cv_ring; cc.cv_amp = fb.cv_amp;
  fast_carry_apply_kernel<<<(unsigned)((fb.carry_cap + 3) / 4), 128, 0, st>>>(cc);
  CUDA_TRY(cudaGetLastError());
  return WAM_OK;
}

// End of a fast call: keep the last slabs of the streams with open readings (see above).
static int fast_carry_save(wam_fsk_batch* b, Group& g, const FastCtx& c, long out_cap_per_window, cudaStream_t st) {
  (void)b;
  FastBuffers& fb = g.fb;
  const FastGeom& q = c.q;
  const int need = std::max(std::max(q.sync_slabs, q.e2_len), 2);
  if (need > kVerifyClasses || q.n != (long)q.n_slabs * q.slab_len) return WAM_OK;  // no window could settle them: the record stays
  const int wslabs = std::min(need, q.n_slabs);
  const int cap = (std::max(1024, (q.ns + 3) / 4) + 31) / 32 * 32;  // (a multiple of the warp size: see the sub-selection in the kernels)
  if (!fb.carry_li || fb.carry_cap != cap) {
    for (void* p : {(void*)fb.carry_li, (void*)fb.carry_count, (void*)fb.cv_iota, (void*)fb.cv_f64, (void*)fb.cv_u32, (void*)fb.cv_ring,
                    (void*)fb.cv_amp, (void*)fb.cv_out_len})
      if (p) cudaFree(p);
    fb.carry_cap = cap;
    CUDA_TRY(cudaMalloc(&fb.carry_li, sizeof(int32_t) * (size_t)cap));
    CUDA_TRY(cudaMalloc(&fb.carry_count, sizeof(int32_t) * 4));
    CUDA_TRY(cudaMemsetAsync(fb.carry_count, 0, sizeof(int32_t) * 4, st));
    CUDA_TRY(cudaMalloc(&fb.cv_iota, sizeof(int32_t) * (size_t)cap));
    std::vector<int32_t> h((size_t)cap);
    for (int i = 0; i < cap; i++) h[(size_t)i] = i;
    CUDA_TRY(cudaMemcpy(fb.cv_iota, h.data(), sizeof(int32_t) * (size_t)cap, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&fb.cv_f64, sizeof(double) * F64_COUNT * (size_t)cap));
    CUDA_TRY(cudaMalloc(&fb.cv_u32, sizeof(uint32_t) * U32_COUNT * (size_t)cap));
    CUDA_TRY(cudaMalloc(&fb.cv_ring, sizeof(uint32_t) * (size_t)g.d.ring_words * (size_t)cap));
    CUDA_TRY(cudaMalloc(&fb.cv_amp, sizeof(float) * (size_t)g.d.amp_phys * (size_t)cap));
    CUDA_TRY(cudaMalloc(&fb.cv_out_len, sizeof(int32_t) * (size_t)cap));
  }
  fb.carry_n = (long)wslabs * q.slab_len;
  fb.carry_out_stride = std::max<long>(out_cap_per_window, 16);
  int rc;
  if ((rc = ensure((void**)&fb.cv_samples, &fb.cv_samples_bytes, sizeof(float) * (size_t)cap * (size_t)fb.carry_n)) != WAM_OK) return rc;
  if ((rc = ensure((void**)&fb.cv_out, &fb.cv_out_bytes, (size_t)cap * (size_t)fb.carry_out_stride)) != WAM_OK) return rc;
  CUDA_TRY(cudaMemsetAsync(fb.carry_count, 0, sizeof(int32_t), st));  // ([1], [2]: running totals for the statistics)
  fast_carry_collect_kernel<<<(unsigned)((q.ns + 255) / 256), 256, 0, st>>>(c, fb.carry_li, fb.carry_count, cap);
  const FastScratch d = {fb.cv_f64, fb.cv_u32, fb.cv_ring, fb.cv_amp, fb.cv_samples, fb.carry_n, fb.cv_out_len};
  fast_carry_gather_kernel<<<(unsigned)cap, 128, 0, st>>>(c, fb.carry_li, fb.carry_count, cap, wslabs, d);
  CUDA_TRY(cudaGetLastError());
  fb.carry_pending = true;
  return WAM_OK;
}

// L: the call's launch description (<= kFastGroupsPerLaunch groups, rows contiguous, TMA descriptors for the whole call).
static int fast_demodulate(wam_fsk_batch* b, DemodLaunch& L, Group* const* lg, const long* tmap_rows, long n, uint32_t flags,
                           cudaStream_t st) {
  const int G = L.n_groups;
  const int W = L.block_begin[G];
  const bool guarded = !(flags & WAM_BATCH_FAST_UNGUARDED);
  const bool tap = (flags & WAM_BATCH_TAP_FAST_DECISION) != 0 && L.g[0].tap != nullptr;
  const bool append = L.g[0].append != 0;
  int rc = fast_streams(b);
  if (rc != WAM_OK) return rc;
  if (b->fast_per_sm < 0) {
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fsk_demod_fast_kernel<false, true>, 32, 0));
    b->fast_per_sm = per_sm;
  }
  // slabs overlap on two streams per group when every CTA of the call is resident at once (see launch_slabbed);
  // otherwise they follow one another on the caller's stream
  const bool one_wave = W <= b->fast_per_sm * b->sm_count && !(flags & WAM_BATCH_NO_SLABS);
  const long slab_len = (flags & WAM_BATCH_NO_SLABS) ? n : fast_slab_len(n);
  const int n_slabs = (int)((n + slab_len - 1) / slab_len);

  FastCtx ctx[kFastGroupsPerLaunch];
  for (int gi = 0; gi < G; gi++) {
    Group& g = *lg[gi];
    const DemodArgs& a = L.g[gi];
    if (g.fb.carry_pending) { rc = fast_carry_settle(b, g, st); if (rc != WAM_OK) return rc; }
    if (g.doubt_state == 2) { rc = fast_clean_doubt(g, st); if (rc != WAM_OK) return rc; }
    g.doubt_state = 1;
    FastGeom q;
    q.ns = (int)g.ids.size(); q.n = n; q.slab_len = slab_len; q.n_slabs = n_slabs;
    q.ph = 2 * g.d.ring_words;
    q.pa = (int)round_up((size_t)g.d.amp_cap, 16);
    q.bh_stride = (long)round_up((size_t)q.ph + (size_t)(n / kTile), 8);
    q.ah_stride = (long)round_up((size_t)q.pa + (size_t)(n / 2), 4);
    q.sync_slabs = (int)((2L * g.d.total_bits + slab_len - 1) / slab_len) + 1;
    {
      // two-stage end-of-data check: amplitude ring before the sync / silent run before the decision, each behind a
      // warm-up.  Amplitudes only pass the full-rate filters (pre-filter, I/Q low-pass): sixteen time constants of the
      // slower one (the float32 state error of 1e-7 shrinks to 1e-14) are warm-up enough, a whole slab is not needed.
      const double decay = -std::log(std::sqrt(std::max(std::max(g.d.lp_a2, g.d.pre_a2), 1e-300)));
      const long warm = decay > 0.0 ? (long)std::ceil(16.0 / decay) : slab_len;
      const int l1 = (int)((2L * g.d.amp_cap + warm + slab_len - 1) / slab_len) + 1;
      const int l2 = (int)((2L * g.d.eod_count + warm + slab_len - 1) / slab_len) + 1;
      const bool whole = n == (long)n_slabs * slab_len;
      const bool ok = whole && l1 <= kVerifyClasses && l2 <= kVerifyClasses && l1 <= n_slabs && l2 <= n_slabs;
      q.e1_len = ok ? l1 : 0; q.e2_len = ok ? l2 : 0;
    }
    q.out_stride = a.out_stride;
    q.sv_stride = (long)kVerifyClasses * slab_len;
    rc = fast_prepare_buffers(b, g, q, st);
    if (rc != WAM_OK) return rc;
    FastCtx& c = ctx[gi];
    FastBuffers& fb = g.fb;
    c.q = q; c.ring_words = g.d.ring_words; c.amp_phys = g.d.amp_phys; c.amp_cap = g.d.amp_cap;
    c.f64 = g.f64; c.u32 = g.u32; c.ck_f64 = fb.ck_f64; c.ck_u32 = fb.ck_u32;
    c.sync_ring = g.sync_ring; c.amp_ring = g.amp_ring; c.bit_hist = fb.bit_hist; c.amp_hist = fb.amp_hist;
    c.slab_list = fb.slab_list; c.slab_count = fb.slab_count;
    c.hard_list = fb.hard_list; c.hard_count = fb.hard_count; c.hard_mark = fb.hard_mark;
    c.item_li = fb.item_li; c.item_slab = fb.item_slab; c.item_count = fb.item_count; c.item_res = fb.item_res;
    c.sv_f64 = fb.sv_f64; c.sv_u32 = fb.sv_u32; c.sv_ring = fb.sv_ring; c.sv_amp = fb.sv_amp;
    c.sv_samples = fb.sv_samples; c.sv_out = fb.sv_out; c.sv_out_len = fb.sv_out_len;
    c.samples = a.samples; c.stride = a.stride; c.ids = a.ids; c.id0 = a.id0; c.row_base = a.row_base;
    c.out = a.out; c.out_len = a.out_len; c.append = append ? 1 : 0;
    fast_prologue_kernel<<<q.ns, 128, 0, st>>>(c);
  }
  CUDA_TRY(cudaGetLastError());

  // ---- fast pass: one launch per group and time slab
  rc = ensure((void**)&b->slab_done, &b->slab_done_bytes, sizeof(int) * (size_t)W);
  if (rc != WAM_OK) return rc;
  CUDA_TRY(cudaMemsetAsync(b->slab_done, 0, sizeof(int) * (size_t)W, st));
  const int n_str = one_wave ? 2 * G : 0;
  if (one_wave) {
    CUDA_TRY(cudaEventRecord(b->slab_fork, st));
    for (int i = 0; i < n_str; i++) CUDA_TRY(cudaStreamWaitEvent(b->slab_streams[i], b->slab_fork, 0));
  }
  for (int slab = 0; slab < n_slabs; slab++) {
    const long t0 = (long)slab * slab_len;
    const long len = std::min(slab_len, n - t0);
    for (int gi = 0; gi < G; gi++) {
      Group& g = *lg[gi];
      const FastGeom& q = ctx[gi].q;
      DemodLaunch Lg;
      memset(&Lg, 0, sizeof(Lg));
      Lg.n_groups = 1;
      Lg.block_begin[1] = L.block_begin[gi + 1] - L.block_begin[gi];
      Lg.g[0] = L.g[gi];
      DemodArgs& a = Lg.g[0];
      a.samples = L.g[gi].samples + t0;
      a.n = len;
      if (slab > 0) a.append = 1;
      if (!make_sample_tmap(&Lg.tmap[0], a.samples, a.stride, len, tmap_rows[gi]))
        return fail(WAM_E_CUDA, "cuTensorMapEncodeTiled failed for a time slab");
      Lg.slab = slab;
      Lg.slab_done = one_wave ? b->slab_done + L.block_begin[gi] : nullptr;
      const size_t ns = (size_t)q.ns;
      a.f64 = slab == 0 ? g.f64 : g.fb.ck_f64 + (size_t)(slab - 1) * F64_COUNT * ns;
      a.u32 = slab == 0 ? g.u32 : g.fb.ck_u32 + (size_t)(slab - 1) * U32_COUNT * ns;
      a.f64_out = g.fb.ck_f64 + (size_t)slab * F64_COUNT * ns;
      a.u32_out = g.fb.ck_u32 + (size_t)slab * U32_COUNT * ns;
      a.bit_hist = g.fb.bit_hist; a.bh_stride = q.bh_stride; a.hist_t0 = q.ph + t0 / kTile;
      a.amp_hist = g.fb.amp_hist; a.ah_stride = q.ah_stride; a.amp_t0 = q.pa + t0 / 2;
      a.slab_list = g.fb.slab_list + (size_t)slab * ns; a.slab_count = g.fb.slab_count + slab;
      cudaStream_t sg = one_wave ? b->slab_streams[2 * gi + (slab % 2)] : st;
      const bool agc = g.d.agc_enabled != 0;
      auto kern = tap ? (agc ? fsk_demod_fast_kernel<true, true> : fsk_demod_fast_kernel<true, false>)
                      : (agc ? fsk_demod_fast_kernel<false, true> : fsk_demod_fast_kernel<false, false>);
      kern<<<Lg.block_begin[1], 32, 0, sg>>>(Lg);
      b->launches++;
    }
  }
  CUDA_TRY(cudaGetLastError());
  if (one_wave)
    for (int i = 0; i < n_str; i++) {
      CUDA_TRY(cudaEventRecord(b->slab_join[i], b->slab_streams[i]));
      CUDA_TRY(cudaStreamWaitEvent(st, b->slab_join[i], 0));
    }
  b->fast_calls++;

  // ---- check of the doubtful decisions (window classes and groups side by side), hard list, epilogue
  if (guarded) {
    for (int gi = 0; gi < G; gi++) fast_collect_kernel<<<dim3(4, (unsigned)std::min(n_slabs, 64)), 256, 0, st>>>(ctx[gi]);
    CUDA_TRY(cudaEventRecord(b->slab_fork, st));
    for (int gi = 0; gi < G; gi++) {
      Group& g = *lg[gi];
      FastCtx& c = ctx[gi];
      const FastGeom& q = c.q;
      for (int cls = 0; cls < kAllClasses; cls++) {
        if (cls < kVerifyClasses && (long)(cls + 1) * slab_len > n && cls > 0) continue;  // no window of the call is this long
        if (cls >= kVerifyClasses && q.e2_len == 0) continue;
        // the two stages follow one another on one stream
        const int sidx = gi * (kVerifyClasses + 1) + std::min(cls, kVerifyClasses);
        cudaStream_t sv = b->slab_streams[sidx];
        if (cls != kStageE2) CUDA_TRY(cudaStreamWaitEvent(sv, b->slab_fork, 0));
        fast_gather_kernel<<<kVerifyCap, 128, 0, sv>>>(c, cls);
        DemodLaunch Lv;
        memset(&Lv, 0, sizeof(Lv));
        DemodArgs& a = Lv.g[0];
        a.d = g.d;
        a.ids = nullptr; a.id0 = 0; a.row_base = 0;
        a.n_local = kAllClasses * kVerifyCap;
        a.f64 = g.fb.sv_f64; a.u32 = g.fb.sv_u32; a.sync_ring = g.fb.sv_ring; a.amp_ring = g.fb.sv_amp;
        a.samples = g.fb.sv_samples; a.stride = q.sv_stride;
        a.n = (long)(cls < kVerifyClasses ? cls + 1 : cls == kStageE1 ? q.e1_len : q.e2_len) * slab_len;
        a.out = g.fb.sv_out; a.out_stride = q.out_stride; a.out_len = g.fb.sv_out_len;
        a.thin_margin = cls == kStageE2 ? 1e-9 : 0.0;
        rc = launch_exact_selected(b, Lv, g.fb.iota + cls * kVerifyCap, g.fb.item_count + (cls == kStageE2 ? kStageE1 : cls),
                                   kVerifyCap, a.n, sv);
        if (rc != WAM_OK) return rc;
        fast_compare_kernel<<<kVerifyCap / 4, 128, 0, sv>>>(c, cls);
        if (cls == kStageE1) continue;
        CUDA_TRY(cudaEventRecord(b->slab_join[sidx], sv));
        CUDA_TRY(cudaStreamWaitEvent(st, b->slab_join[sidx], 0));
      }
    }
    {
      // behind the checks (they may still correct a byte under construction in the checkpoints), beside the hard list
      // and the epilogue: keep the last slabs of the streams that end the call with open readings
      cudaStream_t sc = b->slab_streams[kSlabStreams - 1];
      CUDA_TRY(cudaEventRecord(b->slab_fork, st));
      CUDA_TRY(cudaStreamWaitEvent(sc, b->slab_fork, 0));
      const long wcap = wam_fsk_batch_out_capacity(b, (long)kVerifyClasses * slab_len);
      for (int gi = 0; gi < G; gi++)
        if ((rc = fast_carry_save(b, *lg[gi], ctx[gi], wcap, sc)) != WAM_OK) return rc;
      CUDA_TRY(cudaEventRecord(b->slab_join[kSlabStreams - 1], sc));
    }
    // whole-call float64 run of the hard lists, on the live state and rings
    for (int gi = 0; gi < G; gi++) {
      Group& g = *lg[gi];
      fast_hard_prepare_kernel<<<8, 256, 0, st>>>(ctx[gi]);
      DemodLaunch Lh;
      memset(&Lh, 0, sizeof(Lh));
      Lh.g[0] = L.g[gi];
      Lh.g[0].append = 1;  // fast_hard_prepare_kernel has put the byte counts back to the start of the call
      Lh.g[0].tap = nullptr;
      rc = launch_exact_selected(b, Lh, g.fb.hard_list, g.fb.hard_count, ctx[gi].q.ns, n, st);
      if (rc != WAM_OK) return rc;
    }
  }
  for (int gi = 0; gi < G; gi++) fast_epilogue_kernel<<<ctx[gi].q.ns, 128, 0, st>>>(ctx[gi]);
  if (guarded) CUDA_TRY(cudaStreamWaitEvent(st, b->slab_join[kSlabStreams - 1], 0));  // the carry save
  CUDA_TRY(cudaGetLastError());
  return WAM_OK;
}
