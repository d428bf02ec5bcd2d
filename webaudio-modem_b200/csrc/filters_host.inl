// filters_host.inl — host drivers of the batched IIR / FIR kernels (included by wam_api.cu).

namespace wam {

static void mat_mul(const double* A, const double* B, double* C, int M) {
  double tmp[kMaxM * kMaxM];
  for (int r = 0; r < M; r++)
    for (int c = 0; c < M; c++) {
      double acc = 0;
      for (int k = 0; k < M; k++) acc += A[r * M + k] * B[k * M + c];
      tmp[r * M + c] = acc;
    }
  memcpy(C, tmp, sizeof(double) * (size_t)(M * M));
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// scratch of one IIR call: ticket counter, then per (stream, tile) a flag and two state vectors
static size_t iir_tiles(long n) { return (size_t)((n + kIirTile - 1) / kIirTile); }
size_t iir_scratch_bytes(long n, long n_streams) {
  const size_t recs = iir_tiles(n) * (size_t)std::max<long>(n_streams, 0);
  return 256 + (recs * sizeof(int) + 255) / 256 * 256 + 2 * recs * sizeof(double) * kMaxM;
}

int iir_process_batch_device(IirArgs& ia, const float* d_in, float* d_out, long stride, long n, long n_streams, double* d_state,
                             void* d_scratch, size_t scratch_bytes, cudaStream_t st) {
  const int M = ia.na - 1;
  if (n <= 0 || n_streams <= 0) return WAM_OK;
  if (scratch_bytes < iir_scratch_bytes(n, n_streams) || !d_scratch) return fail(WAM_E_CAPACITY, "IIR scratch smaller than wam_iir_scratch_bytes()");
  const size_t tiles = iir_tiles(n);
  if (tiles * (size_t)n_streams > 0x7fffffffull) return fail(WAM_E_UNSUPPORTED, "wam_iir_process_batch: more than 2^31 tiles in one call");
  // companion matrix A of the feedback taps, T = A^kIirChunk, T^(2^k), then the span and tile transitions
  if (M > 0) {
    double A[kMaxM * kMaxM] = {0};
    for (int c = 0; c < M; c++) A[c] = -ia.a[c + 1];
    for (int r = 1; r < M; r++) A[r * M + (r - 1)] = 1.0;
    double T[kMaxM * kMaxM];
    memcpy(T, A, sizeof(A));
    for (int chunk = kIirChunk; chunk > 1; chunk >>= 1) mat_mul(T, T, T, M);  // kIirChunk is a power of two
    memcpy(ia.tpow[0], T, sizeof(double) * (size_t)(M * M));
    for (int k = 1; k < 5; k++) mat_mul(ia.tpow[k - 1], ia.tpow[k - 1], ia.tpow[k], M);
    mat_mul(ia.tpow[4], ia.tpow[4], ia.tspan[0], M);  // T^32: one span
    memcpy(ia.tspan[1], ia.tspan[0], sizeof(double) * (size_t)(M * M));
    for (int j = 1; j < kIirWarps; j++) mat_mul(ia.tspan[1], ia.tspan[0], ia.tspan[1], M);  // one tile
  }
  const size_t recs = tiles * (size_t)n_streams;
  char* sc = (char*)d_scratch;
  ia.ticket = (unsigned int*)sc;
  ia.tile_flag = (int*)(sc + 256);
  ia.tile_aggr = (double*)(sc + 256 + (recs * sizeof(int) + 255) / 256 * 256);
  ia.tile_incl = ia.tile_aggr + recs * kMaxM;
  CUDA_TRY(cudaMemsetAsync(sc, 0, 256 + recs * sizeof(int), st));
  ia.in = d_in; ia.out = d_out; ia.stride = stride; ia.n = n; ia.n_streams = n_streams; ia.tiles = (int)tiles;
  ia.vec = (stride % 4 == 0 && aligned16(d_in) && aligned16(d_out)) ? 1 : 0;
  ia.state = (ia.nb - 1 + M > 0) ? d_state : nullptr;
  const size_t smem = sizeof(float) * kIirWarps * 32 * kIirPitch;
  const unsigned grid = (unsigned)recs;
  if (ia.nb == 3 && ia.na == 3) iir_scan_kernel<2, 2><<<grid, kIirWarps * 32, smem, st>>>(ia);  // the biquads of FilterFactory
  else iir_scan_kernel<-1, -1><<<grid, kIirWarps * 32, smem, st>>>(ia);
  CUDA_TRY(cudaGetLastError());
  return WAM_OK;
}

int iir_process_batch_host(IirArgs& ia, const float* in, float* out, long stride, long n, long n_streams, double* state) {
  const int stw = (ia.nb - 1) + (ia.na - 1);
  DevBuf din, dout, dst, dsc;
  int rc;
  const size_t nb = sizeof(float) * (size_t)stride * (size_t)n_streams;
  const size_t sb = iir_scratch_bytes(n, n_streams);
  if ((rc = din.alloc(nb)) || (rc = dout.alloc(nb)) || (rc = dsc.alloc(sb))) return rc;
  if (state && stw > 0) {
    if ((rc = dst.alloc(sizeof(double) * (size_t)stw * (size_t)n_streams))) return rc;
    CUDA_TRY(cudaMemcpy(dst.p, state, sizeof(double) * (size_t)stw * (size_t)n_streams, cudaMemcpyHostToDevice));
  }
  CUDA_TRY(cudaMemcpy(din.p, in, nb, cudaMemcpyHostToDevice));
  rc = iir_process_batch_device(ia, (const float*)din.p, (float*)dout.p, stride, n, n_streams, (state && stw > 0) ? (double*)dst.p : nullptr,
                                dsc.p, sb, nullptr);
  if (rc != WAM_OK) return rc;
  CUDA_TRY(cudaMemcpy(out, dout.p, nb, cudaMemcpyDeviceToHost));
  if (state && stw > 0)
    CUDA_TRY(cudaMemcpy(state, dst.p, sizeof(double) * (size_t)stw * (size_t)n_streams, cudaMemcpyDeviceToHost));
  return WAM_OK;
}

// d_state: nullable [n_streams][ntaps - 1], d_state_new: scratch of the same size (required with d_state)
int fir_process_batch_device(const double* d_taps, int ntaps, const float* d_in, float* d_out, long stride, long n,
                             long n_streams, double* d_state, double* d_state_new, cudaStream_t st) {
  if (n <= 0 || n_streams <= 0) return WAM_OK;
  if (ntaps == 0) {  // a filter with no taps outputs zeros
    CUDA_TRY(cudaMemset2DAsync(d_out, sizeof(float) * (size_t)stride, 0, sizeof(float) * (size_t)n, (size_t)n_streams, st));
    return WAM_OK;
  }
  const int hist = ntaps - 1;
  if (d_state && hist > 0 && !d_state_new) return fail(WAM_E_INVALID, "FIR state needs a scratch buffer of the same size");
  FirArgs fa;
  fa.taps = d_taps; fa.ntaps = ntaps; fa.in = d_in; fa.out = d_out; fa.stride = stride; fa.n = n; fa.n_streams = n_streams;
  fa.tiles = (int)((n + kFirTile - 1) / kFirTile);
  if ((size_t)fa.tiles * (size_t)n_streams > 0x7fffffffull) return fail(WAM_E_UNSUPPORTED, "wam_fir_process_batch: more than 2^31 tiles in one call");
  fa.vec = (stride % 4 == 0 && aligned16(d_in) && aligned16(d_out)) ? 1 : 0;
  fa.state = hist > 0 ? d_state : nullptr;
  fa.state_new = (hist > 0 && d_state) ? d_state_new : nullptr;
  const size_t smem = sizeof(double) * (size_t)fir_smem_doubles(ntaps);
  if (smem > 48 * 1024)  // opt-in above the default limit (per device: set on every such call, it is cheap)
    CUDA_TRY(cudaFuncSetAttribute(fir_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  fir_kernel<<<(unsigned)((size_t)fa.tiles * (size_t)n_streams), kFirThreads, smem, st>>>(fa);
  CUDA_TRY(cudaGetLastError());
  if (fa.state_new)
    CUDA_TRY(cudaMemcpyAsync(d_state, d_state_new, sizeof(double) * (size_t)hist * (size_t)n_streams, cudaMemcpyDeviceToDevice, st));
  return WAM_OK;
}

int fir_process_batch_host(const double* taps, int ntaps, const float* in, float* out, long stride, long n,
                           long n_streams, double* state) {
  const size_t nb = sizeof(float) * (size_t)stride * (size_t)n_streams;
  if (ntaps == 0) {
    for (long s = 0; s < n_streams; s++) memset(out + s * stride, 0, sizeof(float) * (size_t)n);
    return WAM_OK;
  }
  const int hist = ntaps - 1;
  DevBuf din, dout, dtaps, dsi, dso;
  int rc;
  if ((rc = din.alloc(nb)) || (rc = dout.alloc(nb)) || (rc = dtaps.alloc(sizeof(double) * (size_t)ntaps))) return rc;
  CUDA_TRY(cudaMemcpy(din.p, in, nb, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(dtaps.p, taps, sizeof(double) * (size_t)ntaps, cudaMemcpyHostToDevice));
  const size_t sb = sizeof(double) * (size_t)hist * (size_t)n_streams;
  if (state && hist > 0) {
    if ((rc = dsi.alloc(sb)) || (rc = dso.alloc(sb))) return rc;
    CUDA_TRY(cudaMemcpy(dsi.p, state, sb, cudaMemcpyHostToDevice));
  }
  rc = fir_process_batch_device((const double*)dtaps.p, ntaps, (const float*)din.p, (float*)dout.p, stride, n, n_streams,
                                (state && hist > 0) ? (double*)dsi.p : nullptr, (double*)dso.p, nullptr);
  if (rc != WAM_OK) return rc;
  CUDA_TRY(cudaStreamSynchronize(nullptr));
  CUDA_TRY(cudaMemcpy(out, dout.p, nb, cudaMemcpyDeviceToHost));
  if (state && hist > 0) CUDA_TRY(cudaMemcpy(state, dsi.p, sb, cudaMemcpyDeviceToHost));
  return WAM_OK;
}

}  // namespace wam
