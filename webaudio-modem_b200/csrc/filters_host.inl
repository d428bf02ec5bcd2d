// filters_host.inl — host drivers of the batched IIR / FIR kernels (included by wam_api.cu).

namespace wam {

static void mat_mul(const double* A, const double* B, double* C, int M) {
  double tmp[(kMaxIirTaps - 1) * (kMaxIirTaps - 1)];
  for (int r = 0; r < M; r++)
    for (int c = 0; c < M; c++) {
      double acc = 0;
      for (int k = 0; k < M; k++) acc += A[r * M + k] * B[k * M + c];
      tmp[r * M + c] = acc;
    }
  memcpy(C, tmp, sizeof(double) * (size_t)(M * M));
}

int iir_process_batch_host(IirArgs& ia, const float* in, float* out, long stride, long n, long n_streams, double* state) {
  const int M = ia.na - 1, NX = ia.nb - 1;
  if (n_streams > 65535) return fail(WAM_E_UNSUPPORTED, "wam_iir_process_batch: at most 65535 streams per call");
  // companion matrix A of the feedback taps, T = A^kIirChunk, then T^(2^k)
  if (M > 0) {
    double A[(kMaxIirTaps - 1) * (kMaxIirTaps - 1)] = {0};
    for (int c = 0; c < M; c++) A[c] = -ia.a[c + 1];
    for (int r = 1; r < M; r++) A[r * M + (r - 1)] = 1.0;
    double T[(kMaxIirTaps - 1) * (kMaxIirTaps - 1)];
    memcpy(T, A, sizeof(A));
    for (int chunk = kIirChunk; chunk > 1; chunk >>= 1) mat_mul(T, T, T, M);  // kIirChunk is a power of two
    memcpy(ia.tpow[0], T, sizeof(double) * (size_t)(M * M));
    for (int k = 1; k < 6; k++) mat_mul(ia.tpow[k - 1], ia.tpow[k - 1], ia.tpow[k], M);
  }
  const long spans = (n + kIirWarpSpan - 1) / kIirWarpSpan;
  const int stw = NX + M;
  DevBuf din, dout, dst_in, dst_out, dcz, dsz, dss;
  int rc;
  const size_t nb = sizeof(float) * (size_t)stride * (size_t)n_streams;
  const size_t mz = sizeof(double) * (size_t)std::max(M, 1);
  if ((rc = din.alloc(nb)) || (rc = dout.alloc(nb)) || (rc = dcz.alloc(mz * 32 * (size_t)spans * (size_t)n_streams)) ||
      (rc = dsz.alloc(mz * (size_t)spans * (size_t)n_streams)) || (rc = dss.alloc(mz * (size_t)spans * (size_t)n_streams)))
    return rc;
  if (state && stw > 0) {
    if ((rc = dst_in.alloc(sizeof(double) * (size_t)stw * (size_t)n_streams)) ||
        (rc = dst_out.alloc(sizeof(double) * (size_t)stw * (size_t)n_streams)))
      return rc;
    CUDA_TRY(cudaMemcpy(dst_in.p, state, sizeof(double) * (size_t)stw * (size_t)n_streams, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(dst_out.p, state, sizeof(double) * (size_t)stw * (size_t)n_streams, cudaMemcpyHostToDevice));
  }
  CUDA_TRY(cudaMemcpy(din.p, in, nb, cudaMemcpyHostToDevice));
  ia.in = (const float*)din.p; ia.out = (float*)dout.p; ia.stride = stride; ia.n = n;
  ia.n_streams = (int)n_streams; ia.spans = (int)spans;
  ia.state_in = (state && stw > 0) ? (const double*)dst_in.p : nullptr;
  ia.state_out = (state && stw > 0) ? (double*)dst_out.p : nullptr;
  ia.chunk_z = (double*)dcz.p; ia.span_z = (double*)dsz.p; ia.span_start = (double*)dss.p;
  const size_t smem = sizeof(float) * 32 * (kIirChunk + 1);
  dim3 grid((unsigned)spans, (unsigned)n_streams);
  iir_span_kernel<0><<<grid, 32, smem>>>(ia);
  CUDA_TRY(cudaGetLastError());
  iir_span_carry_kernel<<<(unsigned)((n_streams + 127) / 128), 128>>>(ia);
  CUDA_TRY(cudaGetLastError());
  iir_span_kernel<2><<<grid, 32, smem>>>(ia);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpy(out, dout.p, nb, cudaMemcpyDeviceToHost));
  if (state && stw > 0)
    CUDA_TRY(cudaMemcpy(state, dst_out.p, sizeof(double) * (size_t)stw * (size_t)n_streams, cudaMemcpyDeviceToHost));
  return WAM_OK;
}

int fir_process_batch_host(const double* taps, int ntaps, const float* in, float* out, long stride, long n,
                           long n_streams, double* state) {
  if (n_streams > 65535) return fail(WAM_E_UNSUPPORTED, "wam_fir_process_batch: at most 65535 streams per call");
  const size_t nb = sizeof(float) * (size_t)stride * (size_t)n_streams;
  if (ntaps == 0) {  // a filter with no taps outputs zeros
    for (long s = 0; s < n_streams; s++) memset(out + s * stride, 0, sizeof(float) * (size_t)n);
    return WAM_OK;
  }
  const int hist = ntaps - 1;
  DevBuf din, dout, dtaps, dsi, dso;
  int rc;
  if ((rc = din.alloc(nb)) || (rc = dout.alloc(nb)) || (rc = dtaps.alloc(sizeof(double) * (size_t)ntaps))) return rc;
  CUDA_TRY(cudaMemcpy(din.p, in, nb, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(dtaps.p, taps, sizeof(double) * (size_t)ntaps, cudaMemcpyHostToDevice));
  FirArgs fa;
  fa.taps = (const double*)dtaps.p; fa.ntaps = ntaps;
  fa.in = (const float*)din.p; fa.out = (float*)dout.p; fa.stride = stride; fa.n = n;
  fa.state_in = nullptr; fa.state_out = nullptr;
  if (state && hist > 0) {
    const size_t sb = sizeof(double) * (size_t)hist * (size_t)n_streams;
    if ((rc = dsi.alloc(sb)) || (rc = dso.alloc(sb))) return rc;
    CUDA_TRY(cudaMemcpy(dsi.p, state, sb, cudaMemcpyHostToDevice));
    fa.state_in = (const double*)dsi.p; fa.state_out = (double*)dso.p;
  }
  const size_t smem = sizeof(double) * (size_t)(ntaps + kFirTile + hist);
  if (smem > 48 * 1024)
    CUDA_TRY(cudaFuncSetAttribute(fir_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((n + kFirTile - 1) / kFirTile), (unsigned)n_streams);
  fir_kernel<<<grid, kFirTile, smem>>>(fa);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpy(out, dout.p, nb, cudaMemcpyDeviceToHost));
  if (state && hist > 0)
    CUDA_TRY(cudaMemcpy(state, dso.p, sizeof(double) * (size_t)hist * (size_t)n_streams, cudaMemcpyDeviceToHost));
  return WAM_OK;
}

}  // namespace wam
