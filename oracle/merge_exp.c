/*
 * merge_exp.c — experiment (test infrastructure, NOT product code): do two FSKCore instances that start from different
 * states on the same stream end up in the same state?  The reference restatement (wam_oracle.c) runs a stream from its
 * beginning (the TRUE run); a second instance (the SPECULATIVE run) starts `warmup` samples before position T from
 * the freshly configured state, only its globalSampleCounter preset to the value the true run would have there if no
 * resetState() had happened since the start of the stream.  After every input sample from T on the two states are
 * compared: (B) every integer of the state machine, the sync ring's and amplitude ring's contents, the silence
 * threshold; (D) the DSP state — AGC gain, filter histories, LO phase, last phase — bitwise and to 1e-12 relative.
 * Reported: samples past T until the states agree and stay in agreement to the end of the stream, or -1.
 * This is the question behind time-parallel demodulation of long streams (speculative chunks validated at hand-over).
 */
#include "wam_oracle.c"

typedef struct {
  int64_t merged_b;       /* samples past T until the state machine state is equal for good (-1: never) */
  int64_t merged_close;   /* ... and the DSP state within 1e-12 relative */
  int64_t merged_exact;   /* ... and the DSP state bitwise equal */
  int64_t gsc_guess_ok;   /* the no-reset guess of globalSampleCounter was the true value at T - warmup */
  int64_t bytes_true, bytes_spec_after_merge_equal; /* decoded bytes of the true run past the merge point / identical ones of the speculative run */
} merge_result;

static int ring_tail_equal(const wamo_ring* a, const wamo_ring* b, long count, int bitwise, double tol) {
  if (a->length < count || b->length < count) return a->length == b->length && count == 0;
  for (long i = 1; i <= count; i++) {
    double va = a->buf[(long)fmod(a->writeIndex - i + 4 * a->maxLength, a->maxLength)];
    double vb = b->buf[(long)fmod(b->writeIndex - i + 4 * b->maxLength, b->maxLength)];
    if (bitwise ? (va != vb) : (fabs(va - vb) > tol * (fabs(va) + 1e-30))) return 0;
  }
  return 1;
}
static int iir_cmp(const wamo_iir* a, const wamo_iir* b, int bitwise, double tol) {
  /* histories in logical order */
  for (int k = 0; k < a->nx; k++) {
    double va = a->x[(a->xIndex - k + 8 * a->nx) % a->nx], vb = b->x[(b->xIndex - k + 8 * b->nx) % b->nx];
    if (bitwise ? (va != vb) : (fabs(va - vb) > tol * (fabs(va) + fabs(vb)) + 1e-300)) return 0;
  }
  for (int k = 0; k < a->ny; k++) {
    double va = a->y[(a->yIndex - k + 8 * a->ny) % a->ny], vb = b->y[(b->yIndex - k + 8 * b->ny) % b->ny];
    if (bitwise ? (va != vb) : (fabs(va - vb) > tol * (fabs(va) + fabs(vb)) + 1e-300)) return 0;
  }
  return 1;
}
static int b_equal(const wamo_fsk* a, const wamo_fsk* b) {
  long nb = (long)(a->nbits * a->downsampledSamplesPerBit);
  return a->started == b->started && a->globalSampleCounter == b->globalSampleCounter &&
         a->bitSampleCounter == b->bitSampleCounter && a->bitAccumulator == b->bitAccumulator &&
         a->bitAccumCount == b->bitAccumCount && a->nextBitSampleIndex == b->nextBitSampleIndex &&
         a->current == b->current && a->bitPosition == b->bitPosition && a->silenceCount == b->silenceCount &&
         a->dsCounter == b->dsCounter && a->silenceThreshold == b->silenceThreshold &&
         ring_tail_equal(a->syncSamples, b->syncSamples, nb, 1, 0) &&
         ring_tail_equal(a->syncAmplitude, b->syncAmplitude, (long)a->syncAmplitude->maxLength, 1, 0);
}
static int d_equal(const wamo_fsk* a, const wamo_fsk* b, int bitwise, double tol) {
#define CMP(x, y) (bitwise ? ((x) == (y)) : (fabs((x) - (y)) <= tol * (fabs(x) + fabs(y)) + 1e-300))
  return CMP(a->agc.currentGain, b->agc.currentGain) && CMP(a->localOscPhase, b->localOscPhase) &&
         CMP(a->lastPhase, b->lastPhase) && CMP(a->iAcc, b->iAcc) && CMP(a->qAcc, b->qAcc) &&
         iir_cmp(a->preFilter, b->preFilter, bitwise, tol) && iir_cmp(a->iqI, b->iqI, bitwise, tol) &&
         iir_cmp(a->iqQ, b->iqQ, bitwise, tol) && iir_cmp(a->postFilter, b->postFilter, bitwise, tol);
#undef CMP
}

int merge_experiment(const wamo_fsk_config* cfg, const float* samples, long n, long T, long warmup, merge_result* r) {
  wamo_fsk* t = wamo_fsk_new(); wamo_fsk* s = wamo_fsk_new();
  wamo_fsk_configure(t, cfg); wamo_fsk_configure(s, cfg);
  uint8_t ob[64];
  long start = T - warmup; if (start < 0) start = 0;
  start -= start & 1;  /* keep the decimator phase */
  float x;
  for (long i = 0; i < start; i++) { x = samples[i]; wamo_fsk_demodulate(t, &x, 1, ob, 64); }
  r->gsc_guess_ok = (t->globalSampleCounter == (double)(start / 2));
  s->globalSampleCounter = (double)(start / 2);
  /* what a real implementation can carry without running the stream: the sync ring's fill level */
  r->merged_b = r->merged_close = r->merged_exact = -1;
  long last_b = -1, last_c = -1, last_e = -1;  /* start of the current run of agreement */
  long tb = 0, sb_same = 0;
  uint8_t* tbytes = (uint8_t*)malloc((size_t)(n / 8 + 64)); uint8_t* sbytes = (uint8_t*)malloc((size_t)(n / 8 + 64));
  long* tpos = (long*)malloc(sizeof(long) * (size_t)(n / 8 + 64)); long* spos = (long*)malloc(sizeof(long) * (size_t)(n / 8 + 64));
  long nt = 0, nsb = 0;
  for (long i = start; i < n; i++) {
    x = samples[i]; long k = wamo_fsk_demodulate(t, &x, 1, ob, 64);
    for (long j = 0; j < k; j++) { tbytes[nt] = ob[j]; tpos[nt++] = i; }
    x = samples[i]; k = wamo_fsk_demodulate(s, &x, 1, ob, 64);
    for (long j = 0; j < k; j++) { sbytes[nsb] = ob[j]; spos[nsb++] = i; }
    if (i < T) continue;
    int be = b_equal(t, s);
    int ce = be && d_equal(t, s, 0, 1e-12), ee = be && d_equal(t, s, 1, 0);
    if (be) { if (last_b < 0) last_b = i; } else last_b = -1;
    if (ce) { if (last_c < 0) last_c = i; } else last_c = -1;
    if (ee) { if (last_e < 0) last_e = i; } else last_e = -1;
  }
  if (last_b >= 0) r->merged_b = last_b - T;
  if (last_c >= 0) r->merged_close = last_c - T;
  if (last_e >= 0) r->merged_exact = last_e - T;
  /* bytes past the merge point */
  if (last_b >= 0) {
    long a = 0, b2 = 0;
    while (a < nt && tpos[a] <= last_b) a++;
    while (b2 < nsb && spos[b2] <= last_b) b2++;
    tb = nt - a;
    while (a < nt && b2 < nsb && tbytes[a] == sbytes[b2] && tpos[a] == spos[b2]) { a++; b2++; sb_same++; }
  }
  r->bytes_true = tb; r->bytes_spec_after_merge_equal = sb_same;
  free(tbytes); free(sbytes); free(tpos); free(spos);
  wamo_fsk_free(t); wamo_fsk_free(s);
  return 0;
}
