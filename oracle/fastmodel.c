/*
 * fastmodel.c — CPU MODEL of the mixed-precision fast demodulator (test infrastructure, NOT product code).
 *
 * The GPU fast path (webaudio-modem_b200/csrc/fsk_demod_fast.cuh) runs FSKCore.demodulateData
 * (/root/reference/src/modems/fsk.ts:190-375) with float32 arithmetic and certifies every DECISION of the
 * state machine against the float32 error: a decimated sample whose filtered phase difference is closer to the
 * slicer threshold than the error bound is "doubtful", and a vote / sync / silence decision that a doubtful sample
 * could turn flags the stream, which is then demodulated again by the exact float64 kernel.  This file is the same
 * algorithm in plain C (float32 DSP with fmaf, per-sample state machine), used
 *   - to measure the float32 error against the oracle's float64 filteredPhaseDiff (how wide the doubt band must be),
 *   - to measure how many streams get flagged per SNR class and why,
 *   - as the specification the CUDA kernel is tested against (tests/test_fastmodel.py).
 * It includes wam_oracle.c for the configuration / filter-design code of the reference restatement.
 */
#include "wam_oracle.c"

typedef struct fm_params {
  double eps0;      /* floor of the doubt band on |filteredPhaseDiff| */
  double kappa;     /* relative float32 error of an I/Q filter output against the recent amplitude scale */
  double eps_amp;   /* relative doubt band of the silence compare amplitude < threshold */
  double bc_delta;  /* a raw phase difference within bc_delta of +-pi may have wrapped the other way */
  int32_t form;     /* 0: direct-form-I float32 biquads, 1: normal (coupled) form */
  int32_t unguarded;/* 1: no doubt tracking at all (plain float32 run) */
} fm_params;

enum { FM_CAUSE_VOTE_START = 0, FM_CAUSE_VOTE_DATA, FM_CAUSE_VOTE_STOP, FM_CAUSE_SYNC, FM_CAUSE_EOD, FM_CAUSE_RANGE, FM_NCAUSE };

typedef struct fm_result {
  int32_t n_out;
  int32_t flag;            /* some decision was doubtful */
  int32_t first_cause, first_index; /* first doubtful decision: cause and decimated index */
  int32_t n_cause[FM_NCAUSE];
  int32_t n_doubt_samples, n_dec;
  int32_t n_bc;            /* branch-cut events */
  int32_t started;
  double syncDetections, eodEvents, gsc, sil_thr;
  double err_max, err_rms; /* |F_fast - F_oracle| over the decimated samples compared (oracle tap given) */
  double ratio_max;        /* max |F_fast - F_oracle| / doubt band */
  int32_t n_wrong_bits, n_wrong_undoubted; /* hard bits that differ from the oracle's; those outside the doubt band */
  double w_index, w_amp, w_S, w_err, w_band, w_pd, w_oamp, w_F; /* the sample of ratio_max */
} fm_result;

typedef struct {
  float k0, k1, k2, sg, om;       /* normal form: y = k0 x + k1 w1 + k2 w2; w' = R w + e1 x */
  float b0, b1, b2, a1, a2;       /* direct form */
} fm_biquad;

typedef struct {
  fm_biquad pre, lp;
  float att, rel; int agc; double attd, reld, gd;
  float cw, sw; double dphi_bias;
  int dspb, nbits, total_bits, check_period, stop_pos, parity, eod_count, min_matched, ring_cap, amp_cap, bc_hold;
  double rho_e;                    /* envelope decay of the post filter's impulse response per decimated sample */
  double gamma_e;
  int* pattern;
  /* A1 */
  float g, pw1, pw2, px1, px2, py1, py2;
  /* A2 */
  float lc, ls, iw1, iw2, qw1, qw2, ow1, ow2, psi, psq, iacc, qacc;
  float ix1, ix2, iy1, iy2, qx1, qx2, qy1, qy2, ox1, ox2, oy1, oy2;
  int dsc, lo_age;
  /* doubt envelope */
  float S, E, rs_prev; int bc_left;
  /* B */
  uint32_t gsc, bsc, next_idx, bit_acc, bit_cnt, sil_cnt; int started, bitpos, current;
  double sil_thr;
  uint8_t *ring, *dring; int ring_len; long ring_pos; int d_win;
  float* aring; int amp_len; long amp_pos;
  uint32_t d_ones, d_zeros; int amp_pending; uint32_t sil_extra;
  long dec_index;
} fm_t;

static void fm_design(fm_biquad* q, const double b[3], const double a[3]) {
  q->b0 = (float)b[0]; q->b1 = (float)b[1]; q->b2 = (float)b[2]; q->a1 = (float)a[1]; q->a2 = (float)a[2];
  const double sg = -a[1] / 2, om2 = a[2] - sg * sg;
  const double om = om2 > 0 ? sqrt(om2) : 0;
  const double c1 = b[1] - b[0] * a[1];
  const double c2 = om > 0 ? (b[2] - b[0] * a[2] + c1 * sg) / om : 0;
  q->k0 = (float)b[0]; q->k1 = (float)c1; q->k2 = (float)c2; q->sg = (float)sg; q->om = (float)om;
}
static inline float fm_bq_normal(const fm_biquad* q, float* w1, float* w2, float x) {
  const float y = fmaf(q->k2, *w2, fmaf(q->k1, *w1, q->k0 * x));
  const float n1 = fmaf(q->sg, *w1, fmaf(-q->om, *w2, x));
  const float n2 = fmaf(q->om, *w1, q->sg * *w2);
  *w1 = n1; *w2 = n2;
  return y;
}
static inline float fm_bq_df(const fm_biquad* q, float* x1, float* x2, float* y1, float* y2, float x) {
  float y = fmaf(q->b1, *x1, q->b0 * x);
  y = fmaf(q->b2, *x2, y);
  y = fmaf(-q->a2, *y2, y);
  y = fmaf(-q->a1, *y1, y);
  *x2 = *x1; *x1 = x; *y2 = *y1; *y1 = y;
  return y;
}

static void fm_reset_a2(fm_t* s) {
  s->lc = 1; s->ls = 0; s->iw1 = s->iw2 = s->qw1 = s->qw2 = s->ow1 = s->ow2 = 0;
  s->ix1 = s->ix2 = s->iy1 = s->iy2 = s->qx1 = s->qx2 = s->qy1 = s->qy2 = s->ox1 = s->ox2 = s->oy1 = s->oy2 = 0;
  s->psi = 1; s->psq = 0;  /* lastPhase = 0 */
  s->iacc = s->qacc = 0; s->dsc = 0; s->lo_age = 0;
  s->E = 0; s->rs_prev = 0; s->bc_left = 0;
}
static void fm_reset_state(fm_t* s) { /* fsk.ts:175-188 */
  fm_reset_a2(s);
  s->gsc = 0; s->bsc = 0; s->bit_acc = 0; s->bit_cnt = 0; s->next_idx = 0; s->current = 0; s->bitpos = 0;
  s->started = 0; s->sil_cnt = 0;
  s->d_ones = s->d_zeros = 0; s->amp_pending = 0; s->sil_extra = 0;
}

static fm_t* fm_new(const wamo_fsk_config* cfg) {
  wamo_fsk* m = wamo_fsk_new();
  wamo_fsk_configure(m, cfg);
  fm_t* s = (fm_t*)calloc(1, sizeof(*s));
  double b[3], a[3];
  wamo_iir_coefficients(m->preFilter, b, a);
  fm_design(&s->pre, b, a);
  wamo_iir_coefficients(m->iqI, b, a);
  fm_design(&s->lp, b, a);
  s->agc = m->has_agc; s->att = (float)m->agc.attackRate; s->rel = (float)m->agc.releaseRate;
  s->attd = m->agc.attackRate; s->reld = m->agc.releaseRate; s->gd = 1.0;
  const double omega = 2 * M_PI * m->centerFreq / cfg->sampleRate;
  s->cw = (float)cos(omega); s->sw = (float)sin(omega);
  /* the float32 LO turns by atan2(sw, cw) per sample instead of omega: a constant offset of the phase difference */
  s->dphi_bias = 2.0 * (atan2((double)s->sw, (double)s->cw) - omega);
  s->dspb = (int)m->downsampledSamplesPerBit; s->nbits = m->nbits; s->total_bits = s->nbits * s->dspb;
  s->check_period = (int)js_round(s->dspb / 4.0);
  s->parity = cfg->parity; s->stop_pos = cfg->parity == 0 ? 9 : 10;
  s->eod_count = (int)ceil(m->samplesForEOD);
  { /* smallest matched with matched / total > threshold */
    int mm = 0; const double total = (double)s->total_bits;
    while (mm <= s->total_bits && !((double)mm / total > cfg->syncThreshold)) mm++;
    s->min_matched = mm;
  }
  s->ring_cap = (int)(m->maxSyncBits * s->dspb * 1.1);
  s->amp_cap = s->dspb * 8;
  s->pattern = (int*)malloc(sizeof(int) * (size_t)(s->nbits + 1));
  for (int i = 0; i < s->nbits; i++) s->pattern[i] = m->preambleSfdBits[i];
  s->ring = (uint8_t*)calloc((size_t)s->ring_cap + 1, 1); s->dring = (uint8_t*)calloc((size_t)s->ring_cap + 1, 1);
  s->aring = (float*)calloc((size_t)s->amp_cap + 1, sizeof(float));
  /* |h_post(n)| <= C * r^n with r the pole radius of the low-pass (run at the decimated rate) */
  s->rho_e = sqrt(a[2]);
  s->gamma_e = sqrt((double)s->lp.k1 * s->lp.k1 + (double)s->lp.k2 * s->lp.k2) / s->rho_e;  /* |h(j)| <= gamma * rho^j */
  s->g = 1.0f; s->sil_thr = 0.01;
  fm_reset_state(s);
  wamo_fsk_free(m);
  return s;
}
static void fm_free(fm_t* s) { free(s->pattern); free(s->ring); free(s->dring); free(s->aring); free(s); }

typedef struct { uint8_t* out; long cap; fm_result* r; const fm_params* p; } fm_ctx;

static void fm_flag(fm_t* s, fm_ctx* c, int cause) {
  if (c->p->unguarded) return;
  if (!c->r->flag) { c->r->flag = 1; c->r->first_cause = cause; c->r->first_index = (int32_t)s->dec_index; }
  c->r->n_cause[cause]++;
}

/* FSKCore.processByte — fsk.ts:346-375; returns 1 when resetState() ran */
static int fm_process_byte(fm_t* s, fm_ctx* c, int bit) {
  const int bp = s->bitpos;
  if (bp == 0) {
    if (bit != 0) { fm_reset_state(s); return 1; }
  } else if (bp >= 1 && bp <= 8) {
    s->current |= bit << (8 - bp);
  } else if (s->parity != 0 && bp == 9) {
  } else if (bp == s->stop_pos) {
    if (bit != 1) { s->started = 0; return 0; }
    if (c->r->n_out < c->cap) c->out[c->r->n_out] = (uint8_t)s->current;
    c->r->n_out++;
    s->current = 0; s->bitpos = -1;
  } else { s->started = 0; return 0; }
  s->bitpos++;
  return 0;
}

/* FSKCore.processDownsampledBit — fsk.ts:278-344 — with doubt tracking.  dbit: the hard bit is doubtful;
 * amplitude compare bands: lo = certainly silent, hi = possibly silent. */
static int fm_decim(fm_t* s, fm_ctx* c, int bit, int dbit, float amp) {
  const fm_params* p = c->p;
  /* ring puts */
  const long pos = s->ring_pos % s->ring_cap;
  if (s->ring_len >= s->total_bits) {  /* the sample leaving the sync window */
    const long old = (s->ring_pos - s->total_bits) % s->ring_cap;
    s->d_win -= s->dring[old];
  }
  s->ring[pos] = (uint8_t)bit; s->dring[pos] = (uint8_t)dbit; s->d_win += dbit;
  s->ring_pos++; if (s->ring_len < s->ring_cap) s->ring_len++;
  s->aring[s->amp_pos % s->amp_cap] = amp; s->amp_pos++; if (s->amp_len < s->amp_cap) s->amp_len++;
  s->dec_index++;
  c->r->n_dec++; c->r->n_doubt_samples += dbit;

  s->gsc++;
  {
    const double a = (double)amp, thr = s->sil_thr;
    const int silent = a < thr;
    const int lo = a < thr * (1.0 - p->eps_amp), hi = a < thr * (1.0 + p->eps_amp);
    const int adoubt = !p->unguarded && (lo != hi);
    if (adoubt) {
      s->amp_pending = 1;
      if (!silent) s->sil_extra += s->sil_cnt + 1; else s->sil_extra += 0;
    } else if (!hi) { s->amp_pending = 0; s->sil_extra = 0; }
    if (silent) s->sil_cnt++; else s->sil_cnt = 0;
    if (s->amp_pending && s->sil_cnt + s->sil_extra >= (uint32_t)s->eod_count) {
      fm_flag(s, c, FM_CAUSE_EOD);
      s->amp_pending = 0; s->sil_extra = 0;  /* count it once */
    }
    if (silent && s->sil_cnt >= (uint32_t)s->eod_count) {
      c->r->eodEvents += 1;
      fm_reset_state(s);
      return 1;
    }
  }
  if (!s->started) {
    if (s->check_period > 0 && s->ring_len >= s->total_bits && s->gsc % (uint32_t)s->check_period == 0 && s->total_bits > 0) {
      int matched = 0, dn = 0;
      for (int j = 1; j < s->nbits; j++)
        for (int k = 0; k < s->dspb; k++) {
          const long q = (s->ring_pos - 1 - ((long)j * s->dspb + k)) % s->ring_cap;
          matched += (s->ring[q] == s->pattern[s->nbits - j]);
          dn += s->dring[q];
        }
      const int sync = matched >= s->min_matched;
      if (!p->unguarded && dn > 0 && ((matched - dn >= s->min_matched) != (matched + dn >= s->min_matched)))
        fm_flag(s, c, FM_CAUSE_SYNC);
      if (sync) {
        s->started = 1; s->current = 0; s->bitpos = 0;
        s->bit_acc = 0; s->bit_cnt = 0; s->bsc = 0; s->next_idx = 0; s->d_ones = s->d_zeros = 0;
        c->r->syncDetections += 1;
        double sum = 0;
        const long first = s->amp_pos - s->amp_len;
        for (long i = 0; i < s->amp_len; i++) sum += (double)s->aring[(first + i) % s->amp_cap];
        s->sil_thr = (sum / (double)s->amp_len) * 0.1;
      }
    }
    return 0;
  }
  s->bit_acc += (uint32_t)bit; s->bit_cnt++; s->bsc++;
  if (dbit) { if (bit) s->d_ones++; else s->d_zeros++; }
  if (s->bsc >= s->next_idx) {
    const int decided = 2u * s->bit_acc > s->bit_cnt;
    if (!p->unguarded && (s->d_ones | s->d_zeros)) {
      const int lo = 2u * (s->bit_acc - s->d_ones) > s->bit_cnt, hi = 2u * (s->bit_acc + s->d_zeros) > s->bit_cnt;
      if (lo != hi) {
        const int bp = s->bitpos;
        if (bp == 0) fm_flag(s, c, FM_CAUSE_VOTE_START);
        else if (bp == s->stop_pos) fm_flag(s, c, FM_CAUSE_VOTE_STOP);
        else if (bp >= 1 && bp <= 8) fm_flag(s, c, FM_CAUSE_VOTE_DATA);
        /* the parity bit is never looked at */
      }
    }
    s->bit_acc = 0; s->bit_cnt = 0; s->d_ones = s->d_zeros = 0;
    s->next_idx += (uint32_t)s->dspb;
    return fm_process_byte(s, c, decided);
  }
  return 0;
}

static void fm_run(fm_t* s, fm_ctx* c, const float* x, long n, const double* oF, const double* oA, long on) {
  const fm_params* p = c->p;
  double e2 = 0; long ne = 0;
  for (long i = 0; i < n; i++) {
    /* ---- A1: AGC (fsk.ts:52-76).  The gain recurrence stays in float64: its branch level > 0.5 is a discontinuity
     * (a float32 gain takes the other branch once in ~1e7 samples and then carries a 1e-3 gain error for hundreds of
     * samples), and an exact gain keeps the float32 store of fsk.ts:55 exact as well ---- */
    float sg = x[i];
    if (s->agc) {
      sg = (float)((double)x[i] * s->gd);
      const double level = fabs((double)sg);
      if (level > 0.0) {
        const double t = 0.5 / level;
        const double rate = level > 0.5 ? s->attd : s->reld;
        double g = s->gd + (t - s->gd) * rate;
        g = g > 10.0 ? 10.0 : g; g = g < 0.1 ? 0.1 : g;
        s->gd = g;
      }
    }
    float pf;
    if (p->form == 1) pf = fm_bq_normal(&s->pre, &s->pw1, &s->pw2, sg);
    else pf = fm_bq_df(&s->pre, &s->px1, &s->px2, &s->py1, &s->py2, sg);
    /* ---- A2 ---- */
    if (s->lo_age == 32) {  /* renormalise the rotation once per tile */
      const float m = fmaf(s->lc, s->lc, s->ls * s->ls);
      const float f = fmaf(-0.5f, m, 1.5f);
      s->lc *= f; s->ls *= f; s->lo_age = 0;
    }
    s->lo_age++;
    const float xi = pf * s->lc, xq = pf * s->ls;
    const float nc = fmaf(s->lc, s->cw, -(s->ls * s->sw)), nsn = fmaf(s->ls, s->cw, s->lc * s->sw);
    s->lc = nc; s->ls = nsn;
    float yi, yq;
    if (p->form == 1) { yi = fm_bq_normal(&s->lp, &s->iw1, &s->iw2, xi); yq = fm_bq_normal(&s->lp, &s->qw1, &s->qw2, xq); }
    else { yi = fm_bq_df(&s->lp, &s->ix1, &s->ix2, &s->iy1, &s->iy2, xi); yq = fm_bq_df(&s->lp, &s->qx1, &s->qx2, &s->qy1, &s->qy2, xq); }
    if (s->dsc == 0) { s->iacc = yi; s->qacc = yq; s->dsc = 1; continue; }
    const float si = s->iacc + yi, sq = s->qacc + yq;
    s->iacc = 0; s->qacc = 0; s->dsc = 0;
    /* phase difference straight from the two phasors: atan2(cross, dot) = wrapped (phase - lastPhase) */
    const float cross = fmaf(sq, s->psi, -(si * s->psq)), dot = fmaf(si, s->psi, sq * s->psq);
    float pd = atan2f(cross, dot) - (float)s->dphi_bias;
    const float pw = fmaf(si, si, sq * sq);
    const float rs = pw > 0 ? 1.0f / sqrtf(pw) : 0.0f;
    const float amp = 0.5f * pw * rs;
    s->psi = si; s->psq = sq;
    float fpd;
    if (p->form == 1) fpd = fm_bq_normal(&s->lp, &s->ow1, &s->ow2, pd);
    else fpd = fm_bq_df(&s->lp, &s->ox1, &s->ox2, &s->oy1, &s->oy2, pd);
    const int bit = fpd > 0.0f;
    /* ---- doubt band ---- */
    int dbit = 0;
    double band = 0;
    if (!p->unguarded) {
      s->S = fmaxf(amp, s->S * (1.0f - 1.0f / 128.0f));
      /* phase error of this phasor ~ kappa * S / (2 amp); the difference carries this one and the previous one */
      const float hrs = 0.5f * rs;  /* 1 / (2 amp) = 1 / |(si, sq)| */
      const float ephi = (float)p->kappa * s->S * (hrs + s->rs_prev);
      s->rs_prev = hrs;
      float et = pw > 0 ? ephi : 10.0f;
      if (fabsf(fabsf(pd) - (float)M_PI) < (float)p->bc_delta + 4.0f * ephi) { c->r->n_bc++; et += 6.3f; }  /* may have wrapped the other way */
      s->E = fmaf((float)s->rho_e, s->E, (float)s->gamma_e * et);
      band = (double)s->E + p->eps0;
      dbit = fabs((double)fpd) < band;
    }
    if (oF && s->dec_index < on) {
      const double err = fabs((double)fpd - oF[s->dec_index]);
      if (err > c->r->err_max) c->r->err_max = err;
      e2 += err * err; ne++;
      if (band > 0 && err / band > c->r->ratio_max) {
        c->r->ratio_max = err / band;
        c->r->w_index = (double)s->dec_index; c->r->w_amp = amp; c->r->w_S = s->S; c->r->w_err = err; c->r->w_band = band;
        c->r->w_pd = pd; c->r->w_oamp = oA[s->dec_index]; c->r->w_F = oF[s->dec_index];
      }
      const int obit = oF[s->dec_index] > 0;
      if (obit != bit) { c->r->n_wrong_bits++; if (!dbit) c->r->n_wrong_undoubted++; }
    }
    if (fm_decim(s, c, bit, dbit, amp)) fm_reset_a2(s);
  }
  c->r->err_rms = ne ? sqrt(e2 / (double)ne) : 0;
  c->r->started = s->started; c->r->gsc = s->gsc; c->r->sil_thr = s->sil_thr;
}

typedef struct {
  const wamo_fsk_config* cfgs; const int32_t* cfg_index; long n_streams; const float* samples; long stride, n;
  const fm_params* p; uint8_t* out; long out_stride; fm_result* res; int compare; int tid, n_threads;
} fm_job;

static void* fm_worker(void* arg) {
  fm_job* j = (fm_job*)arg;
  double *oF = NULL, *oA = NULL; float* tmp = NULL; uint8_t* obytes = NULL;
  const long ndec = j->n / 2 + 2;
  if (j->compare) {
    oF = (double*)malloc(sizeof(double) * (size_t)ndec); oA = (double*)malloc(sizeof(double) * (size_t)ndec);
    tmp = (float*)malloc(sizeof(float) * (size_t)(j->n > 0 ? j->n : 1)); obytes = (uint8_t*)malloc(65536);
  }
  for (long s = j->tid; s < j->n_streams; s += j->n_threads) {
    const wamo_fsk_config* cfg = &j->cfgs[j->cfg_index ? j->cfg_index[s] : 0];
    long on = 0;
    if (j->compare) {
      wamo_fsk* m = wamo_fsk_new();
      wamo_fsk_configure(m, cfg);
      wamo_fsk_set_decim_tap(m, oF, oA, ndec);
      memcpy(tmp, j->samples + s * j->stride, sizeof(float) * (size_t)j->n);
      wamo_fsk_demodulate(m, tmp, j->n, obytes, 65536);
      on = wamo_fsk_decim_tap_count(m);
      wamo_fsk_free(m);
    }
    fm_t* f = fm_new(cfg);
    fm_result* r = &j->res[s];
    memset(r, 0, sizeof(*r));
    fm_ctx c = {j->out + s * j->out_stride, j->out_stride, r, j->p};
    fm_run(f, &c, j->samples + s * j->stride, j->n, j->compare ? oF : NULL, oA, on);
    fm_free(f);
  }
  free(oF); free(oA); free(tmp); free(obytes);
  return NULL;
}

/* Runs the model over n_streams fresh streams.  compare != 0: also run the oracle on every stream and fill the
 * error statistics of fm_result (the comparison of filteredPhaseDiff is only meaningful while both state machines
 * made the same resets: look at streams whose bytes and counters agree). */
int fm_batch(const wamo_fsk_config* cfgs, const int32_t* cfg_index, long n_streams, const float* samples, long stride,
             long n, const fm_params* p, uint8_t* out, long out_stride, fm_result* res, int compare, int n_threads) {
  if (n_threads < 1) n_threads = 1;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)n_threads);
  fm_job* jobs = (fm_job*)malloc(sizeof(fm_job) * (size_t)n_threads);
  for (int t = 0; t < n_threads; t++) {
    jobs[t] = (fm_job){cfgs, cfg_index, n_streams, samples, stride, n, p, out, out_stride, res, compare, t, n_threads};
    pthread_create(&th[t], NULL, fm_worker, &jobs[t]);
  }
  for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
  free(th); free(jobs);
  return 0;
}
