/*
 * wam_oracle.c — CPU ORACLE (test infrastructure, NOT product code).  See wam_oracle.h.
 *
 * Every function cites the reference file:line it restates.  Arithmetic is IEEE float64
 * with JS semantics (no FMA contraction: build with -ffp-contract=off), float32 only at the
 * reference's Float32Array stores.
 */
#include "wam_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * JS helpers
 * ---------------------------------------------------------------------------------------- */
static double js_max(double a, double b) { /* Math.max: NaN-propagating */
  if (isnan(a) || isnan(b)) return NAN;
  return a > b ? a : b;
}
static double js_min(double a, double b) {
  if (isnan(a) || isnan(b)) return NAN;
  return a < b ? a : b;
}
static double js_round(double x) { return floor(x + 0.5); } /* Math.round */
static int is_integral(double x) { return x == floor(x) && isfinite(x); }

/* ------------------------------------------------------------------------------------------
 * RingBuffer — src/utils.ts:6-105
 * The constructor takes `size` as a JS number; `new ArrayType(size)` truncates it (ToIndex)
 * while maxLength keeps the fraction (utils.ts:14-18).  Typed-array element access with a
 * non-integral or out-of-range index reads `undefined` / ignores the write.
 * ---------------------------------------------------------------------------------------- */
struct wamo_ring {
  int kind;          /* 0 Uint8Array, 1 Float32Array */
  double* buf;       /* element values after the typed-array store conversion */
  long buflen;       /* trunc(size) */
  double readIndex, writeIndex, length, maxLength;
};

wamo_ring* wamo_ring_new(int kind, double size) {
  wamo_ring* r = (wamo_ring*)calloc(1, sizeof(*r));
  r->kind = kind;
  r->buflen = (long)trunc(size);
  if (r->buflen < 0) r->buflen = 0;
  r->buf = (double*)calloc((size_t)(r->buflen > 0 ? r->buflen : 1), sizeof(double));
  r->maxLength = size;
  return r;
}
void wamo_ring_free(wamo_ring* r) {
  if (!r) return;
  free(r->buf);
  free(r);
}
static double ring_store_convert(int kind, double v) {
  if (kind == 1) return (double)(float)v; /* Float32Array store */
  /* Uint8Array store: ToUint8 (modular); only 0/1 are ever stored by FSKCore */
  if (!isfinite(v)) return 0.0;
  double t = trunc(v);
  double m = fmod(t, 256.0);
  if (m < 0) m += 256.0;
  return m;
}
void wamo_ring_put(wamo_ring* r, double v) { /* utils.ts:37-47 */
  if (is_integral(r->writeIndex) && r->writeIndex >= 0 && r->writeIndex < (double)r->buflen)
    r->buf[(long)r->writeIndex] = ring_store_convert(r->kind, v);
  r->writeIndex = fmod(r->writeIndex + 1.0, r->maxLength);
  if (r->length < r->maxLength) {
    r->length += 1.0;
  } else {
    r->readIndex = fmod(r->readIndex + 1.0, r->maxLength);
  }
}
int wamo_ring_get(const wamo_ring* r, double index, double* v) { /* utils.ts:28-35 */
  if (index < 0) index += r->length;
  if (index < 0 || index >= r->length) return -1;
  double p = fmod(r->readIndex + index, r->maxLength);
  if (is_integral(p) && p >= 0 && p < (double)r->buflen) {
    *v = r->buf[(long)p];
    return 0;
  }
  return 1; /* undefined */
}
double wamo_ring_length(const wamo_ring* r) { return r->length; }
void wamo_ring_clear(wamo_ring* r) { /* utils.ts:93-97 — buffer contents are kept */
  r->readIndex = 0;
  r->writeIndex = 0;
  r->length = 0;
}

/* ------------------------------------------------------------------------------------------
 * IIRFilter — src/dsp/filters.ts:8-106
 * ---------------------------------------------------------------------------------------- */
struct wamo_iir {
  double *b, *a, *x, *y;
  int nb, na, nx, ny, xIndex, yIndex, order;
};

wamo_iir* wamo_iir_new(const double* b, int nb, const double* a, int na, int* err) {
  if (err) *err = 0;
  if (!b || nb <= 0) { if (err) *err = 1; return NULL; } /* filters.ts:19 */
  if (!a || na <= 0) { if (err) *err = 2; return NULL; } /* filters.ts:20 */
  if (a[0] == 0) { if (err) *err = 3; return NULL; }     /* filters.ts:21 */
  wamo_iir* f = (wamo_iir*)calloc(1, sizeof(*f));
  f->nb = nb; f->na = na;
  f->b = (double*)malloc(sizeof(double) * (size_t)nb);
  f->a = (double*)malloc(sizeof(double) * (size_t)na);
  memcpy(f->b, b, sizeof(double) * (size_t)nb);
  memcpy(f->a, a, sizeof(double) * (size_t)na);
  f->order = (nb > na ? nb : na) - 1;            /* filters.ts:27 */
  if (f->a[0] != 1) {                            /* filters.ts:30-39 */
    double a0 = f->a[0];
    for (int i = 0; i < nb; i++) f->b[i] /= a0;
    for (int i = 1; i < na; i++) f->a[i] /= a0;
    f->a[0] = 1;
  }
  f->nx = nb > f->order + 1 ? nb : f->order + 1; /* filters.ts:94 */
  f->ny = (na - 1) > f->order ? (na - 1) : f->order; /* filters.ts:95 */
  f->x = (double*)calloc((size_t)(f->nx > 0 ? f->nx : 1), sizeof(double));
  f->y = (double*)calloc((size_t)(f->ny > 0 ? f->ny : 1), sizeof(double));
  return f;
}
void wamo_iir_free(wamo_iir* f) {
  if (!f) return;
  free(f->b); free(f->a); free(f->x); free(f->y); free(f);
}
void wamo_iir_reset(wamo_iir* f) { /* filters.ts:92-98 */
  memset(f->x, 0, sizeof(double) * (size_t)f->nx);
  memset(f->y, 0, sizeof(double) * (size_t)f->ny);
  f->xIndex = 0; f->yIndex = 0;
}
double wamo_iir_process(wamo_iir* f, double input) { /* filters.ts:47-76 */
  f->x[f->xIndex] = input;
  double output = 0;
  int xIdx = f->xIndex;
  for (int i = 0; i < f->nb; i++) {
    output += f->b[i] * f->x[xIdx];
    xIdx = xIdx == 0 ? f->nx - 1 : xIdx - 1;
  }
  if (f->ny > 0) {
    int yIdx = f->yIndex == 0 ? f->ny - 1 : f->yIndex - 1;
    for (int i = 1; i < f->na; i++) {
      output -= f->a[i] * f->y[yIdx];
      yIdx = yIdx == 0 ? f->ny - 1 : yIdx - 1;
    }
    f->y[f->yIndex] = output;
    f->yIndex = (f->yIndex + 1) % f->ny;
  }
  f->xIndex = (f->xIndex + 1) % f->nx;
  return output;
}
void wamo_iir_process_buffer(wamo_iir* f, const float* in, float* out, long n) { /* filters.ts:81-87 */
  for (long i = 0; i < n; i++) out[i] = (float)wamo_iir_process(f, (double)in[i]);
}
int wamo_iir_coefficients(const wamo_iir* f, double* b, double* a) {
  memcpy(b, f->b, sizeof(double) * (size_t)f->nb);
  memcpy(a, f->a, sizeof(double) * (size_t)f->na);
  return f->nb | (f->na << 16);
}

/* ------------------------------------------------------------------------------------------
 * FIRFilter — src/dsp/filters.ts:112-167
 * ---------------------------------------------------------------------------------------- */
struct wamo_fir {
  double *c, *d;
  int n, index;
};
wamo_fir* wamo_fir_new(const double* taps, int n) {
  wamo_fir* f = (wamo_fir*)calloc(1, sizeof(*f));
  f->n = n;
  f->c = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
  f->d = (double*)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
  if (n > 0) memcpy(f->c, taps, sizeof(double) * (size_t)n);
  return f;
}
void wamo_fir_free(wamo_fir* f) {
  if (!f) return;
  free(f->c); free(f->d); free(f);
}
double wamo_fir_process(wamo_fir* f, double input) { /* filters.ts:125-140 */
  if (f->n == 0) return 0;
  f->d[f->index] = input;
  double output = 0;
  int di = f->index;
  for (int i = 0; i < f->n; i++) {
    output += f->c[i] * f->d[di];
    di = di == 0 ? f->n - 1 : di - 1;
  }
  f->index = (f->index + 1) % f->n;
  return output;
}
void wamo_fir_process_buffer(wamo_fir* f, const float* in, float* out, long n) { /* filters.ts:145-151 */
  for (long i = 0; i < n; i++) out[i] = (float)wamo_fir_process(f, (double)in[i]);
}
void wamo_fir_reset(wamo_fir* f) { /* filters.ts:156-159 */
  memset(f->d, 0, sizeof(double) * (size_t)(f->n > 0 ? f->n : 1));
  f->index = 0;
}

/* ------------------------------------------------------------------------------------------
 * FilterDesign — src/dsp/filters.ts:172-315
 * ---------------------------------------------------------------------------------------- */
void wamo_design_butterworth_lowpass(double fc, double fs, double b[3], double a[3]) { /* :180-192 */
  double nyquist = fs / 2;
  double normalizedCutoff = fc / nyquist;
  double c = tan(M_PI * normalizedCutoff / 2);
  double c2 = c * c;
  double sqrt2c = M_SQRT2 * c;
  double denom = 1 + sqrt2c + c2;
  b[0] = c2 / denom; b[1] = 2 * c2 / denom; b[2] = c2 / denom;
  a[0] = 1; a[1] = (2 * c2 - 2) / denom; a[2] = (1 - sqrt2c + c2) / denom;
}
void wamo_design_butterworth_highpass(double fc, double fs, double b[3], double a[3]) { /* :200-212 */
  double nyquist = fs / 2;
  double normalizedCutoff = fc / nyquist;
  double c = tan(M_PI * normalizedCutoff / 2);
  double c2 = c * c;
  double sqrt2c = M_SQRT2 * c;
  double denom = 1 + sqrt2c + c2;
  b[0] = 1 / denom; b[1] = -2 / denom; b[2] = 1 / denom;
  a[0] = 1; a[1] = (2 * c2 - 2) / denom; a[2] = (1 - sqrt2c + c2) / denom;
}
void wamo_design_butterworth_bandpass(double f0, double bandwidth, double fs, double b[3], double a[3]) { /* :221-234 */
  double omega = 2 * M_PI * f0 / fs;
  double bw = 2 * M_PI * bandwidth / fs;
  double c = tan(bw / 2);
  double d = 2 * cos(omega);
  double c2 = c * c;
  double denom = 1 + c + c2;
  b[0] = c / denom; b[1] = 0; b[2] = -c / denom;
  a[0] = 1; a[1] = (-d * (1 + c2)) / denom; a[2] = (1 - c + c2) / denom;
}
int wamo_design_sinc_lowpass(double fc, double fs, int numTaps, double* out) { /* :243-265 */
  if (numTaps % 2 == 0) numTaps++;
  double normalizedCutoff = fc / fs;
  double center = (numTaps - 1) / 2.0;
  for (int i = 0; i < numTaps; i++) {
    if ((double)i == center) {
      out[i] = 2 * normalizedCutoff;
    } else {
      double x = M_PI * (i - center);
      out[i] = sin(2 * normalizedCutoff * x) / x;
    }
    out[i] *= 0.54 - 0.46 * cos(2 * M_PI * i / (numTaps - 1));
  }
  return numTaps;
}
/* NOTE (filters.ts:274-286): sincHighpass passes its *own* numTaps to sincLowpass, which
 * increments a private copy when even; the spectral-inversion loop then runs over the
 * caller's (even) numTaps and `center` is fractional, so `lowpass[center] += 1` creates a
 * non-index property and the array elements are unchanged except for negation of the first
 * numTaps entries.  Restated literally. */
int wamo_design_sinc_highpass(double fc, double fs, int numTaps, double* out) { /* :274-286 */
  int n = wamo_design_sinc_lowpass(fc, fs, numTaps, out);
  double center = (numTaps - 1) / 2.0;
  for (int i = 0; i < numTaps; i++) out[i] = -out[i];
  if (is_integral(center) && center >= 0 && center < n) out[(int)center] += 1;
  return n;
}
/* NOTE (filters.ts:296-314): bandpass has exactly numTaps entries (new Array(numTaps)); the
 * i/j loops run to numTaps over arrays that may have numTaps+1 entries (even request). */
int wamo_design_sinc_bandpass(double f0, double bandwidth, double fs, int numTaps, double* out) {
  double lowFreq = f0 - bandwidth / 2;
  double highFreq = f0 + bandwidth / 2;
  double* hp = (double*)malloc(sizeof(double) * (size_t)(numTaps + 2));
  double* lp = (double*)malloc(sizeof(double) * (size_t)(numTaps + 2));
  wamo_design_sinc_highpass(lowFreq, fs, numTaps, hp);
  wamo_design_sinc_lowpass(highFreq, fs, numTaps, lp);
  for (int i = 0; i < numTaps; i++) out[i] = 0;
  for (int i = 0; i < numTaps; i++)
    for (int j = 0; j < numTaps; j++)
      if (i + j < numTaps) out[i + j] += hp[i] * lp[j];
  free(hp); free(lp);
  return numTaps;
}

/* ------------------------------------------------------------------------------------------
 * CRC16 — src/utils/crc16.ts:21-38 ; XModemPacket — src/transports/xmodem/packet.ts:21-54
 * ---------------------------------------------------------------------------------------- */
uint16_t wamo_crc16(const uint8_t* data, long n) {
  uint32_t crc = 0xFFFF;
  for (long k = 0; k < n; k++) {
    crc ^= ((uint32_t)data[k] << 8);
    for (int i = 0; i < 8; i++) {
      if (crc & 0x8000) crc = (crc << 1) ^ 0x1021;
      else crc <<= 1;
      crc &= 0xFFFF;
    }
  }
  return (uint16_t)(crc ^ 0x0000);
}
long wamo_xmodem_serialize(int sequence, const uint8_t* payload, long n, uint8_t* out, long cap) {
  if (sequence < 1 || sequence > 255) return -1; /* packet.ts:22-24 */
  if (n > 255) return -2;                        /* packet.ts:25-27 */
  long total = 4 + n + 2;
  if (!out) return total;
  if (cap < total) return -3;
  uint16_t crc = wamo_crc16(payload, n);
  out[0] = 0x01;
  out[1] = (uint8_t)sequence;
  out[2] = (uint8_t)((~sequence) & 0xFF);
  out[3] = (uint8_t)n;
  memcpy(out + 4, payload, (size_t)n);
  out[4 + n] = (uint8_t)((crc >> 8) & 0xFF);
  out[4 + n + 1] = (uint8_t)(crc & 0xFF);
  return total;
}
void wamo_xmodem_check(const uint8_t* bytes, long n, int expectedSequence, wamo_pkt_result* res) {
  memset(res, 0, sizeof(*res));
  res->payloadOffset = -1;
  res->sequence = -1; res->length = -1; res->crcReceived = -1; res->crcComputed = -1;
  long p = 0;
  /* receiveAllPackets, xmodem.ts:236-252: ignore anything that is neither EOT nor SOH */
  for (;;) {
    if (p >= n) { res->status = WAMO_PKT_NO_SOH; res->bytesConsumed = (int32_t)p; return; }
    uint8_t first = bytes[p++];
    if (first == 0x04) { res->status = WAMO_PKT_EOT; res->bytesConsumed = (int32_t)p; return; }
    if (first == 0x01) break;
  }
  /* receiveAndProcessPacket, xmodem.ts:265-321 */
  if (p + 3 > n) { res->status = WAMO_PKT_INCOMPLETE; res->bytesConsumed = (int32_t)p; return; }
  int seq = bytes[p], nseq = bytes[p + 1], len = bytes[p + 2];
  p += 3;
  res->sequence = seq; res->length = len;
  if (seq + nseq != 255) { res->status = WAMO_PKT_BAD_COMPLEMENT; res->bytesConsumed = (int32_t)p; return; }
  int prevSeq = expectedSequence == 1 ? 255 : expectedSequence - 1; /* xmodem.ts:525-530 */
  if (seq == expectedSequence) {
    if (p + len + 2 > n) { res->status = WAMO_PKT_INCOMPLETE; res->bytesConsumed = (int32_t)p; return; }
    res->payloadOffset = (int32_t)p;
    int crc = (bytes[p + len] << 8) | bytes[p + len + 1];
    int calc = wamo_crc16(bytes + p, len);
    res->crcReceived = crc; res->crcComputed = calc;
    p += len + 2;
    res->bytesConsumed = (int32_t)p;
    res->status = (calc != crc) ? WAMO_PKT_BAD_CRC : WAMO_PKT_OK;
  } else if (seq == prevSeq) {
    if (p + len + 2 > n) { res->status = WAMO_PKT_INCOMPLETE; res->bytesConsumed = (int32_t)p; return; }
    res->payloadOffset = (int32_t)p;
    p += len + 2;
    res->bytesConsumed = (int32_t)p;
    res->status = WAMO_PKT_DUPLICATE;
  } else {
    res->bytesConsumed = (int32_t)p;
    res->status = WAMO_PKT_UNEXPECTED_SEQ;
  }
}

/* XModemTransport.receiveAllPackets + receiveAndProcessPacket over one burst — xmodem.ts:232-321 */
long wamo_xmodem_receive(const uint8_t* bytes, long n, int maxRetries, wamo_xmodem_rx_state* st,
                         uint8_t* replies, int reply_cap, int32_t* n_replies, uint8_t* data, long data_cap) {
  long p = 0;
  int nrep = 0;
#define WAMO_REPLY(c) do { if (nrep < reply_cap) replies[nrep] = (uint8_t)(c); nrep++; } while (0)
  while (!st->done) {
    if (p >= n) break;                      /* waitForByte would block (xmodem.ts:239) */
    const uint8_t first = bytes[p];
    if (first == 0x04) {                    /* EOT: final ACK (xmodem.ts:241-244) */
      p++;
      WAMO_REPLY(0x06);
      st->done = 1;
      break;
    }
    if (first != 0x01) { p++; continue; }   /* ignored byte (xmodem.ts:248-250) */
    /* receiveAndProcessPacket (xmodem.ts:265-321) */
    if (p + 4 > n) break;                   /* header not complete yet: keep the SOH */
    const int seq = bytes[p + 1], nseq = bytes[p + 2], len = bytes[p + 3];
    int error = 0;
    if (seq + nseq != 255) {
      st->packetsDropped++;
      error = 1;
    } else {
      const int prevSeq = st->expectedSequence == 1 ? 255 : st->expectedSequence - 1; /* xmodem.ts:525-530 */
      if (seq == st->expectedSequence) {
        if (p + 4 + len + 2 > n) break;     /* payload not complete yet */
        st->packetsReceived++;
        const int crc = (bytes[p + 4 + len] << 8) | bytes[p + 4 + len + 1];
        if ((int)wamo_crc16(bytes + p + 4, len) != crc) {
          st->packetsDropped++;
          error = 1;
        } else {
          for (int i = 0; i < len; i++)
            if ((long)st->dataLen + i < data_cap) data[st->dataLen + i] = bytes[p + 4 + i];
          st->dataLen += len;
          st->expectedSequence = (st->expectedSequence % 255) + 1;
          st->retries = 0;
          WAMO_REPLY(0x06);
          p += 4 + len + 2;
        }
      } else if (seq == prevSeq) {
        if (p + 4 + len + 2 > n) break;
        st->packetsDropped++;
        WAMO_REPLY(0x06);                   /* duplicate: consumed, ACKed, ignored (xmodem.ts:309-314) */
        p += 4 + len + 2;
      } else {
        st->packetsDropped++;
        error = 1;
      }
    }
    if (error) {                            /* catch block, xmodem.ts:251-260 */
      if (++st->retries > maxRetries) { st->done = 2; p = n; break; }
      p = n;                                /* receive.buffer = [] */
      WAMO_REPLY(0x15);
    }
  }
#undef WAMO_REPLY
  *n_replies = nrep;
  return p;
}

/* ------------------------------------------------------------------------------------------
 * AGCProcessor — src/modems/fsk.ts:38-77
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  double targetLevel, currentGain, attackRate, releaseRate;
} agc_t;
static void agc_init(agc_t* g, double sampleRate) { /* fsk.ts:44-50 */
  g->targetLevel = 0.5;
  g->currentGain = 1.0;
  g->attackRate = 1.0 - exp(-1.0 / (sampleRate * 0.001));
  g->releaseRate = 1.0 - exp(-1.0 / (sampleRate * 0.01));
}
static void agc_process(agc_t* g, float* samples, long n) { /* fsk.ts:52-76 */
  for (long i = 0; i < n; i++) {
    samples[i] = (float)((double)samples[i] * g->currentGain); /* f32 store, fsk.ts:55 */
    double outputLevel = fabs((double)samples[i]);
    if (outputLevel > g->targetLevel) {
      double targetGain = g->targetLevel / outputLevel;
      g->currentGain += (targetGain - g->currentGain) * g->attackRate;
    } else {
      if (outputLevel > 0) {
        double targetGain = g->targetLevel / outputLevel;
        g->currentGain += (targetGain - g->currentGain) * g->releaseRate;
      }
    }
    g->currentGain = js_max(0.1, js_min(10.0, g->currentGain));
  }
}

/* ------------------------------------------------------------------------------------------
 * FSKCore — src/modems/fsk.ts:82-494
 * ---------------------------------------------------------------------------------------- */
struct wamo_fsk {
  /* config (fsk.ts:134) */
  wamo_fsk_config cfg;
  uint8_t* preamble; uint8_t* sfd;
  int ready;
  /* dsp (fsk.ts:87-92) */
  int has_agc; agc_t agc;
  wamo_iir *preFilter, *iqI, *iqQ, *postFilter;
  /* params (fsk.ts:95-99) */
  double samplesPerBit, bitsPerByte, centerFreq, downsampleRatio, downsampledSamplesPerBit;
  /* iqState, downsample (fsk.ts:102-109) */
  double localOscPhase, lastPhase;
  double dsCounter, iAcc, qAcc;
  /* bitSync (fsk.ts:112-115) */
  double globalSampleCounter, bitSampleCounter, bitAccumulator, bitAccumCount, nextBitSampleIndex;
  /* frame (fsk.ts:118-122) */
  int* preambleSfdBits; int nbits; double maxSyncBits; int started;
  wamo_ring *syncSamples, *syncAmplitude;
  /* byteState (fsk.ts:125) */
  int current, bitPosition;
  uint8_t* bytebuf; long nbytebuf, capbytebuf;
  /* silence (fsk.ts:128) */
  double silenceThreshold, samplesForEOD, silenceCount;
  /* debug (fsk.ts:131) */
  double syncDetections, demodulationCalls, totalSamples;
  /* events */
  double eodEvents, errorEvents, configuredEvents;
  int threw;
  float* tap; long tapcap;
  /* optional decimated-rate tap (tests / fast-path model): filteredPhaseDiff and amplitude per decimated sample */
  double* dtapF; double* dtapA; long dtapcap, dtapn;
};

void wamo_default_config(wamo_fsk_config* c) { /* fsk.ts:19-33 */
  static const uint8_t pre[2] = {0x55, 0x55};
  static const uint8_t sfd[1] = {0x7E};
  memset(c, 0, sizeof(*c));
  c->sampleRate = 48000; c->baudRate = 1200;
  c->markFrequency = 1650; c->spaceFrequency = 1850;
  c->preamblePattern = pre; c->preambleLength = 2;
  c->sfdPattern = sfd; c->sfdLength = 1;
  c->startBits = 1; c->stopBits = 1; c->parity = 0;
  c->syncThreshold = 0.85; c->agcEnabled = 1;
  c->preFilterBandwidth = 800; c->adaptiveThreshold = 1;
}

wamo_fsk* wamo_fsk_new(void) {
  wamo_fsk* m = (wamo_fsk*)calloc(1, sizeof(*m));
  m->silenceThreshold = 0.01; /* fsk.ts:128 */
  return m;
}
void wamo_fsk_free(wamo_fsk* m) {
  if (!m) return;
  wamo_iir_free(m->preFilter); wamo_iir_free(m->iqI); wamo_iir_free(m->iqQ); wamo_iir_free(m->postFilter);
  wamo_ring_free(m->syncSamples); wamo_ring_free(m->syncAmplitude);
  free(m->preambleSfdBits); free(m->bytebuf); free(m->preamble); free(m->sfd);
  free(m);
}

static void push_bit(wamo_fsk* m, int bit, int* capbits) {
  if (m->nbits >= *capbits) {
    *capbits = *capbits ? *capbits * 2 : 64;
    m->preambleSfdBits = (int*)realloc(m->preambleSfdBits, sizeof(int) * (size_t)*capbits);
  }
  m->preambleSfdBits[m->nbits++] = bit;
}
static void add_byte_to_pattern(wamo_fsk* m, int byte, int* capbits) { /* fsk.ts:159-173 */
  for (int i = 0; i < m->cfg.startBits; i++) push_bit(m, 0, capbits);
  for (int i = 7; i >= 0; i--) push_bit(m, (byte >> i) & 1, capbits);
  if (m->cfg.parity != 0) {
    int parity = 0;
    for (int i = 0; i < 8; i++) parity ^= (byte >> i) & 1;
    push_bit(m, m->cfg.parity == 1 ? parity : 1 - parity, capbits);
  }
  for (int i = 0; i < m->cfg.stopBits; i++) push_bit(m, 1, capbits);
}

static void reset_state(wamo_fsk* m) { /* fsk.ts:175-188 */
  m->localOscPhase = 0; m->lastPhase = 0;
  m->globalSampleCounter = 0; m->bitSampleCounter = 0; m->bitAccumulator = 0;
  m->bitAccumCount = 0; m->nextBitSampleIndex = 0;
  m->current = 0; m->bitPosition = 0;
  m->started = 0;
  m->silenceCount = 0;
  if (m->iqI) wamo_iir_reset(m->iqI);
  if (m->iqQ) wamo_iir_reset(m->iqQ);
  if (m->postFilter) wamo_iir_reset(m->postFilter);
  m->dsCounter = 0; m->iAcc = 0; m->qAcc = 0;
}

void wamo_fsk_configure(wamo_fsk* m, const wamo_fsk_config* cfg) { /* fsk.ts:133-157 */
  m->cfg = *cfg;
  free(m->preamble); free(m->sfd);
  m->preamble = (uint8_t*)malloc((size_t)(cfg->preambleLength > 0 ? cfg->preambleLength : 1));
  m->sfd = (uint8_t*)malloc((size_t)(cfg->sfdLength > 0 ? cfg->sfdLength : 1));
  if (cfg->preambleLength > 0) memcpy(m->preamble, cfg->preamblePattern, (size_t)cfg->preambleLength);
  if (cfg->sfdLength > 0) memcpy(m->sfd, cfg->sfdPattern, (size_t)cfg->sfdLength);
  m->cfg.preamblePattern = m->preamble; m->cfg.sfdPattern = m->sfd;

  /* calculateParameters, fsk.ts:426-444 */
  double downsampleRatio = 2;
  double downsampleRate = cfg->sampleRate / downsampleRatio;
  m->centerFreq = (cfg->markFrequency + cfg->spaceFrequency) / 2;
  m->samplesPerBit = floor(cfg->sampleRate / cfg->baudRate);
  m->bitsPerByte = 8 + cfg->startBits + cfg->stopBits + (cfg->parity != 0 ? 1 : 0);
  m->downsampleRatio = downsampleRatio;
  m->downsampledSamplesPerBit = floor(downsampleRate / cfg->baudRate);

  /* initializeDSP, fsk.ts:446-462.  NB: when agcEnabled is false an AGC left over from an
   * earlier configure() stays installed (this.dsp.agc is only ever assigned, never cleared). */
  if (cfg->agcEnabled) { m->has_agc = 1; agc_init(&m->agc, cfg->sampleRate); }
  double freqSpan = fabs(cfg->spaceFrequency - cfg->markFrequency);
  double deviation = freqSpan / 2;
  double carsonBandwidth = 2 * (deviation + cfg->baudRate);
  double finalBandwidth = js_max(cfg->preFilterBandwidth, carsonBandwidth);
  double b[3], a[3];
  wamo_iir_free(m->preFilter); wamo_iir_free(m->iqI); wamo_iir_free(m->iqQ); wamo_iir_free(m->postFilter);
  wamo_design_butterworth_bandpass(m->centerFreq, finalBandwidth, cfg->sampleRate, b, a);
  m->preFilter = wamo_iir_new(b, 3, a, 3, NULL);
  wamo_design_butterworth_lowpass(cfg->baudRate, cfg->sampleRate, b, a);
  m->iqI = wamo_iir_new(b, 3, a, 3, NULL);
  m->iqQ = wamo_iir_new(b, 3, a, 3, NULL);
  m->postFilter = wamo_iir_new(b, 3, a, 3, NULL);

  /* frame detection, fsk.ts:143-150 */
  m->nbits = 0;
  int capbits = 0;
  free(m->preambleSfdBits); m->preambleSfdBits = NULL;
  for (int i = 0; i < cfg->preambleLength; i++) add_byte_to_pattern(m, m->preamble[i], &capbits);
  for (int i = 0; i < cfg->sfdLength; i++) add_byte_to_pattern(m, m->sfd[i], &capbits);
  m->maxSyncBits = m->nbits + 32;
  m->samplesForEOD = m->bitsPerByte * m->downsampledSamplesPerBit * 0.7;
  wamo_ring_free(m->syncSamples); wamo_ring_free(m->syncAmplitude);
  m->syncSamples = wamo_ring_new(0, m->maxSyncBits * m->downsampledSamplesPerBit * 1.1);
  m->syncAmplitude = wamo_ring_new(1, m->downsampledSamplesPerBit * 8);

  reset_state(m);
  m->ready = 1;
  m->configuredEvents += 1;
}

static void bytebuf_push(wamo_fsk* m, int v) {
  if (m->nbytebuf >= m->capbytebuf) {
    m->capbytebuf = m->capbytebuf ? m->capbytebuf * 2 : 256;
    m->bytebuf = (uint8_t*)realloc(m->bytebuf, (size_t)m->capbytebuf);
  }
  m->bytebuf[m->nbytebuf++] = (uint8_t)v;
}

static void process_byte(wamo_fsk* m, int bit) { /* fsk.ts:346-375 */
  int bitPosition = m->bitPosition;
  int stopBitPosition = m->cfg.parity == 0 ? 9 : 10;
  if (bitPosition == 0) {
    if (bit != 0) { reset_state(m); return; }
  } else if (bitPosition >= 1 && bitPosition <= 8) {
    m->current |= (bit << (8 - bitPosition));
  } else if (m->cfg.parity != 0 && bitPosition == 9) {
    /* parity bit: skipped, never checked */
  } else if (bitPosition == stopBitPosition) {
    if (bit != 1) { m->started = 0; return; }
    bytebuf_push(m, m->current);
    m->current = 0; m->bitPosition = -1;
  } else {
    m->started = 0;
    return;
  }
  m->bitPosition++;
}

static int process_downsampled_bit(wamo_fsk* m, int bitValue, double amplitude) { /* fsk.ts:278-344 */
  wamo_ring_put(m->syncSamples, bitValue);
  wamo_ring_put(m->syncAmplitude, amplitude);

  m->globalSampleCounter += 1;
  if (amplitude < m->silenceThreshold) {
    m->silenceCount += 1;
    if (m->silenceCount >= m->samplesForEOD) {
      m->eodEvents += 1;
      reset_state(m);
      return 1;
    }
  } else {
    m->silenceCount = 0;
  }

  if (!m->started) {
    double dspb = m->downsampledSamplesPerBit;
    double sampleCount = m->nbits * dspb;
    double sampleCountForBitDecision = js_round(dspb / 4);
    double matched = 0, total = 0;
    /* JS: x % 0 is NaN, NaN === 0 false */
    int due = sampleCountForBitDecision != 0 && fmod(m->globalSampleCounter, sampleCountForBitDecision) == 0;
    if (wamo_ring_length(m->syncSamples) >= sampleCount && due) {
      double len = wamo_ring_length(m->syncSamples);
      for (int j = 0; j < m->nbits; j++) {
        for (double k = 0; k < dspb; k++) {
          double v;
          int st = wamo_ring_get(m->syncSamples, len - (j * dspb + k) - 1, &v);
          if (st < 0) { m->threw = 1; return 0; }
          /* preambleSfdBits[length - j] is `undefined` for j == 0 (fsk.ts:307): a number never
           * === undefined, but an `undefined` ring read (fractional index) does. */
          if (j == 0) { if (st == 1) matched++; }
          else if (st == 0 && v == (double)m->preambleSfdBits[m->nbits - j]) matched++;
          total++;
        }
      }
      double matchRatio = total > 0 ? matched / total : 0;
      if (matchRatio > m->cfg.syncThreshold) {
        m->started = 1;
        m->current = 0; m->bitPosition = 0;
        m->bitAccumulator = 0; m->bitAccumCount = 0; m->bitSampleCounter = 0; m->nextBitSampleIndex = 0;
        m->syncDetections += 1;
        double sum = 0;
        double alen = wamo_ring_length(m->syncAmplitude);
        for (double i = 0; i < alen; i++) {
          double v;
          int st = wamo_ring_get(m->syncAmplitude, i, &v);
          if (st < 0) { m->threw = 1; return 0; }
          sum += (st == 0) ? v : NAN; /* undefined → NaN */
        }
        m->silenceThreshold = (sum / alen) * 0.1;
      }
    }
  } else {
    m->bitAccumulator += bitValue;
    m->bitAccumCount += 1;
    m->bitSampleCounter += 1;
    if (m->bitSampleCounter >= m->nextBitSampleIndex) {
      int bit = m->bitAccumulator > (m->bitAccumCount / 2) ? 1 : 0;
      m->bitAccumulator = 0; m->bitAccumCount = 0;
      m->nextBitSampleIndex += m->downsampledSamplesPerBit;
      process_byte(m, bit);
    }
  }
  return 0;
}

static int process_sample(wamo_fsk* m, double sample) { /* fsk.ts:224-276 */
  double omega = 2 * M_PI * m->centerFreq / m->cfg.sampleRate;
  double i = sample * cos(m->localOscPhase);
  double q = sample * sin(m->localOscPhase);
  m->localOscPhase = fmod(m->localOscPhase + omega, 2 * M_PI);

  i = wamo_iir_process(m->iqI, i);
  q = wamo_iir_process(m->iqQ, q);

  m->iAcc += i;
  m->qAcc += q;
  m->dsCounter += 1;

  if (m->dsCounter >= m->downsampleRatio) {
    double avgI = m->iAcc / m->downsampleRatio;
    double avgQ = m->qAcc / m->downsampleRatio;
    double currentPhase = atan2(avgQ, avgI);
    double amplitude = sqrt(avgI * avgI + avgQ * avgQ);
    double phaseDiff = currentPhase - m->lastPhase;
    if (phaseDiff > M_PI) phaseDiff -= 2 * M_PI;
    else if (phaseDiff < -M_PI) phaseDiff += 2 * M_PI;
    m->lastPhase = currentPhase;
    double filteredPhaseDiff = wamo_iir_process(m->postFilter, phaseDiff);
    int bitValue = filteredPhaseDiff > 0 ? 1 : 0;
    if (m->dtapF && m->dtapn < m->dtapcap) { m->dtapF[m->dtapn] = filteredPhaseDiff; m->dtapA[m->dtapn] = amplitude; m->dtapn++; }
    m->iAcc = 0; m->qAcc = 0; m->dsCounter = 0;
    return process_downsampled_bit(m, bitValue, amplitude);
  }
  return 0;
}

long wamo_fsk_demodulate(wamo_fsk* m, float* samples, long n, uint8_t* out, long cap) { /* fsk.ts:190-222 */
  if (!m->ready) return -1;
  m->demodulationCalls += 1;
  m->totalSamples += (double)n;
  m->threw = 0;
  if (m->has_agc) agc_process(&m->agc, samples, n);
  float* processed = (float*)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
  wamo_iir_process_buffer(m->preFilter, samples, processed, n);
  if (m->tap) memcpy(m->tap, processed, sizeof(float) * (size_t)(n < m->tapcap ? n : m->tapcap));
  for (long i = 0; i < n; i++) {
    process_sample(m, (double)processed[i]);
    if (m->threw) break;
  }
  free(processed);
  if (m->threw) { /* catch block, fsk.ts:218-221: bytes stay queued */
    m->errorEvents += 1;
    return 0;
  }
  long nout = m->nbytebuf < cap ? m->nbytebuf : cap;
  if (nout > 0) memcpy(out, m->bytebuf, (size_t)nout);
  m->nbytebuf = 0;
  return nout;
}

long wamo_fsk_modulate(wamo_fsk* m, const uint8_t* data, long n, float* out, long cap) { /* fsk.ts:377-424 */
  if (!m->ready) return -1;
  long spb = (long)m->samplesPerBit, bpb = (long)m->bitsPerByte;
  long totalBytes = m->cfg.preambleLength + m->cfg.sfdLength + n;
  long paddingSamples = totalBytes > 0 ? spb * 2 : 0;
  long silenceSamples = bpb * spb;
  long totalSamples = totalBytes * bpb * spb + paddingSamples + silenceSamples;
  if (!out) return totalSamples;
  if (cap < totalSamples) return -2;
  memset(out, 0, sizeof(float) * (size_t)totalSamples);
  long sampleIndex = paddingSamples;
  double phase = 0;
  for (long bi = 0; bi < totalBytes; bi++) {
    int byte = bi < m->cfg.preambleLength ? m->preamble[bi]
               : bi < m->cfg.preambleLength + m->cfg.sfdLength ? m->sfd[bi - m->cfg.preambleLength]
               : data[bi - m->cfg.preambleLength - m->cfg.sfdLength];
    int bits[64]; int nb = 0;
    for (int i = 0; i < m->cfg.startBits && nb < 64; i++) bits[nb++] = 0;
    for (int i = 7; i >= 0; i--) bits[nb++] = (byte >> i) & 1;
    if (m->cfg.parity != 0) {
      int parity = 0;
      for (int i = 0; i < 8; i++) parity ^= (byte >> i) & 1;
      bits[nb++] = m->cfg.parity == 1 ? parity : 1 - parity;
    }
    for (int i = 0; i < m->cfg.stopBits && nb < 64; i++) bits[nb++] = 1;
    for (int k = 0; k < nb; k++) {
      double frequency = bits[k] == 1 ? m->cfg.markFrequency : m->cfg.spaceFrequency;
      for (long i = 0; i < spb && sampleIndex < totalSamples; i++) {
        out[sampleIndex++] = (float)sin(phase);
        phase += 2 * M_PI * frequency / m->cfg.sampleRate;
      }
    }
  }
  return totalSamples;
}

void wamo_fsk_reset(wamo_fsk* m) { /* fsk.ts:464-469 */
  reset_state(m);
  if (m->syncSamples) wamo_ring_clear(m->syncSamples);
  m->nbytebuf = 0;
  m->syncDetections = 0; m->demodulationCalls = 0; m->totalSamples = 0;
}

void wamo_fsk_status_get(const wamo_fsk* m, wamo_fsk_status* st) { /* fsk.ts:481-493 */
  st->ready = m->ready;
  st->frameStarted = m->started;
  st->globalSampleCounter = m->globalSampleCounter;
  st->receivedBitsLength = m->syncSamples ? wamo_ring_length(m->syncSamples) : 0;
  st->byteBufferLength = (double)m->nbytebuf;
  st->demodulationCalls = m->demodulationCalls;
  st->syncDetections = m->syncDetections;
  st->silenceThreshold = m->silenceThreshold;
  st->totalSamplesProcessed = m->totalSamples;
  st->eodEvents = m->eodEvents;
  st->errorEvents = m->errorEvents;
  st->configuredEvents = m->configuredEvents;
}
void wamo_fsk_params(const wamo_fsk* m, double out[8]) {
  out[0] = m->samplesPerBit; out[1] = m->downsampledSamplesPerBit; out[2] = m->bitsPerByte;
  out[3] = m->nbits; out[4] = m->centerFreq;
  out[5] = m->syncSamples ? m->syncSamples->maxLength : 0;
  out[6] = m->syncAmplitude ? m->syncAmplitude->maxLength : 0;
  out[7] = m->samplesForEOD;
}
void wamo_fsk_set_prefilter_tap(wamo_fsk* m, float* buf, long cap) { m->tap = buf; m->tapcap = cap; }
void wamo_fsk_set_decim_tap(wamo_fsk* m, double* f, double* amp, long cap) { m->dtapF = f; m->dtapA = amp; m->dtapcap = cap; m->dtapn = 0; }
long wamo_fsk_decim_tap_count(const wamo_fsk* m) { return m->dtapn; }

/* ------------------------------------------------------------------------------------------
 * Multi-threaded batch driver: one FSKCore instance per stream, streams split over pthreads.
 * This is the reported CPU baseline (bench.py), not part of the reference.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  const wamo_fsk_config* cfgs; const int32_t* cfg_index;
  long s0, s1;
  float* samples; long stride, n;
  uint8_t* out; long out_stride; int32_t* out_len; wamo_fsk_status* st;
} batch_job;

static void* batch_worker(void* arg) {
  batch_job* j = (batch_job*)arg;
  for (long s = j->s0; s < j->s1; s++) {
    wamo_fsk* m = wamo_fsk_new();
    wamo_fsk_configure(m, &j->cfgs[j->cfg_index ? j->cfg_index[s] : 0]);
    long nb = wamo_fsk_demodulate(m, j->samples + s * j->stride, j->n, j->out + s * j->out_stride, j->out_stride);
    j->out_len[s] = (int32_t)nb;
    if (j->st) wamo_fsk_status_get(m, &j->st[s]);
    wamo_fsk_free(m);
  }
  return NULL;
}

int wamo_fsk_batch_demodulate(const wamo_fsk_config* cfgs, const int32_t* cfg_index, long n_streams,
                              float* samples, long stream_stride, long n_samples,
                              uint8_t* out, long out_stride, int32_t* out_len,
                              wamo_fsk_status* status_out, int n_threads) {
  if (n_threads < 1) n_threads = 1;
  if (n_threads > n_streams) n_threads = (int)(n_streams > 0 ? n_streams : 1);
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)n_threads);
  batch_job* jobs = (batch_job*)malloc(sizeof(batch_job) * (size_t)n_threads);
  for (int t = 0; t < n_threads; t++) {
    jobs[t] = (batch_job){cfgs, cfg_index, n_streams * t / n_threads, n_streams * (t + 1) / n_threads,
                          samples, stream_stride, n_samples, out, out_stride, out_len, status_out};
    pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
  }
  for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
  free(th); free(jobs);
  return 0;
}
