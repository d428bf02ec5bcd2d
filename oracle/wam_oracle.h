/*
 * wam_oracle.h — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A float64, line-faithful C restatement of the reference's physical-layer path
 * (cho45/WebAudio-Modem):
 *     src/modems/fsk.ts            FSKCore, AGCProcessor
 *     src/dsp/filters.ts           IIRFilter, FIRFilter, FilterDesign, FilterFactory
 *     src/utils.ts                 RingBuffer (incl. the fractional-capacity behaviour)
 *     src/utils/crc16.ts           CRC16
 *     src/transports/xmodem/packet.ts, types.ts, xmodem.ts:232-321 (receive-side checks)
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker / reported CPU baseline.
 * The product (webaudio-modem_b200/libwam.so) never links, loads or calls it.
 *
 * Parity pin: the reference cannot be executed in this image (no Node/V8), so the oracle
 * is pinned against every known-answer expectation in the reference's own tests
 * (tests/test_oracle_*.py cite them file:line).  Transcendentals come from glibc libm
 * instead of V8's fdlibm port (both < 1 ulp); see DESIGN.md "Oracle".
 */
#ifndef WAM_ORACLE_H
#define WAM_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Mirrors FSKConfig (src/modems/fsk.ts:5-17) + BaseModulatorConfig{sampleRate, baudRate}.
 * Same layout as wam_fsk_config in include/wam.h so one ctypes.Structure serves both. */
typedef struct wamo_fsk_config {
  double sampleRate;
  double baudRate;
  double markFrequency;
  double spaceFrequency;
  const uint8_t* preamblePattern;
  int32_t preambleLength;
  const uint8_t* sfdPattern;
  int32_t sfdLength;
  int32_t startBits;
  int32_t stopBits;
  int32_t parity;             /* 0 none, 1 even, 2 odd */
  double syncThreshold;
  int32_t agcEnabled;
  double preFilterBandwidth;
  int32_t adaptiveThreshold;  /* dead flag in the reference (fsk.ts:16,32) */
} wamo_fsk_config;

/* getStatus() (src/modems/fsk.ts:481-493) + event counters */
typedef struct wamo_fsk_status {
  int32_t ready;
  int32_t frameStarted;
  double globalSampleCounter;
  double receivedBitsLength;
  double byteBufferLength;
  double demodulationCalls;
  double syncDetections;
  double silenceThreshold;
  double totalSamplesProcessed;
  double eodEvents;        /* number of emit('eod') so far */
  double errorEvents;      /* number of emit('error') so far */
  double configuredEvents; /* number of emit('configured') so far */
} wamo_fsk_status;

typedef struct wamo_fsk wamo_fsk;

void wamo_default_config(wamo_fsk_config* cfg);           /* DEFAULT_FSK_CONFIG fsk.ts:19-33 */
wamo_fsk* wamo_fsk_new(void);                             /* new FSKCore() */
void wamo_fsk_free(wamo_fsk* m);
void wamo_fsk_configure(wamo_fsk* m, const wamo_fsk_config* cfg);   /* fsk.ts:133-157 */
/* fsk.ts:377-424. Returns number of samples (or -1 if not configured). If out==NULL only sizes. */
long wamo_fsk_modulate(wamo_fsk* m, const uint8_t* data, long n, float* out, long cap);
/* fsk.ts:190-222. samples are mutated in place when AGC is on. Returns number of bytes
 * written to out (<= cap), or -1 if not configured. */
long wamo_fsk_demodulate(wamo_fsk* m, float* samples, long n, uint8_t* out, long cap);
void wamo_fsk_reset(wamo_fsk* m);                         /* fsk.ts:464-469 */
void wamo_fsk_status_get(const wamo_fsk* m, wamo_fsk_status* st);
/* derived parameters, for tests: [spb, dspb, bpb, nbits, centerFreq, syncRingCapacity, ampRingCapacity, samplesForEOD] */
void wamo_fsk_params(const wamo_fsk* m, double out[8]);
/* optional tap: record the pre-filtered f32 samples of the last demodulate call (tests compare
 * "filtered samples" within 1e-4).  buffer owned by caller, cap floats. */
void wamo_fsk_set_prefilter_tap(wamo_fsk* m, float* buf, long cap);
/* optional tap: filteredPhaseDiff (fsk.ts:261) and amplitude (fsk.ts:252) of every decimated sample since the tap was set */
void wamo_fsk_set_decim_tap(wamo_fsk* m, double* f, double* amp, long cap);
long wamo_fsk_decim_tap_count(const wamo_fsk* m);

/* ---- filters.ts ---- */
typedef struct wamo_iir wamo_iir;
/* returns NULL and sets *err to 1,2,3 for the three constructor errors (filters.ts:19-21) */
wamo_iir* wamo_iir_new(const double* b, int nb, const double* a, int na, int* err);
void wamo_iir_free(wamo_iir* f);
double wamo_iir_process(wamo_iir* f, double x);
void wamo_iir_process_buffer(wamo_iir* f, const float* in, float* out, long n);
void wamo_iir_reset(wamo_iir* f);
int wamo_iir_coefficients(const wamo_iir* f, double* b, double* a); /* returns nb | na<<16 */

typedef struct wamo_fir wamo_fir;
wamo_fir* wamo_fir_new(const double* taps, int n);
void wamo_fir_free(wamo_fir* f);
double wamo_fir_process(wamo_fir* f, double x);
void wamo_fir_process_buffer(wamo_fir* f, const float* in, float* out, long n);
void wamo_fir_reset(wamo_fir* f);

void wamo_design_butterworth_lowpass(double fc, double fs, double b[3], double a[3]);   /* filters.ts:180-192 */
void wamo_design_butterworth_highpass(double fc, double fs, double b[3], double a[3]);  /* filters.ts:200-212 */
void wamo_design_butterworth_bandpass(double f0, double bw, double fs, double b[3], double a[3]); /* :221-234 */
/* windowed-sinc designers; return the (possibly incremented) tap count; out must hold numTaps+1 */
int wamo_design_sinc_lowpass(double fc, double fs, int numTaps, double* out);            /* :243-265 */
int wamo_design_sinc_highpass(double fc, double fs, int numTaps, double* out);           /* :274-286 */
int wamo_design_sinc_bandpass(double f0, double bw, double fs, int numTaps, double* out);/* :296-314 */

/* ---- utils.ts RingBuffer (exposed for tests of the fractional-capacity emulation) ---- */
typedef struct wamo_ring wamo_ring;
wamo_ring* wamo_ring_new(int elem_kind /*0 u8, 1 f32*/, double size);
void wamo_ring_free(wamo_ring* r);
void wamo_ring_put(wamo_ring* r, double v);
/* returns 0 value ok, 1 'undefined', -1 throws 'Index out of bounds' */
int wamo_ring_get(const wamo_ring* r, double index, double* v);
double wamo_ring_length(const wamo_ring* r);
void wamo_ring_clear(wamo_ring* r);

/* ---- crc16.ts / packet.ts ---- */
uint16_t wamo_crc16(const uint8_t* data, long n);                                /* crc16.ts:21-38 */
/* XModemPacket.createData + serialize (packet.ts:21-54). returns bytes written or -1/-2 on the
 * two createData errors. */
long wamo_xmodem_serialize(int sequence, const uint8_t* payload, long n, uint8_t* out, long cap);

/* Receive-side classification of one byte stream, following xmodem.ts:232-321:
 * skip bytes until SOH (EOT ends), read seq/~seq/len, check seq+~seq==255, compare with
 * expected sequence, read len+2, CRC over payload. */
enum {
  WAMO_PKT_OK = 0,          /* in-sequence packet, CRC good → ACK */
  WAMO_PKT_DUPLICATE = 1,   /* previous sequence (xmodem.ts:309-314) → ACK, dropped */
  WAMO_PKT_NO_SOH = 2,      /* ran out of bytes before an SOH */
  WAMO_PKT_EOT = 3,         /* EOT seen before SOH (xmodem.ts:242-245) */
  WAMO_PKT_INCOMPLETE = 4,  /* ran out of bytes inside header or payload (would time out) */
  WAMO_PKT_BAD_COMPLEMENT = 5, /* 'Invalid sequence number' (xmodem.ts:270-274) */
  WAMO_PKT_BAD_CRC = 6,     /* 'Invalid CRC' (xmodem.ts:286-290) */
  WAMO_PKT_UNEXPECTED_SEQ = 7 /* xmodem.ts:315-320 */
};
typedef struct wamo_pkt_result {
  int32_t status;
  int32_t sequence;
  int32_t length;
  int32_t payloadOffset;  /* index of payload[0] in the input buffer, -1 if n/a */
  int32_t crcReceived;
  int32_t crcComputed;
  int32_t bytesConsumed;
} wamo_pkt_result;
void wamo_xmodem_check(const uint8_t* bytes, long n, int expectedSequence, wamo_pkt_result* res);

/* Receive side of XModemTransport.receiveAllPackets / receiveAndProcessPacket (xmodem.ts:232-321) over one
 * burst of demodulated bytes, with the receiver state carried between bursts.  No timers: where the
 * reference would wait for more bytes (and eventually time out), the walk stops and reports how many bytes
 * it consumed (an unfinished packet is left unconsumed from its SOH on).  An error (bad complement, bad CRC,
 * unexpected sequence) counts a retry; beyond maxRetries the session fails (xmodem.ts:253-255), otherwise the
 * rest of the burst is discarded (receive.buffer = [], xmodem.ts:257) and a NAK is queued. */
typedef struct wamo_xmodem_rx_state {
  int32_t expectedSequence; /* receive.expectedSequence, starts at 1 */
  int32_t retries;          /* send.retries (xmodem.ts:253,299) */
  int32_t done;             /* 0 running, 1 EOT received and ACKed, 2 failed after max retries */
  int32_t dataLen;          /* reassembled payload bytes so far (receive.data) */
  int32_t packetsReceived;  /* statistics.packetsReceived (xmodem.ts:277) */
  int32_t packetsDropped;   /* statistics.packetsDropped (xmodem.ts:271,287,312,317) */
} wamo_xmodem_rx_state;
/* replies: ACK 0x06 / NAK 0x15 in the order they would be sent (at most reply_cap are stored, *n_replies
 * counts all); data: payloads are appended at st->dataLen (bytes beyond data_cap are dropped, dataLen still
 * advances); returns the number of bytes consumed from the burst. */
long wamo_xmodem_receive(const uint8_t* bytes, long n, int maxRetries, wamo_xmodem_rx_state* st,
                         uint8_t* replies, int reply_cap, int32_t* n_replies, uint8_t* data, long data_cap);

/* ---- multi-threaded batch driver (CPU baseline for bench.py; one FSKCore per stream) ---- */
/* Demodulates n_streams independent streams ([stream][n_samples], stride in floats) with
 * n_threads pthreads.  out: [stream][out_stride] bytes, out_len[stream].  Returns 0. */
int wamo_fsk_batch_demodulate(const wamo_fsk_config* cfgs, const int32_t* cfg_index, long n_streams,
                              float* samples, long stream_stride, long n_samples,
                              uint8_t* out, long out_stride, int32_t* out_len,
                              wamo_fsk_status* status_out /* nullable, per stream */, int n_threads);

#ifdef __cplusplus
}
#endif
#endif
