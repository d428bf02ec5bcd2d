/*
 * wam.h — C ABI of the B200-native FSK physical layer (libwam.so).
 *
 * This is the drop-in boundary for ONE path of cho45/WebAudio-Modem: FSKCore
 * modulate/demodulate (src/modems/fsk.ts) + the filters it uses (src/dsp/filters.ts)
 * + CRC-16 / XModem packet checking for framed blocks (src/utils/crc16.ts,
 * src/transports/xmodem/packet.ts).  Each entry point names the reference interface it
 * replaces.  The reference-side binding (N-API addon + TypeScript class implementing
 * IModulator, src/core.ts:88-117) is shown in INTEGRATION.md.
 *
 * Conventions
 *  - plain C types only; caller owns every buffer; no C++ exceptions cross the ABI.
 *  - every function returns 0 (WAM_OK) or a negative wam_error; wam_last_error() gives a
 *    thread-local message.  WAM_E_NOT_CONFIGURED maps to the reference's
 *    Error('FSK modulator not configured') / Error('FSK demodulator not configured')
 *    (fsk.ts:191-193,378-380).
 *  - one handle = one CUDA device; handles are not re-entrant (FSKCore is not either).
 *  - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *    WAM_E_CUDA.
 */
#ifndef WAM_H
#define WAM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WAM_VERSION 100 /* 0.1.0 */

typedef enum wam_error {
  WAM_OK = 0,
  WAM_E_INVALID = -1,        /* bad argument */
  WAM_E_NOT_CONFIGURED = -2, /* fsk.ts:191-193,378-380 */
  WAM_E_CUDA = -3,           /* CUDA runtime/driver error, or no device */
  WAM_E_NOMEM = -4,
  WAM_E_CAPACITY = -5,       /* caller's output buffer too small */
  WAM_E_UNSUPPORTED = -6,    /* configuration outside what the kernels implement */
  WAM_E_FILTER_B_EMPTY = -10, /* 'Feedforward coefficients (b) cannot be empty'  filters.ts:19 */
  WAM_E_FILTER_A_EMPTY = -11, /* 'Feedback coefficients (a) cannot be empty'     filters.ts:20 */
  WAM_E_FILTER_A0_ZERO = -12, /* 'First feedback coefficient (a[0]) cannot be zero' filters.ts:21 */
  WAM_E_PKT_SEQUENCE = -20,   /* 'Invalid sequence: N. Must be 1-255.'  packet.ts:22-24 */
  WAM_E_PKT_PAYLOAD = -21     /* 'Payload too large: N. Max 255 bytes.' packet.ts:25-27 */
} wam_error;

/* FSKConfig — src/modems/fsk.ts:5-17 (+ sampleRate, baudRate from BaseModulatorConfig).
 * Field names and meaning are the reference's; DEFAULT_FSK_CONFIG via wam_fsk_default_config. */
typedef struct wam_fsk_config {
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
  int32_t parity;             /* 0 'none', 1 'even', 2 'odd' */
  double syncThreshold;
  int32_t agcEnabled;
  double preFilterBandwidth;
  int32_t adaptiveThreshold;  /* accepted and ignored, like the reference (dead flag) */
} wam_fsk_config;

/* FSKCore.getStatus() — src/modems/fsk.ts:481-493, plus the event counts a host needs to
 * re-emit 'eod' / 'error' (fsk.ts:289,219). */
typedef struct wam_fsk_status {
  int32_t ready;
  int32_t frameStarted;
  double globalSampleCounter;
  double receivedBitsLength;
  double byteBufferLength;
  double demodulationCalls;
  double syncDetections;
  double silenceThreshold;
  double totalSamplesProcessed;
  double eodEvents;
  double errorEvents;
  double configuredEvents;
} wam_fsk_status;

int wam_version(void);
const char* wam_last_error(void);
const char* wam_error_string(int code);
int wam_device_count(int* count);

/* DEFAULT_FSK_CONFIG — fsk.ts:19-33 (pattern pointers reference static storage) */
void wam_fsk_default_config(wam_fsk_config* cfg);

/* ---------------------------------------------------------------------------------------
 * Single-stream modem: drop-in for one FSKCore instance (IModulator, src/core.ts:88-117)
 * ------------------------------------------------------------------------------------- */
typedef struct wam_fsk wam_fsk;

int wam_fsk_create(int device, wam_fsk** out);               /* new FSKCore()           fsk.ts:82 */
int wam_fsk_destroy(wam_fsk* m);                             /* dispose()               core.ts:106 */
int wam_fsk_configure(wam_fsk* m, const wam_fsk_config* c);  /* configure()             fsk.ts:133-157 */
int wam_fsk_is_ready(wam_fsk* m);                            /* isReady()               core.ts:270-272 */
/* number of samples modulateData(nbytes) produces (fsk.ts:391-394); negative wam_error */
long wam_fsk_modulate_size(wam_fsk* m, long nbytes);
/* modulateData() — fsk.ts:377-424.  out: host float32[cap]. */
int wam_fsk_modulate(wam_fsk* m, const uint8_t* data, long nbytes, float* out, long cap, long* n_out);
/* demodulateData() — fsk.ts:190-222.  samples: host float32[n]; MUTATED IN PLACE when AGC is
 * enabled, exactly like the reference (fsk.ts:55).  out receives the bytes completed during
 * this call (returned once).  */
int wam_fsk_demodulate(wam_fsk* m, float* samples, long n, uint8_t* out, long cap, long* n_out);
int wam_fsk_reset(wam_fsk* m);                               /* reset()                 fsk.ts:464-469 */
int wam_fsk_status_get(wam_fsk* m, wam_fsk_status* st);      /* getStatus()             fsk.ts:481-493 */

/* ---------------------------------------------------------------------------------------
 * Batched modem: n_streams independent FSKCore instances with device-resident streaming
 * state, one GPU.  (The new entry point named by the north star; equivalent to n_streams
 * reference instances each receiving the same demodulateData/modulateData call.)
 * ------------------------------------------------------------------------------------- */
typedef struct wam_fsk_batch wam_fsk_batch;

enum {
  WAM_BATCH_WRITEBACK_AGC = 1u << 0, /* write the AGC-scaled samples back (reference mutates its input) */
  WAM_BATCH_TAP_PREFILTER = 1u << 1, /* (device API) write pre-filtered f32 samples to tap buffer */
  WAM_BATCH_DEBUG_GENERIC_SM = 1u << 2, /* test hook: per-sample state machine instead of the event-driven one */
  WAM_BATCH_NO_PIPELINE = 1u << 3,    /* test hook: never use the warp-specialised few-stream kernel */
  WAM_BATCH_NO_TMA = 1u << 4,         /* test hook: stage input tiles with cp.async instead of TMA */
  WAM_BATCH_NO_SLABS = 1u << 5,       /* test hook: every warp walks its streams in one pass (no dynamic time slabs) */
  WAM_BATCH_FORCE_SLABS = 1u << 6,    /* test hook: dynamic time slabs even for few streams / short calls */
  WAM_BATCH_EXACT_ONLY = 1u << 7,     /* never take the mixed-precision fast path: float64 kernels only */
  WAM_BATCH_FORCE_FAST = 1u << 8,     /* test hook: fast path even for few streams / short calls */
  WAM_BATCH_FAST_UNGUARDED = 1u << 9, /* test hook: keep the float32 results of flagged streams (no float64 re-run) */
  WAM_BATCH_TAP_FAST_DECISION = 1u << 10 /* test hook (device API, fast path): tap[stream][2k, 2k+1] = filtered phase
                                            difference and doubt band of decimated sample k */
};

/* cfg_index[stream] selects cfgs[]; NULL = all streams use cfgs[0]. */
int wam_fsk_batch_create(int device, long n_streams, const wam_fsk_config* cfgs, int n_cfgs,
                         const int32_t* cfg_index, wam_fsk_batch** out);
int wam_fsk_batch_destroy(wam_fsk_batch* b);
int wam_fsk_batch_reset(wam_fsk_batch* b);                   /* reset() on every stream */
/* new FSKCore() + configure() on every stream again (fresh AGC, filters, rings, counters);
 * asynchronous on cuda_stream. */
int wam_fsk_batch_renew(wam_fsk_batch* b, void* cuda_stream);
/* bytes of output capacity per stream that n_samples of input can never exceed */
long wam_fsk_batch_out_capacity(wam_fsk_batch* b, long n_samples);

/* HOST buffers: samples float32 [n_streams][stream_stride] (n_samples used per stream),
 * out uint8 [n_streams][out_stride], out_len int32 [n_streams].  Copies H2D/D2H inside,
 * pipelined with the kernels.  Equivalent to demodulateData(samples[s]) on every stream. */
int wam_fsk_batch_demodulate(wam_fsk_batch* b, float* samples, long stream_stride, long n_samples,
                             uint8_t* out, long out_stride, int32_t* out_len, uint32_t flags);
/* DEVICE buffers (same shapes), asynchronous on cuda_stream (a cudaStream_t, NULL = default).
 * tap: optional device float32 [n_streams][stream_stride] for WAM_BATCH_TAP_PREFILTER. */
int wam_fsk_batch_demodulate_device(wam_fsk_batch* b, float* d_samples, long stream_stride, long n_samples,
                                    uint8_t* d_out, long out_stride, int32_t* d_out_len, float* d_tap,
                                    void* cuda_stream, uint32_t flags);
/* As wam_fsk_batch_demodulate, the samples arriving as 16-bit PCM: int16 [n_streams][stream_stride],
 * sample = pcm / 32768 (exact in float32), i.e. demodulateData(Float32Array.from(pcm, v => v / 32768)).
 * Half the host->device bytes of the float32 entry; the widening runs on the device. */
int wam_fsk_batch_demodulate_pcm16(wam_fsk_batch* b, const int16_t* samples, long stream_stride, long n_samples,
                                   uint8_t* out, long out_stride, int32_t* out_len, uint32_t flags);
/* Ragged batch: stream s receives demodulateData(samples[s][0 .. n_valid[s])) with 0 <= n_valid[s] <= n_samples;
 * n_valid[s] < 0 means demodulateData() is NOT called on stream s in this round (state and counters untouched,
 * out_len[s] = 0).  This is the entry point of a server that multiplexes many independent audio sessions, each
 * delivering its own block sizes at its own pace (fsk-processor.ts:152-167 calls demodulateData once per
 * 128-sample render quantum per session).  getStatus().demodulationCalls / totalSamplesProcessed are per stream. */
int wam_fsk_batch_demodulate_ragged(wam_fsk_batch* b, float* samples, long stream_stride, long n_samples,
                                    const int32_t* n_valid, uint8_t* out, long out_stride, int32_t* out_len,
                                    uint32_t flags);
int wam_fsk_batch_demodulate_ragged_device(wam_fsk_batch* b, float* d_samples, long stream_stride, long n_samples,
                                           const int32_t* d_n_valid, uint8_t* d_out, long out_stride,
                                           int32_t* d_out_len, void* cuda_stream, uint32_t flags);
/* Session multiplexer: the host adapter for many concurrent block-wise callers.  In the reference every audio
 * session owns an FSKCore and FSKProcessor.process() hands it one 128-sample render quantum at a time
 * (fsk-processor.ts:152-167, :296-322).  Here n_sessions such callers push their blocks into pinned staging
 * (wam_fsk_mux_push, up to max_block samples per session between flushes) and one wam_fsk_mux_flush runs a single
 * ragged batch over the sessions that pushed: one flush = one demodulateData() call per such session over what it
 * pushed; sessions that pushed nothing are not called.  Not re-entrant: push/flush from one thread (or lock). */
typedef struct wam_fsk_mux wam_fsk_mux;
int wam_fsk_mux_create(int device, long n_sessions, const wam_fsk_config* cfgs, int n_cfgs, const int32_t* cfg_index,
                       long max_block, wam_fsk_mux** out);
int wam_fsk_mux_destroy(wam_fsk_mux* m);
int wam_fsk_mux_push(wam_fsk_mux* m, long session, const float* samples, long n);
long wam_fsk_mux_pending(wam_fsk_mux* m, long session);
long wam_fsk_mux_out_capacity(wam_fsk_mux* m);   /* bytes one flush can produce per session at most */
int wam_fsk_mux_flush(wam_fsk_mux* m, uint8_t* out, long out_stride, int32_t* out_len);
wam_fsk_batch* wam_fsk_mux_batch(wam_fsk_mux* m); /* the sessions' FSKCore instances (status, reset); owned by the mux */
/* Send half: one ChunkedModulator (src/webaudio/chunked-modulator.ts:31-87) per session.  wam_fsk_mux_send queues a
 * payload (startModulation; an empty payload resets the session), wam_fsk_mux_modulate runs modulateData() for every
 * session that queued one as ONE batched GPU call, wam_fsk_mux_pull is getNextSamples(sampleCount): returns 1 and
 * fills out / res while the session is sending, 0 when it is not (the reference returns null). */
typedef struct wam_chunk_result {
  long samples;          /* samples written to out (ChunkResult.signal.length) */
  int isComplete;
  long samplesConsumed;
  long totalSamples;
} wam_chunk_result;
int wam_fsk_mux_send(wam_fsk_mux* m, long session, const uint8_t* data, long n);
int wam_fsk_mux_modulate(wam_fsk_mux* m);
int wam_fsk_mux_is_modulating(wam_fsk_mux* m, long session);
int wam_fsk_mux_pull(wam_fsk_mux* m, long session, float* out, long sample_count, wam_chunk_result* res);
/* per-stream getStatus(); st: host array [n_streams] */
int wam_fsk_batch_status(wam_fsk_batch* b, wam_fsk_status* st);
/* profiling aid: per-phase SM cycle counters of the demodulator (A1, A2, B, other), summed over CTAs */
int wam_fsk_batch_debug_phase_cycles(wam_fsk_batch* b, int enable, double* out4, double* per_cta, long n_ctas);
/* kernels launched by this handle so far (bench.py's gpu_launches) */
long wam_fsk_batch_launch_count(wam_fsk_batch* b);
/* Mixed-precision fast path (float32 kernel with certified decisions; streams whose decisions the float32 error
 * could have turned are demodulated again in float64): counters over the batch's life. */
typedef struct wam_fast_stats {
  int64_t fast_calls;          /* demodulate calls served by the fast kernel */
  int64_t flagged_last_call;   /* streams re-run in float64 over the whole of the most recent fast call */
  int64_t flagged_streams;     /* streams flagged at least once */
  int64_t doubtful_samples;    /* decimated samples whose hard bit was inside the float32 error band */
  uint32_t flag_causes;        /* OR of the causes seen: 1 start-bit vote, 2 data-bit vote, 4 stop-bit vote, 8 sync, 16 EOD, 32 range */
  uint32_t error_flags;        /* OR of the streams' error words (1 output overflow, 2 pipeline timeout, 4 slab timeout) */
  /* most recent fast call: doubtful decisions checked by a float64 run of a short window around them, by outcome
   * (confirmed: the fast results stand; refuted or dropped for lack of room: the stream is re-run over the whole call) */
  int64_t windows_confirmed, windows_refuted, windows_dropped;
  /* streams whose open float32 readings at the end of a fast call were re-read in float64 at the start of the next
   * call (over the batch's life), and how many of them the float64 reading changed */
  int64_t carried_settled, carried_corrected;
} wam_fast_stats;
int wam_fsk_batch_fast_stats(wam_fsk_batch* b, wam_fast_stats* out);
/* Channel model of the synthetic workloads (BASELINE config 5: modulate -> AWGN -> demodulate): adds Gaussian noise
 * of standard deviation d_sigma[row] in place to device rows (counter-based Philox4x32-10; seed, seq, row and column
 * fix every value).  n and stride multiples of 4, rows 16-byte aligned.  Not part of FSKCore. */
int wam_awgn_add_device(float* d_samples, long stride, long n_rows, long n, const float* d_sigma,
                        unsigned long long seed, unsigned int seq, void* cuda_stream);
/* Debug: the float64 verification windows of the last fast call of configuration group `group`.  *n_classes = number
 * of window classes C (0..C-3: windows of c + 1 time slabs; C-2 and C-1: the two stages of an end-of-data check);
 * counts (nullable) [C] items per class; items (nullable, items_cap >= C * cap * 3):
 * items[(c * cap + i) * 3 + {0,1,2}] = stream (index inside the group), slab, result (1 = the fast results stand, else
 * 1 << 30 | bits naming what differed).  Returns cap, negative on error. */
int wam_fsk_batch_debug_fast_windows(wam_fsk_batch* b, int group, int* n_classes, int32_t* counts, int32_t* items,
                                     long items_cap);
/* test hook: multiplies the fast kernel's doubt band (1 = calibrated); a wide band flags many decisions */
int wam_fsk_batch_debug_fast_band(wam_fsk_batch* b, double scale);

/* modulateData() for every stream: data uint8 [n_streams][data_stride], data_len[s] bytes each
 * (NULL = nbytes for all).  out float32 [n_streams][out_stride]; out_len[s] samples written.
 * HOST buffers. */
int wam_fsk_batch_modulate(wam_fsk_batch* b, const uint8_t* data, long data_stride, const int32_t* data_len,
                           long nbytes, float* out, long out_stride, int32_t* out_len);
int wam_fsk_batch_modulate_device(wam_fsk_batch* b, const uint8_t* d_data, long data_stride,
                                  const int32_t* d_data_len, long nbytes, float* d_out, long out_stride,
                                  int32_t* d_out_len, void* cuda_stream);

/* ---------------------------------------------------------------------------------------
 * CRC-16 / XModem framed-block check (src/utils/crc16.ts, src/transports/xmodem/packet.ts,
 * receive-side rules src/transports/xmodem/xmodem.ts:232-321)
 * ------------------------------------------------------------------------------------- */
typedef enum wam_pkt_status {
  WAM_PKT_OK = 0,             /* in sequence, CRC good → ACK                      xmodem.ts:276-307 */
  WAM_PKT_DUPLICATE = 1,      /* previous sequence → ACK, dropped                 xmodem.ts:309-314 */
  WAM_PKT_NO_SOH = 2,         /* no SOH in the bytes                              xmodem.ts:236-252 */
  WAM_PKT_EOT = 3,            /* EOT before any SOH                               xmodem.ts:242-245 */
  WAM_PKT_INCOMPLETE = 4,     /* bytes end inside header/payload (would time out) */
  WAM_PKT_BAD_COMPLEMENT = 5, /* 'Invalid sequence number'                        xmodem.ts:270-274 */
  WAM_PKT_BAD_CRC = 6,        /* 'Invalid CRC'                                    xmodem.ts:286-290 */
  WAM_PKT_UNEXPECTED_SEQ = 7  /* 'Unexpected sequence number'                     xmodem.ts:315-320 */
} wam_pkt_status;

typedef struct wam_pkt_result {
  int32_t status;        /* wam_pkt_status */
  int32_t sequence;      /* -1 if not reached */
  int32_t length;
  int32_t payloadOffset; /* index of payload[0] within the stream's bytes, -1 if n/a */
  int32_t crcReceived;
  int32_t crcComputed;
  int32_t bytesConsumed;
} wam_pkt_result;

/* CRC16.calculate — crc16.ts:21-38 (host, scalar; for single packets) */
uint16_t wam_crc16(const uint8_t* data, long n);
/* XModemPacket.createData + serialize — packet.ts:21-54.  Returns bytes written or wam_error. */
long wam_xmodem_serialize(int sequence, const uint8_t* payload, long n, uint8_t* out, long cap);
/* One warp per stream: locate SOH, check seq/~seq, length, CRC-16 over the payload.
 * HOST buffers: bytes [n_streams][stride], len[s]; expected_seq[s] (NULL = 1). */
int wam_xmodem_batch_check(int device, const uint8_t* bytes, long stride, const int32_t* len,
                           const int32_t* expected_seq, long n_streams, wam_pkt_result* results);
int wam_xmodem_batch_check_device(const uint8_t* d_bytes, long stride, const int32_t* d_len,
                                  const int32_t* d_expected_seq, long n_streams, wam_pkt_result* d_results,
                                  void* cuda_stream);
/* Batched receive side of XModemTransport (receiveAllPackets + receiveAndProcessPacket,
 * xmodem.ts:232-321): every stream's burst of demodulated bytes is walked packet by packet with the
 * receiver state carried between calls — sequence tracking, duplicate detection, retry counting, payload
 * reassembly, and the ACK (0x06) / NAK (0x15) bytes the transport would send, in order.  No timers: where
 * the reference would wait for more bytes the walk stops; consumed[s] tells how many bytes were used (an
 * unfinished packet is left unconsumed from its SOH on, present it again with the bytes that follow).  An
 * error (bad complement, bad CRC, unexpected sequence) counts a retry; beyond max_retries the session
 * fails (done = 2, xmodem.ts:253-255), otherwise the rest of the burst is discarded (receive.buffer = [],
 * xmodem.ts:257) and a NAK is queued.  Initialise a state with {1, 0, 0, 0, 0, 0}. */
typedef struct wam_xmodem_rx_state {
  int32_t expectedSequence; /* receive.expectedSequence (starts at 1) */
  int32_t retries;          /* send.retries, xmodem.ts:253,299 */
  int32_t done;             /* 0 running, 1 EOT received and ACKed, 2 failed after max retries */
  int32_t dataLen;          /* reassembled payload bytes so far (receive.data) */
  int32_t packetsReceived;  /* statistics.packetsReceived, xmodem.ts:277 */
  int32_t packetsDropped;   /* statistics.packetsDropped, xmodem.ts:271,287,312,317 */
} wam_xmodem_rx_state;
/* HOST buffers: bytes [n_streams][stride], len[s]; state[s] in/out; replies [n_streams][reply_cap]
 * (n_replies[s] counts all replies, also those beyond reply_cap); data [n_streams][data_stride]: payloads
 * are appended at state[s].dataLen (data may be NULL: only the counters advance). */
int wam_xmodem_batch_receive(int device, const uint8_t* bytes, long stride, const int32_t* len, long n_streams,
                             int max_retries, wam_xmodem_rx_state* state, uint8_t* replies, int reply_cap,
                             int32_t* n_replies, int32_t* consumed, uint8_t* data, long data_stride);
/* Same with DEVICE pointers, asynchronous on cuda_stream (e.g. fed straight from
 * wam_fsk_batch_demodulate_device's output rows). */
int wam_xmodem_batch_receive_device(const uint8_t* d_bytes, long stride, const int32_t* d_len, long n_streams,
                                    int max_retries, wam_xmodem_rx_state* d_state, uint8_t* d_replies,
                                    int reply_cap, int32_t* d_n_replies, int32_t* d_consumed, uint8_t* d_data,
                                    long data_stride, void* cuda_stream);
/* CRC-16 of n_blocks byte blocks on the GPU (warp per block). HOST buffers. */
int wam_crc16_batch(int device, const uint8_t* bytes, long stride, const int32_t* len, long n_blocks,
                    uint16_t* crc_out);

/* ---------------------------------------------------------------------------------------
 * filters.ts — designs (host math) and batched application (GPU)
 * ------------------------------------------------------------------------------------- */
void wam_design_butterworth_lowpass(double fc, double fs, double b[3], double a[3]);   /* filters.ts:180-192 */
void wam_design_butterworth_highpass(double fc, double fs, double b[3], double a[3]);  /* filters.ts:200-212 */
void wam_design_butterworth_bandpass(double f0, double bw, double fs, double b[3], double a[3]); /* :221-234 */
int wam_design_sinc_lowpass(double fc, double fs, int numTaps, double* out);            /* :243-265 */
int wam_design_sinc_highpass(double fc, double fs, int numTaps, double* out);           /* :274-286 */
int wam_design_sinc_bandpass(double f0, double bw, double fs, int numTaps, double* out);/* :296-314 */

/* IIRFilter.processBuffer for n_streams independent filter instances sharing (b, a)
 * (filters.ts:47-87): in/out host float32 [n_streams][stride], n samples each.  state: host
 * double [n_streams][nb + na - 1... ] see wam_iir_state_size(); NULL = fresh filters, state not
 * returned.  Time-parallel (chunked linear-recurrence scan) for long streams. */
long wam_iir_state_size(int nb, int na);
int wam_iir_process_batch(int device, const double* b, int nb, const double* a, int na,
                          const float* in, float* out, long stride, long n, long n_streams, double* state);
/* FIRFilter.processBuffer (filters.ts:125-151): shared-memory-staged FIR.  state: host double
 * [n_streams][ntaps-1] most-recent-first input history, NULL = fresh. */
int wam_fir_process_batch(int device, const double* taps, int ntaps, const float* in, float* out,
                          long stride, long n, long n_streams, double* state);
/* The same with DEVICE buffers on the current device, asynchronous on cuda_stream (the measured entry points:
 * bench_extra.py).  IIR: d_state (nullable) is read and written in place; d_scratch of wam_iir_scratch_bytes(n,
 * n_streams) bytes holds the look-back records of the call.  FIR: d_taps in device memory; d_state (nullable) is updated
 * through d_state_scratch (same size). */
size_t wam_iir_scratch_bytes(long n, long n_streams);
int wam_iir_process_batch_device(const double* b, int nb, const double* a, int na, const float* d_in, float* d_out,
                                 long stride, long n, long n_streams, double* d_state, void* d_scratch,
                                 size_t scratch_bytes, void* cuda_stream);
int wam_fir_process_batch_device(const double* d_taps, int ntaps, const float* d_in, float* d_out, long stride,
                                 long n, long n_streams, double* d_state, double* d_state_scratch, void* cuda_stream);

/* test hook: the device float64 primitives of the discriminator (fast_atan2 / fast_sqrt / fast_rcp)
 * evaluated on host arrays; used by tests/test_gpu_fastmath.py to bound their error against libm. */
int wam_debug_fastmath(int device, const double* y, const double* x, long n, double* out_atan2,
                       double* out_sqrt, double* out_rcp);

/* Pins the calling thread to the CPUs local to `device` (sysfs local_cpulist of its PCIe address): staging buffers
 * allocated and first touched afterwards are NUMA-local to the GPU.  Returns the CPU count of the new mask, 0 when the
 * topology is not exposed or not allowed (nothing changed), negative on a CUDA error. */
int wam_host_bind_near_device(int device);
/* pinned host memory helpers (so the HOST-buffer entry points can overlap copies) */
int wam_host_alloc(void** p, size_t bytes);
int wam_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* WAM_H */
