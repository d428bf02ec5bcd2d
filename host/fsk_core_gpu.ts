/**
 * FSKCoreGPU — drop-in for FSKCore (src/modems/fsk.ts) backed by libwam.so through the N-API addon
 * in host/wam_napi.c.  Implements IModulator<FSKConfig> (src/core.ts:88-117) with the reference's
 * method names, config fields, error messages and events, and adds the batched entry point.
 *
 * NOT EXECUTED in the build environment (no Node toolchain in the image): it is the binding a
 * maintainer drops next to src/modems/fsk.ts; the executable mirror used by the parity tests is
 * webaudio-modem_b200/fsk.py over the same C ABI (include/wam.h).
 */
import { BaseModulator, type ModulationType, type SignalQuality } from '../src/core';
import { DEFAULT_FSK_CONFIG, type FSKConfig } from '../src/modems/fsk';

type Rows = { bytes: Uint8Array; lengths: Int32Array; stride: number };
type Chunk = { signal: Float32Array; isComplete: boolean; samplesConsumed: number; totalSamples: number };

// eslint-disable-next-line @typescript-eslint/no-var-requires
const native = require('./build/Release/wam_napi.node') as {
  // one stream (wam_fsk_*)
  fskCreate(device: number): object;
  fskConfigure(h: object, cfg: FSKConfig): void;
  fskModulate(h: object, data: Uint8Array): Promise<Float32Array>;
  fskDemodulate(h: object, samples: Float32Array): Promise<{ bytes: Uint8Array; eod: number }>;
  fskReset(h: object): void;
  fskStatus(h: object): Record<string, number | boolean>;
  // batch (wam_fsk_batch_*); Int16Array samples take the 16-bit PCM entry (sample = pcm / 32768)
  batchCreate(device: number, nStreams: number, cfgs: FSKConfig[], cfgIndex?: Int32Array): object;
  batchDemodulate(h: object, samples: Float32Array | Int16Array, nSamples: number): Promise<Rows>;
  batchModulate(h: object, data: Uint8Array, nBytes: number, samplesPerRow: number, lengths?: Int32Array):
    Promise<{ samples: Float32Array; lengths: Int32Array; stride: number }>;
  batchStatus(h: object): Record<string, number | boolean>[];
  xmodemBatchCheck(device: number, bytes: Uint8Array, stride: number, lengths: Int32Array, expectedSeq?: Int32Array): Int32Array;
  // session multiplexer (wam_fsk_mux_*): receive half and send half
  muxCreate(device: number, nSessions: number, cfgs: FSKConfig[], cfgIndex: Int32Array | undefined, maxBlock: number): object;
  muxPush(h: object, session: number, quantum: Float32Array): void;
  muxFlush(h: object): Promise<Rows>;
  muxSend(h: object, session: number, data: Uint8Array): void;
  muxModulate(h: object): Promise<void>;
  muxPull(h: object, session: number, sampleCount: number): Chunk | null;
  // batched XModem receiver (wam_xmodem_batch_receive)
  xmodemReceiverCreate(device: number, nSessions: number, maxRetries: number, maxDataBytes?: number): object;
  xmodemReceiverFeed(h: object, bytes: Uint8Array, stride: number, lengths: Int32Array):
    Promise<{ replies: Uint8Array; replyCounts: Int32Array; replyStride: number; consumed: Int32Array; done: Int32Array }>;
  xmodemReceiverData(h: object, session: number): Uint8Array;
};

/** README-only names (README.md:33-38) accepted as aliases of the real FSKConfig fields. */
type FSKConfigInput = Partial<FSKConfig> & { baud?: number; markFreq?: number; spaceFreq?: number };

export class FSKCoreGPU extends BaseModulator<FSKConfig> {
  readonly name = 'FSK';
  readonly type: ModulationType = 'FSK';
  private handle: object | undefined;

  constructor(private readonly device = 0) {
    super();
  }

  configure(config: FSKConfigInput): void {
    const { baud, markFreq, spaceFreq, ...rest } = config;
    this.config = {
      ...DEFAULT_FSK_CONFIG,
      ...(baud !== undefined ? { baudRate: baud } : {}),
      ...(markFreq !== undefined ? { markFrequency: markFreq } : {}),
      ...(spaceFreq !== undefined ? { spaceFrequency: spaceFreq } : {}),
      ...rest,
    } as FSKConfig; // fsk.ts:134 — spread over the defaults, no validation
    this.handle ??= native.fskCreate(this.device);
    native.fskConfigure(this.handle, this.config); // wam_fsk_configure
    this.ready = true;
    this.emit('configured');
  }

  async modulateData(data: Uint8Array): Promise<Float32Array> {
    if (!this.ready || !this.handle) throw new Error('FSK modulator not configured'); // fsk.ts:378-380
    return native.fskModulate(this.handle, data); // wam_fsk_modulate
  }

  /** `samples` is mutated in place when AGC is enabled, exactly like the reference (fsk.ts:55). */
  async demodulateData(samples: Float32Array): Promise<Uint8Array> {
    if (!this.ready || !this.handle) throw new Error('FSK demodulator not configured'); // fsk.ts:191-193
    try {
      const { bytes, eod } = await native.fskDemodulate(this.handle, samples); // wam_fsk_demodulate
      for (let i = 0; i < eod; i++) this.emit('eod'); // fsk.ts:289
      return bytes;
    } catch (error) {
      this.emit('error', { data: error }); // fsk.ts:218-221
      return new Uint8Array(0);
    }
  }

  reset(): void {
    if (this.handle) native.fskReset(this.handle); // fsk.ts:464-469 (ready stays true)
  }

  getStatus() {
    return this.handle ? native.fskStatus(this.handle) : { ready: false }; // fsk.ts:481-493
  }

  getSignalQuality(): SignalQuality {
    return { snr: 0, ber: 0, eyeOpening: 0, phaseJitter: 0, frequencyOffset: 0 }; // fsk.ts:471-479
  }
}

/** Batched entry point: thousands of independent streams with device-resident streaming state. */
export class FSKBatchGPU {
  private readonly handle: object;
  private readonly config: FSKConfig;
  constructor(readonly nStreams: number, configs: FSKConfigInput[] | FSKConfigInput, cfgIndex?: Int32Array, device = 0) {
    const list = (Array.isArray(configs) ? configs : [configs]).map((c) => ({ ...DEFAULT_FSK_CONFIG, ...c }) as FSKConfig);
    if (list.length !== 1 && cfgIndex === undefined) throw new Error('cfgIndex is required with more than one configuration');
    this.config = list[0]; // modulate() supports one configuration per batch
    this.handle = native.batchCreate(device, nStreams, list, cfgIndex);
  }
  /** samples: [nStreams][nSamples] row-major (Float32Array, or Int16Array = 16-bit PCM, half the PCIe bytes);
   *  resolves to the bytes each stream completed in this call. */
  async demodulate(samples: Float32Array | Int16Array, nSamples: number): Promise<Uint8Array[]> {
    const { bytes, lengths, stride } = await native.batchDemodulate(this.handle, samples, nSamples);
    return Array.from(lengths, (n, s) => bytes.subarray(s * stride, s * stride + n));
  }
  /** data: [nStreams][nBytes] row-major, lengths[s] <= nBytes bytes used per stream; resolves to one signal per stream. */
  async modulate(data: Uint8Array, nBytes: number, lengths?: Int32Array): Promise<Float32Array[]> {
    const c = this.config;
    const bitsPerByte = 8 + c.startBits + c.stopBits + (c.parity !== 'none' ? 1 : 0); // fsk.ts:384-386
    const samplesPerBit = Math.floor(c.sampleRate / c.baudRate); // fsk.ts:97
    const perRow = (c.preamblePattern.length + c.sfdPattern.length + nBytes) * bitsPerByte * samplesPerBit + samplesPerBit * 2; // fsk.ts:388-391
    const r = await native.batchModulate(this.handle, data, nBytes, perRow, lengths);
    return Array.from(r.lengths, (n, s) => r.samples.subarray(s * r.stride, s * r.stride + n));
  }
  getStatus() {
    return native.batchStatus(this.handle);
  }
}

/**
 * Session multiplexer (wam_fsk_mux_*): many FSKProcessor.process() callers, each delivering one 128-sample render
 * quantum per call (fsk-processor.ts:152-167), share one ragged GPU batch per tick.  Not executed in this image.
 */
export class FSKSessionMuxGPU {
  private readonly handle: object;
  constructor(readonly nSessions: number, configs: FSKConfigInput[] | FSKConfigInput, cfgIndex?: Int32Array, maxBlock = 1024, device = 0) {
    const list = (Array.isArray(configs) ? configs : [configs]).map((c) => ({ ...DEFAULT_FSK_CONFIG, ...c }) as FSKConfig);
    this.handle = native.muxCreate(device, nSessions, list, cfgIndex, maxBlock); // wam_fsk_mux_create
  }
  /** called from a session's process(): copies the quantum into pinned staging (wam_fsk_mux_push) */
  push(session: number, quantum: Float32Array): void {
    native.muxPush(this.handle, session, quantum);
  }
  /** one ragged batch over the sessions that pushed since the last flush (wam_fsk_mux_flush) */
  async flush(): Promise<Uint8Array[]> {
    const { bytes, lengths, stride } = await native.muxFlush(this.handle);
    return Array.from(lengths, (n: number, s: number) => bytes.subarray(s * stride, s * stride + n));
  }
  /** ChunkedModulator.startModulation(data) of one session (chunked-modulator.ts:31-39); the signal exists after modulate() */
  send(session: number, data: Uint8Array): void {
    native.muxSend(this.handle, session, data);
  }
  /** modulateData() of every session that queued a payload, as one batched GPU call (wam_fsk_mux_modulate) */
  async modulate(): Promise<void> {
    await native.muxModulate(this.handle);
  }
  /** ChunkedModulator.getNextSamples(sampleCount) (chunked-modulator.ts:41-81): the next slice, or null when idle */
  pull(session: number, sampleCount = 128): Chunk | null {
    return native.muxPull(this.handle, session, sampleCount);
  }
}

/**
 * Receive side of XModemTransport for n sessions (wam_xmodem_batch_receive; xmodem.ts:232-321): feed every session's
 * demodulated bytes, get the ACK / NAK bytes to send back and the reassembled payloads.  The transport's timers,
 * its send side and the half-duplex turn-taking stay in XModemTransport on the host.
 */
export class XModemBatchReceiverGPU {
  private readonly handle: object;
  constructor(readonly nSessions: number, maxRetries = 10, device = 0) {
    this.handle = native.xmodemReceiverCreate(device, nSessions, maxRetries);
  }
  /** bursts[s]: the bytes session s demodulated since the last call.  Resolves to the replies per session
   *  (0x06 ACK / 0x15 NAK, in order), how many bytes of each burst were used, and the sessions' done flags. */
  async feed(bursts: Uint8Array[]) {
    const stride = Math.max(1, ...bursts.map((b) => b.length));
    const bytes = new Uint8Array(stride * this.nSessions);
    const lengths = Int32Array.from(bursts, (b) => b.length);
    bursts.forEach((b, s) => bytes.set(b, s * stride));
    const r = await native.xmodemReceiverFeed(this.handle, bytes, stride, lengths);
    return {
      replies: Array.from(r.replyCounts, (n, s) => r.replies.subarray(s * r.replyStride, s * r.replyStride + Math.min(n, r.replyStride))),
      consumed: r.consumed,
      done: r.done, // 0 running, 1 EOT received and ACKed, 2 failed after maxRetries
    };
  }
  received(session: number): Uint8Array {
    return native.xmodemReceiverData(this.handle, session); // assembleData(receive.data), xmodem.ts:323-334
  }
}
