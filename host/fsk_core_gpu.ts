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

// eslint-disable-next-line @typescript-eslint/no-var-requires
const native = require('./build/Release/wam_napi.node') as {
  fskCreate(device: number): object;
  fskConfigure(h: object, cfg: FSKConfig): void;
  fskModulate(h: object, data: Uint8Array): Promise<Float32Array>;
  fskDemodulate(h: object, samples: Float32Array): Promise<{ bytes: Uint8Array; eod: number }>;
  fskReset(h: object): void;
  fskStatus(h: object): Record<string, number | boolean>;
  batchCreate(device: number, nStreams: number, cfgs: FSKConfig[], cfgIndex?: Int32Array): object;
  batchDemodulate(h: object, samples: Float32Array, nSamples: number): Promise<{ bytes: Uint8Array; lengths: Int32Array; stride: number }>;
  batchModulate(h: object, data: Uint8Array, nBytes: number): Promise<{ samples: Float32Array; stride: number }>;
  xmodemBatchCheck(device: number, bytes: Uint8Array, stride: number, lengths: Int32Array, expectedSeq?: Int32Array): Int32Array;
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
  constructor(readonly nStreams: number, configs: FSKConfigInput[] | FSKConfigInput, cfgIndex?: Int32Array, device = 0) {
    const list = (Array.isArray(configs) ? configs : [configs]).map((c) => ({ ...DEFAULT_FSK_CONFIG, ...c }) as FSKConfig);
    this.handle = native.batchCreate(device, nStreams, list, cfgIndex);
  }
  /** samples: [nStreams][nSamples] row-major; resolves to the bytes each stream completed in this call. */
  async demodulate(samples: Float32Array, nSamples: number): Promise<Uint8Array[]> {
    const { bytes, lengths, stride } = await native.batchDemodulate(this.handle, samples, nSamples);
    return Array.from(lengths, (n, s) => bytes.subarray(s * stride, s * stride + n));
  }
  async modulate(data: Uint8Array, nBytes: number) {
    return native.batchModulate(this.handle, data, nBytes);
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
  async feed(bursts: Uint8Array[]): Promise<Uint8Array[]> {
    return native.xmodemReceiverFeed(this.handle, bursts); // replies per session: 0x06 ACK / 0x15 NAK in order
  }
  received(session: number): Uint8Array {
    return native.xmodemReceiverData(this.handle, session); // assembleData(receive.data), xmodem.ts:323-334
  }
}
