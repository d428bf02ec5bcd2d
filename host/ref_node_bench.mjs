// ref_node_bench.mjs — the baseline the north star names: the UNMODIFIED reference FSKCore under
// Node worker_threads, one worker per host core, demodulating its share of synthetic streams.
// NOT EXECUTED in the build environment (no Node in the image or on the GPU boxes); bench.py times the
// C float64 port of the same algorithm (oracle/) on all host cores instead and labels it "port".
//
//   node --experimental-transform-types --import ./host/alias-hook.mjs host/ref_node_bench.mjs <reference-root> [streams]
// (alias-hook.mjs must resolve the reference's '@/...' imports to '<reference-root>/src/...'.)
import { Worker, isMainThread, parentPort, workerData } from 'node:worker_threads';
import os from 'node:os';
import path from 'node:path';

const root = process.argv[2] ?? workerData?.root;
const N_SAMPLES = 48000;

if (isMainThread) {
  const cores = os.availableParallelism();
  const streams = Number(process.argv[3] ?? cores * 8);
  const t0 = performance.now();
  const done = await Promise.all(
    Array.from({ length: cores }, (_, w) => new Promise((resolve) => {
      const worker = new Worker(new URL(import.meta.url), { workerData: { root, w, cores, streams }, execArgv: process.execArgv });
      worker.on('message', resolve);
    })),
  );
  const dt = (performance.now() - t0) / 1e3;
  const bytes = done.reduce((a, b) => a + b, 0);
  console.log(JSON.stringify({ impl: 'reference-node', cores, streams, msamples_per_s: (streams * N_SAMPLES) / dt / 1e6, decoded_bits_per_s: (bytes * 8) / dt }));
} else {
  const { FSKCore, DEFAULT_FSK_CONFIG } = await import(path.join(workerData.root, 'src/modems/fsk.ts'));
  const { w, cores, streams } = workerData;
  let bytes = 0;
  for (let s = w; s < streams; s += cores) {
    const cfg = { ...DEFAULT_FSK_CONFIG, baudRate: 300, markFrequency: s * 2 < streams ? 980 : 1650, spaceFrequency: s * 2 < streams ? 1180 : 1850 };
    const tx = new FSKCore(); tx.configure(cfg);
    const frame = await tx.modulateData(Uint8Array.from({ length: 25 }, (_, i) => (s * 31 + i * 7) & 0xff));
    const x = new Float32Array(N_SAMPLES);
    x.set(frame.subarray(0, Math.min(frame.length, N_SAMPLES - (s % 1280))), s % 1280);
    const sigma = Math.sqrt(0.5 / 10 ** ((-15 + 3 * (s % 16)) / 10));
    for (let i = 0; i < N_SAMPLES; i += 2) { // Box-Muller
      const u = Math.sqrt(-2 * Math.log(Math.random() || 1e-12)), v = 2 * Math.PI * Math.random();
      x[i] += sigma * u * Math.cos(v); if (i + 1 < N_SAMPLES) x[i + 1] += sigma * u * Math.sin(v);
    }
    const rx = new FSKCore(); rx.configure(cfg);
    bytes += (await rx.demodulateData(x)).length;
  }
  parentPort.postMessage(bytes);
}
