// Micro-benchmark (B200): warp-instruction throughput per SM and dependent-chain latency of the pipes the demodulator
// leans on, alone and in the mixes the kernels issue.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ubench_pipes scripts/ubench_pipes.cu && gpurun_out/ubench_pipes
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 2048, CH = 8;

template <int OP, int NCH>
__global__ void k(float* out, float a, float b, long long* cyc) {
  float x[CH]; float2 y[CH]; double z[CH]; unsigned u[CH];
  for (int i = 0; i < CH; i++) { x[i] = threadIdx.x * 1e-3f + i + 1; y[i] = make_float2(x[i], x[i] + 1); z[i] = x[i]; u[i] = threadIdx.x + i; }
  const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
  const double ad = a, bd = b;
  const long long c0 = clock64();
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < NCH; i++) {
      if (OP == 0) x[i] = fmaf(x[i], a, b);                       // FFMA
      if (OP == 1) y[i] = __ffma2_rn(y[i], a2, b2);               // FFMA2
      if (OP == 2) z[i] = fma(z[i], ad, bd);                      // DFMA
      if (OP == 3) { float r; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x[i])); x[i] = r; }  // MUFU.RCP
      if (OP == 4) { z[i] = (double)x[i]; x[i] = (float)z[i]; }   // F2F.F64.F32 + F2F.F32.F64
      if (OP == 5) x[i] = fmaxf(x[i], b) ;                        // FMNMX
      if (OP == 6) { float r; asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x[i])); x[i] = r; }  // MUFU.RSQ
      if (OP == 7) u[i] = u[i] * 3u + 7u;                         // IMAD
      if (OP == 8) x[i] = x[i] > a ? b : x[i];                    // FSETP + FSEL
      if (OP == 9) u[i] = (u[i] & 0xff00ffu) ^ (u[i] >> 3);       // SHF + LOP3
      if (OP == 10) y[i] = __fmul2_rn(y[i], a2);                  // FMUL2
      if (OP == 11) y[i] = __fadd2_rn(y[i], a2);                  // FADD2
      if (OP == 12) z[i] = z[i] + ad;                             // DADD
      if (OP == 13) { x[i] = fmaf(x[i], a, b); float r; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(y[i].x)); y[i].x = r; }  // FFMA + MUFU
      if (OP == 14) { x[i] = fmaf(x[i], a, b); z[i] = fma(z[i], ad, bd); }  // FFMA + DFMA
      if (OP == 15) { x[i] = fmaf(x[i], a, b); u[i] = u[i] * 3u + 7u; }      // FFMA + IMAD
      if (OP == 16) { x[i] = fmaf(x[i], a, b); u[i] = (u[i] & 0xff00ffu) ^ (u[i] >> 3); }  // FFMA + SHF + LOP3
      if (OP == 17) { y[i] = __ffma2_rn(y[i], a2, b2); u[i] = (u[i] & 0xff00ffu) ^ (u[i] >> 3); }  // FFMA2 + SHF + LOP3
      if (OP == 18) u[i] = __float_as_uint((float)(int)u[i]) + 1u; // I2F + IADD
      if (OP == 19) u[i] = (unsigned)__float2int_rn(__uint_as_float((u[i] & 0xffffu) | 0x40000000u)) + 1u;  // F2I
      if (OP == 20) x[i] = x[i] * a;                              // FMUL
      if (OP == 21) u[i] = __popc(u[i]) + u[i];                   // POPC + IADD
      if (OP == 22) { x[i] = fmaf(x[i], a, b); x[i] = fmaxf(x[i], b); }  // FFMA + FMNMX
    }
  }
  const long long c1 = clock64();
  float s = 0;
  for (int i = 0; i < CH; i++) s += x[i] + y[i].x + y[i].y + (float)z[i] + (float)u[i];
  if (s == 12345.678f) out[0] = s;
  if (cyc && threadIdx.x == 0 && blockIdx.x == 0) *cyc = c1 - c0;
}

template <int OP>
void run(const char* name, int ops_per) {
  float* d; cudaMalloc(&d, 4);
  long long* dc; cudaMalloc(&dc, 8);
  int dev_sms = 148; cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, 0);
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const int blocks = dev_sms * 8, threads = 256;
  k<OP, CH><<<blocks, threads>>>(d, 1.0001f, 0.5f, nullptr);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<OP, CH><<<blocks, threads>>>(d, 1.0001f, 0.5f, nullptr);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double warp_instr = (double)blocks * threads / 32 * ITERS * CH * ops_per;
  const double per_sm_per_clk = warp_instr / dev_sms / (ms * 1e-3 * clk_khz * 1e3);
  // latency: one warp, one dependent chain
  k<OP, 1><<<1, 32>>>(d, 1.0001f, 0.5f, dc);
  long long c = 0; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
  printf("%-34s %8.3f ms  %6.3f warp-instr/clk/SM  chain %6.2f clk/iter (%d instr)\n", name, ms, per_sm_per_clk,
         (double)c / ITERS, ops_per);
  cudaFree(d); cudaFree(dc);
}

int main() {
  run<0>("FFMA", 1);
  run<20>("FMUL", 1);
  run<1>("FFMA2 (f32x2)", 1);
  run<10>("FMUL2", 1);
  run<11>("FADD2", 1);
  run<2>("DFMA", 1);
  run<12>("DADD", 1);
  run<3>("MUFU.RCP f32", 1);
  run<6>("MUFU.RSQ f32", 1);
  run<4>("F2F f32->f64 + f64->f32", 2);
  run<5>("FMNMX", 1);
  run<7>("IMAD", 1);
  run<8>("FSETP + FSEL", 2);
  run<9>("SHF + LOP3", 2);
  run<18>("I2F + IADD", 2);
  run<19>("LOP3 + F2I + IADD", 3);
  run<21>("POPC + IADD", 2);
  run<13>("FFMA + MUFU.RCP", 2);
  run<14>("FFMA + DFMA", 2);
  run<15>("FFMA + IMAD", 2);
  run<16>("FFMA + SHF + LOP3", 3);
  run<17>("FFMA2 + SHF + LOP3", 3);
  run<22>("FFMA + FMNMX", 2);
  return 0;
}
