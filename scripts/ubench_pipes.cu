// Micro-benchmark (B200): warp-instruction throughput per SM of the pipes the demodulator leans on.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ubench_pipes scripts/ubench_pipes.cu && gpurun_out/ubench_pipes
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096, CH = 8;

template <int OP>
__global__ void k(float* out, float a, float b) {
  float x[CH]; float2 y[CH]; double z[CH];
  for (int i = 0; i < CH; i++) { x[i] = threadIdx.x * 1e-3f + i; y[i] = make_float2(x[i], x[i] + 1); z[i] = x[i]; }
  const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < CH; i++) {
      if (OP == 0) x[i] = fmaf(x[i], a, b);                       // FFMA (3 register operands)
      if (OP == 1) y[i] = __ffma2_rn(y[i], a2, b2);               // FFMA2
      if (OP == 2) z[i] = fma(z[i], (double)a, (double)b);        // DFMA
      if (OP == 3) { float r; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x[i])); x[i] = r; }  // MUFU.RCP
      if (OP == 4) { z[i] = (double)x[i]; x[i] = (float)z[i] + a; }  // F2F.F64.F32 + F2F.F32.F64 (+ FADD)
      if (OP == 5) x[i] = fmaxf(x[i] + a, b);                      // FADD + FMNMX
      if (OP == 6) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(z[i])); z[i] = r; }  // MUFU.RCP64H
      if (OP == 7) x[i] = __int_as_float(__float_as_int(x[i]) + 3);  // IADD
    }
  }
  float s = 0;
  for (int i = 0; i < CH; i++) s += x[i] + y[i].x + y[i].y + (float)z[i];
  if (s == 12345.678f) out[0] = s;
}

template <int OP>
void run(const char* name, int ops_per) {
  float* d; cudaMalloc(&d, 4);
  int dev_sms = 148; cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, 0);
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const int blocks = dev_sms * 8, threads = 256;
  k<OP><<<blocks, threads>>>(d, 1.0001f, 0.5f);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<OP><<<blocks, threads>>>(d, 1.0001f, 0.5f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double warp_instr = (double)blocks * threads / 32 * ITERS * CH * ops_per;
  const double per_sm_per_clk = warp_instr / dev_sms / (ms * 1e-3 * clk_khz * 1e3);
  printf("%-28s %8.3f ms  %6.3f warp-instr/clk/SM (at %d MHz nominal)\n", name, ms, per_sm_per_clk, clk_khz / 1000);
  cudaFree(d);
}

int main() {
  run<0>("FFMA", 1);
  run<1>("FFMA2 (f32x2)", 1);
  run<2>("DFMA", 1);
  run<3>("MUFU.RCP f32", 1);
  run<4>("F2F f32->f64 + f64->f32 (+FADD)", 2);
  run<5>("FADD + FMNMX", 2);
  run<6>("MUFU.RCP64H", 1);
  run<7>("IADD", 1);
  return 0;
}
