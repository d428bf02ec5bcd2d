// awgn.cuh — additive white Gaussian noise on device-resident sample rows: counter-based Philox4x32-10 + Box-Muller,
// four samples per thread and Philox call, added in place (one 16-byte load and store per four samples).  The channel
// model of BASELINE config 5 (modulate -> AWGN -> demodulate) and of the synthetic workloads; not part of FSKCore.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace wam {

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
  }
  return c;
}

__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
  const float u1 = ((float)(a >> 8) + 0.5f) * (1.0f / 16777216.0f);  // (0, 1)
  const float u2 = (float)(b >> 8) * (1.0f / 16777216.0f);            // [0, 1)
  const float r = sqrtf(-2.0f * __logf(u1));
  float s, c;
  __sincosf(6.283185307179586f * u2, &s, &c);
  return make_float2(r * c, r * s);
}

// samples[row][0..n) += sigma[row] * N(0, 1); stride and n multiples of 4, rows 16-byte aligned.
// Counter = (column / 4, row, stream of calls `seq`), key = seed: every (seed, seq, row, column) has its own noise.
__global__ void awgn_add_kernel(float* __restrict__ samples, long stride, long n_rows, long n, const float* __restrict__ sigma,
                                uint64_t seed, uint32_t seq) {
  const long quads = n / 4;
  const long total = n_rows * quads;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long row = i / quads, q = i - row * quads;
    const uint4 r = philox4x32_10(make_uint4((uint32_t)q, (uint32_t)(q >> 32), (uint32_t)row, seq),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const float2 z0 = box_muller(r.x, r.y), z1 = box_muller(r.z, r.w);
    const float sg = sigma[row];
    float4* p = reinterpret_cast<float4*>(samples + row * stride) + q;
    float4 v = *p;
    v.x = fmaf(sg, z0.x, v.x); v.y = fmaf(sg, z0.y, v.y); v.z = fmaf(sg, z1.x, v.z); v.w = fmaf(sg, z1.y, v.w);
    *p = v;
  }
}

}  // namespace wam
