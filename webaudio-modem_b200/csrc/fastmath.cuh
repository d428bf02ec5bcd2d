// fastmath.cuh — float64 primitives for the demodulator's decimated-rate discriminator.
//
// CUDA's atan2()/sqrt()/division are IEEE-grade but cost ~130 / ~26 / ~20 instructions.  The
// discriminator only needs a few ulp (DESIGN.md "numerics": an error e in the phase flips a hard
// decision with probability ~40*e per decimated sample), so these versions trade the last ulp for
// instruction count:
//   fast_rcp    MUFU.RCP64H seed (2^-23) + two Newton steps            -> <= 2 ulp
//   fast_sqrt   MUFU.RSQ64H seed + two coupled Goldschmidt steps       -> <= 2 ulp
//   fast_atan2  octant fold, 64-interval table rotation (c_k = k/64, atan(c_k) from the host's
//               libm), one reciprocal, degree-7 odd polynomial on |t| <= 1/120 -> <= 6 ulp
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace wam {

constexpr int kAtanTableSize = 65;  // c_k = k / 64, k = 0..64

__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
}

// sqrt(p) for p >= 0 (denormals and zero give 0)
__device__ __forceinline__ double fast_sqrt(double p) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(p));
  double g = p * y;
  double h = 0.5 * y;
  double r = fma(-g, h, 0.5);
  g = fma(g, r, g);
  h = fma(h, r, h);
  r = fma(-g, h, 0.5);
  g = fma(g, r, g);
  return p > 1e-290 ? g : 0.0;
}

// atan2(y, x) with JS/libm quadrant conventions (including signed zeros); NaN/Inf are not handled
// specially (the discriminator never produces them from finite samples).
// tab[k] = {k / 64, atan(k / 64)}, k = 0..64 (device global memory, read through L1).
__device__ __forceinline__ double fast_atan2(double y, double x, const double2* __restrict__ tab) {
  const int hx = __double2hiint(x), hy = __double2hiint(y);
  const double ax = __hiloint2double(hx & 0x7fffffff, __double2loint(x));
  const double ay = __hiloint2double(hy & 0x7fffffff, __double2loint(y));
  const bool swap = ay > ax;
  double mx = swap ? ay : ax;
  double mn = swap ? ax : ay;
  {
    // rare: zero / denormal / tiny magnitudes — renormalise (the angle is scale invariant): one select on the scale
    // factor's high word instead of two on each operand
    const double sc = __hiloint2double(__double2hiint(mx) < 0x06000000 ? 0x78300000 : 0x3ff00000, 0);  // 2^900 : 1
    mx *= sc; mn *= sc;
  }
  // coarse ratio from the reciprocal seed picks the table interval
  double r0;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(mx));
  // round(64 * mn * r0) from the low word of (x + 1.5 * 2^52): no F2I; 0 <= mn * r0 <= 1 + 2^-22, NaN (mx == 0) gives
  // a garbage index that is clamped and whose result is discarded below
  int k = __double2loint(fma(mn * r0, 64.0, 6755399441055744.0));
  k = max(0, min(k, 64));
  const double2 e = __ldg(tab + k);
  // rotate by -atan(c): t = (mn - c*mx) / (mx + c*mn), |t| <= ~1/120
  const double xr = fma(e.x, mn, mx);
  const double yr = fma(-e.x, mx, mn);
  const double t = yr * fast_rcp(xr);
  const double z = t * t;
  double p = fma(z, -0.14285714285714285, 0.2);
  p = fma(z, p, -0.3333333333333333);
  p = p * z;
  double a = e.y + fma(t, p, t);
  a = (mx > 0.0) ? a : 0.0;
  // quadrant: swap -> pi/2 - a ; x < 0 -> pi - a  (sign bit of x, so that atan2(+-0, -0) = +-pi)
  const bool xneg = hx < 0;
  {
    // base = swap ? pi/2 : (xneg ? pi : 0): pi and pi/2 share their low word (0x54442D18), so three integer
    // selects; "base - a" is "base + (-a)": flip a's sign bit instead of selecting between two sums
    const int base_hi = swap ? 0x3FF921FB : (xneg ? 0x400921FB : 0);
    const int base_lo = (swap || xneg) ? 0x54442D18 : 0;
    const int flip = (swap != xneg) ? (int)0x80000000 : 0;
    const double as = __hiloint2double(__double2hiint(a) ^ flip, __double2loint(a));
    a = __hiloint2double(base_hi, base_lo) + as;
  }
  return copysign(a, y);
}

// test hook: out[i] = fast_atan2(y[i], x[i]); out2[i] = fast_sqrt(|x[i]|); out3[i] = fast_rcp(x[i])
__global__ void fastmath_debug_kernel(const double* __restrict__ y, const double* __restrict__ x, long n,
                                      const double2* __restrict__ tab, double* out_atan2, double* out_sqrt,
                                      double* out_rcp) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out_atan2[i] = fast_atan2(y[i], x[i], tab);
  out_sqrt[i] = fast_sqrt(fabs(x[i]));
  out_rcp[i] = fast_rcp(x[i]);
}

}  // namespace wam
