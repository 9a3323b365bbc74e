/* TEST INFRASTRUCTURE.  Restatement of glibc >= 2.27 expf (sysdeps/ieee754/flt-32/e_expf.c with
 * e_exp2f_data.c: N = 32 table, cubic in double; the algorithm of ARM "optimized-routines"
 * v18.x, which glibc 2.27+ ships).  glibc is a third-party dependency of the reference, not under
 * /root/reference: the call site is cnode.cpp:87 (`exp(float)` -> expf, confirmed by `nm`).
 * The CUDA path carries its own device copy of this algorithm (hanabizero_b200/csrc/hz_math.cuh);
 * tests/test_oracle_expf.py + oracle/expf_sweep.c check THIS copy against the libm of the box,
 * bit for bit.  The table is generated, not typed: tab[i] = bits(2^(i/32)) - (i << 47). */
#ifndef HZ_ORACLE_EXPF_GLIBC_H
#define HZ_ORACLE_EXPF_GLIBC_H
#include <math.h>
#include <stdint.h>
#include <string.h>

static const uint64_t hz_exp2f_tab[32] = {
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,
    0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,
    0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,
    0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull,
};

static inline uint32_t hz_asuint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline uint64_t hz_asuint64(double f) { uint64_t u; memcpy(&u, &f, 8); return u; }
static inline double hz_asdouble(uint64_t u) { double f; memcpy(&f, &u, 8); return f; }

/* Two evaluation orders exist in x86-64 glibc, picked at load time by ifunc
 * (sysdeps/x86_64/fpu/multiarch/e_expf.c): __expf_fma (built with -mfma, the compiler contracts
 * `r = InvLn2N*xd - kd` into one fused multiply-subtract) on every CPU with FMA3, and
 * __expf_sse2 (no contraction) otherwise.  variant: 0 = sse2 order, 1 = fma order (reduction
 * fused, polynomial separate), 2 = fma order with the polynomial fused as well.
 * Exhaustive sweep on this box (oracle/expf_sweep.c, all 2^32 inputs): variants 1 and 2 are
 * identical to each other and to libm everywhere; variant 0 differs from them on exactly two
 * inputs, 0x4202422f (x=32.56) and 0xc27c65d9 (x=-63.10), by one ulp.  The CUDA path implements
 * variant 1 (what any FMA-capable host, i.e. every GPU box, runs). */
static inline float hz_expf_glibc(float x, int variant) {
  const double N = 32.0;
  const double InvLn2N = 0x1.71547652b82fep+0 * N;
  const double SHIFT = 0x1.8p+52;
  const double C0 = 0x1.c6af84b912394p-5 / N / N / N;
  const double C1 = 0x1.ebfce50fac4f3p-3 / N / N;
  const double C2 = 0x1.62e42ff0c52d6p-1 / N;
  double xd = (double)x;
  uint32_t abstop = (hz_asuint(x) >> 20) & 0x7ff;
  if (abstop >= (hz_asuint(88.0f) >> 20)) {
    if (hz_asuint(x) == hz_asuint(-INFINITY)) return 0.0f;
    if (abstop >= (hz_asuint(INFINITY) >> 20)) return x + x;
    if (x > 0x1.62e42ep6f) return INFINITY;            /* __math_oflowf(0) */
    if (x < -0x1.9fe368p6f) return 0.0f;               /* __math_uflowf(0) */
    if (x < -0x1.9d1d9ep6f) return 0x1p-149f;          /* __math_may_uflowf(0): 0x1.4p-75f squared */
  }
  double z = InvLn2N * xd;
  volatile double kdv = z + SHIFT; /* math_narrow_eval */
  double kd = kdv;
  uint64_t ki = hz_asuint64(kd);
  kd -= SHIFT;
  double r = variant ? fma(InvLn2N, xd, -kd) : z - kd;
  uint64_t t = hz_exp2f_tab[ki % 32];
  t += ki << (52 - 5);
  double s = hz_asdouble(t);
  double y;
  if (variant == 2) {
    z = fma(C0, r, C1);
    double r2 = r * r;
    y = fma(C2, r, 1.0);
    y = fma(z, r2, y);
  } else {
    z = C0 * r + C1;
    double r2 = r * r;
    y = C2 * r + 1.0;
    y = z * r2 + y;
  }
  y = y * s;
  return (float)y;
}
#endif
