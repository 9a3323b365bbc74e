/* TEST INFRASTRUCTURE.  Exhaustive check of the expf restatement (expf_glibc.h) against the libm
 * of this box over EVERY float bit pattern in [lo, hi) (default: all 2^32).  Build + run:
 *   gcc -O2 -ffp-contract=off -fopenmp expf_sweep.c -lm -o _build/expf_sweep && _build/expf_sweep
 * Prints mismatches per evaluation order (see expf_glibc.h). */
#include <stdio.h>
#include <stdlib.h>
#include "expf_glibc.h"
int main(int argc, char** argv) {
  uint64_t lo = argc > 1 ? strtoull(argv[1], 0, 0) : 0, hi = argc > 2 ? strtoull(argv[2], 0, 0) : (1ull << 32);
  uint64_t bad0 = 0, bad1 = 0, bad2 = 0;
#pragma omp parallel for reduction(+ : bad0, bad1, bad2) schedule(static)
  for (uint64_t u = lo; u < hi; ++u) {
    uint32_t b = (uint32_t)u;
    float x; memcpy(&x, &b, 4);
    float ref = expf(x), a = hz_expf_glibc(x, 0), f = hz_expf_glibc(x, 1), g = hz_expf_glibc(x, 2);
    uint32_t rb = hz_asuint(ref);
    int refnan = ref != ref;
    if (refnan ? !(a != a) : rb != hz_asuint(a)) bad0++;
    if (refnan ? !(f != f) : rb != hz_asuint(f)) bad1++;
    if (refnan ? !(g != g) : rb != hz_asuint(g)) bad2++;
  }
  printf("patterns=%llu mismatch_vs_libm: variant0(sse2)=%llu variant1(fma-reduce)=%llu variant2(fma-all)=%llu\n",
         (unsigned long long)(hi - lo), (unsigned long long)bad0, (unsigned long long)bad1, (unsigned long long)bad2);
  return (bad1 && bad0) ? 1 : 0;
}
