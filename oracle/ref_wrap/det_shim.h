/* TEST INFRASTRUCTURE (oracle). Force-included (-include) when compiling the reference
 * ctree sources for the deterministic oracle build.  It neutralises the only source of
 * non-determinism in the reference tree engine without editing reference sources:
 *   - cmulti_traverse reseeds srand(gettimeofday().tv_usec) on every call
 *     (/root/reference/core/ctree/cnode.cpp:409-411)
 *   - cselect_child breaks epsilon-ties with rand() % ties (cnode.cpp:367-369)
 * With rand() == 0 the pick is the first index attaining the strict maximum — the parity
 * contract of SURVEY.md §7.4-3 / Appendix A.7. */
#ifndef HZ_ORACLE_DET_SHIM_H
#define HZ_ORACLE_DET_SHIM_H
#ifdef __cplusplus
#include <cstdlib>
#endif
#include <stdlib.h>
static inline int hz_oracle_rand0(void) { return 0; }
#define rand hz_oracle_rand0
#define srand(x) ((void)(x))
#endif
