// Host-only helpers of libhzb200 (no device code; compiled by the host compiler alone).
//
// hz_host_random_legal: the host-side stand-in for a caller's policy in a host-driven env loop (EnvPipeline): a
// uniformly random legal move per game from the legal-mask word of the packed rows the env kernel wrote to pinned
// host memory.  Counter-based (seed, game, step), so any step can be regenerated.  The k-th set bit of the mask comes
// from BMI2 `pdep` where the CPU has it (3 ns per game), else from a 256 x 8 select table (same results).
#include <stddef.h>
#include <stdint.h>

#include "../../include/hzb200.h"

namespace hz {
void set_error(const char* fmt, ...);
}

namespace {

struct SelTables {
  uint8_t pop[256];
  uint8_t sel[256][8];   // sel[v][k] = index of the k-th set bit of v
  SelTables() {
    for (int v = 0; v < 256; ++v) {
      int n = 0;
      for (int b = 0; b < 8; ++b) {
        sel[v][b] = 0;
        if ((v >> b) & 1) sel[v][n++] = (uint8_t)b;
      }
      pop[v] = (uint8_t)n;
    }
  }
};
const SelTables kSel;

inline uint64_t mix(uint64_t seed, uint64_t game, uint64_t step_key) {   // splitmix64 finaliser over (seed, game, step)
  uint64_t x = seed + 0x9E3779B97F4A7C15ull * (game + 1ull) + step_key;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x;
}

void pick_table(const uint32_t* rows, int64_t ld, int word, int n, uint32_t amask, uint64_t seed, uint64_t step_key,
                int32_t* out) {
  for (int i = 0; i < n; ++i) {
    const uint32_t m = rows[(size_t)i * ld + word] & amask;
    const uint64_t x = mix(seed, (uint64_t)i, step_key);
    // k-th set bit of m without data-dependent branches: per-byte counts, then the select table
    const uint32_t b0 = m & 255u, b1 = (m >> 8) & 255u, b2 = (m >> 16) & 255u, b3 = m >> 24;
    const uint32_t c0 = kSel.pop[b0], c1 = c0 + kSel.pop[b1], c2 = c1 + kSel.pop[b2], c3 = c2 + kSel.pop[b3];
    const uint32_t k = (uint32_t)(((x >> 32) * c3) >> 32);     // uniform in [0, popcount)
    const uint32_t byte = (k >= c0) + (k >= c1) + (k >= c2);
    const uint32_t base = byte == 0 ? 0u : (byte == 1 ? c0 : (byte == 2 ? c1 : c2));
    const uint32_t bv = (m >> (8 * byte)) & 255u;
    out[i] = c3 ? (int32_t)(8 * byte + kSel.sel[bv][(k - base) & 7u]) : 0;
  }
}

#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target("bmi2,popcnt"))) void pick_bmi2(const uint32_t* rows, int64_t ld, int word, int n, uint32_t amask,
                                                       uint64_t seed, uint64_t step_key, int32_t* out) {
  for (int i = 0; i < n; ++i) {
    const uint32_t m = rows[(size_t)i * ld + word] & amask;
    const uint64_t x = mix(seed, (uint64_t)i, step_key);
    const uint32_t c = (uint32_t)__builtin_popcount(m);
    const uint32_t k = (uint32_t)(((x >> 32) * c) >> 32);      // uniform in [0, popcount)
    out[i] = c ? (int32_t)__builtin_ctz(__builtin_ia32_pdep_si(1u << k, m)) : 0;   // deposit bit k at the k-th set bit
  }
}
#endif

}  // namespace

extern "C" {
#pragma GCC visibility push(default)

int hz_host_random_legal(const uint32_t* rows, int64_t ld_words, int legal_word, int num_games, int num_actions,
                         uint64_t seed, uint32_t step, int32_t* out_actions) {
  if (!rows || !out_actions || num_games < 0 || num_actions <= 0 || num_actions > 32 || legal_word < 0 || ld_words <= legal_word) {
    hz::set_error("hz_host_random_legal: bad argument");
    return HZ_ERR_ARG;
  }
  const uint32_t amask = num_actions >= 32 ? 0xffffffffu : ((1u << num_actions) - 1u);
  const uint64_t step_key = (uint64_t)step << 32;
#if defined(__x86_64__) && defined(__GNUC__)
  static const bool have_bmi2 = __builtin_cpu_supports("bmi2") && __builtin_cpu_supports("popcnt");
  if (have_bmi2) {
    pick_bmi2(rows, ld_words, legal_word, num_games, amask, seed, step_key, out_actions);
    return HZ_OK;
  }
#endif
  pick_table(rows, ld_words, legal_word, num_games, amask, seed, step_key, out_actions);
  return HZ_OK;
}

#pragma GCC visibility pop
}
