"""CPU: the host-only helpers of libhzb200 (hz_host.cpp) need no GPU — hz_host_random_legal must pick, for every game,
the k-th legal move with k drawn from the documented counter-based generator (splitmix64 finaliser over
(seed, game, step)), whichever code path (BMI2 pdep or the select table) the CPU takes."""
import numpy as np


def _expected(mask, seed, game, step):
    M = (1 << 64) - 1
    x = (seed + 0x9E3779B97F4A7C15 * (game + 1) + (step << 32)) & M
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M
    bits = [b for b in range(32) if (mask >> b) & 1]
    if not bits:
        return 0
    return bits[((x >> 32) * len(bits)) >> 32]


def test_host_random_legal_matches_its_definition():
    from hanabizero_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(0)
    for n, a, ld, word in ((257, 20, 4, 0), (64, 11, 29, 25), (33, 32, 1, 0)):
        rows = rng.integers(0, 2 ** 32, size=(n, ld), dtype=np.uint64).astype(np.uint32)
        rows[3, word] = 0                      # a game without legal moves (finished): action 0
        out = np.full(n, -1, np.int32)
        for seed, step in ((0, 0), (12345678901234567, 77), (2 ** 64 - 1, 2 ** 32 - 1)):
            _lib.check(lib.hz_host_random_legal(rows.ctypes.data, ld, word, n, a, seed, step, out.ctypes.data))
            amask = (1 << a) - 1
            want = [_expected(int(rows[i, word]) & amask, seed, i, step) for i in range(n)]
            assert out.tolist() == want
