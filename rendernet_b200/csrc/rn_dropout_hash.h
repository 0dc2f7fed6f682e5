// rn_dropout_hash.h -- the counter-based generator behind rn_dropout_16: element i of dropout call `salt` of a step seeded with
// `seed` is kept iff hash(seed, salt, i) < keep * 2^32.  Stateless (the backward pass and the tests recompute the mask), identical
// on host and device.  SplitMix64 finaliser over a 64-bit counter keyed by (seed, salt).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define RN_HD __host__ __device__ __forceinline__
#else
#define RN_HD static inline
#endif

RN_HD uint32_t rn_dropout_hash(uint32_t seed, uint32_t salt, unsigned long long i) {
  const unsigned long long key = ((unsigned long long)seed << 32) | salt;
  unsigned long long z = (i + 1ULL) * 0x9E3779B97F4A7C15ULL + key * 0xD1B54A32D192ED03ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  z ^= z >> 31;
  return (uint32_t)(z >> 32);
}
// keep probability -> threshold in [1, 2^32]
RN_HD unsigned long long rn_dropout_threshold(float keep) {
  const double t = (double)keep * 4294967296.0;
  return t >= 4294967296.0 ? 4294967296ULL : (unsigned long long)t;
}
