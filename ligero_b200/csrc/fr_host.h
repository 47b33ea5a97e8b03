// Host-side BN254 Fr helpers (table generation, sponge, circuit evaluation).  Uses the plain-C
// emulation of the same carry-chain statements as the device code (fr.cuh), so host and device
// arithmetic are one algorithm.
#pragma once
#include <stdint.h>
#include <string.h>

#include "fr.cuh"

namespace lg {

// canonical integer (< r) in limbs -> Montgomery form
inline Fr fr_to_mont(const Fr& canonical) { return fr_mul(canonical, fr_r2()); }

inline Fr fr_from_u64(uint64_t x) {
  Fr c = fr_zero();
  c.v[0] = (uint32_t)x;
  c.v[1] = (uint32_t)(x >> 32);
  return fr_to_mont(c);
}

inline Fr fr_pow_u64(Fr base, uint64_t e) {
  Fr acc = fr_one();
  while (e) {
    if (e & 1) acc = fr_mul(acc, base);
    base = fr_sqr(base);
    e >>= 1;
  }
  return acc;
}

// exponent given as 8 x u32 little-endian limbs
inline Fr fr_pow_limbs(const Fr& base, const uint32_t e[8]) {
  Fr acc = fr_one();
  for (int i = 255; i >= 0; i--) {
    acc = fr_sqr(acc);
    if ((e[i >> 5] >> (i & 31)) & 1) acc = fr_mul(acc, base);
  }
  return acc;
}

inline Fr fr_inv(const Fr& a) {  // a^(r-2)
  uint32_t e[8] = {LG_P0 - 2u, LG_P1, LG_P2, LG_P3, LG_P4, LG_P5, LG_P6, LG_P7};
  return fr_pow_limbs(a, e);
}

// 2-adic root of unity of order 2^28: 5^((r-1)/2^28)   (arkworks FrConfig::TWO_ADIC_ROOT_OF_UNITY)
inline Fr fr_two_adic_root() {
  // (r-1) >> 28
  uint32_t p[8] = {LG_P0 - 1u, LG_P1, LG_P2, LG_P3, LG_P4, LG_P5, LG_P6, LG_P7};
  uint32_t e[8];
  for (int i = 0; i < 8; i++) {
    uint32_t lo = p[i] >> 28;
    uint32_t hi = (i + 1 < 8) ? (p[i + 1] << 4) : 0u;
    e[i] = lo | hi;
  }
  return fr_pow_limbs(fr_from_u64(5), e);
}

// generator of the size-2^log_n radix-2 domain (GeneralEvaluationDomain::new(2^log_n).group_gen)
inline Fr fr_root_of_unity(int log_n) {
  Fr w = fr_two_adic_root();
  for (int i = log_n; i < 28; i++) w = fr_sqr(w);
  return w;
}

inline bool fr_is_canonical(const Fr& a) {  // a < r
  for (int i = 7; i >= 0; i--) {
    if (a.v[i] < fr_p(i)) return true;
    if (a.v[i] > fr_p(i)) return false;
  }
  return false;
}

}  // namespace lg
