// Host-side ChaCha20 stream (rand_chacha 0.3 ChaCha20Rng: 64-bit counter, stream 0) for the small
// challenge derivations that stay on the host: the t distinct column indices (src/utils.rs:31-55) and
// short field vectors.  The long vectors (4mk elements) are expanded on the device (protocol.cu).
#pragma once
#include <stdint.h>
#include <string.h>

#include <set>
#include <vector>

#include "fr_host.h"

namespace lg {

class ChaChaRng {
 public:
  ChaChaRng(const uint8_t seed[32], int rounds = 20) : rounds_(rounds) { memcpy(key_, seed, 32); }
  uint32_t next_u32() {
    if (pos_ == 16) {
      block();
      pos_ = 0;
    }
    return buf_[pos_++];
  }
  uint64_t next_u64() {
    const uint64_t lo = next_u32();
    const uint64_t hi = next_u32();
    return lo | (hi << 32);
  }

 private:
  static uint32_t rotl(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }
  void block() {
    uint32_t st[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, key_[0], key_[1], key_[2], key_[3],
                       key_[4], key_[5], key_[6], key_[7], (uint32_t)ctr_, (uint32_t)(ctr_ >> 32), 0, 0};
    uint32_t x[16];
    memcpy(x, st, sizeof x);
    auto qr = [&](int a, int b, int c, int d) {
      x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 16); x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 12);
      x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 8);  x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 7);
    };
    for (int r = 0; r < rounds_; r += 2) {
      qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15);
      qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14);
    }
    for (int i = 0; i < 16; i++) buf_[i] = x[i] + st[i];
    ctr_++;
  }
  uint32_t key_[8];
  uint32_t buf_[16];
  uint64_t ctr_ = 0;
  int pos_ = 16;
  int rounds_;
};

// ark-ff `Fp::rand`: 4 x next_u64 -> limbs, mask the top two bits, reject if >= r; the accepted integer is
// the Montgomery representation itself
inline Fr fr_rand(ChaChaRng& rng) {
  for (;;) {
    Fr e;
    for (int i = 0; i < 4; i++) {
      const uint64_t w = rng.next_u64();
      e.v[2 * i] = (uint32_t)w;
      e.v[2 * i + 1] = (uint32_t)(w >> 32);
    }
    e.v[7] &= 0x3fffffffu;
    if (fr_is_canonical(e)) return e;
  }
}

// rand 0.8 `gen_range(0..n)` for usize (widening multiply + zone rejection)
inline uint64_t gen_range(ChaChaRng& rng, uint64_t n) {
  const int lz = __builtin_clzll(n);
  const uint64_t zone = (n << lz) - 1;
  for (;;) {
    const uint64_t v = rng.next_u64();
    const unsigned __int128 pr = (unsigned __int128)v * n;
    if ((uint64_t)pr <= zone) return (uint64_t)(pr >> 64);
  }
}

// get_distinct_indices_from_prng (src/utils.rs:31-55): t ascending indices in [0, n)
inline std::vector<uint64_t> distinct_indices(const uint8_t seed[32], uint64_t n, uint64_t t) {
  ChaChaRng rng(seed);
  const uint64_t to_select = t < n - t ? t : n - t;
  std::set<uint64_t> sel;
  while (sel.size() < to_select) sel.insert(gen_range(rng, n));
  std::vector<uint64_t> out;
  if (to_select == t) {
    out.assign(sel.begin(), sel.end());
  } else {
    for (uint64_t i = 0; i < n; i++)
      if (!sel.count(i)) out.push_back(i);
  }
  return out;
}

}  // namespace lg
