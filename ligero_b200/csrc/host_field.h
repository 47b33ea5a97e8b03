// Host-side field, hash and sponge utilities of the prover driver (product code, no device work):
//   * Fq: BN254 Fr in 4 x 64-bit Montgomery limbs (same bytes as lg::Fr / ark_bn254::Fr) for the
//     host-only parts of LigeroCircuit::prove / verify -- circuit evaluation (src/arithmetic_circuit/
//     mod.rs:247-358), the Fiat-Shamir Poseidon sponge (stays on the host per the design) and the
//     verifier's scalar checks;
//   * SHA-256 for Merkle path verification (ark-crypto-primitives Path::verify);
//   * PoseidonSponge with arkworks' duplex rules.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#endif

namespace lgh {

typedef unsigned __int128 u128;

struct Fq {
  uint64_t l[4];
  bool operator==(const Fq& o) const { return l[0] == o.l[0] && l[1] == o.l[1] && l[2] == o.l[2] && l[3] == o.l[3]; }
  bool operator!=(const Fq& o) const { return !(*this == o); }
  bool operator<(const Fq& o) const {  // arbitrary total order (map keys)
    for (int i = 3; i >= 0; i--)
      if (l[i] != o.l[i]) return l[i] < o.l[i];
    return false;
  }
  bool is_zero() const { return (l[0] | l[1] | l[2] | l[3]) == 0; }
};

static const uint64_t kP[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
static const uint64_t kPInv = 0xc2e1f593efffffffULL;
static const Fq kOne = {{0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL}};
static const Fq kR2 = {{0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL}};
static const Fq kZero = {{0, 0, 0, 0}};

inline bool geq_p(const uint64_t a[4]) {
  for (int i = 3; i >= 0; i--) {
    if (a[i] > kP[i]) return true;
    if (a[i] < kP[i]) return false;
  }
  return true;
}
inline void sub_p(uint64_t a[4]) {
  u128 b = 0;
  for (int i = 0; i < 4; i++) {
    const u128 t = (u128)a[i] - kP[i] - (uint64_t)b;
    a[i] = (uint64_t)t;
    b = (t >> 64) & 1;
  }
}
inline Fq add(const Fq& a, const Fq& b) {
  Fq r;
  u128 c = 0;
  for (int i = 0; i < 4; i++) {
    c += (u128)a.l[i] + b.l[i];
    r.l[i] = (uint64_t)c;
    c >>= 64;
  }
  if (geq_p(r.l)) sub_p(r.l);
  return r;
}
inline Fq sub(const Fq& a, const Fq& b) {
  Fq r;
  u128 br = 0;
  for (int i = 0; i < 4; i++) {
    const u128 t = (u128)a.l[i] - b.l[i] - (uint64_t)br;
    r.l[i] = (uint64_t)t;
    br = (t >> 64) & 1;
  }
  if (br) {
    u128 c = 0;
    for (int i = 0; i < 4; i++) {
      c += (u128)r.l[i] + kP[i];
      r.l[i] = (uint64_t)c;
      c >>= 64;
    }
  }
  return r;
}
inline Fq neg(const Fq& a) { return sub(kZero, a); }
// Montgomery product, "no-carry" CIOS: the modulus' top limb leaves two spare bits, so the running sum fits four limbs
// and the two carry chains (a*b_i and m*r) advance in one pass.  Host side of Fiat-Shamir only (sponge, verifier).
inline Fq mul_portable(const Fq& a, const Fq& b) {
  uint64_t t0 = 0, t1 = 0, t2 = 0, t3 = 0;
#define LGH_MUL_ROUND(bi)                                                \
  {                                                                      \
    u128 A = (u128)a.l[0] * (bi) + t0;                                   \
    const uint64_t m = (uint64_t)A * kPInv;                              \
    u128 C = (u128)m * kP[0] + (uint64_t)A;                              \
    A = (u128)a.l[1] * (bi) + t1 + (uint64_t)(A >> 64);                  \
    C = (u128)m * kP[1] + (uint64_t)A + (uint64_t)(C >> 64);             \
    t0 = (uint64_t)C;                                                    \
    A = (u128)a.l[2] * (bi) + t2 + (uint64_t)(A >> 64);                  \
    C = (u128)m * kP[2] + (uint64_t)A + (uint64_t)(C >> 64);             \
    t1 = (uint64_t)C;                                                    \
    A = (u128)a.l[3] * (bi) + t3 + (uint64_t)(A >> 64);                  \
    C = (u128)m * kP[3] + (uint64_t)A + (uint64_t)(C >> 64);             \
    t2 = (uint64_t)C;                                                    \
    t3 = (uint64_t)(C >> 64) + (uint64_t)(A >> 64);                      \
  }
  LGH_MUL_ROUND(b.l[0]) LGH_MUL_ROUND(b.l[1]) LGH_MUL_ROUND(b.l[2]) LGH_MUL_ROUND(b.l[3])
#undef LGH_MUL_ROUND
  Fq r = {{t0, t1, t2, t3}};
  if (geq_p(r.l)) sub_p(r.l);
  return r;
}
#if defined(__x86_64__) && defined(__GNUC__)
#define LGH_HAVE_MULX 1
// the same CIOS with MULX: the four partial products of a row are issued back to back and their low and high halves
// join the running sum in two add-with-carry chains (20 % faster per product than the compiler's 128-bit code on the
// CPUs of the B200 boxes; used when the CPU has BMI2, checked once at run time)
__attribute__((target("bmi2"))) inline Fq mul_mulx(const Fq& a, const Fq& b) {
  typedef unsigned long long ull;
  ull t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
  const ull a0 = a.l[0], a1 = a.l[1], a2 = a.l[2], a3 = a.l[3];
  const ull q0 = kP[0], q1 = kP[1], q2 = kP[2], q3 = kP[3];
#define LGH_MULX_ROUND(bi)                                                  \
  {                                                                         \
    ull lo0, hi0, lo1, hi1, lo2, hi2, lo3, hi3;                             \
    lo0 = _mulx_u64(a0, (bi), &hi0);                                        \
    lo1 = _mulx_u64(a1, (bi), &hi1);                                        \
    lo2 = _mulx_u64(a2, (bi), &hi2);                                        \
    lo3 = _mulx_u64(a3, (bi), &hi3);                                        \
    unsigned char c = 0, d = 0;                                             \
    c = _addcarry_u64(c, t0, lo0, &t0);                                     \
    c = _addcarry_u64(c, t1, lo1, &t1);                                     \
    c = _addcarry_u64(c, t2, lo2, &t2);                                     \
    c = _addcarry_u64(c, t3, lo3, &t3);                                     \
    c = _addcarry_u64(c, t4, 0, &t4);                                       \
    d = _addcarry_u64(d, t1, hi0, &t1);                                     \
    d = _addcarry_u64(d, t2, hi1, &t2);                                     \
    d = _addcarry_u64(d, t3, hi2, &t3);                                     \
    d = _addcarry_u64(d, t4, hi3, &t4);                                     \
    const ull m = t0 * kPInv;                                               \
    lo0 = _mulx_u64(m, q0, &hi0);                                           \
    lo1 = _mulx_u64(m, q1, &hi1);                                           \
    lo2 = _mulx_u64(m, q2, &hi2);                                           \
    lo3 = _mulx_u64(m, q3, &hi3);                                           \
    c = 0;                                                                  \
    d = 0;                                                                  \
    c = _addcarry_u64(c, t0, lo0, &t0);                                     \
    c = _addcarry_u64(c, t1, lo1, &t1);                                     \
    c = _addcarry_u64(c, t2, lo2, &t2);                                     \
    c = _addcarry_u64(c, t3, lo3, &t3);                                     \
    c = _addcarry_u64(c, t4, 0, &t4);                                       \
    d = _addcarry_u64(d, t1, hi0, &t0);                                     \
    d = _addcarry_u64(d, t2, hi1, &t1);                                     \
    d = _addcarry_u64(d, t3, hi2, &t2);                                     \
    d = _addcarry_u64(d, t4, hi3, &t3);                                     \
    t4 = 0;                                                                 \
  }
  LGH_MULX_ROUND(b.l[0]) LGH_MULX_ROUND(b.l[1]) LGH_MULX_ROUND(b.l[2]) LGH_MULX_ROUND(b.l[3])
#undef LGH_MULX_ROUND
  Fq r = {{t0, t1, t2, t3}};
  if (geq_p(r.l)) sub_p(r.l);
  return r;
}
inline bool cpu_has_mulx() {
  static const bool v = __builtin_cpu_supports("bmi2") != 0;
  return v;
}
// Montgomery product with MULX and the two ADX carry chains (adcx: low halves, adox: high halves): CIOS, the four rounds
// unrolled with the accumulator registers renamed instead of shifted.  LAZY: inputs < 2p, result < 2p (4p^2 < p 2^256), no final
// subtraction.  1.4x the throughput of the compiler's code when independent products are in flight (the three S-boxes of a full
// Poseidon round), the same latency for a dependent chain.  Used by the sponge's width-3 fast path only.
#define LGH_HAVE_ADX 1
__attribute__((target("bmi2,adx"))) static inline void mul_adx_lazy(uint64_t r[4], const uint64_t a[4], const uint64_t b[4]) {
  uint64_t t0, t1, t2, t3, t4, lo, hi, zero;
  static const uint64_t P[4] = {kP[0], kP[1], kP[2], kP[3]};
  const uint64_t pinv = kPInv;
#define LGH_ADX_ROUND(BI, T0, T1, T2, T3, T4)                                                                          \
  "movq " BI "(%[b]), %%rdx\n\t"                                                                                        \
  "xorq %[zero], %[zero]\n\t"                                                                                           \
  "mulx 0(%[a]), %[lo], %[hi]\n\t adcx %[lo], %[" T0 "]\n\t adox %[hi], %[" T1 "]\n\t"                                   \
  "mulx 8(%[a]), %[lo], %[hi]\n\t adcx %[lo], %[" T1 "]\n\t adox %[hi], %[" T2 "]\n\t"                                   \
  "mulx 16(%[a]), %[lo], %[hi]\n\t adcx %[lo], %[" T2 "]\n\t adox %[hi], %[" T3 "]\n\t"                                  \
  "mulx 24(%[a]), %[lo], %[hi]\n\t adcx %[lo], %[" T3 "]\n\t movq %[zero], %[" T4 "]\n\t adox %[hi], %[" T4 "]\n\t"       \
  "adcx %[zero], %[" T4 "]\n\t"                                                                                          \
  "movq %[" T0 "], %%rdx\n\t imulq %[pinv], %%rdx\n\t"                                                                   \
  "xorq %[zero], %[zero]\n\t"                                                                                           \
  "mulx 0(%[p]), %[lo], %[hi]\n\t adcx %[lo], %[" T0 "]\n\t adox %[hi], %[" T1 "]\n\t"                                   \
  "mulx 8(%[p]), %[lo], %[hi]\n\t adcx %[lo], %[" T1 "]\n\t adox %[hi], %[" T2 "]\n\t"                                   \
  "mulx 16(%[p]), %[lo], %[hi]\n\t adcx %[lo], %[" T2 "]\n\t adox %[hi], %[" T3 "]\n\t"                                  \
  "mulx 24(%[p]), %[lo], %[hi]\n\t adcx %[lo], %[" T3 "]\n\t adox %[hi], %[" T4 "]\n\t"                                  \
  "adcx %[zero], %[" T4 "]\n\t"
  asm volatile(
      "xorq %[t0], %[t0]\n\t xorq %[t1], %[t1]\n\t xorq %[t2], %[t2]\n\t xorq %[t3], %[t3]\n\t"
      LGH_ADX_ROUND("0", "t0", "t1", "t2", "t3", "t4")
      LGH_ADX_ROUND("8", "t1", "t2", "t3", "t4", "t0")
      LGH_ADX_ROUND("16", "t2", "t3", "t4", "t0", "t1")
      LGH_ADX_ROUND("24", "t3", "t4", "t0", "t1", "t2")
      : [t0] "=&r"(t0), [t1] "=&r"(t1), [t2] "=&r"(t2), [t3] "=&r"(t3), [t4] "=&r"(t4), [lo] "=&r"(lo), [hi] "=&r"(hi),
        [zero] "=&r"(zero)
      : [a] "r"(a), [b] "r"(b), [p] "r"(P), [pinv] "r"(pinv)
      : "rdx", "cc", "memory");
#undef LGH_ADX_ROUND
  // after four rounds the value sits in (t4, t0, t1, t2); t3 == 0
  r[0] = t4;
  r[1] = t0;
  r[2] = t1;
  r[3] = t2;
}
inline bool cpu_has_adx() {
  static const bool v = __builtin_cpu_supports("bmi2") != 0 && __builtin_cpu_supports("adx") != 0;
  return v;
}
#endif
inline Fq mul(const Fq& a, const Fq& b) {
#ifdef LGH_HAVE_MULX
  if (cpu_has_mulx()) return mul_mulx(a, b);
#endif
  return mul_portable(a, b);
}
inline Fq from_mont(const Fq& a) {
  const Fq one = {{1, 0, 0, 0}};
  return mul(a, one);
}
inline Fq to_mont(const Fq& canonical) { return mul(canonical, kR2); }
inline Fq from_u64(uint64_t x) {
  const Fq c = {{x, 0, 0, 0}};
  return to_mont(c);
}
inline Fq pow_u64(Fq b, uint64_t e) {
  if (e == 17) {  // the test sponge's S-box: 4 squarings + 1 product
    const Fq b2 = mul(b, b), b4 = mul(b2, b2), b8 = mul(b4, b4);
    return mul(mul(b8, b8), b);
  }
  if (e == 5) {
    const Fq b2 = mul(b, b);
    return mul(mul(b2, b2), b);
  }
  Fq acc = kOne;
  bool first = true;
  while (e) {
    if (e & 1) {
      acc = first ? b : mul(acc, b);
      first = false;
    }
    e >>= 1;
    if (e) b = mul(b, b);
  }
  return acc;
}
inline Fq pow_limbs(const Fq& b, const uint64_t e[4]) {
  Fq acc = kOne;
  for (int i = 255; i >= 0; i--) {
    acc = mul(acc, acc);
    if ((e[i >> 6] >> (i & 63)) & 1) acc = mul(acc, b);
  }
  return acc;
}
inline Fq inv(const Fq& a) {
  const uint64_t e[4] = {kP[0] - 2, kP[1], kP[2], kP[3]};
  return pow_limbs(a, e);
}
// generator of the radix-2 domain of size 2^log_n (GeneralEvaluationDomain::element(1))
inline Fq root_of_unity(int log_n) {
  const uint64_t pm1[4] = {kP[0] - 1, kP[1], kP[2], kP[3]};
  uint64_t e[4];
  for (int i = 0; i < 4; i++) e[i] = (pm1[i] >> 28) | (i + 1 < 4 ? pm1[i + 1] << 36 : 0);
  Fq w = pow_limbs(from_u64(5), e);
  for (int i = log_n; i < 28; i++) w = mul(w, w);
  return w;
}
// little-endian bytes mod r -> Montgomery (F::from_le_bytes_mod_order for inputs < 2^256)
inline Fq from_le_bytes_mod_order(const uint8_t* bytes, size_t len) {
  // len <= 31 in every call site (31-byte packing), so the value is already < r
  Fq c = kZero;
  for (size_t i = 0; i < len && i < 32; i++) c.l[i / 8] |= (uint64_t)bytes[i] << (8 * (i % 8));
  if (geq_p(c.l)) sub_p(c.l);
  return to_mont(c);
}

// ---- SHA-256 -------------------------------------------------------------------------------------
inline uint32_t ror32(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
inline void sha256(const uint8_t* msg, size_t len, uint8_t out[32]) {
  static const uint32_t K[64] = {
      0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
      0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
      0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
      0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
      0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
      0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
  uint32_t st[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
  std::vector<uint8_t> buf(msg, msg + len);
  buf.push_back(0x80);
  while (buf.size() % 64 != 56) buf.push_back(0);
  const uint64_t bits = (uint64_t)len * 8;
  for (int i = 7; i >= 0; i--) buf.push_back((uint8_t)(bits >> (8 * i)));
  for (size_t off = 0; off < buf.size(); off += 64) {
    uint32_t w[64];
    for (int i = 0; i < 16; i++)
      w[i] = ((uint32_t)buf[off + 4 * i] << 24) | ((uint32_t)buf[off + 4 * i + 1] << 16) | ((uint32_t)buf[off + 4 * i + 2] << 8) | buf[off + 4 * i + 3];
    for (int i = 16; i < 64; i++) {
      const uint32_t s0 = ror32(w[i - 15], 7) ^ ror32(w[i - 15], 18) ^ (w[i - 15] >> 3);
      const uint32_t s1 = ror32(w[i - 2], 17) ^ ror32(w[i - 2], 19) ^ (w[i - 2] >> 10);
      w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
    for (int i = 0; i < 64; i++) {
      const uint32_t t1 = h + (ror32(e, 6) ^ ror32(e, 11) ^ ror32(e, 25)) + ((e & f) ^ (~e & g)) + K[i] + w[i];
      const uint32_t t2 = (ror32(a, 2) ^ ror32(a, 13) ^ ror32(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
      h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
  }
  for (int i = 0; i < 8; i++) {
    out[4 * i] = (uint8_t)(st[i] >> 24);
    out[4 * i + 1] = (uint8_t)(st[i] >> 16);
    out[4 * i + 2] = (uint8_t)(st[i] >> 8);
    out[4 * i + 3] = (uint8_t)st[i];
  }
}

// ---- Poseidon sponge (ark-crypto-primitives PoseidonSponge; parameters are an input) ---------------
struct PoseidonConfig {
  int full_rounds = 0, partial_rounds = 0;
  uint64_t alpha = 0;
  int rate = 0, capacity = 0;
  std::vector<Fq> mds;  // (rate+capacity)^2 row-major
  std::vector<Fq> ark;  // (full+partial) x (rate+capacity)
};

class PoseidonSponge {
 public:
  explicit PoseidonSponge(const PoseidonConfig& cfg) : cfg_(cfg), state_(cfg.rate + cfg.capacity, kZero) {
    // the permutation runs once per `rate` absorbed elements on the host (Fiat-Shamir is sequential): entries 0 and 1
    // of the MDS matrix (the test sponge's is all 0/1) cost nothing or one addition instead of a product
    for (const Fq& e : cfg_.mds) mds_kind_.push_back(e.is_zero() ? 0 : (e == kOne ? 1 : 2));
  }
  void absorb_field(const std::vector<Fq>& elems) {
    if (elems.empty()) return;
    if (absorbing_) {
      int idx = index_;
      if (idx == cfg_.rate) {
        permute();
        idx = 0;
      }
      absorb_internal(idx, elems);
    } else {
      permute();
      absorb_internal(0, elems);
    }
  }
  // Absorb for Vec<u8>: u64_le(len) || bytes packed 31 bytes per element
  void absorb_bytes(const uint8_t* data, size_t len) {
    std::vector<uint8_t> buf(8 + len);
    const uint64_t l = len;
    memcpy(buf.data(), &l, 8);
    memcpy(buf.data() + 8, data, len);
    std::vector<Fq> elems;
    for (size_t off = 0; off < buf.size(); off += 31) elems.push_back(from_le_bytes_mod_order(buf.data() + off, std::min<size_t>(31, buf.size() - off)));
    absorb_field(elems);
  }
  std::vector<Fq> squeeze_field(size_t n) {
    std::vector<Fq> out;
    int idx;
    if (absorbing_) {
      permute();
      idx = 0;
    } else {
      idx = index_;
      if (idx == cfg_.rate) {
        permute();
        idx = 0;
      }
    }
    for (;;) {
      const size_t need = n - out.size();
      if (idx + need <= (size_t)cfg_.rate) {
        for (size_t i = 0; i < need; i++) out.push_back(state_[cfg_.capacity + idx + i]);
        absorbing_ = false;
        index_ = idx + (int)need;
        return out;
      }
      for (int i = idx; i < cfg_.rate; i++) out.push_back(state_[cfg_.capacity + i]);
      if (out.size() != n) permute();
      idx = 0;
    }
  }
  std::vector<uint8_t> squeeze_bytes(size_t n) {
    const size_t usable = 31;
    const size_t ne = (n + usable - 1) / usable;
    std::vector<uint8_t> out;
    for (const Fq& e : squeeze_field(ne)) {
      const Fq c = from_mont(e);
      const uint8_t* b = (const uint8_t*)c.l;
      out.insert(out.end(), b, b + usable);
    }
    out.resize(n);
    return out;
  }

 private:
  void absorb_internal(int start, const std::vector<Fq>& elems) {
    size_t pos = 0;
    for (;;) {
      const size_t rem = elems.size() - pos;
      if (start + rem <= (size_t)cfg_.rate) {
        for (size_t i = 0; i < rem; i++) state_[cfg_.capacity + start + i] = add(state_[cfg_.capacity + start + i], elems[pos + i]);
        absorbing_ = true;
        index_ = start + (int)rem;
        return;
      }
      const int take = cfg_.rate - start;
      for (int i = 0; i < take; i++) state_[cfg_.capacity + start + i] = add(state_[cfg_.capacity + start + i], elems[pos + i]);
      permute();
      pos += take;
      start = 0;
    }
  }
#ifdef LGH_HAVE_ADX
  // Width 3, alpha = 17, MDS entries all 0 or 1 (the reference's test sponge): the whole permutation on lazily reduced
  // values (< 2p) with the ADX product; the three S-boxes of a full round advance level by level, so three independent
  // products are always in flight (a partial round is one dependent chain of five and stays at the product's latency).
  // 20.5 k sequential permutations per 2^24-gate proof run here; the state is canonical again on exit.
  static inline void csub2p_(uint64_t x[4]) {  // x < 4p -> x < 2p
    static const uint64_t k2P[4] = {0x87c3eb27e0000002ULL, 0x5067d090f372e122ULL, 0x70a08b6d0302b0baULL, 0x60c89ce5c2634053ULL};
    uint64_t t[4];
    u128 b = 0;
    for (int i = 0; i < 4; i++) {
      const u128 d = (u128)x[i] - k2P[i] - (uint64_t)b;
      t[i] = (uint64_t)d;
      b = (d >> 64) & 1;
    }
    const uint64_t keep = (uint64_t)0 - (uint64_t)b;  // all ones when x < 2p
    for (int i = 0; i < 4; i++) x[i] = (x[i] & keep) | (t[i] & ~keep);
  }
  static inline void addl_(uint64_t r[4], const uint64_t a[4], const uint64_t b[4]) {  // a, b < 2p: no overflow, r < 4p
    u128 c = 0;
    for (int i = 0; i < 4; i++) {
      c += (u128)a[i] + b[i];
      r[i] = (uint64_t)c;
      c >>= 64;
    }
  }
  bool fast3_ok() const {
    if (cfg_.rate + cfg_.capacity != 3 || cfg_.alpha != 17 || !cpu_has_adx()) return false;
    for (uint8_t k : mds_kind_)
      if (k > 1) return false;
    static const bool off = getenv("LG_SPONGE_GENERIC") != nullptr;  // tests: force the generic path
    return !off;
  }
  __attribute__((target("bmi2,adx"))) void permute3_adx() {
    uint64_t st[3][4];
    for (int i = 0; i < 3; i++) memcpy(st[i], state_[i].l, 32);
    const int half = cfg_.full_rounds / 2, rounds = cfg_.full_rounds + cfg_.partial_rounds;
    const Fq* ark = cfg_.ark.data();
    for (int rnd = 0; rnd < rounds; rnd++, ark += 3) {
      for (int i = 0; i < 3; i++) {
        addl_(st[i], st[i], ark[i].l);  // < 2p + p
        csub2p_(st[i]);
      }
      if (rnd < half || rnd >= half + cfg_.partial_rounds) {
        uint64_t a2[3][4], a4[3][4], a8[3][4], a16[3][4];
        for (int i = 0; i < 3; i++) mul_adx_lazy(a2[i], st[i], st[i]);
        for (int i = 0; i < 3; i++) mul_adx_lazy(a4[i], a2[i], a2[i]);
        for (int i = 0; i < 3; i++) mul_adx_lazy(a8[i], a4[i], a4[i]);
        for (int i = 0; i < 3; i++) mul_adx_lazy(a16[i], a8[i], a8[i]);
        for (int i = 0; i < 3; i++) mul_adx_lazy(st[i], a16[i], st[i]);
      } else {
        uint64_t a2[4], a4[4], a8[4], a16[4];
        mul_adx_lazy(a2, st[0], st[0]);
        mul_adx_lazy(a4, a2, a2);
        mul_adx_lazy(a8, a4, a4);
        mul_adx_lazy(a16, a8, a8);
        mul_adx_lazy(st[0], a16, st[0]);
      }
      uint64_t nxt[3][4];
      for (int i = 0; i < 3; i++) {  // 0/1 matrix: sums of the selected state words, reduced below 2p after every addition
        bool first = true;
        const uint8_t* kind = mds_kind_.data() + (size_t)i * 3;
        for (int j = 0; j < 3; j++) {
          if (!kind[j]) continue;
          if (first) {
            memcpy(nxt[i], st[j], 32);
            first = false;
          } else {
            addl_(nxt[i], nxt[i], st[j]);
            csub2p_(nxt[i]);
          }
        }
        if (first) memset(nxt[i], 0, 32);
      }
      memcpy(st, nxt, sizeof(st));
    }
    for (int i = 0; i < 3; i++) {
      if (geq_p(st[i])) sub_p(st[i]);  // < 2p -> canonical
      memcpy(state_[i].l, st[i], 32);
    }
  }
#endif
  void permute() {
#ifdef LGH_HAVE_ADX
    if (fast3_ok()) {
      permute3_adx();
      return;
    }
#endif
    const int t = cfg_.rate + cfg_.capacity;
    const int half = cfg_.full_rounds / 2;
    Fq* st = state_.data();
    nxt_.resize(t);
    Fq* nxt = nxt_.data();
    const Fq* ark = cfg_.ark.data();
    for (int rnd = 0; rnd < cfg_.full_rounds + cfg_.partial_rounds; rnd++, ark += t) {
      for (int i = 0; i < t; i++) st[i] = add(st[i], ark[i]);
      if (rnd < half || rnd >= half + cfg_.partial_rounds) {
        for (int i = 0; i < t; i++) st[i] = pow_u64(st[i], cfg_.alpha);
      } else {
        st[0] = pow_u64(st[0], cfg_.alpha);
      }
      for (int i = 0; i < t; i++) {
        Fq acc = kZero;
        bool first = true;
        const Fq* row = cfg_.mds.data() + (size_t)i * t;
        const uint8_t* kind = mds_kind_.data() + (size_t)i * t;
        for (int j = 0; j < t; j++) {
          if (kind[j] == 0) continue;
          const Fq term = kind[j] == 1 ? st[j] : mul(st[j], row[j]);
          acc = first ? term : add(acc, term);
          first = false;
        }
        nxt[i] = acc;
      }
      for (int i = 0; i < t; i++) st[i] = nxt[i];
    }
  }
  PoseidonConfig cfg_;
  std::vector<uint8_t> mds_kind_;  // per MDS entry: 0 = zero, 1 = one, 2 = general
  std::vector<Fq> state_, nxt_;
  bool absorbing_ = true;
  int index_ = 0;
};

}  // namespace lgh
