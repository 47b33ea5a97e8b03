/* CPU oracle in C (TEST INFRASTRUCTURE ONLY -- never linked into or called by the product).
 *
 * A fast restatement of the reference's encode + commit + test arithmetic with the REFERENCE'S OWN
 * SCHEDULE (per-row iFFT_k, zero-pad, FFT_n; full transpose; one BLAKE2s per column; SHA-256 tree),
 * used (a) as the checker at sizes the Python oracle cannot reach and (b) as the CPU baseline that
 * bench.py times on the GPU box's host cores ("kind": "port").  It is validated against
 * oracle/ligero_oracle.py (tests/test_oracle_c.py), which is itself pinned only by the reference's
 * in-tree known answers -- PARITY UNPINNED, see DESIGN.md.
 *
 * Follows: src/ligero/mod.rs:521-551 (encode, column hash, Merkle), 998-1008 (RS helpers),
 * src/matrices/mod.rs:100-110,138-149,163-171, src/utils.rs:23-55; third-party algorithms restated
 * from their specifications: arkworks radix-2 FFT (ark-poly 0.5), BLAKE2s (RFC 7693), SHA-256
 * (FIPS 180-4), ChaCha20 (rand_chacha 0.3), Montgomery arithmetic over BN254 Fr (ark-ff 0.5).
 *
 * Build: make -C oracle   (gcc -O3 -shared -fPIC -pthread)
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t l[4]; } fr_t; /* Montgomery form, little-endian limbs (ark_bn254::Fr layout) */

static const uint64_t P[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
static const uint64_t PINV = 0xc2e1f593efffffffULL; /* -p^-1 mod 2^64 */
static const fr_t ONE = {{0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL}};
static const fr_t R2 = {{0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL}};

static inline int geq_p(const uint64_t a[4]) {
  for (int i = 3; i >= 0; i--) {
    if (a[i] > P[i]) return 1;
    if (a[i] < P[i]) return 0;
  }
  return 1;
}
static inline void sub_p(uint64_t a[4]) {
  u128 b = 0;
  for (int i = 0; i < 4; i++) {
    u128 t = (u128)a[i] - P[i] - (uint64_t)b;
    a[i] = (uint64_t)t;
    b = (t >> 64) & 1;
  }
}
static inline fr_t fr_add(fr_t a, fr_t b) {
  fr_t r;
  u128 c = 0;
  for (int i = 0; i < 4; i++) {
    c += (u128)a.l[i] + b.l[i];
    r.l[i] = (uint64_t)c;
    c >>= 64;
  }
  if (geq_p(r.l)) sub_p(r.l);
  return r;
}
static inline fr_t fr_sub(fr_t a, fr_t b) {
  fr_t r;
  u128 br = 0;
  for (int i = 0; i < 4; i++) {
    u128 t = (u128)a.l[i] - b.l[i] - (uint64_t)br;
    r.l[i] = (uint64_t)t;
    br = (t >> 64) & 1;
  }
  if (br) {
    u128 c = 0;
    for (int i = 0; i < 4; i++) {
      c += (u128)r.l[i] + P[i];
      r.l[i] = (uint64_t)c;
      c >>= 64;
    }
  }
  return r;
}
static inline fr_t fr_mul(fr_t a, fr_t b) { /* CIOS, 4 x 64-bit limbs */
  uint64_t t[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) {
      c += (u128)a.l[j] * b.l[i] + t[j];
      t[j] = (uint64_t)c;
      c >>= 64;
    }
    c += t[4];
    t[4] = (uint64_t)c;
    t[5] = (uint64_t)(c >> 64);
    uint64_t m = t[0] * PINV;
    c = (u128)m * P[0] + t[0];
    c >>= 64;
    for (int j = 1; j < 4; j++) {
      c += (u128)m * P[j] + t[j];
      t[j - 1] = (uint64_t)c;
      c >>= 64;
    }
    c += t[4];
    t[3] = (uint64_t)c;
    t[4] = t[5] + (uint64_t)(c >> 64);
  }
  fr_t r = {{t[0], t[1], t[2], t[3]}};
  if (t[4] || geq_p(r.l)) sub_p(r.l);
  return r;
}
static inline fr_t fr_from_mont(fr_t a) {
  fr_t one = {{1, 0, 0, 0}};
  return fr_mul(a, one);
}
static inline fr_t fr_to_mont(fr_t a) { return fr_mul(a, R2); }
static inline int fr_is_zero(fr_t a) { return (a.l[0] | a.l[1] | a.l[2] | a.l[3]) == 0; }
static fr_t fr_pow(fr_t b, const uint64_t e[4]) {
  fr_t acc = ONE;
  for (int i = 255; i >= 0; i--) {
    acc = fr_mul(acc, acc);
    if ((e[i >> 6] >> (i & 63)) & 1) acc = fr_mul(acc, b);
  }
  return acc;
}
static fr_t fr_inv(fr_t a) {
  uint64_t e[4] = {P[0] - 2, P[1], P[2], P[3]};
  return fr_pow(a, e);
}
static fr_t fr_from_u64(uint64_t x) {
  fr_t c = {{x, 0, 0, 0}};
  return fr_to_mont(c);
}
static fr_t root_of_unity(int log_n) { /* 5^((p-1)/2^28) squared down to order 2^log_n */
  uint64_t e[4];
  uint64_t pm1[4] = {P[0] - 1, P[1], P[2], P[3]};
  for (int i = 0; i < 4; i++) e[i] = (pm1[i] >> 28) | (i + 1 < 4 ? pm1[i + 1] << 36 : 0);
  fr_t w = fr_pow(fr_from_u64(5), e);
  for (int i = log_n; i < 28; i++) w = fr_mul(w, w);
  return w;
}

/* ---------------------------------------------------------------------------------------------
 * radix-2 FFT as ark-poly's Radix2EvaluationDomain computes it: natural order in, natural out.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int log_n;
  size_t n;
  fr_t* tw;     /* w^i, i < n/2 */
  fr_t* tw_inv; /* w^-i */
  fr_t n_inv;
} domain_t;

static void domain_init(domain_t* d, int log_n) {
  d->log_n = log_n;
  d->n = (size_t)1 << log_n;
  size_t h = d->n / 2 ? d->n / 2 : 1;
  d->tw = (fr_t*)malloc(h * sizeof(fr_t));
  d->tw_inv = (fr_t*)malloc(h * sizeof(fr_t));
  fr_t w = root_of_unity(log_n), wi = fr_inv(w);
  d->tw[0] = d->tw_inv[0] = ONE;
  for (size_t i = 1; i < h; i++) {
    d->tw[i] = fr_mul(d->tw[i - 1], w);
    d->tw_inv[i] = fr_mul(d->tw_inv[i - 1], wi);
  }
  d->n_inv = fr_inv(fr_from_u64(d->n));
}
static void domain_free(domain_t* d) {
  free(d->tw);
  free(d->tw_inv);
}
static void bitrev_permute(fr_t* a, int log_n) {
  size_t n = (size_t)1 << log_n;
  for (size_t i = 0; i < n; i++) {
    size_t j = 0;
    for (int b = 0; b < log_n; b++) j |= ((i >> b) & 1) << (log_n - 1 - b);
    if (i < j) {
      fr_t t = a[i];
      a[i] = a[j];
      a[j] = t;
    }
  }
}
static void fft_core(fr_t* a, int log_n, const fr_t* tw) {
  size_t n = (size_t)1 << log_n;
  bitrev_permute(a, log_n);
  for (size_t len = 2; len <= n; len <<= 1) {
    size_t half = len >> 1, step = n / len;
    for (size_t s = 0; s < n; s += len)
      for (size_t j = 0; j < half; j++) {
        fr_t u = a[s + j], v = fr_mul(a[s + j + half], tw[j * step]);
        a[s + j] = fr_add(u, v);
        a[s + j + half] = fr_sub(u, v);
      }
  }
}
static void domain_fft(const domain_t* d, fr_t* a) { fft_core(a, d->log_n, d->tw); }
static void domain_ifft(const domain_t* d, fr_t* a) {
  fft_core(a, d->log_n, d->tw_inv);
  for (size_t i = 0; i < d->n; i++) a[i] = fr_mul(a[i], d->n_inv);
}

/* ---------------------------------------------------------------------------------------------
 * BLAKE2s-256 and SHA-256
 * ------------------------------------------------------------------------------------------- */
static const uint32_t B2S_IV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au, 0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
static const uint8_t B2S_SIGMA[10][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0}};
static inline uint32_t ror32(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
typedef struct {
  uint32_t h[8];
  uint64_t t;
  uint8_t buf[64];
  size_t buflen;
} b2s_t;
static void b2s_compress(b2s_t* S, const uint8_t block[64], int last) {
  uint32_t m[16], v[16];
  for (int i = 0; i < 16; i++) memcpy(&m[i], block + 4 * i, 4);
  for (int i = 0; i < 8; i++) { v[i] = S->h[i]; v[i + 8] = B2S_IV[i]; }
  v[12] ^= (uint32_t)S->t;
  v[13] ^= (uint32_t)(S->t >> 32);
  if (last) v[14] = ~v[14];
#define G(a, b, c, d, x, y)                  \
  v[a] = v[a] + v[b] + (x); v[d] = ror32(v[d] ^ v[a], 16); v[c] = v[c] + v[d]; v[b] = ror32(v[b] ^ v[c], 12); \
  v[a] = v[a] + v[b] + (y); v[d] = ror32(v[d] ^ v[a], 8);  v[c] = v[c] + v[d]; v[b] = ror32(v[b] ^ v[c], 7);
  for (int r = 0; r < 10; r++) {
    const uint8_t* s = B2S_SIGMA[r];
    G(0, 4, 8, 12, m[s[0]], m[s[1]]) G(1, 5, 9, 13, m[s[2]], m[s[3]]) G(2, 6, 10, 14, m[s[4]], m[s[5]]) G(3, 7, 11, 15, m[s[6]], m[s[7]])
    G(0, 5, 10, 15, m[s[8]], m[s[9]]) G(1, 6, 11, 12, m[s[10]], m[s[11]]) G(2, 7, 8, 13, m[s[12]], m[s[13]]) G(3, 4, 9, 14, m[s[14]], m[s[15]])
  }
#undef G
  for (int i = 0; i < 8; i++) S->h[i] ^= v[i] ^ v[i + 8];
}
static void b2s_init(b2s_t* S) {
  memcpy(S->h, B2S_IV, 32);
  S->h[0] ^= 0x01010020u;
  S->t = 0;
  S->buflen = 0;
}
static void b2s_update(b2s_t* S, const uint8_t* in, size_t len) {
  while (len) {
    if (S->buflen == 64) { /* buffer full and more input follows: not the last block */
      S->t += 64;
      b2s_compress(S, S->buf, 0);
      S->buflen = 0;
    }
    size_t take = 64 - S->buflen;
    if (take > len) take = len;
    memcpy(S->buf + S->buflen, in, take);
    S->buflen += take;
    in += take;
    len -= take;
  }
}
static void b2s_final(b2s_t* S, uint8_t out[32]) {
  S->t += S->buflen;
  memset(S->buf + S->buflen, 0, 64 - S->buflen);
  b2s_compress(S, S->buf, 1);
  memcpy(out, S->h, 32);
}

static const uint32_t SHA_K[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
    0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
    0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
    0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
    0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
    0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
static void sha256_block(uint32_t st[8], const uint8_t blk[64]) {
  uint32_t w[64];
  for (int i = 0; i < 16; i++) w[i] = ((uint32_t)blk[4 * i] << 24) | ((uint32_t)blk[4 * i + 1] << 16) | ((uint32_t)blk[4 * i + 2] << 8) | blk[4 * i + 3];
  for (int i = 16; i < 64; i++) {
    uint32_t s0 = ror32(w[i - 15], 7) ^ ror32(w[i - 15], 18) ^ (w[i - 15] >> 3);
    uint32_t s1 = ror32(w[i - 2], 17) ^ ror32(w[i - 2], 19) ^ (w[i - 2] >> 10);
    w[i] = w[i - 16] + s0 + w[i - 7] + s1;
  }
  uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
  for (int i = 0; i < 64; i++) {
    uint32_t t1 = h + (ror32(e, 6) ^ ror32(e, 11) ^ ror32(e, 25)) + ((e & f) ^ (~e & g)) + SHA_K[i] + w[i];
    uint32_t t2 = (ror32(a, 2) ^ ror32(a, 13) ^ ror32(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
    h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
  }
  st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}
void ref_sha256(const uint8_t* msg, size_t len, uint8_t out[32]) {
  uint32_t st[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
  size_t off = 0;
  for (; off + 64 <= len; off += 64) sha256_block(st, msg + off);
  uint8_t tail[128];
  size_t rem = len - off;
  memset(tail, 0, sizeof tail);
  memcpy(tail, msg + off, rem);
  tail[rem] = 0x80;
  size_t tl = rem + 9 <= 64 ? 64 : 128;
  uint64_t bits = (uint64_t)len * 8;
  for (int i = 0; i < 8; i++) tail[tl - 1 - i] = (uint8_t)(bits >> (8 * i));
  sha256_block(st, tail);
  if (tl == 128) sha256_block(st, tail + 64);
  for (int i = 0; i < 8; i++) { out[4 * i] = st[i] >> 24; out[4 * i + 1] = st[i] >> 16; out[4 * i + 2] = st[i] >> 8; out[4 * i + 3] = st[i]; }
}
void ref_blake2s(const uint8_t* msg, size_t len, uint8_t out[32]) {
  b2s_t S;
  b2s_init(&S);
  b2s_update(&S, msg, len);
  b2s_final(&S, out);
}

/* ---------------------------------------------------------------------------------------------
 * thread pool helper: run fn(lo, hi, arg) over [0, n) split across `threads`
 * ------------------------------------------------------------------------------------------- */
typedef void (*range_fn)(size_t lo, size_t hi, void* arg);
typedef struct { range_fn fn; size_t lo, hi; void* arg; } job_t;
static void* job_main(void* p) {
  job_t* j = (job_t*)p;
  j->fn(j->lo, j->hi, j->arg);
  return NULL;
}
static void parallel_for(size_t n, int threads, range_fn fn, void* arg) {
  if (threads < 1) threads = 1;
  if ((size_t)threads > n) threads = (int)(n ? n : 1);
  if (threads == 1) {
    fn(0, n, arg);
    return;
  }
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * threads);
  job_t* jobs = (job_t*)malloc(sizeof(job_t) * threads);
  for (int t = 0; t < threads; t++) {
    jobs[t].fn = fn;
    jobs[t].lo = n * t / threads;
    jobs[t].hi = n * (t + 1) / threads;
    jobs[t].arg = arg;
    pthread_create(&th[t], NULL, job_main, &jobs[t]);
  }
  for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
  free(th);
  free(jobs);
}
static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* ---------------------------------------------------------------------------------------------
 * encode (src/ligero/mod.rs:521-533) and commit (536-551)
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const fr_t* msg;
  fr_t* u;
  size_t k, n;
  const domain_t *dk, *dn;
} enc_arg;
static void encode_range(size_t lo, size_t hi, void* p) {
  enc_arg* a = (enc_arg*)p;
  fr_t* coeffs = (fr_t*)malloc(a->k * sizeof(fr_t));
  for (size_t i = lo; i < hi; i++) {
    memcpy(coeffs, a->msg + i * a->k, a->k * sizeof(fr_t));
    domain_ifft(a->dk, coeffs);                         /* reed_solomon_interpolate (998-1002) */
    fr_t* row = a->u + i * a->n;
    memcpy(row, coeffs, a->k * sizeof(fr_t));           /* resize(n, 0) */
    memset(row + a->k, 0, (a->n - a->k) * sizeof(fr_t));
    domain_fft(a->dn, row);                             /* reed_solomon_evaluate (1004-1008) */
  }
  free(coeffs);
}
typedef struct {
  const fr_t* u;
  size_t rows, n;
  uint8_t* leaves;
  int prefix;
} hash_arg;
static void hash_range(size_t lo, size_t hi, void* p) {
  hash_arg* a = (hash_arg*)p;
  fr_t* col = (fr_t*)malloc((a->rows ? a->rows : 1) * sizeof(fr_t));
  for (size_t j = lo; j < hi; j++) {
    for (size_t i = 0; i < a->rows; i++) col[i] = fr_from_mont(a->u[i * a->n + j]); /* columns() + serialize */
    b2s_t S;
    b2s_init(&S);
    if (a->prefix) {
      uint64_t len = a->rows;
      b2s_update(&S, (const uint8_t*)&len, 8);
    }
    b2s_update(&S, (const uint8_t*)col, a->rows * 32);
    b2s_final(&S, a->leaves + 32 * j);
  }
  free(col);
}
static void merkle(const uint8_t* leaves, size_t n, uint8_t* nodes, int leaf_prefix) {
  size_t half = n / 2;
  for (size_t i = 0; i < half; i++) {
    uint8_t buf[80];
    size_t len = 0;
    for (int s = 0; s < 2; s++) {
      if (leaf_prefix) {
        uint64_t l = 32;
        memcpy(buf + len, &l, 8);
        len += 8;
      }
      memcpy(buf + len, leaves + 32 * (2 * i + s), 32);
      len += 32;
    }
    ref_sha256(buf, len, nodes + 32 * (half - 1 + i));
  }
  for (size_t i = half - 1; i-- > 0;) {
    uint8_t buf[64];
    memcpy(buf, nodes + 32 * (2 * i + 1), 32);
    memcpy(buf + 32, nodes + 32 * (2 * i + 2), 32);
    ref_sha256(buf, 64, nodes + 32 * i);
  }
}

/* returns 0 on success.  u_out (rows*n), leaves_out (n*32), nodes_out ((n-1)*32) may be NULL.
 * secs[0] = encode seconds, secs[1] = hash+tree seconds. */
int ref_commit(const uint64_t* preenc_u, size_t rows, size_t k, unsigned rho_inv, int threads, int col_prefix, int leaf_prefix,
               uint64_t* u_out, uint8_t* leaves_out, uint8_t* nodes_out, uint8_t root[32], double secs[2]) {
  int log_k = 0, log_rho = 0;
  while (((size_t)1 << log_k) < k) log_k++;
  while ((1u << log_rho) < rho_inv) log_rho++;
  if (((size_t)1 << log_k) != k || (1u << log_rho) != rho_inv || rows == 0) return 1;
  size_t n = k * rho_inv;
  domain_t dk, dn;
  domain_init(&dk, log_k);
  domain_init(&dn, log_k + log_rho);
  fr_t* u = u_out ? (fr_t*)u_out : (fr_t*)malloc(rows * n * sizeof(fr_t));
  uint8_t* leaves = leaves_out ? leaves_out : (uint8_t*)malloc(n * 32);
  uint8_t* nodes = nodes_out ? nodes_out : (uint8_t*)malloc(n * 32);
  if (!u || !leaves || !nodes) return 2;
  double t0 = now_s();
  enc_arg ea = {(const fr_t*)preenc_u, u, k, n, &dk, &dn};
  parallel_for(rows, threads, encode_range, &ea);
  double t1 = now_s();
  hash_arg ha = {u, rows, n, leaves, col_prefix};
  parallel_for(n, threads, hash_range, &ha);
  merkle(leaves, n, nodes, leaf_prefix);
  double t2 = now_s();
  if (root) memcpy(root, nodes, 32);
  if (secs) { secs[0] = t1 - t0; secs[1] = t2 - t1; }
  if (!u_out) free(u);
  if (!leaves_out) free(leaves);
  if (!nodes_out) free(nodes);
  domain_free(&dk);
  domain_free(&dn);
  return 0;
}

/* plain batched transforms for cross-checks: dir 0 = fft, 1 = ifft; rows of length 2^log_n in place */
int ref_fft_rows(uint64_t* data, size_t rows, int log_n, int inverse) {
  domain_t d;
  domain_init(&d, log_n);
  for (size_t i = 0; i < rows; i++) {
    if (inverse) domain_ifft(&d, (fr_t*)data + i * d.n);
    else domain_fft(&d, (fr_t*)data + i * d.n);
  }
  domain_free(&d);
  return 0;
}

/* element-wise helpers exposed for tests */
void ref_fr_mul(const uint64_t* a, const uint64_t* b, uint64_t* out, size_t count) {
  for (size_t i = 0; i < count; i++) ((fr_t*)out)[i] = fr_mul(((const fr_t*)a)[i], ((const fr_t*)b)[i]);
}

/* ---------------------------------------------------------------------------------------------
 * ChaCha20 challenge expansion (src/utils.rs:23-55)
 * ------------------------------------------------------------------------------------------- */
static inline uint32_t rol32(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }
static void chacha_block(const uint32_t key[8], uint64_t counter, int rounds, uint32_t out[16]) {
  uint32_t st[16] = {0x61707865, 0x3320646e, 0x79622d32, 0x6b206574, key[0], key[1], key[2], key[3], key[4], key[5], key[6], key[7],
                     (uint32_t)counter, (uint32_t)(counter >> 32), 0, 0};
  uint32_t x[16];
  memcpy(x, st, sizeof x);
#define QR(a, b, c, d) \
  x[a] += x[b]; x[d] = rol32(x[d] ^ x[a], 16); x[c] += x[d]; x[b] = rol32(x[b] ^ x[c], 12); \
  x[a] += x[b]; x[d] = rol32(x[d] ^ x[a], 8);  x[c] += x[d]; x[b] = rol32(x[b] ^ x[c], 7);
  for (int r = 0; r < rounds; r += 2) {
    QR(0, 4, 8, 12) QR(1, 5, 9, 13) QR(2, 6, 10, 14) QR(3, 7, 11, 15)
    QR(0, 5, 10, 15) QR(1, 6, 11, 12) QR(2, 7, 8, 13) QR(3, 4, 9, 14)
  }
#undef QR
  for (int i = 0; i < 16; i++) out[i] = x[i] + st[i];
}
typedef struct { uint32_t key[8]; uint64_t ctr; uint32_t buf[16]; int pos; } rng_t;
static void rng_init(rng_t* r, const uint8_t seed[32]) { memcpy(r->key, seed, 32); r->ctr = 0; r->pos = 16; }
static uint32_t rng_u32(rng_t* r) {
  if (r->pos == 16) { chacha_block(r->key, r->ctr++, 20, r->buf); r->pos = 0; }
  return r->buf[r->pos++];
}
static uint64_t rng_u64(rng_t* r) { uint64_t lo = rng_u32(r); uint64_t hi = rng_u32(r); return lo | (hi << 32); }

/* get_field_elements_from_prng: out = count Fr (the sampled integer IS the Montgomery representation) */
void ref_expand_fr(const uint8_t seed[32], size_t count, uint64_t* out) {
  rng_t r;
  rng_init(&r, seed);
  for (size_t i = 0; i < count;) {
    uint64_t l[4];
    for (int j = 0; j < 4; j++) l[j] = rng_u64(&r);
    l[3] &= 0xFFFFFFFFFFFFFFFFULL >> 2;
    if (!geq_p(l)) { memcpy(out + 4 * i, l, 32); i++; }
  }
}
static int cmp_u64(const void* a, const void* b) { uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b; return x < y ? -1 : x > y; }
/* get_distinct_indices_from_prng: writes t ascending indices */
void ref_expand_indices(const uint8_t seed[32], size_t n, size_t t, uint64_t* out) {
  rng_t r;
  rng_init(&r, seed);
  size_t to_select = t < n - t ? t : n - t;
  uint8_t* mark = (uint8_t*)calloc(n, 1);
  size_t have = 0;
  int lz = __builtin_clzll((unsigned long long)n);
  uint64_t zone = ((uint64_t)n << lz) - 1;
  while (have < to_select) {
    uint64_t v = rng_u64(&r);
    u128 pr = (u128)v * n;
    if ((uint64_t)pr <= zone) {
      size_t idx = (size_t)(pr >> 64);
      if (!mark[idx]) { mark[idx] = 1; have++; }
    }
  }
  size_t o = 0;
  for (size_t i = 0; i < n; i++)
    if ((mark[i] != 0) == (to_select == t)) out[o++] = i;
  (void)cmp_u64;
  free(mark);
}

/* ---------------------------------------------------------------------------------------------
 * tests: row combinations and polynomials (reference schedules)
 * ------------------------------------------------------------------------------------------- */
/* DenseMatrix::row_mul (src/matrices/mod.rs:138-149): out[c] = sum_i r[i] * M[i][c] */
typedef struct { const fr_t* m; const fr_t* r; size_t rows, cols; fr_t* out; } rm_arg;
static void row_mul_range(size_t lo, size_t hi, void* p) {
  rm_arg* a = (rm_arg*)p;
  for (size_t c = lo; c < hi; c++) {
    fr_t acc = {{0, 0, 0, 0}};
    for (size_t i = 0; i < a->rows; i++) acc = fr_add(acc, fr_mul(a->m[i * a->cols + c], a->r[i]));
    a->out[c] = acc;
  }
}
void ref_row_mul(const uint64_t* m, const uint64_t* r, size_t rows, size_t cols, int threads, uint64_t* out) {
  rm_arg a = {(const fr_t*)m, (const fr_t*)r, rows, cols, (fr_t*)out};
  parallel_for(cols, threads, row_mul_range, &a);
}

/* SparseMatrix::row_mul (src/matrices/mod.rs:100-110) on CSR: out[col] += r[row] * val */
void ref_sparse_row_mul(const uint64_t* row_ptr, const uint64_t* col_idx, const uint64_t* vals, const uint64_t* r, size_t rows,
                        size_t cols, uint64_t* out) {
  fr_t* o = (fr_t*)out;
  memset(o, 0, cols * sizeof(fr_t));
  for (size_t i = 0; i < rows; i++)
    for (uint64_t e = row_ptr[i]; e < row_ptr[i + 1]; e++)
      o[col_idx[e]] = fr_add(o[col_idx[e]], fr_mul(((const fr_t*)r)[i], ((const fr_t*)vals)[e]));
}

/* sum_i a_i(x) * b_i(x) for `count` pairs of polynomials given by their k evaluations on the small
 * domain (a: rows of U_pre, b: rows of r_a); the reference interpolates both (ifft_k), multiplies the
 * coefficient polynomials (FFT-based, size 2k) and sums (src/ligero/mod.rs:723-736).  out: 2k coeffs. */
typedef struct { const fr_t *a, *b; size_t k; const domain_t *dk, *d2; fr_t* partial; } lin_arg;
static void lin_range(size_t lo, size_t hi, void* p) {
  lin_arg* a = (lin_arg*)p;
  size_t k = a->k, k2 = 2 * k;
  fr_t* x = (fr_t*)malloc(k2 * sizeof(fr_t));
  fr_t* y = (fr_t*)malloc(k2 * sizeof(fr_t));
  fr_t* acc = (fr_t*)calloc(k2, sizeof(fr_t));
  for (size_t i = lo; i < hi; i++) {
    memcpy(x, a->a + i * k, k * sizeof(fr_t));
    memcpy(y, a->b + i * k, k * sizeof(fr_t));
    domain_ifft(a->dk, x);
    domain_ifft(a->dk, y);
    memset(x + k, 0, k * sizeof(fr_t));
    memset(y + k, 0, k * sizeof(fr_t));
    domain_fft(a->d2, x);
    domain_fft(a->d2, y);
    for (size_t j = 0; j < k2; j++) acc[j] = fr_add(acc[j], fr_mul(x[j], y[j]));
  }
  memcpy(a->partial, acc, k2 * sizeof(fr_t)); /* each worker owns its slot */
  free(x); free(y); free(acc);
}
void ref_linear_poly(const uint64_t* u_pre, const uint64_t* r_a, size_t count, size_t k, int threads, uint64_t* out) {
  int log_k = 0;
  while (((size_t)1 << log_k) < k) log_k++;
  domain_t dk, d2;
  domain_init(&dk, log_k);
  domain_init(&d2, log_k + 1);
  if (threads < 1) threads = 1;
  if ((size_t)threads > count) threads = (int)count;
  size_t k2 = 2 * k;
  fr_t* partial = (fr_t*)calloc((size_t)(threads + 1) * k2, sizeof(fr_t));
  /* run ranges explicitly so each worker owns slot t */
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * threads);
  job_t* jobs = (job_t*)malloc(sizeof(job_t) * threads);
  lin_arg* args = (lin_arg*)malloc(sizeof(lin_arg) * threads);
  for (int t = 0; t < threads; t++) {
    args[t] = (lin_arg){(const fr_t*)u_pre, (const fr_t*)r_a, k, &dk, &d2, partial + (size_t)t * k2};
    jobs[t] = (job_t){lin_range, count * t / threads, count * (t + 1) / threads, &args[t]};
    pthread_create(&th[t], NULL, job_main, &jobs[t]);
  }
  for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
  fr_t* acc = (fr_t*)calloc(k2, sizeof(fr_t));
  for (int t = 0; t < threads; t++)
    for (size_t j = 0; j < k2; j++) acc[j] = fr_add(acc[j], partial[(size_t)t * k2 + j]);
  domain_ifft(&d2, acc); /* evaluations of the sum on the 2k domain -> coefficients (exact: deg < 2k-1) */
  memcpy(out, acc, k2 * sizeof(fr_t));
  free(acc); free(partial); free(th); free(jobs); free(args);
  domain_free(&dk);
  domain_free(&d2);
}

/* sum_i r_i (x_i(v) y_i(v) - z_i(v)), i < m, rows given as evaluations on the small domain
 * (src/ligero/mod.rs:842-848).  out: 2k coefficients. */
void ref_quadratic_poly(const uint64_t* u_pre, const uint64_t* r, size_t m, size_t k, uint64_t* out) {
  int log_k = 0;
  while (((size_t)1 << log_k) < k) log_k++;
  domain_t dk, d2;
  domain_init(&dk, log_k);
  domain_init(&d2, log_k + 1);
  size_t k2 = 2 * k;
  fr_t* x = (fr_t*)malloc(k2 * sizeof(fr_t));
  fr_t* y = (fr_t*)malloc(k2 * sizeof(fr_t));
  fr_t* z = (fr_t*)malloc(k2 * sizeof(fr_t));
  fr_t* acc = (fr_t*)calloc(k2, sizeof(fr_t));
  const fr_t* U = (const fr_t*)u_pre;
  for (size_t i = 0; i < m; i++) {
    const fr_t* src[3] = {U + i * k, U + (m + i) * k, U + (2 * m + i) * k};
    fr_t* dst[3] = {x, y, z};
    for (int s = 0; s < 3; s++) {
      memcpy(dst[s], src[s], k * sizeof(fr_t));
      domain_ifft(&dk, dst[s]);
      memset(dst[s] + k, 0, k * sizeof(fr_t));
      domain_fft(&d2, dst[s]);
    }
    fr_t ri = ((const fr_t*)r)[i];
    for (size_t j = 0; j < k2; j++) acc[j] = fr_add(acc[j], fr_mul(ri, fr_sub(fr_mul(x[j], y[j]), z[j])));
  }
  domain_ifft(&d2, acc);
  memcpy(out, acc, k2 * sizeof(fr_t));
  free(x); free(y); free(z); free(acc);
  domain_free(&dk);
  domain_free(&d2);
}
