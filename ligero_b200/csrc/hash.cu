// Column hashing and the Merkle tree for sm_100a.
//
//  * leaf_j = BLAKE2s-256( u64_le(R) || canonical-LE-32B(U[0][j]) || ... || U[R-1][j] )
//      = FieldToBytesColHasher<Fr, Blake2s256> of the reference's LigeroMTTestParams
//        (src/ligero/types.rs:15-46, call site src/ligero/mod.rs:536-542; transpose of
//        src/matrices/mod.rs:163-167 is never materialised: a thread walks its column down the rows,
//        a warp reads 32 consecutive 32-byte elements of one row = 1 KiB fully coalesced).
//  * inner nodes = SHA-256 two-to-one (ark-crypto-primitives MerkleTree over TestMerkleTreeParams;
//        call site src/ligero/mod.rs:544-551): bottom level hashes u64_le(32)||L||u64_le(32)||R,
//        upper levels hash L||R; heap layout, node 0 = root.
// Both length-prefix conventions are runtime switches (SURVEY App. A.4/A.5 are recollections).
#include "lg_internal.h"

namespace lg {

__device__ __forceinline__ Fr ld_fr_g(const Fr* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  const uint4 a = q[0], b = q[1];
  Fr r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}

// ------------------------------------------------------------------------------------------------
// BLAKE2s (RFC 7693)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t rotr32(uint32_t x, int n) { return __funnelshift_r(x, x, n); }

// The xor + rotate pairs (LOP3, SHF) can only run on the half-rate ALU pipe, which bounds the kernel; the
// additions are therefore written as multiply-adds by 1 so that they issue on the FMA pipe (IMAD) instead of
// taking ALU slots as IADD3 (ptxas otherwise splits them about evenly between the two pipes).
// The multiplier is read from memory at run time: a literal 1 is folded back into IADD3 by ptxas.
__device__ uint32_t g_b2s_one = 1;
#define add_fma(a, b) ((a) * b2s_one + (b))

#define B2S_G(a, b, c, d, x, y)                                \
  do {                                                         \
    a = FMA ? add_fma(add_fma(a, b), (x)) : a + b + (x);       \
    d = rotr32(d ^ a, 16);                                     \
    c = FMA ? add_fma(c, d) : c + d;                           \
    b = rotr32(b ^ c, 12);                                     \
    a = FMA ? add_fma(add_fma(a, b), (y)) : a + b + (y);       \
    d = rotr32(d ^ a, 8);                                      \
    c = FMA ? add_fma(c, d) : c + d;                           \
    b = rotr32(b ^ c, 7);                                      \
  } while (0)

#define B2S_ROUND(s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, s12, s13, s14, s15) \
  do {                                                                                   \
    B2S_G(v0, v4, v8, v12, m[s0], m[s1]);                                                \
    B2S_G(v1, v5, v9, v13, m[s2], m[s3]);                                                \
    B2S_G(v2, v6, v10, v14, m[s4], m[s5]);                                               \
    B2S_G(v3, v7, v11, v15, m[s6], m[s7]);                                               \
    B2S_G(v0, v5, v10, v15, m[s8], m[s9]);                                               \
    B2S_G(v1, v6, v11, v12, m[s10], m[s11]);                                             \
    B2S_G(v2, v7, v8, v13, m[s12], m[s13]);                                              \
    B2S_G(v3, v4, v9, v14, m[s14], m[s15]);                                              \
  } while (0)

// FMA: additions as multiply-adds on the FMA pipe (enough warps to fill the ALU pipe); otherwise plain additions
// (three-input IADD3: fewer instructions and a shorter dependent chain, for the few-warps regime)
template <bool FMA>
__device__ __forceinline__ void blake2s_compress(uint32_t (&h)[8], const uint32_t (&m)[16], uint64_t t, bool last,
                                                 const uint32_t b2s_one) {
  uint32_t v0 = h[0], v1 = h[1], v2 = h[2], v3 = h[3], v4 = h[4], v5 = h[5], v6 = h[6], v7 = h[7];
  uint32_t v8 = 0x6A09E667u, v9 = 0xBB67AE85u, v10 = 0x3C6EF372u, v11 = 0xA54FF53Au;
  uint32_t v12 = 0x510E527Fu ^ (uint32_t)t, v13 = 0x9B05688Cu ^ (uint32_t)(t >> 32);
  uint32_t v14 = last ? ~0x1F83D9ABu : 0x1F83D9ABu, v15 = 0x5BE0CD19u;
  B2S_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15);
  B2S_ROUND(14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3);
  B2S_ROUND(11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4);
  B2S_ROUND(7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8);
  B2S_ROUND(9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13);
  B2S_ROUND(2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9);
  B2S_ROUND(12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11);
  B2S_ROUND(13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10);
  B2S_ROUND(6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5);
  B2S_ROUND(10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0);
  h[0] ^= v0 ^ v8; h[1] ^= v1 ^ v9; h[2] ^= v2 ^ v10; h[3] ^= v3 ^ v11;
  h[4] ^= v4 ^ v12; h[5] ^= v5 ^ v13; h[6] ^= v6 ^ v14; h[7] ^= v7 ^ v15;
}

__device__ __forceinline__ void blake2s_init(uint32_t (&h)[8]) {
  h[0] = 0x6A09E667u ^ 0x01010020u;  // digest 32 bytes, no key, fanout = depth = 1
  h[1] = 0xBB67AE85u; h[2] = 0x3C6EF372u; h[3] = 0xA54FF53Au;
  h[4] = 0x510E527Fu; h[5] = 0x9B05688Cu; h[6] = 0x1F83D9ABu; h[7] = 0x5BE0CD19u;
}

// One thread per column.  `col` points at the column's first element; consecutive rows are `stride`
// elements apart.  MONT: elements are in Montgomery form and are converted in registers.
// The 64-byte blocks [b0, b1) of the column message are compressed; block b holds rows 2b and 2b+1
// (shifted by the 8-byte length prefix, whose spill-over travels in c0/c1).  A column can therefore be
// hashed in row tiles: the carried state is (h, c0, c1) plus, when a tile ends on an odd row, that row's
// element (`pend`, raw as stored in U), which opens the next tile's first block.  This is what lets a commit
// overlap the hashing of one row tile with the encoding of the next (and, across GPUs, with the rows that
// are still arriving over NVLink).  Rows at or beyond `row_lim` are not read.
constexpr int kHashL2Ahead = 6;  // blocks (pairs of rows) prefetched into L2 ahead of the register prefetch

template <bool PREFIX, bool MONT, bool FMA = true>
__device__ __forceinline__ void hash_column_blocks(const Fr* col, size_t stride, size_t rows, size_t row_lim, uint64_t b0,
                                                   uint64_t b1, bool have_pend, uint32_t (&h)[8], uint32_t& c0,
                                                   uint32_t& c1, Fr& pend) {
  const uint64_t total = (PREFIX ? 8ull : 0ull) + 32ull * rows;
  const uint64_t nblocks = total == 0 ? 1 : (total + 63) / 64;
  const uint32_t b2s_one = g_b2s_one;
  Fr n0 = fr_zero(), n1 = fr_zero();
  if (have_pend) n0 = pend;
  else if (2 * b0 < row_lim) n0 = ld_fr_g(col + 2 * b0 * stride);
  if (2 * b0 + 1 < row_lim) n1 = ld_fr_g(col + (2 * b0 + 1) * stride);
  for (uint64_t b = b0; b < b1; b++) {
    Fr e0 = n0, e1 = n1;
    const size_t r2 = 2 * (b + 1);
    n0 = fr_zero();
    n1 = fr_zero();
    if (r2 < row_lim) n0 = ld_fr_g(col + r2 * stride);          // prefetch the next block's two elements
    if (r2 + 1 < row_lim) n1 = ld_fr_g(col + (r2 + 1) * stride);
    // ... and pull the rows of a few blocks further down into L2: consecutive rows of a column are a whole row
    // pitch apart (new DRAM page, often a new TLB entry), and with few columns per SM nothing else hides that
    if (r2 + 2 * kHashL2Ahead + 1 < row_lim) {
      asm volatile("prefetch.global.L2 [%0];" ::"l"(col + (r2 + 2 * kHashL2Ahead) * stride));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(col + (r2 + 2 * kHashL2Ahead + 1) * stride));
    }
    if (MONT) {
      e0 = fr_from_mont(e0);
      e1 = fr_from_mont(e1);
    }
    uint32_t m[16];
    if (PREFIX) {
      m[0] = c0; m[1] = c1;
#pragma unroll
      for (int i = 0; i < 8; i++) m[2 + i] = e0.v[i];
#pragma unroll
      for (int i = 0; i < 6; i++) m[10 + i] = e1.v[i];
      c0 = e1.v[6]; c1 = e1.v[7];
    } else {
#pragma unroll
      for (int i = 0; i < 8; i++) { m[i] = e0.v[i]; m[8 + i] = e1.v[i]; }
    }
    const bool last = (b + 1 == nblocks);
    const uint64_t t = last ? total : 64 * (b + 1);
    blake2s_compress<FMA>(h, m, t, last, b2s_one);
  }
  pend = n0;  // row 2*b1 if the tile ends on it (odd row_lim), else unused
}

template <bool PREFIX, bool MONT>
__device__ __forceinline__ void hash_one_column(const Fr* col, size_t stride, size_t rows, uint32_t (&h)[8]) {
  blake2s_init(h);
  const uint64_t total = (PREFIX ? 8ull : 0ull) + 32ull * rows;
  const uint64_t nblocks = total == 0 ? 1 : (total + 63) / 64;
  uint32_t c0 = (uint32_t)rows, c1 = (uint32_t)((uint64_t)rows >> 32);  // carry words (PREFIX): u64_le(R) first
  Fr pend = fr_zero();
  hash_column_blocks<PREFIX, MONT>(col, stride, rows, rows, 0, nblocks, false, h, c0, c1, pend);
}

// rows [row0, row_end) of every column.  `state` carries (h[8], c0, c1, pend[8]) per physical column between
// tiles, word-major so a warp's accesses coalesce; the tile that ends at `rows` writes the leaves.
constexpr int kHashStateWords = 18;
template <bool PREFIX, bool FMA>
__global__ void __launch_bounds__(64) hash_columns_kernel(const Fr* __restrict__ u, size_t rows, int log_k, int rho,
                                                          size_t row0, size_t row_end, uint32_t* __restrict__ state,
                                                          uint8_t* __restrict__ leaves, size_t col0, size_t col1) {
  const size_t pc = col0 + (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // physical column = plane * k + c
  const size_t k = (size_t)1 << log_k, ncols = (size_t)rho * k;
  if (pc >= col1) return;
  const size_t s = pc >> log_k, c = pc & (k - 1);
  const uint64_t total = (PREFIX ? 8ull : 0ull) + 32ull * rows;
  const uint64_t nblocks = total == 0 ? 1 : (total + 63) / 64;
  const bool first = row0 == 0, last = row_end >= rows, have_pend = (row0 & 1) != 0;
  uint32_t h[8], c0, c1;
  Fr pend = fr_zero();
  if (first) {
    blake2s_init(h);
    c0 = (uint32_t)rows;
    c1 = (uint32_t)((uint64_t)rows >> 32);
  } else {
#pragma unroll
    for (int i = 0; i < 8; i++) h[i] = state[(size_t)i * ncols + pc];
    c0 = state[8 * ncols + pc];
    c1 = state[9 * ncols + pc];
    if (have_pend) {
#pragma unroll
      for (int i = 0; i < 8; i++) pend.v[i] = state[(size_t)(10 + i) * ncols + pc];
    }
  }
  const size_t lim = row_end < rows ? row_end : rows;
  const uint64_t b0 = row0 / 2, b1 = last ? nblocks : row_end / 2;
  // every plane of a committed matrix holds plain integers (Matrix): no conversion here
  hash_column_blocks<PREFIX, false, FMA>(u + s * rows * k + c, k, rows, lim, b0, b1, have_pend, h, c0, c1, pend);
  if (!last) {
#pragma unroll
    for (int i = 0; i < 8; i++) state[(size_t)i * ncols + pc] = h[i];
    state[8 * ncols + pc] = c0;
    state[9 * ncols + pc] = c1;
    if (row_end & 1) {
#pragma unroll
      for (int i = 0; i < 8; i++) state[(size_t)(10 + i) * ncols + pc] = pend.v[i];
    }
    return;
  }
  uint4* dst = reinterpret_cast<uint4*>(leaves + 32 * ((size_t)rho * c + s));  // logical column rho*c + s
  dst[0] = make_uint4(h[0], h[1], h[2], h[3]);
  dst[1] = make_uint4(h[4], h[5], h[6], h[7]);
}

// columns given explicitly (t opened columns, each `rows` contiguous elements): used by the verifier
template <bool PREFIX>
__global__ void hash_column_list_kernel(const Fr* __restrict__ cols, size_t rows, size_t count,
                                        uint8_t* __restrict__ digests) {
  const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= count) return;
  uint32_t h[8];
  hash_one_column<PREFIX, true>(cols + j * rows, 1, rows, h);
  uint4* dst = reinterpret_cast<uint4*>(digests + 32 * j);
  dst[0] = make_uint4(h[0], h[1], h[2], h[3]);
  dst[1] = make_uint4(h[4], h[5], h[6], h[7]);
}

// ------------------------------------------------------------------------------------------------
// Four lanes per column ("quad" kernel) for FEW columns.
// BLAKE2s over one column is a sequential chain of R/2 compressions.  With n/G columns per GPU (8 192 at G = 8:
// 1.7 warps per SM under the thread-per-column kernel) nothing hides that chain, and one warp issuing all eight
// G functions of a round is held by its own sub-partition's ALU pipe (640 xor/rotate warp instructions at 2 cycles
// each per compression).  Here lane i of a group of four owns state column i (v[i], v[4+i], v[8+i], v[12+i]): a round
// is one G on the columns, three shuffles to diagonalise, one G on the diagonals, three shuffles back -- a quarter of
// the ALU work per warp, spread over four times as many warps and therefore over all four sub-partitions, so the
// pace is the dependent chain itself (measured, 16 388 rows: 1 580 cycles per compression alone, 1 800 with 8 192
// columns on the GPU, against 2 270 for the thread-per-column kernel at any column count up to 16 384).  The message block of a column is staged through shared memory (double
// buffered, one LDG.128 per lane per block, issued one block ahead; each lane then reads the four words its two G's
// need per round at per-lane addresses kept in registers).  The shuffles and the per-lane LDS share the SM's one
// LSU/crossbar (about 1.7 cycles per warp instruction), which makes this variant SLOWER than thread-per-column from
// 16 384 columns on (12.0 against 9.5 ms); hash_columns_range() picks by column count (Ctx::hash_quad_max).
// ------------------------------------------------------------------------------------------------
#define LG_B2S_SIGMA                                                                                                 \
  {                                                                                                                  \
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},     \
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},     \
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},     \
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},     \
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0}      \
  }
__constant__ uint8_t kB2sSigma[10][16] = LG_B2S_SIGMA;
__constant__ uint32_t kB2sIV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au,
                                   0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};

constexpr int kQuadWarps = 2;          // warps per CTA (1, 2, 4 measured: 7.52, 7.52, 7.69 ms at 8 192 columns)
constexpr int kQuadColWords = 36;      // shared-memory words per column: 2 buffers x 16 message words + 4 padding
constexpr int kQuadL2Ahead = 6;

__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts_v2(uint32_t addr, uint32_t x, uint32_t y) {
  asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(x), "r"(y) : "memory");
}

#define B2S_GQ(x, y)               \
  do {                             \
    a = a + b + (x);               \
    d = rotr32(d ^ a, 16);         \
    c = c + d;                     \
    b = rotr32(b ^ c, 12);         \
    a = a + b + (y);               \
    d = rotr32(d ^ a, 8);          \
    c = c + d;                     \
    b = rotr32(b ^ c, 7);          \
  } while (0)

// One compression of the quad's column.  Lane j keeps b_j for the whole round: the diagonal step of G index i runs
// in lane j = i + 1 on (a_{j-1}, b_j, c_{j+1}, d_{j+2}), so what travels between the two half rounds is a, c and d --
// the words a G finishes first -- and the shuffle of the word it finishes last (b, which also opens the next G) is
// never on the dependent chain.
// maddr[r][0..3] = shared addresses (buffer 0) of this lane's message words for the column step (x, y) and the
// diagonal step (x, y) of round r; boff = byte offset of the buffer in use.  (Reading all 16 words with four LDS.128
// and picking the lane's word with selects was measured too: 8.26 ms against 7.70 ms at 8 192 columns -- the selects
// land on the ALU pipe.)
__device__ __forceinline__ void blake2s_compress_quad(uint32_t& h_lo, uint32_t& h_hi, const uint32_t (&maddr)[10][4],
                                                      uint32_t boff, uint32_t iv_c, uint32_t iv_d, uint32_t t_sel,
                                                      int i) {
  uint32_t a = h_lo, b = h_hi, c = iv_c, d = iv_d ^ t_sel;
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t x0 = lds_u32(maddr[r][0] + boff), y0 = lds_u32(maddr[r][1] + boff);
    const uint32_t x1 = lds_u32(maddr[r][2] + boff), y1 = lds_u32(maddr[r][3] + boff);
    B2S_GQ(x0, y0);
    a = __shfl_sync(0xffffffffu, a, (i + 3) & 3, 4);
    d = __shfl_sync(0xffffffffu, d, (i + 2) & 3, 4);
    c = __shfl_sync(0xffffffffu, c, (i + 1) & 3, 4);
    B2S_GQ(x1, y1);
    a = __shfl_sync(0xffffffffu, a, (i + 1) & 3, 4);
    d = __shfl_sync(0xffffffffu, d, (i + 2) & 3, 4);
    c = __shfl_sync(0xffffffffu, c, (i + 3) & 3, 4);
  }
  h_lo ^= a ^ c;
  h_hi ^= b ^ d;
}

template <bool PREFIX>
__global__ void __launch_bounds__(32 * kQuadWarps) hash_columns_quad_kernel(const Fr* __restrict__ u, size_t rows, int log_k,
                                                                            int rho, size_t row0, size_t row_end,
                                                                            uint32_t* __restrict__ state,
                                                                            uint8_t* __restrict__ leaves, size_t col0, size_t col1) {
  __shared__ __align__(16) uint32_t sm[kQuadWarps * 8 * kQuadColWords];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t k = (size_t)1 << log_k, ncols = (size_t)rho * k;
  const size_t pc0 = col0 + ((size_t)blockIdx.x * kQuadWarps + warp) * 8;  // first of this warp's 8 physical columns
  if (pc0 >= col1) return;  // whole warps only (k, col0 and col1 are multiples of 8)
  const size_t s = pc0 >> log_k, c0 = pc0 & (k - 1);
  const Fr* base = u + s * rows * k + c0;
  const uint64_t total = (PREFIX ? 8ull : 0ull) + 32ull * rows;
  const uint64_t nblocks = total == 0 ? 1 : (total + 63) / 64;
  const uint32_t sm_warp = (uint32_t)__cvta_generic_to_shared(sm + warp * 8 * kQuadColWords);

  // ---- compute role: lane (q, i) = G index i of column q
  const int q = lane >> 2, i = lane & 3;
  const uint32_t sm_col = sm_warp + 4u * kQuadColWords * q;
  uint32_t maddr[10][4];
#pragma unroll
  for (int r = 0; r < 10; r++) {
    maddr[r][0] = sm_col + 4u * kB2sSigma[r][2 * i];
    maddr[r][1] = sm_col + 4u * kB2sSigma[r][2 * i + 1];
    maddr[r][2] = sm_col + 4u * kB2sSigma[r][8 + 2 * ((i + 3) & 3)];  // lane j runs diagonal G index j - 1
    maddr[r][3] = sm_col + 4u * kB2sSigma[r][8 + 2 * ((i + 3) & 3) + 1];
  }
  const uint32_t iv_c = kB2sIV[i], iv_d = kB2sIV[4 + i];
  // row tile [row0, row_end): same carried state as the thread-per-column kernel (h, carry words, pending odd row),
  // word-major in `state`; the tile that ends at `rows` writes the leaves
  const bool first = row0 == 0, last_tile = row_end >= rows, have_pend = (row0 & 1) != 0;
  const size_t row_lim = row_end < rows ? row_end : rows;
  const uint64_t b0 = row0 / 2, b1 = last_tile ? nblocks : row_end / 2;
  uint32_t h_lo = iv_c ^ (i == 0 ? 0x01010020u : 0u), h_hi = iv_d;
  if (!first) {
    h_lo = state[(size_t)i * ncols + pc0 + q];
    h_hi = state[(size_t)(4 + i) * ncols + pc0 + q];
  }

  // ---- load role: lane (lrow, lcol, lhalf) fetches 16 bytes: half `lhalf` of the element of column lcol in row
  // 2b + lrow -- lanes 0..15 read 256 contiguous bytes of one row, lanes 16..31 of the next
  const int lrow = lane >> 4, lcol = (lane >> 1) & 7, lhalf = lane & 1;
  const uint4* src = reinterpret_cast<const uint4*>(base + lcol) + lhalf;  // + row * (2k) uint4
  const size_t pitch4 = 2 * k;
  const uint32_t sm_lcol = sm_warp + 4u * kQuadColWords * lcol;
  // message words this lane supplies: with the 8-byte length prefix a block is [carry(2) | row 2b (8) | row 2b+1 (6)]
  // and the last two words of row 2b+1 are carried into the next block by the lane that loaded them
  const bool carrier = PREFIX && lrow == 1 && lhalf == 1;
  const uint32_t w0 = PREFIX ? (2 + 8 * lrow + 4 * lhalf) : (8 * lrow + 4 * lhalf);
  const uint32_t st0 = sm_lcol + 4u * w0;                            // first two words of the 16 bytes
  const uint32_t st1 = carrier ? sm_lcol : st0 + 8u;                 // second two (or the carried pair -> words 0,1)
  uint32_t cy0 = (uint32_t)rows, cy1 = (uint32_t)((uint64_t)rows >> 32);  // block 0 opens with u64_le(R)
  if (!first && carrier) {
    cy0 = state[8 * ncols + pc0 + lcol];
    cy1 = state[9 * ncols + pc0 + lcol];
  }

  uint4 v = make_uint4(0, 0, 0, 0);
  if (have_pend && lrow == 0) {  // the row that opened the previous tile's unfinished block
    const size_t w = (size_t)(10 + 4 * lhalf) * ncols + pc0 + lcol;
    v = make_uint4(state[w], state[w + ncols], state[w + 2 * ncols], state[w + 3 * ncols]);
  } else if (2 * b0 + lrow < row_lim) {
    v = src[(2 * b0 + lrow) * pitch4];
  }
  for (uint64_t b = b0; b < b1; b++) {
    const uint32_t boff = (uint32_t)(b & 1) * 64u;
    if (carrier) {
      sts_v2(st0 + boff, v.x, v.y);
      sts_v2(st1 + boff, cy0, cy1);
      cy0 = v.z;
      cy1 = v.w;
    } else {
      sts_v2(st0 + boff, v.x, v.y);
      sts_v2(st1 + boff, v.z, v.w);
    }
    const size_t rn = 2 * (b + 1) + lrow;  // this lane's row of the next block
    v = make_uint4(0, 0, 0, 0);
    if (rn < row_lim) v = src[rn * pitch4];
    if (rn + 2 * kQuadL2Ahead < row_lim) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + (rn + 2 * kQuadL2Ahead) * pitch4));
    __syncwarp();
    const bool last = (b + 1 == nblocks);
    const uint64_t t = last ? total : 64 * (b + 1);
    const uint32_t t_sel = i == 0 ? (uint32_t)t : i == 1 ? (uint32_t)(t >> 32) : (i == 2 && last) ? 0xffffffffu : 0u;
    blake2s_compress_quad(h_lo, h_hi, maddr, boff, iv_c, iv_d, t_sel, i);
  }
  if (!last_tile) {
    state[(size_t)i * ncols + pc0 + q] = h_lo;
    state[(size_t)(4 + i) * ncols + pc0 + q] = h_hi;
    if (carrier) {
      state[8 * ncols + pc0 + lcol] = cy0;
      state[9 * ncols + pc0 + lcol] = cy1;
    }
    if ((row_end & 1) && lrow == 0) {  // v holds row 2*b1 = row_end - 1, the pending odd row
      const size_t w = (size_t)(10 + 4 * lhalf) * ncols + pc0 + lcol;
      state[w] = v.x;
      state[w + ncols] = v.y;
      state[w + 2 * ncols] = v.z;
      state[w + 3 * ncols] = v.w;
    }
    return;
  }
  const size_t col = c0 + q;
  uint32_t* dst = reinterpret_cast<uint32_t*>(leaves + 32 * ((size_t)rho * col + s));  // logical column rho*c + s
  dst[i] = h_lo;
  dst[4 + i] = h_hi;
}

static void launch_quad(cudaStream_t st, const Fr* u, size_t rows, int log_k, int rho_inv, size_t row0, size_t row_end,
                        uint32_t* state, uint8_t* leaves, bool len_prefix, size_t col0, size_t col1) {
  const unsigned grid = (unsigned)(((col1 - col0) / 8 + kQuadWarps - 1) / kQuadWarps), bs = 32 * kQuadWarps;
  if (len_prefix) hash_columns_quad_kernel<true><<<grid, bs, 0, st>>>(u, rows, log_k, rho_inv, row0, row_end, state, leaves, col0, col1);
  else hash_columns_quad_kernel<false><<<grid, bs, 0, st>>>(u, rows, log_k, rho_inv, row0, row_end, state, leaves, col0, col1);
}

static void launch_thread(cudaStream_t st, const Fr* u, size_t rows, int log_k, int rho_inv, size_t row0, size_t row_end,
                          uint32_t* state, uint8_t* leaves, bool len_prefix, bool fma, size_t col0, size_t col1) {
  const unsigned bs = 64;
  const unsigned grid = (unsigned)((col1 - col0 + bs - 1) / bs);
  if (len_prefix && fma)
    hash_columns_kernel<true, true><<<grid, bs, 0, st>>>(u, rows, log_k, rho_inv, row0, row_end, state, leaves, col0, col1);
  else if (len_prefix)
    hash_columns_kernel<true, false><<<grid, bs, 0, st>>>(u, rows, log_k, rho_inv, row0, row_end, state, leaves, col0, col1);
  else if (fma)
    hash_columns_kernel<false, true><<<grid, bs, 0, st>>>(u, rows, log_k, rho_inv, row0, row_end, state, leaves, col0, col1);
  else
    hash_columns_kernel<false, false><<<grid, bs, 0, st>>>(u, rows, log_k, rho_inv, row0, row_end, state, leaves, col0, col1);
}

int hash_columns_range(Ctx* ctx, cudaStream_t st, const Fr* u, size_t rows, int log_k, int rho_inv, size_t row0,
                       size_t row_end, uint32_t* state, uint8_t* leaves, bool len_prefix) {
  if (row0 >= row_end || row_end > rows || ((row0 > 0 || row_end < rows) && !state))
    return set_error(ctx, ERR_INVALID, "column hashing tile out of range, or a partial tile without a state buffer");
  const size_t n = (size_t)rho_inv << log_k;
  if (log_k >= 3 && n <= ctx->hash_quad_max) {
    launch_quad(st, u, rows, log_k, rho_inv, row0, row_end, state, leaves, len_prefix, 0, n);
    ctx->launches++;
    LG_CUDA(ctx, cudaGetLastError());
    return OK;
  }
  // at most one warp per SM sub-partition: the warp is bound by its own dependent chain and instruction count, so
  // plain additions win (8.16 against 9.52 ms at 16 384 columns x 16 388 rows); with more warps the ALU pipe is the
  // bound and the additions belong on the FMA pipe (26.5 against 30.0 ms at 65 536 columns)
  const bool fma = n > (size_t)128 * ctx->sm_count;
  // (Measured and dropped: giving the thread-per-column kernel a whole number of warps per sub-partition, 3 x 592 x 32 =
  // 56 832 of 65 536 columns, and hashing the other 8 704 four lanes per column on a second stream.  The four-lane
  // chains starve next to warps that saturate the ALU pipe: 37.8 ms against 26.5 ms.)
  launch_thread(st, u, rows, log_k, rho_inv, row0, row_end, state, leaves, len_prefix, fma, 0, n);
  ctx->launches++;
  LG_CUDA(ctx, cudaGetLastError());
  return OK;
}

size_t hash_state_words(size_t n) { return (size_t)kHashStateWords * n; }

int hash_columns(Ctx* ctx, const Fr* u, size_t rows, int log_k, int rho_inv, uint8_t* leaves, bool len_prefix) {
  return hash_columns_range(ctx, ctx->stream, u, rows, log_k, rho_inv, 0, rows, nullptr, leaves, len_prefix);
}

int hash_column_list(Ctx* ctx, const Fr* cols, size_t rows, size_t count, uint8_t* digests, bool len_prefix) {
  if (count == 0) return OK;
  const unsigned bs = 32;
  const unsigned grid = (unsigned)((count + bs - 1) / bs);
  if (len_prefix) hash_column_list_kernel<true><<<grid, bs, 0, ctx->stream>>>(cols, rows, count, digests);
  else hash_column_list_kernel<false><<<grid, bs, 0, ctx->stream>>>(cols, rows, count, digests);
  ctx->launches++;
  LG_CUDA(ctx, cudaGetLastError());
  return OK;
}

// ------------------------------------------------------------------------------------------------
// SHA-256 (FIPS 180-4)
// ------------------------------------------------------------------------------------------------
__constant__ uint32_t kSha256K[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98,
    0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786,
    0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8,
    0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13,
    0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819,
    0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a,
    0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7,
    0xc67178f2};

__device__ __forceinline__ void sha256_compress(uint32_t (&st)[8], uint32_t (&w)[16]) {
  uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#pragma unroll
  for (int i = 0; i < 64; i++) {
    if (i >= 16) {
      const uint32_t w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
      const uint32_t s0 = rotr32(w15, 7) ^ rotr32(w15, 18) ^ (w15 >> 3);
      const uint32_t s1 = rotr32(w2, 17) ^ rotr32(w2, 19) ^ (w2 >> 10);
      w[i & 15] = w[i & 15] + s0 + w[(i + 9) & 15] + s1;
    }
    const uint32_t S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25);
    const uint32_t ch = (e & f) ^ (~e & g);
    const uint32_t t1 = h + S1 + ch + kSha256K[i] + w[i & 15];
    const uint32_t S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22);
    const uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
    const uint32_t t2 = S0 + mj;
    h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
  }
  st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}

__device__ __forceinline__ void sha256_init(uint32_t (&st)[8]) {
  st[0] = 0x6a09e667u; st[1] = 0xbb67ae85u; st[2] = 0x3c6ef372u; st[3] = 0xa54ff53au;
  st[4] = 0x510e527fu; st[5] = 0x9b05688cu; st[6] = 0x1f83d9abu; st[7] = 0x5be0cd19u;
}

// digest words as stored in memory (byte order) <-> big-endian message words
__device__ __forceinline__ void load_digest_be(const uint8_t* p, uint32_t (&d)[8]) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  const uint4 a = q[0], b = q[1];
  d[0] = __byte_perm(a.x, 0, 0x0123); d[1] = __byte_perm(a.y, 0, 0x0123);
  d[2] = __byte_perm(a.z, 0, 0x0123); d[3] = __byte_perm(a.w, 0, 0x0123);
  d[4] = __byte_perm(b.x, 0, 0x0123); d[5] = __byte_perm(b.y, 0, 0x0123);
  d[6] = __byte_perm(b.z, 0, 0x0123); d[7] = __byte_perm(b.w, 0, 0x0123);
}
__device__ __forceinline__ void store_state_be(uint8_t* p, const uint32_t (&st)[8]) {
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4(__byte_perm(st[0], 0, 0x0123), __byte_perm(st[1], 0, 0x0123), __byte_perm(st[2], 0, 0x0123),
                    __byte_perm(st[3], 0, 0x0123));
  q[1] = make_uint4(__byte_perm(st[4], 0, 0x0123), __byte_perm(st[5], 0, 0x0123), __byte_perm(st[6], 0, 0x0123),
                    __byte_perm(st[7], 0, 0x0123));
}

// SHA-256( [u64_le(32)] L [u64_le(32)] R ) -> out
template <bool PREFIX>
__device__ __forceinline__ void sha256_two_to_one(const uint8_t* left, const uint8_t* right, uint8_t* out) {
  uint32_t L[8], Rr[8], st[8], w[16];
  load_digest_be(left, L);
  load_digest_be(right, Rr);
  sha256_init(st);
  if (PREFIX) {
    w[0] = 0x20000000u; w[1] = 0;  // bytes 20 00 00 00 00 00 00 00
#pragma unroll
    for (int i = 0; i < 8; i++) w[2 + i] = L[i];
    w[10] = 0x20000000u; w[11] = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) w[12 + i] = Rr[i];
    sha256_compress(st, w);
#pragma unroll
    for (int i = 0; i < 4; i++) w[i] = Rr[4 + i];
    w[4] = 0x80000000u;
#pragma unroll
    for (int i = 5; i < 15; i++) w[i] = 0;
    w[15] = 80 * 8;
    sha256_compress(st, w);
  } else {
#pragma unroll
    for (int i = 0; i < 8; i++) { w[i] = L[i]; w[8 + i] = Rr[i]; }
    sha256_compress(st, w);
    w[0] = 0x80000000u;
#pragma unroll
    for (int i = 1; i < 15; i++) w[i] = 0;
    w[15] = 64 * 8;
    sha256_compress(st, w);
  }
  store_state_be(out, st);
}

// bottom inner level: nodes[base + i] = H(leaf[2i], leaf[2i+1])
template <bool PREFIX>
__global__ void merkle_bottom_kernel(const uint8_t* __restrict__ leaves, uint8_t* __restrict__ nodes, size_t half) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= half) return;
  sha256_two_to_one<PREFIX>(leaves + 64 * i, leaves + 64 * i + 32, nodes + 32 * (half - 1 + i));
}

// one inner level: `width` nodes starting at heap index width-1, children at 2i+1, 2i+2
__global__ void merkle_level_kernel(uint8_t* nodes, size_t width) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= width) return;
  const size_t idx = width - 1 + i;
  sha256_two_to_one<false>(nodes + 32 * (2 * idx + 1), nodes + 32 * (2 * idx + 2), nodes + 32 * idx);
}

// all levels of width <= blockDim.x in one CTA
__global__ void merkle_top_kernel(uint8_t* nodes, size_t width) {
  for (size_t w = width; w >= 1; w >>= 1) {
    if (threadIdx.x < w) {
      const size_t idx = w - 1 + threadIdx.x;
      sha256_two_to_one<false>(nodes + 32 * (2 * idx + 1), nodes + 32 * (2 * idx + 2), nodes + 32 * idx);
    }
    __syncthreads();
  }
}

int merkle_build(Ctx* ctx, const uint8_t* leaves, size_t n, uint8_t* nodes, bool leaf_len_prefix, cudaStream_t st) {
  if (!st) st = ctx->stream;
  if (n < 2 || (n & (n - 1))) return set_error(ctx, ERR_INVALID, "merkle tree needs a power-of-two number (>1) of leaves");
  const size_t half = n / 2;
  const unsigned bs = 128;
  if (leaf_len_prefix) merkle_bottom_kernel<true><<<(unsigned)((half + bs - 1) / bs), bs, 0, st>>>(leaves, nodes, half);
  else merkle_bottom_kernel<false><<<(unsigned)((half + bs - 1) / bs), bs, 0, st>>>(leaves, nodes, half);
  ctx->launches++;
  size_t w = half / 2;
  for (; w > 256; w >>= 1) {
    merkle_level_kernel<<<(unsigned)((w + bs - 1) / bs), bs, 0, st>>>(nodes, w);
    ctx->launches++;
  }
  if (w >= 1) {
    merkle_top_kernel<<<1, 256, 0, st>>>(nodes, w);
    ctx->launches++;
  }
  LG_CUDA(ctx, cudaGetLastError());
  return OK;
}

}  // namespace lg
