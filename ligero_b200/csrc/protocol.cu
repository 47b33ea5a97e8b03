// Kernels for the three Ligero tests and the openings (sm_100a):
//   K10  ChaCha20 challenge expansion with rejection sampling      (src/utils.rs:23-29)
//   K5   r^T * U_pre                                               (src/matrices/mod.rs:138-149, call mod.rs:658)
//   K6   r^T * A on the structured constraint matrix               (src/matrices/mod.rs:100-110, call mod.rs:722)
//   K7   linear-test polynomial  sum_i r_i(x) * p_i(x)              (src/ligero/mod.rs:723-736)
//   K8   quadratic-test polynomial sum_i r_i (p_x p_y - p_z)        (src/ligero/mod.rs:842-848)
//   K9   column + authentication-path gather                        (src/ligero/mod.rs:935-955)
// K7/K8 use the identity of SURVEY App. D: a polynomial of degree < 2k-1 is fixed by its values on the
// size-2k domain, and U already holds p_i on that domain (coset planes 0 and rho_inv/2), so the 12m+3m
// size-2k transforms of the reference collapse into column-wise dot products plus ONE inverse NTT.
#include <cmath>
#include <cstring>

#include "fr_host.h"
#include "lg_internal.h"

namespace lg {

__device__ __forceinline__ Fr p_ld(const Fr* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  const uint4 a = q[0], b = q[1];
  Fr r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ void p_st(Fr* p, const Fr& x) {
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
  q[1] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
}

// ------------------------------------------------------------------------------------------------
// K10: ChaCha20 (rand_chacha 0.3: 64-bit block counter, stream 0) -> Fr by rejection sampling.
// Candidate c consumes stream words [8c, 8c+8) whether accepted or not (F::rand draws 4 u64 per try),
// so candidates sit at fixed stream positions: block b = c/2 yields candidates 2b and 2b+1.
// ------------------------------------------------------------------------------------------------
struct ChaChaKey { uint32_t k[8]; };

__device__ __forceinline__ uint32_t rotl32(uint32_t x, int n) { return __funnelshift_l(x, x, n); }
#define CC_QR(a, b, c, d)                                   \
  a += b; d = rotl32(d ^ a, 16); c += d; b = rotl32(b ^ c, 12); \
  a += b; d = rotl32(d ^ a, 8);  c += d; b = rotl32(b ^ c, 7);

__device__ __forceinline__ void chacha20_block(const ChaChaKey& key, uint64_t counter, uint32_t (&out)[16]) {
  const uint32_t s0 = 0x61707865u, s1 = 0x3320646eu, s2 = 0x79622d32u, s3 = 0x6b206574u;
  uint32_t x0 = s0, x1 = s1, x2 = s2, x3 = s3, x4 = key.k[0], x5 = key.k[1], x6 = key.k[2], x7 = key.k[3];
  uint32_t x8 = key.k[4], x9 = key.k[5], x10 = key.k[6], x11 = key.k[7];
  uint32_t x12 = (uint32_t)counter, x13 = (uint32_t)(counter >> 32), x14 = 0, x15 = 0;
#pragma unroll
  for (int r = 0; r < 10; r++) {
    CC_QR(x0, x4, x8, x12) CC_QR(x1, x5, x9, x13) CC_QR(x2, x6, x10, x14) CC_QR(x3, x7, x11, x15)
    CC_QR(x0, x5, x10, x15) CC_QR(x1, x6, x11, x12) CC_QR(x2, x7, x8, x13) CC_QR(x3, x4, x9, x14)
  }
  out[0] = x0 + s0; out[1] = x1 + s1; out[2] = x2 + s2; out[3] = x3 + s3;
  out[4] = x4 + key.k[0]; out[5] = x5 + key.k[1]; out[6] = x6 + key.k[2]; out[7] = x7 + key.k[3];
  out[8] = x8 + key.k[4]; out[9] = x9 + key.k[5]; out[10] = x10 + key.k[6]; out[11] = x11 + key.k[7];
  out[12] = x12 + (uint32_t)counter; out[13] = x13 + (uint32_t)(counter >> 32); out[14] = x14; out[15] = x15;
}

// candidate = 8 words, top two bits masked; accepted iff < r.  Returns the masked limbs in `e`.
__device__ __forceinline__ bool candidate(const uint32_t* w, Fr& e) {
#pragma unroll
  for (int i = 0; i < 8; i++) e.v[i] = w[i];
  e.v[7] &= 0x3fffffffu;
  const uint32_t p[8] = {LG_P0, LG_P1, LG_P2, LG_P3, LG_P4, LG_P5, LG_P6, LG_P7};
  bool lt = false, decided = false;
#pragma unroll
  for (int i = 7; i >= 0; i--) {
    if (!decided && e.v[i] != p[i]) {
      lt = e.v[i] < p[i];
      decided = true;
    }
  }
  return lt;  // equal to r is rejected
}

constexpr int kExpandThreads = 256;

// pass 1: accepted candidates per CTA
__global__ void __launch_bounds__(kExpandThreads) expand_count_kernel(ChaChaKey key, uint64_t nblocks, uint32_t* cta_counts) {
  const uint64_t b = (uint64_t)blockIdx.x * kExpandThreads + threadIdx.x;
  uint32_t acc = 0;
  if (b < nblocks) {
    uint32_t w[16];
    chacha20_block(key, b, w);
    Fr e;
    acc = (uint32_t)candidate(w, e) + (uint32_t)candidate(w + 8, e);
  }
  // CTA reduction
  __shared__ uint32_t warp_sums[kExpandThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < kExpandThreads / 32; i++) s += warp_sums[i];
    cta_counts[blockIdx.x] = s;
  }
}

// exclusive scan of the CTA counts (one CTA, 1024 threads, sequential chunks); total -> offsets[n]
__global__ void __launch_bounds__(1024) expand_scan_kernel(const uint32_t* counts, uint64_t* offsets, uint64_t n) {
  __shared__ uint64_t sh[1024];
  const uint64_t per = (n + 1023) / 1024;
  const uint64_t lo = (uint64_t)threadIdx.x * per, hi = lo + per < n ? lo + per : n;
  uint64_t s = 0;
  for (uint64_t i = lo; i < hi; i++) s += counts[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint64_t run = 0;
    for (int i = 0; i < 1024; i++) {
      const uint64_t t = sh[i];
      sh[i] = run;
      run += t;
    }
    offsets[n] = run;
  }
  __syncthreads();
  uint64_t run = sh[threadIdx.x];
  for (uint64_t i = lo; i < hi; i++) {
    offsets[i] = run;
    run += counts[i];
  }
}

// pass 2: recompute the block, scan inside the CTA, scatter the first `count` accepted values
__global__ void __launch_bounds__(kExpandThreads) expand_write_kernel(ChaChaKey key, uint64_t nblocks, const uint64_t* cta_offsets,
                                                                     uint64_t count, Fr* out) {
  const uint64_t b = (uint64_t)blockIdx.x * kExpandThreads + threadIdx.x;
  Fr e0 = fr_zero(), e1 = fr_zero();
  bool a0 = false, a1 = false;
  if (b < nblocks) {
    uint32_t w[16];
    chacha20_block(key, b, w);
    a0 = candidate(w, e0);
    a1 = candidate(w + 8, e1);
  }
  const uint32_t mine = (uint32_t)a0 + (uint32_t)a1;
  // inclusive warp scan
  uint32_t incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if ((threadIdx.x & 31) >= o) incl += t;
  }
  __shared__ uint32_t warp_tot[kExpandThreads / 32];
  if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
  __syncthreads();
  uint32_t before = 0;
  for (int i = 0; i < (int)(threadIdx.x >> 5); i++) before += warp_tot[i];
  uint64_t pos = cta_offsets[blockIdx.x] + before + (incl - mine);
  if (a0) {
    if (pos < count) p_st(out + pos, e0);
    pos++;
  }
  if (a1 && pos < count) p_st(out + pos, e1);
}

int expand_fr(Ctx* ctx, const uint8_t seed[32], size_t count, Fr* out_dev) {
  if (count == 0) return OK;
  phase_mark(ctx, PH_BEGIN);
  ChaChaKey key;
  memcpy(key.k, seed, 32);
  const double p_acc = 0.7561;  // r / 2^254
  double cand = (double)count / p_acc;
  cand += 8.0 * sqrt(cand) + 64.0;
  for (int attempt = 0; attempt < 4; attempt++) {
    const uint64_t nblocks = ((uint64_t)cand + 1) / 2 + 1;
    const uint64_t nctas = (nblocks + kExpandThreads - 1) / kExpandThreads;
    if (nctas > 0x7fffffffull) return set_error(ctx, ERR_INVALID, "challenge vector too long");
    void* scratch;
    LG_TRY(ctx_scratch(ctx, nctas * 4 + (nctas + 1) * 8 + 64, &scratch));
    uint64_t* offsets = (uint64_t*)scratch;              // nctas + 1
    uint32_t* counts = (uint32_t*)(offsets + nctas + 1);  // nctas
    expand_count_kernel<<<(unsigned)nctas, kExpandThreads, 0, ctx->stream>>>(key, nblocks, counts);
    expand_scan_kernel<<<1, 1024, 0, ctx->stream>>>(counts, offsets, nctas);
    expand_write_kernel<<<(unsigned)nctas, kExpandThreads, 0, ctx->stream>>>(key, nblocks, offsets, count, out_dev);
    ctx->launches += 3;
    uint64_t total = 0;
    LG_CUDA(ctx, cudaMemcpyAsync(&total, offsets + nctas, 8, cudaMemcpyDeviceToHost, ctx->stream));
    LG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (total >= count) {
      phase_mark(ctx, PH_EXPAND);
      return OK;
    }
    cand *= 1.25;  // astronomically unlikely; enlarge and redo
  }
  return set_error(ctx, ERR_STATE, "rejection sampling did not produce enough field elements");
}

// ------------------------------------------------------------------------------------------------
// column-wise reductions over the rows of a coset plane
//   MODE 0: out[c] = sum_i w[i]      * X[i][c]                       (w: vector of `rows`)
//   MODE 1: out[c] = sum_i W[i][c]   * X[i][c]                       (W: rows x k matrix)
//   MODE 2: out[c] = sum_i w[i] * (X[i][c] * Y[i][c] - Z[i][c])      (i < rows = m)
// grid = (k / 128, slabs): each CTA owns 128 columns of one slab of rows; partial[slab][c] is reduced
// by col_reduce_final.  Reads are fully coalesced (a warp reads 1 KiB of one row).
// ------------------------------------------------------------------------------------------------
// PLAIN: X (Y, Z) hold plain integers (coset plane >= 1 of a committed matrix): products with a Montgomery-form
// weight then come out plain as well and col_reduce_final puts the R back; MODE 2 first lifts x so that x*y is plain.
template <int MODE, bool PLAIN>
__global__ void __launch_bounds__(128) col_reduce_kernel(const Fr* __restrict__ W, const Fr* __restrict__ X,
                                                         const Fr* __restrict__ Y, const Fr* __restrict__ Z, size_t rows,
                                                         size_t k, size_t rows_per_slab, Fr* __restrict__ partial) {
  const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= k) return;
  const size_t i0 = (size_t)blockIdx.y * rows_per_slab;
  const size_t i1 = i0 + rows_per_slab < rows ? i0 + rows_per_slab : rows;
  Fr acc = fr_zero();
  for (size_t i = i0; i < i1; i++) {
    const Fr x = p_ld(X + i * k + c);
    Fr term;
    if (MODE == 0) {
      term = fr_mul(x, p_ld(W + i));
    } else if (MODE == 1) {
      term = fr_mul(x, p_ld(W + i * k + c));
    } else {
      const Fr y = p_ld(Y + i * k + c), z = p_ld(Z + i * k + c);
      term = fr_mul(fr_sub(fr_mul(PLAIN ? fr_mul(x, fr_r2()) : x, y), z), p_ld(W + i));
    }
    acc = fr_add(acc, term);
  }
  p_st(partial + (size_t)blockIdx.y * k + c, acc);
}

// out[c * out_stride + out_offset] = sum_slab partial[slab][c]
__global__ void col_reduce_final_kernel(const Fr* __restrict__ partial, size_t k, size_t slabs, Fr* __restrict__ out,
                                        size_t out_stride, size_t out_offset, int to_mont) {
  const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= k) return;
  Fr acc = p_ld(partial + c);
  for (size_t s = 1; s < slabs; s++) acc = fr_add(acc, p_ld(partial + s * k + c));
  if (to_mont) acc = fr_mul(acc, fr_r2());
  p_st(out + c * out_stride + out_offset, acc);
}

// mode as above; `out` gets k values at stride/offset (so two calls can interleave even/odd points)
int col_reduce(Ctx* ctx, int mode, const Fr* W, const Fr* X, const Fr* Y, const Fr* Z, size_t rows, size_t k, Fr* out,
               size_t out_stride, size_t out_offset, bool x_plain) {
  if (rows == 0 || k == 0) return set_error(ctx, ERR_INVALID, "empty reduction");
  // enough slabs to fill the machine: ~8 CTAs (of 128 threads) per SM
  const size_t col_ctas = (k + 127) / 128;
  size_t slabs = ((size_t)ctx->sm_count * 8 + col_ctas - 1) / col_ctas;
  if (slabs > rows) slabs = rows;
  if (slabs < 1) slabs = 1;
  if (slabs > 65535) slabs = 65535;
  const size_t rows_per_slab = (rows + slabs - 1) / slabs;
  slabs = (rows + rows_per_slab - 1) / rows_per_slab;
  // partials live after any NTT temporary in the scratch: use a dedicated allocation to stay simple
  Fr* partial;
  LG_CUDA(ctx, cudaMallocAsync(&partial, slabs * k * sizeof(Fr), ctx->stream));
  dim3 grid((unsigned)col_ctas, (unsigned)slabs);
  switch (mode * 2 + (x_plain ? 1 : 0)) {
    case 0: col_reduce_kernel<0, false><<<grid, 128, 0, ctx->stream>>>(W, X, Y, Z, rows, k, rows_per_slab, partial); break;
    case 1: col_reduce_kernel<0, true><<<grid, 128, 0, ctx->stream>>>(W, X, Y, Z, rows, k, rows_per_slab, partial); break;
    case 2: col_reduce_kernel<1, false><<<grid, 128, 0, ctx->stream>>>(W, X, Y, Z, rows, k, rows_per_slab, partial); break;
    case 3: col_reduce_kernel<1, true><<<grid, 128, 0, ctx->stream>>>(W, X, Y, Z, rows, k, rows_per_slab, partial); break;
    case 4: col_reduce_kernel<2, false><<<grid, 128, 0, ctx->stream>>>(W, X, Y, Z, rows, k, rows_per_slab, partial); break;
    default: col_reduce_kernel<2, true><<<grid, 128, 0, ctx->stream>>>(W, X, Y, Z, rows, k, rows_per_slab, partial); break;
  }
  col_reduce_final_kernel<<<(unsigned)col_ctas, 128, 0, ctx->stream>>>(partial, k, slabs, out, out_stride, out_offset,
                                                                     x_plain ? 1 : 0);
  ctx->launches += 2;
  LG_CUDA(ctx, cudaFreeAsync(partial, ctx->stream));
  LG_CUDA(ctx, cudaGetLastError());
  return OK;
}

// ------------------------------------------------------------------------------------------------
// K6: r_a = r^T A for A = [[I_{3mk}, -(Px;Py;Pz)], [0, Padd]]  (src/ligero/mod.rs:423-432).
// r_a[0 .. 3mk) = r[0 .. 3mk) (identity block) and stays where it is; only the last mk entries are
// computed, from the CSC form of the right-hand block: one thread per column gathers its few entries.
// value ids: 0 -> +1, 1 -> -1, v >= 2 -> consts[v - 2].
// ------------------------------------------------------------------------------------------------
__global__ void spmv_right_block_kernel(const uint32_t* __restrict__ col_ptr, const uint32_t* __restrict__ row_idx,
                                        const uint32_t* __restrict__ val_id, const Fr* __restrict__ consts,
                                        const Fr* __restrict__ r, size_t mk, Fr* __restrict__ out) {
  const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= mk) return;
  Fr acc = fr_zero();
  for (uint32_t e = col_ptr[c]; e < col_ptr[c + 1]; e++) {
    const Fr x = p_ld(r + row_idx[e]);
    const uint32_t v = val_id[e];
    if (v == 0) acc = fr_add(acc, x);
    else if (v == 1) acc = fr_sub(acc, x);
    else acc = fr_add(acc, fr_mul(x, p_ld(consts + (v - 2))));
  }
  p_st(out + c, acc);
}

int spmv_right_block(Ctx* ctx, const uint32_t* col_ptr, const uint32_t* row_idx, const uint32_t* val_id, const Fr* consts,
                     const Fr* r, size_t mk, Fr* out) {
  // `out` may alias r + 3mk: every thread reads all its inputs ... from anywhere in r, so write to a
  // separate buffer first when aliasing (the caller passes a distinct buffer)
  spmv_right_block_kernel<<<(unsigned)((mk + 127) / 128), 128, 0, ctx->stream>>>(col_ptr, row_idx, val_id, consts, r, mk, out);
  ctx->launches++;
  LG_CUDA(ctx, cudaGetLastError());
  return OK;
}

// ------------------------------------------------------------------------------------------------
// K9: gather t columns (all rows) and their authentication paths
// ------------------------------------------------------------------------------------------------
__global__ void gather_columns_kernel(const Fr* __restrict__ u, size_t rows, int log_k, int rho, const uint64_t* __restrict__ idx,
                                      size_t t, Fr* __restrict__ out) {
  const size_t k = (size_t)1 << log_k;
  const size_t tot = t * rows;
  for (size_t f = (size_t)blockIdx.x * blockDim.x + threadIdx.x; f < tot; f += (size_t)gridDim.x * blockDim.x) {
    const size_t q = f / rows, i = f % rows;
    const size_t j = idx[q], s = j % rho, c = j / rho;
    Fr x = p_ld(u + s * rows * k + i * k + c);
    x = fr_mul(x, fr_r2());  // the planes hold plain integers (Matrix): back to Montgomery form
    p_st(out + f, x);
  }
}
// sib[q] = leaf[idx ^ 1]; auth[q][d] for d = 0 .. log2(n)-2, root side first
__global__ void gather_paths_kernel(const uint8_t* __restrict__ leaves, const uint8_t* __restrict__ nodes, size_t n, int log_n,
                                    const uint64_t* __restrict__ idx, size_t t, uint8_t* __restrict__ sib, uint8_t* __restrict__ auth) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= t) return;
  const size_t j = idx[q];
  const uint4* s = reinterpret_cast<const uint4*>(leaves + 32 * (j ^ 1));
  uint4* so = reinterpret_cast<uint4*>(sib + 32 * q);
  so[0] = s[0];
  so[1] = s[1];
  size_t cur = n / 2 - 1 + j / 2;  // bottom-level inner node above the leaf
  for (int d = log_n - 2; d >= 0; d--) {
    const size_t sibling = (cur & 1) ? cur + 1 : cur - 1;
    const uint4* a = reinterpret_cast<const uint4*>(nodes + 32 * sibling);
    uint4* ao = reinterpret_cast<uint4*>(auth + 32 * (q * (size_t)(log_n - 1) + d));
    ao[0] = a[0];
    ao[1] = a[1];
    cur = (cur - 1) / 2;
  }
}

int gather_open(Ctx* ctx, const Matrix& m, const uint64_t* idx_dev, size_t t, Fr* cols_dev, uint8_t* sib_dev, uint8_t* auth_dev) {
  if (t == 0) return OK;
  int log_n = 0;
  while (((size_t)1 << log_n) < m.n) log_n++;
  gather_columns_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(m.u, m.rows, m.log_k, m.rho_inv, idx_dev, t, cols_dev);
  ctx->launches++;
  if (sib_dev && auth_dev) {  // columns only: the verifier's gather from its own re-encoded matrix
    gather_paths_kernel<<<(unsigned)((t + 63) / 64), 64, 0, ctx->stream>>>(m.leaves, m.nodes, m.n, log_n, idx_dev, t, sib_dev, auth_dev);
    ctx->launches++;
  }
  LG_CUDA(ctx, cudaGetLastError());
  return OK;
}

// ------------------------------------------------------------------------------------------------
// The verifier's per-column checks (src/ligero/mod.rs:705-707, 822-829, 909-932) for t opened columns held as
// t x rows contiguous Montgomery elements: one CTA per column, partial sums per thread, tree reduction in shared memory.
//   MODE 0: out[q] = sum_i w[i] * col_q[i]                                   (Test-Interleaved, w = r)
//   MODE 1: out[q] = sum_i w[q*rows + i] * col_q[i]                          (Test-Linear, w = the same columns of R_A)
//   MODE 2: out[q] = sum_{i<m} w[i] * (col_q[i] * col_q[m+i] - col_q[2m+i])  (Test-Quadratic, rows = 4m)
// ------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256) column_check_kernel(const Fr* __restrict__ cols, const Fr* __restrict__ w, size_t rows,
                                                           Fr* __restrict__ out) {
  __shared__ Fr part[256];
  const size_t q = blockIdx.x;
  const Fr* col = cols + q * rows;
  const size_t count = MODE == 2 ? rows / 4 : rows;
  Fr acc = fr_zero();
  for (size_t i = threadIdx.x; i < count; i += blockDim.x) {
    Fr term;
    if (MODE == 0) term = fr_mul(p_ld(w + i), p_ld(col + i));
    else if (MODE == 1) term = fr_mul(p_ld(w + q * rows + i), p_ld(col + i));
    else term = fr_mul(p_ld(w + i), fr_sub(fr_mul(p_ld(col + i), p_ld(col + count + i)), p_ld(col + 2 * count + i)));
    acc = fr_add(acc, term);
  }
  part[threadIdx.x] = acc;
  __syncthreads();
  for (unsigned s = blockDim.x / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) part[threadIdx.x] = fr_add(part[threadIdx.x], part[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) p_st(out + q, part[0]);
}

int column_checks(Ctx* ctx, int mode, const Fr* cols, const Fr* w, size_t rows, size_t t, Fr* out) {
  if (t == 0) return OK;
  if (mode == 0) column_check_kernel<0><<<(unsigned)t, 256, 0, ctx->stream>>>(cols, w, rows, out);
  else if (mode == 1) column_check_kernel<1><<<(unsigned)t, 256, 0, ctx->stream>>>(cols, w, rows, out);
  else column_check_kernel<2><<<(unsigned)t, 256, 0, ctx->stream>>>(cols, w, rows, out);
  ctx->launches++;
  LG_CUDA(ctx, cudaGetLastError());
  return OK;
}

}  // namespace lg
