// C ABI (include/ligero_b200.h) over the kernels: context, commit, read-backs, microbenchmarks.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/ligero_b200.h"
#include "fr_host.h"
#include "lg_internal.h"

#include "capi_types.h"

namespace lg {

int set_error(Ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->last_error = msg;
  return code;
}

void phase_mark(Ctx* ctx, int phase_ended) {
  if (!ctx->timing) return;
  cudaEvent_t e;
  if (!ctx->event_pool.empty()) {
    e = ctx->event_pool.back();
    ctx->event_pool.pop_back();
  } else if (cudaEventCreate(&e) != cudaSuccess) {
    return;
  }
  cudaEventRecord(e, ctx->stream);
  ctx->marks.emplace_back(phase_ended, e);
}

int ctx_scratch(Ctx* ctx, size_t bytes, void** out) {
  if (bytes > ctx->scratch_bytes) {
    if (ctx->scratch) {
      LG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      LG_CUDA(ctx, cudaFree(ctx->scratch));
      ctx->scratch = nullptr;
      ctx->scratch_bytes = 0;
    }
    cudaError_t e = cudaMalloc(&ctx->scratch, bytes);
    if (e != cudaSuccess) return set_error(ctx, ERR_NOMEM, std::string("scratch cudaMalloc: ") + cudaGetErrorString(e));
    ctx->scratch_bytes = bytes;
  }
  *out = ctx->scratch;
  return OK;
}

int ctx_host_stage(Ctx* ctx, size_t bytes, void** out) {
  if (bytes > ctx->host_stage_bytes) {
    if (ctx->host_stage) {
      LG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      LG_CUDA(ctx, cudaFreeHost(ctx->host_stage));
      ctx->host_stage = nullptr;
      ctx->host_stage_bytes = 0;
    }
    cudaError_t e = cudaHostAlloc(&ctx->host_stage, bytes, cudaHostAllocDefault);
    if (e != cudaSuccess) return set_error(ctx, ERR_NOMEM, std::string("pinned staging cudaHostAlloc: ") + cudaGetErrorString(e));
    ctx->host_stage_bytes = bytes;
  }
  *out = ctx->host_stage;
  return OK;
}

// true if p is device (or managed) memory
bool is_device_ptr(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// logical row gather: out[i][rho*c + s] = plane[s][row0+i][c]
__global__ void gather_rows_kernel(const Fr* __restrict__ u, size_t rows, int log_k, int rho, size_t row0,
                                   size_t nrows, Fr* __restrict__ out) {
  const size_t k = (size_t)1 << log_k, n = k * rho;
  const size_t tot = nrows * n;
  for (size_t f = (size_t)blockIdx.x * blockDim.x + threadIdx.x; f < tot; f += (size_t)gridDim.x * blockDim.x) {
    const size_t i = f / n, j = f % n;
    const size_t s = j % rho, c = j / rho;
    const uint4* src = reinterpret_cast<const uint4*>(u + s * rows * k + (row0 + i) * k + c);
    const uint4 a = src[0], b = src[1];
    Fr x;
    x.v[0] = a.x; x.v[1] = a.y; x.v[2] = a.z; x.v[3] = a.w;
    x.v[4] = b.x; x.v[5] = b.y; x.v[6] = b.z; x.v[7] = b.w;
    x = fr_mul(x, fr_r2());  // the planes hold plain integers (Matrix): back to Montgomery form
    uint4* dst = reinterpret_cast<uint4*>(out + f);
    dst[0] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
    dst[1] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
  }
}

// ---- integer-peak microbenchmarks --------------------------------------------------------------
__global__ void __launch_bounds__(256) bench_fr_mul_kernel(Fr* io, int iters) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  Fr a = io[t], b = io[t + (size_t)gridDim.x * blockDim.x];
  Fr c = fr_add(a, b), d = fr_sub(a, b);
  for (int i = 0; i < iters; i++) {  // 4 independent multiply chains per thread
    a = fr_mul(a, b);
    b = fr_mul(b, c);
    c = fr_mul(c, d);
    d = fr_mul(d, a);
  }
  io[t] = fr_add(fr_add(a, b), fr_add(c, d));
}

// the encoder's multiplier: table-constant products (fr_lazy.cuh), 4 independent chains per thread;
// BFLY: whole forward butterflies (conditional subtraction, product, add, subtract) instead of bare products
template <bool BFLY>
__global__ void __launch_bounds__(256) bench_fr_shoup_kernel(Fr* io, int iters) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  Fr a = io[t], b = io[t + (size_t)gridDim.x * blockDim.x];
  FrTw tw;
  tw.w = fr_from_mont(a);
  tw.w.v[7] &= 0x0fffffffu;  // < r
  tw.p = fr_shoup_quotient(tw.w);
  Fr c = fr_add(a, b), d = fr_sub(a, b);
  for (int i = 0; i < iters; i++) {
    if (BFLY) {
      lz_bfly_dit(a, b, tw);
      lz_bfly_dit(c, d, tw);
      lz_bfly_dit(b, a, tw);
      lz_bfly_dit(d, c, tw);
    } else {
      a = fr_mul_shoup(a, tw);
      b = fr_mul_shoup(b, tw);
      c = fr_mul_shoup(c, tw);
      d = fr_mul_shoup(d, tw);
    }
  }
  io[t] = fr_add(fr_add(fr_normalize(a), fr_normalize(b)), fr_add(fr_normalize(c), fr_normalize(d)));
}

// Raw rate of the multiplier's own instruction: 32x32->64 multiply-accumulates in the carry-chained form the Shoup body
// issues (mad.lo.cc / madc.hi.cc / madc.lo.cc / madc.hi -> IMAD.WIDE.U32 + IMAD.WIDE.U32.X), 32 independent chains of two
// per thread and iteration (64 wide multiplies), multiplicands rewritten every iteration so that nothing is loop
// invariant (ptxas hoists invariant products: the round-1 probe measured 64-bit additions).  2048 threads per SM.
__global__ void __launch_bounds__(256) bench_imad_wide_kernel(uint64_t* io, int iters) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t a[8], b[8], lo[8], hi[8];
  const uint64_t s0 = io[t];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    a[i] = (uint32_t)(s0 >> i) * 2654435761u + i;
    b[i] = (uint32_t)(s0 >> (i + 8)) * 40503u + 7 * i + 1;
    lo[i] = (uint32_t)s0 + i;
    hi[i] = (uint32_t)(s0 >> 32) ^ i;
  }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
#pragma unroll
      for (int i = 0; i < 8; i += 2)
        asm volatile("mad.lo.cc.u32 %0, %4, %5, %0; madc.hi.cc.u32 %1, %4, %5, %1; madc.lo.cc.u32 %2, %4, %6, %2; madc.hi.u32 %3, %4, %6, %3;"
                     : "+r"(lo[i]), "+r"(hi[i]), "+r"(lo[i + 1]), "+r"(hi[i + 1])
                     : "r"(a[j]), "r"(b[i]), "r"(b[i + 1]));
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {  // 16 ALU-pipe instructions per 64 multiplies
      a[i] ^= hi[i];
      b[i] += lo[(i + 3) & 7];
    }
  }
  uint64_t r = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) r += ((uint64_t)hi[i] << 32) + lo[i] + a[i] + b[i];
  io[t] = r;
}

static int matrix_alloc(lg_ctx* ctx, size_t rows, size_t k, uint32_t rho_inv, lg_matrix** out, void* external_u = nullptr) {
  Ctx* c = &ctx->c;
  if (!out) return set_error(c, ERR_INVALID, "null output handle");
  if (rows == 0) return set_error(c, ERR_INVALID, "rows must be > 0");
  if (k < 2 || (k & (k - 1))) return set_error(c, ERR_INVALID, "k must be a power of two >= 2");
  if (rho_inv < 2 || (rho_inv & (rho_inv - 1))) return set_error(c, ERR_INVALID, "rho_inv must be a power of two >= 2");
  int log_k = 0;
  while (((size_t)1 << log_k) < k) log_k++;
  lg_matrix* h = new (std::nothrow) lg_matrix();
  if (!h) return set_error(c, ERR_NOMEM, "host allocation failed");
  h->owner = ctx;
  Matrix& m = h->m;
  m.ctx = c;
  m.rows = rows;
  m.log_k = log_k;
  m.rho_inv = (int)rho_inv;
  m.k = k;
  m.n = k * rho_inv;
  cudaError_t e = cudaSuccess;
  if (external_u) {
    m.u = (Fr*)external_u;
    m.owns_u = false;
  } else {
    e = cudaMalloc(&m.u, m.n * rows * sizeof(Fr));
  }
  if (e == cudaSuccess) e = cudaMalloc(&m.leaves, m.n * 32);
  if (e == cudaSuccess) e = cudaMalloc(&m.nodes, m.n * 32);
  if (e != cudaSuccess) {
    cudaGetLastError();
    if (m.u && m.owns_u) cudaFree(m.u);
    if (m.leaves) cudaFree(m.leaves);
    if (m.nodes) cudaFree(m.nodes);
    delete h;
    return set_error(c, ERR_NOMEM, std::string("matrix cudaMalloc: ") + cudaGetErrorString(e));
  }
  *out = h;
  return OK;
}

// stage a host matrix on the device (second scratch region after the NTT temporary)
static int stage_input(lg_ctx* ctx, const uint64_t* src, size_t elems, const Fr** dev, void** to_free) {
  Ctx* c = &ctx->c;
  *to_free = nullptr;
  if (is_device_ptr(src)) {
    *dev = reinterpret_cast<const Fr*>(src);
    return OK;
  }
  void* d = nullptr;
  cudaError_t e = cudaMalloc(&d, elems * sizeof(Fr));
  if (e != cudaSuccess) return set_error(c, ERR_NOMEM, std::string("input staging cudaMalloc: ") + cudaGetErrorString(e));
  e = cudaMemcpyAsync(d, src, elems * sizeof(Fr), cudaMemcpyHostToDevice, c->stream);
  if (e != cudaSuccess) {
    cudaFree(d);
    return set_error(c, ERR_CUDA, std::string("H2D copy: ") + cudaGetErrorString(e));
  }
  *dev = reinterpret_cast<const Fr*>(d);
  *to_free = d;
  return OK;
}

// second stream (lowest priority: it only fills what the encoder leaves free) + per-column state for tile-wise hashing
int hash_pipeline_setup(Ctx* c, size_t n) {
  if (!c->hash_stream) {
    int lo = 0, hi = 0;
    LG_CUDA(c, cudaDeviceGetStreamPriorityRange(&lo, &hi));
    LG_CUDA(c, cudaStreamCreateWithPriority(&c->hash_stream, cudaStreamNonBlocking, lo));
    LG_CUDA(c, cudaEventCreateWithFlags(&c->ev_encoded, cudaEventDisableTiming));
    LG_CUDA(c, cudaEventCreateWithFlags(&c->ev_hashed, cudaEventDisableTiming));
  }
  if (c->hash_state_words < hash_state_words(n)) {
    LG_CUDA(c, cudaStreamSynchronize(c->stream));
    LG_CUDA(c, cudaStreamSynchronize(c->hash_stream));
    if (c->hash_state) cudaFree(c->hash_state);
    c->hash_state = nullptr;
    LG_CUDA(c, cudaMalloc(&c->hash_state, hash_state_words(n) * sizeof(uint32_t)));
    c->hash_state_words = hash_state_words(n);
  }
  return OK;
}

// Row-tile pipeline of the commit.
//  * host input: tile i+1 is uploaded on a copy stream while tile i is being encoded (pinned host memory makes
//    the copies truly asynchronous; pageable memory still works);
//  * `hash`: the BLAKE2s column hashing of tile i runs on a second, high-priority stream while tile i+1 is being
//    encoded.  The encoder is bound by the integer-multiply pipe and the hash by the ALU pipe, so the two kernels
//    share an SM instead of queueing; the per-column hash state travels between tiles in ctx->hash_state.
// Every tile is encoded with an OutMap that drops its rows at their final position in the plane layout, so no
// tile ever needs a second pass.
static int encode_tiled(lg_matrix* h, const uint64_t* src, bool host, bool hash) {
  Matrix& m = h->m;
  Ctx* c = m.ctx;
  const size_t row_bytes = m.k * sizeof(Fr);
  size_t tile_rows = (((size_t)256 << 20) / row_bytes) & ~(size_t)1;
  if (tile_rows < 2) tile_rows = 2;
  if (tile_rows > m.rows) tile_rows = m.rows;
  const size_t tile_bytes = tile_rows * row_bytes;
  if (host && !c->copy_stream) {
    LG_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
      LG_CUDA(c, cudaEventCreateWithFlags(&c->ev_copied[i], cudaEventDisableTiming));
      LG_CUDA(c, cudaEventCreateWithFlags(&c->ev_consumed[i], cudaEventDisableTiming));
    }
  }
  if (hash) LG_TRY(hash_pipeline_setup(c, m.n));
  if (host && c->stage_bytes < tile_bytes) {
    LG_CUDA(c, cudaStreamSynchronize(c->stream));
    LG_CUDA(c, cudaStreamSynchronize(c->copy_stream));
    for (int i = 0; i < 2; i++) {
      if (c->stage[i]) cudaFree(c->stage[i]);
      c->stage[i] = nullptr;
      LG_CUDA(c, cudaMalloc(&c->stage[i], tile_bytes));
    }
    c->stage_bytes = tile_bytes;
  }
  const size_t cos_bytes = m.log_k > 10 ? (size_t)(m.rho_inv - 1) * tile_bytes : 0;
  if (c->tile_cosets_bytes < cos_bytes) {
    LG_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->tile_cosets) cudaFree(c->tile_cosets);
    c->tile_cosets = nullptr;
    LG_CUDA(c, cudaMalloc(&c->tile_cosets, cos_bytes));
    c->tile_cosets_bytes = cos_bytes;
  }
  OutMap map{};
  map.base[0] = m.u;
  map.log_kg = m.log_k;
  map.rows_total = m.rows;
  if (host) {
    // the copy stream must not overwrite a staging tile that an earlier call is still reading
    LG_CUDA(c, cudaEventRecord(c->ev_consumed[0], c->stream));
    LG_CUDA(c, cudaEventRecord(c->ev_consumed[1], c->stream));
  }
  int t = 0;
  for (size_t row0 = 0; row0 < m.rows; row0 += tile_rows, t++) {
    const size_t nr = row0 + tile_rows <= m.rows ? tile_rows : m.rows - row0;
    const Fr* tile = reinterpret_cast<const Fr*>(src) + row0 * m.k;
    const int b = t & 1;
    if (host) {
      LG_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->ev_consumed[b], 0));
      LG_CUDA(c, cudaMemcpyAsync(c->stage[b], (const uint8_t*)src + row0 * row_bytes, nr * row_bytes,
                                 cudaMemcpyHostToDevice, c->copy_stream));
      LG_CUDA(c, cudaEventRecord(c->ev_copied[b], c->copy_stream));
      LG_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_copied[b], 0));
      tile = (const Fr*)c->stage[b];
    }
    map.m = map.m_g = (uint32_t)nr;  // one block: grow(i) = row0 + i
    map.i0 = (uint32_t)row0;
    LG_TRY(encode_rows(c, tile, nr, m.log_k, m.rho_inv, nullptr, (Fr*)c->tile_cosets, &map, true));
    if (host) LG_CUDA(c, cudaEventRecord(c->ev_consumed[b], c->stream));
    if (hash) {
      LG_CUDA(c, cudaEventRecord(c->ev_encoded, c->stream));
      LG_CUDA(c, cudaStreamWaitEvent(c->hash_stream, c->ev_encoded, 0));
      LG_TRY(hash_columns_range(c, c->hash_stream, m.u, m.rows, m.log_k, m.rho_inv, row0, row0 + nr, c->hash_state,
                                m.leaves, h->owner->col_len_prefix));
    }
  }
  if (hash) {
    LG_TRY(merkle_build(c, m.leaves, m.n, m.nodes, h->owner->leaf_len_prefix, c->hash_stream));
    LG_CUDA(c, cudaEventRecord(c->ev_hashed, c->hash_stream));
    LG_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_hashed, 0));
  }
  return OK;
}

// The same upload/encode overlap for a rank's row shard (multi-GPU): the local matrix arrives from the host in row
// tiles while the previous tile is encoded and scattered to the column shards through `map` (row_off selects the
// tile's rows inside the local [X_g; Y_g; Z_g; W_g] order).
static int encode_sharded_tiled(lg_ctx* ctx, const uint64_t* src, size_t rows, int log_k, int rho_inv, OutMap map, bool plain) {
  Ctx* c = &ctx->c;
  const size_t k = (size_t)1 << log_k, row_bytes = k * sizeof(Fr);
  size_t tile_rows = (((size_t)256 << 20) / row_bytes) & ~(size_t)1;
  if (tile_rows < 2) tile_rows = 2;
  if (tile_rows > rows) tile_rows = rows;
  const size_t tile_bytes = tile_rows * row_bytes;
  if (!c->copy_stream) {
    LG_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
      LG_CUDA(c, cudaEventCreateWithFlags(&c->ev_copied[i], cudaEventDisableTiming));
      LG_CUDA(c, cudaEventCreateWithFlags(&c->ev_consumed[i], cudaEventDisableTiming));
    }
  }
  if (c->stage_bytes < tile_bytes) {
    LG_CUDA(c, cudaStreamSynchronize(c->stream));
    LG_CUDA(c, cudaStreamSynchronize(c->copy_stream));
    for (int i = 0; i < 2; i++) {
      if (c->stage[i]) cudaFree(c->stage[i]);
      c->stage[i] = nullptr;
      LG_CUDA(c, cudaMalloc(&c->stage[i], tile_bytes));
    }
    c->stage_bytes = tile_bytes;
  }
  const size_t cos_bytes = log_k > 10 ? (size_t)(rho_inv - 1) * tile_bytes : 0;
  if (c->tile_cosets_bytes < cos_bytes) {
    LG_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->tile_cosets) cudaFree(c->tile_cosets);
    c->tile_cosets = nullptr;
    LG_CUDA(c, cudaMalloc(&c->tile_cosets, cos_bytes));
    c->tile_cosets_bytes = cos_bytes;
  }
  // The two staging tiles alternate ACROSS calls (Ctx::stage_seq): the multi-GPU committer encodes its rows run by run,
  // one call per run, and the upload of run j+1 must overlap the encoding of run j.  A tile is overwritten only after the
  // encode that read it (ev_consumed, recorded after every use here and in encode_tiled).
  for (size_t row0 = 0; row0 < rows; row0 += tile_rows) {
    const size_t nr = row0 + tile_rows <= rows ? tile_rows : rows - row0;
    const int b = (int)(c->stage_seq++ & 1);
    LG_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->ev_consumed[b], 0));
    LG_CUDA(c, cudaMemcpyAsync(c->stage[b], (const uint8_t*)src + row0 * row_bytes, nr * row_bytes, cudaMemcpyHostToDevice,
                               c->copy_stream));
    LG_CUDA(c, cudaEventRecord(c->ev_copied[b], c->copy_stream));
    LG_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_copied[b], 0));
    map.row_off = (uint32_t)row0;
    LG_TRY(encode_rows(c, (const Fr*)c->stage[b], nr, log_k, rho_inv, nullptr, (Fr*)c->tile_cosets, &map, plain));
    LG_CUDA(c, cudaEventRecord(c->ev_consumed[b], c->stream));
  }
  return OK;
}

static bool tiled_shape(const Matrix& m) {
  return m.rows * m.k * sizeof(Fr) >= ((size_t)64 << 20) && m.rows < ((size_t)1 << 31);
}

static int do_encode(lg_matrix* h, const uint64_t* preenc_u) {
  Matrix& m = h->m;
  Ctx* c = m.ctx;
  if (!preenc_u) return set_error(c, ERR_INVALID, "null input matrix");
  if (!is_device_ptr(preenc_u) && tiled_shape(m)) return encode_tiled(h, preenc_u, true, false);
  const Fr* dev;
  void* to_free;
  LG_TRY(stage_input(h->owner, preenc_u, m.rows * m.k, &dev, &to_free));
  int s = encode_rows(c, dev, m.rows, m.log_k, m.rho_inv, m.u, m.u + m.rows * m.k, nullptr, true);
  if (to_free) {
    cudaStreamSynchronize(c->stream);
    cudaFree(to_free);
  }
  return s;
}

static int do_hash(lg_matrix* h, uint8_t root_out[32]) {
  Matrix& m = h->m;
  Ctx* c = m.ctx;
  phase_mark(c, PH_BEGIN);
  LG_TRY(hash_columns(c, m.u, m.rows, m.log_k, m.rho_inv, m.leaves, h->owner->col_len_prefix));
  phase_mark(c, PH_HASH);
  LG_TRY(merkle_build(c, m.leaves, m.n, m.nodes, h->owner->leaf_len_prefix));
  phase_mark(c, PH_MERKLE);
  if (root_out) {
    LG_CUDA(c, cudaMemcpyAsync(root_out, m.nodes, 32, cudaMemcpyDeviceToHost, c->stream));
    LG_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  return OK;
}

// encode + hash + Merkle tree.  Large matrices go through the row-tile pipeline with the hashing overlapped
// (ctx->overlap, on by default; lg_ctx_set_overlap(ctx, 0) serialises the kernels, e.g. to time them one by one)
static int do_commit(lg_matrix* h, const uint64_t* preenc_u, uint8_t root_out[32]) {
  Matrix& m = h->m;
  Ctx* c = m.ctx;
  if (!preenc_u) return set_error(c, ERR_INVALID, "null input matrix");
  // host input goes through the row-tile pipeline anyway (upload overlapped with encoding); hashing tile by tile
  // there finishes most of the column hashes before the last tile is encoded.  With the matrix already in HBM the
  // plain sequence is faster: both kernels saturate the SM's issue port, so sharing an SM gains nothing (measured)
  if (c->overlap && tiled_shape(m) && !is_device_ptr(preenc_u)) {
    LG_TRY(encode_tiled(h, preenc_u, true, true));
    if (root_out) {
      LG_CUDA(c, cudaMemcpyAsync(root_out, m.nodes, 32, cudaMemcpyDeviceToHost, c->stream));
      LG_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    return OK;
  }
  LG_TRY(do_encode(h, preenc_u));
  return do_hash(h, root_out);
}

}  // namespace lg

using namespace lg;

extern "C" {

int lg_version(void) { return 1; }

int lg_ctx_create(int device, lg_ctx** out) {
  if (!out) return ERR_INVALID;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
    cudaGetLastError();
    return ERR_CUDA;  // no usable GPU: the product has no CPU fallback
  }
  if (cudaSetDevice(device) != cudaSuccess) return ERR_CUDA;
  lg_ctx* ctx = new (std::nothrow) lg_ctx();
  if (!ctx) return ERR_NOMEM;
  ctx->c.device = device;
  // the main stream gets the greatest priority: the overlapped column hashing (lowest priority, its own stream)
  // then only fills the registers and issue slots the encoder leaves free
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  // (one step below the greatest priority, which is kept for the latency-bound hash stream of the multi-GPU pipeline)
  const int prio_main = prio_hi < prio_lo ? prio_hi + 1 : prio_hi;
  if (cudaStreamCreateWithPriority(&ctx->c.stream, cudaStreamNonBlocking, prio_main) != cudaSuccess) {
    delete ctx;
    return ERR_CUDA;
  }
  cudaDeviceGetAttribute(&ctx->c.sm_count, cudaDevAttrMultiProcessorCount, device);
  {
    // the tests allocate their temporaries (up to two 4mk-element vectors) stream-ordered; keep freed blocks in the
    // pool instead of returning them to the driver at every synchronisation, or each proof pays the allocation again
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
  }
  if (const char* e = getenv("LG_OVERLAP")) ctx->c.overlap = atoi(e) != 0;  // tuning hook; see lg_ctx_set_overlap
  if (const char* e = getenv("LG_NTT_PERSIST")) ctx->c.persist_groups = atoi(e) == 0 ? 0 : 3;
  if (const char* e = getenv("LG_NTT_GROUPS"))
    if (ctx->c.persist_groups && atoi(e) == 2) ctx->c.persist_groups = 2;
  if (const char* e = getenv("LG_HASH_QUAD_MAX")) ctx->c.hash_quad_max = (size_t)strtoull(e, nullptr, 10);
  *out = ctx;
  return OK;
}

int lg_ctx_destroy(lg_ctx* ctx) {
  if (!ctx) return OK;
  cudaSetDevice(ctx->c.device);
  cudaStreamSynchronize(ctx->c.stream);
  for (auto& kv : ctx->c.tables) {
    cudaFree(kv.second.w_fwd);  // w_inv and scale live in the same allocation
    if (kv.second.scale4) cudaFree(kv.second.scale4);
  }
  if (ctx->c.scratch) cudaFree(ctx->c.scratch);
  if (ctx->c.copy_stream) {
    cudaStreamSynchronize(ctx->c.copy_stream);
    for (int i = 0; i < 2; i++) {
      if (ctx->c.stage[i]) cudaFree(ctx->c.stage[i]);
      if (ctx->c.ev_copied[i]) cudaEventDestroy(ctx->c.ev_copied[i]);
      if (ctx->c.ev_consumed[i]) cudaEventDestroy(ctx->c.ev_consumed[i]);
    }
    cudaStreamDestroy(ctx->c.copy_stream);
  }
  if (ctx->c.open_stream) {
    cudaStreamSynchronize(ctx->c.open_stream);
    cudaStreamDestroy(ctx->c.open_stream);
    cudaEventDestroy(ctx->c.ev_gathered);
  }
  if (ctx->c.hash_stream_hi) {
    cudaStreamSynchronize(ctx->c.hash_stream_hi);
    cudaStreamDestroy(ctx->c.hash_stream_hi);
  }
  if (ctx->c.hash_stream) {
    cudaStreamSynchronize(ctx->c.hash_stream);
    cudaStreamDestroy(ctx->c.hash_stream);
    cudaEventDestroy(ctx->c.ev_encoded);
    cudaEventDestroy(ctx->c.ev_hashed);
    if (ctx->c.hash_state) cudaFree(ctx->c.hash_state);
  }
  if (ctx->c.tile_cosets) cudaFree(ctx->c.tile_cosets);
  if (ctx->c.host_stage) cudaFreeHost(ctx->c.host_stage);
  for (auto& m : ctx->c.marks) cudaEventDestroy(m.second);
  for (auto& e : ctx->c.event_pool) cudaEventDestroy(e);
  cudaStreamDestroy(ctx->c.stream);
  delete ctx;
  return OK;
}

const char* lg_last_error(const lg_ctx* ctx) { return ctx ? ctx->c.last_error.c_str() : "null context"; }

int lg_ctx_sync(lg_ctx* ctx) {
  if (!ctx) return ERR_INVALID;
  LG_CUDA(&ctx->c, cudaStreamSynchronize(ctx->c.stream));
  return OK;
}

uint64_t lg_ctx_launches(const lg_ctx* ctx) { return ctx ? ctx->c.launches : 0; }

int lg_ctx_set_formats(lg_ctx* ctx, int col_len_prefix, int leaf_len_prefix) {
  if (!ctx) return ERR_INVALID;
  ctx->col_len_prefix = col_len_prefix != 0;
  ctx->leaf_len_prefix = leaf_len_prefix != 0;
  return OK;
}

void* lg_ctx_stream(const lg_ctx* ctx) { return ctx ? (void*)ctx->c.stream : nullptr; }

int lg_ctx_set_overlap(lg_ctx* ctx, int enabled) {
  if (!ctx) return ERR_INVALID;
  ctx->c.overlap = enabled != 0;
  return OK;
}

int lg_ctx_set_hash_quad_max(lg_ctx* ctx, size_t max_columns) {
  if (!ctx) return ERR_INVALID;
  ctx->c.hash_quad_max = max_columns;
  return OK;
}

int lg_ctx_set_timing(lg_ctx* ctx, int enabled) {
  if (!ctx) return ERR_INVALID;
  ctx->c.timing = enabled != 0;
  return OK;
}

int lg_ctx_phase_ms(lg_ctx* ctx, double* ms_out, uint64_t* count_out, int n) {
  if (!ctx) return ERR_INVALID;
  Ctx* c = &ctx->c;
  LG_CUDA(c, cudaStreamSynchronize(c->stream));
  for (int i = 0; i < n; i++) {
    if (ms_out) ms_out[i] = 0;
    if (count_out) count_out[i] = 0;
  }
  for (size_t i = 1; i < c->marks.size(); i++) {
    const int ph = c->marks[i].first;
    if (ph < 0 || ph >= n) continue;
    float ms = 0;
    if (cudaEventElapsedTime(&ms, c->marks[i - 1].second, c->marks[i].second) == cudaSuccess) {
      if (ms_out) ms_out[ph] += ms;
      if (count_out) count_out[ph] += 1;
    }
  }
  for (auto& m : c->marks) c->event_pool.push_back(m.second);
  c->marks.clear();
  return OK;
}

int lg_encode(lg_ctx* ctx, const uint64_t* preenc_u, size_t rows, size_t k, uint32_t rho_inv, lg_matrix** out) {
  if (!ctx) return ERR_INVALID;
  cudaSetDevice(ctx->c.device);
  lg_matrix* h = nullptr;
  LG_TRY(matrix_alloc(ctx, rows, k, rho_inv, &h));
  int s = do_encode(h, preenc_u);
  if (s != OK) {
    lg_matrix_free(h);
    return s;
  }
  *out = h;
  return OK;
}

int lg_matrix_wrap(lg_ctx* ctx, uint64_t* u_dev, size_t rows, size_t k, uint32_t rho_inv, lg_matrix** out) {
  if (!ctx) return ERR_INVALID;
  cudaSetDevice(ctx->c.device);
  if (!u_dev || !is_device_ptr(u_dev)) return set_error(&ctx->c, ERR_INVALID, "lg_matrix_wrap needs a device buffer");
  return matrix_alloc(ctx, rows, k, rho_inv, out, u_dev);
}

int lg_matrix_encode(lg_matrix* m, const uint64_t* preenc_u) {
  if (!m) return ERR_INVALID;
  cudaSetDevice(m->m.ctx->device);
  return do_encode(m, preenc_u);
}

int lg_matrix_hash(lg_matrix* m, uint8_t root_out[32]) {
  if (!m) return ERR_INVALID;
  cudaSetDevice(m->m.ctx->device);
  return do_hash(m, root_out);
}

int lg_commit(lg_ctx* ctx, const uint64_t* preenc_u, size_t rows, size_t k, uint32_t rho_inv, uint8_t root_out[32],
              lg_matrix** out) {
  if (!ctx) return ERR_INVALID;
  lg_matrix* h = nullptr;
  cudaSetDevice(ctx->c.device);
  LG_TRY(matrix_alloc(ctx, rows, k, rho_inv, &h));
  int s = do_commit(h, preenc_u, root_out);
  if (s != OK) {
    lg_matrix_free(h);
    return s;
  }
  *out = h;
  return OK;
}

int lg_recommit(lg_matrix* m, const uint64_t* preenc_u, uint8_t root_out[32]) {
  if (!m) return ERR_INVALID;
  cudaSetDevice(m->m.ctx->device);
  return do_commit(m, preenc_u, root_out);
}

int lg_matrix_free(lg_matrix* m) {
  if (!m) return OK;
  cudaSetDevice(m->m.ctx->device);
  cudaStreamSynchronize(m->m.ctx->stream);
  if (m->m.owns_u && m->m.u) cudaFree(m->m.u);
  if (m->m.leaves) cudaFree(m->m.leaves);
  if (m->m.nodes) cudaFree(m->m.nodes);
  delete m;
  return OK;
}

int lg_matrix_dims(const lg_matrix* m, size_t* rows, size_t* k, size_t* n) {
  if (!m) return ERR_INVALID;
  if (rows) *rows = m->m.rows;
  if (k) *k = m->m.k;
  if (n) *n = m->m.n;
  return OK;
}

int lg_matrix_create(lg_ctx* ctx, size_t rows, size_t k, uint32_t rho_inv, lg_matrix** out) {
  if (!ctx) return ERR_INVALID;
  cudaSetDevice(ctx->c.device);
  return matrix_alloc(ctx, rows, k, rho_inv, out);
}

int lg_ipc_export(const lg_matrix* m, uint8_t handle_out[64]) {
  if (!m || !handle_out) return ERR_INVALID;
  Ctx* c = m->m.ctx;
  cudaSetDevice(c->device);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  cudaIpcMemHandle_t h;
  LG_CUDA(c, cudaIpcGetMemHandle(&h, m->m.u));
  memcpy(handle_out, &h, 64);
  return OK;
}

int lg_ipc_open(lg_ctx* ctx, const uint8_t handle[64], void** ptr_out) {
  if (!ctx || !handle || !ptr_out) return ERR_INVALID;
  Ctx* c = &ctx->c;
  cudaSetDevice(c->device);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  LG_CUDA(c, cudaIpcOpenMemHandle(ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
  return OK;
}

int lg_ipc_close(lg_ctx* ctx, void* ptr) {
  if (!ctx) return ERR_INVALID;
  Ctx* c = &ctx->c;
  cudaSetDevice(c->device);
  LG_CUDA(c, cudaStreamSynchronize(c->stream));
  LG_CUDA(c, cudaIpcCloseMemHandle(ptr));
  return OK;
}

int lg_encode_sharded(lg_ctx* ctx, const uint64_t* msg_local, size_t m_g, size_t k, uint32_t rho_inv, void* const* shard_u,
                      int world, size_t m, size_t i0, uint64_t* cosets_scratch, int plain) {
  if (!ctx || !shard_u) return ERR_INVALID;
  Ctx* c = &ctx->c;
  cudaSetDevice(c->device);
  if (world < 1 || world > kMaxRanks || (world & (world - 1))) return set_error(c, ERR_INVALID, "world must be a power of two <= 8");
  if (k < 2 || (k & (k - 1)) || k % world) return set_error(c, ERR_INVALID, "k must be a power of two divisible by world");
  if (i0 + m_g > m) return set_error(c, ERR_INVALID, "row slice out of range");
  if (m_g == 0) return OK;
  if (!msg_local) return set_error(c, ERR_INVALID, "null input matrix");
  int log_k = 0, log_w = 0;
  while (((size_t)1 << log_k) < k) log_k++;
  while ((1 << log_w) < world) log_w++;
  if (log_k > 10 && !cosets_scratch) return set_error(c, ERR_INVALID, "rows longer than 1024 need the local coset scratch buffer");
  OutMap map{};
  for (int g = 0; g < world; g++) {
    if (!shard_u[g]) return set_error(c, ERR_INVALID, "null shard pointer");
    map.base[g] = (Fr*)shard_u[g];
  }
  map.log_kg = log_k - log_w;
  map.m = (uint32_t)m;
  map.m_g = (uint32_t)m_g;
  map.i0 = (uint32_t)i0;
  map.rows_total = 4ull * m;
  const size_t rows = 4 * m_g;
  if (!is_device_ptr(msg_local) && rows * k * sizeof(Fr) >= ((size_t)64 << 20) && rows < ((size_t)1 << 31))
    return encode_sharded_tiled(ctx, msg_local, rows, log_k, (int)rho_inv, map, plain != 0);
  const Fr* dev;
  void* to_free;
  LG_TRY(stage_input(ctx, msg_local, rows * k, &dev, &to_free));
  int s = encode_rows(c, dev, rows, log_k, (int)rho_inv, nullptr, (Fr*)cosets_scratch, &map, plain != 0);
  if (to_free) {
    cudaStreamSynchronize(c->stream);
    cudaFree(to_free);
  }
  return s;
}

int lg_encode_sharded_rows(lg_ctx* ctx, const uint64_t* msg_rows, size_t nrows, size_t row_base, size_t rows_total, size_t k,
                           uint32_t rho_inv, void* const* shard_u, int world, uint64_t* cosets_scratch, int plain) {
  if (!ctx || !shard_u) return ERR_INVALID;
  Ctx* c = &ctx->c;
  cudaSetDevice(c->device);
  if (world < 1 || world > kMaxRanks || (world & (world - 1))) return set_error(c, ERR_INVALID, "world must be a power of two <= 8");
  if (k < 2 || (k & (k - 1)) || k % world) return set_error(c, ERR_INVALID, "k must be a power of two divisible by world");
  if (row_base + nrows > rows_total) return set_error(c, ERR_INVALID, "row range out of bounds");
  if (nrows == 0) return OK;
  if (!msg_rows) return set_error(c, ERR_INVALID, "null input matrix");
  int log_k = 0, log_w = 0;
  while (((size_t)1 << log_k) < k) log_k++;
  while ((1 << log_w) < world) log_w++;
  if (log_k > 10 && !cosets_scratch) return set_error(c, ERR_INVALID, "rows longer than 1024 need the local coset scratch buffer");
  OutMap map{};
  for (int g = 0; g < world; g++) {
    if (!shard_u[g]) return set_error(c, ERR_INVALID, "null shard pointer");
    map.base[g] = (Fr*)shard_u[g];
  }
  map.log_kg = log_k - log_w;
  map.m = map.m_g = (uint32_t)nrows;  // one block of consecutive global rows: grow(i) = row_base + i
  map.i0 = (uint32_t)row_base;
  map.rows_total = rows_total;
  // host rows go through the staged upload (copy stream + two staging tiles, reused across calls) from 1 MiB on: the
  // pipeline steps of the multi-GPU committer are runs of a few tens of MiB, and a cudaMalloc / cudaFree per run would
  // serialise the device
  if (!is_device_ptr(msg_rows) && nrows * k * sizeof(Fr) >= ((size_t)1 << 20) && nrows < ((size_t)1 << 31))
    return encode_sharded_tiled(ctx, msg_rows, nrows, log_k, (int)rho_inv, map, plain != 0);
  const Fr* dev;
  void* to_free;
  LG_TRY(stage_input(ctx, msg_rows, nrows * k, &dev, &to_free));
  int s = encode_rows(c, dev, nrows, log_k, (int)rho_inv, nullptr, (Fr*)cosets_scratch, &map, plain != 0);
  if (to_free) {
    cudaStreamSynchronize(c->stream);
    cudaFree(to_free);
  }
  return s;
}

int lg_matrix_hash_rows(lg_matrix* h, size_t row0, size_t row_end) {
  if (!h) return ERR_INVALID;
  Matrix& m = h->m;
  Ctx* c = m.ctx;
  cudaSetDevice(c->device);
  LG_TRY(hash_pipeline_setup(c, m.n));
  // everything enqueued on the context stream so far (the rows of this tile) happens before the tile is hashed
  LG_CUDA(c, cudaEventRecord(c->ev_encoded, c->stream));
  LG_CUDA(c, cudaStreamWaitEvent(c->hash_stream, c->ev_encoded, 0));
  return hash_columns_range(c, c->hash_stream, m.u, m.rows, m.log_k, m.rho_inv, row0, row_end, c->hash_state, m.leaves,
                            h->owner->col_len_prefix);
}

int lg_matrix_hash_finish(lg_matrix* h, uint8_t root_out[32]) {
  if (!h) return ERR_INVALID;
  Matrix& m = h->m;
  Ctx* c = m.ctx;
  cudaSetDevice(c->device);
  LG_TRY(hash_pipeline_setup(c, m.n));
  LG_TRY(merkle_build(c, m.leaves, m.n, m.nodes, h->owner->leaf_len_prefix, c->hash_stream));
  LG_CUDA(c, cudaEventRecord(c->ev_hashed, c->hash_stream));
  LG_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_hashed, 0));
  if (root_out) {
    LG_CUDA(c, cudaMemcpyAsync(root_out, m.nodes, 32, cudaMemcpyDeviceToHost, c->stream));
    LG_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  return OK;
}

void* lg_matrix_u_dev(const lg_matrix* m) { return m ? (void*)m->m.u : nullptr; }
void* lg_matrix_leaves_dev(const lg_matrix* m) { return m ? (void*)m->m.leaves : nullptr; }
void* lg_matrix_nodes_dev(const lg_matrix* m) { return m ? (void*)m->m.nodes : nullptr; }

int lg_matrix_read_rows(const lg_matrix* h, size_t row0, size_t nrows, uint64_t* out) {
  if (!h || !out) return ERR_INVALID;
  const Matrix& m = h->m;
  Ctx* c = m.ctx;
  if (row0 + nrows > m.rows) return set_error(c, ERR_INVALID, "row range out of bounds");
  if (nrows == 0) return OK;
  cudaSetDevice(c->device);
  Fr* tmp;
  LG_CUDA(c, cudaMalloc(&tmp, nrows * m.n * sizeof(Fr)));
  gather_rows_kernel<<<1024, 256, 0, c->stream>>>(m.u, m.rows, m.log_k, m.rho_inv, row0, nrows, tmp);
  c->launches++;
  cudaError_t e = cudaMemcpyAsync(out, tmp, nrows * m.n * sizeof(Fr), cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  cudaFree(tmp);
  if (e != cudaSuccess) return set_error(c, ERR_CUDA, cudaGetErrorString(e));
  return OK;
}

int lg_matrix_read_leaves(const lg_matrix* h, uint8_t* out) {
  if (!h || !out) return ERR_INVALID;
  Ctx* c = h->m.ctx;
  cudaSetDevice(c->device);
  LG_CUDA(c, cudaMemcpyAsync(out, h->m.leaves, h->m.n * 32, cudaMemcpyDeviceToHost, c->stream));
  LG_CUDA(c, cudaStreamSynchronize(c->stream));
  return OK;
}

int lg_matrix_read_nodes(const lg_matrix* h, uint8_t* out) {
  if (!h || !out) return ERR_INVALID;
  Ctx* c = h->m.ctx;
  cudaSetDevice(c->device);
  LG_CUDA(c, cudaMemcpyAsync(out, h->m.nodes, (h->m.n - 1) * 32, cudaMemcpyDeviceToHost, c->stream));
  LG_CUDA(c, cudaStreamSynchronize(c->stream));
  return OK;
}

int lg_intt(lg_ctx* ctx, const uint64_t* in, uint64_t* out, size_t rows, size_t size) {
  if (!ctx || !in || !out) return ERR_INVALID;
  Ctx* c = &ctx->c;
  if (size < 2 || (size & (size - 1))) return set_error(c, ERR_INVALID, "size must be a power of two >= 2");
  cudaSetDevice(c->device);
  int log_k = 0;
  while (((size_t)1 << log_k) < size) log_k++;
  const size_t elems = rows * size;
  const bool in_dev = is_device_ptr(in), out_dev = is_device_ptr(out);
  Fr *din = (Fr*)in, *dout = (Fr*)out;
  Fr* stage = nullptr;
  if (!in_dev || !out_dev) {
    LG_CUDA(c, cudaMalloc(&stage, elems * sizeof(Fr)));
    if (!in_dev) {
      LG_CUDA(c, cudaMemcpyAsync(stage, in, elems * sizeof(Fr), cudaMemcpyHostToDevice, c->stream));
      din = stage;
    }
    if (!out_dev) dout = stage;
  }
  int s = intt_rows(c, din, dout, rows, log_k);
  if (s == OK && !out_dev) {
    cudaError_t e = cudaMemcpyAsync(out, dout, elems * sizeof(Fr), cudaMemcpyDeviceToHost, c->stream);
    if (e != cudaSuccess) s = set_error(c, ERR_CUDA, cudaGetErrorString(e));
  }
  cudaError_t e2 = cudaStreamSynchronize(c->stream);
  if (s == OK && e2 != cudaSuccess) s = set_error(c, ERR_CUDA, cudaGetErrorString(e2));
  if (stage) cudaFree(stage);
  return s;
}

int lg_bench_shoup_peak(lg_ctx* ctx, double* shoup_mul_per_s, double* butterfly_per_s) {
  if (!ctx) return ERR_INVALID;
  if (shoup_mul_per_s) *shoup_mul_per_s = ctx->c.last_shoup_peak[0];
  if (butterfly_per_s) *butterfly_per_s = ctx->c.last_shoup_peak[1];
  return OK;
}

int lg_bench_int_peak(lg_ctx* ctx, double ms_target, double* fr_mul_per_s, double* imad_wide_per_s) {
  if (!ctx) return ERR_INVALID;
  Ctx* c = &ctx->c;
  cudaSetDevice(c->device);
  const int blocks = c->sm_count * 8, threads = 256;  // 2048 threads / SM
  const size_t nthreads = (size_t)blocks * threads;
  Fr* buf;
  LG_CUDA(c, cudaMalloc(&buf, 2 * nthreads * sizeof(Fr)));
  LG_CUDA(c, cudaMemsetAsync(buf, 0x17, 2 * nthreads * sizeof(Fr), c->stream));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float ms = 0;
  // Fr multiplications
  int iters = 64;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0, c->stream);
    bench_fr_mul_kernel<<<blocks, threads, 0, c->stream>>>(buf, iters);
    cudaEventRecord(e1, c->stream);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    c->launches++;
    if (rep == 0 && ms > 0) iters = (int)(iters * ms_target / ms) + 1;
  }
  if (fr_mul_per_s) *fr_mul_per_s = 4.0 * iters * (double)nthreads / (ms * 1e-3);
  // table-constant (Shoup) products and whole lazy butterflies: what the encoder actually executes
  for (int which = 0; which < 2; which++) {
    iters = 64;
    for (int rep = 0; rep < 3; rep++) {
      cudaEventRecord(e0, c->stream);
      if (which) bench_fr_shoup_kernel<true><<<blocks, threads, 0, c->stream>>>(buf, iters);
      else bench_fr_shoup_kernel<false><<<blocks, threads, 0, c->stream>>>(buf, iters);
      cudaEventRecord(e1, c->stream);
      cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
      c->launches++;
      if (rep == 0 && ms > 0) iters = (int)(iters * ms_target / ms) + 1;
    }
    c->last_shoup_peak[which] = 4.0 * iters * (double)nthreads / (ms * 1e-3);
  }
  // raw IMAD.WIDE.U32(.X) in the multiplier's carry-chained form
  iters = 64;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0, c->stream);
    bench_imad_wide_kernel<<<blocks, threads, 0, c->stream>>>((uint64_t*)buf, iters);
    cudaEventRecord(e1, c->stream);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    c->launches++;
    if (rep == 0 && ms > 0) iters = (int)(iters * ms_target / ms) + 1;
  }
  if (imad_wide_per_s) *imad_wide_per_s = 64.0 * iters * (double)nthreads / (ms * 1e-3);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaError_t e = cudaGetLastError();
  cudaFree(buf);
  if (e != cudaSuccess) return set_error(c, ERR_CUDA, cudaGetErrorString(e));
  return OK;
}

}  // extern "C"
