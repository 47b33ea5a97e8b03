// Internal declarations shared by the kernels and the C ABI (not installed; see include/ligero_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <new>
#include <stdexcept>
#include <mutex>
#include <set>
#include <string>
#include <vector>

#include "fr.cuh"
#include "fr_lazy.cuh"

namespace lg {

// ---- error plumbing: the C ABI never unwinds; every internal call returns a status ---------------
enum Status : int {
  OK = 0,
  ERR_INVALID = 1,      // bad argument (shape, null pointer, unsupported size)
  ERR_CUDA = 2,         // CUDA runtime error (see lg_last_error)
  ERR_NOMEM = 3,
  ERR_UNSUPPORTED = 4,
  ERR_STATE = 5,
};

struct Ctx;
int set_error(Ctx* ctx, int code, const std::string& msg);
// nothing unwinds across the C ABI: allocation failures of the host containers become LG_ERR_NOMEM
template <class F>
inline int guard(F&& body) {
  try {
    return body();
  } catch (const std::bad_alloc&) {
    return ERR_NOMEM;
  } catch (const std::exception&) {
    return ERR_STATE;
  }
}
#define LG_CUDA(ctx, expr)                                                                         \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return lg::set_error((ctx), lg::ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)
#define LG_TRY(expr)            \
  do {                          \
    int _s = (expr);            \
    if (_s != lg::OK) return _s; \
  } while (0)

// ---- NTT tables for one (log_k, rho_inv) ------------------------------------------------------------
struct NttTables {
  int log_k = 0;
  int rho_inv = 0;
  // {plain integer, quotient multiplier} entries (fr_lazy.cuh); one allocation, w_fwd is its base
  FrTw* w_fwd = nullptr;   // omega_k^i,  i < max(1, k/2)
  FrTw* w_inv = nullptr;   // omega_k^-i
  FrTw* scale = nullptr;   // (rho_inv-1) tables of k: scale[(s-1)*k + pos] = g^(s*bitrev(pos)) / k,  g = omega_{rho_inv*k}
  // the same scale factors for the persistent encoder (log_k >= 10): per coset and per 1024-element chunk four planes of
  // 1024 16-byte pieces {w.lo, w.hi, p.lo, p.hi}; the entry of chunk position 4t+e sits in slot e*256+t, so the 256 threads
  // of a group read consecutive slots (own allocation)
  uint4* scale4 = nullptr;
  FrTw kinv;               // 1/k
};

struct Ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  int sm_count = 148;
  std::string last_error;
  std::map<std::pair<int, int>, NttTables> tables;  // (log_k, rho_inv | plain << 16)
  // kernels whose dynamic shared-memory limit has been raised on THIS device (the attribute is per device)
  std::set<const void*> smem_configured;
  // kernel launch counter (bench.py's "gpu_launches")
  uint64_t launches = 0;
  // scratch reused across calls
  void* scratch = nullptr;
  size_t scratch_bytes = 0;
  // host-input pipeline: a copy stream, two staging tiles and the coset intermediate of one tile, so the
  // H2D upload of row tile i+1 overlaps the encoding of tile i
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_consumed[2] = {nullptr, nullptr};
  void* stage[2] = {nullptr, nullptr};
  size_t stage_bytes = 0;
  uint64_t stage_seq = 0;  // staging tiles alternate across calls of the sharded encoder (capi.cu)
  void* tile_cosets = nullptr;
  size_t tile_cosets_bytes = 0;
  // commit pipeline: column hashing of row tile i (high-priority stream) overlaps the encoding of tile i+1
  double last_shoup_peak[2] = {0, 0};  // lg_bench_int_peak: table-constant products/s, lazy butterflies/s
  bool overlap = true;
  // groups of 256 threads per SM in the persistent encoder (ntt.cu): 3 = every register of the SM, 2 = capped at 80
  // registers so that four CTAs of a co-resident column-hash kernel fit beside it (multi-GPU block pipeline);
  // 0 = the one-CTA-per-chunk kernel.  LG_NTT_PERSIST=0 / LG_NTT_GROUPS=2 in the environment set the default.
  int persist_groups = 3;
  // multi-GPU block pipeline (capi_shard.cu): the column hash of the row blocks that have landed runs on this stream,
  // one priority step ABOVE the context stream -- a column's BLAKE2s chain is latency bound, so its few warps must be
  // resident and scheduled first; the encoder fills the rest of the SM
  cudaStream_t hash_stream_hi = nullptr;
  // when set, encode_rows records this event right after launching the shared-memory kernel, i.e. before the last
  // strided pass (the NVLink-bound scatter of a sharded encode): the multi-GPU pipeline starts a block's column hash there
  cudaEvent_t ev_after_local = nullptr;
  cudaStream_t hash_stream = nullptr;
  cudaEvent_t ev_encoded = nullptr, ev_hashed = nullptr;
  uint32_t* hash_state = nullptr;
  size_t hash_state_words = 0;
  // column hashes (whole or row tiles) of at most this many columns use the four-lanes-per-column kernel (hash.cu);
  // lg_ctx_set_hash_quad_max / LG_HASH_QUAD_MAX override, 0 disables
  size_t hash_quad_max = 8192;
  // the prover's opened columns go device -> pinned proof buffer on this stream, behind the next test (capi_host.cu)
  cudaStream_t open_stream = nullptr;
  cudaEvent_t ev_gathered = nullptr;
  // pinned host staging for large device-to-host results (opened columns), grown on demand
  void* host_stage = nullptr;
  size_t host_stage_bytes = 0;
  // optional per-phase device timing (CUDA events on `stream`): bench.py's live kernel durations
  bool timing = false;
  std::vector<std::pair<int, cudaEvent_t>> marks;  // (phase that ENDS at this event, event)
  std::vector<cudaEvent_t> event_pool;
};

// phases reported by lg_ctx_phase_ms
enum Phase : int { PH_BEGIN = -1, PH_NTT_STRIDED_INV = 0, PH_NTT_LOCAL = 1, PH_NTT_STRIDED_FWD = 2, PH_HASH = 3, PH_MERKLE = 4,
                   PH_EXPAND = 5, PH_TESTS = 6, PH_OPEN = 7, PH_COUNT = 8 };
void phase_mark(Ctx* ctx, int phase_ended);

int ctx_scratch(Ctx* ctx, size_t bytes, void** out);
// second stream + events + per-column BLAKE2s state for tile-wise column hashing of an n-column matrix (capi.cu)
int hash_pipeline_setup(Ctx* c, size_t n);
int ctx_host_stage(Ctx* ctx, size_t bytes, void** out);  // pinned, reused across calls (one host thread per context)
// plain: the coset scale factors carry an extra R^-1, so the coset planes come out as plain integers
int get_tables(Ctx* ctx, int log_k, int rho_inv, const NttTables** out, bool plain = false);

// ---- committed matrix ------------------------------------------------------------------------------
// Physical layout of U (R x n, n = rho_inv*k) in HBM: rho_inv "coset planes", each R x k row-major:
//     plane[s][i][c] = U[i][rho_inv*c + s] = p_i(g^s * omega_k^c)
// plane 0 is the message itself (systematic positions, src/ligero/mod.rs:89).
// All planes hold the PLAIN integers (a, not the caller's Montgomery form a*R): the column hash needs exactly
// those bytes.  The encoder produces the coset planes that way for free by folding R^-1 into the coset scale
// table (the transform is linear) and converts plane 0 while copying it, so the hash kernel -- a latency-bound
// chain per column -- carries no Montgomery reduction at all.  Everything that reads U back (openings, row
// read-back, the test reductions) converts on the fly.
struct Matrix {
  Ctx* ctx = nullptr;
  size_t rows = 0;      // R
  int log_k = 0;
  int rho_inv = 0;
  size_t k = 0, n = 0;
  Fr* u = nullptr;          // rho_inv * R * k
  uint8_t* leaves = nullptr;  // n * 32, logical column order
  uint8_t* nodes = nullptr;   // (n-1) * 32, heap order, node 0 = root
  bool owns_u = true;
};

// ---- where a codeword element goes -------------------------------------------------------------------
// Single GPU: the local plane layout.  Multi GPU: the encode kernels store every finished element
// straight into the column shard of the rank that owns its message index c (peer memory over NVLink,
// mapped with CUDA IPC), at its GLOBAL row position -- the all-to-all is fused into the last NTT pass.
//   element (plane s, local row i, column c)  ->  base[c >> log_kg] + ((s*rows_total + grow(i)) << log_kg) + (c mod kg)
//   grow(i) = (i / m_g) * m + i0 + i % m_g     (local rows are [X_g; Y_g; Z_g; W_g], m_g rows per block;
//                                              i = kernel row + row_off when a row tile of them is encoded)
constexpr int kMaxRanks = 8;
struct OutMap {
  Fr* base[kMaxRanks];
  int log_kg;
  uint32_t m, m_g, i0;
  uint32_t row_off;  // the kernel's row 0 is local row row_off (a row tile of the local matrix)
  unsigned long long rows_total;
};
#if defined(__CUDACC__)
__device__ __forceinline__ uint32_t outmap_row(const OutMap& o, uint32_t i_local) {
  const uint32_t il = i_local + o.row_off;
  const uint32_t b = il / o.m_g;
  return b * o.m + o.i0 + (il - b * o.m_g);
}
__device__ __forceinline__ Fr* outmap_ptr(const OutMap& o, uint32_t s, uint32_t grow, uint32_t c) {
  return o.base[c >> o.log_kg] + ((((unsigned long long)s * o.rows_total + grow) << o.log_kg) + (c & ((1u << o.log_kg) - 1u)));
}
#endif

// ---- kernel launchers (all stream-ordered on ctx->stream) -------------------------------------------
// Reed-Solomon row encoding: msg (R x k, Montgomery, row-major, device) -> planes (a2+a3)
// plane0 (nullable) receives a copy of the message; cosets receives the rho_inv-1 planes s = 1..rho_inv-1
// map (nullable): final destination of every element (multi-GPU); `cosets` is then only the local
// intermediate of rows longer than one CTA tile and plane0 is ignored
// plain_cosets: all planes (incl. the plane-0 copy) as plain integers (the committed matrix, see Matrix)
int encode_rows(Ctx* ctx, const Fr* msg, size_t rows, int log_k, int rho_inv, Fr* plane0, Fr* cosets,
                const OutMap* map = nullptr, bool plain_cosets = false);
// protocol.cu
int expand_fr(Ctx* ctx, const uint8_t seed[32], size_t count, Fr* out_dev);
// x_plain: X (and Y, Z) hold plain integers (a coset plane s >= 1 of a committed matrix); the result is Montgomery
int col_reduce(Ctx* ctx, int mode, const Fr* W, const Fr* X, const Fr* Y, const Fr* Z, size_t rows, size_t k, Fr* out,
               size_t out_stride, size_t out_offset, bool x_plain = false);
int spmv_right_block(Ctx* ctx, const uint32_t* col_ptr, const uint32_t* row_idx, const uint32_t* val_id, const Fr* consts,
                     const Fr* r, size_t mk, Fr* out);
struct Matrix;
int gather_open(Ctx* ctx, const Matrix& m, const uint64_t* idx_dev, size_t t, Fr* cols_dev, uint8_t* sib_dev, uint8_t* auth_dev);
// verifier's per-column checks on t opened columns (t x rows contiguous, Montgomery); mode 0/1/2 = interleaved / linear /
// quadratic (protocol.cu); out: t elements
int column_checks(Ctx* ctx, int mode, const Fr* cols, const Fr* w, size_t rows, size_t t, Fr* out);
// batched inverse NTT of `rows` rows of length 2^log_k, natural order in and out, includes 1/k
int intt_rows(Ctx* ctx, const Fr* in, Fr* out, size_t rows, int log_k);
// column hashing (a4+a5): leaves[j] = BLAKE2s(u64le(R) || canonical LE bytes of column j)
int hash_columns(Ctx* ctx, const Fr* u_planes, size_t rows, int log_k, int rho_inv, uint8_t* leaves,
                 bool len_prefix);
// Merkle tree (a6)
int merkle_build(Ctx* ctx, const uint8_t* leaves, size_t n, uint8_t* nodes, bool leaf_len_prefix,
                 cudaStream_t st = nullptr);
// the same column hash over the row tile [row0, row_end) only, carrying the BLAKE2s state of every column in
// `state` (hash_state_words(n) words) between tiles; tiles come in row order and the one that ends at `rows`
// writes the leaves
size_t hash_state_words(size_t n);
int hash_columns_range(Ctx* ctx, cudaStream_t st, const Fr* u_planes, size_t rows, int log_k, int rho_inv, size_t row0,
                       size_t row_end, uint32_t* state, uint8_t* leaves, bool len_prefix);
// BLAKE2s of `count` explicit columns (each `rows` contiguous Montgomery elements)
int hash_column_list(Ctx* ctx, const Fr* cols, size_t rows, size_t count, uint8_t* digests, bool len_prefix);

// ---- evaluation trace + witness layout on the device (trace.cu; a1 of SURVEY 8a) --------------------------
// Built once per circuit by lg_ligero_new.  Gates are sorted by level (Add before Mul inside a level); the high bit
// of gate_node marks a Mul.  `pos` = the node's slot in each of the X/Y/Z/W blocks (its index once the constants
// other than node 0 are dropped, src/ligero/mod.rs:483-504).
struct TraceSegment {
  bool narrow;      // true: levels [lv0, lv1) walked by one CTA; false: one level = gates [g0, g1), thread per gate
  size_t lv0, lv1, g0, g1;
};
struct TraceSchedule {
  size_t n_nodes = 0, n_gates = 0, n_consts = 0, n_levels = 0, mk = 0;
  uint32_t *gate_node = nullptr, *gate_l = nullptr, *gate_r = nullptr, *gate_pos = nullptr;  // n_gates each
  uint32_t* level_start = nullptr;                                                           // n_levels + 1
  uint32_t *const_node = nullptr, *const_pos = nullptr;                                      // n_consts
  Fr* const_val = nullptr;
  Fr* vals = nullptr;                                                                        // n_nodes (value table)
  std::vector<TraceSegment> segments;
};
// out = [X;Y;Z;W] (4*mk elements, device); the variables come as device arrays (node, slot, value)
int trace_run(Ctx* ctx, const TraceSchedule& t, const uint32_t* var_node, const uint32_t* var_pos, const Fr* var_val,
              size_t n_vars, Fr* out);
void trace_free(TraceSchedule& t);

}  // namespace lg
