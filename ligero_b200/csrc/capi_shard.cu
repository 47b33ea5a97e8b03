// Multi-GPU prover behind the C ABI (include/ligero_b200.h, "lg_shard_*" / "lg_mgpu_*"): one lg_shard per GPU.
//
// Sharding (reference: the single call LigeroCircuit::prove, src/ligero/mod.rs:435-578):
//   rows    : every rank Reed-Solomon-encodes a share of the rows of [X;Y;Z;W];
//   exchange: each finished codeword element is stored from inside the encode kernels straight into the column shard
//             of the rank that owns its message index, over NVLink peer memory (ntt.cu, OutMap);
//   columns : rank h hashes the columns [h n/G, (h+1) n/G), builds their Merkle subtree and evaluates the three tests
//             on them (column-parallel: nothing to reduce), and serves the openings of its columns.
// Everything the ranks exchange besides U itself (subtree roots, test evaluations, authentication paths) travels
// through a per-rank MAILBOX in device memory that every peer maps: a sender stores its payload into each peer's
// mailbox and then raises a flag there (release, system scope); a one-warp wait kernel on the receiver's stream spins
// on its own flags (acquire).  All ranks run the same sequence of such collectives, numbered by an epoch counter, so
// no NCCL call sits on the data path and the same code serves one process per GPU (CUDA IPC mappings, handles passed
// by the host's launcher: lg_shard_handles / lg_shard_connect) and one process driving all GPUs (lg_mgpu_*: direct
// peer access, one host thread per GPU).
//
// Commit pipeline: a rank's rows are encoded in `steps` = 4 * sub consecutive-row runs, the j-th run of every rank
// belonging to the same block of global rows.  After run j each rank raises flag j everywhere; the column owner waits
// for all G flags ON ITS HASH STREAM (one priority step above the encoder's) and hashes that row block while the next
// run is being encoded and delivered.  A column hash is one sequential BLAKE2s chain per column, ~7 ms for a rank of 8
// GPUs at 2^24 gates however few columns it owns: behind the encoder it costs only its tail.
#include <array>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include "capi_types.h"
#include "fr_host.h"
#include "host_field.h"

using namespace lg;

// one consecutive run of global rows encoded by this rank in one pipeline step
struct ShardRun {
  size_t row_base, nrows;      // global rows [row_base, row_base + nrows) belong to this rank
  size_t blk0, blk1;           // the step's whole row block [blk0, blk1) (all ranks together)
  size_t local_off;            // first local row of the run inside the rank's local matrix
};

namespace {
typedef ShardRun Run;

constexpr size_t kMailFlags = 0;      // unsigned long long flag[kMaxRanks]
constexpr size_t kMailErr = 64;       // unsigned long long: epoch of a wait that timed out
constexpr size_t kMailRoots = 128;    // [2][kMaxRanks][32]
constexpr size_t kMailHeader = 1024;

struct Peers {
  uint8_t* mail[kMaxRanks];
};

// payload -> the same offset of every peer's mailbox, then flag `epoch` there.  One CTA.
__global__ void __launch_bounds__(1024) shard_post_kernel(Peers peers, int world, int me, size_t dst_off, const uint4* __restrict__ src,
                                                          size_t n16, unsigned long long epoch) {
  for (int h = 0; h < world; h++) {
    uint4* dst = reinterpret_cast<uint4*>(peers.mail[h] + dst_off);
    for (size_t i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = src[i];
  }
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < world) {
    // max, not store: with two encode streams a later run's flag may be raised before an earlier one's
    atomicMax_system(reinterpret_cast<unsigned long long*>(peers.mail[threadIdx.x] + kMailFlags) + me, epoch);
  }
}

// authentication paths of the opened columns this rank owns -> slot q of every peer's path area, then flag.
// entry layout per opened column: (depth_local + 1) digests: leaf sibling, then the path inside the subtree, root side first
__global__ void __launch_bounds__(256) shard_post_paths_kernel(Peers peers, int world, int me, size_t dst_off, const uint8_t* __restrict__ leaves,
                                                               const uint8_t* __restrict__ nodes, size_t n_local, int log_n_local,
                                                               const uint32_t* __restrict__ slot, const uint32_t* __restrict__ leaf,
                                                               uint32_t count, unsigned long long epoch) {
  const int per = log_n_local;  // 1 sibling + (log_n_local - 1) path digests
  for (uint32_t w = threadIdx.x; w < count * (uint32_t)per; w += blockDim.x) {
    const uint32_t e = w / per, d = w % per;
    const size_t j = leaf[e];
    const uint8_t* src;
    if (d == 0) {
      src = leaves + 32 * (j ^ 1);
    } else {
      // path digest d-1 (root side first) = sibling of the ancestor at depth d of the subtree
      size_t cur = n_local / 2 - 1 + j / 2;  // bottom-level inner node above the leaf, depth log_n_local - 1
      for (int up = log_n_local - 1; up > d; up--) cur = (cur - 1) / 2;
      src = nodes + 32 * ((cur & 1) ? cur + 1 : cur - 1);
    }
    const uint4 a = reinterpret_cast<const uint4*>(src)[0], b = reinterpret_cast<const uint4*>(src)[1];
    for (int h = 0; h < world; h++) {
      uint4* dst = reinterpret_cast<uint4*>(peers.mail[h] + dst_off + 32 * ((size_t)slot[e] * per + d));
      dst[0] = a;
      dst[1] = b;
    }
  }
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < world) {
    // max, not store: with two encode streams a later run's flag may be raised before an earlier one's
    atomicMax_system(reinterpret_cast<unsigned long long*>(peers.mail[threadIdx.x] + kMailFlags) + me, epoch);
  }
}

// spin until every rank has raised `epoch` in this rank's mailbox (bounded: a peer that died must not hang the GPU)
__global__ void shard_wait_kernel(uint8_t* mail, int world, unsigned long long epoch, long long max_cycles) {
  if ((int)threadIdx.x < world) {
    volatile unsigned long long* f = reinterpret_cast<volatile unsigned long long*>(mail + kMailFlags);
    const long long t0 = clock64();
    while (f[threadIdx.x] < epoch) {
      if (clock64() - t0 > max_cycles) {
        *reinterpret_cast<volatile unsigned long long*>(mail + kMailErr) = epoch;
        break;
      }
      __nanosleep(64);
    }
  }
  __threadfence_system();
}

// opened columns straight from their owners' shards (peer loads over NVLink): out[q][i] = U[i][idx[q]], Montgomery form
struct ShardBases {
  const Fr* u[kMaxRanks];
};
__global__ void gather_columns_sharded_kernel(ShardBases b, size_t rows, size_t kg, int log_kg, int rho, const uint64_t* __restrict__ idx,
                                              size_t t, Fr* __restrict__ out) {
  const size_t tot = t * rows;
  for (size_t f = (size_t)blockIdx.x * blockDim.x + threadIdx.x; f < tot; f += (size_t)gridDim.x * blockDim.x) {
    const size_t q = f / rows, i = f % rows;
    const size_t j = idx[q], s = j % rho, c = j / rho;
    const uint4* p = reinterpret_cast<const uint4*>(b.u[c >> log_kg] + ((s * rows + i) * kg + (c & (kg - 1))));
    const uint4 lo = p[0], hi = p[1];
    Fr x;
    x.v[0] = lo.x; x.v[1] = lo.y; x.v[2] = lo.z; x.v[3] = lo.w;
    x.v[4] = hi.x; x.v[5] = hi.y; x.v[6] = hi.z; x.v[7] = hi.w;
    x = fr_mul(x, fr_r2());  // the planes hold plain integers (Matrix)
    uint4* o = reinterpret_cast<uint4*>(out + f);
    o[0] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
    o[1] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
  }
}

// rows of a 4m x k device matrix picked by a list of (source row, count) runs, concatenated
__global__ void gather_row_runs_kernel(const Fr* __restrict__ full, size_t k, const uint32_t* __restrict__ src_row, size_t nrows,
                                       Fr* __restrict__ out) {
  const size_t tot = nrows * k;
  for (size_t f = (size_t)blockIdx.x * blockDim.x + threadIdx.x; f < tot; f += (size_t)gridDim.x * blockDim.x) {
    const size_t i = f / k, c = f % k;
    const uint4* p = reinterpret_cast<const uint4*>(full + (size_t)src_row[i] * k + c);
    uint4* o = reinterpret_cast<uint4*>(out + f);
    o[0] = p[0];
    o[1] = p[1];
  }
}

}  // namespace

struct lg_shard {
  lg_ctx* ctx = nullptr;
  int rank = 0, world = 1, log_w = 0;
  size_t m = 0, k = 0, kg = 0, rows = 0, n_local = 0;
  uint32_t rho = 8;
  int log_k = 0, log_n_local = 0;
  size_t t_max = 0;
  int sub = 1;                 // sub-blocks per X/Y/Z/W block: the commit pipeline has 4*sub steps
  int pipeline = -1;           // block pipeline: 0 off, 1 eager, 2 deferred (see encode_runs); -1: by world size
  std::vector<ShardRun> runs;
  size_t rows_local = 0, max_run = 0;
  lg_matrix* cols = nullptr;   // this rank's column shard of U (rows x kg, rho planes)
  lg_matrix* rhat = nullptr;   // the rows of r_a on the 2k domain, same sharding (2 planes)
  uint8_t* mail = nullptr;
  size_t mail_bytes = 0, slice_off = 0, slice_cap = 0, path_off = 0, path_cap = 0;
  void* peer_u[kMaxRanks] = {};
  void* peer_rhat[kMaxRanks] = {};
  Peers peers = {};
  bool connected = false, via_ipc = false;
  unsigned long long epoch = 0;
  Fr* scratch = nullptr;       // coset intermediate of one run (k > 1024)
  size_t scratch_bytes = 0;
  // two-stream encoding (odd runs on alt_stream): the NVLink-bound last pass of run j overlaps the shared-memory kernel
  // of run j+1; needs its own coset intermediate and its own NTT temporary (Ctx::scratch is swapped per run)
  int two_stream = -1;         // -1: by world size
  cudaStream_t alt_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  Fr* scratch2 = nullptr;
  size_t scratch2_bytes = 0;
  void* ctx_scratch2 = nullptr;
  size_t ctx_scratch2_bytes = 0;
  cudaEvent_t ev_step = nullptr, ev_done = nullptr, ev_local[2] = {nullptr, nullptr};
  uint8_t* host_pin = nullptr; // pinned: header read-back (roots, err)
  uint32_t* local_src_rows = nullptr;  // device: global row of every local row (gather of r_a / the witness)
  Fr* pre_full = nullptr;      // 4mk, device trace output (lg_shard_prove)
  Fr* pre_local = nullptr;     // rows_local x k
  uint8_t subtree_roots[kMaxRanks][32] = {};
  double last_ms[4] = {0, 0, 0, 0};
};

namespace {

int sfail(lg_shard* s, int code, const std::string& msg) { return set_error(&s->ctx->c, code, msg); }

void block_slice(size_t len, int world, int rank, size_t* a, size_t* b) {  // first len % world ranks get one extra row
  const size_t base = len / world, extra = len % world;
  *a = rank * base + ((size_t)rank < extra ? rank : extra);
  *b = *a + base + ((size_t)rank < extra ? 1 : 0);
}

void build_runs(lg_shard* s) {
  s->runs.clear();
  size_t off = 0;
  s->max_run = 0;
  for (int b = 0; b < 4; b++)
    for (int u = 0; u < s->sub; u++) {
      size_t s0, s1, a, e;
      block_slice(s->m, s->sub, u, &s0, &s1);
      block_slice(s1 - s0, s->world, s->rank, &a, &e);
      Run r;
      r.blk0 = b * s->m + s0;
      r.blk1 = b * s->m + s1;
      r.row_base = r.blk0 + a;
      r.nrows = e - a;
      r.local_off = off;
      off += r.nrows;
      if (r.nrows > s->max_run) s->max_run = r.nrows;
      s->runs.push_back(r);
    }
  s->rows_local = off;
}

int post(lg_shard* s, cudaStream_t st, size_t dst_off, const void* src_dev, size_t bytes) {
  Ctx* c = &s->ctx->c;
  s->epoch++;
  shard_post_kernel<<<1, 1024, 0, st>>>(s->peers, s->world, s->rank, dst_off, (const uint4*)src_dev, bytes / 16, s->epoch);
  c->launches++;
  LG_CUDA(c, cudaGetLastError());
  return OK;
}
int wait_all(lg_shard* s, cudaStream_t st, unsigned long long epoch = 0) {
  Ctx* c = &s->ctx->c;
  shard_wait_kernel<<<1, 32, 0, st>>>(s->mail, s->world, epoch ? epoch : s->epoch, (long long)2e10);
  c->launches++;
  LG_CUDA(c, cudaGetLastError());
  return OK;
}
int check_err(lg_shard* s) {  // after a stream synchronisation
  unsigned long long e = 0;
  Ctx* c = &s->ctx->c;
  LG_CUDA(c, cudaMemcpy(&e, s->mail + kMailErr, 8, cudaMemcpyDeviceToHost));
  if (e) return sfail(s, ERR_STATE, "a peer rank did not reach collective " + std::to_string(e) + " in time");
  return OK;
}

int ensure_buf(lg_shard* s, Fr** buf, size_t* have, size_t bytes) {
  Ctx* c = &s->ctx->c;
  if (bytes <= *have) return OK;
  LG_CUDA(c, cudaStreamSynchronize(c->stream));
  if (s->alt_stream) LG_CUDA(c, cudaStreamSynchronize(s->alt_stream));
  if (*buf) cudaFree(*buf);
  *buf = nullptr;
  *have = 0;
  cudaError_t e = cudaMalloc(buf, bytes);
  if (e != cudaSuccess) return sfail(s, ERR_NOMEM, std::string("shard scratch cudaMalloc: ") + cudaGetErrorString(e));
  *have = bytes;
  return OK;
}
int ensure_scratch(lg_shard* s, size_t bytes) { return ensure_buf(s, &s->scratch, &s->scratch_bytes, bytes); }

// encode every run of `local` (host or device, the rank's rows in run order) into the peers' shards `dst`.
// hash != nullptr: the block pipeline -- a flag after every run, and the owner hashes block j on its hash stream
//   mode 1 (eager)   : as soon as block j has landed everywhere, i.e. beside the shared-memory kernel of run j+1, which
//                      then runs two groups per SM and leaves the hash a third of the registers;
//   mode 2 (deferred): when the shared-memory kernel of run j+1 is done, i.e. beside the last strided pass of run j+1,
//                      whose warps mostly wait for their NVLink stores -- the encoder keeps all three groups.
int encode_runs(lg_shard* s, const uint64_t* local, uint32_t rho, void* const* dst, bool plain, lg_matrix* hash, int mode) {
  Ctx* c = &s->ctx->c;
  if (s->log_k > 10) LG_TRY(ensure_scratch(s, (size_t)(rho - 1) * s->max_run * s->k * sizeof(Fr)));
  const int groups_saved = c->persist_groups;
  static int overlap_groups = -1;  // LG_SHARD_GROUPS=3 keeps the three-group encoder in the eager mode too
  if (overlap_groups < 0) {
    const char* e = getenv("LG_SHARD_GROUPS");
    overlap_groups = (e && atoi(e) == 3) ? 3 : 2;
  }
  auto hash_block = [&](const Run& r, unsigned long long epoch, cudaEvent_t after) -> int {
    LG_CUDA(c, cudaStreamWaitEvent(c->hash_stream_hi, after, 0));
    LG_TRY(wait_all(s, c->hash_stream_hi, epoch));
    return hash_columns_range(c, c->hash_stream_hi, hash->m.u, hash->m.rows, hash->m.log_k, hash->m.rho_inv, r.blk0, r.blk1,
                              c->hash_state, hash->m.leaves, s->ctx->col_len_prefix);
  };
  unsigned long long prev_epoch = 0;
  // two streams (device input, rows longer than one chunk, eager or no hashing): odd runs go to alt_stream
  const int ts_default = (s->world >= 8 && hash) ? 1 : 0;
  const bool two = (s->two_stream < 0 ? ts_default : s->two_stream) != 0 && mode != 2 && s->log_k > 10 && s->runs.size() > 1 &&
                   (!local || is_device_ptr(local));
  cudaStream_t main_stream = c->stream;
  void* ctx_scr = c->scratch;
  size_t ctx_scr_bytes = c->scratch_bytes;
  if (two) {
    if (!s->alt_stream) {
      int prio = 0;
      LG_CUDA(c, cudaStreamGetPriority(main_stream, &prio));
      LG_CUDA(c, cudaStreamCreateWithPriority(&s->alt_stream, cudaStreamNonBlocking, prio));
      LG_CUDA(c, cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming));
      LG_CUDA(c, cudaEventCreateWithFlags(&s->ev_join, cudaEventDisableTiming));
    }
    LG_TRY(ensure_buf(s, &s->scratch2, &s->scratch2_bytes, (size_t)(rho - 1) * s->max_run * s->k * sizeof(Fr)));
    LG_TRY(ensure_buf(s, (Fr**)&s->ctx_scratch2, &s->ctx_scratch2_bytes, s->max_run * s->k * sizeof(Fr)));
    LG_CUDA(c, cudaEventRecord(s->ev_fork, main_stream));
    LG_CUDA(c, cudaStreamWaitEvent(s->alt_stream, s->ev_fork, 0));
  }
  auto restore = [&]() {
    if (two) {  // hand the (possibly grown) NTT temporary back and rejoin the streams
      if (c->stream == s->alt_stream) {
        s->ctx_scratch2 = c->scratch;
        s->ctx_scratch2_bytes = c->scratch_bytes;
      } else {
        ctx_scr = c->scratch;
        ctx_scr_bytes = c->scratch_bytes;
      }
      c->stream = main_stream;
      c->scratch = ctx_scr;
      c->scratch_bytes = ctx_scr_bytes;
    }
  };
  for (size_t j = 0; j < s->runs.size(); j++) {
    const Run& r = s->runs[j];
    if (hash && mode == 1 && groups_saved == 3) c->persist_groups = (j == 0) ? 3 : overlap_groups;
    cudaEvent_t ev_loc = s->ev_local[j & 1];
    if (hash && mode == 2) c->ev_after_local = ev_loc;
    Fr* coset_scratch = s->scratch;
    if (two) {
      restore();
      if (j & 1) {
        c->stream = s->alt_stream;
        c->scratch = s->ctx_scratch2;
        c->scratch_bytes = s->ctx_scratch2_bytes;
        coset_scratch = s->scratch2;
      }
    }
    int st = OK;
    if (r.nrows)
      st = lg_encode_sharded_rows(s->ctx, local + r.local_off * s->k * 4, r.nrows, r.row_base, s->rows, s->k, rho, dst, s->world,
                                  (uint64_t*)coset_scratch, plain ? 1 : 0);
    else if (hash && mode == 2)
      st = cudaEventRecord(ev_loc, c->stream) == cudaSuccess ? OK : ERR_CUDA;
    c->persist_groups = groups_saved;
    c->ev_after_local = nullptr;
    if (st != OK) {
      restore();
      return st;
    }
    if (hash) {
      s->epoch++;
      shard_post_kernel<<<1, 32, 0, c->stream>>>(s->peers, s->world, s->rank, 0, nullptr, 0, s->epoch);
      c->launches++;
      LG_CUDA(c, cudaGetLastError());
      if (mode == 2) {
        if (j > 0) LG_TRY(hash_block(s->runs[j - 1], prev_epoch, ev_loc));
        prev_epoch = s->epoch;
      }
      if (mode == 1 || j + 1 == s->runs.size()) {
        LG_CUDA(c, cudaEventRecord(s->ev_step, c->stream));
        LG_TRY(hash_block(r, s->epoch, s->ev_step));
      }
    }
  }
  if (two) {
    restore();
    LG_CUDA(c, cudaEventRecord(s->ev_join, s->alt_stream));
    LG_CUDA(c, cudaStreamWaitEvent(main_stream, s->ev_join, 0));
  }
  return OK;
}

int commit_async(lg_shard* s, const uint64_t* local) {
  Ctx* c = &s->ctx->c;
  if (!s->connected) return sfail(s, ERR_STATE, "lg_shard_connect first");
  if (!local && s->rows_local) return sfail(s, ERR_INVALID, "null input matrix");
  Matrix& m = s->cols->m;
  // Measured on 8 x B200 (2^24-gate shape, ms per step; profiles/r2_mgpu_sweep.txt):
  //   2 GPUs: a rank owns 32 768 columns, its hash saturates the ALU pipe and sharing the SMs with it costs the encoder more
  //           than the overlap saves (77 pipelined vs 67 plain);   4 GPUs: 40.6 vs 38.7, still a loss;
  //   8 GPUs: the hash is a 7.5 ms latency chain over 8 192 columns; eager pipeline + two encode streams + 16 steps: 21.5
  //           vs 25.0 plain (eager alone 26.3, two streams alone 24.5, deferred 26.3).
  // Host input is the exception at every size: the upload paces the encoder anyway (lg_shard_set_pipeline(1): 79 vs 89 ms
  // end to end at 2 GPUs, 39 vs 46 at 8).
  const int pipe = s->pipeline < 0 ? (s->world >= 8 ? 1 : 0) : s->pipeline;
  if (pipe) {
    LG_TRY(hash_pipeline_setup(c, m.n));
    // the hash stream must not start on a new commitment before the previous one's consumers are done
    LG_CUDA(c, cudaEventRecord(s->ev_step, c->stream));
    LG_CUDA(c, cudaStreamWaitEvent(c->hash_stream_hi, s->ev_step, 0));
    LG_TRY(encode_runs(s, local, s->rho, s->peer_u, true, s->cols, pipe));
    LG_TRY(merkle_build(c, m.leaves, m.n, m.nodes, s->ctx->leaf_len_prefix, c->hash_stream_hi));
    LG_CUDA(c, cudaEventRecord(s->ev_done, c->hash_stream_hi));
    LG_CUDA(c, cudaStreamWaitEvent(c->stream, s->ev_done, 0));
  } else {
    LG_TRY(encode_runs(s, local, s->rho, s->peer_u, true, nullptr, 0));
    LG_TRY(post(s, c->stream, 0, nullptr, 0));  // every rank's rows have been stored ...
    LG_TRY(wait_all(s, c->stream));             // ... everywhere
    LG_TRY(hash_columns(c, m.u, m.rows, m.log_k, m.rho_inv, m.leaves, s->ctx->col_len_prefix));
    LG_TRY(merkle_build(c, m.leaves, m.n, m.nodes, s->ctx->leaf_len_prefix));
  }
  // subtree root (node 0) -> everyone; the flag doubles as "this rank is done reading its shard's previous contents"
  const size_t off = kMailRoots + ((s->epoch + 1) & 1) * kMaxRanks * 32 + (size_t)s->rank * 32;
  LG_TRY(post(s, c->stream, off, m.nodes, 32));
  LG_TRY(wait_all(s, c->stream));
  LG_CUDA(c, cudaMemcpyAsync(s->host_pin, s->mail + kMailRoots + (s->epoch & 1) * kMaxRanks * 32, kMaxRanks * 32, cudaMemcpyDeviceToHost,
                             c->stream));
  return OK;
}

void fold_roots(const lg_shard* s, const uint8_t* roots, uint8_t out[32]) {
  std::vector<std::array<uint8_t, 32>> level(s->world);
  for (int g = 0; g < s->world; g++) memcpy(level[g].data(), roots + 32 * g, 32);
  while (level.size() > 1) {
    std::vector<std::array<uint8_t, 32>> up(level.size() / 2);
    for (size_t i = 0; i < up.size(); i++) {
      uint8_t buf[64];
      memcpy(buf, level[2 * i].data(), 32);
      memcpy(buf + 32, level[2 * i + 1].data(), 32);
      lgh::sha256(buf, 64, up[i].data());
    }
    level.swap(up);
  }
  memcpy(out, level[0].data(), 32);
}

int finish_root(lg_shard* s, uint8_t root_out[32]) {
  Ctx* c = &s->ctx->c;
  LG_CUDA(c, cudaStreamSynchronize(c->stream));
  LG_TRY(check_err(s));
  memcpy(s->subtree_roots, s->host_pin, (size_t)s->world * 32);
  if (root_out) fold_roots(s, s->host_pin, root_out);
  return OK;
}

// all-gather of `count` Fr per rank (device) -> host, rank order
int gather_slices(lg_shard* s, const Fr* part_dev, size_t count, Fr* host_out) {
  Ctx* c = &s->ctx->c;
  if ((size_t)s->world * count * sizeof(Fr) > s->slice_cap) return sfail(s, ERR_INVALID, "slice larger than the mailbox");
  const size_t par = (s->epoch + 1) & 1;
  LG_TRY(post(s, c->stream, s->slice_off + par * s->slice_cap + (size_t)s->rank * count * sizeof(Fr), part_dev, count * sizeof(Fr)));
  LG_TRY(wait_all(s, c->stream));
  LG_CUDA(c, cudaMemcpyAsync(host_out, s->mail + s->slice_off + par * s->slice_cap, (size_t)s->world * count * sizeof(Fr),
                             cudaMemcpyDeviceToHost, c->stream));
  LG_CUDA(c, cudaStreamSynchronize(c->stream));
  return check_err(s);
}

struct OpenedHost {
  std::vector<uint64_t> idx;
  std::vector<uint64_t> cols;   // t * rows * 4 limbs
  std::vector<uint8_t> sib, auth;
};

// open_columns (src/ligero/mod.rs:935-955) over the sharded commitment; want_cols = false: take part in the
// collectives but do not fetch the columns (a rank whose proof nobody reads)
int open_sharded(lg_shard* s, size_t n, size_t t, lg_sponge* sponge, bool want_cols, OpenedHost& o) {
  Ctx* c = &s->ctx->c;
  uint8_t seed[32];
  LG_TRY(lg_sponge_squeeze_bytes(sponge, seed, 32));
  o.idx.resize(t);
  LG_TRY(lg_expand_indices(seed, n, t, o.idx.data()));
  if (t > s->t_max) return sfail(s, ERR_INVALID, "more openings than the shard was created for");
  const int per = s->log_n_local;
  int log_n = 0;
  while (((size_t)1 << log_n) < n) log_n++;
  const size_t depth = (size_t)log_n - 1, ntop = (size_t)s->log_w;
  // this rank's share of the opened columns: (slot q, leaf inside the subtree)
  std::vector<uint32_t> mine;
  for (size_t q = 0; q < t; q++)
    if ((int)(o.idx[q] / s->n_local) == s->rank) {
      mine.push_back((uint32_t)q);
      mine.push_back((uint32_t)(o.idx[q] % s->n_local));
    }
  const uint32_t cnt = (uint32_t)(mine.size() / 2);
  std::vector<uint32_t> packed(2 * (size_t)cnt + 2);
  for (uint32_t e = 0; e < cnt; e++) {
    packed[e] = mine[2 * e];
    packed[cnt + e] = mine[2 * e + 1];
  }
  uint32_t* d_list = nullptr;
  LG_CUDA(c, cudaMallocAsync((void**)&d_list, packed.size() * 4, c->stream));
  LG_CUDA(c, cudaMemcpyAsync(d_list, packed.data(), packed.size() * 4, cudaMemcpyHostToDevice, c->stream));
  s->epoch++;
  const size_t par = s->epoch & 1;
  const Matrix& m = s->cols->m;
  shard_post_paths_kernel<<<1, 256, 0, c->stream>>>(s->peers, s->world, s->rank, s->path_off + par * s->path_cap, m.leaves, m.nodes,
                                                    s->n_local, s->log_n_local, d_list, d_list + cnt, cnt, s->epoch);
  c->launches++;
  LG_CUDA(c, cudaGetLastError());
  LG_TRY(wait_all(s, c->stream));
  std::vector<uint8_t> entries(t * (size_t)per * 32);
  LG_CUDA(c, cudaMemcpyAsync(entries.data(), s->mail + s->path_off + par * s->path_cap, entries.size(), cudaMemcpyDeviceToHost, c->stream));
  uint64_t* d_idx = nullptr;
  Fr* d_cols = nullptr;
  void* stage = nullptr;
  if (want_cols) {
    LG_CUDA(c, cudaMallocAsync((void**)&d_idx, t * 8, c->stream));
    LG_CUDA(c, cudaMemcpyAsync(d_idx, o.idx.data(), t * 8, cudaMemcpyHostToDevice, c->stream));
    LG_CUDA(c, cudaMallocAsync((void**)&d_cols, t * s->rows * sizeof(Fr), c->stream));
    ShardBases b{};
    for (int g = 0; g < s->world; g++) b.u[g] = (const Fr*)s->peer_u[g];
    phase_mark(c, PH_BEGIN);
    gather_columns_sharded_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(b, s->rows, s->kg, s->log_k - s->log_w, (int)s->rho, d_idx, t,
                                                                          d_cols);
    c->launches++;
    LG_CUDA(c, cudaGetLastError());
    phase_mark(c, PH_OPEN);
    LG_TRY(ctx_host_stage(c, t * s->rows * sizeof(Fr), &stage));
    LG_CUDA(c, cudaMemcpyAsync(stage, d_cols, t * s->rows * sizeof(Fr), cudaMemcpyDeviceToHost, c->stream));
  }
  LG_CUDA(c, cudaStreamSynchronize(c->stream));
  cudaFreeAsync(d_list, c->stream);
  if (d_idx) cudaFreeAsync(d_idx, c->stream);
  if (d_cols) cudaFreeAsync(d_cols, c->stream);
  LG_TRY(check_err(s));
  if (want_cols) {
    o.cols.resize(t * s->rows * 4);
    memcpy(o.cols.data(), stage, t * s->rows * sizeof(Fr));
  }
  // top log2(G) levels of every path from the gathered subtree roots (root side first)
  std::vector<std::vector<std::array<uint8_t, 32>>> levels;
  levels.emplace_back(s->world);
  for (int g = 0; g < s->world; g++) memcpy(levels[0][g].data(), s->subtree_roots[g], 32);
  while (levels.back().size() > 1) {
    const auto& lv = levels.back();
    std::vector<std::array<uint8_t, 32>> up(lv.size() / 2);
    for (size_t i = 0; i < up.size(); i++) {
      uint8_t buf[64];
      memcpy(buf, lv[2 * i].data(), 32);
      memcpy(buf + 32, lv[2 * i + 1].data(), 32);
      lgh::sha256(buf, 64, up[i].data());
    }
    levels.push_back(std::move(up));
  }
  o.sib.assign(t * 32, 0);
  o.auth.assign(t * depth * 32 + 1, 0);
  for (size_t q = 0; q < t; q++) {
    const uint8_t* e = entries.data() + q * (size_t)per * 32;
    memcpy(o.sib.data() + 32 * q, e, 32);
    size_t pos = o.idx[q] / s->n_local;
    std::vector<const uint8_t*> top;
    for (size_t lv = 0; lv + 1 < levels.size(); lv++) {  // bottom (subtree roots) upwards
      top.push_back(levels[lv][pos ^ 1].data());
      pos >>= 1;
    }
    for (size_t d = 0; d < ntop; d++) memcpy(o.auth.data() + 32 * (q * depth + d), top[ntop - 1 - d], 32);
    for (int d = 1; d < per; d++) memcpy(o.auth.data() + 32 * (q * depth + ntop + (size_t)(d - 1)), e + 32 * (size_t)d, 32);
  }
  return OK;
}

}  // namespace

struct lg_mgpu {
  int world = 0;
  lg_ctx* ctx[kMaxRanks] = {};
  std::string error;
};
struct lg_mligero {
  lg_mgpu* mg = nullptr;
  lg_ligero* L[kMaxRanks] = {};
  lg_shard* sh[kMaxRanks] = {};
};

// run fn(rank) on one host thread per GPU; first failure wins
template <class F>
static int on_all(lg_mgpu* g, F fn) {
  int st[kMaxRanks] = {};
  std::vector<std::thread> th;
  for (int i = 0; i < g->world; i++) th.emplace_back([&, i]() { st[i] = fn(i); });
  for (auto& t : th) t.join();
  for (int i = 0; i < g->world; i++)
    if (st[i] != OK) {
      g->error = "rank " + std::to_string(i) + ": " + lg_last_error(g->ctx[i]);
      return st[i];
    }
  return OK;
}

extern "C" {

static int lg_shard_create_impl(lg_ctx* ctx, size_t m, size_t k, uint32_t rho_inv, int rank, int world, size_t t_max, int sub_blocks,
                    lg_shard** out) {
  if (!ctx || !out) return ERR_INVALID;
  Ctx* c = &ctx->c;
  cudaSetDevice(c->device);
  if (world < 1 || world > kMaxRanks || (world & (world - 1)) || rank < 0 || rank >= world)
    return set_error(c, ERR_INVALID, "world must be a power of two <= 8 and 0 <= rank < world");
  if (m == 0 || k < 2 || (k & (k - 1)) || k % world || rho_inv < 2 || (rho_inv & (rho_inv - 1)) || (rho_inv * k / world) < 2)
    return set_error(c, ERR_INVALID, "shape not shardable: k a power of two divisible by world, rho_inv a power of two >= 2");
  if (sub_blocks < 1) {  // automatic: the eager pipeline of 8 GPUs wants finer steps (measured 23.1 ms with 4 steps, 21.9 with 8, 21.5 with 16)
    sub_blocks = world >= 8 ? 4 : 1;
    if (const char* e = getenv("LG_SHARD_SUB"))
      if (atoi(e) >= 1) sub_blocks = atoi(e);
  }
  if ((size_t)sub_blocks > m) sub_blocks = (int)m;
  lg_shard* s = new (std::nothrow) lg_shard();
  if (!s) return ERR_NOMEM;
  s->ctx = ctx;
  s->rank = rank;
  s->world = world;
  while ((1 << s->log_w) < world) s->log_w++;
  s->m = m;
  s->k = k;
  s->kg = k / world;
  s->rho = rho_inv;
  s->rows = 4 * m;
  s->n_local = (size_t)rho_inv * s->kg;
  while (((size_t)1 << s->log_k) < k) s->log_k++;
  while (((size_t)1 << s->log_n_local) < s->n_local) s->log_n_local++;
  s->t_max = t_max;
  s->sub = sub_blocks;
  if (const char* e = getenv("LG_SHARD_TWO_STREAM")) s->two_stream = atoi(e) != 0 ? 1 : 0;
  if (const char* e = getenv("LG_SHARD_PIPELINE")) s->pipeline = atoi(e) < 0 ? -1 : (atoi(e) > 2 ? 2 : atoi(e));
  build_runs(s);
  int st = lg_matrix_create(ctx, s->rows, s->kg, rho_inv, &s->cols);
  if (st == OK && t_max) st = lg_matrix_create(ctx, s->rows, s->kg, 2, &s->rhat);
  s->slice_off = kMailHeader;
  s->slice_cap = 2 * k * sizeof(Fr);
  s->path_off = s->slice_off + 2 * s->slice_cap;
  s->path_cap = (t_max ? t_max : 1) * (size_t)s->log_n_local * 32;
  s->mail_bytes = s->path_off + 2 * s->path_cap;
  cudaError_t e = cudaSuccess;
  if (st == OK) {
    e = cudaMalloc((void**)&s->mail, s->mail_bytes);
    if (e == cudaSuccess) e = cudaMemset(s->mail, 0, s->mail_bytes);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_step, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_local[0], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_local[1], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&s->host_pin, kMailHeader, cudaHostAllocDefault);
    if (e == cudaSuccess && !c->hash_stream_hi) {
      int lo = 0, hi = 0;
      cudaDeviceGetStreamPriorityRange(&lo, &hi);
      e = cudaStreamCreateWithPriority(&c->hash_stream_hi, cudaStreamNonBlocking, hi);
    }
    if (e == cudaSuccess && s->rows_local) {
      std::vector<uint32_t> src(s->rows_local);
      for (const Run& r : s->runs)
        for (size_t i = 0; i < r.nrows; i++) src[r.local_off + i] = (uint32_t)(r.row_base + i);
      e = cudaMalloc((void**)&s->local_src_rows, src.size() * 4);
      if (e == cudaSuccess) e = cudaMemcpy(s->local_src_rows, src.data(), src.size() * 4, cudaMemcpyHostToDevice);
    }
  }
  if (st != OK || e != cudaSuccess) {
    if (e != cudaSuccess) st = set_error(c, ERR_CUDA, std::string("lg_shard_create: ") + cudaGetErrorString(e));
    lg_shard_free(s);
    return st;
  }
  *out = s;
  return OK;
}
int lg_shard_create(lg_ctx* ctx, size_t m, size_t k, uint32_t rho_inv, int rank, int world, size_t t_max, int sub_blocks,
                    lg_shard** out) {
  return lg::guard([&]() { return lg_shard_create_impl(ctx, m, k, rho_inv, rank, world, t_max, sub_blocks, out); });
}

int lg_shard_free(lg_shard* s) {
  if (!s) return OK;
  Ctx* c = &s->ctx->c;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  if (c->hash_stream_hi) cudaStreamSynchronize(c->hash_stream_hi);
  if (s->via_ipc)
    for (int g = 0; g < s->world; g++)
      if (g != s->rank) {
        if (s->peer_u[g]) cudaIpcCloseMemHandle(s->peer_u[g]);
        if (s->peer_rhat[g]) cudaIpcCloseMemHandle(s->peer_rhat[g]);
        if (s->peers.mail[g]) cudaIpcCloseMemHandle(s->peers.mail[g]);
      }
  if (s->cols) lg_matrix_free(s->cols);
  if (s->rhat) lg_matrix_free(s->rhat);
  if (s->mail) cudaFree(s->mail);
  if (s->scratch) cudaFree(s->scratch);
  if (s->scratch2) cudaFree(s->scratch2);
  if (s->ctx_scratch2) cudaFree(s->ctx_scratch2);
  if (s->alt_stream) {
    cudaStreamSynchronize(s->alt_stream);
    cudaStreamDestroy(s->alt_stream);
    cudaEventDestroy(s->ev_fork);
    cudaEventDestroy(s->ev_join);
  }
  if (s->local_src_rows) cudaFree(s->local_src_rows);
  if (s->pre_full) cudaFree(s->pre_full);
  if (s->pre_local) cudaFree(s->pre_local);
  if (s->host_pin) cudaFreeHost(s->host_pin);
  if (s->ev_step) cudaEventDestroy(s->ev_step);
  if (s->ev_done) cudaEventDestroy(s->ev_done);
  for (int i = 0; i < 2; i++)
    if (s->ev_local[i]) cudaEventDestroy(s->ev_local[i]);
  delete s;
  return OK;
}

int lg_shard_handles(lg_shard* s, uint8_t out[3 * 64]) {
  if (!s || !out) return ERR_INVALID;
  Ctx* c = &s->ctx->c;
  cudaSetDevice(c->device);
  cudaIpcMemHandle_t h;
  memset(out, 0, 3 * 64);
  LG_CUDA(c, cudaIpcGetMemHandle(&h, s->cols->m.u));
  memcpy(out, &h, 64);
  if (s->rhat) {
    LG_CUDA(c, cudaIpcGetMemHandle(&h, s->rhat->m.u));
    memcpy(out + 64, &h, 64);
  }
  LG_CUDA(c, cudaIpcGetMemHandle(&h, s->mail));
  memcpy(out + 128, &h, 64);
  return OK;
}

int lg_shard_connect(lg_shard* s, const uint8_t* all_handles) {
  if (!s || !all_handles) return ERR_INVALID;
  Ctx* c = &s->ctx->c;
  cudaSetDevice(c->device);
  if (s->connected) return sfail(s, ERR_STATE, "already connected");
  for (int g = 0; g < s->world; g++) {
    if (g == s->rank) {
      s->peer_u[g] = s->cols->m.u;
      s->peer_rhat[g] = s->rhat ? s->rhat->m.u : nullptr;
      s->peers.mail[g] = s->mail;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, all_handles + (size_t)g * 192, 64);
    LG_CUDA(c, cudaIpcOpenMemHandle(&s->peer_u[g], h, cudaIpcMemLazyEnablePeerAccess));
    if (s->rhat) {
      memcpy(&h, all_handles + (size_t)g * 192 + 64, 64);
      LG_CUDA(c, cudaIpcOpenMemHandle(&s->peer_rhat[g], h, cudaIpcMemLazyEnablePeerAccess));
    }
    memcpy(&h, all_handles + (size_t)g * 192 + 128, 64);
    void* p = nullptr;
    LG_CUDA(c, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    s->peers.mail[g] = (uint8_t*)p;
  }
  s->via_ipc = true;
  s->connected = true;
  return OK;
}

int lg_shard_connect_local(lg_shard* const* shards, int world) {
  if (!shards || world < 1 || world > kMaxRanks) return ERR_INVALID;
  for (int g = 0; g < world; g++)
    if (!shards[g] || shards[g]->world != world || shards[g]->rank != g) return ERR_INVALID;
  for (int g = 0; g < world; g++) {
    lg_shard* s = shards[g];
    Ctx* c = &s->ctx->c;
    cudaSetDevice(c->device);
    for (int h = 0; h < world; h++) {
      if (h != g && shards[h]->ctx->c.device != c->device) {
        int can = 0;
        cudaDeviceCanAccessPeer(&can, c->device, shards[h]->ctx->c.device);
        if (!can) return sfail(s, ERR_UNSUPPORTED, "no peer access between the devices");
        cudaError_t e = cudaDeviceEnablePeerAccess(shards[h]->ctx->c.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return sfail(s, ERR_CUDA, cudaGetErrorString(e));
        cudaGetLastError();
      }
      s->peer_u[h] = shards[h]->cols->m.u;
      s->peer_rhat[h] = shards[h]->rhat ? shards[h]->rhat->m.u : nullptr;
      s->peers.mail[h] = shards[h]->mail;
    }
    s->via_ipc = false;
    s->connected = true;
  }
  return OK;
}

int lg_shard_set_pipeline(lg_shard* s, int enabled) {
  if (!s) return ERR_INVALID;
  s->pipeline = enabled < 0 ? -1 : (enabled > 2 ? 2 : enabled);
  return OK;
}

int lg_shard_layout(const lg_shard* s, size_t* rows_local, size_t* n_runs, size_t* row_base, size_t* nrows) {
  if (!s) return ERR_INVALID;
  if (rows_local) *rows_local = s->rows_local;
  if (n_runs) *n_runs = s->runs.size();
  for (size_t j = 0; j < s->runs.size(); j++) {
    if (row_base) row_base[j] = s->runs[j].row_base;
    if (nrows) nrows[j] = s->runs[j].nrows;
  }
  return OK;
}

lg_matrix* lg_shard_matrix(lg_shard* s) { return s ? s->cols : nullptr; }

int lg_shard_commit_async(lg_shard* s, const uint64_t* msg_local) {
  if (!s) return ERR_INVALID;
  cudaSetDevice(s->ctx->c.device);
  return commit_async(s, msg_local);
}

int lg_shard_root(lg_shard* s, uint8_t root_out[32], uint8_t* subtree_roots_out) {
  if (!s) return ERR_INVALID;
  cudaSetDevice(s->ctx->c.device);
  LG_TRY(finish_root(s, root_out));
  if (subtree_roots_out) memcpy(subtree_roots_out, s->subtree_roots, (size_t)s->world * 32);
  return OK;
}

int lg_shard_commit(lg_shard* s, const uint64_t* msg_local, uint8_t root_out[32]) {
  if (!s) return ERR_INVALID;
  cudaSetDevice(s->ctx->c.device);
  LG_TRY(commit_async(s, msg_local));
  return finish_root(s, root_out);
}

// the commit-and-test transcript (src/ligero/mod.rs:457-578) on this rank's rows of a ready pre-encoding matrix;
// every rank runs the same Fiat-Shamir transcript on the gathered values.  out may be NULL (this rank only helps).
static int lg_shard_prove_matrix_impl(lg_shard* s, lg_ligero* L, const uint64_t* local_rows, lg_sponge* sponge, lg_proof** out) {
  if (!s || !L || !sponge) return ERR_INVALID;
  Ctx* c = &s->ctx->c;
  cudaSetDevice(c->device);
  size_t m, k, n, t;
  LG_TRY(lg_ligero_params(L, &m, &k, &n, &t, nullptr));
  if (m != s->m || k != s->k || n != (size_t)s->rho * k || t > s->t_max || !s->rhat)
    return sfail(s, ERR_INVALID, "the shard was created for another circuit shape");
  const lg_constraints* A = nullptr;
  LG_TRY(lg_ligero_constraints(L, &A));
  typedef std::chrono::steady_clock Clock;
  const Clock::time_point t0 = Clock::now();
  const size_t rows = s->rows, kg = s->kg;
  uint8_t root[32], seed[32];
  LG_TRY(commit_async(s, local_rows));                                        // 521-551
  LG_TRY(finish_root(s, root));
  LG_TRY(lg_sponge_absorb_bytes(sponge, root, 32));                           // 560
  s->last_ms[0] = std::chrono::duration<double, std::milli>(Clock::now() - t0).count();
  // Test-Interleaved (646-669)
  LG_TRY(lg_sponge_squeeze_bytes(sponge, seed, 32));
  Fr *d_r = nullptr, *d_part = nullptr;
  LG_CUDA(c, cudaMallocAsync((void**)&d_r, rows * sizeof(Fr), c->stream));
  LG_CUDA(c, cudaMallocAsync((void**)&d_part, 2 * kg * sizeof(Fr), c->stream));
  LG_TRY(expand_fr(c, seed, rows, d_r));
  LG_TRY(col_reduce(c, 0, d_r, s->cols->m.u, nullptr, nullptr, rows, kg, d_part, 1, 0, true));
  std::vector<Fr> lc(k);
  LG_TRY(gather_slices(s, d_part, kg, lc.data()));
  cudaFreeAsync(d_r, c->stream);
  LG_TRY(lg_sponge_absorb_fr(sponge, (const uint64_t*)lc.data(), k));
  OpenedHost op[3];
  LG_TRY(open_sharded(s, n, t, sponge, out != nullptr, op[0]));
  // Test-Linear-Constraints (712-747): r_a replicated; its rows go to the odd points of the 2k domain through the same
  // row-sharded encode + NVLink scatter as the witness (rho = 2, Montgomery form kept)
  LG_TRY(lg_sponge_squeeze_bytes(sponge, seed, 32));
  Fr *d_ra = nullptr, *d_ra_local = nullptr;
  LG_CUDA(c, cudaMallocAsync((void**)&d_ra, rows * k * sizeof(Fr), c->stream));
  LG_TRY(lg_linear_ra(s->ctx, A, seed, (uint64_t*)d_ra));
  if (s->rows_local) {
    LG_CUDA(c, cudaMallocAsync((void**)&d_ra_local, s->rows_local * k * sizeof(Fr), c->stream));
    gather_row_runs_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(d_ra, k, s->local_src_rows, s->rows_local, d_ra_local);
    c->launches++;
    LG_CUDA(c, cudaGetLastError());
  }
  cudaFreeAsync(d_ra, c->stream);
  LG_TRY(encode_runs(s, (const uint64_t*)d_ra_local, 2, s->peer_rhat, false, nullptr, 0));
  if (d_ra_local) cudaFreeAsync(d_ra_local, c->stream);
  LG_TRY(post(s, c->stream, 0, nullptr, 0));
  LG_TRY(wait_all(s, c->stream));
  LG_TRY(lg_linear_evals(s->cols, (const uint64_t*)s->rhat->m.u, (const uint64_t*)(s->rhat->m.u + rows * kg), (uint64_t*)d_part));
  std::vector<Fr> ev(2 * k), poly(2 * k);
  LG_TRY(gather_slices(s, d_part, 2 * kg, ev.data()));
  size_t lin_len = 0, quad_len = 0;
  std::vector<Fr> lin(2 * k), quad(2 * k);
  LG_TRY(lg_poly_from_evals(s->ctx, (const uint64_t*)ev.data(), 2 * k, (uint64_t*)lin.data(), &lin_len));
  LG_TRY(lg_sponge_absorb_fr(sponge, (const uint64_t*)lin.data(), lin_len));
  LG_TRY(open_sharded(s, n, t, sponge, out != nullptr, op[1]));
  // Test-Quadratic-Constraints (832-859)
  LG_TRY(lg_sponge_squeeze_bytes(sponge, seed, 32));
  Fr* d_rq = nullptr;
  LG_CUDA(c, cudaMallocAsync((void**)&d_rq, m * sizeof(Fr), c->stream));
  LG_TRY(expand_fr(c, seed, m, d_rq));
  LG_TRY(lg_quadratic_evals(s->cols, (const uint64_t*)d_rq, (uint64_t*)d_part));
  cudaFreeAsync(d_rq, c->stream);
  LG_TRY(gather_slices(s, d_part, 2 * kg, ev.data()));
  cudaFreeAsync(d_part, c->stream);
  LG_TRY(lg_poly_from_evals(s->ctx, (const uint64_t*)ev.data(), 2 * k, (uint64_t*)quad.data(), &quad_len));
  LG_TRY(lg_sponge_absorb_fr(sponge, (const uint64_t*)quad.data(), quad_len));
  LG_TRY(open_sharded(s, n, t, sponge, out != nullptr, op[2]));
  s->last_ms[1] = std::chrono::duration<double, std::milli>(Clock::now() - t0).count();
  if (!out) return OK;
  int log_n = 0;
  while (((size_t)1 << log_n) < n) log_n++;
  const uint64_t* cols3[3] = {op[0].cols.data(), op[1].cols.data(), op[2].cols.data()};
  const uint64_t* idx3[3] = {op[0].idx.data(), op[1].idx.data(), op[2].idx.data()};
  const uint8_t* sib3[3] = {op[0].sib.data(), op[1].sib.data(), op[2].sib.data()};
  const uint8_t* auth3[3] = {op[0].auth.data(), op[1].auth.data(), op[2].auth.data()};
  return lg_proof_assemble(root, (const uint64_t*)lc.data(), k, (const uint64_t*)lin.data(), lin_len, (const uint64_t*)quad.data(), quad_len,
                           t, rows, (size_t)log_n - 1, cols3, idx3, sib3, auth3, out);
}
int lg_shard_prove_matrix(lg_shard* s, lg_ligero* L, const uint64_t* local_rows, lg_sponge* sponge, lg_proof** out) {
  return lg::guard([&]() { return lg_shard_prove_matrix_impl(s, L, local_rows, sponge, out); });
}

// LigeroCircuit::prove over the shards from the variable assignment: every rank runs the (cheap, replicated) evaluation
// trace on its own GPU, keeps its rows of [X;Y;Z;W] and goes on with lg_shard_prove_matrix
static int lg_shard_prove_impl(lg_shard* s, lg_ligero* L, const size_t* var_idx, const uint64_t* var_vals, size_t n_vars, int bump, lg_sponge* sponge,
                   lg_proof** out) {
  if (!s || !L || !sponge) return ERR_INVALID;
  Ctx* c = &s->ctx->c;
  cudaSetDevice(c->device);
  if (!s->pre_full) LG_CUDA(c, cudaMalloc((void**)&s->pre_full, s->rows * s->k * sizeof(Fr)));
  if (!s->pre_local && s->rows_local) LG_CUDA(c, cudaMalloc((void**)&s->pre_local, s->rows_local * s->k * sizeof(Fr)));
  LG_TRY(lg_ligero_witness_matrix_dev(L, var_idx, var_vals, n_vars, bump, (uint64_t*)s->pre_full));
  if (s->rows_local) {
    gather_row_runs_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(s->pre_full, s->k, s->local_src_rows, s->rows_local, s->pre_local);
    c->launches++;
    LG_CUDA(c, cudaGetLastError());
  }
  return lg_shard_prove_matrix(s, L, (const uint64_t*)s->pre_local, sponge, out);
}
int lg_shard_prove(lg_shard* s, lg_ligero* L, const size_t* var_idx, const uint64_t* var_vals, size_t n_vars, int bump, lg_sponge* sponge,
                   lg_proof** out) {
  return lg::guard([&]() { return lg_shard_prove_impl(s, L, var_idx, var_vals, n_vars, bump, sponge, out); });
}

int lg_shard_last_ms(const lg_shard* s, double ms_out[4]) {
  if (!s || !ms_out) return ERR_INVALID;
  memcpy(ms_out, s->last_ms, sizeof(s->last_ms));
  return OK;
}

// =====================================================================================================
// One process, all GPUs: the reference's single call (LigeroCircuit::prove) reaching G devices.
// =====================================================================================================
int lg_mgpu_create(const int* dev_ids, int n_dev, lg_mgpu** out) {
  if (!out || n_dev < 1 || n_dev > kMaxRanks || (n_dev & (n_dev - 1))) return ERR_INVALID;
  lg_mgpu* g = new (std::nothrow) lg_mgpu();
  if (!g) return ERR_NOMEM;
  g->world = n_dev;
  for (int i = 0; i < n_dev; i++) {
    int st = lg_ctx_create(dev_ids ? dev_ids[i] : i, &g->ctx[i]);
    if (st != OK) {
      lg_mgpu_destroy(g);
      return st;
    }
  }
  *out = g;
  return OK;
}

int lg_mgpu_destroy(lg_mgpu* g) {
  if (!g) return OK;
  for (int i = 0; i < g->world; i++)
    if (g->ctx[i]) lg_ctx_destroy(g->ctx[i]);
  delete g;
  return OK;
}

const char* lg_mgpu_last_error(const lg_mgpu* g) { return g ? g->error.c_str() : "null handle"; }
lg_ctx* lg_mgpu_ctx(lg_mgpu* g, int i) { return (g && i >= 0 && i < g->world) ? g->ctx[i] : nullptr; }

// encode + commit of a whole rows x k host (or device-0 ... any addressable) matrix over all GPUs: root only
static int lg_mgpu_commit_impl(lg_mgpu* g, const uint64_t* preenc_u, size_t rows, size_t k, uint32_t rho_inv, uint8_t root_out[32]) {
  if (!g || !preenc_u || rows % 4) return ERR_INVALID;
  lg_shard* sh[kMaxRanks] = {};
  int st = on_all(g, [&](int i) { return lg_shard_create(g->ctx[i], rows / 4, k, rho_inv, i, g->world, 0, 0, &sh[i]); });
  if (st == OK) st = lg_shard_connect_local(sh, g->world);
  if (st == OK) {
    uint8_t roots[kMaxRanks][32];
    st = on_all(g, [&](int i) {
      // this rank's rows, run by run, straight from the caller's matrix (pageable or pinned host memory)
      lg_shard* s = sh[i];
      std::vector<uint64_t> local(s->rows_local * k * 4);
      for (const Run& r : s->runs)
        memcpy(local.data() + r.local_off * k * 4, preenc_u + r.row_base * k * 4, r.nrows * k * sizeof(Fr));
      return lg_shard_commit(s, local.data(), roots[i]);
    });
    if (st == OK && root_out) memcpy(root_out, roots[0], 32);
  }
  for (int i = 0; i < g->world; i++)
    if (sh[i]) lg_shard_free(sh[i]);
  return st;
}
int lg_mgpu_commit(lg_mgpu* g, const uint64_t* preenc_u, size_t rows, size_t k, uint32_t rho_inv, uint8_t root_out[32]) {
  return lg::guard([&]() { return lg_mgpu_commit_impl(g, preenc_u, rows, k, rho_inv, root_out); });
}

static int lg_mgpu_ligero_new_impl(lg_mgpu* g, const lg_circuit* circuit, const size_t* outputs, size_t n_outputs, size_t lambda, lg_mligero** out) {
  if (!g || !circuit || !out) return ERR_INVALID;
  lg_mligero* ml = new (std::nothrow) lg_mligero();
  if (!ml) return ERR_NOMEM;
  ml->mg = g;
  int st = on_all(g, [&](int i) {
    LG_TRY(lg_ligero_new(g->ctx[i], circuit, outputs, n_outputs, lambda, &ml->L[i]));
    LG_TRY(lg_ligero_set_trace_mode(ml->L[i], 1));
    size_t m, k, n, t;
    LG_TRY(lg_ligero_params(ml->L[i], &m, &k, &n, &t, nullptr));
    return lg_shard_create(g->ctx[i], m, k, (uint32_t)(n / k), i, g->world, t, 0, &ml->sh[i]);
  });
  if (st == OK) st = lg_shard_connect_local(ml->sh, g->world);
  if (st != OK) {
    lg_mgpu_ligero_free(ml);
    return st;
  }
  *out = ml;
  return OK;
}
int lg_mgpu_ligero_new(lg_mgpu* g, const lg_circuit* circuit, const size_t* outputs, size_t n_outputs, size_t lambda, lg_mligero** out) {
  return lg::guard([&]() { return lg_mgpu_ligero_new_impl(g, circuit, outputs, n_outputs, lambda, out); });
}

int lg_mgpu_ligero_free(lg_mligero* ml) {
  if (!ml) return OK;
  for (int i = 0; i < ml->mg->world; i++) {
    if (ml->sh[i]) lg_shard_free(ml->sh[i]);
    if (ml->L[i]) lg_ligero_free(ml->L[i]);
  }
  delete ml;
  return OK;
}

// LigeroCircuit::prove (bump != 0) / prove_inner over all GPUs; the caller's sponge is advanced exactly as by lg_prove
static int lg_mgpu_prove_impl(lg_mligero* ml, const size_t* var_idx, const uint64_t* var_vals, size_t n_vars, int bump, lg_sponge* sponge,
                  lg_proof** out) {
  if (!ml || !sponge || !out) return ERR_INVALID;
  lg_mgpu* g = ml->mg;
  lg_sponge* sp[kMaxRanks] = {};
  sp[0] = sponge;
  for (int i = 1; i < g->world; i++) LG_TRY(lg_sponge_clone(sponge, &sp[i]));
  int st = on_all(g, [&](int i) { return lg_shard_prove(ml->sh[i], ml->L[i], var_idx, var_vals, n_vars, bump, sp[i], i == 0 ? out : nullptr); });
  for (int i = 1; i < g->world; i++) lg_sponge_free(sp[i]);
  return st;
}
int lg_mgpu_prove(lg_mligero* ml, const size_t* var_idx, const uint64_t* var_vals, size_t n_vars, int bump, lg_sponge* sponge,
                  lg_proof** out) {
  return lg::guard([&]() { return lg_mgpu_prove_impl(ml, var_idx, var_vals, n_vars, bump, sponge, out); });
}

}  // extern "C"
