// Host driver: the reference's public surface rebuilt over the device primitives.
//   ArithmeticCircuit  (src/arithmetic_circuit/mod.rs)        -> lg_circuit_*
//   LigeroCircuit::new (src/ligero/mod.rs:147-433)            -> lg_ligero_new
//   LigeroCircuit::prove / prove_inner (435-578, 646-669, 712-747, 832-859, 935-955) -> lg_prove
//   LigeroCircuit::verify (613-644, 671-708, 749-830, 861-933, 957-996)              -> lg_verify
//   PoseidonSponge (ark-crypto-primitives; stays on the host)  -> lg_sponge_*
// Same names, argument meaning and failure behaviour as the reference (its panics become LG_ERR_*
// codes with the panic text in lg_last_error).  All heavy arithmetic goes through the C-ABI entry
// points of capi.cu / capi_protocol.cu on the GPU; this file only sequences the Fiat-Shamir transcript.
#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "capi_types.h"
#include "circuit_host.h"
#include "host_field.h"
#include "host_prng.h"

using namespace lg;
using lgh::Fq;
using lgh::N_ADD;
using lgh::N_CONST;
using lgh::N_MUL;
using lgh::N_VAR;
using lgh::Node;  // circuit_host.h

namespace {

typedef std::array<uint8_t, 32> Digest;

// pinned host buffers for opened columns, recycled across proofs (cudaHostAlloc of 82 MB costs more than the copy it serves)
struct PinnedPool {
  std::mutex mu;
  std::multimap<size_t, void*> idle;
  void* get(size_t bytes, size_t* got) {
    {
      std::lock_guard<std::mutex> g(mu);
      auto it = idle.lower_bound(bytes);
      if (it != idle.end() && it->first <= 2 * bytes + 4096) {
        void* p = it->second;
        *got = it->first;
        idle.erase(it);
        return p;
      }
    }
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    *got = bytes;
    return p;
  }
  void put(void* p, size_t bytes) {
    std::lock_guard<std::mutex> g(mu);
    size_t held = 0;
    for (auto& kv : idle) held += kv.first;
    if (held + bytes > ((size_t)1 << 30)) {  // keep at most 1 GiB idle
      cudaFreeHost(p);
      return;
    }
    idle.emplace(bytes, p);
  }
};
PinnedPool& pinned_pool() {
  static PinnedPool* pool = new PinnedPool();  // leaked on purpose: the CUDA runtime may be gone at static destruction
  return *pool;
}

// t opened columns of `rows` elements each, stored column after column in ONE buffer: the prover's buffer comes from the
// pinned pool and the device-to-host copy of an opening lands in it directly (no staging copy, no per-column vectors: 82 MB
// per opening at 2^24 gates); a deserialised proof owns a vector.  Columns of unequal length (malformed proofs) are kept
// apart so that they serialise back unchanged and never verify.
struct Opened {
  Fq* cols = nullptr;
  size_t t = 0, rows = 0;
  size_t pinned_bytes = 0;  // > 0: cols belongs to the pinned pool
  std::vector<Fq> own;
  std::vector<std::vector<Fq>> ragged;
  bool is_ragged = false;
  std::vector<uint64_t> leaf_index;
  std::vector<Digest> sibling;
  std::vector<std::vector<Digest>> auth;

  Opened() = default;
  Opened(const Opened&) = delete;
  Opened& operator=(const Opened&) = delete;
  Opened(Opened&& o) noexcept { *this = std::move(o); }
  Opened& operator=(Opened&& o) noexcept {
    if (this == &o) return *this;
    release();
    const bool was_own = o.pinned_bytes == 0 && o.cols != nullptr;
    own = std::move(o.own);
    cols = was_own ? own.data() : o.cols;
    t = o.t;
    rows = o.rows;
    pinned_bytes = o.pinned_bytes;
    ragged = std::move(o.ragged);
    is_ragged = o.is_ragged;
    leaf_index = std::move(o.leaf_index);
    sibling = std::move(o.sibling);
    auth = std::move(o.auth);
    o.cols = nullptr;
    o.pinned_bytes = 0;
    o.t = o.rows = 0;
    return *this;
  }
  ~Opened() { release(); }
  void release() {
    if (pinned_bytes) pinned_pool().put(cols, pinned_bytes);
    cols = nullptr;
    pinned_bytes = 0;
  }
  // false: out of (pinned) memory
  bool alloc(size_t t_, size_t rows_, bool pinned) {
    release();
    t = t_;
    rows = rows_;
    is_ragged = false;
    if (pinned && t * rows) {
      cols = (Fq*)pinned_pool().get(t * rows * sizeof(Fq), &pinned_bytes);
      if (cols) return true;
      pinned_bytes = 0;
    }
    own.assign(t * rows, lgh::kZero);
    cols = own.data();
    return true;
  }
  const Fq* col(size_t q) const { return cols + q * rows; }
  size_t n_columns() const { return is_ragged ? ragged.size() : t; }
};

}  // namespace

struct lg_circuit {
  lgh::RawVec<Node> nodes;  // (a vector whose resize() does not zero: LigeroCircuit::new fills its copy on all host threads)
  std::vector<Fq> const_values;
  std::vector<std::string> labels;
  std::map<Fq, size_t> constants;            // value -> node index
  std::map<std::string, size_t> variables;   // label -> node index
  std::string error;
  size_t push(Node n) {
    nodes.push_back(n);
    return nodes.size() - 1;
  }
};

struct lg_sponge {
  lgh::PoseidonSponge s;
  explicit lg_sponge(const lgh::PoseidonConfig& c) : s(c) {}
};

struct lg_ligero {
  lg_ctx* ctx = nullptr;
  lg_circuit circuit;  // formatted copy: constant 1 at node 0
  std::vector<size_t> outputs;
  size_t one_index = 0;
  bool one_found = false;
  size_t m = 0, k = 0, n = 0, t = 0, sol_len = 0;
  lg_constraints* a = nullptr;
  // evaluation trace on the device (trace.cu): level schedule, reachability from the outputs, witness slots
  lg::TraceSchedule trace;
  lgh::RawVec<uint32_t> index_map;  // node -> slot in the X/Y/Z/W blocks (0xffffffff for dropped constants)
  lgh::RawVec<uint8_t> reach;       // node feeds an output
  lgh::RawVec<uint32_t> var_nodes;  // every Variable node
  bool all_gates_reach = true;
  int trace_mode = -1;              // -1: by circuit shape, 0: host evaluator, 1: device
  // host wall clock of the last prove, ms: trace+layout, commit, interleaved test, linear test, quadratic test,
  // the three openings together, whole call (every phase ends on a device-to-host copy, so these are synchronised)
  double prove_ms[7] = {0, 0, 0, 0, 0, 0, 0};
  // device buffers kept between proofs of this circuit (allocating and freeing 36 GiB per proof at 2^24 gates costs
  // more than the three tests together): the codeword matrix with its leaves and tree, and the [X;Y;Z;W] matrix of
  // the device trace.  lg_ligero_release_buffers / lg_ligero_free return them.
  lg_matrix* u_cache = nullptr;
  uint64_t* pre_cache = nullptr;
  // the verifier's device scratch, kept the same way (cudaMalloc / cudaFree of the 4 GiB and 2 GiB blocks per call made
  // lg_verify at 2^24 gates jitter between 0.33 s and 2 s): see VSlot.  r_a shares pre_cache (same size, never live
  // at the same time as the prover's matrix).
  void* vbuf[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};
enum VSlot { V_R = 0, V_CHK, V_PLANES, V_COLS, V_RCOLS, V_IDX, V_RQ, V_DIG };

struct lg_proof {
  Digest root;
  std::vector<Fq> preenc_u_lc;
  Opened interleaved;
  std::vector<Fq> linear_poly;
  Opened linear;
  std::vector<Fq> quadratic_poly;
  Opened quadratic;
};

namespace {

int fail(lg_ctx* ctx, int code, const std::string& msg) { return set_error(ctx ? &ctx->c : nullptr, code, msg); }

using lgh::bump_index;  // src/ligero/mod.rs:230-242 (circuit_host.h)

size_t calculate_t(size_t sec_param, size_t d_num, size_t d_den, size_t codeword_len) {
  const double residual = (double)codeword_len / std::pow(2.0, 254);
  const double rhs = std::log2(std::pow(2.0, -(double)sec_param) - residual);
  const double nom = rhs - 1.0;
  const double denom = std::log2(1.0 - 0.5 * (double)d_num / (double)d_den);
  const size_t t = (size_t)std::ceil(nom / denom);
  return t < codeword_len ? t : codeword_len;
}

// iterative version of inner_evaluate (src/arithmetic_circuit/mod.rs:247-271): same values, no recursion
int evaluation_trace(const lg_circuit& c, const std::vector<std::pair<size_t, Fq>>& vars, const std::vector<size_t>& outputs,
                     std::vector<Fq>& vals, std::vector<uint8_t>& set, std::string& err) {
  const size_t N = c.nodes.size();
  vals.assign(N, lgh::kZero);
  set.assign(N, 0);
  for (size_t i = 0; i < N; i++)
    if (c.nodes[i].type == N_CONST) {
      vals[i] = c.const_values[c.nodes[i].l];
      set[i] = 1;
    }
  for (auto& v : vars) {
    if (v.first >= N || c.nodes[v.first].type != N_VAR) {
      err = "Value supplied for non-variable node";
      return ERR_INVALID;
    }
    vals[v.first] = v.second;
    set[v.first] = 1;
  }
  std::vector<size_t> stack;
  for (size_t out : outputs) {
    if (out >= N) {
      err = "output node not in circuit";
      return ERR_INVALID;
    }
    stack.push_back(out);
    while (!stack.empty()) {
      const size_t i = stack.back();
      if (set[i]) {
        stack.pop_back();
        continue;
      }
      const Node& nd = c.nodes[i];
      if (nd.type == N_VAR) {
        err = "Uninitialised variable";
        return ERR_INVALID;
      }
      if (!set[nd.l]) {
        stack.push_back(nd.l);
        continue;
      }
      if (!set[nd.r]) {
        stack.push_back(nd.r);
        continue;
      }
      vals[i] = nd.type == N_ADD ? lgh::add(vals[nd.l], vals[nd.r]) : lgh::mul(vals[nd.l], vals[nd.r]);
      set[i] = 1;
      stack.pop_back();
    }
  }
  return OK;
}

// host radix-2 transform, natural order (only for the verifier's single size-2k FFT / size-k pieces)
void host_fft(std::vector<Fq>& a, bool inverse) {
  const size_t n = a.size();
  int log_n = 0;
  while (((size_t)1 << log_n) < n) log_n++;
  for (size_t i = 0; i < n; i++) {
    size_t j = 0;
    for (int b = 0; b < log_n; b++) j |= ((i >> b) & 1) << (log_n - 1 - b);
    if (i < j) std::swap(a[i], a[j]);
  }
  Fq w = lgh::root_of_unity(log_n);
  if (inverse) w = lgh::inv(w);
  std::vector<Fq> tw(n / 2 ? n / 2 : 1);
  tw[0] = lgh::kOne;
  for (size_t i = 1; i < tw.size(); i++) tw[i] = lgh::mul(tw[i - 1], w);
  for (size_t len = 2; len <= n; len <<= 1) {
    const size_t half = len / 2, step = n / len;
    for (size_t s = 0; s < n; s += len)
      for (size_t j = 0; j < half; j++) {
        const Fq u = a[s + j], v = lgh::mul(a[s + j + half], tw[j * step]);
        a[s + j] = lgh::add(u, v);
        a[s + j + half] = lgh::sub(u, v);
      }
  }
  if (inverse) {
    const Fq ninv = lgh::inv(lgh::from_u64(n));
    for (auto& x : a) x = lgh::mul(x, ninv);
  }
}

// ---- constraint matrix (right-hand block of A) straight into CSC ------------------------------------
// value ids of the constants of the circuit: 0 -> +1, 1 -> -1, v >= 2 -> table[v - 2].  Canonical numbering shared by the
// host and the device builder: constant nodes in node order, each registering its value c and then -c (an Add gate uses c,
// a Mul gate -c: mod.rs:324-336, 343-355).  vidp / vidn are indexed like const_values.
void constant_value_ids(const lg_circuit& c, std::vector<uint32_t>& vidp, std::vector<uint32_t>& vidn, std::vector<Fq>& table) {
  std::map<Fq, uint32_t> ids;
  const Fq minus_one = lgh::neg(lgh::kOne);
  auto vid_of = [&](const Fq& v) -> uint32_t {
    if (v == lgh::kOne) return 0;
    if (v == minus_one) return 1;
    auto it = ids.find(v);
    if (it != ids.end()) return it->second;
    const uint32_t id = (uint32_t)table.size() + 2;
    table.push_back(v);
    ids[v] = id;
    return id;
  };
  vidp.assign(c.const_values.size(), 0);
  vidn.assign(c.const_values.size(), 0);
  std::vector<uint8_t> done(c.const_values.size(), 0);
  for (const Node& nd : c.nodes)
    if (nd.type == N_CONST && !done[nd.l]) {
      done[nd.l] = 1;
      vidp[nd.l] = vid_of(c.const_values[nd.l]);
      vidn[nd.l] = vid_of(lgh::neg(c.const_values[nd.l]));
    }
}

// ark_bn254::Fr is always canonical; a C caller's limbs might not be: every value must be < r
bool limbs_canonical(const uint64_t* vals, size_t count) {
  for (size_t i = 0; i < count; i++)
    if (lgh::geq_p(vals + 4 * i)) return false;
  return true;
}

double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
bool debug_timing() {
  const char* e = getenv("LG_DEBUG_TIMING");
  return e && atoi(e) != 0;
}

int build_constraints(lg_ligero* L, const lgh::NodeArrays& arr, std::string& err) {
  const lg_circuit& c = L->circuit;
  const auto& nodes = c.nodes;
  const size_t mk = L->m * L->k;
  std::vector<uint32_t> vidp, vidn;
  std::vector<Fq> table;
  constant_value_ids(c, vidp, vidn, table);
  // the reference panics on these (mod.rs:325, 345, 369-414): refuse them before building anything
  {
    const size_t i = lgh::first_gate_of_two_constants(arr, lgh::host_threads());
    if (i != SIZE_MAX) {
      err = nodes[i].type == N_ADD ? "Add(constant, constant) is not supported (the reference panics at src/ligero/mod.rs:325)"
                                   : "Mul(constant, constant) is not supported (the reference panics at src/ligero/mod.rs:345)";
      return ERR_UNSUPPORTED;
    }
  }
  for (size_t o : L->outputs) {
    const Node& nd = nodes[o];
    if (nd.type != N_ADD && nd.type != N_MUL) {
      err = "The output node must be an addition or multiplication gate";
      return ERR_INVALID;
    }
  }
  // large circuits: CSC built on the device (constraints.cu); LG_CSC_DEVICE=0 / 1 forces the host / device builder
  bool on_device = nodes.size() >= ((size_t)1 << 16);
  if (const char* e = getenv("LG_CSC_DEVICE")) on_device = atoi(e) != 0;
  if (on_device) {
    const double t1 = now_ms();
    std::vector<uint32_t> outs(L->outputs.begin(), L->outputs.end());
    const int s = lg::build_constraints_device(L->ctx, arr.type.get(), arr.l.get(), arr.r.get(), arr.n, vidp.data(), vidn.data(), vidp.size(),
                                               outs.data(), outs.size(), mk, table.empty() ? nullptr : (const uint64_t*)table.data(),
                                               table.size(), &L->a);
    if (debug_timing()) fprintf(stderr, "[lg] constraint matrix on the device: build %.1f ms\n", now_ms() - t1);
    if (s != OK) err = lg_last_error(L->ctx);
    return s;
  }
  // index_map: node -> position after dropping every constant but node 0 (mod.rs:179-194)
  std::vector<uint32_t> index_map(nodes.size(), 0xffffffffu);
  index_map[0] = 0;
  size_t seen = 0;
  for (size_t i = 1; i < nodes.size(); i++) {
    if (nodes[i].type == N_CONST) seen++;
    else index_map[i] = (uint32_t)(i - seen);
  }
  struct Triplet {
    uint32_t col, row, vid;
  };
  std::vector<Triplet> trip;
  size_t row = 0;  // row inside each P matrix
  auto add_gate = [&](size_t l, size_t r, uint32_t own_col) {  // P_add row
    const bool lc = nodes[l].type == N_CONST, rc = nodes[r].type == N_CONST;
    const uint32_t base = (uint32_t)(3 * mk + row);
    if (lc) {
      trip.push_back({0, base, vidp[nodes[l].l]});
      trip.push_back({index_map[r], base, 0});
    } else if (rc) {
      trip.push_back({index_map[l], base, 0});
      trip.push_back({0, base, vidp[nodes[r].l]});
    } else {
      trip.push_back({index_map[l], base, 0});
      trip.push_back({index_map[r], base, 0});
    }
    trip.push_back({own_col, base, 1});
  };
  auto mul_gate = [&](size_t l, size_t r, uint32_t own_col) {  // rows of -P_x, -P_y, -P_z
    const bool lc = nodes[l].type == N_CONST, rc = nodes[r].type == N_CONST;
    const uint32_t rx = (uint32_t)row, ry = (uint32_t)(mk + row), rz = (uint32_t)(2 * mk + row);
    if (lc) {
      trip.push_back({0, rx, vidn[nodes[l].l]});
      trip.push_back({index_map[r], ry, 1});
    } else if (rc) {
      trip.push_back({index_map[l], rx, 1});
      trip.push_back({0, ry, vidn[nodes[r].l]});
    } else {
      trip.push_back({index_map[l], rx, 1});
      trip.push_back({index_map[r], ry, 1});
    }
    trip.push_back({own_col, rz, 1});
  };
  for (size_t i = 0; i < nodes.size(); i++) {
    const Node& nd = nodes[i];
    if (nd.type == N_CONST && i != 0) continue;  // constants other than node 0 get no row
    if (row >= mk) {
      err = "internal: more rows than m*k";
      return ERR_STATE;
    }
    if (nd.type == N_ADD) add_gate(nd.l, nd.r, index_map[i]);
    else if (nd.type == N_MUL) mul_gate(nd.l, nd.r, index_map[i]);
    row++;
  }
  for (size_t o : L->outputs) {  // o = 1 for each output node (mod.rs:369-414)
    const Node& nd = nodes[o];
    if (row >= mk) {
      err = "internal: more rows than m*k";
      return ERR_STATE;
    }
    if (nd.type == N_ADD) add_gate(nd.l, nd.r, 0);
    else mul_gate(nd.l, nd.r, 0);
    row++;
  }
  // CSC by a counting sort on the column (stable: entries of a column keep the order in which the rows produced them)
  std::vector<uint32_t> col_ptr(mk + 1, 0), row_idx(trip.size()), val_id(trip.size());
  for (const auto& t : trip) col_ptr[t.col + 1]++;
  for (size_t cidx = 0; cidx < mk; cidx++) col_ptr[cidx + 1] += col_ptr[cidx];
  {
    std::vector<uint32_t> cursor(col_ptr.begin(), col_ptr.end() - 1);
    for (const auto& t : trip) {
      const uint32_t e = cursor[t.col]++;
      row_idx[e] = t.row;
      val_id[e] = t.vid;
    }
  }
  return lg_constraints_create(L->ctx, mk, col_ptr.data(), row_idx.data(), val_id.data(), trip.size(),
                               table.empty() ? nullptr : (const uint64_t*)table.data(), table.size(), &L->a);
}

// ---- level schedule of the circuit for the device evaluator (trace.cu) ----------------------------------
constexpr size_t kNarrowLevel = 2048;  // levels of at most this many gates are walked by a single CTA

template <class V, class T = typename V::value_type>
int upload(lg_ctx* ctx, const V& v, T** out) {
  *out = nullptr;
  if (v.empty()) return OK;
  if (cudaMalloc((void**)out, v.size() * sizeof(T)) != cudaSuccess) {
    cudaGetLastError();
    return fail(ctx, ERR_NOMEM, "out of device memory for the circuit schedule");
  }
  if (cudaMemcpy(*out, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess)
    return fail(ctx, ERR_CUDA, "schedule upload failed");
  return OK;
}

int build_trace(lg_ligero* L, const lgh::NodeArrays& arr) {
  const lg_circuit& c = L->circuit;
  const auto& nodes = c.nodes;
  const size_t N = nodes.size();
  cudaSetDevice(L->ctx->c.device);
  // slot map, reachability from the outputs, levels and the gates sorted by (level, Add before Mul): circuit_host.h
  lgh::Schedule sch;
  lgh::build_schedule(arr, L->outputs.data(), L->outputs.size(), sch, lgh::host_threads());
  L->index_map = std::move(sch.index_map);
  L->reach = std::move(sch.reach);
  L->all_gates_reach = sch.all_gates_reach;
  L->var_nodes = std::move(sch.var_nodes);
  const size_t depth = sch.depth;
  const std::vector<uint32_t>& level_start = sch.level_start;
  lg::TraceSchedule& t = L->trace;
  t.n_nodes = N;
  t.n_gates = sch.n_gates;
  t.n_levels = sch.depth;
  t.mk = L->m * L->k;
  const lgh::RawVec<uint32_t>&gnode = sch.gate_node, &gl = sch.gate_l, &gr = sch.gate_r, &gpos = sch.gate_pos, &cnode = sch.const_nodes;
  std::vector<uint32_t> cpos(cnode.size());
  std::vector<Fq> cval(cnode.size());
  for (size_t j = 0; j < cnode.size(); j++) {
    cpos[j] = L->index_map[cnode[j]];
    cval[j] = c.const_values[nodes[cnode[j]].l];
  }
  t.n_consts = cnode.size();
  // segments: wide levels get a launch each, runs of narrow levels share a single-CTA launch
  t.segments.clear();
  for (size_t l = 0; l < depth;) {
    const size_t width = level_start[l + 1] - level_start[l];
    if (width > kNarrowLevel) {
      t.segments.push_back({false, l, l + 1, level_start[l], level_start[l + 1]});
      l++;
    } else {
      size_t e = l;
      while (e < depth && (size_t)(level_start[e + 1] - level_start[e]) <= kNarrowLevel) e++;
      t.segments.push_back({true, l, e, level_start[l], level_start[e]});
      l = e;
    }
  }
  LG_TRY(upload(L->ctx, gnode, &t.gate_node));
  LG_TRY(upload(L->ctx, gl, &t.gate_l));
  LG_TRY(upload(L->ctx, gr, &t.gate_r));
  LG_TRY(upload(L->ctx, gpos, &t.gate_pos));
  LG_TRY(upload(L->ctx, level_start, &t.level_start));
  LG_TRY(upload(L->ctx, cnode, &t.const_node));
  LG_TRY(upload(L->ctx, cpos, &t.const_pos));
  LG_TRY(upload(L->ctx, cval, (Fq**)&t.const_val));
  if (cudaMalloc((void**)&t.vals, N * sizeof(Fq)) != cudaSuccess) {
    cudaGetLastError();
    return fail(L->ctx, ERR_NOMEM, "out of device memory for the value table");
  }
  return OK;
}

// by default the trace runs on the device when the circuit is wide enough for it: a level costs a launch (or a
// CTA-wide barrier), so deep and thin circuits (R1CS-compiled ones) are evaluated faster by the host loop
bool trace_on_device(const lg_ligero* L) {
  if (L->trace_mode >= 0) return L->trace_mode != 0;
  return L->trace.n_gates >= ((size_t)1 << 16) && L->trace.n_levels * 256 <= L->trace.n_gates;
}

// ---- openings -----------------------------------------------------------------------------------------
int open_columns(lg_ligero* L, lg_matrix* U, lgh::PoseidonSponge& sponge, Opened& out) {  // mod.rs:935-955
  const std::vector<uint8_t> seed = sponge.squeeze_bytes(32);
  std::vector<uint64_t> idx(L->t);
  LG_TRY(lg_expand_indices(seed.data(), L->n, L->t, idx.data()));
  const size_t rows = 4 * L->m, t = L->t;
  int log_n = 0;
  while (((size_t)1 << log_n) < L->n) log_n++;
  const size_t depth = (size_t)(log_n - 1);
  Ctx* c = &L->ctx->c;
  cudaSetDevice(c->device);
  for (size_t q = 0; q < t; q++)
    if (idx[q] >= U->m.n) return fail(L->ctx, ERR_INVALID, "column index out of range");
  // The t x R elements (82 MB at 2^24 gates) go device -> pinned proof buffer on their own stream: nothing the
  // transcript does next depends on them (the openings are not absorbed), so the copy runs behind the next test and
  // lg_prove_matrix waits for it once, at the end.
  if (!c->open_stream) {
    LG_CUDA(c, cudaStreamCreateWithFlags(&c->open_stream, cudaStreamNonBlocking));
    LG_CUDA(c, cudaEventCreateWithFlags(&c->ev_gathered, cudaEventDisableTiming));
  }
  out.alloc(t, rows, true);
  uint64_t* d_idx = nullptr;
  Fr* d_cols = nullptr;
  uint8_t *d_sib = nullptr, *d_auth = nullptr;
  LG_CUDA(c, cudaMallocAsync((void**)&d_idx, t * 8, c->stream));
  LG_CUDA(c, cudaMallocAsync((void**)&d_cols, t * rows * sizeof(Fr), c->stream));
  LG_CUDA(c, cudaMallocAsync((void**)&d_sib, t * 32, c->stream));
  LG_CUDA(c, cudaMallocAsync((void**)&d_auth, t * depth * 32 + 32, c->stream));
  LG_CUDA(c, cudaMemcpyAsync(d_idx, idx.data(), t * 8, cudaMemcpyHostToDevice, c->stream));
  lg::phase_mark(c, lg::PH_BEGIN);
  LG_TRY(lg::gather_open(c, U->m, d_idx, t, d_cols, d_sib, d_auth));
  lg::phase_mark(c, lg::PH_OPEN);
  LG_CUDA(c, cudaEventRecord(c->ev_gathered, c->stream));
  LG_CUDA(c, cudaStreamWaitEvent(c->open_stream, c->ev_gathered, 0));
  LG_CUDA(c, cudaMemcpyAsync(out.cols, d_cols, t * rows * sizeof(Fr), cudaMemcpyDeviceToHost, c->open_stream));
  LG_CUDA(c, cudaFreeAsync(d_cols, c->open_stream));  // released when the copy is done
  std::vector<uint8_t> sib(t * 32), auth(t * depth * 32 + 1);
  LG_CUDA(c, cudaMemcpyAsync(sib.data(), d_sib, t * 32, cudaMemcpyDeviceToHost, c->stream));
  if (depth) LG_CUDA(c, cudaMemcpyAsync(auth.data(), d_auth, t * depth * 32, cudaMemcpyDeviceToHost, c->stream));
  LG_CUDA(c, cudaFreeAsync(d_idx, c->stream));
  LG_CUDA(c, cudaFreeAsync(d_sib, c->stream));
  LG_CUDA(c, cudaFreeAsync(d_auth, c->stream));
  LG_CUDA(c, cudaStreamSynchronize(c->stream));
  out.leaf_index = idx;
  out.sibling.resize(t);
  out.auth.assign(t, std::vector<Digest>(depth));
  for (size_t q = 0; q < t; q++) {
    memcpy(out.sibling[q].data(), sib.data() + 32 * q, 32);
    for (size_t d = 0; d < depth; d++) memcpy(out.auth[q][d].data(), auth.data() + 32 * (q * depth + d), 32);
  }
  return OK;
}

bool verify_path(const lg_ctx* ctx, const Digest& root, const Digest& leaf, const Digest& sibling, const std::vector<Digest>& auth,
                 uint64_t index) {
  uint8_t buf[80];
  size_t len = 0;
  const Digest* lr[2] = {(index & 1) ? &sibling : &leaf, (index & 1) ? &leaf : &sibling};
  for (int s = 0; s < 2; s++) {
    if (ctx->leaf_len_prefix) {
      const uint64_t l = 32;
      memcpy(buf + len, &l, 8);
      len += 8;
    }
    memcpy(buf + len, lr[s]->data(), 32);
    len += 32;
  }
  Digest cur;
  lgh::sha256(buf, len, cur.data());
  index >>= 1;
  for (size_t d = auth.size(); d-- > 0;) {
    uint8_t b2[64];
    if (index & 1) {
      memcpy(b2, auth[d].data(), 32);
      memcpy(b2 + 32, cur.data(), 32);
    } else {
      memcpy(b2, cur.data(), 32);
      memcpy(b2 + 32, auth[d].data(), 32);
    }
    lgh::sha256(b2, 64, cur.data());
    index >>= 1;
  }
  return cur == root;
}

// verify_column_openings (mod.rs:957-996): re-derive the indices, hash the columns on the GPU, check paths
// plain cudaMalloc'ed buffer released at scope exit (the verifier returns early on every failed check)
struct DevMem {
  void* p = nullptr;
  bool owned = true;
  ~DevMem() {
    if (p && owned) cudaFree(p);
  }
  int alloc(Ctx* c, size_t bytes) {
    LG_CUDA(c, cudaMalloc(&p, bytes ? bytes : 1));
    return OK;
  }
  // a buffer that lives in *slot (owned by the lg_ligero object) across calls; every slot has one size per circuit
  int cached(Ctx* c, void** slot, size_t bytes) {
    if (!*slot) LG_CUDA(c, cudaMalloc(slot, bytes ? bytes : 1));
    p = *slot;
    owned = false;
    return OK;
  }
};

// the verifier's R_A columns, tile by tile: planes of `nr` encoded rows (Montgomery form, plane stride nr*k) -> out[q][row0 + i]
__global__ void gather_tile_columns_kernel(const Fr* __restrict__ planes, size_t nr, int log_k, int rho, const uint64_t* __restrict__ idx,
                                           size_t t, size_t row0, size_t rows_total, Fr* __restrict__ out) {
  const size_t k = (size_t)1 << log_k, tot = t * nr;
  for (size_t f = (size_t)blockIdx.x * blockDim.x + threadIdx.x; f < tot; f += (size_t)gridDim.x * blockDim.x) {
    const size_t q = f / nr, i = f % nr;
    const size_t j = idx[q], s = j % rho, c = j / rho;
    const uint4* src = reinterpret_cast<const uint4*>(planes + s * nr * k + i * k + c);
    uint4* dst = reinterpret_cast<uint4*>(out + q * rows_total + row0 + i);
    dst[0] = src[0];
    dst[1] = src[1];
  }
}

// cols_keep (optional): receives the device copy of the t opened columns (t x rows, contiguous) for the per-column checks
int verify_openings(lg_ligero* L, const Opened& o, const Digest& root, lgh::PoseidonSponge& sponge, bool* ok,
                    DevMem* cols_keep = nullptr) {
  *ok = false;
  const std::vector<uint8_t> seed = sponge.squeeze_bytes(32);
  std::vector<uint64_t> idx(L->t);
  LG_TRY(lg_expand_indices(seed.data(), L->n, L->t, idx.data()));
  const size_t rows = 4 * L->m;
  if (o.is_ragged || o.t != L->t || o.rows != rows || o.leaf_index.size() != L->t || o.sibling.size() != L->t || o.auth.size() != L->t)
    return OK;
  int log_n = 0;
  while (((size_t)1 << log_n) < L->n) log_n++;
  for (size_t q = 0; q < L->t; q++)
    if (o.auth[q].size() != (size_t)(log_n - 1)) return OK;
  const size_t flat_elems = L->t * rows;
  Ctx* c = &L->ctx->c;
  DevMem mcols, mdig;  // per-circuit scratch of the lg_ligero object: nothing to release on the early returns below
  LG_TRY(mcols.cached(c, &L->vbuf[V_COLS], flat_elems * sizeof(Fr)));
  LG_TRY(mdig.cached(c, &L->vbuf[V_DIG], L->t * 32));
  Fr* dcols = (Fr*)mcols.p;
  uint8_t* ddig = (uint8_t*)mdig.p;
  std::vector<uint8_t> dig(L->t * 32);
  cudaMemcpyAsync(dcols, o.cols, flat_elems * sizeof(Fr), cudaMemcpyHostToDevice, c->stream);
  int s = hash_column_list(c, dcols, rows, L->t, ddig, L->ctx->col_len_prefix);
  cudaMemcpyAsync(dig.data(), ddig, dig.size(), cudaMemcpyDeviceToHost, c->stream);
  cudaError_t e = cudaStreamSynchronize(c->stream);
  if (cols_keep) {
    cols_keep->p = dcols;
    cols_keep->owned = false;
  }
  if (s != OK) return s;
  if (e != cudaSuccess) return set_error(c, ERR_CUDA, cudaGetErrorString(e));
  for (size_t q = 0; q < L->t; q++) {
    if (o.leaf_index[q] != idx[q]) return OK;
    Digest leaf;
    memcpy(leaf.data(), dig.data() + 32 * q, 32);
    if (!verify_path(L->ctx, root, leaf, o.sibling[q], o.auth[q], idx[q])) return OK;
  }
  *ok = true;
  return OK;
}

// ---- serialisation (arkworks CanonicalSerialize-compatible layout; the reference defines none) -------
// writes into the caller's buffer, or only counts when there is none (one code path for the size query and the copy:
// a 2^24-gate proof is 247 MB)
struct Writer {
  uint8_t* p;
  size_t cap, pos = 0;
  bool ok = true;
  Writer(uint8_t* buf, size_t capacity) : p(buf), cap(capacity) {}
  void raw(const void* src, size_t n) {
    if (p) {
      if (pos + n > cap) ok = false;
      else memcpy(p + pos, src, n);
    }
    pos += n;
  }
  void u64(uint64_t v) { raw(&v, 8); }  // little endian host
  void fr(const Fq& x) {
    if (p) {
      const Fq c = lgh::from_mont(x);
      raw(c.l, 32);
    } else {
      pos += 32;
    }
  }
  void frs(const std::vector<Fq>& v) {
    u64(v.size());
    if (!p) {
      pos += 32 * v.size();
      return;
    }
    for (auto& x : v) fr(x);
  }
  void digest(const Digest& d) {
    u64(32);
    raw(d.data(), 32);
  }
  void frs(const Fq* v, size_t n) {
    u64(n);
    if (!p) {
      pos += 32 * n;
      return;
    }
    for (size_t i = 0; i < n; i++) fr(v[i]);
  }
  void opened(const Opened& o) {
    u64(o.n_columns());
    if (o.is_ragged)
      for (auto& col : o.ragged) frs(col);
    else
      for (size_t q = 0; q < o.t; q++) frs(o.col(q), o.rows);
    u64(o.leaf_index.size());
    for (size_t q = 0; q < o.leaf_index.size(); q++) {  // Path { leaf_sibling_hash, auth_path, leaf_index }
      digest(o.sibling[q]);
      u64(o.auth[q].size());
      for (auto& d : o.auth[q]) digest(d);
      u64(o.leaf_index[q]);
    }
  }
};
struct Reader {
  const uint8_t* p;
  size_t len, pos = 0;
  bool ok = true;
  uint64_t u64() {
    if (pos + 8 > len) {
      ok = false;
      return 0;
    }
    uint64_t v;
    memcpy(&v, p + pos, 8);
    pos += 8;
    return v;
  }
  Fq fr() {
    Fq c = lgh::kZero;
    if (pos + 32 > len) {
      ok = false;
      return c;
    }
    memcpy(c.l, p + pos, 32);
    pos += 32;
    if (lgh::geq_p(c.l)) {
      ok = false;
      return lgh::kZero;
    }
    return lgh::to_mont(c);
  }
  std::vector<Fq> frs() {
    const uint64_t n = u64();
    std::vector<Fq> v;
    if (!ok || n > (len - pos) / 32) {
      ok = false;
      return v;
    }
    v.reserve(n);
    for (uint64_t i = 0; i < n && ok; i++) v.push_back(fr());
    return v;
  }
  Digest digest() {
    Digest d{};
    if (u64() != 32 || pos + 32 > len) {
      ok = false;
      return d;
    }
    memcpy(d.data(), p + pos, 32);
    pos += 32;
    return d;
  }
  Opened opened() {
    Opened o;
    uint64_t nc = u64();
    if (!ok || nc > len) {
      ok = false;
      return o;
    }
    std::vector<std::vector<Fq>> colv;
    for (uint64_t i = 0; i < nc && ok; i++) colv.push_back(frs());
    bool same = true;
    for (auto& cv : colv) same = same && cv.size() == colv[0].size();
    if (same) {
      o.alloc(colv.size(), colv.empty() ? 0 : colv[0].size(), false);
      for (size_t q = 0; q < colv.size(); q++) std::copy(colv[q].begin(), colv[q].end(), o.cols + q * o.rows);
    } else {
      o.is_ragged = true;
      o.ragged = std::move(colv);
    }
    uint64_t np = u64();
    if (!ok || np > len) {
      ok = false;
      return o;
    }
    for (uint64_t i = 0; i < np && ok; i++) {
      o.sibling.push_back(digest());
      uint64_t na = u64();
      if (!ok || na > len) {
        ok = false;
        return o;
      }
      std::vector<Digest> a;
      for (uint64_t j = 0; j < na && ok; j++) a.push_back(digest());
      o.auth.push_back(a);
      o.leaf_index.push_back(u64());
    }
    return o;
  }
};

}  // namespace

extern "C" {

// ---------------------------------------------------------------------------------------------------
// ArithmeticCircuit
// ---------------------------------------------------------------------------------------------------
int lg_circuit_new(lg_circuit** out) {
  if (!out) return ERR_INVALID;
  *out = new (std::nothrow) lg_circuit();
  return *out ? OK : ERR_NOMEM;
}
int lg_circuit_free(lg_circuit* c) {
  delete c;
  return OK;
}
const char* lg_circuit_last_error(const lg_circuit* c) { return c ? c->error.c_str() : "null circuit"; }

static int lg_circuit_constant_impl(lg_circuit* c, const uint64_t value[4], size_t* index_out) {  // mod.rs:76-84
  if (!c || !value) return ERR_INVALID;
  if (lgh::geq_p(value)) {
    c->error = "constant is not a canonical Montgomery representative (limbs >= r)";
    return ERR_INVALID;
  }
  Fq v;
  memcpy(v.l, value, 32);
  auto it = c->constants.find(v);
  size_t idx;
  if (it != c->constants.end()) {
    idx = it->second;
  } else {
    c->const_values.push_back(v);
    idx = c->push({N_CONST, c->const_values.size() - 1, 0});
    c->constants[v] = idx;
  }
  if (index_out) *index_out = idx;
  return OK;
}
int lg_circuit_constant(lg_circuit* c, const uint64_t value[4], size_t* index_out) {
  return lg::guard([&]() { return lg_circuit_constant_impl(c, value, index_out); });
}

static int lg_circuit_new_variable_impl(lg_circuit* c, const char* label, size_t* index_out) {  // mod.rs:92-109
  if (!c) return ERR_INVALID;
  const std::string lab = label ? std::string(label) : "var_" + std::to_string(c->variables.size());
  if (c->variables.count(lab)) {
    c->error = "Variable label already in use: " + lab;
    return ERR_INVALID;
  }
  c->labels.push_back(lab);
  const size_t idx = c->push({N_VAR, c->labels.size() - 1, 0});
  c->variables[lab] = idx;
  if (index_out) *index_out = idx;
  return OK;
}
int lg_circuit_new_variable(lg_circuit* c, const char* label, size_t* index_out) {
  return lg::guard([&]() { return lg_circuit_new_variable_impl(c, label, index_out); });
}

int lg_circuit_get_variable(const lg_circuit* c, const char* label, size_t* index_out) {
  if (!c || !label || !index_out) return ERR_INVALID;
  auto it = c->variables.find(label);
  if (it == c->variables.end()) return ERR_INVALID;  // "Variable not in circuit"
  *index_out = it->second;
  return OK;
}

static int push_gate(lg_circuit* c, uint8_t type, size_t l, size_t r, size_t* index_out) {  // mod.rs:125-145
  if (!c) return ERR_INVALID;
  if (l >= c->nodes.size() || r >= c->nodes.size()) {
    c->error = type == N_ADD ? "operand to Add not in circuit" : "operand to Mul not in circuit";
    return ERR_INVALID;
  }
  const size_t idx = c->push({type, l, r});
  if (index_out) *index_out = idx;
  return OK;
}
int lg_circuit_add(lg_circuit* c, size_t l, size_t r, size_t* index_out) { return push_gate(c, N_ADD, l, r, index_out); }
int lg_circuit_mul(lg_circuit* c, size_t l, size_t r, size_t* index_out) { return push_gate(c, N_MUL, l, r, index_out); }

int lg_circuit_counts(const lg_circuit* c, size_t* nodes, size_t* constants, size_t* variables, size_t* gates) {
  if (!c) return ERR_INVALID;
  if (nodes) *nodes = c->nodes.size();
  if (constants) *constants = c->constants.size();
  if (variables) *variables = c->variables.size();
  if (gates) {
    size_t g = 0;
    for (auto& n : c->nodes) g += (n.type == N_ADD || n.type == N_MUL);
    *gates = g;
  }
  return OK;
}

int lg_circuit_node(const lg_circuit* c, size_t index, int* type, size_t* left, size_t* right, uint64_t value[4]) {
  if (!c || index >= c->nodes.size()) return ERR_INVALID;
  const Node& n = c->nodes[index];
  if (type) *type = n.type;
  if (left) *left = (n.type == N_ADD || n.type == N_MUL) ? n.l : 0;
  if (right) *right = (n.type == N_ADD || n.type == N_MUL) ? n.r : 0;
  if (value && n.type == N_CONST) memcpy(value, c->const_values[n.l].l, 32);
  return OK;
}

// evaluation_trace_multioutput + evaluate_multioutput (mod.rs:325-400): values of the output nodes
static int lg_circuit_evaluate_impl(const lg_circuit* c, const size_t* var_idx, const uint64_t* var_vals, size_t n_vars, const size_t* outputs,
                        size_t n_outputs, uint64_t* out_vals, size_t* n_values_out) {
  if (!c || (n_vars && (!var_idx || !var_vals)) || !outputs || !out_vals) return ERR_INVALID;
  if (!limbs_canonical(var_vals, n_vars)) return ERR_INVALID;
  std::vector<std::pair<size_t, Fq>> vars(n_vars);
  for (size_t i = 0; i < n_vars; i++) {
    vars[i].first = var_idx[i];
    memcpy(vars[i].second.l, var_vals + 4 * i, 32);
  }
  std::vector<Fq> vals;
  std::vector<uint8_t> set;
  std::string err;
  const std::vector<size_t> outs(outputs, outputs + n_outputs);
  int s = evaluation_trace(*c, vars, outs, vals, set, err);
  if (s != OK) {
    const_cast<lg_circuit*>(c)->error = err;
    return s;
  }
  // the reference filters the trace by membership in `outputs` (mod.rs:384-389): values come in NODE order, one per
  // distinct output node
  std::vector<size_t> sorted(outs);
  std::sort(sorted.begin(), sorted.end());
  sorted.erase(std::unique(sorted.begin(), sorted.end()), sorted.end());
  for (size_t i = 0; i < sorted.size(); i++) memcpy(out_vals + 4 * i, vals[sorted[i]].l, 32);
  if (n_values_out) *n_values_out = sorted.size();
  return OK;
}
int lg_circuit_evaluate(const lg_circuit* c, const size_t* var_idx, const uint64_t* var_vals, size_t n_vars, const size_t* outputs,
                        size_t n_outputs, uint64_t* out_vals, size_t* n_values_out) {
  return lg::guard([&]() { return lg_circuit_evaluate_impl(c, var_idx, var_vals, n_vars, outputs, n_outputs, out_vals, n_values_out); });
}

// from_constraint_system (mod.rs:455-520).  Matrices as ConstraintSystem::to_matrices yields them: CSR with
// coefficient Fr values and column indices (column 0 = the constant one), n_cols = instance + witness variables.
static int lg_circuit_from_r1cs_impl(size_t n_constraints, size_t n_cols, const uint64_t* const row_ptr[3], const uint64_t* const col_idx[3],
                         const uint64_t* const coeffs[3], lg_circuit** out, size_t* outputs) {
  if (!out || !outputs || !row_ptr || !col_idx || !coeffs || n_cols == 0) return ERR_INVALID;
  lg_circuit* c = new (std::nothrow) lg_circuit();
  if (!c) return ERR_NOMEM;
  size_t one;
  lg_circuit_constant(c, lgh::kOne.l, &one);
  for (size_t i = 1; i < n_cols; i++) lg_circuit_new_variable(c, nullptr, nullptr);
  std::vector<size_t> rows[3];
  for (int mtx = 0; mtx < 3; mtx++) {
    for (size_t r = 0; r < n_constraints; r++) {  // compile_sparse_scalar_product (501-520)
      std::vector<std::pair<size_t, size_t>> consts;
      for (uint64_t e = row_ptr[mtx][r]; e < row_ptr[mtx][r + 1]; e++) {
        size_t ci;
        lg_circuit_constant(c, coeffs[mtx] + 4 * e, &ci);
        consts.push_back({ci, (size_t)col_idx[mtx][e]});
      }
      if (consts.empty()) {  // add_nodes(empty).reduce().unwrap() panics in the reference
        delete c;
        return ERR_INVALID;
      }
      std::vector<size_t> products;
      for (auto& cv : consts) {
        if (cv.second >= n_cols) {
          delete c;
          return ERR_INVALID;
        }
        if (cv.first == 0 || cv.second == 0) products.push_back(cv.first + cv.second);
        else {
          size_t p;
          lg_circuit_mul(c, cv.first, cv.second, &p);
          products.push_back(p);
        }
      }
      size_t acc = products[0];
      for (size_t i = 1; i < products.size(); i++) lg_circuit_add(c, acc, products[i], &acc);
      rows[mtx].push_back(acc);
    }
  }
  std::vector<size_t> ab(n_constraints), mc(n_constraints);
  for (size_t r = 0; r < n_constraints; r++) lg_circuit_mul(c, rows[0][r], rows[1][r], &ab[r]);
  size_t minus_one;
  const Fq m1 = lgh::neg(lgh::kOne);
  lg_circuit_constant(c, m1.l, &minus_one);
  for (size_t r = 0; r < n_constraints; r++) lg_circuit_mul(c, rows[2][r], minus_one, &mc[r]);
  for (size_t r = 0; r < n_constraints; r++) {
    size_t a1, a2;
    lg_circuit_add(c, ab[r], mc[r], &a1);
    lg_circuit_add(c, a1, one, &a2);
    outputs[r] = a2;
  }
  *out = c;
  return OK;
}
int lg_circuit_from_r1cs(size_t n_constraints, size_t n_cols, const uint64_t* const row_ptr[3], const uint64_t* const col_idx[3],
                         const uint64_t* const coeffs[3], lg_circuit** out, size_t* outputs) {
  return lg::guard([&]() { return lg_circuit_from_r1cs_impl(n_constraints, n_cols, row_ptr, col_idx, coeffs, out, outputs); });
}

// read_constraint_system's R1CS half (src/reader.rs:6-19): the iden3 .r1cs v1 container -> the three
// matrices as ConstraintSystem::to_matrices yields them (terms of a row merged per wire, zero coefficients dropped, wires
// ascending, as ark-relations' sorted LinearCombination does) -> from_constraint_system.  The witness half of the
// reference (a wasm witness calculator run by ark-circom) is out of scope: callers pass the assignment.
static int lg_circuit_from_r1cs_bytes_impl(const uint8_t* data, size_t len, lg_circuit** out, size_t* outputs, size_t outputs_cap,
                               size_t* n_constraints_out, size_t* n_wires_out) {
  if (!data || !out) return ERR_INVALID;
  auto rd32 = [&](size_t off, uint32_t& v) {
    if (off + 4 > len) return false;
    memcpy(&v, data + off, 4);
    return true;
  };
  auto rd64 = [&](size_t off, uint64_t& v) {
    if (off + 8 > len) return false;
    memcpy(&v, data + off, 8);
    return true;
  };
  uint32_t version = 0, nsec = 0;
  if (len < 12 || memcmp(data, "r1cs", 4) != 0 || !rd32(4, version) || !rd32(8, nsec) || version != 1) return ERR_INVALID;
  size_t off = 12, hdr_off = 0, hdr_len = 0, body_off = 0, body_len = 0, map_len = 0;
  bool have_map = false;
  for (uint32_t sct = 0; sct < nsec; sct++) {
    uint32_t typ;
    uint64_t size;
    if (!rd32(off, typ) || !rd64(off + 4, size) || size > len - off - 12) return ERR_INVALID;
    off += 12;
    if (typ == 1) {
      hdr_off = off;
      hdr_len = size;
    } else if (typ == 2) {
      body_off = off;
      body_len = size;
    } else if (typ == 3) {  // wire -> label map: 8 bytes per wire
      have_map = true;
      map_len = size;
    }
    off += size;
  }
  if (!hdr_off || !body_off) return ERR_INVALID;
  uint32_t fs = 0, n_wires = 0, n_constraints = 0;
  if (hdr_len < 4 || !rd32(hdr_off, fs) || fs != 32 || hdr_len < 4 + 32 + 16 + 8 + 4) return ERR_INVALID;
  if (memcmp(data + hdr_off + 4, lgh::kP, 32) != 0) return ERR_UNSUPPORTED;  // only BN254 Fr
  rd32(hdr_off + 4 + 32, n_wires);
  rd32(hdr_off + 4 + 32 + 16 + 8, n_constraints);
  if (n_wires == 0 || n_constraints == 0) return ERR_INVALID;
  // every count that drives an allocation must be justified by the file: a constraint takes at least 12 bytes of the
  // body, and a wire 8 bytes of the wire-to-label section (files without that section: one byte of file per wire)
  if ((size_t)n_constraints > body_len / 12) return ERR_INVALID;
  if ((size_t)n_wires > (have_map ? map_len / 8 : len)) return ERR_INVALID;
  std::vector<uint64_t> row_ptr[3], col_idx[3];
  std::vector<Fq> coeff[3];
  size_t o = body_off;
  const size_t end = body_off + body_len;
  for (int mtx = 0; mtx < 3; mtx++) row_ptr[mtx].push_back(0);
  std::map<uint32_t, Fq> acc;
  for (uint32_t r = 0; r < n_constraints; r++)
    for (int mtx = 0; mtx < 3; mtx++) {
      uint32_t nt;
      if (o + 4 > end || !rd32(o, nt)) return ERR_INVALID;
      o += 4;
      if ((size_t)nt * 36 > end - o) return ERR_INVALID;
      acc.clear();
      for (uint32_t e = 0; e < nt; e++, o += 36) {
        uint32_t w;
        rd32(o, w);
        if (w >= n_wires) return ERR_INVALID;
        Fq raw;
        memcpy(raw.l, data + o + 4, 32);
        while (lgh::geq_p(raw.l)) lgh::sub_p(raw.l);  // coefficients are canonical in practice; reduce anything else
        const Fq c = lgh::to_mont(raw);
        auto it = acc.find(w);
        if (it == acc.end()) acc[w] = c;
        else it->second = lgh::add(it->second, c);
      }
      for (const auto& kv : acc) {
        if (kv.second.is_zero()) continue;
        col_idx[mtx].push_back(kv.first);
        coeff[mtx].push_back(kv.second);
      }
      row_ptr[mtx].push_back(col_idx[mtx].size());
    }
  if (n_constraints_out) *n_constraints_out = n_constraints;
  if (n_wires_out) *n_wires_out = n_wires;
  if (!outputs || outputs_cap < n_constraints) return ERR_INVALID;
  const uint64_t* rp[3] = {row_ptr[0].data(), row_ptr[1].data(), row_ptr[2].data()};
  const uint64_t* ci[3] = {col_idx[0].data(), col_idx[1].data(), col_idx[2].data()};
  const uint64_t* cf[3] = {(const uint64_t*)coeff[0].data(), (const uint64_t*)coeff[1].data(), (const uint64_t*)coeff[2].data()};
  return lg_circuit_from_r1cs(n_constraints, n_wires, rp, ci, cf, out, outputs);
}
int lg_circuit_from_r1cs_bytes(const uint8_t* data, size_t len, lg_circuit** out, size_t* outputs, size_t outputs_cap,
                               size_t* n_constraints_out, size_t* n_wires_out) {
  return lg::guard([&]() { return lg_circuit_from_r1cs_bytes_impl(data, len, out, outputs, outputs_cap, n_constraints_out, n_wires_out); });
}

// Seeded random Add/Mul circuit of exactly `gates` gates for the synthetic configurations: two input
// variables, gate type by a fair coin, operands drawn uniformly from all earlier non-constant nodes (depth O(log gates)
// with overwhelming probability), every node feeds the single output, the output is an Add gate of value 1
// (Add(sum, constant 1 - sum)), no gate has two constant operands.  sol_len of the LigeroCircuit is gates + 4.
// Generator: splitmix64(seed); not part of the reference (its tests build circuits by hand).
static int lg_circuit_synthetic_impl(size_t gates, uint64_t seed, lg_circuit** out, size_t* output, size_t var_idx[2], uint64_t var_vals[8]) {
  if (!out || !output || !var_idx || !var_vals || gates < 4) return ERR_INVALID;
  lg_circuit* c = new (std::nothrow) lg_circuit();
  if (!c) return ERR_NOMEM;
  uint64_t st = seed;
  auto next = [&]() {
    uint64_t z = (st += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
  };
  auto below = [&](uint64_t n) { return (uint64_t)(((unsigned __int128)next() * n) >> 64); };
  size_t one;
  lg_circuit_constant(c, lgh::kOne.l, &one);
  std::vector<Fq> vals;      // value of every node
  std::vector<uint8_t> used;  // node is an operand of a later gate
  vals.reserve(gates + 8);
  used.reserve(gates + 8);
  vals.push_back(lgh::kOne);
  used.push_back(1);
  for (int v = 0; v < 2; v++) {
    Fq x;
    for (int j = 0; j < 4; j++) x.l[j] = next();
    x.l[3] &= (1ULL << 60) - 1;  // < 2^252 < r: a valid Montgomery residue
    lg_circuit_new_variable(c, nullptr, &var_idx[v]);
    memcpy(var_vals + 4 * v, x.l, 32);
    vals.push_back(x);
    used.push_back(0);
  }
  size_t n_gates = 0, unused = 2;
  auto gate = [&](bool mul, size_t l, size_t r) {
    size_t idx;
    push_gate(c, mul ? N_MUL : N_ADD, l, r, &idx);
    vals.push_back(mul ? lgh::mul(vals[l], vals[r]) : lgh::add(vals[l], vals[r]));
    used.push_back(0);
    for (size_t o : {l, r})
      if (!used[o]) {
        used[o] = 1;
        unused--;
      }
    unused++;
    n_gates++;
    return idx;
  };
  // random part: afterwards a balanced sum over the `unused` nodes takes unused - 1 Add gates, then one closing Add
  while (n_gates + unused < gates) {
    const size_t hi = vals.size();
    size_t l = 1 + below(hi - 1), r = 1 + below(hi - 1);
    if (n_gates + unused + 1 == gates) {
      // exactly one more is wanted: a gate that consumes exactly one unused node (total grows by 1, not by 2)
      size_t u = hi - 1;  // the newest node is always unused
      size_t v = 1;
      while (v < hi && (!used[v] || v == u)) v++;
      l = u;
      r = v;
    }
    gate((next() & 1) != 0, l, r);
  }
  std::vector<size_t> level;
  for (size_t i = 1; i < vals.size(); i++)
    if (!used[i]) level.push_back(i);
  while (level.size() > 1) {
    std::vector<size_t> nxt;
    for (size_t i = 0; i + 1 < level.size(); i += 2) nxt.push_back(gate(false, level[i], level[i + 1]));
    if (level.size() & 1) nxt.push_back(level.back());
    level.swap(nxt);
  }
  const Fq k = lgh::sub(lgh::kOne, vals[level[0]]);
  size_t kc;
  lg_circuit_constant(c, k.l, &kc);
  size_t o;
  push_gate(c, N_ADD, level[0], kc, &o);
  *output = o;
  *out = c;
  return OK;
}
int lg_circuit_synthetic(size_t gates, uint64_t seed, lg_circuit** out, size_t* output, size_t var_idx[2], uint64_t var_vals[8]) {
  return lg::guard([&]() { return lg_circuit_synthetic_impl(gates, seed, out, output, var_idx, var_vals); });
}

// ---------------------------------------------------------------------------------------------------
// sponge
// ---------------------------------------------------------------------------------------------------
static int lg_sponge_new_impl(int full_rounds, int partial_rounds, uint64_t alpha, const uint64_t* mds, const uint64_t* ark, int rate, int capacity,
                  lg_sponge** out) {
  if (!out || !mds || !ark || rate < 1 || capacity < 0 || full_rounds < 0 || partial_rounds < 0) return ERR_INVALID;
  lgh::PoseidonConfig cfg;
  cfg.full_rounds = full_rounds;
  cfg.partial_rounds = partial_rounds;
  cfg.alpha = alpha;
  cfg.rate = rate;
  cfg.capacity = capacity;
  const size_t t = (size_t)rate + capacity;
  cfg.mds.resize(t * t);
  cfg.ark.resize((size_t)(full_rounds + partial_rounds) * t);
  memcpy(cfg.mds.data(), mds, cfg.mds.size() * 32);
  memcpy(cfg.ark.data(), ark, cfg.ark.size() * 32);
  *out = new (std::nothrow) lg_sponge(cfg);
  return *out ? OK : ERR_NOMEM;
}
int lg_sponge_new(int full_rounds, int partial_rounds, uint64_t alpha, const uint64_t* mds, const uint64_t* ark, int rate, int capacity,
                  lg_sponge** out) {
  return lg::guard([&]() { return lg_sponge_new_impl(full_rounds, partial_rounds, alpha, mds, ark, rate, capacity, out); });
}

// ark_poly_commit::test_sponge() with the deterministic ark_std::test_rng() (ChaCha12 from the fixed seed)
static int lg_sponge_test_impl(lg_sponge** out) {
  if (!out) return ERR_INVALID;
  const uint8_t seed[32] = {1, 0, 0, 0, 23, 0, 0, 0, 200, 1, 0, 0, 210, 30, 0, 0};
  ChaChaRng rng(seed, 12);
  lgh::PoseidonConfig cfg;
  cfg.full_rounds = 8;
  cfg.partial_rounds = 31;
  cfg.alpha = 17;
  cfg.rate = 2;
  cfg.capacity = 1;
  const Fq one = lgh::kOne, zero = lgh::kZero;
  cfg.mds = {one, zero, one, one, one, zero, zero, one, one};
  for (int i = 0; i < (8 + 31) * 3; i++) {
    const Fr e = fr_rand(rng);
    Fq q;
    memcpy(q.l, e.v, 32);
    cfg.ark.push_back(q);
  }
  *out = new (std::nothrow) lg_sponge(cfg);
  return *out ? OK : ERR_NOMEM;
}
int lg_sponge_test(lg_sponge** out) {
  return lg::guard([&]() { return lg_sponge_test_impl(out); });
}
int lg_sponge_clone(const lg_sponge* s, lg_sponge** out) {
  if (!s || !out) return ERR_INVALID;
  *out = new (std::nothrow) lg_sponge(*s);
  return *out ? OK : ERR_NOMEM;
}
int lg_sponge_free(lg_sponge* s) {
  delete s;
  return OK;
}
static int lg_sponge_absorb_bytes_impl(lg_sponge* s, const uint8_t* data, size_t len) {
  if (!s || (!data && len)) return ERR_INVALID;
  s->s.absorb_bytes(data, len);
  return OK;
}
int lg_sponge_absorb_bytes(lg_sponge* s, const uint8_t* data, size_t len) {
  return lg::guard([&]() { return lg_sponge_absorb_bytes_impl(s, data, len); });
}
static int lg_sponge_absorb_fr_impl(lg_sponge* s, const uint64_t* elems, size_t count) {
  if (!s || (!elems && count)) return ERR_INVALID;
  std::vector<Fq> v(count);
  memcpy(v.data(), elems, count * 32);
  s->s.absorb_field(v);
  return OK;
}
int lg_sponge_absorb_fr(lg_sponge* s, const uint64_t* elems, size_t count) {
  return lg::guard([&]() { return lg_sponge_absorb_fr_impl(s, elems, count); });
}
static int lg_sponge_squeeze_bytes_impl(lg_sponge* s, uint8_t* out, size_t len) {
  if (!s || !out) return ERR_INVALID;
  const std::vector<uint8_t> b = s->s.squeeze_bytes(len);
  memcpy(out, b.data(), len);
  return OK;
}
int lg_sponge_squeeze_bytes(lg_sponge* s, uint8_t* out, size_t len) {
  return lg::guard([&]() { return lg_sponge_squeeze_bytes_impl(s, out, len); });
}

// ---------------------------------------------------------------------------------------------------
// LigeroCircuit
// ---------------------------------------------------------------------------------------------------
static int lg_ligero_new_impl(lg_ctx* ctx, const lg_circuit* circuit, const size_t* outputs, size_t n_outputs, size_t lambda, lg_ligero** out) {
  if (!ctx || !circuit || !out || (!outputs && n_outputs)) return ERR_INVALID;
  if (circuit->nodes.empty()) return fail(ctx, ERR_INVALID, "empty circuit");
  lg_ligero* L = new (std::nothrow) lg_ligero();
  if (!L) return ERR_NOMEM;
  L->ctx = ctx;
  lg_circuit& c = L->circuit;  // the formatted copy; its nodes are written below, already in their final places
  c.const_values = circuit->const_values;
  c.labels = circuit->labels;
  c.constants = circuit->constants;
  c.variables = circuit->variables;
  c.error = circuit->error;
  auto it = c.constants.find(lgh::kOne);
  if (it != c.constants.end()) {
    L->one_index = it->second;
    L->one_found = true;
  } else {
    L->one_index = 1;
    L->one_found = false;
  }
  const size_t oi = L->one_index;
  const bool of = L->one_found;
  Node one_node{N_CONST, 0, 0};
  if (oi != 0) {  // insert_one (mod.rs:244-271)
    if (of) {
      // keep the existing const_values slot of the constant 1
      for (size_t j = 0; j < c.const_values.size(); j++)
        if (c.const_values[j] == lgh::kOne) one_node.l = j;
    } else {
      c.const_values.push_back(lgh::kOne);
      one_node.l = c.const_values.size() - 1;
    }
  }
  // the caller's constant 1 (if any) removed, the constant 1 in front, the operands of every gate bumped: one pass
  lgh::format_nodes(circuit->nodes.data(), circuit->nodes.size(), oi, of, one_node, c.nodes, lgh::host_threads());
  if (oi != 0) {
    for (auto& kv : c.constants) kv.second = bump_index(oi, of, kv.second);
    c.constants[lgh::kOne] = 0;
    for (auto& kv : c.variables) kv.second = bump_index(oi, of, kv.second);
  }
  L->sol_len = 1 + c.nodes.size() - c.constants.size() + n_outputs;  // mod.rs:171
  L->m = (size_t)std::ceil(std::sqrt((double)L->sol_len));             // compute_dimensions 275-279
  size_t k = 1;
  while (k < L->m) k <<= 1;
  L->k = k;
  L->n = 8 * k;                                                        // reed_solomon_parameters 283-294
  L->t = calculate_t(lambda, L->n - k + 1, L->n, L->n);
  if (k < 2) {
    delete L;
    return fail(ctx, ERR_UNSUPPORTED, "k < 2 is not supported by the device path");
  }
  for (size_t i = 0; i < n_outputs; i++) {
    const size_t o = bump_index(oi, of, outputs[i]);
    if (o >= c.nodes.size()) {
      delete L;
      return fail(ctx, ERR_INVALID, "output node not in circuit");
    }
    L->outputs.push_back(o);
  }
  if (c.nodes.size() >= 0x7fffffffu) {
    delete L;
    return fail(ctx, ERR_UNSUPPORTED, "circuits of 2^31 nodes or more are not supported");
  }
  std::string err;
  const double t_c00 = now_ms();
  lgh::NodeArrays arr;  // type / left / right as flat arrays: what every pass below reads instead of the 24-byte nodes
  lgh::pack_nodes(c.nodes.data(), c.nodes.size(), arr, lgh::host_threads());
  const double t_c0 = now_ms();
  int s = build_constraints(L, arr, err);
  const double t_c1 = now_ms();
  if (s != OK) {
    if (!err.empty()) fail(ctx, s, err);
    lg_ligero_free(L);
    return s;
  }
  s = build_trace(L, arr);
  if (debug_timing())
    fprintf(stderr, "[lg] LigeroCircuit::new: node arrays %.1f ms, constraint matrix %.1f ms, trace schedule %.1f ms (%zu nodes, %d host threads)\n",
            t_c0 - t_c00, t_c1 - t_c0, now_ms() - t_c1, c.nodes.size(), lgh::host_threads());
  if (s != OK) {
    lg_ligero_free(L);
    return s;
  }
  *out = L;
  return OK;
}
int lg_ligero_new(lg_ctx* ctx, const lg_circuit* circuit, const size_t* outputs, size_t n_outputs, size_t lambda, lg_ligero** out) {
  return lg::guard([&]() { return lg_ligero_new_impl(ctx, circuit, outputs, n_outputs, lambda, out); });
}

int lg_ligero_release_buffers(lg_ligero* L) {
  if (!L) return ERR_INVALID;
  cudaSetDevice(L->ctx->c.device);
  if (L->u_cache) lg_matrix_free(L->u_cache);
  L->u_cache = nullptr;
  if (L->pre_cache) {
    cudaStreamSynchronize(L->ctx->c.stream);
    cudaFree(L->pre_cache);
  }
  L->pre_cache = nullptr;
  for (void*& b : L->vbuf) {
    if (b) cudaFree(b);
    b = nullptr;
  }
  return OK;
}

int lg_ligero_free(lg_ligero* L) {
  if (!L) return OK;
  lg_ligero_release_buffers(L);
  if (L->a) lg_constraints_free(L->a);
  cudaSetDevice(L->ctx->c.device);
  lg::trace_free(L->trace);
  delete L;
  return OK;
}

int lg_ligero_set_trace_mode(lg_ligero* L, int mode) {
  if (!L || mode < -1 || mode > 1) return ERR_INVALID;
  L->trace_mode = mode;
  return OK;
}

int lg_ligero_prove_ms(const lg_ligero* L, double ms_out[7]) {
  if (!L || !ms_out) return ERR_INVALID;
  for (int i = 0; i < 7; i++) ms_out[i] = L->prove_ms[i];
  return OK;
}

int lg_ligero_trace_info(const lg_ligero* L, size_t* gates, size_t* levels, size_t* launches, int* on_device) {
  if (!L) return ERR_INVALID;
  if (gates) *gates = L->trace.n_gates;
  if (levels) *levels = L->trace.n_levels;
  if (launches) *launches = L->trace.segments.size();
  if (on_device) *on_device = trace_on_device(L) ? 1 : 0;
  return OK;
}

// a1 on the device: evaluation trace by levels + scatter into [X;Y;Z;W], all in HBM (out_dev: Fr[4*m*k], device)
static int lg_ligero_witness_matrix_dev_impl(lg_ligero* L, const size_t* var_idx, const uint64_t* var_vals, size_t n_vars, int bump,
                                 uint64_t* out_dev) {
  if (!L || !out_dev || (n_vars && (!var_idx || !var_vals))) return ERR_INVALID;
  if (!limbs_canonical(var_vals, n_vars)) return fail(L->ctx, ERR_INVALID, "variable values must be canonical Montgomery limbs (< r)");
  if (!lg::is_device_ptr(out_dev)) return fail(L->ctx, ERR_INVALID, "lg_ligero_witness_matrix_dev needs a device buffer");
  const lg_circuit& c = L->circuit;
  const size_t N = c.nodes.size();
  std::map<uint32_t, Fq> given;  // a repeated index keeps its last value, as the host evaluator does
  for (size_t i = 0; i < n_vars; i++) {
    const size_t idx = bump ? bump_index(L->one_index, L->one_found, var_idx[i]) : var_idx[i];
    if (idx >= N || c.nodes[idx].type != N_VAR) return fail(L->ctx, ERR_INVALID, "Value supplied for non-variable node");
    Fq v;
    memcpy(v.l, var_vals + 4 * i, 32);
    given[(uint32_t)idx] = v;
  }
  bool missing_unreached = false;
  for (const uint32_t i : L->var_nodes)
    if (!given.count(i)) {
      if (L->reach[i]) return fail(L->ctx, ERR_INVALID, "Uninitialised variable");
      missing_unreached = true;
    }
  if (missing_unreached || !L->all_gates_reach)
    return fail(L->ctx, ERR_INVALID,
                "Uninitialised variable. Make sure the circuit only contains nodes upon which the final output truly depends");
  lg::Ctx* cx = &L->ctx->c;
  cudaSetDevice(cx->device);
  std::vector<uint32_t> vnode, vpos;
  std::vector<Fq> vval;
  for (const auto& kv : given) {
    vnode.push_back(kv.first);
    vpos.push_back(L->index_map[kv.first]);
    vval.push_back(kv.second);
  }
  const size_t nv = vnode.size();
  uint8_t* buf = nullptr;
  if (nv) {
    LG_CUDA(cx, cudaMalloc((void**)&buf, nv * (8 + sizeof(Fq))));
    LG_CUDA(cx, cudaMemcpyAsync(buf, vval.data(), nv * sizeof(Fq), cudaMemcpyHostToDevice, cx->stream));
    LG_CUDA(cx, cudaMemcpyAsync(buf + nv * sizeof(Fq), vnode.data(), nv * 4, cudaMemcpyHostToDevice, cx->stream));
    LG_CUDA(cx, cudaMemcpyAsync(buf + nv * (sizeof(Fq) + 4), vpos.data(), nv * 4, cudaMemcpyHostToDevice, cx->stream));
  }
  int s = lg::trace_run(cx, L->trace, (const uint32_t*)(buf + nv * sizeof(Fq)), (const uint32_t*)(buf + nv * (sizeof(Fq) + 4)),
                        (const lg::Fr*)buf, nv, (lg::Fr*)out_dev);
  if (buf) {
    cudaStreamSynchronize(cx->stream);  // the staging vectors above go out of scope
    cudaFree(buf);
  }
  return s;
}
int lg_ligero_witness_matrix_dev(lg_ligero* L, const size_t* var_idx, const uint64_t* var_vals, size_t n_vars, int bump,
                                 uint64_t* out_dev) {
  return lg::guard([&]() { return lg_ligero_witness_matrix_dev_impl(L, var_idx, var_vals, n_vars, bump, out_dev); });
}

int lg_ligero_params(const lg_ligero* L, size_t* m, size_t* k, size_t* n, size_t* t, size_t* sol_len) {
  if (!L) return ERR_INVALID;
  if (m) *m = L->m;
  if (k) *k = L->k;
  if (n) *n = L->n;
  if (t) *t = L->t;
  if (sol_len) *sol_len = L->sol_len;
  return OK;
}

// the pre-encoding matrix [X;Y;Z;W] (mod.rs:476-516); out: Fr[4*m*k] host
static int lg_ligero_witness_matrix_impl(lg_ligero* L, const size_t* var_idx, const uint64_t* var_vals, size_t n_vars, int bump, uint64_t* out) {
  if (!L || !out || (n_vars && (!var_idx || !var_vals))) return ERR_INVALID;
  if (!limbs_canonical(var_vals, n_vars)) return fail(L->ctx, ERR_INVALID, "variable values must be canonical Montgomery limbs (< r)");
  const lg_circuit& c = L->circuit;
  std::vector<std::pair<size_t, Fq>> vars(n_vars);
  for (size_t i = 0; i < n_vars; i++) {
    vars[i].first = bump ? bump_index(L->one_index, L->one_found, var_idx[i]) : var_idx[i];
    memcpy(vars[i].second.l, var_vals + 4 * i, 32);
  }
  std::vector<Fq> sol;
  std::vector<uint8_t> set;
  std::string err;
  int s = evaluation_trace(c, vars, L->outputs, sol, set, err);
  if (s != OK) return fail(L->ctx, s, err);
  for (size_t i = 0; i < sol.size(); i++)
    if (!set[i])
      return fail(L->ctx, ERR_INVALID,
                  "Uninitialised variable. Make sure the circuit only contains nodes upon which the final output truly depends");
  const size_t mk = L->m * L->k;
  Fq* X = (Fq*)out;
  memset(X, 0, 4 * mk * sizeof(Fq));
  Fq *Y = X + mk, *Z = Y + mk, *W = Z + mk;
  size_t pos = 0;
  for (size_t i = 0; i < c.nodes.size(); i++) {
    const Node& nd = c.nodes[i];
    if (nd.type == N_CONST && i != 0) continue;
    if (pos >= mk) return fail(L->ctx, ERR_STATE, "internal: witness longer than m*k");
    W[pos] = sol[i];
    if (nd.type == N_MUL) {
      X[pos] = sol[nd.l];
      Y[pos] = sol[nd.r];
      Z[pos] = sol[i];
    }
    pos++;
  }
  return OK;
}
int lg_ligero_witness_matrix(lg_ligero* L, const size_t* var_idx, const uint64_t* var_vals, size_t n_vars, int bump, uint64_t* out) {
  return lg::guard([&]() { return lg_ligero_witness_matrix_impl(L, var_idx, var_vals, n_vars, bump, out); });
}

// prove_inner on a ready pre-encoding matrix (host or device): the commit-and-test transcript
static int lg_prove_matrix_impl(lg_ligero* L, const uint64_t* preenc_u, lg_sponge* sponge, lg_proof** out) {
  if (!L || !preenc_u || !sponge || !out) return ERR_INVALID;
  lg_ctx* ctx = L->ctx;
  lgh::PoseidonSponge& sp = sponge->s;
  const size_t rows = 4 * L->m, k = L->k;
  lg_proof* P = new (std::nothrow) lg_proof();
  if (!P) return ERR_NOMEM;
  typedef std::chrono::steady_clock Clock;
  const Clock::time_point t_begin = Clock::now();
  Clock::time_point t_last = t_begin;
  double open_ms = 0;
  auto lap = [&]() {
    const Clock::time_point now = Clock::now();
    const double ms = std::chrono::duration<double, std::milli>(now - t_last).count();
    t_last = now;
    return ms;
  };
  lg_matrix* U = L->u_cache;
  int s;
  if (U) {
    s = lg_recommit(U, preenc_u, P->root.data());  // mod.rs:521-551 into the resident buffers
  } else {
    s = lg_commit(ctx, preenc_u, rows, k, 8, P->root.data(), &U);
    if (s == OK) L->u_cache = U;
  }
  auto done = [&](int code) {
    // the three openings' device-to-host copies run on their own stream (open_columns): wait for them before the proof
    // (or its buffers) leaves this call
    lg::Ctx* cx = &ctx->c;
    if (cx->open_stream && cudaStreamSynchronize(cx->open_stream) != cudaSuccess && code == OK)
      code = fail(ctx, ERR_CUDA, "opened-column copy failed");
    if (code != OK) {
      delete P;
    } else {
      L->prove_ms[5] = open_ms;
      L->prove_ms[6] = std::chrono::duration<double, std::milli>(Clock::now() - t_begin).count();
      *out = P;
    }
    return code;
  };
  if (s != OK) return done(s);
  sp.absorb_bytes(P->root.data(), 32);  // 560
  L->prove_ms[1] = lap();
  // Test-Interleaved (646-669)
  std::vector<uint8_t> seed = sp.squeeze_bytes(32);
  std::vector<Fq> r(rows);
  if ((s = lg_expand_fr(ctx, seed.data(), rows, (uint64_t*)r.data())) != OK) return done(s);
  P->preenc_u_lc.resize(k);
  if ((s = lg_row_combine(U, (const uint64_t*)r.data(), (uint64_t*)P->preenc_u_lc.data())) != OK) return done(s);
  sp.absorb_field(P->preenc_u_lc);
  L->prove_ms[2] = lap();
  if ((s = open_columns(L, U, sp, P->interleaved)) != OK) return done(s);
  open_ms += lap();
  // Test-Linear-Constraints (712-747)
  seed = sp.squeeze_bytes(32);
  std::vector<Fq> poly(2 * k);
  size_t len = 0;
  if ((s = lg_linear_test_seeded(U, L->a, seed.data(), (uint64_t*)poly.data(), &len)) != OK) return done(s);
  P->linear_poly.assign(poly.begin(), poly.begin() + len);
  sp.absorb_field(P->linear_poly);
  L->prove_ms[3] = lap();
  if ((s = open_columns(L, U, sp, P->linear)) != OK) return done(s);
  open_ms += lap();
  // Test-Quadratic-Constraints (832-859)
  seed = sp.squeeze_bytes(32);
  std::vector<Fq> rq(L->m);
  if ((s = lg_expand_fr(ctx, seed.data(), L->m, (uint64_t*)rq.data())) != OK) return done(s);
  if ((s = lg_quadratic_test(U, (const uint64_t*)rq.data(), (uint64_t*)poly.data(), &len)) != OK) return done(s);
  P->quadratic_poly.assign(poly.begin(), poly.begin() + len);
  sp.absorb_field(P->quadratic_poly);
  L->prove_ms[4] = lap();
  if ((s = open_columns(L, U, sp, P->quadratic)) != OK) return done(s);
  open_ms += lap();
  return done(OK);
}
int lg_prove_matrix(lg_ligero* L, const uint64_t* preenc_u, lg_sponge* sponge, lg_proof** out) {
  return lg::guard([&]() { return lg_prove_matrix_impl(L, preenc_u, sponge, out); });
}

// LigeroCircuit::prove (bump = 1: indices refer to the caller's circuit) / prove_inner (bump = 0)
static int lg_prove_impl(lg_ligero* L, const size_t* var_idx, const uint64_t* var_vals, size_t n_vars, int bump, lg_sponge* sponge, lg_proof** out) {
  if (!L || !sponge || !out) return ERR_INVALID;
  const auto t0 = std::chrono::steady_clock::now();
  auto trace_ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); };
  int s;
  if (trace_on_device(L)) {  // wide circuit: trace + layout in HBM, nothing but the variables crosses PCIe
    lg::Ctx* cx = &L->ctx->c;
    cudaSetDevice(cx->device);
    if (!L->pre_cache) LG_CUDA(cx, cudaMalloc((void**)&L->pre_cache, 4 * L->m * L->k * sizeof(Fq)));
    uint64_t* pre_dev = L->pre_cache;
    s = lg_ligero_witness_matrix_dev(L, var_idx, var_vals, n_vars, bump, pre_dev);
    double tr = 0;
    if (s == OK) {
      cudaStreamSynchronize(cx->stream);
      tr = trace_ms();
      s = lg_prove_matrix(L, pre_dev, sponge, out);
    }
    if (s == OK) {
      L->prove_ms[0] = tr;
      L->prove_ms[6] += tr;
    }
    return s;
  }
  std::vector<Fq> pre(4 * L->m * L->k);
  LG_TRY(lg_ligero_witness_matrix(L, var_idx, var_vals, n_vars, bump, (uint64_t*)pre.data()));
  const double tr = trace_ms();
  s = lg_prove_matrix(L, (const uint64_t*)pre.data(), sponge, out);
  if (s == OK) {
    L->prove_ms[0] = tr;
    L->prove_ms[6] += tr;
  }
  return s;
}
int lg_prove(lg_ligero* L, const size_t* var_idx, const uint64_t* var_vals, size_t n_vars, int bump, lg_sponge* sponge, lg_proof** out) {
  return lg::guard([&]() { return lg_prove_impl(L, var_idx, var_vals, n_vars, bump, sponge, out); });
}

static int lg_prove_with_labels_impl(lg_ligero* L, const char* const* labels, const uint64_t* var_vals, size_t n_vars, lg_sponge* sponge,
                         lg_proof** out) {
  if (!L || (!labels && n_vars)) return ERR_INVALID;
  std::vector<size_t> idx(n_vars);
  for (size_t i = 0; i < n_vars; i++) {
    auto it = L->circuit.variables.find(labels[i]);
    if (it == L->circuit.variables.end()) return fail(L->ctx, ERR_INVALID, std::string("Variable not found: ") + labels[i]);
    idx[i] = it->second;
  }
  return lg_prove(L, idx.data(), var_vals, n_vars, 0, sponge, out);
}
int lg_prove_with_labels(lg_ligero* L, const char* const* labels, const uint64_t* var_vals, size_t n_vars, lg_sponge* sponge,
                         lg_proof** out) {
  return lg::guard([&]() { return lg_prove_with_labels_impl(L, labels, var_vals, n_vars, sponge, out); });
}

static int lg_verify_impl(lg_ligero* L, const lg_proof* P, lg_sponge* sponge, int* accepted) {
  if (!L || !P || !sponge || !accepted) return ERR_INVALID;
  *accepted = 0;
  lg_ctx* ctx = L->ctx;
  Ctx* c = &ctx->c;
  lgh::PoseidonSponge& sp = sponge->s;
  const size_t m = L->m, k = L->k, n = L->n, rows = 4 * m;
  int log_k = 0;
  while (((size_t)1 << log_k) < k) log_k++;
  const bool vt = getenv("LG_VERIFY_TIMING") != nullptr;
  auto vt0 = std::chrono::steady_clock::now();
  auto vlap = [&](const char* what) {
    if (!vt) return;
    cudaStreamSynchronize(c->stream);
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[lg_verify] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(now - vt0).count());
    vt0 = now;
  };
  sp.absorb_bytes(P->root.data(), 32);
  // ---- verify_interleaved (671-708)
  std::vector<uint8_t> seed = sp.squeeze_bytes(32);
  cudaSetDevice(c->device);
  DevMem r_dev, chk;
  LG_TRY(r_dev.cached(c, &L->vbuf[V_R], rows * sizeof(Fr)));
  LG_TRY(chk.cached(c, &L->vbuf[V_CHK], 2 * L->t * sizeof(Fr)));  // [0, t): column checks, [t, 2t): polynomial values
  std::vector<Fq> got(L->t);
  // every per-column check below is a dot product over an opened column: one CTA per column on the device
  auto run_checks = [&](int mode, const DevMem& cols, const void* w) -> int {
    LG_TRY(lg::column_checks(c, mode, (const Fr*)cols.p, (const Fr*)w, rows, L->t, (Fr*)chk.p));
    LG_CUDA(c, cudaMemcpyAsync(got.data(), chk.p, L->t * sizeof(Fr), cudaMemcpyDeviceToHost, c->stream));
    LG_CUDA(c, cudaStreamSynchronize(c->stream));
    return OK;
  };
  // a test polynomial (degree < kk) at the t opened indices of the n-point domain, from its values on a kk-point domain
  // (kk = k or 2k): one device encode of a single row of length kk at rate kk/n into the cached tile buffer, then a gather
  // of the t entries.  Same values as the reference's full-domain evaluation read at those indices (703-707, 826-829,
  // 928-932); no allocation, only the t results cross to the host.
  std::vector<Fq> poly_at(L->t);
  auto on_large_domain = [&](const std::vector<Fq>& evals, size_t kk, const std::vector<uint64_t>& at) -> int {
    const size_t tile = rows < 1024 ? rows : 1024;  // 8 * tile * k >= 32k elements: room for the kk inputs and the n outputs
    DevMem planes, at_dev;
    LG_TRY(planes.cached(c, &L->vbuf[V_PLANES], 8 * tile * k * sizeof(Fr)));
    LG_TRY(at_dev.cached(c, &L->vbuf[V_IDX], L->t * sizeof(uint64_t)));
    int log_kk = 0;
    while (((size_t)1 << log_kk) < kk) log_kk++;
    const int rho = (int)(n / kk);
    Fr* in = (Fr*)planes.p;
    Fr* p0 = in + kk;
    LG_CUDA(c, cudaMemcpyAsync(in, evals.data(), kk * sizeof(Fr), cudaMemcpyHostToDevice, c->stream));
    LG_CUDA(c, cudaMemcpyAsync(at_dev.p, at.data(), L->t * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream));
    LG_TRY(lg::encode_rows(c, in, 1, log_kk, rho, p0, p0 + kk, nullptr, false));
    Fr* res = (Fr*)chk.p + L->t;
    gather_tile_columns_kernel<<<64, 256, 0, c->stream>>>(p0, 1, log_kk, rho, (const uint64_t*)at_dev.p, L->t, 0, 1, res);
    c->launches++;
    LG_CUDA(c, cudaGetLastError());
    LG_CUDA(c, cudaMemcpyAsync(poly_at.data(), res, L->t * sizeof(Fr), cudaMemcpyDeviceToHost, c->stream));
    LG_CUDA(c, cudaStreamSynchronize(c->stream));
    return OK;
  };
  LG_TRY(expand_fr(c, seed.data(), rows, (Fr*)r_dev.p));
  sp.absorb_field(P->preenc_u_lc);
  bool ok;
  {
    DevMem cols;
    vlap("expand r + absorb lc");
    LG_TRY(verify_openings(L, P->interleaved, P->root, sp, &ok, &cols));
    if (!ok) return OK;
    vlap("openings (interleaved)");
    std::vector<Fq> msg(P->preenc_u_lc);  // reed_solomon_interpolate: msg.resize(k) pads or truncates (998-1002)
    msg.resize(k, lgh::kZero);
    LG_TRY(on_large_domain(msg, k, P->interleaved.leaf_index));  // reed_solomon(preenc_u_lc) at the opened indices
    LG_TRY(run_checks(0, cols, r_dev.p));                        // <r, column> (705-707)
    for (size_t q = 0; q < L->t; q++)
      if (poly_at[q] != got[q]) return OK;
    vlap("encode lc + checks");
  }
  // ---- verify_linear (749-830)
  seed = sp.squeeze_bytes(32);
  // r_polys on the large domain (816-819).  The reference encodes all 4m rows of r_a at the full rate and then reads t
  // columns.  Same values here, but the rows are encoded in tiles of at most 1024 once the opened indices are known, and
  // the t columns of each tile are gathered right away: a 2 GiB tile at 2^24 gates instead of a 32 GiB matrix allocated
  // next to the prover's resident one.  Only r_a itself (4mk elements) stays resident in between.
  DevMem ra_dev;
  LG_TRY(ra_dev.cached(c, (void**)&L->pre_cache, 4 * m * k * sizeof(Fr)));
  {
    int s = expand_fr(c, seed.data(), 4 * m * k, (Fr*)ra_dev.p);
    if (s == OK) s = lg_sparse_row_mul(ctx, L->a, (const uint64_t*)ra_dev.p, (uint64_t*)ra_dev.p);
    if (s != OK) return s;
    vlap("r_a");
  }
  // columns idx of R_A (t x rows, Montgomery) from row tiles of r_a
  auto gather_ra_columns = [&](const uint64_t* idx_dev, Fr* out) -> int {
    const size_t tile = rows < 1024 ? rows : 1024;
    DevMem planes;
    LG_TRY(planes.cached(c, &L->vbuf[V_PLANES], 8 * tile * k * sizeof(Fr)));
    int log_k = 0;
    while (((size_t)1 << log_k) < k) log_k++;
    for (size_t row0 = 0; row0 < rows; row0 += tile) {
      const size_t nr = row0 + tile <= rows ? tile : rows - row0;
      Fr* p0 = (Fr*)planes.p;
      LG_TRY(lg::encode_rows(c, (const Fr*)ra_dev.p + row0 * k, nr, log_k, 8, p0, p0 + nr * k, nullptr, false));
      gather_tile_columns_kernel<<<c->sm_count * 4, 256, 0, c->stream>>>(p0, nr, log_k, 8, idx_dev, L->t, row0, rows, out);
      c->launches++;
      LG_CUDA(c, cudaGetLastError());
    }
    LG_CUDA(c, cudaStreamSynchronize(c->stream));
    return OK;
  };
  const std::vector<Fq>& ql = P->linear_poly;
  const size_t deg_l = ql.empty() ? 0 : ql.size() - 1;
  if (deg_l >= 2 * k - 1) return OK;
  std::vector<Fq> ie(ql);
  ie.resize(2 * k, lgh::kZero);
  host_fft(ie, false);
  {
    Fq sum = lgh::kZero;
    for (size_t i = 0; i < 2 * k; i += 2) sum = lgh::add(sum, ie[i]);
    if (!sum.is_zero()) return OK;
  }
  sp.absorb_field(ql);
  vlap("host fft + absorb linear");
  {
    DevMem cols, rcols, idx_dev;
    int s = verify_openings(L, P->linear, P->root, sp, &ok, &cols);
    if (s != OK || !ok) return s;
    // the same columns of R_A, gathered on the device, then <R_A column, U column> per opened column (822-829)
    s = rcols.cached(c, &L->vbuf[V_RCOLS], L->t * rows * sizeof(Fr));
    if (s == OK) s = idx_dev.cached(c, &L->vbuf[V_IDX], L->t * sizeof(uint64_t));
    if (s == OK && cudaMemcpyAsync(idx_dev.p, P->linear.leaf_index.data(), L->t * sizeof(uint64_t), cudaMemcpyHostToDevice,
                                   c->stream) != cudaSuccess)
      s = fail(ctx, ERR_CUDA, "index upload failed");
    if (s == OK) s = gather_ra_columns((const uint64_t*)idx_dev.p, (Fr*)rcols.p);
    vlap("R_A row tiles: encode + gather");
    if (s == OK) s = run_checks(1, cols, rcols.p);
    vlap("checks (linear)");
    if (s != OK) return s;
    LG_TRY(on_large_domain(ie, 2 * k, P->linear.leaf_index));
    for (size_t q = 0; q < L->t; q++)
      if (got[q] != poly_at[q]) return OK;
  }
  // ---- verify_quadratic_constraints (861-933)
  seed = sp.squeeze_bytes(32);
  DevMem rq_dev;
  LG_TRY(rq_dev.cached(c, &L->vbuf[V_RQ], m * sizeof(Fr)));
  LG_TRY(expand_fr(c, seed.data(), m, (Fr*)rq_dev.p));
  const std::vector<Fq>& qq = P->quadratic_poly;
  const size_t deg_q = qq.empty() ? 0 : qq.size() - 1;
  if (deg_q >= 2 * k - 1) return OK;
  std::vector<Fq> iq(qq);
  iq.resize(2 * k, lgh::kZero);
  host_fft(iq, false);
  for (size_t cc = 0; cc < k; cc++)
    if (!iq[2 * cc].is_zero()) return OK;
  sp.absorb_field(qq);
  {
    DevMem cols;
    LG_TRY(verify_openings(L, P->quadratic, P->root, sp, &ok, &cols));
    if (!ok) return OK;
    LG_TRY(run_checks(2, cols, rq_dev.p));  // sum_i r_i (x_i y_i - z_i) per opened column (909-932)
    vlap("quadratic: fft, absorb, openings, checks");
    LG_TRY(on_large_domain(iq, 2 * k, P->quadratic.leaf_index));
    for (size_t q = 0; q < L->t; q++)
      if (poly_at[q] != got[q]) return OK;
    vlap("quadratic polynomial on the large domain");
  }
  *accepted = 1;
  return OK;
}
int lg_verify(lg_ligero* L, const lg_proof* P, lg_sponge* sponge, int* accepted) {
  return lg::guard([&]() { return lg_verify_impl(L, P, sponge, accepted); });
}

// ---------------------------------------------------------------------------------------------------
// proof container
// ---------------------------------------------------------------------------------------------------
int lg_ligero_constraints(const lg_ligero* L, const lg_constraints** out) {
  if (!L || !out) return ERR_INVALID;
  *out = L->a;
  return OK;
}

static int lg_proof_assemble_impl(const uint8_t root[32], const uint64_t* preenc_u_lc, size_t k, const uint64_t* linear_poly,
                      size_t linear_len, const uint64_t* quadratic_poly, size_t quadratic_len, size_t t, size_t rows,
                      size_t depth, const uint64_t* const cols[3], const uint64_t* const idx[3],
                      const uint8_t* const sib[3], const uint8_t* const auth[3], lg_proof** out) {
  if (!root || !out || (k && !preenc_u_lc) || (linear_len && !linear_poly) || (quadratic_len && !quadratic_poly) || !cols ||
      !idx || !sib || !auth)
    return ERR_INVALID;
  lg_proof* P = new (std::nothrow) lg_proof();
  if (!P) return ERR_NOMEM;
  memcpy(P->root.data(), root, 32);
  P->preenc_u_lc.resize(k);
  if (k) memcpy(P->preenc_u_lc.data(), preenc_u_lc, k * 32);
  P->linear_poly.resize(linear_len);
  if (linear_len) memcpy(P->linear_poly.data(), linear_poly, linear_len * 32);
  P->quadratic_poly.resize(quadratic_len);
  if (quadratic_len) memcpy(P->quadratic_poly.data(), quadratic_poly, quadratic_len * 32);
  Opened* parts[3] = {&P->interleaved, &P->linear, &P->quadratic};
  for (int p = 0; p < 3; p++) {
    if (t && (!cols[p] || !idx[p] || !sib[p] || (depth && !auth[p]))) {
      delete P;
      return ERR_INVALID;
    }
    Opened& o = *parts[p];
    o.alloc(t, rows, false);
    if (t * rows) memcpy(o.cols, cols[p], t * rows * 32);
    o.leaf_index.assign(idx[p], idx[p] + t);
    o.sibling.resize(t);
    o.auth.assign(t, std::vector<Digest>(depth));
    for (size_t q = 0; q < t; q++) {
      memcpy(o.sibling[q].data(), sib[p] + 32 * q, 32);
      for (size_t d = 0; d < depth; d++) memcpy(o.auth[q][d].data(), auth[p] + 32 * (q * depth + d), 32);
    }
  }
  *out = P;
  return OK;
}
int lg_proof_assemble(const uint8_t root[32], const uint64_t* preenc_u_lc, size_t k, const uint64_t* linear_poly,
                      size_t linear_len, const uint64_t* quadratic_poly, size_t quadratic_len, size_t t, size_t rows,
                      size_t depth, const uint64_t* const cols[3], const uint64_t* const idx[3],
                      const uint8_t* const sib[3], const uint8_t* const auth[3], lg_proof** out) {
  return lg::guard([&]() { return lg_proof_assemble_impl(root, preenc_u_lc, k, linear_poly, linear_len, quadratic_poly, quadratic_len, t, rows, depth, cols, idx, sib, auth, out); });
}

int lg_proof_free(lg_proof* p) {
  delete p;
  return OK;
}
static int lg_proof_serialize_impl(const lg_proof* P, uint8_t* buf, size_t cap, size_t* len_out) {
  if (!P || !len_out) return ERR_INVALID;
  Writer w(buf, cap);
  w.digest(P->root);
  w.frs(P->preenc_u_lc);
  w.opened(P->interleaved);
  w.frs(P->linear_poly);
  w.opened(P->linear);
  w.frs(P->quadratic_poly);
  w.opened(P->quadratic);
  *len_out = w.pos;
  return w.ok ? OK : ERR_INVALID;  // buffer too small
}
int lg_proof_serialize(const lg_proof* P, uint8_t* buf, size_t cap, size_t* len_out) {
  return lg::guard([&]() { return lg_proof_serialize_impl(P, buf, cap, len_out); });
}
static int lg_proof_deserialize_impl(const uint8_t* buf, size_t len, lg_proof** out) {
  if (!buf || !out) return ERR_INVALID;
  Reader rd{buf, len};
  lg_proof* P = new (std::nothrow) lg_proof();
  if (!P) return ERR_NOMEM;
  P->root = rd.digest();
  P->preenc_u_lc = rd.frs();
  P->interleaved = rd.opened();
  P->linear_poly = rd.frs();
  P->linear = rd.opened();
  P->quadratic_poly = rd.frs();
  P->quadratic = rd.opened();
  if (!rd.ok || rd.pos != len) {
    delete P;
    return ERR_INVALID;
  }
  *out = P;
  return OK;
}
int lg_proof_deserialize(const uint8_t* buf, size_t len, lg_proof** out) {
  return lg::guard([&]() { return lg_proof_deserialize_impl(buf, len, out); });
}

}  // extern "C"
