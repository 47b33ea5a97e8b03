// Host-only passes over an ArithmeticCircuit's node array that LigeroCircuit::new needs before anything reaches the GPU
// (src/ligero/mod.rs:147-228): the compact copy of the nodes, the check for the gates the reference panics on, the
// node -> witness-slot map (mod.rs:179-194), reachability from the outputs (mod.rs:476-478) and the level schedule of the
// device evaluator (trace.cu).  At 2^24 gates the node array is 400 MB; every pass here either runs on all host threads
// over contiguous chunks (results are independent of the thread count) or streams 9 bytes per node instead of 24.
// No CUDA in this file: tests/host/circuit_host_test.cpp checks it against straightforward loops on the CPU.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <utility>
#include <vector>

namespace lgh {

enum NodeType : uint8_t { N_VAR = 0, N_CONST = 1, N_ADD = 2, N_MUL = 3 };
struct Node {
  uint8_t type;
  uint64_t l, r;  // operands; for N_CONST l = index into const_values; for N_VAR l = index into labels
};

// allocator whose resize() leaves trivially constructible elements uninitialised: a 64 MB vector is first touched by
// the threads that fill it instead of being zeroed by one
template <class T>
struct DefaultInit : std::allocator<T> {
  template <class U>
  struct rebind {
    using other = DefaultInit<U>;
  };
  DefaultInit() = default;
  template <class U>
  DefaultInit(const DefaultInit<U>&) noexcept {}
  template <class U>
  void construct(U* p) {
    ::new ((void*)p) U;
  }
  template <class U, class... A>
  void construct(U* p, A&&... a) {
    ::new ((void*)p) U(std::forward<A>(a)...);
  }
};
template <class T>
using RawVec = std::vector<T, DefaultInit<T>>;

// threads for the passes below: LG_HOST_THREADS, else the hardware concurrency, at most 16
inline int host_threads() {
  static const int n = [] {
    if (const char* e = getenv("LG_HOST_THREADS")) {
      const int v = atoi(e);
      if (v >= 1) return v > 64 ? 64 : v;
    }
    const unsigned h = std::thread::hardware_concurrency();
    return (int)std::min(16u, std::max(1u, h));
  }();
  return n;
}

// number of contiguous chunks parallel_chunks() splits [0, n) into
inline int chunk_count(size_t n, int threads) { return (threads <= 1 || n < ((size_t)1 << 16)) ? 1 : threads; }

// f(chunk, lo, hi) for every chunk of [0, n), chunk c before chunk c + 1 in index order; chunk 0 runs on the caller.
// f must not throw.  If a thread cannot be started its chunk runs on the caller as well.
template <class F>
void parallel_chunks(size_t n, int chunks, F&& f) {
  if (chunks <= 1) {
    f(0, (size_t)0, n);
    return;
  }
  const size_t per = (n + (size_t)chunks - 1) / (size_t)chunks;
  auto lo_of = [&](int t) { return std::min(n, (size_t)t * per); };
  std::vector<std::thread> pool;
  pool.reserve((size_t)chunks - 1);
  int started = 1;
  try {
    for (; started < chunks; started++) {
      const size_t lo = lo_of(started), hi = lo_of(started + 1);
      const int t = started;
      pool.emplace_back([&f, t, lo, hi] { f(t, lo, hi); });
    }
  } catch (...) {
  }
  f(0, (size_t)0, lo_of(1));
  for (int t = started; t < chunks; t++) f(t, lo_of(t), lo_of(t + 1));
  for (auto& th : pool) th.join();
}

// the reference's bump_index (src/ligero/mod.rs:230-242): where node `index` of the caller's circuit sits once the
// constant 1 has been made node 0
inline size_t bump_index(size_t one_index, bool one_found, size_t index) {
  if (one_found) {
    if (index < one_index) return index + 1;
    if (index == one_index) return 0;
    return index;
  }
  return index + 1;
}

// The formatted copy of the node array that LigeroCircuit::new keeps (insert_one, mod.rs:244-271), written in one
// threaded pass instead of copy + erase + insert + a pass over the gates: node i of the caller's circuit lands at
// bump_index(i) with its operands bumped, `one` becomes node 0 (the caller's constant 1 is dropped when it had one).
// one_index == 0 means the circuit already starts with the constant 1: a plain copy.
inline void format_nodes(const Node* src, size_t n, size_t one_index, bool one_found, const Node& one, RawVec<Node>& dst,
                         int threads) {
  const bool plain = one_index == 0;
  dst.resize(plain || one_found ? n : n + 1);
  Node* d = dst.data();
  if (!plain) d[0] = one;
  parallel_chunks(n, chunk_count(n, threads), [=](int, size_t lo, size_t hi) {
    if (plain) {
      if (hi > lo) memcpy((void*)(d + lo), (const void*)(src + lo), (hi - lo) * sizeof(Node));
      return;
    }
    for (size_t i = lo; i < hi; i++) {
      if (one_found && i == one_index) continue;
      Node nd = src[i];
      if (nd.type >= N_ADD) {
        nd.l = bump_index(one_index, one_found, nd.l);
        nd.r = bump_index(one_index, one_found, nd.r);
      }
      d[bump_index(one_index, one_found, i)] = nd;
    }
  });
}

// type / left / right of every node as three flat arrays (9 bytes per node; what constraints.cu takes as well)
struct NodeArrays {
  size_t n = 0;
  std::unique_ptr<uint8_t[]> type;
  std::unique_ptr<uint32_t[]> l, r;
};

// precondition: n < 2^32 and every operand index < 2^32 (the callers refuse circuits of 2^31 nodes or more)
inline void pack_nodes(const Node* nodes, size_t n, NodeArrays& a, int threads) {
  a.n = n;
  a.type.reset(new uint8_t[n ? n : 1]);
  a.l.reset(new uint32_t[n ? n : 1]);
  a.r.reset(new uint32_t[n ? n : 1]);
  uint8_t* type = a.type.get();
  uint32_t *l = a.l.get(), *r = a.r.get();
  parallel_chunks(n, chunk_count(n, threads), [=](int, size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; i++) {
      type[i] = nodes[i].type;
      l[i] = (uint32_t)nodes[i].l;
      r[i] = (uint32_t)nodes[i].r;
    }
  });
}

// lowest index of an Add or Mul gate whose operands are both constants (the reference panics on them, mod.rs:325, 345),
// or SIZE_MAX.  Operands must be valid node indices.
inline size_t first_gate_of_two_constants(const NodeArrays& a, int threads) {
  const int chunks = chunk_count(a.n, threads);
  std::vector<size_t> first((size_t)chunks, SIZE_MAX);
  const uint8_t* type = a.type.get();
  const uint32_t *l = a.l.get(), *r = a.r.get();
  size_t* out = first.data();
  parallel_chunks(a.n, chunks, [=](int t, size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; i++)
      if (type[i] >= N_ADD && type[l[i]] == N_CONST && type[r[i]] == N_CONST) {
        out[t] = i;
        return;
      }
  });
  for (size_t v : first)
    if (v != SIZE_MAX) return v;
  return SIZE_MAX;
}

struct Schedule {
  RawVec<uint32_t> index_map;  // node -> slot in the X/Y/Z/W blocks; 0xffffffff for the constants other than node 0
  RawVec<uint8_t> reach;       // node is an output or feeds one
  bool all_gates_reach = true;
  RawVec<uint32_t> var_nodes, const_nodes;  // every Variable / Constant node, ascending
  size_t n_gates = 0;
  uint32_t depth = 0;                 // number of levels; a gate's level is 1 + the larger level of its operands
  std::vector<uint32_t> level_start;  // depth + 1 offsets into the gate arrays
  // gates ordered by (level, Add before Mul, node index): node | Mul flag in bit 31, operands, witness slot
  RawVec<uint32_t> gate_node, gate_l, gate_r, gate_pos;
};

namespace detail {

// Level of every gate = 1 + the larger level of its operands (operands always precede a gate: ArithmeticCircuit only
// appends).  One forward sweep settles them, but it is a chain of random reads that one core's miss queue bounds (10 ns
// per gate), so the sweep goes block by block: inside a block of 2^17 nodes all threads take the gates whose operands both
// lie before the block (their levels are final), and the few gates with an operand inside the block are finished by
// one thread in index order afterwards.  `level` must be zero on entry.  false if a level does not fit LT.
template <class LT>
bool compute_levels(const NodeArrays& a, LT* level, int threads, uint32_t* depth_out, size_t* gates_out) {
  const size_t N = a.n;
  const uint8_t* type = a.type.get();
  const uint32_t *L = a.l.get(), *R = a.r.get();
  const uint32_t level_max = (uint32_t)(LT)~(LT)0;
  const int chunks = chunk_count(N, threads);
  uint32_t depth = 0;
  size_t n_gates = 0;
  if (chunks == 1) {
    for (size_t i = 0; i < N; i++) {
      if (type[i] < N_ADD) continue;
      const uint32_t v = 1u + std::max<uint32_t>(level[L[i]], level[R[i]]);
      if (v > level_max) return false;
      level[i] = (LT)v;
      depth = std::max(depth, v);
      n_gates++;
    }
    *depth_out = depth;
    *gates_out = n_gates;
    return true;
  }
  constexpr size_t kBlock = (size_t)1 << 17;
  std::vector<uint32_t> deferred(kBlock);  // gates left to the sequential step; thread t writes from seg_lo[t] on
  std::vector<size_t> seg_lo((size_t)chunks, 0), n_def((size_t)chunks, 0), gates((size_t)chunks, 0);
  std::vector<uint32_t> deepest((size_t)chunks, 0);
  std::vector<uint8_t> over((size_t)chunks, 0);
  uint32_t* def_all = deferred.data();
  size_t *p_lo = seg_lo.data(), *p_nd = n_def.data(), *p_g = gates.data();
  uint32_t* p_deep = deepest.data();
  uint8_t* p_over = over.data();
  for (size_t b0 = 0; b0 < N; b0 += kBlock) {
    const size_t bn = std::min(kBlock, N - b0);
    parallel_chunks(bn, chunks, [=](int t, size_t lo, size_t hi) {
      uint32_t* def = def_all + lo;
      size_t nd = 0, g = 0;
      uint32_t dp = p_deep[t];
      for (size_t j = lo; j < hi; j++) {
        const size_t i = b0 + j;
        if (type[i] < N_ADD) continue;
        g++;
        const uint32_t l = L[i], r = R[i];
        if (l >= b0 || r >= b0) {
          def[nd++] = (uint32_t)i;
          continue;
        }
        const uint32_t v = 1u + std::max<uint32_t>(level[l], level[r]);
        if (v > level_max) {
          p_over[t] = 1;
          continue;
        }
        level[i] = (LT)v;
        dp = std::max(dp, v);
      }
      p_lo[t] = lo;
      p_nd[t] = nd;
      p_g[t] += g;
      p_deep[t] = dp;
    });
    for (int t = 0; t < chunks; t++) {
      if (over[(size_t)t]) return false;
      const uint32_t* def = def_all + seg_lo[(size_t)t];
      for (size_t q = 0; q < n_def[(size_t)t]; q++) {
        const size_t i = def[q];
        const uint32_t v = 1u + std::max<uint32_t>(level[L[i]], level[R[i]]);
        if (v > level_max) return false;
        level[i] = (LT)v;
        depth = std::max(depth, v);
      }
    }
  }
  for (int t = 0; t < chunks; t++) {
    depth = std::max(depth, deepest[(size_t)t]);
    n_gates += gates[(size_t)t];
  }
  *depth_out = depth;
  *gates_out = n_gates;
  return true;
}

// levels + counting sort, with the level of a node held in LT; false if a level does not fit (nothing is kept then)
template <class LT>
bool schedule_gates(const NodeArrays& a, Schedule& s, int threads) {
  const size_t N = a.n;
  const uint8_t* type = a.type.get();
  const uint32_t *L = a.l.get(), *R = a.r.get();
  RawVec<LT> level(N);
  {
    const int chunks = chunk_count(N, threads);
    LT* lv = level.data();
    parallel_chunks(N, chunks, [=](int, size_t lo, size_t hi) { memset(lv + lo, 0, (hi - lo) * sizeof(LT)); });
  }
  uint32_t depth = 0;
  size_t n_gates = 0;
  if (!compute_levels<LT>(a, level.data(), threads, &depth, &n_gates)) return false;
  s.depth = depth;
  s.n_gates = n_gates;
  const size_t B = 2 * (size_t)depth;  // bucket of a gate: 2 * (level - 1) + (Mul ? 1 : 0)
  // per-chunk counts cost chunks * B words: deep, thin circuits (B close to N) are counted by one chunk
  int chunks = chunk_count(N, threads);
  if ((size_t)chunks * B > N / 2 + 1024) chunks = 1;
  std::vector<std::vector<uint32_t>> cursor((size_t)chunks, std::vector<uint32_t>(B, 0));
  {
    std::vector<uint32_t>* cur = cursor.data();
    const LT* lv = level.data();
    parallel_chunks(N, chunks, [=](int t, size_t lo, size_t hi) {
      uint32_t* c = cur[t].data();
      for (size_t i = lo; i < hi; i++)
        if (type[i] >= N_ADD) c[2 * ((size_t)lv[i] - 1) + (type[i] == N_MUL)]++;
    });
  }
  // exclusive prefix over (bucket, chunk): chunk t's gates of a bucket come after those of the chunks before it, which
  // keeps every bucket in ascending node order
  s.level_start.assign((size_t)depth + 1, 0);
  uint32_t run = 0;
  for (size_t b = 0; b < B; b++) {
    if ((b & 1) == 0) s.level_start[b / 2] = run;
    for (int t = 0; t < chunks; t++) {
      const uint32_t c = cursor[(size_t)t][b];
      cursor[(size_t)t][b] = run;
      run += c;
    }
  }
  s.level_start[depth] = run;
  s.gate_node.resize(n_gates);
  s.gate_l.resize(n_gates);
  s.gate_r.resize(n_gates);
  s.gate_pos.resize(n_gates);
  {
    std::vector<uint32_t>* cur = cursor.data();
    const LT* lv = level.data();
    uint32_t *gn = s.gate_node.data(), *gl = s.gate_l.data(), *gr = s.gate_r.data(), *gp = s.gate_pos.data();
    const uint32_t* imap = s.index_map.data();
    parallel_chunks(N, chunks, [=](int t, size_t lo, size_t hi) {
      uint32_t* c = cur[t].data();
      for (size_t i = lo; i < hi; i++) {
        if (type[i] < N_ADD) continue;
        const uint32_t g = c[2 * ((size_t)lv[i] - 1) + (type[i] == N_MUL)]++;
        gn[g] = (uint32_t)i | (type[i] == N_MUL ? 0x80000000u : 0u);
        gl[g] = L[i];
        gr[g] = R[i];
        gp[g] = imap[i];
      }
    });
  }
  return true;
}

}  // namespace detail

// precondition: a.n >= 1, a.n < 2^31, operands of every gate are indices of earlier nodes, outputs are node indices
inline void build_schedule(const NodeArrays& a, const size_t* outputs, size_t n_outputs, Schedule& s, int threads) {
  const size_t N = a.n;
  const uint8_t* type = a.type.get();
  const uint32_t *L = a.l.get(), *R = a.r.get();
  const int chunks = chunk_count(N, threads);
  // node -> witness slot (mod.rs:483-504: constants other than node 0 own no slot), and the lists of constants / variables
  std::vector<size_t> nconst((size_t)chunks + 1, 0), nvar((size_t)chunks + 1, 0);
  {
    size_t *pc = nconst.data(), *pv = nvar.data();
    parallel_chunks(N, chunks, [=](int t, size_t lo, size_t hi) {
      size_t c = 0, v = 0;
      for (size_t i = lo; i < hi; i++) {
        c += type[i] == N_CONST;
        v += type[i] == N_VAR;
      }
      pc[t + 1] = c;
      pv[t + 1] = v;
    });
  }
  for (int t = 0; t < chunks; t++) {
    nconst[(size_t)t + 1] += nconst[(size_t)t];
    nvar[(size_t)t + 1] += nvar[(size_t)t];
  }
  s.index_map.resize(N);
  s.const_nodes.resize(nconst[(size_t)chunks]);
  s.var_nodes.resize(nvar[(size_t)chunks]);
  {
    const size_t *pc = nconst.data(), *pv = nvar.data();
    uint32_t *imap = s.index_map.data(), *cn = s.const_nodes.data(), *vn = s.var_nodes.data();
    const size_t zero_is_const = type[0] == N_CONST ? 1 : 0;
    parallel_chunks(N, chunks, [=](int t, size_t lo, size_t hi) {
      size_t c = pc[t], v = pv[t];
      for (size_t i = lo; i < hi; i++) {
        if (type[i] == N_CONST) {
          cn[c++] = (uint32_t)i;
          imap[i] = 0xffffffffu;
        } else {
          imap[i] = (uint32_t)(i - (c - zero_is_const));  // constants among nodes 1..i
          if (type[i] == N_VAR) vn[v++] = (uint32_t)i;
        }
      }
    });
    imap[0] = 0;
  }
  // which nodes feed an output (the reference panics at prove time on any that does not, mod.rs:476-478)
  s.reach.resize(N);
  memset(s.reach.data(), 0, N);
  for (size_t o = 0; o < n_outputs; o++) s.reach[outputs[o]] = 1;
  s.all_gates_reach = true;
  {
    uint8_t* reach = s.reach.data();
    for (size_t i = N; i-- > 0;) {
      if (type[i] < N_ADD) continue;
      if (reach[i]) reach[L[i]] = reach[R[i]] = 1;
      else s.all_gates_reach = false;
    }
  }
  if (!detail::schedule_gates<uint16_t>(a, s, threads)) detail::schedule_gates<uint32_t>(a, s, threads);
}

}  // namespace lgh
