// CPU check of ligero_b200/csrc/circuit_host.h: the threaded passes of LigeroCircuit::new against straightforward
// single-threaded loops (the statement of what each array is), on random circuits of many shapes and thread counts.
//   circuit_host_test <seed>   ->  prints "CIRCUIT_HOST_OK <cases>" or the first mismatch and exits 1
#include <cstdio>
#include <string>

#include "../../ligero_b200/csrc/circuit_host.h"

using namespace lgh;

static uint64_t st;
static uint64_t next() {
  uint64_t z = (st += 0x9e3779b97f4a7c15ULL);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return z ^ (z >> 31);
}
static uint64_t below(uint64_t n) { return (uint64_t)(((unsigned __int128)next() * n) >> 64); }

// shape 0: random operands (shallow and wide); 1: a chain (every gate uses the previous node: depth = gates);
// 2: operands among the last 8 nodes (deep); 3: no gates at all
static std::vector<Node> make_circuit(size_t n, int shape, unsigned const_pct, unsigned var_pct) {
  std::vector<Node> v;
  v.push_back({N_CONST, 0, 0});
  v.push_back({N_VAR, 0, 0});
  while (v.size() < n) {
    const size_t i = v.size();
    const unsigned roll = (unsigned)below(100);
    if (shape == 3 || roll < const_pct) v.push_back({(uint8_t)(roll & 1 ? N_CONST : N_VAR), below(5), 0});
    else if (roll < const_pct + var_pct) v.push_back({N_VAR, i, 0});
    else {
      uint64_t l, r;
      if (shape == 1) l = i - 1, r = below(i);
      else if (shape == 2) l = i - 1 - below(std::min<size_t>(i, 8)), r = i - 1 - below(std::min<size_t>(i, 8));
      else l = below(i), r = below(i);
      v.push_back({(uint8_t)(next() & 1 ? N_MUL : N_ADD), l, r});
    }
  }
  return v;
}

#define CHECK(cond, what)                                                                      \
  do {                                                                                         \
    if (!(cond)) {                                                                             \
      printf("MISMATCH %s: n=%zu shape=%d threads=%d\n", what, nodes.size(), shape, threads);  \
      return false;                                                                            \
    }                                                                                          \
  } while (0)

static bool check(const std::vector<Node>& nodes, int shape, int threads, const std::vector<size_t>& outputs) {
  const size_t N = nodes.size();
  NodeArrays a;
  pack_nodes(nodes.data(), N, a, threads);
  for (size_t i = 0; i < N; i++) CHECK(a.type[i] == nodes[i].type && a.l[i] == nodes[i].l && a.r[i] == nodes[i].r, "pack");
  // --- the plain statements
  size_t bad = SIZE_MAX;
  for (size_t i = 0; i < N && bad == SIZE_MAX; i++)
    if (nodes[i].type >= N_ADD && nodes[nodes[i].l].type == N_CONST && nodes[nodes[i].r].type == N_CONST) bad = i;
  CHECK(first_gate_of_two_constants(a, threads) == bad, "first gate of two constants");
  std::vector<uint32_t> index_map(N, 0xffffffffu);
  index_map[0] = 0;
  size_t seen = 0;
  for (size_t i = 1; i < N; i++) {
    if (nodes[i].type == N_CONST) seen++;
    else index_map[i] = (uint32_t)(i - seen);
  }
  std::vector<uint8_t> reach(N, 0);
  for (size_t o : outputs) reach[o] = 1;
  bool all = true;
  for (size_t i = N; i-- > 0;) {
    if (nodes[i].type < N_ADD) continue;
    if (reach[i]) reach[nodes[i].l] = reach[nodes[i].r] = 1;
    else all = false;
  }
  std::vector<uint32_t> level(N, 0);
  uint32_t depth = 0;
  size_t n_gates = 0;
  for (size_t i = 0; i < N; i++) {
    if (nodes[i].type < N_ADD) continue;
    level[i] = 1 + std::max(level[nodes[i].l], level[nodes[i].r]);
    depth = std::max(depth, level[i]);
    n_gates++;
  }
  // gates ordered by (level, Add before Mul, node index)
  std::vector<uint32_t> order;
  for (size_t i = 0; i < N; i++)
    if (nodes[i].type >= N_ADD) order.push_back((uint32_t)i);
  std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) {
    if (level[x] != level[y]) return level[x] < level[y];
    return (nodes[x].type == N_MUL) < (nodes[y].type == N_MUL);
  });
  std::vector<uint32_t> gnode, gl, gr, gpos, level_start((size_t)depth + 1, 0);
  for (uint32_t i : order) {
    gnode.push_back(i | (nodes[i].type == N_MUL ? 0x80000000u : 0u));
    gl.push_back((uint32_t)nodes[i].l);
    gr.push_back((uint32_t)nodes[i].r);
    gpos.push_back(index_map[i]);
    level_start[level[i]]++;  // gates of level lv counted at [lv]; turned into offsets below
  }
  for (size_t lv = 1; lv <= depth; lv++) level_start[lv] += level_start[lv - 1];  // now [lv] = end of level lv = start of lv + 1
  std::vector<uint32_t> vars, consts;
  for (size_t i = 0; i < N; i++) {
    if (nodes[i].type == N_VAR) vars.push_back((uint32_t)i);
    if (nodes[i].type == N_CONST) consts.push_back((uint32_t)i);
  }
  // --- the threaded builder
  Schedule s;
  build_schedule(a, outputs.data(), outputs.size(), s, threads);
  CHECK(s.index_map.size() == N && std::equal(index_map.begin(), index_map.end(), s.index_map.begin()), "index_map");
  CHECK(s.reach.size() == N && std::equal(reach.begin(), reach.end(), s.reach.begin()), "reach");
  CHECK(s.all_gates_reach == all, "all_gates_reach");
  CHECK(s.n_gates == n_gates && s.depth == depth, "gate count / depth");
  CHECK(s.level_start == level_start, "level_start");
  CHECK(s.gate_node.size() == gnode.size() && std::equal(gnode.begin(), gnode.end(), s.gate_node.begin()), "gate_node");
  CHECK(s.gate_l.size() == gl.size() && std::equal(gl.begin(), gl.end(), s.gate_l.begin()), "gate_l");
  CHECK(s.gate_r.size() == gr.size() && std::equal(gr.begin(), gr.end(), s.gate_r.begin()), "gate_r");
  CHECK(s.gate_pos.size() == gpos.size() && std::equal(gpos.begin(), gpos.end(), s.gate_pos.begin()), "gate_pos");
  CHECK(s.var_nodes.size() == vars.size() && std::equal(vars.begin(), vars.end(), s.var_nodes.begin()), "var_nodes");
  CHECK(s.const_nodes.size() == consts.size() && std::equal(consts.begin(), consts.end(), s.const_nodes.begin()), "const_nodes");
  return true;
}

// format_nodes against the reference's sequence: copy, remove the constant 1 if the circuit has one, put it in front,
// bump the operands of every gate (mod.rs:244-271)
static bool check_format(const std::vector<Node>& nodes, int shape, int threads) {
  const size_t N = nodes.size();
  const Node one{N_CONST, 77, 0};
  // (one_index, one_found): already in front; found at a constant somewhere inside; not found
  std::vector<std::pair<size_t, bool>> cases = {{0, true}, {1, false}};
  for (size_t i = N / 2; i < N; i++)
    if (nodes[i].type == N_CONST) {
      cases.push_back({i, true});
      break;
    }
  for (size_t i = 1; i < N; i++)
    if (nodes[i].type == N_CONST) {
      cases.push_back({i, true});
      break;
    }
  for (auto [oi, of] : cases) {
    std::vector<Node> want = nodes;
    if (oi != 0) {
      if (of) want.erase(want.begin() + oi);
      want.insert(want.begin(), one);
      for (auto& nd : want)
        if (nd.type == N_ADD || nd.type == N_MUL) {
          nd.l = bump_index(oi, of, nd.l);
          nd.r = bump_index(oi, of, nd.r);
        }
    }
    RawVec<Node> got;
    format_nodes(nodes.data(), N, oi, of, one, got, threads);
    CHECK(got.size() == want.size(), "format_nodes size");
    for (size_t i = 0; i < want.size(); i++)
      CHECK(got[i].type == want[i].type && got[i].l == want[i].l && got[i].r == want[i].r, "format_nodes");
  }
  return true;
}

int main(int argc, char** argv) {
  st = argc > 1 ? strtoull(argv[1], nullptr, 10) : 1;
  int cases = 0;
  // sizes on both sides of the 65 536-node threshold below which everything runs on one chunk; a 70 000-gate chain has
  // more than 65 535 levels (the 16-bit level table overflows and the 32-bit one takes over)
  const size_t sizes[] = {2, 3, 7, 100, 5000, 65535, 65536, 70001, 200003, 600011};  // the last two span 2 and 5 level blocks
  for (size_t n : sizes)
    for (int shape = 0; shape < 4; shape++)
      for (int threads : {1, 2, 3, 8}) {
        const unsigned const_pct = shape == 1 ? 0 : 10, var_pct = shape == 1 ? 0 : 5;
        const std::vector<Node> nodes = make_circuit(n, shape, const_pct, var_pct);
        std::vector<size_t> outputs;
        for (size_t i = nodes.size(); i-- > 0 && outputs.size() < 2;)
          if (nodes[i].type >= N_ADD) outputs.push_back(i);
        if (!check(nodes, shape, threads, outputs)) return 1;
        if (!check_format(nodes, shape, threads)) return 1;
        cases++;
      }
  printf("CIRCUIT_HOST_OK %d\n", cases);
  return 0;
}
