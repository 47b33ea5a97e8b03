// Evaluation trace and witness layout on the device (sm_100a)  --  SURVEY 8(a) row a1 / 8(f)-3.
//
//   reference: ArithmeticCircuit::evaluation_trace_multioutput (src/arithmetic_circuit/mod.rs:325-358, recursive
//   inner_evaluate 247-271) followed by the layout loop of prove_inner (src/ligero/mod.rs:476-516, as_matrix 1015-1017):
//       w = [1, every non-constant node's value ...],   (x, y, z) = (left, right, out) at Mul nodes, else 0,
//       each padded to m*k and cut into m rows of k:  preenc_u = [X; Y; Z; W].
//
// A node's value depends only on its operands, so the recursion is replaced by a LEVEL schedule built once per circuit
// (host_driver, lg_ligero_new): level(gate) = 1 + max(level(left), level(right)); gates sorted by level, Add before Mul
// inside a level so warps do not diverge.  One launch evaluates one wide level, thread per gate, and writes the
// gate's value BOTH into the value table and straight into its slot of the W block (and X, Y, Z for a Mul), so the
// 4mk-element matrix is produced in HBM where the encoder reads it: no host trace, no H2D of the matrix (4 GiB at
// 2^24 gates).  Runs of narrow levels (R1CS-compiled circuits are deep and thin) are walked by ONE CTA that
// separates levels with __syncthreads(): a launch per level would cost more than the level.
// Values are the same Montgomery products/sums as the host evaluator's, so the matrix is bit-identical.
//
// Algorithmic traffic per gate: 12 B of schedule + 2 x 32 B operand gathers + 32 B value + 32 B (Add) or 128 B (Mul)
// of witness = 140 / 236 B; one Montgomery product per Mul gate.  HBM/L2-gather bound.
#include "lg_internal.h"

namespace lg {

namespace {

__device__ __forceinline__ Fr t_ld(const Fr* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  const uint4 a = q[0], b = q[1];
  Fr r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ void t_st(Fr* p, const Fr& x) {
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
  q[1] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
}

constexpr uint32_t kMulBit = 0x80000000u;

// constants and variables: value table + (for nodes that own a witness slot) the W block
__global__ void trace_init_kernel(const uint32_t* __restrict__ node, const uint32_t* __restrict__ pos,
                                  const Fr* __restrict__ val, size_t count, Fr* __restrict__ vals, Fr* __restrict__ W) {
  const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= count) return;
  const Fr v = t_ld(val + j);
  t_st(vals + node[j], v);
  const uint32_t p = pos[j];
  if (p != 0xffffffffu) t_st(W + p, v);
}

__device__ __forceinline__ void trace_gate(const uint32_t* __restrict__ gnode, const uint32_t* __restrict__ gl,
                                           const uint32_t* __restrict__ gr, const uint32_t* __restrict__ gpos, size_t g,
                                           Fr* vals, Fr* __restrict__ X, size_t mk) {
  const uint32_t nd = gnode[g];
  const Fr a = t_ld(vals + gl[g]), b = t_ld(vals + gr[g]);
  const uint32_t p = gpos[g];
  Fr v;
  if (nd & kMulBit) {
    v = fr_mul(a, b);
    t_st(X + p, a);
    t_st(X + mk + p, b);
    t_st(X + 2 * mk + p, v);
  } else {
    v = fr_add(a, b);
  }
  t_st(vals + (nd & ~kMulBit), v);
  t_st(X + 3 * mk + p, v);
}

// one wide level: gates [g0, g1)
__global__ void __launch_bounds__(256) trace_level_kernel(const uint32_t* __restrict__ gnode, const uint32_t* __restrict__ gl,
                                                          const uint32_t* __restrict__ gr, const uint32_t* __restrict__ gpos,
                                                          size_t g0, size_t g1, Fr* vals, Fr* __restrict__ X, size_t mk) {
  const size_t g = g0 + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g < g1) trace_gate(gnode, gl, gr, gpos, g, vals, X, mk);
}

// a run of narrow levels [lv0, lv1) in one CTA; level_start[l] .. level_start[l+1] are the gates of level l
__global__ void __launch_bounds__(1024) trace_narrow_kernel(const uint32_t* __restrict__ gnode, const uint32_t* __restrict__ gl,
                                                            const uint32_t* __restrict__ gr, const uint32_t* __restrict__ gpos,
                                                            const uint32_t* __restrict__ level_start, size_t lv0, size_t lv1,
                                                            Fr* vals, Fr* __restrict__ X, size_t mk) {
  for (size_t l = lv0; l < lv1; l++) {
    const size_t g1 = level_start[l + 1];
    for (size_t g = (size_t)level_start[l] + threadIdx.x; g < g1; g += blockDim.x) trace_gate(gnode, gl, gr, gpos, g, vals, X, mk);
    __syncthreads();  // the next level reads what this one wrote (global memory, same CTA)
  }
}

}  // namespace

int trace_run(Ctx* ctx, const TraceSchedule& t, const uint32_t* var_node, const uint32_t* var_pos, const Fr* var_val,
              size_t n_vars, Fr* out) {
  cudaStream_t st = ctx->stream;
  const size_t mk = t.mk;
  LG_CUDA(ctx, cudaMemsetAsync(out, 0, 4 * mk * sizeof(Fr), st));
  Fr* W = out + 3 * mk;
  if (t.n_consts) {
    trace_init_kernel<<<(unsigned)((t.n_consts + 255) / 256), 256, 0, st>>>(t.const_node, t.const_pos, t.const_val, t.n_consts,
                                                                            t.vals, W);
    ctx->launches++;
  }
  if (n_vars) {
    trace_init_kernel<<<(unsigned)((n_vars + 255) / 256), 256, 0, st>>>(var_node, var_pos, var_val, n_vars, t.vals, W);
    ctx->launches++;
  }
  for (const auto& seg : t.segments) {
    if (seg.narrow) {
      trace_narrow_kernel<<<1, 1024, 0, st>>>(t.gate_node, t.gate_l, t.gate_r, t.gate_pos, t.level_start, seg.lv0, seg.lv1,
                                              t.vals, out, mk);
    } else {
      const size_t width = seg.g1 - seg.g0;
      trace_level_kernel<<<(unsigned)((width + 255) / 256), 256, 0, st>>>(t.gate_node, t.gate_l, t.gate_r, t.gate_pos, seg.g0,
                                                                          seg.g1, t.vals, out, mk);
    }
    ctx->launches++;
  }
  LG_CUDA(ctx, cudaGetLastError());
  return OK;
}

void trace_free(TraceSchedule& t) {
  cudaFree(t.gate_node);
  cudaFree(t.gate_l);
  cudaFree(t.gate_r);
  cudaFree(t.gate_pos);
  cudaFree(t.level_start);
  cudaFree(t.const_node);
  cudaFree(t.const_pos);
  cudaFree(t.const_val);
  cudaFree(t.vals);
  t = TraceSchedule();
}

}  // namespace lg
