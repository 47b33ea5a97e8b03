// Constraint matrix on the device (reference: generate_matrices, src/ligero/mod.rs:296-433).
//
// Only the right-hand mk-column block of A = [[I, -(Px;Py;Pz)], [0, Padd]] is stored, in CSC form (lg_constraints).  Every
// Add / Mul gate and every output contributes exactly three entries, so the builder is: two prefix sums over the node
// array (constants dropped so far -> witness slot = row inside each P block; gates so far -> position of the gate's
// three entries), one thread per gate writing its (column, row, value id) triplets in the reference's order, a stable
// radix sort of the entry numbers by column (CUB: set-up code, not the proving path), and a gather.  The resulting arrays
// are identical to the host builder's (build_constraints in capi_host.cu), which stays as the checker and for small circuits.
#include <cub/cub.cuh>

#include "capi_types.h"

namespace lg {

namespace {

constexpr uint8_t kVar = 0, kConst = 1, kAdd = 2, kMul = 3;

struct NodeArrays {
  const uint8_t* type;
  const uint32_t *l, *r;
  const uint32_t* cexcl;  // constants (other than node 0) among nodes [0, i)
  const uint32_t *vidp, *vidn;  // value id of +c / -c per entry of const_values
};

__global__ void node_flags_kernel(const uint8_t* __restrict__ type, size_t n, uint32_t* __restrict__ is_const, uint32_t* __restrict__ is_gate) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint8_t t = type[i];
  is_const[i] = (t == kConst && i != 0) ? 1u : 0u;
  is_gate[i] = (t == kAdd || t == kMul) ? 1u : 0u;
}

__device__ __forceinline__ void put(uint32_t* cols, uint32_t* rows, uint32_t* vids, uint32_t* counts, size_t e, uint32_t col, uint32_t row,
                                    uint32_t vid) {
  cols[e] = col;
  rows[e] = row;
  vids[e] = vid;
  atomicAdd(&counts[col + 1], 1u);
}

// the three entries of one gate (mod.rs:318-367 for a node, 369-414 for an output: own column 0)
__device__ __forceinline__ void emit_gate(const NodeArrays& a, uint8_t t, uint32_t l, uint32_t r, uint32_t own_col, uint32_t row, uint32_t mk,
                                          size_t e, uint32_t* cols, uint32_t* rows, uint32_t* vids, uint32_t* counts) {
  const bool lc = a.type[l] == kConst, rc = a.type[r] == kConst;
  const uint32_t il = l - a.cexcl[l], ir = r - a.cexcl[r];  // witness slots of the operands (node 0 -> 0)
  if (t == kAdd) {
    const uint32_t base = 3u * mk + row;
    if (lc) {
      put(cols, rows, vids, counts, e, 0, base, a.vidp[a.l[l]]);
      put(cols, rows, vids, counts, e + 1, ir, base, 0);
    } else if (rc) {
      put(cols, rows, vids, counts, e, il, base, 0);
      put(cols, rows, vids, counts, e + 1, 0, base, a.vidp[a.l[r]]);
    } else {
      put(cols, rows, vids, counts, e, il, base, 0);
      put(cols, rows, vids, counts, e + 1, ir, base, 0);
    }
    put(cols, rows, vids, counts, e + 2, own_col, base, 1);
  } else {
    if (lc) {
      put(cols, rows, vids, counts, e, 0, row, a.vidn[a.l[l]]);
      put(cols, rows, vids, counts, e + 1, ir, mk + row, 1);
    } else if (rc) {
      put(cols, rows, vids, counts, e, il, row, 1);
      put(cols, rows, vids, counts, e + 1, 0, mk + row, a.vidn[a.l[r]]);
    } else {
      put(cols, rows, vids, counts, e, il, row, 1);
      put(cols, rows, vids, counts, e + 1, ir, mk + row, 1);
    }
    put(cols, rows, vids, counts, e + 2, own_col, 2u * mk + row, 1);
  }
}

__global__ void emit_nodes_kernel(NodeArrays a, const uint32_t* __restrict__ gexcl, size_t n, uint32_t mk, uint32_t* cols, uint32_t* rows,
                                  uint32_t* vids, uint32_t* counts) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint8_t t = a.type[i];
  if (t != kAdd && t != kMul) return;
  const uint32_t slot = (uint32_t)i - a.cexcl[i];
  emit_gate(a, t, a.l[i], a.r[i], slot, slot, mk, 3 * (size_t)gexcl[i], cols, rows, vids, counts);
}

__global__ void emit_outputs_kernel(NodeArrays a, const uint32_t* __restrict__ outputs, size_t n_out, uint32_t first_row, size_t first_entry,
                                    uint32_t mk, uint32_t* cols, uint32_t* rows, uint32_t* vids, uint32_t* counts) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_out) return;
  const uint32_t o = outputs[q];
  emit_gate(a, a.type[o], a.l[o], a.r[o], 0, first_row + (uint32_t)q, mk, first_entry + 3 * q, cols, rows, vids, counts);
}

__global__ void iota_kernel(uint32_t* v, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = (uint32_t)i;
}
__global__ void gather2_kernel(const uint32_t* __restrict__ perm, const uint32_t* __restrict__ rows, const uint32_t* __restrict__ vids, size_t n,
                               uint32_t* __restrict__ row_idx, uint32_t* __restrict__ val_id) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t p = perm[i];
  row_idx[i] = rows[p];
  val_id[i] = vids[p];
}

struct Buf {  // device scratch freed at scope exit
  void* p = nullptr;
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 4); }
  ~Buf() {
    if (p) cudaFree(p);
  }
  template <class T>
  T* as() {
    return (T*)p;
  }
};

}  // namespace

// type / l / r: host node arrays (n nodes; for a constant node l = its index in const_values); vidp / vidn: host, one per
// const_values entry; outputs: host node indices.  Validation (gates with two constant operands, output node types) is the
// caller's.  On success *out owns the device CSC.
int build_constraints_device(lg_ctx* ctx, const uint8_t* type, const uint32_t* l, const uint32_t* r, size_t n, const uint32_t* vidp,
                             const uint32_t* vidn, size_t n_const_values, const uint32_t* outputs, size_t n_out, size_t mk,
                             const uint64_t* const_table, size_t n_table, lg_constraints** out) {
  Ctx* c = &ctx->c;
  cudaSetDevice(c->device);
  cudaStream_t st = c->stream;
  if (4 * mk > 0xffffffffull || n >= 0x7fffffffull) return set_error(c, ERR_UNSUPPORTED, "circuit too large for 32-bit constraint indices");
  Buf d_type, d_l, d_r, d_vidp, d_vidn, d_out, d_isc, d_isg, d_cexcl, d_gexcl, d_tmp;
#define CK(expr)                                                                                          \
  do {                                                                                                    \
    cudaError_t _e = (expr);                                                                              \
    if (_e != cudaSuccess) return set_error(c, _e == cudaErrorMemoryAllocation ? ERR_NOMEM : ERR_CUDA,     \
                                            std::string("constraint builder: ") + cudaGetErrorString(_e)); \
  } while (0)
  CK(d_type.alloc(n));
  CK(d_l.alloc(n * 4));
  CK(d_r.alloc(n * 4));
  CK(d_vidp.alloc(n_const_values * 4));
  CK(d_vidn.alloc(n_const_values * 4));
  CK(d_out.alloc(n_out * 4));
  CK(d_isc.alloc(n * 4));
  CK(d_isg.alloc(n * 4));
  CK(d_cexcl.alloc((n + 1) * 4));
  CK(d_gexcl.alloc((n + 1) * 4));
  CK(cudaMemcpyAsync(d_type.p, type, n, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_l.p, l, n * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_r.p, r, n * 4, cudaMemcpyHostToDevice, st));
  if (n_const_values) {
    CK(cudaMemcpyAsync(d_vidp.p, vidp, n_const_values * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_vidn.p, vidn, n_const_values * 4, cudaMemcpyHostToDevice, st));
  }
  if (n_out) CK(cudaMemcpyAsync(d_out.p, outputs, n_out * 4, cudaMemcpyHostToDevice, st));
  const unsigned nb = (unsigned)((n + 255) / 256);
  node_flags_kernel<<<nb, 256, 0, st>>>(d_type.as<uint8_t>(), n, d_isc.as<uint32_t>(), d_isg.as<uint32_t>());
  c->launches++;
  size_t tmp_bytes = 0, need = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_isc.as<uint32_t>(), d_cexcl.as<uint32_t>(), n, st);
  // entries: 3 per gate and per output
  // (sizes known only after the gate count: two-phase; the sort's temporary is sized below)
  CK(d_tmp.alloc(tmp_bytes));
  CK(cub::DeviceScan::ExclusiveSum(d_tmp.p, tmp_bytes, d_isc.as<uint32_t>(), d_cexcl.as<uint32_t>(), n, st));
  CK(cub::DeviceScan::ExclusiveSum(d_tmp.p, tmp_bytes, d_isg.as<uint32_t>(), d_gexcl.as<uint32_t>(), n, st));
  c->launches += 2;
  uint32_t last[4];  // cexcl[n-1], isc[n-1], gexcl[n-1], isg[n-1]
  CK(cudaMemcpyAsync(&last[0], d_cexcl.as<uint32_t>() + (n - 1), 4, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(&last[1], d_isc.as<uint32_t>() + (n - 1), 4, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(&last[2], d_gexcl.as<uint32_t>() + (n - 1), 4, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(&last[3], d_isg.as<uint32_t>() + (n - 1), 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  const size_t n_consts = (size_t)last[0] + last[1], n_gates = (size_t)last[2] + last[3];
  const size_t n_rows_nodes = n - n_consts;  // every node but the dropped constants owns a row of each P block
  if (n_rows_nodes + n_out > mk) return set_error(c, ERR_STATE, "internal: more rows than m*k");
  const size_t nnz = 3 * (n_gates + n_out);
  if (nnz > 0xfffffff0ull) return set_error(c, ERR_UNSUPPORTED, "too many constraint entries for 32-bit offsets");

  lg_constraints* a = new (std::nothrow) lg_constraints();
  if (!a) return ERR_NOMEM;
  a->owner = ctx;
  a->mk = mk;
  a->nnz = nnz;
  a->n_consts = n_table;
  auto fail_free = [&](int code, const std::string& msg) {
    lg_constraints_free(a);
    return set_error(c, code, msg);
  };
#define CKA(expr)                                                                                         \
  do {                                                                                                    \
    cudaError_t _e = (expr);                                                                              \
    if (_e != cudaSuccess) return fail_free(_e == cudaErrorMemoryAllocation ? ERR_NOMEM : ERR_CUDA,        \
                                            std::string("constraint builder: ") + cudaGetErrorString(_e)); \
  } while (0)
  CKA(cudaMalloc(&a->col_ptr, (mk + 1) * 4));
  CKA(cudaMalloc(&a->row_idx, (nnz ? nnz : 1) * 4));
  CKA(cudaMalloc(&a->val_id, (nnz ? nnz : 1) * 4));
  CKA(cudaMalloc(&a->consts, (n_table ? n_table : 1) * sizeof(Fr)));
  if (n_table) CKA(cudaMemcpyAsync(a->consts, const_table, n_table * sizeof(Fr), cudaMemcpyHostToDevice, st));
  Buf cols, rows, vids, cols2, perm, perm2, counts, sort_tmp;
  CKA(cols.alloc(nnz * 4));
  CKA(rows.alloc(nnz * 4));
  CKA(vids.alloc(nnz * 4));
  CKA(cols2.alloc(nnz * 4));
  CKA(perm.alloc(nnz * 4));
  CKA(perm2.alloc(nnz * 4));
  CKA(counts.alloc((mk + 1) * 4));
  CKA(cudaMemsetAsync(counts.p, 0, (mk + 1) * 4, st));
  NodeArrays na{d_type.as<uint8_t>(), d_l.as<uint32_t>(), d_r.as<uint32_t>(), d_cexcl.as<uint32_t>(), d_vidp.as<uint32_t>(),
                d_vidn.as<uint32_t>()};
  emit_nodes_kernel<<<nb, 256, 0, st>>>(na, d_gexcl.as<uint32_t>(), n, (uint32_t)mk, cols.as<uint32_t>(), rows.as<uint32_t>(),
                                        vids.as<uint32_t>(), counts.as<uint32_t>());
  c->launches++;
  if (n_out) {
    emit_outputs_kernel<<<(unsigned)((n_out + 255) / 256), 256, 0, st>>>(na, d_out.as<uint32_t>(), n_out, (uint32_t)n_rows_nodes, 3 * n_gates,
                                                                        (uint32_t)mk, cols.as<uint32_t>(), rows.as<uint32_t>(),
                                                                        vids.as<uint32_t>(), counts.as<uint32_t>());
    c->launches++;
  }
  if (nnz) {
    iota_kernel<<<(unsigned)((nnz + 255) / 256), 256, 0, st>>>(perm.as<uint32_t>(), nnz);
    c->launches++;
    int end_bit = 1;
    while (((size_t)1 << end_bit) < mk) end_bit++;
    cub::DeviceRadixSort::SortPairs(nullptr, need, cols.as<uint32_t>(), cols2.as<uint32_t>(), perm.as<uint32_t>(), perm2.as<uint32_t>(), nnz, 0,
                                    end_bit, st);
    CKA(sort_tmp.alloc(need));
    CKA(cub::DeviceRadixSort::SortPairs(sort_tmp.p, need, cols.as<uint32_t>(), cols2.as<uint32_t>(), perm.as<uint32_t>(), perm2.as<uint32_t>(),
                                        nnz, 0, end_bit, st));  // stable: entries of a column keep the reference's order
    c->launches++;
    gather2_kernel<<<(unsigned)((nnz + 255) / 256), 256, 0, st>>>(perm2.as<uint32_t>(), rows.as<uint32_t>(), vids.as<uint32_t>(), nnz, a->row_idx,
                                                                 a->val_id);
    c->launches++;
  }
  // col_ptr = inclusive prefix sums of the per-column counts (counts[0] = 0)
  size_t scan_bytes = 0;
  cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, counts.as<uint32_t>(), a->col_ptr, mk + 1, st);
  Buf scan_tmp;
  CKA(scan_tmp.alloc(scan_bytes));
  CKA(cub::DeviceScan::InclusiveSum(scan_tmp.p, scan_bytes, counts.as<uint32_t>(), a->col_ptr, mk + 1, st));
  c->launches++;
  CKA(cudaGetLastError());
  CKA(cudaStreamSynchronize(st));
#undef CK
#undef CKA
  *out = a;
  return OK;
}

}  // namespace lg

extern "C" int lg_constraints_read(const lg_constraints* a, size_t* mk, size_t* nnz, size_t* n_consts, uint32_t* col_ptr, uint32_t* row_idx,
                                   uint32_t* val_id, uint64_t* const_table) {
  if (!a) return lg::ERR_INVALID;
  lg::Ctx* c = &a->owner->c;
  cudaSetDevice(c->device);
  if (mk) *mk = a->mk;
  if (nnz) *nnz = a->nnz;
  if (n_consts) *n_consts = a->n_consts;
  if (col_ptr) LG_CUDA(c, cudaMemcpy(col_ptr, a->col_ptr, (a->mk + 1) * 4, cudaMemcpyDeviceToHost));
  if (row_idx && a->nnz) LG_CUDA(c, cudaMemcpy(row_idx, a->row_idx, a->nnz * 4, cudaMemcpyDeviceToHost));
  if (val_id && a->nnz) LG_CUDA(c, cudaMemcpy(val_id, a->val_id, a->nnz * 4, cudaMemcpyDeviceToHost));
  if (const_table && a->n_consts) LG_CUDA(c, cudaMemcpy(const_table, a->consts, a->n_consts * sizeof(lg::Fr), cudaMemcpyDeviceToHost));
  return lg::OK;
}
