// C ABI, part 2: challenge expansion, the three tests and the openings (include/ligero_b200.h).
#include <cstring>
#include <vector>

#include "capi_types.h"
#include "fr_host.h"
#include "host_prng.h"

using namespace lg;

namespace {

// device copy of a host-or-device input; frees itself (stream-ordered) when it owns the buffer
struct DevIn {
  Ctx* c;
  const void* ptr = nullptr;
  void* owned = nullptr;
  int init(Ctx* ctx, const void* src, size_t bytes) {
    c = ctx;
    if (is_device_ptr(src)) {
      ptr = src;
      return OK;
    }
    LG_CUDA(c, cudaMallocAsync(&owned, bytes ? bytes : 1, c->stream));
    LG_CUDA(c, cudaMemcpyAsync(owned, src, bytes, cudaMemcpyHostToDevice, c->stream));
    ptr = owned;
    return OK;
  }
  ~DevIn() {
    if (owned) cudaFreeAsync(owned, c->stream);
  }
};

struct DevBuf {
  Ctx* c = nullptr;
  void* p = nullptr;
  int alloc(Ctx* ctx, size_t bytes) {
    c = ctx;
    LG_CUDA(c, cudaMallocAsync(&p, bytes ? bytes : 1, c->stream));
    return OK;
  }
  ~DevBuf() {
    if (p) cudaFreeAsync(p, c->stream);
  }
};

__global__ void copy_fr_kernel(const Fr* __restrict__ src, Fr* __restrict__ dst, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const uint4* s = reinterpret_cast<const uint4*>(src + i);
    uint4* d = reinterpret_cast<uint4*>(dst + i);
    d[0] = s[0];
    d[1] = s[1];
  }
}

// evaluations on the 2k domain (natural order) -> coefficients on the host, trailing zeros trimmed
int finish_poly(Ctx* c, Fr* qhat_dev, int log_k, uint64_t* coeffs_out, size_t* len_out) {
  const size_t k2 = (size_t)2 << log_k;
  LG_TRY(intt_rows(c, qhat_dev, qhat_dev, 1, log_k + 1));
  std::vector<Fr> host(k2);
  LG_CUDA(c, cudaMemcpyAsync(host.data(), qhat_dev, k2 * sizeof(Fr), cudaMemcpyDeviceToHost, c->stream));
  LG_CUDA(c, cudaStreamSynchronize(c->stream));
  size_t len = k2;
  while (len > 0 && fr_is_zero(host[len - 1])) len--;  // DensePolynomial::from_coefficients_vec
  if (coeffs_out) memcpy(coeffs_out, host.data(), len * sizeof(Fr));
  if (len_out) *len_out = len;
  return OK;
}

}  // namespace

extern "C" {

int lg_expand_fr(lg_ctx* ctx, const uint8_t seed[32], size_t count, uint64_t* out) {
  if (!ctx || !seed || (!out && count)) return ERR_INVALID;
  Ctx* c = &ctx->c;
  cudaSetDevice(c->device);
  if (count == 0) return OK;
  if (is_device_ptr(out)) {
    LG_TRY(expand_fr(c, seed, count, (Fr*)out));
    LG_CUDA(c, cudaStreamSynchronize(c->stream));
    return OK;
  }
  DevBuf d;
  LG_TRY(d.alloc(c, count * sizeof(Fr)));
  LG_TRY(expand_fr(c, seed, count, (Fr*)d.p));
  LG_CUDA(c, cudaMemcpyAsync(out, d.p, count * sizeof(Fr), cudaMemcpyDeviceToHost, c->stream));
  LG_CUDA(c, cudaStreamSynchronize(c->stream));
  return OK;
}

static int lg_expand_indices_impl(const uint8_t seed[32], size_t n, size_t t, uint64_t* idx_out) {
  if (!seed || !idx_out || n == 0 || t > n) return ERR_INVALID;
  const std::vector<uint64_t> v = distinct_indices(seed, n, t);
  memcpy(idx_out, v.data(), v.size() * sizeof(uint64_t));
  return OK;
}
int lg_expand_indices(const uint8_t seed[32], size_t n, size_t t, uint64_t* idx_out) {
  return lg::guard([&]() { return lg_expand_indices_impl(seed, n, t, idx_out); });
}

int lg_row_combine(lg_matrix* h, const uint64_t* r, uint64_t* out) {
  if (!h || !r || !out) return ERR_INVALID;
  Matrix& m = h->m;
  Ctx* c = m.ctx;
  cudaSetDevice(c->device);
  DevIn rin;
  LG_TRY(rin.init(c, r, m.rows * sizeof(Fr)));
  DevBuf res;
  LG_TRY(res.alloc(c, m.k * sizeof(Fr)));
  phase_mark(c, PH_BEGIN);
  LG_TRY(col_reduce(c, 0, (const Fr*)rin.ptr, m.u, nullptr, nullptr, m.rows, m.k, (Fr*)res.p, 1, 0, true));  // plane 0 = U_pre
  phase_mark(c, PH_TESTS);
  LG_CUDA(c, cudaMemcpyAsync(out, res.p, m.k * sizeof(Fr), cudaMemcpyDeviceToHost, c->stream));
  LG_CUDA(c, cudaStreamSynchronize(c->stream));
  return OK;
}

static int lg_constraints_create_impl(lg_ctx* ctx, size_t mk, const uint32_t* col_ptr, const uint32_t* row_idx, const uint32_t* val_id,
                          size_t nnz, const uint64_t* const_table, size_t n_consts, lg_constraints** out) {
  if (!ctx || !out || !col_ptr || mk == 0 || (nnz && (!row_idx || !val_id)) || (n_consts && !const_table)) return ERR_INVALID;
  Ctx* c = &ctx->c;
  cudaSetDevice(c->device);
  if (col_ptr[0] != 0 || col_ptr[mk] != nnz) return set_error(c, ERR_INVALID, "col_ptr must run from 0 to nnz");
  if (4 * mk > 0xffffffffull) return set_error(c, ERR_INVALID, "constraint matrix too large for 32-bit row indices");
  for (size_t e = 0; e < nnz; e++) {
    if (row_idx[e] >= 4 * mk) return set_error(c, ERR_INVALID, "row index out of range");
    if (val_id[e] >= n_consts + 2) return set_error(c, ERR_INVALID, "value id out of range");
  }
  lg_constraints* a = new (std::nothrow) lg_constraints();
  if (!a) return ERR_NOMEM;
  a->owner = ctx;
  a->mk = mk;
  a->nnz = nnz;
  a->n_consts = n_consts;
  cudaError_t e = cudaMalloc(&a->col_ptr, (mk + 1) * 4);
  if (e == cudaSuccess) e = cudaMalloc(&a->row_idx, (nnz ? nnz : 1) * 4);
  if (e == cudaSuccess) e = cudaMalloc(&a->val_id, (nnz ? nnz : 1) * 4);
  if (e == cudaSuccess) e = cudaMalloc(&a->consts, (n_consts ? n_consts : 1) * sizeof(Fr));
  if (e == cudaSuccess) e = cudaMemcpyAsync(a->col_ptr, col_ptr, (mk + 1) * 4, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess && nnz) e = cudaMemcpyAsync(a->row_idx, row_idx, nnz * 4, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess && nnz) e = cudaMemcpyAsync(a->val_id, val_id, nnz * 4, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess && n_consts)
    e = cudaMemcpyAsync(a->consts, const_table, n_consts * sizeof(Fr), cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  if (e != cudaSuccess) {
    lg_constraints_free(a);
    return set_error(c, ERR_CUDA, std::string("constraint upload: ") + cudaGetErrorString(e));
  }
  *out = a;
  return OK;
}
int lg_constraints_create(lg_ctx* ctx, size_t mk, const uint32_t* col_ptr, const uint32_t* row_idx, const uint32_t* val_id,
                          size_t nnz, const uint64_t* const_table, size_t n_consts, lg_constraints** out) {
  return lg::guard([&]() { return lg_constraints_create_impl(ctx, mk, col_ptr, row_idx, val_id, nnz, const_table, n_consts, out); });
}

int lg_constraints_free(lg_constraints* a) {
  if (!a) return OK;
  cudaSetDevice(a->owner->c.device);
  cudaFree(a->col_ptr);
  cudaFree(a->row_idx);
  cudaFree(a->val_id);
  cudaFree(a->consts);
  delete a;
  return OK;
}

// r_a = r_linear^T A (device, 4mk elements).  r_lin_dev is overwritten in its last mk entries.
static int compute_r_a_inplace(Ctx* c, const lg_constraints* a, Fr* r_lin_dev) {
  DevBuf last;
  LG_TRY(last.alloc(c, a->mk * sizeof(Fr)));
  LG_TRY(spmv_right_block(c, a->col_ptr, a->row_idx, a->val_id, a->consts, r_lin_dev, a->mk, (Fr*)last.p));
  copy_fr_kernel<<<c->sm_count * 4, 256, 0, c->stream>>>((const Fr*)last.p, r_lin_dev + 3 * a->mk, a->mk);
  c->launches++;
  LG_CUDA(c, cudaGetLastError());
  return OK;
}

int lg_sparse_row_mul(lg_ctx* ctx, const lg_constraints* a, const uint64_t* r_linear, uint64_t* out) {
  if (!ctx || !a || !r_linear || !out) return ERR_INVALID;
  Ctx* c = &ctx->c;
  cudaSetDevice(c->device);
  const size_t total = 4 * a->mk;
  DevBuf buf;
  LG_TRY(buf.alloc(c, total * sizeof(Fr)));
  LG_CUDA(c, cudaMemcpyAsync(buf.p, r_linear, total * sizeof(Fr), cudaMemcpyDefault, c->stream));
  LG_TRY(compute_r_a_inplace(c, a, (Fr*)buf.p));
  LG_CUDA(c, cudaMemcpyAsync(out, buf.p, total * sizeof(Fr), cudaMemcpyDefault, c->stream));
  LG_CUDA(c, cudaStreamSynchronize(c->stream));
  return OK;
}

// q on the 2k domain restricted to this matrix's columns (natural order: index 2c + half), on the device:
//   q(zeta^(2c))   = sum_i r_even[i][c] U[i][rho c]             (plane 0)
//   q(zeta^(2c+1)) = sum_i r_odd[i][c]  U[i][rho c + rho/2]     (plane rho/2)
static int linear_evals_dev(Matrix& m, const Fr* r_even, const Fr* r_odd, Fr* qhat) {
  Ctx* c = m.ctx;
  const size_t plane = m.rows * m.k;
  phase_mark(c, PH_BEGIN);
  // the planes of the committed matrix hold plain integers (x_plain)
  LG_TRY(col_reduce(c, 1, r_even, m.u, nullptr, nullptr, m.rows, m.k, qhat, 2, 0, true));
  LG_TRY(col_reduce(c, 1, r_odd, m.u + (size_t)(m.rho_inv / 2) * plane, nullptr, nullptr, m.rows, m.k, qhat, 2, 1, true));
  phase_mark(c, PH_TESTS);
  return OK;
}

static int quadratic_evals_dev(Matrix& m, const Fr* r_quad, Fr* qhat) {
  Ctx* c = m.ctx;
  const size_t mm = m.rows / 4, plane = m.rows * m.k;
  phase_mark(c, PH_BEGIN);
  for (int half = 0; half < 2; half++) {
    const Fr* p = m.u + (size_t)(half ? m.rho_inv / 2 : 0) * plane;
    LG_TRY(col_reduce(c, 2, r_quad, p, p + mm * m.k, p + 2 * mm * m.k, mm, m.k, qhat, 2, half, true));
  }
  phase_mark(c, PH_TESTS);
  return OK;
}

// shared tail of the linear test once r_a is on the device (and may be clobbered)
static int linear_test_core(lg_matrix* h, Fr* r_a, uint64_t* coeffs_out, size_t* len_out) {
  Matrix& m = h->m;
  Ctx* c = m.ctx;
  const size_t plane = m.rows * m.k;
  // r_i on the odd points of the 2k domain: LDE of every row of r_a by the coset zeta = omega_{2k}
  DevBuf odd, qhat;
  LG_TRY(odd.alloc(c, plane * sizeof(Fr)));
  LG_TRY(qhat.alloc(c, 2 * m.k * sizeof(Fr)));
  LG_TRY(encode_rows(c, r_a, m.rows, m.log_k, 2, nullptr, (Fr*)odd.p));
  LG_TRY(linear_evals_dev(m, r_a, (const Fr*)odd.p, (Fr*)qhat.p));
  return finish_poly(c, (Fr*)qhat.p, m.log_k, coeffs_out, len_out);
}

static int lg_linear_test_impl(lg_matrix* h, const lg_constraints* a, const uint64_t* r_linear, uint64_t* coeffs_out, size_t* len_out) {
  if (!h || !a || !r_linear) return ERR_INVALID;
  Matrix& m = h->m;
  Ctx* c = m.ctx;
  cudaSetDevice(c->device);
  if (m.rows % 4 || a->mk * 4 != m.rows * m.k) return set_error(c, ERR_INVALID, "constraint matrix does not match the committed matrix");
  if (m.rho_inv < 2) return set_error(c, ERR_UNSUPPORTED, "the tests need the 2k domain inside the codeword (rho_inv >= 2)");
  DevBuf ra;
  LG_TRY(ra.alloc(c, 4 * a->mk * sizeof(Fr)));
  LG_CUDA(c, cudaMemcpyAsync(ra.p, r_linear, 4 * a->mk * sizeof(Fr), cudaMemcpyDefault, c->stream));
  LG_TRY(compute_r_a_inplace(c, a, (Fr*)ra.p));
  return linear_test_core(h, (Fr*)ra.p, coeffs_out, len_out);
}
int lg_linear_test(lg_matrix* h, const lg_constraints* a, const uint64_t* r_linear, uint64_t* coeffs_out, size_t* len_out) {
  return lg::guard([&]() { return lg_linear_test_impl(h, a, r_linear, coeffs_out, len_out); });
}

static int lg_linear_test_seeded_impl(lg_matrix* h, const lg_constraints* a, const uint8_t seed[32], uint64_t* coeffs_out, size_t* len_out) {
  if (!h || !a || !seed) return ERR_INVALID;
  Matrix& m = h->m;
  Ctx* c = m.ctx;
  cudaSetDevice(c->device);
  if (m.rows % 4 || a->mk * 4 != m.rows * m.k) return set_error(c, ERR_INVALID, "constraint matrix does not match the committed matrix");
  if (m.rho_inv < 2) return set_error(c, ERR_UNSUPPORTED, "the tests need the 2k domain inside the codeword (rho_inv >= 2)");
  DevBuf ra;
  LG_TRY(ra.alloc(c, 4 * a->mk * sizeof(Fr)));
  LG_TRY(expand_fr(c, seed, 4 * a->mk, (Fr*)ra.p));  // get_field_elements_from_prng(4mk) on the device
  LG_TRY(compute_r_a_inplace(c, a, (Fr*)ra.p));
  return linear_test_core(h, (Fr*)ra.p, coeffs_out, len_out);
}
int lg_linear_test_seeded(lg_matrix* h, const lg_constraints* a, const uint8_t seed[32], uint64_t* coeffs_out, size_t* len_out) {
  return lg::guard([&]() { return lg_linear_test_seeded_impl(h, a, seed, coeffs_out, len_out); });
}

static int lg_quadratic_test_impl(lg_matrix* h, const uint64_t* r_quad, uint64_t* coeffs_out, size_t* len_out) {
  if (!h || !r_quad) return ERR_INVALID;
  Matrix& m = h->m;
  Ctx* c = m.ctx;
  cudaSetDevice(c->device);
  if (m.rows % 4) return set_error(c, ERR_INVALID, "rows must be 4m");
  if (m.rho_inv < 2) return set_error(c, ERR_UNSUPPORTED, "the tests need the 2k domain inside the codeword (rho_inv >= 2)");
  const size_t mm = m.rows / 4, plane = m.rows * m.k;
  DevIn rin;
  LG_TRY(rin.init(c, r_quad, mm * sizeof(Fr)));
  DevBuf qhat;
  LG_TRY(qhat.alloc(c, 2 * m.k * sizeof(Fr)));
  LG_TRY(quadratic_evals_dev(m, (const Fr*)rin.ptr, (Fr*)qhat.p));
  return finish_poly(c, (Fr*)qhat.p, m.log_k, coeffs_out, len_out);
}
int lg_quadratic_test(lg_matrix* h, const uint64_t* r_quad, uint64_t* coeffs_out, size_t* len_out) {
  return lg::guard([&]() { return lg_quadratic_test_impl(h, r_quad, coeffs_out, len_out); });
}

// ---- the same tests in pieces, for a column-sharded matrix (one rank's columns) -----------
int lg_linear_ra(lg_ctx* ctx, const lg_constraints* a, const uint8_t seed[32], uint64_t* r_a_dev) {
  if (!ctx || !a || !seed || !r_a_dev) return ERR_INVALID;
  Ctx* c = &ctx->c;
  cudaSetDevice(c->device);
  if (!is_device_ptr(r_a_dev)) return set_error(c, ERR_INVALID, "lg_linear_ra writes a device buffer of 4mk elements");
  LG_TRY(expand_fr(c, seed, 4 * a->mk, (Fr*)r_a_dev));  // get_field_elements_from_prng(4mk) on the device
  return compute_r_a_inplace(c, a, (Fr*)r_a_dev);
}

int lg_linear_evals(lg_matrix* h, const uint64_t* r_even_dev, const uint64_t* r_odd_dev, uint64_t* evals_out) {
  if (!h || !r_even_dev || !r_odd_dev || !evals_out) return ERR_INVALID;
  Matrix& m = h->m;
  Ctx* c = m.ctx;
  cudaSetDevice(c->device);
  if (m.rho_inv < 2) return set_error(c, ERR_UNSUPPORTED, "the tests need the 2k domain inside the codeword (rho_inv >= 2)");
  if (!is_device_ptr(r_even_dev) || !is_device_ptr(r_odd_dev)) return set_error(c, ERR_INVALID, "r matrices must be device buffers");
  DevBuf qhat;
  LG_TRY(qhat.alloc(c, 2 * m.k * sizeof(Fr)));
  LG_TRY(linear_evals_dev(m, (const Fr*)r_even_dev, (const Fr*)r_odd_dev, (Fr*)qhat.p));
  LG_CUDA(c, cudaMemcpyAsync(evals_out, qhat.p, 2 * m.k * sizeof(Fr), cudaMemcpyDefault, c->stream));
  LG_CUDA(c, cudaStreamSynchronize(c->stream));
  return OK;
}

int lg_quadratic_evals(lg_matrix* h, const uint64_t* r_quad, uint64_t* evals_out) {
  if (!h || !r_quad || !evals_out) return ERR_INVALID;
  Matrix& m = h->m;
  Ctx* c = m.ctx;
  cudaSetDevice(c->device);
  if (m.rows % 4) return set_error(c, ERR_INVALID, "rows must be 4m");
  if (m.rho_inv < 2) return set_error(c, ERR_UNSUPPORTED, "the tests need the 2k domain inside the codeword (rho_inv >= 2)");
  DevIn rin;
  LG_TRY(rin.init(c, r_quad, (m.rows / 4) * sizeof(Fr)));
  DevBuf qhat;
  LG_TRY(qhat.alloc(c, 2 * m.k * sizeof(Fr)));
  LG_TRY(quadratic_evals_dev(m, (const Fr*)rin.ptr, (Fr*)qhat.p));
  LG_CUDA(c, cudaMemcpyAsync(evals_out, qhat.p, 2 * m.k * sizeof(Fr), cudaMemcpyDefault, c->stream));
  LG_CUDA(c, cudaStreamSynchronize(c->stream));
  return OK;
}

static int lg_poly_from_evals_impl(lg_ctx* ctx, const uint64_t* evals, size_t size, uint64_t* coeffs_out, size_t* len_out) {
  if (!ctx || !evals || size < 4 || (size & (size - 1))) return ERR_INVALID;
  Ctx* c = &ctx->c;
  cudaSetDevice(c->device);
  int log_k = 0;
  while (((size_t)2 << log_k) < size) log_k++;
  DevBuf q;
  LG_TRY(q.alloc(c, size * sizeof(Fr)));
  LG_CUDA(c, cudaMemcpyAsync(q.p, evals, size * sizeof(Fr), cudaMemcpyDefault, c->stream));
  return finish_poly(c, (Fr*)q.p, log_k, coeffs_out, len_out);
}
int lg_poly_from_evals(lg_ctx* ctx, const uint64_t* evals, size_t size, uint64_t* coeffs_out, size_t* len_out) {
  return lg::guard([&]() { return lg_poly_from_evals_impl(ctx, evals, size, coeffs_out, len_out); });
}

int lg_open(lg_matrix* h, const uint64_t* idx, size_t t, uint64_t* cols_out, uint8_t* sib_out, uint8_t* auth_out) {
  if (!h || (!idx && t)) return ERR_INVALID;
  Matrix& m = h->m;
  Ctx* c = m.ctx;
  cudaSetDevice(c->device);
  if (t == 0) return OK;
  for (size_t i = 0; i < t; i++)
    if (idx[i] >= m.n) return set_error(c, ERR_INVALID, "column index out of range");
  int log_n = 0;
  while (((size_t)1 << log_n) < m.n) log_n++;
  const size_t auth_bytes = t * (size_t)(log_n - 1) * 32;
  DevIn di;
  LG_TRY(di.init(c, idx, t * 8));
  DevBuf cols, sib, auth;
  LG_TRY(cols.alloc(c, t * m.rows * sizeof(Fr)));
  LG_TRY(sib.alloc(c, t * 32));
  LG_TRY(auth.alloc(c, auth_bytes));
  phase_mark(c, PH_BEGIN);
  LG_TRY(gather_open(c, m, (const uint64_t*)di.ptr, t, (Fr*)cols.p, (uint8_t*)sib.p, (uint8_t*)auth.p));
  phase_mark(c, PH_OPEN);
  if (cols_out) LG_CUDA(c, cudaMemcpyAsync(cols_out, cols.p, t * m.rows * sizeof(Fr), cudaMemcpyDeviceToHost, c->stream));
  if (sib_out) LG_CUDA(c, cudaMemcpyAsync(sib_out, sib.p, t * 32, cudaMemcpyDeviceToHost, c->stream));
  if (auth_out && auth_bytes) LG_CUDA(c, cudaMemcpyAsync(auth_out, auth.p, auth_bytes, cudaMemcpyDeviceToHost, c->stream));
  LG_CUDA(c, cudaStreamSynchronize(c->stream));
  return OK;
}

}  // extern "C"
