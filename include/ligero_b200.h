/* ligero_b200 -- C ABI of the B200-native Ligero commit-and-test back end.
 *
 * Drop-in boundary for the numeric body of NP-Eng/ligero's `LigeroCircuit::prove_inner`
 * (reference: src/ligero/mod.rs:457-578).  The reference has no FFI seam of its own (pure Rust,
 * compile-time generics only); these are the entry points a Rust `extern "C"` block in
 * `src/ligero/mod.rs` would bind, one per private helper it replaces.
 *
 * Conventions
 *   - every function returns 0 on success or an LG_ERR_* code; nothing unwinds across the boundary;
 *     `lg_last_error(ctx)` gives the text of the last failure on that context;
 *   - Fr = BN254 scalar field element = `uint64_t[4]`, little-endian limbs, MONTGOMERY form -- the
 *     in-memory layout of `ark_bn254::Fr` (`Fp256<MontBackend<FrConfig,4>>`), so `&[Fr]` passes
 *     zero-copy;  "Fr[n]" below means `const uint64_t*` pointing at 4*n limbs;
 *   - input pointers may be host or device memory (detected with cudaPointerGetAttributes) unless a
 *     parameter says otherwise; output pointers are host memory unless named `*_dev`;
 *   - a context owns one CUDA stream; calls on one context are stream-ordered and the functions that
 *     return host data synchronise that stream before returning.  One context per host thread;
 *   - LIFETIMES: a context outlives everything created from it (lg_matrix, lg_constraints, lg_ligero, lg_shard): their free
 *     functions use the context's stream, so destroy the children first, the context last;
 *   - Fr inputs must be canonical Montgomery limbs (value < r), as ark_bn254::Fr always is; the host-driver entry points
 *     that take field elements (lg_circuit_constant, the var_vals of lg_prove / lg_circuit_evaluate / witness layout) return
 *     LG_ERR_INVALID otherwise; the bulk device entry points (lg_commit & co.) do not scan their input;
 *   - only the `LigeroMTTestParams` instantiation is implemented (Blake2s-256 column hash over
 *     canonical bytes, identity leaf hash, SHA-256 two-to-one: src/ligero/types.rs:15-46).
 */
#ifndef LIGERO_B200_H
#define LIGERO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LG_OK 0
#define LG_ERR_INVALID 1
#define LG_ERR_CUDA 2
#define LG_ERR_NOMEM 3
#define LG_ERR_UNSUPPORTED 4
#define LG_ERR_STATE 5

typedef struct lg_ctx lg_ctx;       /* per-GPU context: stream, twiddle/scale tables, scratch */
typedef struct lg_matrix lg_matrix; /* committed matrix: U (R x n), leaf digests and Merkle nodes resident in HBM */

/* ---- context ------------------------------------------------------------------------------- */
int lg_version(void);
int lg_ctx_create(int device, lg_ctx** out);
int lg_ctx_destroy(lg_ctx* ctx);
const char* lg_last_error(const lg_ctx* ctx);
int lg_ctx_sync(lg_ctx* ctx);
/* number of kernels this context has launched so far (bench.py's "gpu_launches") */
uint64_t lg_ctx_launches(const lg_ctx* ctx);
/* the two length-prefix conventions recalled from arkworks; default 1,1 */
int lg_ctx_set_formats(lg_ctx* ctx, int col_len_prefix, int leaf_len_prefix);
/* the stream the context launches on, as a cudaStream_t (for CUDA-event timing by the caller) */
void* lg_ctx_stream(const lg_ctx* ctx);
/* per-phase device timing: when enabled, every phase boundary records a CUDA event on the context
 * stream; lg_ctx_phase_ms synchronises, writes the accumulated milliseconds and interval counts per
 * phase (LG_PHASE_*) since the last query, and resets.  n = number of entries in the output arrays. */
#define LG_PHASE_NTT_STRIDED_INV 0 /* register-only DIF passes over HBM (rows longer than one CTA tile) */
#define LG_PHASE_NTT_LOCAL 1       /* shared-memory iNTT tail + all coset NTTs */
#define LG_PHASE_NTT_STRIDED_FWD 2 /* register-only DIT passes over HBM */
#define LG_PHASE_HASH 3            /* BLAKE2s column hashing */
#define LG_PHASE_MERKLE 4          /* SHA-256 tree */
#define LG_PHASE_EXPAND 5          /* ChaCha20 challenge expansion */
#define LG_PHASE_TESTS 6           /* row combination / linear / quadratic tests */
#define LG_PHASE_OPEN 7            /* column gather */
#define LG_PHASE_COUNT 8
int lg_ctx_set_timing(lg_ctx* ctx, int enabled);
/* lg_commit / lg_recommit of a large matrix run as a row-tile pipeline in which the column hashing of one
 * tile overlaps the encoding of the next on a second stream (default).  enabled = 0 serialises the kernels
 * (same results; used to time each kernel on its own). */
int lg_ctx_set_overlap(lg_ctx* ctx, int enabled);
/* Column hashing has two kernels with identical results: one thread per column (enough columns to fill the
 * SMs' ALU pipes) and four lanes per column (few columns, e.g. one rank's column range on 4-8 GPUs, where the
 * sequential BLAKE2s chain of a column sets the pace).  Hashes (whole or by row tiles) of at most max_columns columns use the
 * latter; 0 disables it.  Default 8192 (LG_HASH_QUAD_MAX in the environment overrides). */
int lg_ctx_set_hash_quad_max(lg_ctx* ctx, size_t max_columns);
int lg_ctx_phase_ms(lg_ctx* ctx, double* ms_out, uint64_t* count_out, int n);

/* ---- encode + commit: replaces src/ligero/mod.rs:521-551 ------------------------------------ */
/* (reed_solomon_interpolate 998-1002, reed_solomon_evaluate 1004-1008 per row; DenseMatrix::columns
 *  src/matrices/mod.rs:163-167; H::evaluate per column 536-542; create_merkle_tree 544-549; root 551)
 *   preenc_u : Fr[rows*k], row-major pre-encoding matrix [X;Y;Z;W] (host or device)
 *   k        : power of two >= 2;  rho_inv : power of two >= 2 (the reference hard-codes 8, mod.rs:284)
 *   root_out : 32 bytes (may be NULL)
 * U, the leaves and the tree stay resident on the device behind *out until lg_matrix_free. */
int lg_commit(lg_ctx* ctx, const uint64_t* preenc_u, size_t rows, size_t k, uint32_t rho_inv, uint8_t root_out[32],
              lg_matrix** out);
/* same work into an existing handle of identical shape (steady-state proving: no allocation) */
int lg_recommit(lg_matrix* m, const uint64_t* preenc_u, uint8_t root_out[32]);
int lg_matrix_free(lg_matrix* m);
int lg_matrix_dims(const lg_matrix* m, size_t* rows, size_t* k, size_t* n);
/* encode only (a2+a3), result left on the device; lg_matrix_hash then does a4-a6 on it */
int lg_encode(lg_ctx* ctx, const uint64_t* preenc_u, size_t rows, size_t k, uint32_t rho_inv, lg_matrix** out);
int lg_matrix_hash(lg_matrix* m, uint8_t root_out[32]);
/* a handle over a CALLER-OWNED device buffer of rho_inv*rows*k Fr in the plane layout
 * plane[s][i][c] = U[i][rho_inv*c + s] (multi-GPU: the row shard that is encoded into, and the column
 * shard that arrives over NVLink); lg_matrix_free leaves the buffer alone */
int lg_matrix_wrap(lg_ctx* ctx, uint64_t* u_dev, size_t rows, size_t k, uint32_t rho_inv, lg_matrix** out);
/* a2+a3 into an existing handle (stream-ordered, returns without synchronising) */
int lg_matrix_encode(lg_matrix* m, const uint64_t* preenc_u);

/* ---- multi-GPU: one process per GPU, rows sharded for encoding, column ranges for hashing ------ */
/* an un-encoded handle of the given shape (U allocated with cudaMalloc so that it can be exported) */
int lg_matrix_create(lg_ctx* ctx, size_t rows, size_t k, uint32_t rho_inv, lg_matrix** out);
/* CUDA IPC: export the U buffer of a handle / map a peer's exported buffer into this process
 * (peer access over NVLink is enabled lazily by the driver) */
int lg_ipc_export(const lg_matrix* m, uint8_t handle_out[64]);
int lg_ipc_open(lg_ctx* ctx, const uint8_t handle[64], void** ptr_out);
int lg_ipc_close(lg_ctx* ctx, void* ptr);
/* Encode this rank's rows and scatter every finished codeword element -- from inside the last NTT pass,
 * over peer memory -- into the column shard of the rank that owns its message index:
 *   msg_local : Fr[4*m_g*k], the rows {b*m + i0 + i : b < 4, i < m_g} of the 4m x k matrix, [X_g;Y_g;Z_g;W_g]
 *   shard_u   : world device pointers; shard_u[h] = U buffer of rank h's column shard, a plane-layout
 *               matrix of 4m rows x (k/world) columns (this rank's own plus lg_ipc_open'ed peers)
 *   cosets_scratch : device Fr[(rho_inv-1)*4*m_g*k], needed when k > 1024 (may be NULL otherwise)
 * Stream-ordered; the caller synchronises all ranks (e.g. an NCCL all-reduce on the context stream)
 * before hashing the shards. */
/*   plain          : 1 = the shards belong to a COMMITTED matrix (plain integers inside);
 *                    0 = keep the Montgomery form (e.g. the rows of r_a extended to the 2k domain for the linear test) */
int lg_encode_sharded(lg_ctx* ctx, const uint64_t* msg_local, size_t m_g, size_t k, uint32_t rho_inv, void* const* shard_u,
                      int world, size_t m, size_t i0, uint64_t* cosets_scratch, int plain);
/* The same, for `nrows` local rows that are the CONSECUTIVE global rows [row_base, row_base + nrows) of the
 * rows_total-row matrix: lets a rank encode its share block by block (X, then Y, Z, W). */
int lg_encode_sharded_rows(lg_ctx* ctx, const uint64_t* msg_rows, size_t nrows, size_t row_base, size_t rows_total, size_t k,
                           uint32_t rho_inv, void* const* shard_u, int world, uint64_t* cosets_scratch, int plain);
/* Column hashing in row tiles (src/ligero/mod.rs:536-542, one BLAKE2s stream per column): hash rows
 * [row0, row_end) of every column on the context's second stream, ordered after everything enqueued on the
 * context stream so far; tiles must come in row order and cover [0, rows).  lg_matrix_hash_finish builds the
 * tree (544-551) and orders the context stream after it.  Both are asynchronous unless root_out is given.
 * Lets the hashing of the rows that have arrived overlap the encoding (and NVLink delivery) of the rest. */
int lg_matrix_hash_rows(lg_matrix* m, size_t row0, size_t row_end);
int lg_matrix_hash_finish(lg_matrix* m, uint8_t root_out[32]);

/* device addresses of the resident arrays (stream-ordered interop with the caller's own kernels /
 * collectives): U in the plane layout, n*32 leaf bytes, (n-1)*32 node bytes (node 0 = root) */
void* lg_matrix_u_dev(const lg_matrix* m);
void* lg_matrix_leaves_dev(const lg_matrix* m);
void* lg_matrix_nodes_dev(const lg_matrix* m);

/* read-backs (parity tests, debugging): U rows in the reference's logical column order */
int lg_matrix_read_rows(const lg_matrix* m, size_t row0, size_t nrows, uint64_t* out /* Fr[nrows*n] */);
int lg_matrix_read_leaves(const lg_matrix* m, uint8_t* out /* n*32 */);
int lg_matrix_read_nodes(const lg_matrix* m, uint8_t* out /* (n-1)*32, heap order, node 0 = root */);

/* ---- batched inverse NTT (DensePolynomial coefficients from evaluations; small_domain.ifft) -- */
/* in/out : Fr[rows * size], natural order, includes 1/size; host or device, may alias */
int lg_intt(lg_ctx* ctx, const uint64_t* in, uint64_t* out, size_t rows, size_t size);

/* ---- challenge expansion: replaces src/utils.rs:23-55 --------------------------------------- */
/* get_field_elements_from_prng(count, seed): ChaCha20Rng(seed) -> F::rand x count, rejection sampled;
 * the accepted 254-bit integer is used as the Montgomery representation.  out: Fr[count], host or device. */
int lg_expand_fr(lg_ctx* ctx, const uint8_t seed[32], size_t count, uint64_t* out);
/* get_distinct_indices_from_prng(n, t, seed): t ascending distinct indices in [0, n) (host, tiny) */
int lg_expand_indices(const uint8_t seed[32], size_t n, size_t t, uint64_t* idx_out);

/* ---- Test-Interleaved: DenseMatrix::row_mul, src/matrices/mod.rs:138-149 (call mod.rs:658) ---- */
/* out[c] = sum_i r[i] * U_pre[i][c];  r: Fr[rows] (host or device), out: Fr[k] (host) */
int lg_row_combine(lg_matrix* m, const uint64_t* r, uint64_t* out);

/* ---- constraint matrix A = [[I, -(Px;Py;Pz)], [0, Padd]]  (src/ligero/mod.rs:296-433) ---------- */
/* Only the right-hand block (4mk rows x mk columns) is uploaded, in CSC form: column c holds entries
 * [col_ptr[c], col_ptr[c+1]) with row index row_idx[e] in [0, 4mk) and value id val_id[e]:
 * 0 -> +1, 1 -> -1, v >= 2 -> const_table[v - 2] (Fr).  One upload per circuit. */
typedef struct lg_constraints lg_constraints;
int lg_constraints_create(lg_ctx* ctx, size_t mk, const uint32_t* col_ptr, const uint32_t* row_idx, const uint32_t* val_id,
                          size_t nnz, const uint64_t* const_table, size_t n_consts, lg_constraints** out);
int lg_constraints_free(lg_constraints* a);
/* read the CSC back (parity tests: the device builder of LigeroCircuit::new against the host builder); every output
 * pointer is nullable; capacities: col_ptr mk+1, row_idx / val_id nnz, const_table Fr[n_consts] */
int lg_constraints_read(const lg_constraints* a, size_t* mk, size_t* nnz, size_t* n_consts, uint32_t* col_ptr, uint32_t* row_idx,
                        uint32_t* val_id, uint64_t* const_table);
/* SparseMatrix::row_mul, src/matrices/mod.rs:100-110 (call mod.rs:722): out = r^T A, Fr[4mk] each */
int lg_sparse_row_mul(lg_ctx* ctx, const lg_constraints* a, const uint64_t* r_linear, uint64_t* out);

/* ---- Test-Linear-Constraints polynomial: src/ligero/mod.rs:719-736 ----------------------------- */
/* q = sum_i p_i * r_i with r_i = ifft_k(rows of r^T A); coeffs_out: Fr[2k] capacity (host), *len_out =
 * number of coefficients after stripping trailing zeros (DensePolynomial semantics).  The seeded form
 * expands r_linear = get_field_elements_from_prng(4mk, seed) on the device. */
int lg_linear_test(lg_matrix* m, const lg_constraints* a, const uint64_t* r_linear, uint64_t* coeffs_out, size_t* len_out);
int lg_linear_test_seeded(lg_matrix* m, const lg_constraints* a, const uint8_t seed[32], uint64_t* coeffs_out, size_t* len_out);

/* ---- Test-Quadratic-Constraints polynomial: src/ligero/mod.rs:839-848 -------------------------- */
/* q = sum_{i<m} r[i] (p_x,i p_y,i - p_z,i);  r_quad: Fr[m] */
int lg_quadratic_test(lg_matrix* m, const uint64_t* r_quad, uint64_t* coeffs_out, size_t* len_out);

/* The two tests in pieces, for a matrix that holds only a RANGE of columns (one rank of a column-sharded
 * commitment): every rank evaluates q on the 2k-domain points of its own columns
 * (index 2c + half, natural order), the slices are concatenated in rank order, and one inverse NTT of size 2k
 * gives the coefficients with trailing zeros trimmed (src/ligero/mod.rs:723-736, 842-848).
 *   lg_linear_ra      : r_a = r_linear^T A from the 32-byte seed, into a DEVICE buffer of 4mk elements (719-722)
 *   lg_linear_evals   : r_even = this rank's columns of the rows of r_a, r_odd = the same rows extended to the odd
 *                       points of the 2k domain (both DEVICE, rows x k of this matrix, Montgomery) -> 2k' values
 *   lg_quadratic_evals: r_quad (m elements, host or device) -> 2k' values
 *   lg_poly_from_evals: `size` = 2k values in natural order (host or device) -> coefficients, *len_out <= size */
int lg_linear_ra(lg_ctx* ctx, const lg_constraints* a, const uint8_t seed[32], uint64_t* r_a_dev);
int lg_linear_evals(lg_matrix* m, const uint64_t* r_even_dev, const uint64_t* r_odd_dev, uint64_t* evals_out);
int lg_quadratic_evals(lg_matrix* m, const uint64_t* r_quad, uint64_t* evals_out);
int lg_poly_from_evals(lg_ctx* ctx, const uint64_t* evals, size_t size, uint64_t* coeffs_out, size_t* len_out);

/* ---- openings: src/ligero/mod.rs:935-955 (DenseMatrix::column + MerkleTree::generate_proof) ---- */
/* cols_out: Fr[t*rows] (column q = U[:, idx[q]]);  sib_out: t*32 bytes (leaf_sibling_hash);
 * auth_out: t*(log2(n)-1)*32 bytes, per column the sibling digests from just below the root downwards */
int lg_open(lg_matrix* m, const uint64_t* idx, size_t t, uint64_t* cols_out, uint8_t* sib_out, uint8_t* auth_out);

/* =====================================================================================================
 * Host driver: the reference's public API over the primitives above (capi_host.cu).  Same names,
 * argument meaning and failure behaviour as NP-Eng/ligero; panics become LG_ERR_* + lg_last_error.
 * ===================================================================================================== */

/* ---- ArithmeticCircuit: src/arithmetic_circuit/mod.rs ------------------------------------------- */
typedef struct lg_circuit lg_circuit;
int lg_circuit_new(lg_circuit** out);                                                      /* new, 65-72 */
int lg_circuit_free(lg_circuit* c);
const char* lg_circuit_last_error(const lg_circuit* c);
int lg_circuit_constant(lg_circuit* c, const uint64_t value[4], size_t* index_out);        /* constant, 76-84 */
int lg_circuit_new_variable(lg_circuit* c, const char* label /* NULL -> "var_N" */, size_t* index_out); /* 92-109 */
int lg_circuit_get_variable(const lg_circuit* c, const char* label, size_t* index_out);    /* 115-117 */
int lg_circuit_add(lg_circuit* c, size_t left, size_t right, size_t* index_out);           /* add, 125-131 */
int lg_circuit_mul(lg_circuit* c, size_t left, size_t right, size_t* index_out);           /* mul, 139-145 */
int lg_circuit_counts(const lg_circuit* c, size_t* nodes, size_t* constants, size_t* variables, size_t* gates); /* 38-63 */
/* node inspection: type 0 = Variable, 1 = Constant, 2 = Add, 3 = Mul */
int lg_circuit_node(const lg_circuit* c, size_t index, int* type, size_t* left, size_t* right, uint64_t value[4]);
/* evaluate_multioutput, 325-400: out_vals (capacity Fr[n_outputs]) receives the values of the DISTINCT output nodes in
 * node-index order, as the reference's filter over the trace yields them; *n_values_out (nullable) = how many */
int lg_circuit_evaluate(const lg_circuit* c, const size_t* var_idx, const uint64_t* var_vals, size_t n_vars, const size_t* outputs,
                        size_t n_outputs, uint64_t* out_vals, size_t* n_values_out);
/* from_constraint_system, 455-520.  A, B, C as ConstraintSystem::to_matrices yields them, in CSR form:
 * row_ptr[i][n_constraints+1], col_idx[i][nnz], coeffs[i] = Fr[nnz]; column 0 is the constant one and
 * n_cols = num_instance_variables + num_witness_variables.  outputs: size_t[n_constraints]. */
int lg_circuit_from_r1cs(size_t n_constraints, size_t n_cols, const uint64_t* const row_ptr[3], const uint64_t* const col_idx[3],
                         const uint64_t* const coeffs[3], lg_circuit** out, size_t* outputs);

/* ---- Fiat-Shamir sponge (host): ark-crypto-primitives PoseidonSponge ------------------------------ */
typedef struct lg_sponge lg_sponge;
/* PoseidonConfig::new(full, partial, alpha, mds, ark, rate, capacity); mds: Fr[(rate+cap)^2] row-major,
 * ark: Fr[(full+partial)*(rate+cap)] */
int lg_sponge_new(int full_rounds, int partial_rounds, uint64_t alpha, const uint64_t* mds, const uint64_t* ark, int rate, int capacity,
                  lg_sponge** out);
/* ark_poly_commit::test_sponge() with the deterministic ark_std::test_rng() (src/ligero/tests.rs:151,399) */
int lg_sponge_test(lg_sponge** out);
int lg_sponge_clone(const lg_sponge* s, lg_sponge** out);
int lg_sponge_free(lg_sponge* s);
int lg_sponge_absorb_bytes(lg_sponge* s, const uint8_t* data, size_t len); /* absorb(&Vec<u8>) */
int lg_sponge_absorb_fr(lg_sponge* s, const uint64_t* elems, size_t count); /* absorb(&Vec<F>) */
int lg_sponge_squeeze_bytes(lg_sponge* s, uint8_t* out, size_t len);

/* The R1CS half of read_constraint_system (src/reader.rs:6-19): an iden3 .r1cs v1 file image (BN254 Fr only,
 * LG_ERR_UNSUPPORTED for another prime) -> from_constraint_system.  outputs: size_t[outputs_cap >= nConstraints]; call
 * once with outputs = NULL to learn n_constraints / n_wires (returns LG_ERR_INVALID after filling them).  The witness
 * half of the reference (ark-circom's wasm witness calculator) is out of scope: the assignment is an input. */
int lg_circuit_from_r1cs_bytes(const uint8_t* data, size_t len, lg_circuit** out, size_t* outputs, size_t outputs_cap,
                               size_t* n_constraints_out, size_t* n_wires_out);
/* Seeded random Add/Mul circuit of exactly `gates` (>= 4) gates for the synthetic benchmark configurations (SURVEY 8d:
 * two variables, fair-coin gate types, depth O(log gates), every node feeds the single Add output of value 1, no gate
 * with two constant operands; sol_len = gates + 4).  Not part of the reference, whose tests build circuits by hand.
 * output: the output node; var_idx[2] / var_vals[8]: the satisfying assignment (Montgomery limbs). */
int lg_circuit_synthetic(size_t gates, uint64_t seed, lg_circuit** out, size_t* output, size_t var_idx[2], uint64_t var_vals[8]);

/* ---- LigeroCircuit: src/ligero/mod.rs ---------------------------------------------------------------- */
typedef struct lg_ligero lg_ligero;
typedef struct lg_proof lg_proof;
/* LigeroCircuit::new(circuit, outputs, lambda), 147-228 (the circuit is copied; constraint matrix A is
 * built and uploaded once).  LG_ERR_UNSUPPORTED for gates with two constant operands (the reference panics). */
int lg_ligero_new(lg_ctx* ctx, const lg_circuit* circuit, const size_t* outputs, size_t n_outputs, size_t lambda, lg_ligero** out);
int lg_ligero_free(lg_ligero* l);
/* A LigeroCircuit keeps the device buffers of its last proof (codeword matrix, leaves, tree, [X;Y;Z;W]) so that the next
 * proof of the same circuit re-encodes into them, and the verifier's scratch (r_a in the [X;Y;Z;W] block, one tile of
 * 1024 encoded rows, the opened columns: 6.3 GiB at 2^24 gates) so that lg_verify allocates nothing after its first call;
 * this returns all of them to the driver (lg_ligero_free does so too). */
int lg_ligero_release_buffers(lg_ligero* l);
int lg_ligero_params(const lg_ligero* l, size_t* m, size_t* k, size_t* n, size_t* t, size_t* sol_len);
/* witness layout of prove_inner, 476-516: out = Fr[4*m*k] (host) = [X;Y;Z;W].  bump != 0: indices refer to
 * the caller's circuit (as in `prove`), 0: to the formatted circuit (as in `prove_inner`). */
int lg_ligero_witness_matrix(lg_ligero* l, const size_t* var_idx, const uint64_t* var_vals, size_t n_vars, int bump, uint64_t* out);
/* The same matrix produced in HBM: evaluation_trace_multioutput (src/arithmetic_circuit/mod.rs:325-358)
 * as a level-by-level device evaluation that writes every value straight into its slot of [X;Y;Z;W];
 * out_dev = Fr[4*m*k] on the device.  Only the variable assignment crosses PCIe.  Same failure behaviour as above. */
int lg_ligero_witness_matrix_dev(lg_ligero* l, const size_t* var_idx, const uint64_t* var_vals, size_t n_vars, int bump,
                                 uint64_t* out_dev);
/* Where lg_prove runs the trace: -1 (default) = on the device when the circuit is wide (>= 2^16 gates and >= 256 gates
 * per level on average; a level costs a launch or a CTA barrier, so deep thin circuits are evaluated by the host loop),
 * 0 = host evaluator + upload, 1 = device.  The resulting matrix is the same either way. */
int lg_ligero_set_trace_mode(lg_ligero* l, int mode);
/* gates, levels, kernel launches of one device trace, and whether lg_prove currently uses it */
int lg_ligero_trace_info(const lg_ligero* l, size_t* gates, size_t* levels, size_t* launches, int* on_device);
/* host wall clock of the last successful lg_prove / lg_prove_matrix on this circuit, ms: [0] trace + layout (lg_prove
 * only), [1] commit, [2] interleaved test, [3] linear test, [4] quadratic test, [5] the three openings, [6] whole call */
int lg_ligero_prove_ms(const lg_ligero* l, double ms_out[7]);
/* prove, 435-455 (bump != 0) / prove_inner, 457-578 (bump == 0); the sponge is advanced in place */
int lg_prove(lg_ligero* l, const size_t* var_idx, const uint64_t* var_vals, size_t n_vars, int bump, lg_sponge* sponge, lg_proof** out);
/* prove_with_labels, 580-611 */
int lg_prove_with_labels(lg_ligero* l, const char* const* labels, const uint64_t* var_vals, size_t n_vars, lg_sponge* sponge,
                         lg_proof** out);
/* the commit-and-test transcript on a ready pre-encoding matrix (Fr[4*m*k], host or device) */
int lg_prove_matrix(lg_ligero* l, const uint64_t* preenc_u, lg_sponge* sponge, lg_proof** out);
/* verify, 613-644: *accepted = 1 iff every check passes (0 otherwise; errors only for resource failures).  Uses the
 * scratch of `l` (see lg_ligero_release_buffers): one lg_prove / lg_verify at a time per lg_ligero, like everything on a ctx. */
int lg_verify(lg_ligero* l, const lg_proof* proof, lg_sponge* sponge, int* accepted);
int lg_proof_free(lg_proof* p);
/* borrowed handle of the circuit's constraint matrix A on the device (owned by the lg_ligero) */
int lg_ligero_constraints(const lg_ligero* l, const lg_constraints** out);
/* A proof from its parts (LigeroProof, src/ligero/mod.rs:96-144), for provers that run the transcript
 * themselves (the multi-GPU prover): three openings in the order interleaved, linear, quadratic; each has t
 * columns of `rows` elements, t leaf indices, t sibling digests and t x depth path digests (root side first). */
int lg_proof_assemble(const uint8_t root[32], const uint64_t* preenc_u_lc, size_t k, const uint64_t* linear_poly,
                      size_t linear_len, const uint64_t* quadratic_poly, size_t quadratic_len, size_t t, size_t rows,
                      size_t depth, const uint64_t* const cols[3], const uint64_t* const idx[3],
                      const uint8_t* const sib[3], const uint8_t* const auth[3], lg_proof** out);
/* Proof wire format (the reference defines none: LigeroProof has no derives, mod.rs:96-144): the layout
 * arkworks' CanonicalSerialize would give the same structs -- little endian, Vec<T> = u64 length + items,
 * Fr = 32 canonical bytes, digests = Vec<u8>, Path = {leaf_sibling_hash, auth_path, leaf_index: u64}:
 *   u_root | preenc_u_lc, columns, paths | linear poly coeffs, columns, paths | quadratic poly coeffs, columns, paths
 * buf may be NULL to query the length. */
int lg_proof_serialize(const lg_proof* p, uint8_t* buf, size_t cap, size_t* len_out);
int lg_proof_deserialize(const uint8_t* buf, size_t len, lg_proof** out);

/* =====================================================================================================
 * Multi-GPU prover behind the boundary (capi_shard.cu).  The reference's prover is one call,
 * LigeroCircuit::prove (src/ligero/mod.rs:435-455); these entry points let that one call reach G GPUs.
 *
 * lg_shard = one GPU's part: it encodes a share of the rows of [X;Y;Z;W] and owns the columns
 * [rank*n/G, (rank+1)*n/G) of the committed matrix (hash, Merkle subtree, the three tests and the openings of
 * those columns).  Ranks exchange data only through peer-mapped device memory over NVLink (codeword elements from
 * inside the encode kernels; roots, test evaluations and authentication paths through per-rank mailboxes with
 * device-side flags), so the same code runs as one process per GPU (CUDA IPC handles carried by the host's
 * launcher: lg_shard_handles / lg_shard_connect) and as one process driving every GPU (lg_mgpu_*).
 * Every rank of a group must make the same sequence of lg_shard_commit* / lg_shard_prove* calls.
 * ===================================================================================================== */
typedef struct lg_shard lg_shard;
/* m, k, rho_inv: shape of the commitment (4m rows); t_max: openings per test (0: commitment only, no prover
 * buffers); sub_blocks >= 1: each of the X, Y, Z, W blocks is encoded in that many pipeline steps (see
 * lg_shard_layout), 0: chosen by world size (4 at 8 GPUs, else 1).  world a power of two <= 8, k divisible by world. */
int lg_shard_create(lg_ctx* ctx, size_t m, size_t k, uint32_t rho_inv, int rank, int world, size_t t_max, int sub_blocks,
                    lg_shard** out);
int lg_shard_free(lg_shard* s);
/* three CUDA IPC handles (codeword shard, r-hat shard, mailbox), 64 bytes each */
int lg_shard_handles(lg_shard* s, uint8_t out[192]);
/* all_handles: world x 192 bytes in rank order (this rank's own entry is ignored) */
int lg_shard_connect(lg_shard* s, const uint8_t* all_handles);
/* the same for shards that live in ONE process (enables peer access between their devices) */
int lg_shard_connect_local(lg_shard* const* shards, int world);
/* hash the row blocks that have arrived behind the encoding of the next: 0 off, 1 eager (beside the next block's
 * shared-memory kernel), 2 deferred (beside the next block's last strided pass), -1 (default) by world size: eager at
 * 8 GPUs, where a rank's column hash is a latency chain, off below (LG_SHARD_PIPELINE in the environment overrides).
 * Every rank of a group must use the same setting. */
int lg_shard_set_pipeline(lg_shard* s, int enabled);
/* This rank's rows: 4*sub_blocks runs of consecutive global rows, in the order its local matrix stores them
 * (with sub_blocks = 1: [X_g; Y_g; Z_g; W_g]).  row_base / nrows: capacity 4*sub_blocks each, nullable. */
int lg_shard_layout(const lg_shard* s, size_t* rows_local, size_t* n_runs, size_t* row_base, size_t* nrows);
/* the rank's column shard as a matrix handle (rows x k/world columns; read-backs, tests) */
lg_matrix* lg_shard_matrix(lg_shard* s);
/* encode + commit (src/ligero/mod.rs:521-551) of this rank's rows, Fr[rows_local * k], host or device.
 * _async enqueues everything and returns; lg_shard_root synchronises and folds the G subtree roots into the root. */
int lg_shard_commit_async(lg_shard* s, const uint64_t* msg_local);
int lg_shard_root(lg_shard* s, uint8_t root_out[32], uint8_t* subtree_roots_out /* world*32, nullable */);
int lg_shard_commit(lg_shard* s, const uint64_t* msg_local, uint8_t root_out[32]);
/* prove_inner (457-578) on this rank's rows of a ready pre-encoding matrix / prove (435-455) from the variable
 * assignment (the evaluation trace runs replicated on every GPU).  Every rank advances its own copy of the sponge
 * through the same transcript and gets the same proof; out may be NULL for ranks whose proof nobody reads. */
int lg_shard_prove_matrix(lg_shard* s, lg_ligero* l, const uint64_t* local_rows, lg_sponge* sponge, lg_proof** out);
int lg_shard_prove(lg_shard* s, lg_ligero* l, const size_t* var_idx, const uint64_t* var_vals, size_t n_vars, int bump,
                   lg_sponge* sponge, lg_proof** out);
/* host wall clock of the last lg_shard_prove_matrix, ms: [0] commit, [1] whole call */
int lg_shard_last_ms(const lg_shard* s, double ms_out[4]);

/* One process, G GPUs: what a Rust `LigeroCircuit::prove` binds to reach the whole box in one call. */
typedef struct lg_mgpu lg_mgpu;
typedef struct lg_mligero lg_mligero;
int lg_mgpu_create(const int* dev_ids /* NULL: 0..n_dev-1 */, int n_dev, lg_mgpu** out);
int lg_mgpu_destroy(lg_mgpu* g);
const char* lg_mgpu_last_error(const lg_mgpu* g);
lg_ctx* lg_mgpu_ctx(lg_mgpu* g, int i);
/* encode + commit of a whole rows x k host matrix over all GPUs (root only) */
int lg_mgpu_commit(lg_mgpu* g, const uint64_t* preenc_u, size_t rows, size_t k, uint32_t rho_inv, uint8_t root_out[32]);
/* LigeroCircuit::new on every GPU + the shards of its commitment */
int lg_mgpu_ligero_new(lg_mgpu* g, const lg_circuit* circuit, const size_t* outputs, size_t n_outputs, size_t lambda, lg_mligero** out);
int lg_mgpu_ligero_free(lg_mligero* ml);
/* LigeroCircuit::prove / prove_inner over all GPUs: same proof bytes as lg_prove, same sponge advance */
int lg_mgpu_prove(lg_mligero* ml, const size_t* var_idx, const uint64_t* var_vals, size_t n_vars, int bump, lg_sponge* sponge,
                  lg_proof** out);

/* ---- measured integer roofline ----------------------------------------------------------------- */
/* Runs dependent-chain microbenchmarks at full occupancy for ~`ms_target` milliseconds each and
 * reports sustained Montgomery multiplications/s and IMAD.WIDE.U32 (32x32+64) operations/s. */
int lg_bench_int_peak(lg_ctx* ctx, double ms_target, double* fr_mul_per_s, double* imad_wide_per_s);
/* measured by the same lg_bench_int_peak call: the encoder's own multiplier (table-constant products with a
 * precomputed quotient, fr_lazy.cuh) and whole lazily reduced butterflies, per second per GPU */
int lg_bench_shoup_peak(lg_ctx* ctx, double* shoup_mul_per_s, double* butterfly_per_s);

#ifdef __cplusplus
}
#endif
#endif /* LIGERO_B200_H */
