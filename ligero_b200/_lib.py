"""ctypes binding of include/ligero_b200.h.  Loading fails loudly: there is no CPU fallback."""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_size_t, c_uint8, c_uint32, c_uint64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libligero_b200.so")

_lib = None


class LigeroB200Error(RuntimeError):
    pass


_ERR_NAMES = {1: "LG_ERR_INVALID", 2: "LG_ERR_CUDA", 3: "LG_ERR_NOMEM", 4: "LG_ERR_UNSUPPORTED", 5: "LG_ERR_STATE"}

u64p = POINTER(c_uint64)
u8p = POINTER(c_uint8)

# name -> (restype, argtypes): exactly the symbols include/ligero_b200.h declares
SIGNATURES = {
    "lg_version": (c_int, []),
    "lg_ctx_create": (c_int, [c_int, POINTER(c_void_p)]),
    "lg_ctx_destroy": (c_int, [c_void_p]),
    "lg_last_error": (c_char_p, [c_void_p]),
    "lg_ctx_sync": (c_int, [c_void_p]),
    "lg_ctx_launches": (c_uint64, [c_void_p]),
    "lg_ctx_set_formats": (c_int, [c_void_p, c_int, c_int]),
    "lg_ctx_stream": (c_void_p, [c_void_p]),
    "lg_ctx_set_timing": (c_int, [c_void_p, c_int]),
    "lg_ctx_set_overlap": (c_int, [c_void_p, c_int]),
    "lg_ctx_set_hash_quad_max": (c_int, [c_void_p, c_size_t]),
    "lg_circuit_from_r1cs_bytes": (c_int, [c_void_p, c_size_t, POINTER(c_void_p), c_void_p, c_size_t, POINTER(c_size_t),
                                           POINTER(c_size_t)]),
    "lg_circuit_synthetic": (c_int, [c_size_t, c_uint64, POINTER(c_void_p), POINTER(c_size_t), POINTER(c_size_t), c_void_p]),
    "lg_ligero_witness_matrix_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
    "lg_ligero_prove_ms": (c_int, [c_void_p, POINTER(c_double)]),
    "lg_ligero_release_buffers": (c_int, [c_void_p]),
    "lg_ligero_set_trace_mode": (c_int, [c_void_p, c_int]),
    "lg_ligero_trace_info": (c_int, [c_void_p, POINTER(c_size_t), POINTER(c_size_t), POINTER(c_size_t), POINTER(c_int)]),
    "lg_ctx_phase_ms": (c_int, [c_void_p, POINTER(c_double), POINTER(c_uint64), c_int]),
    "lg_commit": (c_int, [c_void_p, c_void_p, c_size_t, c_size_t, c_uint32, c_void_p, POINTER(c_void_p)]),
    "lg_recommit": (c_int, [c_void_p, c_void_p, c_void_p]),
    "lg_matrix_free": (c_int, [c_void_p]),
    "lg_matrix_dims": (c_int, [c_void_p, POINTER(c_size_t), POINTER(c_size_t), POINTER(c_size_t)]),
    "lg_encode": (c_int, [c_void_p, c_void_p, c_size_t, c_size_t, c_uint32, POINTER(c_void_p)]),
    "lg_matrix_hash": (c_int, [c_void_p, c_void_p]),
    "lg_matrix_wrap": (c_int, [c_void_p, c_void_p, c_size_t, c_size_t, c_uint32, POINTER(c_void_p)]),
    "lg_matrix_encode": (c_int, [c_void_p, c_void_p]),
    "lg_matrix_create": (c_int, [c_void_p, c_size_t, c_size_t, c_uint32, POINTER(c_void_p)]),
    "lg_ipc_export": (c_int, [c_void_p, c_void_p]),
    "lg_ipc_open": (c_int, [c_void_p, c_void_p, POINTER(c_void_p)]),
    "lg_ipc_close": (c_int, [c_void_p, c_void_p]),
    "lg_encode_sharded": (c_int, [c_void_p, c_void_p, c_size_t, c_size_t, c_uint32, POINTER(c_void_p), c_int, c_size_t,
                                  c_size_t, c_void_p, c_int]),
    "lg_encode_sharded_rows": (c_int, [c_void_p, c_void_p, c_size_t, c_size_t, c_size_t, c_size_t, c_uint32,
                                       POINTER(c_void_p), c_int, c_void_p, c_int]),
    "lg_linear_ra": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "lg_linear_evals": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "lg_quadratic_evals": (c_int, [c_void_p, c_void_p, c_void_p]),
    "lg_poly_from_evals": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, POINTER(c_size_t)]),
    "lg_ligero_constraints": (c_int, [c_void_p, POINTER(c_void_p)]),
    "lg_proof_assemble": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_void_p, c_size_t, c_size_t, c_size_t,
                                  c_size_t, POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p),
                                  POINTER(c_void_p)]),
    "lg_matrix_hash_rows": (c_int, [c_void_p, c_size_t, c_size_t]),
    "lg_matrix_hash_finish": (c_int, [c_void_p, c_void_p]),
    "lg_matrix_u_dev": (c_void_p, [c_void_p]),
    "lg_matrix_leaves_dev": (c_void_p, [c_void_p]),
    "lg_matrix_nodes_dev": (c_void_p, [c_void_p]),
    "lg_matrix_read_rows": (c_int, [c_void_p, c_size_t, c_size_t, c_void_p]),
    "lg_matrix_read_leaves": (c_int, [c_void_p, c_void_p]),
    "lg_matrix_read_nodes": (c_int, [c_void_p, c_void_p]),
    "lg_expand_fr": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "lg_expand_indices": (c_int, [c_void_p, c_size_t, c_size_t, c_void_p]),
    "lg_row_combine": (c_int, [c_void_p, c_void_p, c_void_p]),
    "lg_constraints_create": (c_int, [c_void_p, c_size_t, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_size_t,
                                      POINTER(c_void_p)]),
    "lg_constraints_free": (c_int, [c_void_p]),
    "lg_constraints_read": (c_int, [c_void_p, POINTER(c_size_t), POINTER(c_size_t), POINTER(c_size_t), c_void_p, c_void_p, c_void_p,
                                    c_void_p]),
    "lg_sparse_row_mul": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "lg_linear_test": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, POINTER(c_size_t)]),
    "lg_linear_test_seeded": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, POINTER(c_size_t)]),
    "lg_quadratic_test": (c_int, [c_void_p, c_void_p, c_void_p, POINTER(c_size_t)]),
    "lg_open": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p]),
    "lg_circuit_new": (c_int, [POINTER(c_void_p)]),
    "lg_circuit_free": (c_int, [c_void_p]),
    "lg_circuit_last_error": (c_char_p, [c_void_p]),
    "lg_circuit_constant": (c_int, [c_void_p, c_void_p, POINTER(c_size_t)]),
    "lg_circuit_new_variable": (c_int, [c_void_p, c_char_p, POINTER(c_size_t)]),
    "lg_circuit_get_variable": (c_int, [c_void_p, c_char_p, POINTER(c_size_t)]),
    "lg_circuit_add": (c_int, [c_void_p, c_size_t, c_size_t, POINTER(c_size_t)]),
    "lg_circuit_mul": (c_int, [c_void_p, c_size_t, c_size_t, POINTER(c_size_t)]),
    "lg_circuit_counts": (c_int, [c_void_p, POINTER(c_size_t), POINTER(c_size_t), POINTER(c_size_t), POINTER(c_size_t)]),
    "lg_circuit_node": (c_int, [c_void_p, c_size_t, POINTER(c_int), POINTER(c_size_t), POINTER(c_size_t), c_void_p]),
    "lg_circuit_evaluate": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_void_p, POINTER(c_size_t)]),
    "lg_circuit_from_r1cs": (c_int, [c_size_t, c_size_t, POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p),
                                     POINTER(c_void_p), c_void_p]),
    "lg_sponge_new": (c_int, [c_int, c_int, c_uint64, c_void_p, c_void_p, c_int, c_int, POINTER(c_void_p)]),
    "lg_sponge_test": (c_int, [POINTER(c_void_p)]),
    "lg_sponge_clone": (c_int, [c_void_p, POINTER(c_void_p)]),
    "lg_sponge_free": (c_int, [c_void_p]),
    "lg_sponge_absorb_bytes": (c_int, [c_void_p, c_void_p, c_size_t]),
    "lg_sponge_absorb_fr": (c_int, [c_void_p, c_void_p, c_size_t]),
    "lg_sponge_squeeze_bytes": (c_int, [c_void_p, c_void_p, c_size_t]),
    "lg_ligero_new": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_size_t, POINTER(c_void_p)]),
    "lg_ligero_free": (c_int, [c_void_p]),
    "lg_ligero_params": (c_int, [c_void_p, POINTER(c_size_t), POINTER(c_size_t), POINTER(c_size_t), POINTER(c_size_t),
                                 POINTER(c_size_t)]),
    "lg_ligero_witness_matrix": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
    "lg_prove": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p, POINTER(c_void_p)]),
    "lg_prove_with_labels": (c_int, [c_void_p, POINTER(c_char_p), c_void_p, c_size_t, c_void_p, POINTER(c_void_p)]),
    "lg_prove_matrix": (c_int, [c_void_p, c_void_p, c_void_p, POINTER(c_void_p)]),
    "lg_verify": (c_int, [c_void_p, c_void_p, c_void_p, POINTER(c_int)]),
    "lg_proof_free": (c_int, [c_void_p]),
    "lg_proof_serialize": (c_int, [c_void_p, c_void_p, c_size_t, POINTER(c_size_t)]),
    "lg_proof_deserialize": (c_int, [c_void_p, c_size_t, POINTER(c_void_p)]),
    "lg_shard_create": (c_int, [c_void_p, c_size_t, c_size_t, c_uint32, c_int, c_int, c_size_t, c_int, POINTER(c_void_p)]),
    "lg_shard_free": (c_int, [c_void_p]),
    "lg_shard_handles": (c_int, [c_void_p, c_void_p]),
    "lg_shard_connect": (c_int, [c_void_p, c_void_p]),
    "lg_shard_connect_local": (c_int, [POINTER(c_void_p), c_int]),
    "lg_shard_set_pipeline": (c_int, [c_void_p, c_int]),
    "lg_shard_layout": (c_int, [c_void_p, POINTER(c_size_t), POINTER(c_size_t), c_void_p, c_void_p]),
    "lg_shard_matrix": (c_void_p, [c_void_p]),
    "lg_shard_commit_async": (c_int, [c_void_p, c_void_p]),
    "lg_shard_root": (c_int, [c_void_p, c_void_p, c_void_p]),
    "lg_shard_commit": (c_int, [c_void_p, c_void_p, c_void_p]),
    "lg_shard_prove_matrix": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, POINTER(c_void_p)]),
    "lg_shard_prove": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p, POINTER(c_void_p)]),
    "lg_shard_last_ms": (c_int, [c_void_p, POINTER(c_double)]),
    "lg_mgpu_create": (c_int, [c_void_p, c_int, POINTER(c_void_p)]),
    "lg_mgpu_destroy": (c_int, [c_void_p]),
    "lg_mgpu_last_error": (c_char_p, [c_void_p]),
    "lg_mgpu_ctx": (c_void_p, [c_void_p, c_int]),
    "lg_mgpu_commit": (c_int, [c_void_p, c_void_p, c_size_t, c_size_t, c_uint32, c_void_p]),
    "lg_mgpu_ligero_new": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_size_t, POINTER(c_void_p)]),
    "lg_mgpu_ligero_free": (c_int, [c_void_p]),
    "lg_mgpu_prove": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p, POINTER(c_void_p)]),
    "lg_intt": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_size_t]),
    "lg_bench_int_peak": (c_int, [c_void_p, c_double, POINTER(c_double), POINTER(c_double)]),
    "lg_bench_shoup_peak": (c_int, [c_void_p, POINTER(c_double), POINTER(c_double)]),
}


def load():
    """dlopen libligero_b200.so (building nothing: run `python -m ligero_b200.build` first)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LigeroB200Error(
            f"{LIB_PATH} is missing: build it with `python -m ligero_b200.build` (nvcc, sm_100a). "
            "ligero_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int, ctx=None, what: str = ""):
    if status == 0:
        return
    msg = ""
    if ctx is not None:
        raw = load().lg_last_error(ctx)
        msg = raw.decode() if raw else ""
    raise LigeroB200Error(f"{what or 'ligero_b200 call'} failed: {_ERR_NAMES.get(status, status)} {msg}")
