"""ctypes wrapper of the C oracle (oracle/ligero_ref.c).  TEST INFRASTRUCTURE ONLY: imported by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by the
product package."""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, c_double, c_int, c_size_t, c_uint, c_void_p

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libligero_ref.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "ligero_ref.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE] + (["-B"] if force else []), check=True, capture_output=True)
    return LIB_PATH


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.ref_commit.restype = c_int
        _lib.ref_commit.argtypes = [c_void_p, c_size_t, c_size_t, c_uint, c_int, c_int, c_int, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_void_p]
        _lib.ref_fft_rows.argtypes = [c_void_p, c_size_t, c_int, c_int]
        _lib.ref_fr_mul.argtypes = [c_void_p, c_void_p, c_void_p, c_size_t]
        _lib.ref_expand_fr.argtypes = [c_void_p, c_size_t, c_void_p]
        _lib.ref_expand_indices.argtypes = [c_void_p, c_size_t, c_size_t, c_void_p]
        _lib.ref_row_mul.argtypes = [c_void_p, c_void_p, c_size_t, c_size_t, c_int, c_void_p]
        _lib.ref_sparse_row_mul.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_size_t, c_void_p]
        _lib.ref_linear_poly.argtypes = [c_void_p, c_void_p, c_size_t, c_size_t, c_int, c_void_p]
        _lib.ref_quadratic_poly.argtypes = [c_void_p, c_void_p, c_size_t, c_size_t, c_void_p]
        _lib.ref_sha256.argtypes = [c_void_p, c_size_t, c_void_p]
        _lib.ref_blake2s.argtypes = [c_void_p, c_size_t, c_void_p]
    return _lib


def _p(a):
    return c_void_p(a.ctypes.data) if a is not None else c_void_p(0)


def commit(preenc_u: np.ndarray, rows: int, k: int, rho_inv: int = 8, threads: int = 0, col_prefix: bool = True,
           leaf_prefix: bool = True, want_u: bool = False, want_tree: bool = False):
    """Reference-schedule encode+commit on the CPU.  preenc_u: uint64[rows*k,4] Montgomery limbs.
    Returns dict(root, secs_encode, secs_hash[, u, leaves, nodes])."""
    lib = load()
    threads = threads or os.cpu_count() or 1
    n = k * rho_inv
    a = np.ascontiguousarray(preenc_u, dtype=np.uint64)
    u = np.empty((rows * n, 4), dtype=np.uint64) if want_u else None
    leaves = np.empty((n, 32), dtype=np.uint8) if want_tree else None
    nodes = np.empty((n - 1 if n > 1 else 1, 32), dtype=np.uint8) if want_tree else None
    nodes_buf = np.empty((n, 32), dtype=np.uint8) if want_tree else None
    root = np.zeros(32, dtype=np.uint8)
    secs = np.zeros(2, dtype=np.float64)
    st = lib.ref_commit(_p(a), rows, k, rho_inv, threads, int(col_prefix), int(leaf_prefix), _p(u), _p(leaves),
                        _p(nodes_buf), _p(root), _p(secs))
    if st != 0:
        raise ValueError(f"ref_commit failed: {st}")
    out = dict(root=bytes(root), secs_encode=float(secs[0]), secs_hash=float(secs[1]), threads=threads)
    if want_u:
        out["u"] = u.reshape(rows, n, 4)
    if want_tree:
        out["leaves"] = leaves
        out["nodes"] = nodes_buf[: n - 1]
    return out


def fft_rows(data: np.ndarray, rows: int, size: int, inverse: bool = False) -> np.ndarray:
    lib = load()
    a = np.array(data, dtype=np.uint64, copy=True).reshape(rows * size, 4)
    lib.ref_fft_rows(_p(a), rows, size.bit_length() - 1, int(inverse))
    return a


def expand_fr(seed: bytes, count: int) -> np.ndarray:
    lib = load()
    out = np.empty((count, 4), dtype=np.uint64)
    s = np.frombuffer(seed, dtype=np.uint8).copy()
    lib.ref_expand_fr(_p(s), count, _p(out))
    return out


def expand_indices(seed: bytes, n: int, t: int) -> np.ndarray:
    lib = load()
    out = np.empty(t, dtype=np.uint64)
    s = np.frombuffer(seed, dtype=np.uint8).copy()
    lib.ref_expand_indices(_p(s), n, t, _p(out))
    return out


def row_mul(m: np.ndarray, r: np.ndarray, rows: int, cols: int, threads: int = 0) -> np.ndarray:
    lib = load()
    out = np.empty((cols, 4), dtype=np.uint64)
    lib.ref_row_mul(_p(np.ascontiguousarray(m)), _p(np.ascontiguousarray(r)), rows, cols, threads or os.cpu_count() or 1, _p(out))
    return out


def sparse_row_mul(row_ptr, col_idx, vals, r, rows: int, cols: int) -> np.ndarray:
    lib = load()
    out = np.empty((cols, 4), dtype=np.uint64)
    lib.ref_sparse_row_mul(_p(np.ascontiguousarray(row_ptr, dtype=np.uint64)), _p(np.ascontiguousarray(col_idx, dtype=np.uint64)),
                           _p(np.ascontiguousarray(vals, dtype=np.uint64)), _p(np.ascontiguousarray(r, dtype=np.uint64)),
                           rows, cols, _p(out))
    return out


def linear_poly(u_pre: np.ndarray, r_a: np.ndarray, count: int, k: int, threads: int = 0) -> np.ndarray:
    lib = load()
    out = np.empty((2 * k, 4), dtype=np.uint64)
    lib.ref_linear_poly(_p(np.ascontiguousarray(u_pre)), _p(np.ascontiguousarray(r_a)), count, k, threads or os.cpu_count() or 1, _p(out))
    return out


def quadratic_poly(u_pre: np.ndarray, r: np.ndarray, m: int, k: int) -> np.ndarray:
    lib = load()
    out = np.empty((2 * k, 4), dtype=np.uint64)
    lib.ref_quadratic_poly(_p(np.ascontiguousarray(u_pre)), _p(np.ascontiguousarray(r)), m, k, _p(out))
    return out


def sha256(data: bytes) -> bytes:
    lib = load()
    out = np.zeros(32, dtype=np.uint8)
    buf = np.frombuffer(data, dtype=np.uint8).copy() if data else np.zeros(1, dtype=np.uint8)
    lib.ref_sha256(_p(buf), len(data), _p(out))
    return bytes(out)


def blake2s(data: bytes) -> bytes:
    lib = load()
    out = np.zeros(32, dtype=np.uint8)
    buf = np.frombuffer(data, dtype=np.uint8).copy() if data else np.zeros(1, dtype=np.uint8)
    lib.ref_blake2s(_p(buf), len(data), _p(out))
    return bytes(out)
