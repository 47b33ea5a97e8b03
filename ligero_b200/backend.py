"""Thin Python mirror of the C ABI: device context, committed matrix, Fr <-> limb conversion.

Everything numeric happens in libligero_b200.so on the GPU; this module only marshals buffers.
"""
from __future__ import annotations

import ctypes
import weakref
from ctypes import byref, c_double, c_size_t, c_void_p
from typing import Iterable, List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import LigeroB200Error, check

BN254_R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
_MONT_R = pow(2, 256, BN254_R)
_MONT_RINV = pow(_MONT_R, -1, BN254_R)
_MASK64 = (1 << 64) - 1


def fr_to_limbs(values: Iterable[int]) -> np.ndarray:
    """canonical integers -> uint64[len,4] Montgomery limbs (the layout of ark_bn254::Fr)."""
    vals = list(values)
    out = np.empty((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        m = (v % BN254_R) * _MONT_R % BN254_R
        out[i, 0] = m & _MASK64
        out[i, 1] = (m >> 64) & _MASK64
        out[i, 2] = (m >> 128) & _MASK64
        out[i, 3] = m >> 192
    return out


def limbs_to_fr(arr: np.ndarray) -> List[int]:
    """uint64[...,4] Montgomery limbs -> canonical integers."""
    a = np.ascontiguousarray(arr, dtype=np.uint64).reshape(-1, 4)
    out = []
    for row in a.tolist():
        m = row[0] | (row[1] << 64) | (row[2] << 128) | (row[3] << 192)
        out.append(m * _MONT_RINV % BN254_R)
    return out


def _ptr(x) -> c_void_p:
    """host numpy array, torch tensor (host or CUDA) or raw int address -> void*"""
    if x is None:
        return c_void_p(0)
    if isinstance(x, np.ndarray):
        assert x.flags["C_CONTIGUOUS"]
        return c_void_p(x.ctypes.data)
    if hasattr(x, "data_ptr"):
        assert x.is_contiguous()
        if getattr(x, "is_cuda", False):
            # the library launches on its own non-blocking stream, which does not wait for torch's: whatever torch still
            # has in flight for this tensor must be finished before the pointer is handed over
            import torch
            torch.cuda.current_stream(x.device).synchronize()
        return c_void_p(x.data_ptr())
    return c_void_p(int(x))


class Context:
    """One per GPU / host thread (lg_ctx)."""

    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        h = c_void_p()
        st = self.lib.lg_ctx_create(device, byref(h))
        if st != 0:
            raise LigeroB200Error(
                f"lg_ctx_create(device={device}) failed with status {st}: a CUDA device is required "
                "(ligero_b200 has no CPU fallback)")
        self.handle = h
        self.device = device
        # objects that hold device memory of this context (committed matrices, constraint matrices, LigeroCircuits, shards):
        # close() frees them BEFORE the context, and their own finalizers do nothing once the context is gone -- the native
        # free functions dereference the context (include/ligero_b200.h: a context outlives everything created from it)
        self._children = weakref.WeakSet()

    def _adopt(self, child):
        self._children.add(child)

    def close(self):
        if getattr(self, "handle", None):
            for child in list(self._children):
                try:
                    child.free()
                except Exception:
                    pass
            self.lib.lg_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        check(self.lib.lg_ctx_sync(self.handle), self.handle, "lg_ctx_sync")

    @property
    def launches(self) -> int:
        return int(self.lib.lg_ctx_launches(self.handle))

    @property
    def stream(self) -> int:
        return int(self.lib.lg_ctx_stream(self.handle) or 0)

    PHASES = ("ntt_strided_inv", "ntt_local", "ntt_strided_fwd", "hash", "merkle", "expand", "tests", "open")

    def set_timing(self, enabled: bool = True):
        check(self.lib.lg_ctx_set_timing(self.handle, int(enabled)), self.handle)

    def set_overlap(self, enabled: bool = True):
        """Overlap column hashing with encoding inside commit/recommit (default on)."""
        check(self.lib.lg_ctx_set_overlap(self.handle, int(enabled)), self.handle)

    def set_hash_quad_max(self, max_columns: int):
        """Whole-matrix column hashes of at most `max_columns` columns use the four-lanes-per-column kernel (0: never)."""
        check(self.lib.lg_ctx_set_hash_quad_max(self.handle, int(max_columns)), self.handle)

    def phase_ms(self):
        """{phase: (accumulated ms, intervals)} since the last call (synchronises the stream)."""
        n = len(self.PHASES)
        ms = (c_double * n)()
        cnt = (ctypes.c_uint64 * n)()
        check(self.lib.lg_ctx_phase_ms(self.handle, ms, cnt, n), self.handle)
        return {p: (ms[i], int(cnt[i])) for i, p in enumerate(self.PHASES)}

    def set_formats(self, col_len_prefix: bool = True, leaf_len_prefix: bool = True):
        check(self.lib.lg_ctx_set_formats(self.handle, int(col_len_prefix), int(leaf_len_prefix)), self.handle)

    # ---- encode + commit -------------------------------------------------------------------
    def commit(self, preenc_u, rows: int, k: int, rho_inv: int = 8) -> "CommittedMatrix":
        """lg_commit: RS-encode the rows, hash the columns, build the Merkle tree."""
        root = np.zeros(32, dtype=np.uint8)
        h = c_void_p()
        check(self.lib.lg_commit(self.handle, _ptr(preenc_u), rows, k, rho_inv, _ptr(root), byref(h)),
              self.handle, "lg_commit")
        return CommittedMatrix(self, h, bytes(root))

    def encode(self, preenc_u, rows: int, k: int, rho_inv: int = 8) -> "CommittedMatrix":
        h = c_void_p()
        check(self.lib.lg_encode(self.handle, _ptr(preenc_u), rows, k, rho_inv, byref(h)), self.handle, "lg_encode")
        return CommittedMatrix(self, h, None)

    def wrap(self, u_dev, rows: int, k: int, rho_inv: int = 8) -> "CommittedMatrix":
        """lg_matrix_wrap: handle over a caller-owned device buffer (torch tensor) in the plane layout."""
        h = c_void_p()
        check(self.lib.lg_matrix_wrap(self.handle, _ptr(u_dev), rows, k, rho_inv, byref(h)), self.handle, "lg_matrix_wrap")
        cm = CommittedMatrix(self, h, None)
        cm._keepalive = u_dev
        return cm

    def intt(self, evals, rows: int, size: int, out=None):
        if out is None:
            out = np.empty((rows * size, 4), dtype=np.uint64)
        check(self.lib.lg_intt(self.handle, _ptr(evals), _ptr(out), rows, size), self.handle, "lg_intt")
        return out

    # ---- challenges ------------------------------------------------------------------------
    def expand_fr(self, seed: bytes, count: int, out=None):
        """get_field_elements_from_prng on the device; returns uint64[count,4] (host) unless `out` given."""
        assert len(seed) == 32
        if out is None:
            out = np.empty((count, 4), dtype=np.uint64)
        s = np.frombuffer(seed, dtype=np.uint8).copy()
        check(self.lib.lg_expand_fr(self.handle, _ptr(s), count, _ptr(out)), self.handle, "lg_expand_fr")
        return out

    def expand_indices(self, seed: bytes, n: int, t: int) -> np.ndarray:
        assert len(seed) == 32
        out = np.empty(t, dtype=np.uint64)
        s = np.frombuffer(seed, dtype=np.uint8).copy()
        check(self.lib.lg_expand_indices(_ptr(s), n, t, _ptr(out)), self.handle, "lg_expand_indices")
        return out

    def constraints(self, mk: int, col_ptr, row_idx, val_id, const_table=None) -> "Constraints":
        return Constraints(self, mk, col_ptr, row_idx, val_id, const_table)

    def int_peak(self, ms_target: float = 50.0):
        a, b = c_double(), c_double()
        check(self.lib.lg_bench_int_peak(self.handle, ms_target, byref(a), byref(b)), self.handle, "lg_bench_int_peak")
        c, d = c_double(), c_double()
        check(self.lib.lg_bench_shoup_peak(self.handle, byref(c), byref(d)), self.handle, "lg_bench_shoup_peak")
        return {"fr_mul_per_s": a.value, "imad_wide_per_s": b.value, "shoup_mul_per_s": c.value,
                "butterfly_per_s": d.value}


class Constraints:
    """lg_constraints: CSC of the right-hand block of A on the device."""

    def __init__(self, ctx: Context, mk: int, col_ptr, row_idx, val_id, const_table=None):
        self.ctx, self.mk = ctx, mk
        cp = np.ascontiguousarray(col_ptr, dtype=np.uint32)
        ri = np.ascontiguousarray(row_idx, dtype=np.uint32)
        vi = np.ascontiguousarray(val_id, dtype=np.uint32)
        ct = np.ascontiguousarray(const_table, dtype=np.uint64).reshape(-1, 4) if const_table is not None else np.zeros((0, 4), np.uint64)
        h = c_void_p()
        check(ctx.lib.lg_constraints_create(ctx.handle, mk, _ptr(cp), _ptr(ri) if len(ri) else None, _ptr(vi) if len(vi) else None,
                                            len(ri), _ptr(ct) if len(ct) else None, len(ct), byref(h)),
              ctx.handle, "lg_constraints_create")
        self.handle = h
        ctx._adopt(self)

    def free(self):
        if self.handle:
            if self.ctx.handle:                      # the context is still alive (see Context.close)
                self.ctx.lib.lg_constraints_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def row_mul(self, r_linear) -> np.ndarray:
        out = np.empty((4 * self.mk, 4), dtype=np.uint64)
        check(self.ctx.lib.lg_sparse_row_mul(self.ctx.handle, self.handle, _ptr(r_linear), _ptr(out)), self.ctx.handle,
              "lg_sparse_row_mul")
        return out


class CommittedMatrix:
    """lg_matrix: U, leaf digests and Merkle nodes resident in HBM."""

    def __init__(self, ctx: Context, handle: c_void_p, root: Optional[bytes]):
        self.ctx, self.handle, self.root = ctx, handle, root
        r, k, n = c_size_t(), c_size_t(), c_size_t()
        check(ctx.lib.lg_matrix_dims(handle, byref(r), byref(k), byref(n)), ctx.handle)
        self.rows, self.k, self.n = r.value, k.value, n.value
        ctx._adopt(self)

    def free(self):
        if self.handle:
            if self.ctx.handle:                      # the context is still alive (see Context.close)
                self.ctx.lib.lg_matrix_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def recommit(self, preenc_u) -> bytes:
        root = np.zeros(32, dtype=np.uint8)
        check(self.ctx.lib.lg_recommit(self.handle, _ptr(preenc_u), _ptr(root)), self.ctx.handle, "lg_recommit")
        self.root = bytes(root)
        return self.root

    def encode(self, preenc_u):
        check(self.ctx.lib.lg_matrix_encode(self.handle, _ptr(preenc_u)), self.ctx.handle, "lg_matrix_encode")

    def hash_async(self):
        """column hashing + tree, no synchronisation (root stays on the device)."""
        check(self.ctx.lib.lg_matrix_hash(self.handle, None), self.ctx.handle, "lg_matrix_hash")

    def hash_rows(self, row0: int, row_end: int):
        """lg_matrix_hash_rows: hash rows [row0, row_end) of every column on the second stream (tiles in row order)."""
        check(self.ctx.lib.lg_matrix_hash_rows(self.handle, row0, row_end), self.ctx.handle, "lg_matrix_hash_rows")

    def hash_finish(self) -> bytes:
        """lg_matrix_hash_finish: tree over the leaves of the last tile; returns the root."""
        root = np.zeros(32, dtype=np.uint8)
        check(self.ctx.lib.lg_matrix_hash_finish(self.handle, _ptr(root)), self.ctx.handle, "lg_matrix_hash_finish")
        self.root = bytes(root)
        return self.root

    def hash(self) -> bytes:
        root = np.zeros(32, dtype=np.uint8)
        check(self.ctx.lib.lg_matrix_hash(self.handle, _ptr(root)), self.ctx.handle, "lg_matrix_hash")
        self.root = bytes(root)
        return self.root

    # ---- tests and openings ----------------------------------------------------------------
    def row_combine(self, r) -> np.ndarray:
        out = np.empty((self.k, 4), dtype=np.uint64)
        check(self.ctx.lib.lg_row_combine(self.handle, _ptr(r), _ptr(out)), self.ctx.handle, "lg_row_combine")
        return out

    def linear_test(self, constraints: "Constraints", r_linear=None, seed: Optional[bytes] = None) -> np.ndarray:
        out = np.empty((2 * self.k, 4), dtype=np.uint64)
        ln = c_size_t()
        if seed is not None:
            s = np.frombuffer(seed, dtype=np.uint8).copy()
            st = self.ctx.lib.lg_linear_test_seeded(self.handle, constraints.handle, _ptr(s), _ptr(out), byref(ln))
        else:
            st = self.ctx.lib.lg_linear_test(self.handle, constraints.handle, _ptr(r_linear), _ptr(out), byref(ln))
        check(st, self.ctx.handle, "lg_linear_test")
        return out[: ln.value]

    def quadratic_test(self, r_quad) -> np.ndarray:
        out = np.empty((2 * self.k, 4), dtype=np.uint64)
        ln = c_size_t()
        check(self.ctx.lib.lg_quadratic_test(self.handle, _ptr(r_quad), _ptr(out), byref(ln)), self.ctx.handle, "lg_quadratic_test")
        return out[: ln.value]

    def open(self, idx):
        """(columns uint64[t, rows, 4], leaf_sibling uint8[t,32], auth uint8[t, log2(n)-1, 32])"""
        ii = np.ascontiguousarray(idx, dtype=np.uint64)
        t = len(ii)
        depth = self.n.bit_length() - 2
        cols = np.empty((t, self.rows, 4), dtype=np.uint64)
        sib = np.empty((t, 32), dtype=np.uint8)
        auth = np.empty((t, max(depth, 0), 32), dtype=np.uint8)
        check(self.ctx.lib.lg_open(self.handle, _ptr(ii), t, _ptr(cols), _ptr(sib), _ptr(auth) if depth > 0 else None),
              self.ctx.handle, "lg_open")
        return cols, sib, auth

    def read_rows(self, row0: int, nrows: int) -> np.ndarray:
        out = np.empty((nrows, self.n, 4), dtype=np.uint64)
        check(self.ctx.lib.lg_matrix_read_rows(self.handle, row0, nrows, _ptr(out)), self.ctx.handle, "read_rows")
        return out

    def read_leaves(self) -> np.ndarray:
        out = np.empty((self.n, 32), dtype=np.uint8)
        check(self.ctx.lib.lg_matrix_read_leaves(self.handle, _ptr(out)), self.ctx.handle, "read_leaves")
        return out

    def read_nodes(self) -> np.ndarray:
        out = np.empty((self.n - 1, 32), dtype=np.uint8)
        check(self.ctx.lib.lg_matrix_read_nodes(self.handle, _ptr(out)), self.ctx.handle, "read_nodes")
        return out
