"""Python mirror of the reference's public API (same names and argument meaning as NP-Eng/ligero):

    ArithmeticCircuit  -- src/arithmetic_circuit/mod.rs
    LigeroCircuit      -- src/ligero/mod.rs   (new / prove / prove_with_labels / verify)
    PoseidonSponge     -- ark-crypto-primitives sponge (host side of Fiat-Shamir), test_sponge()
    LigeroProof        -- src/ligero/mod.rs:96-144 plus the wire format the reference lacks

Everything is a thin ctypes veneer over libligero_b200.so (C++ host driver + sm_100a kernels); field
elements cross the boundary as canonical Python ints and are converted to Montgomery limbs here.
"""
from __future__ import annotations

import ctypes
import struct
from ctypes import byref, c_char_p, c_int, c_size_t, c_void_p
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import LigeroB200Error, check
from .backend import BN254_R, Context, _ptr, fr_to_limbs, limbs_to_fr

DEFAULT_SECURITY_LEVEL = 128      # src/lib.rs:8
CHACHA_SEED_BYTES = 32            # src/lib.rs:9


def _one_fr(v: int) -> np.ndarray:
    return fr_to_limbs([v])


class ArithmeticCircuit:
    def __init__(self):
        self.lib = _lib.load()
        h = c_void_p()
        check(self.lib.lg_circuit_new(byref(h)), None, "lg_circuit_new")
        self.handle = h

    def __del__(self):
        try:
            if self.handle:
                self.lib.lg_circuit_free(self.handle)
                self.handle = None
        except Exception:
            pass

    def _check(self, st, what):
        if st != 0:
            msg = self.lib.lg_circuit_last_error(self.handle)
            raise LigeroB200Error(f"{what}: {msg.decode() if msg else st}")

    def _idx_call(self, fn, *args, what=""):
        out = c_size_t()
        self._check(fn(self.handle, *args, byref(out)), what)
        return out.value

    def constant(self, value: int) -> int:
        return self._idx_call(self.lib.lg_circuit_constant, _ptr(_one_fr(value)), what="constant")

    def new_variable_with_label(self, label: str) -> int:
        return self._idx_call(self.lib.lg_circuit_new_variable, label.encode(), what="new_variable_with_label")

    def new_variable(self) -> int:
        return self._idx_call(self.lib.lg_circuit_new_variable, None, what="new_variable")

    def new_variables(self, num: int) -> List[int]:
        return [self.new_variable() for _ in range(num)]

    def get_variable(self, label: str) -> int:
        out = c_size_t()
        if self.lib.lg_circuit_get_variable(self.handle, label.encode(), byref(out)) != 0:
            raise LigeroB200Error("Variable not in circuit")
        return out.value

    def add(self, left: int, right: int) -> int:
        return self._idx_call(self.lib.lg_circuit_add, left, right, what="add")

    def mul(self, left: int, right: int) -> int:
        return self._idx_call(self.lib.lg_circuit_mul, left, right, what="mul")

    def add_nodes(self, indices: Sequence[int]) -> int:
        it = list(indices)
        if not it:
            raise LigeroB200Error("add_nodes of an empty list")
        acc = it[0]
        for i in it[1:]:
            acc = self.add(acc, i)
        return acc

    def mul_nodes(self, indices: Sequence[int]) -> int:
        it = list(indices)
        acc = it[0]
        for i in it[1:]:
            acc = self.mul(acc, i)
        return acc

    def pow(self, node: int, exponent: int) -> int:
        cur = node
        for b in bin(exponent)[3:]:
            cur = self.mul(cur, cur)
            if b == "1":
                cur = self.mul(cur, node)
        return cur

    def indicator(self, node: int) -> int:
        return self.pow(node, BN254_R - 1)

    def minus(self, node: int) -> int:
        return self.mul(self.constant(BN254_R - 1), node)

    def scalar_product(self, left: Sequence[int], right: Sequence[int]) -> int:
        return self.add_nodes([self.mul(l, r) for l, r in zip(left, right)])

    @staticmethod
    def from_r1cs_bytes(data: bytes):
        """read_constraint_system's R1CS half (src/reader.rs): iden3 .r1cs v1 image -> (circuit, outputs, n_wires)."""
        lib = _lib.load()
        buf = np.frombuffer(data, dtype=np.uint8)
        nc, nw, h = c_size_t(), c_size_t(), c_void_p()
        lib.lg_circuit_from_r1cs_bytes(_ptr(buf), len(buf), byref(h), None, 0, byref(nc), byref(nw))   # sizes only
        if nc.value == 0:
            raise LigeroB200Error("not an iden3 .r1cs v1 file over BN254 Fr")
        outputs = np.zeros(nc.value, dtype=np.uint64)
        st = lib.lg_circuit_from_r1cs_bytes(_ptr(buf), len(buf), byref(h), _ptr(outputs), nc.value, byref(nc), byref(nw))
        if st != 0:
            raise LigeroB200Error(f"from_r1cs_bytes failed with status {st} (malformed file, or an empty row of A, B or C)")
        c = ArithmeticCircuit.__new__(ArithmeticCircuit)
        c.lib, c.handle = lib, h
        return c, [int(x) for x in outputs], nw.value

    @staticmethod
    def from_r1cs_file(path: str):
        with open(path, "rb") as f:
            return ArithmeticCircuit.from_r1cs_bytes(f.read())

    @classmethod
    def synthetic(cls, gates: int, seed: int = 1):
        """lg_circuit_synthetic: seeded random Add/Mul circuit of exactly `gates` gates (SURVEY 8d rules).
        Returns (circuit, output node, [(variable index, value)])."""
        self = cls.__new__(cls)
        self.lib = _lib.load()
        h, out = c_void_p(), c_size_t()
        vidx = (c_size_t * 2)()
        vals = np.zeros((2, 4), dtype=np.uint64)
        check(self.lib.lg_circuit_synthetic(gates, seed, byref(h), byref(out), vidx, _ptr(vals)), None, "lg_circuit_synthetic")
        self.handle = h
        v = limbs_to_fr(vals)
        return self, out.value, [(int(vidx[0]), v[0]), (int(vidx[1]), v[1])]

    def _counts(self):
        a, b, c, d = c_size_t(), c_size_t(), c_size_t(), c_size_t()
        self.lib.lg_circuit_counts(self.handle, byref(a), byref(b), byref(c), byref(d))
        return a.value, b.value, c.value, d.value

    def num_nodes(self): return self._counts()[0]
    def num_constants(self): return self._counts()[1]
    def num_variables(self): return self._counts()[2]
    def num_gates(self): return self._counts()[3]
    def last(self): return self.num_nodes() - 1

    def node(self, index: int):
        t, l, r = c_int(), c_size_t(), c_size_t()
        v = np.zeros((1, 4), dtype=np.uint64)
        self._check(self.lib.lg_circuit_node(self.handle, index, byref(t), byref(l), byref(r), _ptr(v)), "node")
        if t.value == 0:
            return ("var",)
        if t.value == 1:
            return ("const", limbs_to_fr(v)[0])
        return ("add" if t.value == 2 else "mul", l.value, r.value)

    def evaluate_multioutput(self, vars: Sequence[Tuple[int, int]], outputs: Sequence[int]) -> List[int]:
        idx = np.array([i for i, _ in vars], dtype=np.uint64)
        vals = fr_to_limbs([v for _, v in vars])
        outs = np.array(list(outputs), dtype=np.uint64)
        res = np.zeros((max(len(outs), 1), 4), dtype=np.uint64)
        count = c_size_t()
        self._check(self.lib.lg_circuit_evaluate(self.handle, _ptr(idx), _ptr(vals), len(idx), _ptr(outs), len(outs), _ptr(res),
                                                 byref(count)), "evaluate")
        return limbs_to_fr(res[: count.value])      # node order, one value per distinct output (mod.rs:384-389)

    def evaluate_node(self, vars, node: int) -> int:
        return self.evaluate_multioutput(vars, [node])[0]

    def evaluate(self, vars) -> int:
        return self.evaluate_node(vars, self.last())

    @staticmethod
    def from_constraint_system(a_rows, b_rows, c_rows, num_vars_incl_one: int):
        """rows: lists of [(coeff:int, column:int)] as ConstraintSystem::to_matrices yields them."""
        lib = _lib.load()
        n = len(a_rows)
        keep = []
        rp, ci, cf = (c_void_p * 3)(), (c_void_p * 3)(), (c_void_p * 3)()
        for i, mat in enumerate((a_rows, b_rows, c_rows)):
            ptr = np.zeros(n + 1, dtype=np.uint64)
            cols, vals = [], []
            for r, row in enumerate(mat):
                for coeff, col in row:
                    cols.append(col)
                    vals.append(coeff)
                ptr[r + 1] = len(cols)
            cols = np.array(cols, dtype=np.uint64)
            vals = fr_to_limbs(vals)
            keep += [ptr, cols, vals]
            rp[i], ci[i], cf[i] = ptr.ctypes.data, cols.ctypes.data, vals.ctypes.data
        out = c_void_p()
        outputs = np.zeros(n, dtype=np.uint64)
        st = lib.lg_circuit_from_r1cs(n, num_vars_incl_one, rp, ci, cf, byref(out), _ptr(outputs))
        if st != 0:
            raise LigeroB200Error("from_constraint_system failed (empty R1CS row or column out of range)")
        c = ArithmeticCircuit.__new__(ArithmeticCircuit)
        c.lib, c.handle = lib, out
        return c, [int(x) for x in outputs]


class PoseidonSponge:
    def __init__(self, handle):
        self.lib = _lib.load()
        self.handle = handle

    @staticmethod
    def test_sponge() -> "PoseidonSponge":
        lib = _lib.load()
        h = c_void_p()
        check(lib.lg_sponge_test(byref(h)), None, "lg_sponge_test")
        return PoseidonSponge(h)

    @staticmethod
    def new(full_rounds: int, partial_rounds: int, alpha: int, mds: Sequence[Sequence[int]], ark: Sequence[Sequence[int]],
            rate: int, capacity: int) -> "PoseidonSponge":
        lib = _lib.load()
        h = c_void_p()
        m = fr_to_limbs([x for row in mds for x in row])
        a = fr_to_limbs([x for row in ark for x in row])
        check(lib.lg_sponge_new(full_rounds, partial_rounds, alpha, _ptr(m), _ptr(a), rate, capacity, byref(h)), None, "lg_sponge_new")
        return PoseidonSponge(h)

    def clone(self) -> "PoseidonSponge":
        h = c_void_p()
        check(self.lib.lg_sponge_clone(self.handle, byref(h)), None, "lg_sponge_clone")
        return PoseidonSponge(h)

    def absorb_bytes(self, data: bytes):
        buf = np.frombuffer(data, dtype=np.uint8).copy() if data else np.zeros(1, np.uint8)
        check(self.lib.lg_sponge_absorb_bytes(self.handle, _ptr(buf), len(data)), None, "absorb")

    def absorb_field_elements(self, elems: Sequence[int]):
        a = fr_to_limbs(elems)
        check(self.lib.lg_sponge_absorb_fr(self.handle, _ptr(a) if len(a) else None, len(a)), None, "absorb")

    def absorb_fr(self, limbs):
        """absorb(&Vec<F>) from Montgomery-form limbs (uint64[count, 4]) as they come back from the device"""
        a = np.ascontiguousarray(limbs, dtype=np.uint64).reshape(-1, 4)
        check(self.lib.lg_sponge_absorb_fr(self.handle, _ptr(a) if len(a) else None, len(a)), None, "absorb")

    def squeeze_bytes(self, n: int) -> bytes:
        out = np.zeros(n, dtype=np.uint8)
        check(self.lib.lg_sponge_squeeze_bytes(self.handle, _ptr(out), n), None, "squeeze")
        return bytes(out)

    def __del__(self):
        try:
            if self.handle:
                self.lib.lg_sponge_free(self.handle)
                self.handle = None
        except Exception:
            pass


class LigeroProof:
    def __init__(self, handle):
        self.lib = _lib.load()
        self.handle = handle

    def to_bytes(self) -> bytes:
        n = c_size_t()
        check(self.lib.lg_proof_serialize(self.handle, None, 0, byref(n)), None, "lg_proof_serialize")
        buf = np.zeros(n.value, dtype=np.uint8)
        check(self.lib.lg_proof_serialize(self.handle, _ptr(buf), n.value, byref(n)), None, "lg_proof_serialize")
        return bytes(buf)

    @staticmethod
    def from_bytes(data: bytes) -> "LigeroProof":
        lib = _lib.load()
        h = c_void_p()
        buf = np.frombuffer(data, dtype=np.uint8).copy()
        st = lib.lg_proof_deserialize(_ptr(buf), len(data), byref(h))
        if st != 0:
            raise LigeroB200Error("malformed proof bytes")
        return LigeroProof(h)

    def __del__(self):
        try:
            if self.handle:
                self.lib.lg_proof_free(self.handle)
                self.handle = None
        except Exception:
            pass


class LigeroCircuit:
    """LigeroCircuit::new(circuit, outputs, lambda) on a device context."""

    def __init__(self, ctx: Context, circuit: ArithmeticCircuit, outputs: Sequence[int], lam: int = DEFAULT_SECURITY_LEVEL):
        self.ctx = ctx
        self.lib = ctx.lib
        outs = np.array(list(outputs), dtype=np.uint64)
        h = c_void_p()
        check(self.lib.lg_ligero_new(ctx.handle, circuit.handle, _ptr(outs), len(outs), lam, byref(h)), ctx.handle, "LigeroCircuit::new")
        self.handle = h
        m, k, n, t, s = c_size_t(), c_size_t(), c_size_t(), c_size_t(), c_size_t()
        self.lib.lg_ligero_params(h, byref(m), byref(k), byref(n), byref(t), byref(s))
        self.m, self.k, self.n, self.t, self.sol_len = m.value, k.value, n.value, t.value, s.value
        ctx._adopt(self)

    def free(self):
        """lg_ligero_free; a no-op once the context is closed (Context.close frees its circuits first)"""
        if getattr(self, "handle", None):
            if self.ctx.handle:
                self.lib.lg_ligero_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def witness_matrix(self, var_assignment: Sequence[Tuple[int, int]], bump: bool = True) -> np.ndarray:
        idx = np.array([i for i, _ in var_assignment], dtype=np.uint64)
        vals = fr_to_limbs([v for _, v in var_assignment])
        out = np.zeros((4 * self.m * self.k, 4), dtype=np.uint64)
        check(self.lib.lg_ligero_witness_matrix(self.handle, _ptr(idx), _ptr(vals), len(idx), int(bump), _ptr(out)),
              self.ctx.handle, "witness layout")
        return out

    def witness_matrix_device(self, var_assignment: Sequence[Tuple[int, int]], bump: bool = True):
        """lg_ligero_witness_matrix_dev: evaluation trace + [X;Y;Z;W] layout on the GPU; returns a torch int64 tensor
        (4*m*k, 4) resident in HBM (Montgomery limbs), usable as the input of Context.commit / prove_matrix."""
        import torch
        idx = np.array([i for i, _ in var_assignment], dtype=np.uint64)
        vals = fr_to_limbs([v for _, v in var_assignment])
        out = torch.empty((4 * self.m * self.k, 4), dtype=torch.int64, device=f"cuda:{self.ctx.device}")
        check(self.lib.lg_ligero_witness_matrix_dev(self.handle, _ptr(idx), _ptr(vals), len(idx), int(bump), _ptr(out)),
              self.ctx.handle, "device witness layout")
        return out

    def constraints_csc(self):
        """(col_ptr, row_idx, val_id, const_table) of the right-hand block of A as LigeroCircuit::new built it (read back
        from the device; lg_constraints_read)"""
        a = c_void_p()
        check(self.lib.lg_ligero_constraints(self.handle, byref(a)), self.ctx.handle, "lg_ligero_constraints")
        mk, nnz, nc = c_size_t(), c_size_t(), c_size_t()
        check(self.lib.lg_constraints_read(a, byref(mk), byref(nnz), byref(nc), None, None, None, None), self.ctx.handle)
        col_ptr = np.zeros(mk.value + 1, dtype=np.uint32)
        row_idx = np.zeros(max(nnz.value, 1), dtype=np.uint32)
        val_id = np.zeros(max(nnz.value, 1), dtype=np.uint32)
        table = np.zeros((max(nc.value, 1), 4), dtype=np.uint64)
        check(self.lib.lg_constraints_read(a, None, None, None, _ptr(col_ptr), _ptr(row_idx), _ptr(val_id), _ptr(table)),
              self.ctx.handle, "lg_constraints_read")
        return col_ptr, row_idx[: nnz.value], val_id[: nnz.value], table[: nc.value]

    def set_trace_mode(self, mode: int):
        """-1: device trace for wide circuits (default), 0: host evaluator, 1: device."""
        check(self.lib.lg_ligero_set_trace_mode(self.handle, mode), self.ctx.handle, "set_trace_mode")

    def release_buffers(self):
        """return the device buffers kept between proofs of this circuit (lg_ligero_release_buffers)"""
        check(self.lib.lg_ligero_release_buffers(self.handle), self.ctx.handle)

    PROVE_PHASES = ("trace", "commit", "interleaved", "linear", "quadratic", "openings", "total")

    def prove_ms(self) -> Dict[str, float]:
        """host wall clock (ms) of the phases of the last prove on this circuit"""
        ms = (ctypes.c_double * 7)()
        check(self.lib.lg_ligero_prove_ms(self.handle, ms), self.ctx.handle)
        return {p: ms[i] for i, p in enumerate(self.PROVE_PHASES)}

    def trace_info(self) -> Dict[str, int]:
        g, lv, la, dev = c_size_t(), c_size_t(), c_size_t(), c_int()
        check(self.lib.lg_ligero_trace_info(self.handle, byref(g), byref(lv), byref(la), byref(dev)), self.ctx.handle)
        return {"gates": g.value, "levels": lv.value, "launches": la.value, "on_device": bool(dev.value)}

    def prove(self, var_assignment: Sequence[Tuple[int, int]], sponge: PoseidonSponge) -> LigeroProof:
        idx = np.array([i for i, _ in var_assignment], dtype=np.uint64)
        vals = fr_to_limbs([v for _, v in var_assignment])
        h = c_void_p()
        check(self.lib.lg_prove(self.handle, _ptr(idx), _ptr(vals), len(idx), 1, sponge.handle, byref(h)), self.ctx.handle, "prove")
        return LigeroProof(h)

    def prove_with_labels(self, var_assignment: Sequence[Tuple[str, int]], sponge: PoseidonSponge) -> LigeroProof:
        labels = (c_char_p * len(var_assignment))(*[l.encode() for l, _ in var_assignment])
        vals = fr_to_limbs([v for _, v in var_assignment])
        h = c_void_p()
        check(self.lib.lg_prove_with_labels(self.handle, labels, _ptr(vals), len(var_assignment), sponge.handle, byref(h)),
              self.ctx.handle, "prove_with_labels")
        return LigeroProof(h)

    def prove_matrix(self, preenc_u, sponge: PoseidonSponge) -> LigeroProof:
        h = c_void_p()
        check(self.lib.lg_prove_matrix(self.handle, _ptr(preenc_u), sponge.handle, byref(h)), self.ctx.handle, "prove_matrix")
        return LigeroProof(h)

    def verify(self, proof: LigeroProof, sponge: PoseidonSponge) -> bool:
        ok = c_int()
        check(self.lib.lg_verify(self.handle, proof.handle, sponge.handle, byref(ok)), self.ctx.handle, "verify")
        return bool(ok.value)
