"""Proof wire format, oracle side (TEST INFRASTRUCTURE ONLY).

The reference defines no serialisation (``LigeroProof`` has no derives, src/ligero/mod.rs:96-144); this
is the layout arkworks' ``CanonicalSerialize`` would produce for the same structs: little
endian, ``Vec<T>`` = u64 length + items, ``Fr`` = 32 canonical bytes, digests are ``Vec<u8>``,
``Path`` = {leaf_sibling_hash, auth_path, leaf_index: u64}.  The product implements the same layout in
C++ (ligero_b200/csrc/capi_host.cu); tests compare the two byte strings.
"""
from __future__ import annotations

import struct
from typing import List

from . import ligero_oracle as O


def _u64(v: int) -> bytes:
    return struct.pack("<Q", v)


def _frs(v: List[int]) -> bytes:
    return _u64(len(v)) + b"".join(x.to_bytes(32, "little") for x in v)


def _digest(d: bytes) -> bytes:
    return _u64(len(d)) + d


def _opened(o: "O.OpenedColumns") -> bytes:
    out = _u64(len(o.columns)) + b"".join(_frs(c) for c in o.columns)
    out += _u64(len(o.paths))
    for p in o.paths:
        out += _digest(p.leaf_sibling_hash) + _u64(len(p.auth_path)) + b"".join(_digest(d) for d in p.auth_path) + _u64(p.leaf_index)
    return out


def serialize_proof(p: "O.LigeroProof") -> bytes:
    return (_digest(p.u_root) + _frs(p.preenc_u_lc) + _opened(p.interleaved) + _frs(p.linear_poly) + _opened(p.linear)
            + _frs(p.quadratic_poly) + _opened(p.quadratic))


class _Reader:
    def __init__(self, data: bytes):
        self.d, self.pos = data, 0

    def u64(self) -> int:
        v = struct.unpack_from("<Q", self.d, self.pos)[0]
        self.pos += 8
        return v

    def frs(self) -> List[int]:
        n = self.u64()
        out = [int.from_bytes(self.d[self.pos + 32 * i: self.pos + 32 * i + 32], "little") for i in range(n)]
        self.pos += 32 * n
        return out

    def digest(self) -> bytes:
        n = self.u64()
        out = self.d[self.pos: self.pos + n]
        self.pos += n
        return out

    def opened(self) -> "O.OpenedColumns":
        cols = [self.frs() for _ in range(self.u64())]
        paths = []
        for _ in range(self.u64()):
            sib = self.digest()
            auth = [self.digest() for _ in range(self.u64())]
            paths.append(O.MerklePath(sib, auth, self.u64()))
        return O.OpenedColumns(cols, paths)


def deserialize_proof(data: bytes) -> "O.LigeroProof":
    r = _Reader(data)
    root = r.digest()
    lc = r.frs()
    inter = r.opened()
    lin = r.frs()
    lin_o = r.opened()
    quad = r.frs()
    quad_o = r.opened()
    assert r.pos == len(data)
    return O.LigeroProof(root, lc, inter, lin, lin_o, quad, quad_o)
