"""CPU oracle (TEST INFRASTRUCTURE ONLY) -- a literal Python restatement of the reference
Ligero prover/verifier, NP-Eng/ligero, over Python big integers.

  * This file is the *checker*: only ``tests/``, ``__graft_entry__.smoke()`` and the
    ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The product
    (``ligero_b200``) never does; it fails loudly when the CUDA library is missing.
  * PARITY UNPINNED: the reference is Rust and no Rust toolchain exists in this
    image, and the reference's own tests pin no value on the hot path.  What *is* pinned by the
    reference (A-matrix tables, ``row_mul`` known answers, ``num_nodes()==15``, ``(m,k)=(4,4)``,
    prove->verify accept / perturbed reject) is asserted in ``tests/test_oracle_*.py``.  The
    arithmetic lives in un-vendored crates (ark-ff/ark-poly/ark-crypto-primitives/ark-serialize
    0.5.0-alpha, ark-poly-commit git HungryCatsStudio/poly-commit@release-0.5, rand_chacha 0.3,
    rand 0.8, blake2 0.10, sha2); their published algorithms are restated below and every
    recollection that could silently differ is a named switch in ``Formats``.

Citations ``ref:`` are file:line under /root/reference.
"""
from __future__ import annotations

import hashlib
import math
import struct
from dataclasses import dataclass, field as dc_field
from typing import Dict, List, Optional, Sequence, Tuple

# --------------------------------------------------------------------------------------------
# Fields.  BN254 Fr is the GPU field; BLS12-377 Fq only for the A-matrix test.
# --------------------------------------------------------------------------------------------
BN254_R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
BLS12_377_Q = 258664426012969094010652733694893533536393512754914660539884262666720468348340822774968888139573360124440321458177


@dataclass(frozen=True)
class Field:
    p: int
    generator: int          # multiplicative generator (ark FpConfig::GENERATOR)
    two_adicity: int

    @property
    def bits(self) -> int:
        return self.p.bit_length()

    @property
    def limbs(self) -> int:
        return (self.bits + 63) // 64

    @property
    def mont_r(self) -> int:
        return pow(2, 64 * self.limbs, self.p)

    def two_adic_root(self) -> int:
        return pow(self.generator, (self.p - 1) >> self.two_adicity, self.p)

    def to_mont(self, x: int) -> int:
        return x * self.mont_r % self.p

    def from_mont(self, x: int) -> int:
        return x * pow(self.mont_r, -1, self.p) % self.p


FR = Field(BN254_R, 5, 28)
FQ377 = Field(BLS12_377_Q, 15, 46)
P = BN254_R


# --------------------------------------------------------------------------------------------
# Format switches: every [RECALLED] third-party convention in one place.
# --------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class Formats:
    col_len_prefix: bool = True       # A.4: u64_le(len) before the column elements
    leaf_len_prefix: bool = True      # A.5: ByteDigestConverter -> u64_le(32) || digest at bottom level
    rand_is_montgomery: bool = True   # A.6: F::rand output is used as the Montgomery representation
    chacha_rounds: int = 20           # A.6: ChaCha20Rng


DEFAULT_FORMATS = Formats()


# --------------------------------------------------------------------------------------------
# ChaCha (rand_chacha 0.3: 64-bit block counter, 64-bit stream id = 0, words consumed in order)
# --------------------------------------------------------------------------------------------
def _rotl32(x: int, n: int) -> int:
    return ((x << n) | (x >> (32 - n))) & 0xFFFFFFFF


def chacha_block(key_words: Sequence[int], counter: int, rounds: int = 20, stream: int = 0) -> List[int]:
    st = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574, *key_words,
          counter & 0xFFFFFFFF, (counter >> 32) & 0xFFFFFFFF, stream & 0xFFFFFFFF, (stream >> 32) & 0xFFFFFFFF]
    x = list(st)

    def qr(a, b, c, d):
        x[a] = (x[a] + x[b]) & 0xFFFFFFFF; x[d] = _rotl32(x[d] ^ x[a], 16)
        x[c] = (x[c] + x[d]) & 0xFFFFFFFF; x[b] = _rotl32(x[b] ^ x[c], 12)
        x[a] = (x[a] + x[b]) & 0xFFFFFFFF; x[d] = _rotl32(x[d] ^ x[a], 8)
        x[c] = (x[c] + x[d]) & 0xFFFFFFFF; x[b] = _rotl32(x[b] ^ x[c], 7)

    for _ in range(rounds // 2):
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
    return [(x[i] + st[i]) & 0xFFFFFFFF for i in range(16)]


class ChaChaRng:
    """``ChaCha{8,12,20}Rng::from_seed`` + ``next_u32``/``next_u64`` (rand_core BlockRng: the
    u32 stream is consumed contiguously; next_u64 = lo word then hi word)."""

    def __init__(self, seed: bytes, rounds: int = 20):
        assert len(seed) == 32
        self.key = list(struct.unpack("<8I", seed))
        self.rounds = rounds
        self.counter = 0
        self.buf: List[int] = []
        self.pos = 0

    def next_u32(self) -> int:
        if self.pos == len(self.buf):
            self.buf = chacha_block(self.key, self.counter, self.rounds)
            self.counter += 1
            self.pos = 0
        w = self.buf[self.pos]
        self.pos += 1
        return w

    def next_u64(self) -> int:
        lo = self.next_u32()
        hi = self.next_u32()
        return lo | (hi << 32)


def field_rand(rng: ChaChaRng, f: Field = FR, fmt: Formats = DEFAULT_FORMATS) -> int:
    """ark-ff ``impl Distribution<Fp> for Standard``: N x next_u64 -> limbs, mask the
    top (64N - bits) bits, reject if >= p; the accepted integer IS the Montgomery representation.
    Returns the canonical value."""
    shave = 64 * f.limbs - f.bits
    while True:
        limbs = [rng.next_u64() for _ in range(f.limbs)]
        limbs[-1] &= (0xFFFFFFFFFFFFFFFF >> shave)
        v = sum(l << (64 * i) for i, l in enumerate(limbs))
        if v < f.p:
            return f.from_mont(v) if fmt.rand_is_montgomery else v


def get_field_elements_from_prng(n: int, seed: bytes, f: Field = FR, fmt: Formats = DEFAULT_FORMATS) -> List[int]:
    """ref: src/utils.rs:23-29"""
    rng = ChaChaRng(seed, fmt.chacha_rounds)
    return [field_rand(rng, f, fmt) for _ in range(n)]


def gen_range_usize(rng: ChaChaRng, n: int) -> int:
    """rand 0.8 ``UniformInt<usize>::sample_single`` (widening-multiply with zone rejection)."""
    lz = 64 - n.bit_length()
    zone = ((n << lz) - 1) & 0xFFFFFFFFFFFFFFFF
    while True:
        v = rng.next_u64()
        prod = v * n
        hi, lo = prod >> 64, prod & 0xFFFFFFFFFFFFFFFF
        if lo <= zone:
            return hi


def get_distinct_indices_from_prng(n: int, t: int, seed: bytes, fmt: Formats = DEFAULT_FORMATS) -> List[int]:
    """ref: src/utils.rs:31-55 (BTreeSet => ascending; complement when t > n/2)."""
    rng = ChaChaRng(seed, fmt.chacha_rounds)
    to_select = min(t, n - t)
    sel = set()
    while len(sel) < to_select:
        sel.add(gen_range_usize(rng, n))
    if to_select == t:
        return sorted(sel)
    return [i for i in range(n) if i not in sel]


# --------------------------------------------------------------------------------------------
# Radix-2 evaluation domains and dense polynomials (ark-poly)
# --------------------------------------------------------------------------------------------
def _bitrev_permute(a: List[int]) -> None:
    n = len(a)
    j = 0
    for i in range(1, n):
        bit = n >> 1
        while j & bit:
            j ^= bit
            bit >>= 1
        j |= bit
        if i < j:
            a[i], a[j] = a[j], a[i]


def _ntt_inplace(a: List[int], omega: int, p: int) -> None:
    n = len(a)
    _bitrev_permute(a)
    length = 2
    while length <= n:
        w_len = pow(omega, n // length, p)
        half = length >> 1
        tw = [1] * half
        for i in range(1, half):
            tw[i] = tw[i - 1] * w_len % p
        for start in range(0, n, length):
            for j in range(half):
                u = a[start + j]
                v = a[start + j + half] * tw[j] % p
                a[start + j] = (u + v) % p
                a[start + j + half] = (u - v) % p
        length <<= 1


class Domain:
    """``GeneralEvaluationDomain::new(size)`` -> Radix2 (BN254 Fr has no mixed-radix parameters)."""

    def __init__(self, min_size: int, f: Field = FR):
        size = 1
        while size < min_size:
            size <<= 1
        self.f = f
        self.size = size
        self.log = size.bit_length() - 1
        if self.log > f.two_adicity:
            raise ValueError("field cannot accommodate FFT of this size")
        self.group_gen = pow(f.two_adic_root(), 1 << (f.two_adicity - self.log), f.p)
        self.group_gen_inv = pow(self.group_gen, -1, f.p)
        self.size_inv = pow(size, -1, f.p)

    def element(self, i: int) -> int:
        return pow(self.group_gen, i, self.f.p)

    def fft(self, coeffs: Sequence[int]) -> List[int]:
        assert len(coeffs) <= self.size
        a = list(coeffs) + [0] * (self.size - len(coeffs))
        _ntt_inplace(a, self.group_gen, self.f.p)
        return a

    def ifft(self, evals: Sequence[int]) -> List[int]:
        assert len(evals) <= self.size
        a = list(evals) + [0] * (self.size - len(evals))
        _ntt_inplace(a, self.group_gen_inv, self.f.p)
        return [x * self.size_inv % self.f.p for x in a]


def poly_trim(c: List[int]) -> List[int]:
    """DensePolynomial::from_coefficients_vec strips trailing zeros."""
    n = len(c)
    while n and c[n - 1] == 0:
        n -= 1
    return c[:n]


def poly_mul(a: List[int], b: List[int], f: Field = FR) -> List[int]:
    """Exact product (arkworks does it by FFT over a domain of size >= len(a)+len(b)-1)."""
    if not a or not b:
        return []
    d = Domain(len(a) + len(b) - 1, f)
    ea, eb = d.fft(a), d.fft(b)
    return poly_trim(d.ifft([x * y % f.p for x, y in zip(ea, eb)]))


def poly_add(a: List[int], b: List[int], p: int = P) -> List[int]:
    n = max(len(a), len(b))
    return poly_trim([((a[i] if i < len(a) else 0) + (b[i] if i < len(b) else 0)) % p for i in range(n)])


def poly_sub(a: List[int], b: List[int], p: int = P) -> List[int]:
    n = max(len(a), len(b))
    return poly_trim([((a[i] if i < len(a) else 0) - (b[i] if i < len(b) else 0)) % p for i in range(n)])


def poly_scale(a: List[int], s: int, p: int = P) -> List[int]:
    if not a or s % p == 0:
        return []
    return poly_trim([x * s % p for x in a])


def poly_degree(a: List[int]) -> int:
    return 0 if not a else len(a) - 1


def poly_eval(a: List[int], x: int, p: int = P) -> int:
    acc = 0
    for c in reversed(a):
        acc = (acc * x + c) % p
    return acc


# --------------------------------------------------------------------------------------------
# Serialisation, column hash, Merkle tree
# --------------------------------------------------------------------------------------------
def fr_to_bytes(x: int, f: Field = FR) -> bytes:
    return x.to_bytes(8 * f.limbs, "little")


def column_hash(col: Sequence[int], f: Field = FR, fmt: Formats = DEFAULT_FORMATS) -> bytes:
    """FieldToBytesColHasher<F, Blake2s256>::evaluate; ref call site: src/ligero/mod.rs:536-542."""
    h = hashlib.blake2s(digest_size=32)
    if fmt.col_len_prefix:
        h.update(struct.pack("<Q", len(col)))
    for x in col:
        h.update(fr_to_bytes(x, f))
    return h.digest()


def _leaf_ser(d: bytes, fmt: Formats) -> bytes:
    return (struct.pack("<Q", len(d)) + d) if fmt.leaf_len_prefix else d


class MerkleTree:
    """ark-crypto-primitives MerkleTree over TestMerkleTreeParams (identity leaf hash, SHA-256
    two-to-one).  Heap layout: node 0 = root.  ref call site: src/ligero/mod.rs:544-551."""

    def __init__(self, leaves: Sequence[bytes], fmt: Formats = DEFAULT_FORMATS):
        n = len(leaves)
        assert n > 1 and n & (n - 1) == 0, "need a power-of-two number (>1) of leaves"
        self.fmt = fmt
        self.leaves = list(leaves)
        self.n = n
        nodes: List[bytes] = [b""] * (n - 1)
        base = n // 2 - 1                       # first index of the bottom inner level
        for i in range(n // 2):
            nodes[base + i] = hashlib.sha256(_leaf_ser(leaves[2 * i], fmt) + _leaf_ser(leaves[2 * i + 1], fmt)).digest()
        for i in range(base - 1, -1, -1):
            nodes[i] = hashlib.sha256(nodes[2 * i + 1] + nodes[2 * i + 2]).digest()
        self.nodes = nodes

    def root(self) -> bytes:
        return self.nodes[0]

    def generate_proof(self, index: int) -> "MerklePath":
        sib_leaf = self.leaves[index ^ 1]
        cur = self.n // 2 - 1 + index // 2      # bottom-level inner node above the leaf
        path = []
        while cur != 0:
            sib = cur + 1 if cur % 2 == 1 else cur - 1
            path.append(self.nodes[sib])
            cur = (cur - 1) // 2
        path.reverse()                          # root-side first (A.5)
        return MerklePath(sib_leaf, path, index)


@dataclass
class MerklePath:
    leaf_sibling_hash: bytes
    auth_path: List[bytes]
    leaf_index: int

    def verify(self, root: bytes, leaf: bytes, fmt: Formats = DEFAULT_FORMATS) -> bool:
        idx = self.leaf_index
        l, r = (leaf, self.leaf_sibling_hash) if idx % 2 == 0 else (self.leaf_sibling_hash, leaf)
        cur = hashlib.sha256(_leaf_ser(l, fmt) + _leaf_ser(r, fmt)).digest()
        idx >>= 1
        for sib in reversed(self.auth_path):
            cur = hashlib.sha256(cur + sib).digest() if idx % 2 == 0 else hashlib.sha256(sib + cur).digest()
            idx >>= 1
        return cur == root


# --------------------------------------------------------------------------------------------
# Poseidon sponge (host side of Fiat-Shamir) -- parameters are an INPUT.
# --------------------------------------------------------------------------------------------
ARK_TEST_RNG_SEED = bytes([1, 0, 0, 0, 23, 0, 0, 0, 200, 1, 0, 0, 210, 30, 0, 0] + [0] * 16)


@dataclass
class PoseidonConfig:
    full_rounds: int
    partial_rounds: int
    alpha: int
    mds: List[List[int]]
    ark: List[List[int]]
    rate: int
    capacity: int


def test_sponge_config(f: Field = FR) -> PoseidonConfig:
    """ark_poly_commit::test_sponge with the deterministic ark_std::test_rng() (StdRng = ChaCha12
    seeded with ARK_TEST_RNG_SEED; ``DETERMINISTIC_TEST_RNG=1`` behaviour)."""
    rng = ChaChaRng(ARK_TEST_RNG_SEED, 12)
    ark = [[field_rand(rng, f) for _ in range(3)] for _ in range(8 + 31)]
    mds = [[1, 0, 1], [1, 1, 0], [0, 1, 1]]
    return PoseidonConfig(8, 31, 17, mds, ark, 2, 1)


class PoseidonSponge:
    def __init__(self, cfg: PoseidonConfig, f: Field = FR):
        self.cfg, self.f = cfg, f
        self.state = [0] * (cfg.rate + cfg.capacity)
        self.mode = ("absorbing", 0)

    def clone(self) -> "PoseidonSponge":
        s = PoseidonSponge(self.cfg, self.f)
        s.state = list(self.state)
        s.mode = self.mode
        return s

    def _permute(self) -> None:
        c, p = self.cfg, self.f.p
        st = self.state
        half = c.full_rounds // 2
        for rnd in range(c.full_rounds + c.partial_rounds):
            st = [(x + k) % p for x, k in zip(st, c.ark[rnd])]
            if rnd < half or rnd >= half + c.partial_rounds:
                st = [pow(x, c.alpha, p) for x in st]
            else:
                st[0] = pow(st[0], c.alpha, p)
            st = [sum(st[j] * c.mds[i][j] for j in range(len(st))) % p for i in range(len(st))]
        self.state = st

    def _absorb_internal(self, start: int, elems: List[int]) -> None:
        c, p = self.cfg, self.f.p
        rem = elems
        while True:
            if start + len(rem) <= c.rate:
                for i, e in enumerate(rem):
                    self.state[c.capacity + i + start] = (self.state[c.capacity + i + start] + e) % p
                self.mode = ("absorbing", start + len(rem))
                return
            take = c.rate - start
            for i, e in enumerate(rem[:take]):
                self.state[c.capacity + i + start] = (self.state[c.capacity + i + start] + e) % p
            self._permute()
            rem = rem[take:]
            start = 0

    def absorb_field_elements(self, elems: Sequence[int]) -> None:
        elems = [e % self.f.p for e in elems]
        if not elems:
            return
        kind, idx = self.mode
        if kind == "absorbing":
            if idx == self.cfg.rate:
                self._permute()
                idx = 0
            self._absorb_internal(idx, elems)
        else:
            self._permute()
            self._absorb_internal(0, elems)

    def absorb_bytes(self, data: bytes) -> None:
        """Absorb for Vec<u8>: u64_le(len) || bytes, packed 31 bytes/elem little-endian."""
        buf = struct.pack("<Q", len(data)) + data
        chunk = (self.f.bits - 1) // 8
        self.absorb_field_elements([int.from_bytes(buf[i:i + chunk], "little") for i in range(0, len(buf), chunk)])

    def squeeze_field_elements(self, n: int) -> List[int]:
        c = self.cfg
        out: List[int] = []
        kind, idx = self.mode
        if kind == "absorbing":
            self._permute()
            idx = 0
        elif idx == c.rate:
            self._permute()
            idx = 0
        while True:
            if idx + (n - len(out)) <= c.rate:
                need = n - len(out)
                out += self.state[c.capacity + idx: c.capacity + idx + need]
                self.mode = ("squeezing", idx + need)
                return out
            out += self.state[c.capacity + idx: c.capacity + c.rate]
            if len(out) != n:
                self._permute()
            idx = 0

    def squeeze_bytes(self, n: int) -> bytes:
        usable = (self.f.bits - 1) // 8
        ne = (n + usable - 1) // usable
        out = b"".join(e.to_bytes(8 * self.f.limbs, "little")[:usable] for e in self.squeeze_field_elements(ne))
        return out[:n]


# --------------------------------------------------------------------------------------------
# ArithmeticCircuit (ref: src/arithmetic_circuit/mod.rs) -- front end needed to feed the path
# --------------------------------------------------------------------------------------------
VAR, CONST, ADD, MUL = 0, 1, 2, 3


class ArithmeticCircuit:
    def __init__(self, f: Field = FR):
        self.f = f
        self.nodes: List[tuple] = []          # (VAR,label) | (CONST,value) | (ADD,l,r) | (MUL,l,r)
        self.constants: Dict[int, int] = {}
        self.variables: Dict[str, int] = {}

    # counts (ref 38-63)
    def num_nodes(self): return len(self.nodes)
    def num_constants(self): return len(self.constants)
    def num_variables(self): return len(self.variables)
    def last(self): return len(self.nodes) - 1
    def num_gates(self): return sum(1 for n in self.nodes if n[0] in (ADD, MUL))

    def _push(self, node) -> int:
        self.nodes.append(node)
        return len(self.nodes) - 1

    def constant(self, value: int) -> int:                       # ref 76-84
        value %= self.f.p
        if value in self.constants:
            return self.constants[value]
        idx = self._push((CONST, value))
        self.constants[value] = idx
        return idx

    def new_variable_with_label(self, label: str) -> int:        # ref 92-100
        idx = self._push((VAR, label))
        if label in self.variables:
            raise ValueError(f"Variable label already in use: {label}")
        self.variables[label] = idx
        return idx

    def new_variable(self) -> int:                               # ref 107-109
        return self.new_variable_with_label(f"var_{self.num_variables()}")

    def new_variables(self, n: int) -> List[int]:
        return [self.new_variable() for _ in range(n)]

    def get_variable(self, label: str) -> int:
        return self.variables[label]

    def add(self, l: int, r: int) -> int:                        # ref 125-131
        assert l < len(self.nodes) and r < len(self.nodes)
        return self._push((ADD, l, r))

    def mul(self, l: int, r: int) -> int:                        # ref 139-145
        assert l < len(self.nodes) and r < len(self.nodes)
        return self._push((MUL, l, r))

    def add_nodes(self, idx: Sequence[int]) -> int:              # ref 148-153 (left fold; empty panics)
        it = list(idx)
        if not it:
            raise ValueError("add_nodes of empty list (reference panics)")
        acc = it[0]
        for i in it[1:]:
            acc = self.add(acc, i)
        return acc

    def mul_nodes(self, idx: Sequence[int]) -> int:
        it = list(idx)
        acc = it[0]
        for i in it[1:]:
            acc = self.mul(acc, i)
        return acc

    def pow(self, node: int, e: int) -> int:                     # ref 164-203 square-and-multiply
        assert node < len(self.nodes) and e >= 1
        bits = bin(e)[2:]
        cur = node
        for b in bits[1:]:
            cur = self._push((MUL, cur, cur))
            if b == "1":
                cur = self._push((MUL, cur, node))
        return cur

    def indicator(self, node: int) -> int:                       # ref 206-221
        return self.pow(node, self.f.p - 1)

    def minus(self, node: int) -> int:                           # ref 224-227
        return self.mul(self.constant(self.f.p - 1), node)

    def scalar_product(self, left: Sequence[int], right: Sequence[int]) -> int:
        prods = [self._push((MUL, l, r)) for l, r in zip(left, right)]
        return self.add_nodes(prods)

    # evaluation (ref 247-358); iterative DFS instead of recursion, same visiting semantics
    def evaluation_trace_multioutput(self, vars_: Sequence[Tuple[int, int]], outputs: Sequence[int]) -> List[Optional[int]]:
        p = self.f.p
        vals: List[Optional[int]] = [n[1] if n[0] == CONST else None for n in self.nodes]
        for idx, v in vars_:
            if self.nodes[idx][0] != VAR:
                raise ValueError("Value supplied for non-variable node")
            vals[idx] = v % p
        for out in outputs:
            stack = [out]
            while stack:
                i = stack[-1]
                if vals[i] is not None:
                    stack.pop()
                    continue
                n = self.nodes[i]
                if n[0] == VAR:
                    raise ValueError("Uninitialised variable")
                l, r = n[1], n[2]
                if vals[l] is None:
                    stack.append(l)
                    continue
                if vals[r] is None:
                    stack.append(r)
                    continue
                vals[i] = (vals[l] + vals[r]) % p if n[0] == ADD else vals[l] * vals[r] % p
                stack.pop()
        return vals

    def evaluation_trace(self, vars_, node):
        return self.evaluation_trace_multioutput(vars_, [node])

    def evaluate_node(self, vars_, node):
        return self.evaluation_trace(vars_, node)[node]

    def evaluate(self, vars_):
        return self.evaluate_node(vars_, self.last())

    def evaluate_multioutput(self, vars_, outputs):
        tr = self.evaluation_trace_multioutput(vars_, outputs)
        oset = set(outputs)
        return [v for i, v in enumerate(tr) if i in oset and v is not None]

    # R1CS -> circuit (ref 455-520)
    @staticmethod
    def from_constraint_system(a_rows, b_rows, c_rows, num_vars_incl_one: int, f: Field = FR):
        """``a_rows`` etc. are lists of rows ``[(coeff, column)]`` exactly as
        ``ConstraintSystem::to_matrices`` yields them (column 0 = the constant one)."""
        circ = ArithmeticCircuit(f)
        one = circ.constant(1)
        circ.new_variables(num_vars_incl_one - 1)

        def compile_row(row):                                    # ref 501-520
            consts = [(circ.constant(c), v) for c, v in row]
            prods = [(ci + vi) if (ci == 0 or vi == 0) else circ.mul(ci, vi) for ci, vi in consts]
            return circ.add_nodes(prods)

        a = [compile_row(r) for r in a_rows]
        b = [compile_row(r) for r in b_rows]
        c = [compile_row(r) for r in c_rows]
        ab = [circ.mul(x, y) for x, y in zip(a, b)]
        minus_one = circ.constant(f.p - 1)
        mc = [circ.mul(x, minus_one) for x in c]
        outputs = [circ.add_nodes([x, y, one]) for x, y in zip(ab, mc)]
        return circ, outputs


def filter_constants(nodes: List[tuple]):
    """ref: src/arithmetic_circuit/mod.rs:546-607"""
    constants: Dict[int, int] = {}
    filtered: Dict[int, int] = {}
    removed = 0
    for i, n in enumerate(nodes):
        if n[0] == CONST:
            if n[1] in constants:
                removed += 1
            else:
                constants[n[1]] = i - removed
                filtered[i] = i - removed
        else:
            filtered[i] = i - removed
    out = []
    for i, n in enumerate(nodes):
        if n[0] == CONST:
            if i in filtered:
                out.append(n)
        elif n[0] == VAR:
            out.append(n)
        else:
            def upd(j):
                return constants[nodes[j][1]] if nodes[j][0] == CONST else filtered[j]
            out.append((n[0], upd(n[1]), upd(n[2])))
    return out, constants


def read_r1cs(path: str, f: Field = FR):
    """iden3 .r1cs v1 -> (A, B, C rows as [(coeff, wire)], n_wires).  Stands in for
    ark-circom + ``ConstraintSystem::to_matrices`` (ref: src/reader.rs:6-19): rows keep wires in
    ascending order with zero coefficients dropped, as ark-relations' sorted LinearCombination does."""
    data = open(path, "rb").read()
    assert data[:4] == b"r1cs"
    _version, nsec = struct.unpack_from("<II", data, 4)
    off = 12
    secs = {}
    for _ in range(nsec):
        typ, size = struct.unpack_from("<IQ", data, off)
        off += 12
        secs[typ] = data[off:off + size]
        off += size
    hdr = secs[1]
    fs = struct.unpack_from("<I", hdr, 0)[0]
    prime = int.from_bytes(hdr[4:4 + fs], "little")
    assert prime == f.p
    n_wires, _npo, _npi, _nprv = struct.unpack_from("<IIII", hdr, 4 + fs)
    n_constraints = struct.unpack_from("<I", hdr, 4 + fs + 16 + 8)[0]
    body = secs[2]
    o = 0
    mats = ([], [], [])
    for _ in range(n_constraints):
        for m in mats:
            nt = struct.unpack_from("<I", body, o)[0]
            o += 4
            acc: Dict[int, int] = {}
            for _ in range(nt):
                w = struct.unpack_from("<I", body, o)[0]
                o += 4
                c = int.from_bytes(body[o:o + fs], "little") % f.p
                o += fs
                acc[w] = (acc.get(w, 0) + c) % f.p
            m.append([(c, w) for w, c in sorted(acc.items()) if c != 0])
    return mats[0], mats[1], mats[2], n_wires


# --------------------------------------------------------------------------------------------
# Matrices (ref: src/matrices/mod.rs)
# --------------------------------------------------------------------------------------------
class SparseMatrix:
    def __init__(self, num_cols: int, rows: Optional[List[List[Tuple[int, int]]]] = None):
        self.num_cols = num_cols
        self.rows = rows if rows is not None else []

    def __eq__(self, o):
        return self.num_cols == o.num_cols and self.rows == o.rows

    def num_rows(self): return len(self.rows)

    @staticmethod
    def identity(size): return SparseMatrix(size, [[(1, i)] for i in range(size)])

    @staticmethod
    def zero(nr, nc): return SparseMatrix(nc, [[] for _ in range(nr)])

    def h_stack(self, other):
        assert self.num_rows() == other.num_rows()
        sh = self.num_cols
        return SparseMatrix(self.num_cols + other.num_cols,
                            [a + [(v, j + sh) for v, j in b] for a, b in zip(self.rows, other.rows)])

    def v_stack(self, other):
        assert self.num_cols == other.num_cols
        return SparseMatrix(self.num_cols, self.rows + other.rows)

    def neg(self, p: int = P):
        return SparseMatrix(self.num_cols, [[((-v) % p, j) for v, j in row] for row in self.rows])

    def row_mul(self, vec: Sequence[int], p: int = P) -> List[int]:   # ref 100-110
        res = [0] * self.num_cols
        for c, row in zip(vec, self.rows):
            for v, col in row:
                res[col] = (res[col] + c * v) % p
        return res


def dense_row_mul(rows: Sequence[Sequence[int]], vec: Sequence[int], p: int = P) -> List[int]:  # ref 138-149
    res = [0] * len(rows[0])
    for c, row in zip(vec, rows):
        for j, x in enumerate(row):
            res[j] = (res[j] + x * c) % p
    return res


def scalar_product(a, b, p: int = P) -> int:                            # ref src/utils.rs:13-15
    return sum(x * y for x, y in zip(a, b)) % p


# --------------------------------------------------------------------------------------------
# calculate_t (ark-poly-commit linear_codes/utils.rs)
# --------------------------------------------------------------------------------------------
def calculate_t(sec_param: int, distance: Tuple[int, int], codeword_len: int, f: Field = FR) -> int:
    residual = codeword_len / 2.0 ** f.bits
    rhs = math.log2(2.0 ** (-sec_param) - residual)
    nom = rhs - 1.0
    denom = math.log2(1.0 - 0.5 * distance[0] / distance[1])
    t = math.ceil(nom / denom)
    return t if t < codeword_len else codeword_len


# --------------------------------------------------------------------------------------------
# LigeroCircuit (ref: src/ligero/mod.rs)
# --------------------------------------------------------------------------------------------
@dataclass
class OpenedColumns:
    columns: List[List[int]]
    paths: List[MerklePath]


@dataclass
class LigeroProof:
    u_root: bytes
    preenc_u_lc: List[int]
    interleaved: OpenedColumns
    linear_poly: List[int]
    linear: OpenedColumns
    quadratic_poly: List[int]
    quadratic: OpenedColumns


def next_pow2(x: int) -> int:
    return 1 if x <= 1 else 1 << (x - 1).bit_length()


RHO_INV = 8   # hard-coded in the reference: src/ligero/mod.rs:284


class LigeroCircuit:
    def __init__(self, circuit: ArithmeticCircuit, outputs: Sequence[int], lam: int = 128,
                 fmt: Formats = DEFAULT_FORMATS):
        """ref: src/ligero/mod.rs:147-228"""
        f = circuit.f
        self.f, self.fmt = f, fmt
        # deep-ish copy so the caller's circuit is not mutated (reference takes ownership)
        c = ArithmeticCircuit(f)
        c.nodes = list(circuit.nodes)
        c.constants = dict(circuit.constants)
        c.variables = dict(circuit.variables)
        if 1 in c.constants:
            one_index, one_found = c.constants[1], True
        else:
            one_index, one_found = 1, False
        self.one_index, self.one_found = one_index, one_found
        if one_index != 0:
            self._insert_one(c, one_index, one_found)
        self.circuit = c
        sol_len = 1 + c.num_nodes() - c.num_constants() + len(outputs)          # ref 171
        self.sol_len = sol_len
        m = math.ceil(math.sqrt(float(sol_len)))                                 # ref 275-279
        k = next_pow2(m)
        n = RHO_INV * k                                                          # ref 283-294
        t = calculate_t(lam, (n - k + 1, n), n, f)
        self.m, self.k, self.n, self.t = m, k, n, t
        index_map = {0: 0}
        seen = 0
        for i, node in enumerate(c.nodes):
            if i == 0:
                continue
            if node[0] == CONST:
                seen += 1
            else:
                index_map[i] = i - seen
        self.index_map = index_map
        self.outputs = [self.bump_index(one_index, one_found, o) for o in outputs]
        self.a = self.generate_matrices(c, self.outputs, m * k, index_map)
        self.large_domain = Domain(n, f)
        self.small_domain = Domain(k, f)
        self.intermediate_domain = Domain(2 * k, f)

    @staticmethod
    def bump_index(one_index, one_found, index):                                # ref 230-242
        if one_found:
            if index < one_index:
                return index + 1
            if index == one_index:
                return 0
            return index
        return index + 1

    @classmethod
    def _insert_one(cls, c: ArithmeticCircuit, one_index, one_found):           # ref 244-271
        if one_found:
            del c.nodes[one_index]
        c.nodes.insert(0, (CONST, 1))
        b = lambda i: cls.bump_index(one_index, one_found, i)
        c.nodes = [(n[0], b(n[1]), b(n[2])) if n[0] in (ADD, MUL) else n for n in c.nodes]
        c.constants = {v: b(i) for v, i in c.constants.items()}
        c.constants[1] = 0
        c.variables = {l: b(i) for l, i in c.variables.items()}

    @staticmethod
    def generate_matrices(c: ArithmeticCircuit, outputs, num_cols, index_map) -> SparseMatrix:  # ref 296-433
        p = c.f.p
        nodes = c.nodes
        px, py, pz, padd = [], [], [], []
        neg1 = p - 1

        def add_row(l, r, own_col):
            if nodes[l][0] == CONST and nodes[r][0] == CONST:
                raise ValueError("Add(constant, constant) is not supported (the reference panics, mod.rs:325)")
            if nodes[l][0] == CONST:
                row = [(nodes[l][1], 0), (1, index_map[r])]
            elif nodes[r][0] == CONST:
                row = [(1, index_map[l]), (nodes[r][1], 0)]
            else:
                row = [(1, index_map[l]), (1, index_map[r])]
            row.append((neg1, own_col))
            return row

        def mul_rows(l, r):
            if nodes[l][0] == CONST and nodes[r][0] == CONST:
                # the reference hits `index_map.get(r_node).unwrap()` on None here (mod.rs:345; TODO at 148-150)
                raise ValueError("Mul(constant, constant) is not supported (the reference panics)")
            if nodes[l][0] == CONST:
                return [(nodes[l][1], 0)], [(1, index_map[r])]
            if nodes[r][0] == CONST:
                return [(1, index_map[l])], [(nodes[r][1], 0)]
            return [(1, index_map[l])], [(1, index_map[r])]

        for i, node in enumerate(nodes):
            if node[0] == VAR:
                px.append([]); py.append([]); pz.append([]); padd.append([])
            elif node[0] == ADD:
                px.append([]); py.append([]); pz.append([])
                padd.append(add_row(node[1], node[2], index_map[i]))
            elif node[0] == MUL:
                padd.append([])
                x, y = mul_rows(node[1], node[2])
                px.append(x); py.append(y); pz.append([(1, index_map[i])])
            elif i == 0:
                px.append([]); py.append([]); pz.append([]); padd.append([])
        for o in outputs:
            node = nodes[o]
            if node[0] == ADD:
                px.append([]); py.append([]); pz.append([])
                padd.append(add_row(node[1], node[2], 0))
            elif node[0] == MUL:
                padd.append([])
                x, y = mul_rows(node[1], node[2])
                px.append(x); py.append(y); pz.append([(1, 0)])
            else:
                raise ValueError("The output node must be an addition or multiplication gate")
        pad = num_cols - len(px)
        for mat in (px, py, pz, padd):
            mat.extend([] for _ in range(pad))
        upper_right = SparseMatrix(num_cols, px + py + pz).neg(p)
        upper = SparseMatrix.identity(3 * num_cols).h_stack(upper_right)
        lower = SparseMatrix.zero(num_cols, 3 * num_cols).h_stack(SparseMatrix(num_cols, padd))
        return upper.v_stack(lower)

    # ---- RS helpers (ref 998-1017)
    def reed_solomon_interpolate(self, msg):
        return self.small_domain.ifft(list(msg) + [0] * (self.k - len(msg)))

    def reed_solomon_evaluate(self, coeffs):
        return self.large_domain.fft(coeffs)

    def reed_solomon(self, msg):
        return self.reed_solomon_evaluate(self.reed_solomon_interpolate(msg))

    # ---- witness layout (ref 476-516)
    def witness_matrix(self, var_assignment) -> List[List[int]]:
        c, m, k = self.circuit, self.m, self.k
        sol = c.evaluation_trace_multioutput(var_assignment, self.outputs)
        if any(v is None for v in sol):
            raise ValueError("Uninitialised variable. Make sure the circuit only contains nodes upon "
                             "which the final output truly depends")
        x, y, z, w = [], [], [], []
        for i, (val, node) in enumerate(zip(sol, c.nodes)):
            if node[0] == CONST and i != 0:
                continue
            w.append(val)
            if node[0] == MUL:
                x.append(sol[node[1]]); y.append(sol[node[2]]); z.append(val)
            else:
                x.append(0); y.append(0); z.append(0)
        rows = []
        for vec in (x, y, z, w):
            vec = vec + [0] * (m * k - len(vec))
            assert len(vec) == m * k
            rows += [vec[i * k:(i + 1) * k] for i in range(m)]
        return rows

    # ---- prover (ref 435-578)
    def prove(self, var_assignment, sponge: PoseidonSponge) -> LigeroProof:
        va = [(self.bump_index(self.one_index, self.one_found, i), v) for i, v in var_assignment]
        return self.prove_inner(va, sponge)

    def prove_with_labels(self, var_assignment, sponge) -> LigeroProof:
        return self.prove_inner([(self.circuit.variables[l], v) for l, v in var_assignment], sponge)

    def prove_inner(self, var_assignment, sponge, trace: Optional[dict] = None) -> LigeroProof:
        preenc_u = self.witness_matrix(var_assignment)
        return self.prove_matrix(preenc_u, sponge, trace)

    def encode(self, preenc_u):
        coeffs = [self.reed_solomon_interpolate(r) for r in preenc_u]          # ref 521-526
        u = [self.reed_solomon_evaluate(r) for r in coeffs]                    # ref 528-533
        return coeffs, u

    def commit(self, u) -> MerkleTree:
        n_rows = len(u)
        leaves = [column_hash([u[i][j] for i in range(n_rows)], self.f, self.fmt) for j in range(self.n)]  # 536-542
        return MerkleTree(leaves, self.fmt)                                    # ref 544-549

    def prove_matrix(self, preenc_u, sponge, trace: Optional[dict] = None) -> LigeroProof:
        p, m, k = self.f.p, self.m, self.k
        coeffs, u = self.encode(preenc_u)
        tree = self.commit(u)
        u_root = tree.root()
        u_polys = [poly_trim(list(c)) for c in coeffs]                         # ref 555-558
        sponge.absorb_bytes(u_root)                                            # ref 560
        # Test-Interleaved (ref 646-669)
        seed = sponge.squeeze_bytes(32)
        r_int = get_field_elements_from_prng(4 * m, seed, self.f, self.fmt)
        lc = dense_row_mul(preenc_u, r_int, p)
        sponge.absorb_field_elements(lc)
        inter = self.open_columns(u, tree, sponge)
        # Test-Linear-Constraints (ref 712-747)
        seed_l = sponge.squeeze_bytes(32)
        r_lin = get_field_elements_from_prng(4 * m * k, seed_l, self.f, self.fmt)
        r_a = self.a.row_mul(r_lin, p)
        r_polys = [poly_trim(self.small_domain.ifft(r_a[i * k:(i + 1) * k])) for i in range(4 * m)]
        lin: List[int] = []
        for up, rp in zip(u_polys, r_polys):
            lin = poly_add(lin, poly_mul(up, rp, self.f), p)
        sponge.absorb_field_elements(lin)
        lin_open = self.open_columns(u, tree, sponge)
        # Test-Quadratic-Constraints (ref 832-859)
        seed_q = sponge.squeeze_bytes(32)
        r_q = get_field_elements_from_prng(m, seed_q, self.f, self.fmt)
        quad: List[int] = []
        for i in range(m):
            term = poly_scale(poly_sub(poly_mul(u_polys[i], u_polys[m + i], self.f), u_polys[2 * m + i], p), r_q[i], p)
            quad = poly_add(quad, term, p)
        sponge.absorb_field_elements(quad)
        quad_open = self.open_columns(u, tree, sponge)
        if trace is not None:
            trace.update(dict(preenc_u=preenc_u, coeffs=coeffs, u=u, leaves=tree.leaves, nodes=tree.nodes,
                              seed_interleaved=seed, r_interleaved=r_int, seed_linear=seed_l, r_linear=r_lin,
                              r_a=r_a, seed_quadratic=seed_q, r_quadratic=r_q))
        return LigeroProof(u_root, lc, inter, lin, lin_open, quad, quad_open)

    def open_columns(self, u, tree: MerkleTree, sponge) -> OpenedColumns:      # ref 935-955
        seed = sponge.squeeze_bytes(32)
        idx = get_distinct_indices_from_prng(self.n, self.t, seed, self.fmt)
        cols = [[row[j] for row in u] for j in idx]
        paths = [tree.generate_proof(j) for j in idx]
        return OpenedColumns(cols, paths)

    # ---- verifier (ref 613-644, 671-708, 749-830, 861-933, 957-996)
    def verify(self, proof: LigeroProof, sponge: PoseidonSponge) -> bool:
        sponge.absorb_bytes(proof.u_root)
        return (self._verify_interleaved(proof, sponge)
                and self._verify_linear(proof, sponge)
                and self._verify_quadratic(proof, sponge))

    def _verify_openings(self, oc: OpenedColumns, root, sponge) -> bool:
        seed = sponge.squeeze_bytes(32)
        idx = get_distinct_indices_from_prng(self.n, self.t, seed, self.fmt)
        if not (len(idx) == len(oc.columns) == len(oc.paths)):
            # izip! would silently truncate; treat a short proof as a failure of the t-opening check
            return False
        for col, i, path in zip(oc.columns, idx, oc.paths):
            h = column_hash(col, self.f, self.fmt)
            if path.leaf_index != i or not path.verify(root, h, self.fmt):
                return False
        return True

    def _verify_interleaved(self, proof, sponge) -> bool:
        p, m = self.f.p, self.m
        seed = sponge.squeeze_bytes(32)
        r = get_field_elements_from_prng(4 * m, seed, self.f, self.fmt)
        sponge.absorb_field_elements(proof.preenc_u_lc)
        if not self._verify_openings(proof.interleaved, proof.u_root, sponge):
            return False
        w = self.reed_solomon(proof.preenc_u_lc)
        for path, col in zip(proof.interleaved.paths, proof.interleaved.columns):
            assert len(col) == len(r)   # scalar_product_checked
            if w[path.leaf_index] != scalar_product(r, col, p):
                return False
        return True

    def _verify_linear(self, proof, sponge) -> bool:
        p, m, k, n = self.f.p, self.m, self.k, self.n
        q = proof.linear_poly
        seed = sponge.squeeze_bytes(32)
        r_lin = get_field_elements_from_prng(4 * m * k, seed, self.f, self.fmt)
        r_a = self.a.row_mul(r_lin, p)
        r_polys = [poly_trim(self.small_domain.ifft(r_a[i * k:(i + 1) * k])) for i in range(4 * m)]
        if poly_degree(q) >= 2 * k - 1:
            return False
        inter_evals = self.intermediate_domain.fft(list(q) + [0] * (2 * k - len(q)))
        cof = n // (2 * k)
        if sum(inter_evals[0::2]) % p != 0:
            return False
        sponge.absorb_field_elements(q)
        if not self._verify_openings(proof.linear, proof.u_root, sponge):
            return False
        r_evals = [self.reed_solomon_evaluate(rp) for rp in r_polys]
        for path, col in zip(proof.linear.paths, proof.linear.columns):
            j = path.leaf_index
            ev = inter_evals[j // cof] if j % cof == 0 else poly_eval(q, self.large_domain.element(j), p)
            if sum(re[j] * col[i] for i, re in enumerate(r_evals)) % p != ev:
                return False
        return True

    def _verify_quadratic(self, proof, sponge) -> bool:
        p, m, k, n = self.f.p, self.m, self.k, self.n
        q = proof.quadratic_poly
        seed = sponge.squeeze_bytes(32)
        r = get_field_elements_from_prng(m, seed, self.f, self.fmt)
        if poly_degree(q) >= 2 * k - 1:
            return False
        inter_evals = self.intermediate_domain.fft(list(q) + [0] * (2 * k - len(q)))
        if any(inter_evals[2 * c] != 0 for c in range(k)):
            return False
        cof = n // (2 * k)
        sponge.absorb_field_elements(q)
        if not self._verify_openings(proof.quadratic, proof.u_root, sponge):
            return False
        for path, col in zip(proof.quadratic.paths, proof.quadratic.columns):
            j = path.leaf_index
            lhs = inter_evals[j // cof] if j % cof == 0 else poly_eval(q, self.large_domain.element(j), p)
            rhs = sum(r[i] * (col[i] * col[i + m] - col[i + 2 * m]) for i in range(m)) % p
            if lhs != rhs:
                return False
        return True


# --------------------------------------------------------------------------------------------
# Reference test fixtures (ref: src/arithmetic_circuit/tests.rs:17-105)
# --------------------------------------------------------------------------------------------
def generate_bls12_377_circuit(f: Field = FQ377) -> ArithmeticCircuit:
    c = ArithmeticCircuit(f)
    one = c.constant(1)
    x = c.new_variable_with_label("x")
    y = c.new_variable_with_label("y")
    y2 = c.pow(y, 2)
    my2 = c.minus(y2)
    x3 = c.pow(x, 3)
    c.add_nodes([x3, one, my2, one])
    return c


def generate_lemniscate_circuit(f: Field = FR) -> ArithmeticCircuit:
    c = ArithmeticCircuit(f)
    one = c.constant(1)
    x = c.new_variable()
    y = c.new_variable()
    a = c.constant(120)
    b = c.constant(80)
    x2 = c.mul(x, x)
    y2 = c.mul(y, y)
    ax2 = c.mul(a, x2)
    by2 = c.mul(b, y2)
    max2 = c.minus(ax2)
    s = c.add(x2, y2)
    d = c.add(by2, max2)
    s2 = c.mul(s, s)
    c.add_nodes([s2, d, one])
    return c


def generate_3_by_3_determinant_circuit(f: Field = FR) -> ArithmeticCircuit:
    c = ArithmeticCircuit(f)
    one = c.constant(1)
    v = c.new_variables(9)
    det = c.new_variable()
    aei = c.mul_nodes([v[0], v[4], v[8]])
    bfg = c.mul_nodes([v[1], v[5], v[6]])
    cdh = c.mul_nodes([v[2], v[3], v[7]])
    ceg = c.mul_nodes([v[2], v[4], v[6]])
    bdi = c.mul_nodes([v[1], v[3], v[8]])
    afh = c.mul_nodes([v[0], v[5], v[7]])
    s1 = c.add_nodes([aei, bfg, cdh])
    s2 = c.add_nodes([ceg, bdi, afh])
    ms2 = c.minus(s2)
    mdet = c.minus(det)
    c.add_nodes([s1, ms2, mdet, one])
    return c


def synthetic_circuit(gates: int, seed: int = 1, f: Field = FR):
    """Seeded random Add/Mul circuit of exactly ``gates`` gates satisfying SURVEY 8(d)'s rules:
    every node feeds the output (a chain: gate i consumes gate i-1 and one earlier node), depth is
    linear but the evaluator here is iterative; the last gate is ``Add(prev, const)`` making the
    output 1; no gate has two constant operands.  Returns (circuit, outputs, var_assignment)."""
    import random
    rnd = random.Random(seed)
    p = f.p
    c = ArithmeticCircuit(f)
    c.constant(1)
    v0, v1 = c.new_variable(), c.new_variable()
    vals = {0: 1, v0: rnd.randrange(1, p), v1: rnd.randrange(1, p)}
    assign = [(v0, vals[v0]), (v1, vals[v1])]
    prev = c.mul(v0, v1)
    vals[prev] = vals[v0] * vals[v1] % p
    nonconst = [v0, v1, prev]
    for _ in range(gates - 2):
        other = nonconst[rnd.randrange(len(nonconst))]
        if rnd.random() < 0.5:
            g = c.add(prev, other)
            vals[g] = (vals[prev] + vals[other]) % p
        else:
            g = c.mul(prev, other)
            vals[g] = vals[prev] * vals[other] % p
        nonconst.append(g)
        prev = g
    k = c.constant((1 - vals[prev]) % p)
    out = c.add(prev, k)
    return c, [out], assign
