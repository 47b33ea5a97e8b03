"""The C oracle (CPU baseline + large-size checker) agrees with the Python oracle and hashlib."""
import hashlib
import random

import numpy as np

from ligero_b200 import fr_to_limbs, limbs_to_fr
from oracle import cref
from oracle import ligero_oracle as O

P = O.P


def test_hashes_vs_hashlib():
    rnd = random.Random(3)
    for L in [0, 1, 31, 32, 55, 56, 63, 64, 65, 119, 120, 127, 128, 129, 1000]:
        d = bytes(rnd.randrange(256) for _ in range(L))
        assert cref.sha256(d) == hashlib.sha256(d).digest()
        assert cref.blake2s(d) == hashlib.blake2s(d).digest()


def test_commit_vs_python_oracle():
    rnd = random.Random(4)
    for R, k, rho in [(5, 16, 8), (2, 2, 8), (3, 64, 4), (1, 8, 2)]:
        msg = [[rnd.randrange(P) for _ in range(k)] for _ in range(R)]
        dk, dn = O.Domain(k), O.Domain(rho * k)
        u = [dn.fft(dk.ifft(r)) for r in msg]
        leaves = [O.column_hash([u[i][j] for i in range(R)]) for j in range(rho * k)]
        tree = O.MerkleTree(leaves)
        res = cref.commit(fr_to_limbs([x for r in msg for x in r]), R, k, rho, threads=3, want_u=True, want_tree=True)
        assert res["root"] == tree.root()
        assert [bytes(x) for x in res["leaves"]] == leaves
        assert [bytes(x) for x in res["nodes"]] == tree.nodes
        for i in range(R):
            assert limbs_to_fr(res["u"][i]) == u[i]


def test_commit_thread_count_invariant():
    a = np.random.default_rng(1).integers(0, 2 ** 62, size=(24 * 64, 4), dtype=np.uint64)
    a[:, 3] &= (1 << 60) - 1
    assert cref.commit(a, 24, 64, 8, threads=1)["root"] == cref.commit(a, 24, 64, 8, threads=7)["root"]


def test_prng_expansion():
    seed = bytes(range(32))
    assert limbs_to_fr(cref.expand_fr(seed, 300)) == O.get_field_elements_from_prng(300, seed)
    for n, t in [(32, 32), (1024, 156), (64, 40), (16, 3), (65536, 156)]:
        assert list(cref.expand_indices(seed, n, t)) == O.get_distinct_indices_from_prng(n, t, seed)


def test_test_polynomials():
    rnd = random.Random(5)
    m, k = 3, 8
    U = [[rnd.randrange(P) for _ in range(k)] for _ in range(4 * m)]
    ra = [[rnd.randrange(P) for _ in range(k)] for _ in range(4 * m)]
    dk = O.Domain(k)
    lin = []
    for a, b in zip(U, ra):
        lin = O.poly_add(lin, O.poly_mul(O.poly_trim(dk.ifft(a)), O.poly_trim(dk.ifft(b))))
    flatU = fr_to_limbs([x for r in U for x in r])
    got = O.poly_trim(limbs_to_fr(cref.linear_poly(flatU, fr_to_limbs([x for r in ra for x in r]), 4 * m, k, threads=2)))
    assert got == lin
    rq = [rnd.randrange(P) for _ in range(m)]
    quad = []
    for i in range(m):
        t = O.poly_sub(O.poly_mul(O.poly_trim(dk.ifft(U[i])), O.poly_trim(dk.ifft(U[m + i]))), O.poly_trim(dk.ifft(U[2 * m + i])))
        quad = O.poly_add(quad, O.poly_scale(t, rq[i]))
    assert O.poly_trim(limbs_to_fr(cref.quadratic_poly(flatU, fr_to_limbs(rq), m, k))) == quad
    r4 = [rnd.randrange(P) for _ in range(4 * m)]
    assert limbs_to_fr(cref.row_mul(flatU, fr_to_limbs(r4), 4 * m, k)) == O.dense_row_mul(U, r4)


def test_reference_row_mul_known_answers():
    """src/matrices/mod.rs:180-207"""
    m = fr_to_limbs([1, 2, 8, 3, 4, 5])
    v = fr_to_limbs([P - 5, 17])
    assert limbs_to_fr(cref.row_mul(m, v, 2, 3)) == [46, 58, 45]
    # sparse: rows [(1,0),(8,2)], [(4,1),(5,2)]
    got = cref.sparse_row_mul([0, 2, 4], [0, 2, 1, 2], fr_to_limbs([1, 8, 4, 5]), v, 2, 3)
    assert limbs_to_fr(got) == [P - 5, 68, 45]
