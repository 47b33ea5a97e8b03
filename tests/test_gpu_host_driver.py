"""The C++ host driver (LigeroCircuit::new / prove / verify mirror) on the GPU: proofs are byte-identical to
the oracle's, both verifiers accept each other's proofs, perturbed witnesses are rejected."""
import hashlib

import numpy as np
import pytest

import ligero_b200 as lb
from ligero_b200 import fr_to_limbs, limbs_to_fr
from oracle import ligero_oracle as O
from oracle import wire
from tests.golden_util import expected, load_r1cs

pytestmark = pytest.mark.gpu
P = O.P


def osponge():
    return O.PoseidonSponge(O.test_sponge_config())


def mirror_circuit(oc: "O.ArithmeticCircuit") -> "lb.ArithmeticCircuit":
    """rebuild an oracle circuit node by node through the product's API"""
    c = lb.ArithmeticCircuit()
    for n in oc.nodes:
        if n[0] == O.VAR:
            idx = c.new_variable_with_label(n[1])
        elif n[0] == O.CONST:
            idx = c.constant(n[1])
        elif n[0] == O.ADD:
            idx = c.add(n[1], n[2])
        else:
            idx = c.mul(n[1], n[2])
    assert c.num_nodes() == oc.num_nodes() and c.num_constants() == oc.num_constants()
    return c


CASES = {
    "lemniscate": lambda: (O.generate_lemniscate_circuit(), None, [(1, 8), (2, 4)]),
    "determinant": lambda: (O.generate_3_by_3_determinant_circuit(), None,
                            [(1, 2), (2, 0), (3, P - 1), (4, 3), (5, 5), (6, 2), (7, P - 4), (8, 1), (9, 4), (10, 13)]),
    "synthetic_600": lambda: O.synthetic_circuit(600, seed=3),
}


@pytest.mark.parametrize("name", list(CASES))
def test_proof_bytes_equal_oracle_and_golden(gpu_ctx, name):
    oc, outs, assign = CASES[name]()
    outs = outs or [oc.last()]
    olc = O.LigeroCircuit(oc, outs)
    want = wire.serialize_proof(olc.prove(assign, osponge()))
    lc = lb.LigeroCircuit(gpu_ctx, mirror_circuit(oc), outs)
    assert (lc.m, lc.k, lc.n, lc.t, lc.sol_len) == (olc.m, olc.k, olc.n, olc.t, olc.sol_len)
    proof = lc.prove(assign, lb.PoseidonSponge.test_sponge())
    got = proof.to_bytes()
    assert got == want
    assert hashlib.sha256(got).hexdigest() == expected()[name]["proof_sha256"]
    # cross verification
    assert lc.verify(proof, lb.PoseidonSponge.test_sponge())
    assert olc.verify(wire.deserialize_proof(got), osponge())
    assert lc.verify(lb.LigeroProof.from_bytes(want), lb.PoseidonSponge.test_sponge())
    # perturbed witness: rejected by both verifiers (src/ligero/tests.rs:160-170)
    bad = list(assign)
    bad[0] = (bad[0][0], (bad[0][1] + 1) % P)
    bad_proof = lc.prove(bad, lb.PoseidonSponge.test_sponge())
    assert not lc.verify(bad_proof, lb.PoseidonSponge.test_sponge())
    assert not olc.verify(wire.deserialize_proof(bad_proof.to_bytes()), osponge())


def test_tampered_proofs_rejected(gpu_ctx):
    oc = O.generate_lemniscate_circuit()
    lc = lb.LigeroCircuit(gpu_ctx, mirror_circuit(oc), [oc.last()])
    blob = bytearray(lc.prove([(1, 8), (2, 4)], lb.PoseidonSponge.test_sponge()).to_bytes())
    assert lc.verify(lb.LigeroProof.from_bytes(bytes(blob)), lb.PoseidonSponge.test_sponge())
    for pos in (9, 60, len(blob) // 2, len(blob) - 40):
        t = bytearray(blob)
        t[pos] ^= 1
        try:
            pr = lb.LigeroProof.from_bytes(bytes(t))
        except lb.LigeroB200Error:
            continue        # malformed encodings are refused outright
        assert not lc.verify(pr, lb.PoseidonSponge.test_sponge())
    # a different sponge state must not verify
    sp = lb.PoseidonSponge.test_sponge()
    sp.absorb_bytes(b"x")
    assert not lc.verify(lb.LigeroProof.from_bytes(bytes(blob)), sp)


def test_multioutput_with_labels(gpu_ctx):
    """src/ligero/tests.rs:246-361"""
    c = lb.ArithmeticCircuit()
    x = c.new_variable_with_label("x")
    y = c.new_variable_with_label("y")
    c1, c2, c3 = c.constant(P - 8), c.constant(P - 63), c.constant(P - 6)
    x2 = c.mul(x, x)
    y3 = c.pow(y, 3)
    s = c.add(x, y)
    outs = [c.add(x2, c1), c.add(y3, c2), c.add(s, c3)]
    lc = lb.LigeroCircuit(gpu_ctx, c, outs)
    assert lc.m * lc.k == 16
    proof = lc.prove_with_labels([("x", 3), ("y", 4)], lb.PoseidonSponge.test_sponge())
    assert lc.verify(proof, lb.PoseidonSponge.test_sponge())
    with pytest.raises(lb.LigeroB200Error):
        lc.prove_with_labels([("x", 3), ("nope", 4)], lb.PoseidonSponge.test_sponge())


def test_reference_failure_modes(gpu_ctx):
    # constant x constant gate: the reference panics in generate_matrices (mod.rs:345)
    c = lb.ArithmeticCircuit()
    c.constant(1)
    v = c.new_variable()
    a, b = c.constant(27), c.constant(P - 1)
    g = c.mul(a, b)
    out = c.add(g, v)
    with pytest.raises(lb.LigeroB200Error):
        lb.LigeroCircuit(gpu_ctx, c, [out])
    # unreachable node: "Uninitialised variable ..." (mod.rs:476-478)
    c = lb.ArithmeticCircuit()
    c.constant(1)
    v, w = c.new_variable(), c.new_variable()
    out = c.add(v, v)
    lc = lb.LigeroCircuit(gpu_ctx, c, [out])
    with pytest.raises(lb.LigeroB200Error):
        lc.prove([(v, 5)], lb.PoseidonSponge.test_sponge())
    # value for a non-variable node
    with pytest.raises(lb.LigeroB200Error):
        lc.prove([(out, 5), (w, 1)], lb.PoseidonSponge.test_sponge())
    # duplicate label, operand out of range
    with pytest.raises(lb.LigeroB200Error):
        c.new_variable_with_label("var_0")
    with pytest.raises(lb.LigeroB200Error):
        c.add(0, 999)


def test_circom_multiplication_and_poseidon_from_golden(gpu_ctx):
    """BASELINE.json configs 1-2 through from_constraint_system, against the committed golden digests."""
    exp = expected()
    for name in ("multiplication", "poseidon"):
        a, b, c, nw, wit = load_r1cs(name)
        circ, outs = lb.ArithmeticCircuit.from_constraint_system(a, b, c, nw)
        va = list(enumerate(wit))[1:]
        assert all(v == 1 for v in circ.evaluate_multioutput(va, outs))
        lc = lb.LigeroCircuit(gpu_ctx, circ, outs)
        e = exp[name]
        assert (lc.m, lc.k, lc.n, lc.t, lc.sol_len) == (e["m"], e["k"], e["n"], e["t"], e["sol_len"])
        proof = lc.prove(va, lb.PoseidonSponge.test_sponge())
        blob = proof.to_bytes()
        assert len(blob) == e["proof_len"]
        assert hashlib.sha256(blob).hexdigest() == e["proof_sha256"]
        assert lc.verify(proof, lb.PoseidonSponge.test_sponge())
        bad = list(va)
        bad[0] = (bad[0][0], (bad[0][1] + 1) % P)
        assert not lc.verify(lc.prove(bad, lb.PoseidonSponge.test_sponge()), lb.PoseidonSponge.test_sponge())
    # cube: constant x constant gate after compilation -> refused like the reference
    oc_nodes = 15
    cube_a = [[(P - 1, 1)], [(1, 1)]]
    cube_b = [[(1, 1)], [(1, 2)]]
    cube_c = [[(P - 1, 2)], [(27, 0)]]
    circ, outs = lb.ArithmeticCircuit.from_constraint_system(cube_a, cube_b, cube_c, 3)
    assert circ.num_nodes() == oc_nodes            # src/arithmetic_circuit/tests.rs:239
    assert circ.evaluate_multioutput([(1, 3), (2, 9)], outs) == [1, 1]
    with pytest.raises(lb.LigeroB200Error):
        lb.LigeroCircuit(gpu_ctx, circ, outs)


def test_sponge_matches_oracle(gpu_ctx):
    a, b = lb.PoseidonSponge.test_sponge(), osponge()
    a.absorb_bytes(b"\x01" * 32)
    b.absorb_bytes(b"\x01" * 32)
    assert a.squeeze_bytes(32) == b.squeeze_bytes(32)
    elems = [3, 5, P - 1, 0, 12345]
    a.absorb_field_elements(elems)
    b.absorb_field_elements(elems)
    assert a.squeeze_bytes(32) == b.squeeze_bytes(32)
    assert a.squeeze_bytes(32) == b.squeeze_bytes(32)
    a.absorb_field_elements([])
    b.absorb_field_elements([])
    c = a.clone()
    assert a.squeeze_bytes(7) == b.squeeze_bytes(7) == c.squeeze_bytes(7)


def test_witness_matrix_matches_oracle(gpu_ctx):
    oc, outs, assign = O.synthetic_circuit(300, seed=9)
    olc = O.LigeroCircuit(oc, outs)
    va = [(olc.bump_index(olc.one_index, olc.one_found, i), v) for i, v in assign]
    want = [x for r in olc.witness_matrix(va) for x in r]
    lc = lb.LigeroCircuit(gpu_ctx, mirror_circuit(oc), outs)
    assert limbs_to_fr(lc.witness_matrix(assign)) == want


def test_repeated_squaring_10_proof_equals_oracle(gpu_ctx):
    """BASELINE config 2: circom/repeated_squaring_10 through from_constraint_system on the GPU prover: proof bytes equal
    the oracle's, both verifiers accept, a wrong output is rejected."""
    from tests.util import repeated_squaring_r1cs
    a, b, c, nw, wit = repeated_squaring_r1cs(10, 3)
    va = list(enumerate(wit))[1:]
    oc, oouts = O.ArithmeticCircuit.from_constraint_system(a, b, c, nw)
    olc = O.LigeroCircuit(oc, oouts)
    want = wire.serialize_proof(olc.prove(va, osponge()))
    circ, outs = lb.ArithmeticCircuit.from_constraint_system(a, b, c, nw)
    assert outs == oouts and circ.num_nodes() == oc.num_nodes()
    lc = lb.LigeroCircuit(gpu_ctx, circ, outs)
    proof = lc.prove(va, lb.PoseidonSponge.test_sponge())
    assert proof.to_bytes() == want
    assert lc.verify(proof, lb.PoseidonSponge.test_sponge())
    assert olc.verify(wire.deserialize_proof(proof.to_bytes()), osponge())
    bad = list(va)
    bad[0] = (1, (wit[1] + 1) % P)
    assert not lc.verify(lc.prove(bad, lb.PoseidonSponge.test_sponge()), lb.PoseidonSponge.test_sponge())


def test_rust_parity_dump_when_present(gpu_ctx):
    """GPU prover against the REAL reference: when rust/ligero-parity-dump/run.sh has produced
    tests/golden/rust_parity_dump.json, the serialized GPU proofs of lemniscate / determinant / poseidon must be the
    reference's bytes (src/ligero/tests.rs:197-243, 365-415 with DETERMINISTIC_TEST_RNG=1).  Skipped without the file."""
    import json
    from tests.test_golden import rust_dump_path
    path = rust_dump_path()
    if path is None:
        pytest.skip("no reference-produced dump: parity unpinned (see rust/ligero-parity-dump/run.sh)")
    dump = json.load(open(path))
    for name in ("lemniscate", "determinant"):
        oc, _, assign = CASES[name]()
        lc = lb.LigeroCircuit(gpu_ctx, mirror_circuit(oc), [oc.last()])
        assert lc.prove(assign, lb.PoseidonSponge.test_sponge()).to_bytes().hex() == dump[name]["proof_hex"], name
    a, b, c, nw, wit = load_r1cs("poseidon")
    circ, outs = lb.ArithmeticCircuit.from_constraint_system(a, b, c, nw)
    lc = lb.LigeroCircuit(gpu_ctx, circ, outs)
    blob = lc.prove(list(enumerate(wit))[1:], lb.PoseidonSponge.test_sponge()).to_bytes()
    assert blob.hex() == dump["poseidon"]["proof_hex"], "poseidon"
