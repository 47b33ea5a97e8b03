"""Regenerates tests/golden/*.  Run HERE (needs /root/reference for the circom fixtures):

    python tests/golden/make_golden.py

Outputs
  poseidon_r1cs.npz   -- the reference's circom/poseidon/poseidon.r1cs as CSR arrays + witness.json values
                         (inputs only: lets the GPU box build the circuit without /root/reference)
  multiplication_r1cs.npz -- same for circom/multiplication.r1cs
  expected.json       -- oracle outputs on the fixtures and on the in-code test circuits: shapes, Merkle
                         root, SHA-256 of the serialised proof, leading coefficients.  These are REGRESSION
                         pins of the oracle (ligero_oracle.py at generation time), not reference outputs:
                         the reference cannot be executed here (no Rust toolchain) -- parity stays unpinned.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ligero_oracle as O  # noqa: E402
from oracle import wire  # noqa: E402

REF = "/root/reference"
P = O.P


def r1cs_to_npz(path_r1cs, out, witness=None):
    a, b, c, nw = O.read_r1cs(path_r1cs)
    d = {"n_wires": np.array([nw], dtype=np.uint64)}
    for name, mat in (("a", a), ("b", b), ("c", c)):
        ptr, cols, vals = [0], [], []
        for row in mat:
            for coeff, col in row:
                cols.append(col)
                vals.append([(coeff >> (64 * i)) & ((1 << 64) - 1) for i in range(4)])   # canonical limbs
            ptr.append(len(cols))
        d[name + "_ptr"] = np.array(ptr, dtype=np.uint64)
        d[name + "_col"] = np.array(cols, dtype=np.uint64)
        d[name + "_val"] = np.array(vals, dtype=np.uint64).reshape(-1, 4)
    if witness is not None:
        d["witness"] = np.array([[(w >> (64 * i)) & ((1 << 64) - 1) for i in range(4)] for w in witness], dtype=np.uint64)
    np.savez_compressed(out, **d)


def summarize(lc, proof):
    blob = wire.serialize_proof(proof)
    return {"m": lc.m, "k": lc.k, "n": lc.n, "t": lc.t, "sol_len": lc.sol_len, "u_root": proof.u_root.hex(),
            "proof_sha256": hashlib.sha256(blob).hexdigest(), "proof_len": len(blob),
            "preenc_u_lc_0": hex(proof.preenc_u_lc[0]), "linear_poly_len": len(proof.linear_poly),
            "quadratic_poly_len": len(proof.quadratic_poly),
            "interleaved_indices_head": [p.leaf_index for p in proof.interleaved.paths[:8]]}


def sponge():
    return O.PoseidonSponge(O.test_sponge_config())


def main():
    expected = {}
    wit = [int(s) for s in json.load(open(f"{REF}/circom/poseidon/witness.json"))]
    r1cs_to_npz(f"{REF}/circom/poseidon/poseidon.r1cs", os.path.join(HERE, "poseidon_r1cs.npz"), wit)
    r1cs_to_npz(f"{REF}/circom/multiplication.r1cs", os.path.join(HERE, "multiplication_r1cs.npz"), [1, 6, 3, 2])
    # poseidon (BASELINE.json config 1)
    a, b, c, nw = O.read_r1cs(f"{REF}/circom/poseidon/poseidon.r1cs")
    circ, outs = O.ArithmeticCircuit.from_constraint_system(a, b, c, nw)
    lc = O.LigeroCircuit(circ, outs)
    va = list(enumerate(wit))[1:]
    proof = lc.prove(va, sponge())
    assert lc.verify(proof, sponge())
    expected["poseidon"] = summarize(lc, proof)
    # multiplication (config 2)
    a, b, c, nw = O.read_r1cs(f"{REF}/circom/multiplication.r1cs")
    circ, outs = O.ArithmeticCircuit.from_constraint_system(a, b, c, nw)
    lc = O.LigeroCircuit(circ, outs)
    proof = lc.prove([(1, 6), (2, 3), (3, 2)], sponge())
    assert lc.verify(proof, sponge())
    expected["multiplication"] = summarize(lc, proof)
    # in-code circuits of the reference's tests
    circ = O.generate_lemniscate_circuit()
    lc = O.LigeroCircuit(circ, [circ.last()])
    expected["lemniscate"] = summarize(lc, lc.prove([(1, 8), (2, 4)], sponge()))
    circ = O.generate_3_by_3_determinant_circuit()
    lc = O.LigeroCircuit(circ, [circ.last()])
    vals = [(1, 2), (2, 0), (3, P - 1), (4, 3), (5, 5), (6, 2), (7, P - 4), (8, 1), (9, 4), (10, 13)]
    expected["determinant"] = summarize(lc, lc.prove(vals, sponge()))
    circ, outs, assign = O.synthetic_circuit(600, seed=3)
    lc = O.LigeroCircuit(circ, outs)
    expected["synthetic_600"] = summarize(lc, lc.prove(assign, sponge()))
    # primitive vectors
    seed = bytes(range(32))
    expected["expand_fr_head"] = [hex(x) for x in O.get_field_elements_from_prng(4, seed)]
    expected["indices_1024_156_head"] = O.get_distinct_indices_from_prng(1024, 156, seed)[:10]
    expected["test_sponge_ark_0"] = hex(O.test_sponge_config().ark[0][0])
    json.dump(expected, open(os.path.join(HERE, "expected.json"), "w"), indent=1)
    print(json.dumps(expected, indent=1)[:1500])


if __name__ == "__main__":
    main()
