"""BASELINE configs 3 and 4 pinned to the ORACLE, not only to self-verification (src/ligero/mod.rs:521-551, 646-669,
712-747, 832-859, 935-955): the pieces of a whole lg_prove proof of the seeded synthetic circuits are recomputed from
the same witness matrix by the C restatement of the reference's schedule (oracle/ligero_ref.c: per-row iFFT_k + FFT_n,
BLAKE2s per column, SHA-256 tree, per-row polynomial products for the tests) with the constraint matrix A built by the
Python restatement of generate_matrices (oracle/ligero_oracle.py), replaying the Fiat-Shamir transcript.

2^20 gates: everything (root, preenc_u_lc, both polynomials, every opened column and path).
2^24 gates: root against the CPU-oracle pin of the benchmark matrix, plus sampled rows / columns (test_gpu_full_size.py)."""
import json
import os

import numpy as np
import pytest

import ligero_b200 as lb
from ligero_b200 import limbs_to_fr
from oracle import cref, wire
from oracle import ligero_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def to_oracle_circuit(circ):
    oc = O.ArithmeticCircuit(O.FR)
    for i in range(circ.num_nodes()):
        nd = circ.node(i)
        if nd[0] == "var":
            oc.new_variable()
        elif nd[0] == "const":
            oc.constant(nd[1])
        elif nd[0] == "add":
            oc.add(nd[1], nd[2])
        else:
            oc.mul(nd[1], nd[2])
    return oc


def csr_of(a):
    """CSR arrays (row_ptr, col_idx, Montgomery-limb values) of an oracle SparseMatrix for cref.sparse_row_mul"""
    row_ptr = np.zeros(len(a.rows) + 1, dtype=np.uint64)
    cols, vals, table = [], [], {}
    for i, row in enumerate(a.rows):
        for v, j in row:
            cols.append(j)
            vals.append(table.setdefault(v, len(table)))
        row_ptr[i + 1] = len(cols)
    uniq = lb.fr_to_limbs(sorted(table, key=table.get))
    return row_ptr, np.array(cols, dtype=np.uint64), uniq[np.array(vals, dtype=np.int64)]


def check_pieces(gpu_ctx, log_gates):
    gates = 1 << log_gates
    circ, out, assign = lb.ArithmeticCircuit.synthetic(gates, 2024)
    lc = lb.LigeroCircuit(gpu_ctx, circ, [out])
    proof = wire.deserialize_proof(lc.prove(assign, lb.PoseidonSponge.test_sponge()).to_bytes())
    pre = lc.witness_matrix_device(assign).cpu().numpy().view(np.uint64)         # [X;Y;Z;W], 4mk x 4 limbs
    m, k, n, t = lc.m, lc.k, lc.n, lc.t
    OL = O.LigeroCircuit(to_oracle_circuit(circ), [out])
    assert (OL.m, OL.k, OL.n, OL.t) == (m, k, n, t)
    # commitment (521-551)
    ref = cref.commit(pre, 4 * m, k, 8, want_u=True, want_tree=True)
    assert proof.u_root == ref["root"]
    tree = O.MerkleTree([bytes(x) for x in ref["leaves"]])
    assert tree.root() == ref["root"]
    sponge = O.PoseidonSponge(O.test_sponge_config())
    sponge.absorb_bytes(ref["root"])

    def check_opening(opened):
        idx = cref.expand_indices(sponge.squeeze_bytes(32), n, t)
        assert [p.leaf_index for p in opened.paths] == [int(j) for j in idx]
        for q, j in enumerate(idx):
            assert opened.columns[q] == limbs_to_fr(ref["u"][:, int(j)]), f"opened column {j}"
            want = tree.generate_proof(int(j))
            got = opened.paths[q]
            assert (got.leaf_sibling_hash, got.auth_path) == (want.leaf_sibling_hash, want.auth_path), f"path {j}"

    # Test-Interleaved (646-669)
    r = cref.expand_fr(sponge.squeeze_bytes(32), 4 * m)
    want_lc = limbs_to_fr(cref.row_mul(pre, r, 4 * m, k))
    assert proof.preenc_u_lc == want_lc
    sponge.absorb_field_elements(want_lc)
    check_opening(proof.interleaved)
    # Test-Linear-Constraints (712-747): r_a = r^T A with the oracle's A
    r_lin = cref.expand_fr(sponge.squeeze_bytes(32), 4 * m * k)
    row_ptr, col_idx, vals = csr_of(OL.a)
    r_a = cref.sparse_row_mul(row_ptr, col_idx, vals, r_lin, 4 * m * k, 4 * m * k)
    want_lin = O.poly_trim(limbs_to_fr(cref.linear_poly(pre, r_a, 4 * m, k)))
    assert proof.linear_poly == want_lin
    sponge.absorb_field_elements(want_lin)
    check_opening(proof.linear)
    # Test-Quadratic-Constraints (832-859)
    r_q = cref.expand_fr(sponge.squeeze_bytes(32), m)
    want_quad = O.poly_trim(limbs_to_fr(cref.quadratic_poly(pre, r_q, m, k)))
    assert proof.quadratic_poly == want_quad
    sponge.absorb_field_elements(want_quad)
    check_opening(proof.quadratic)


def test_2p16_gate_proof_pieces_equal_the_oracle(gpu_ctx):
    check_pieces(gpu_ctx, 16)


def test_2p20_gate_proof_pieces_equal_the_oracle(gpu_ctx):
    """BASELINE config 3 (m = 1025, k = 2048, n = 16384, t = 156)"""
    check_pieces(gpu_ctx, 20)


@pytest.mark.parametrize("log_gates", [20, 24])
def test_benchmark_matrix_root_equals_the_cpu_oracle_pin(gpu_ctx, log_gates):
    """bench.py's matrix (ligero_b200.synthetic, the same at every GPU count): the GPU root equals the root the CPU oracle
    computed over the WHOLE matrix (scripts/pin_full_size_root.py -> tests/golden/full_size_root.json)."""
    import bench
    from ligero_b200.synthetic import matrix_rows_torch
    pin = json.load(open(os.path.join(ROOT, "tests", "golden", "full_size_root.json")))[f"2^{log_gates}/seed{bench.SEED}"]
    R, k, n, m = bench.shape_for_gates(log_gates)
    assert (pin["rows"], pin["k"]) == (R, k)
    msg = matrix_rows_torch(bench.SEED, range(R), k, "cuda")
    cm = gpu_ctx.commit(msg, R, k, 8)
    try:
        assert cm.root.hex() == pin["root"]
    finally:
        cm.free()
