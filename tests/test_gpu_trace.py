"""Evaluation trace + witness layout on the device: lg_ligero_witness_matrix_dev produces the
matrix of the host evaluator (which tests/test_gpu_host_driver.py pins to the oracle) bit for bit, proofs built from
it are byte-identical, the reference's failure modes are kept, and the 2^20-gate synthetic configuration proves and
verifies end to end."""
import numpy as np
import pytest

import ligero_b200 as lb
from ligero_b200 import limbs_to_fr
from oracle import ligero_oracle as O
from tests.golden_util import load_r1cs
from tests.test_gpu_host_driver import mirror_circuit, osponge

pytestmark = pytest.mark.gpu
P = O.P


def dev_matrix(lc, assign, bump=True):
    t = lc.witness_matrix_device(assign, bump)
    return t.cpu().numpy().view(np.uint64)


@pytest.mark.parametrize("gates,seed", [(4, 1), (37, 2), (600, 3), (5000, 4), (70000, 5)])
def test_device_trace_equals_host_trace(gpu_ctx, gates, seed):
    circ, out, assign = lb.ArithmeticCircuit.synthetic(gates, seed)
    lc = lb.LigeroCircuit(gpu_ctx, circ, [out])
    assert lc.sol_len == gates + 4
    host = lc.witness_matrix(assign)
    assert np.array_equal(dev_matrix(lc, assign), host)
    info = lc.trace_info()
    assert info["gates"] == gates and 1 <= info["launches"] <= info["levels"]


def test_device_trace_matches_oracle_layout(gpu_ctx):
    """against the oracle's prove_inner layout directly (x, y, z at Mul gates, w = every non-constant node)"""
    oc, outs, assign = O.synthetic_circuit(300, seed=9)          # a chain: 300 levels, one gate each
    olc = O.LigeroCircuit(oc, outs)
    lc = lb.LigeroCircuit(gpu_ctx, mirror_circuit(oc), outs)
    got = limbs_to_fr(dev_matrix(lc, assign))
    want = [x for row in olc.witness_matrix(assign) for x in row]
    assert got == want


def test_device_trace_on_r1cs_circuits(gpu_ctx):
    """deep, thin circuits compiled from R1CS (from_constraint_system): many narrow levels in one CTA"""
    for name in ("multiplication", "poseidon"):
        a, b, c, nw, wit = load_r1cs(name)
        circ, outs = lb.ArithmeticCircuit.from_constraint_system(a, b, c, nw)
        lc = lb.LigeroCircuit(gpu_ctx, circ, outs)
        assign = list(enumerate(wit))[1:]
        assert np.array_equal(dev_matrix(lc, assign), lc.witness_matrix(assign))
        assert not lc.trace_info()["on_device"]                   # lg_prove keeps the host loop for such shapes
        lc.set_trace_mode(1)
        p_dev = lc.prove(assign, lb.PoseidonSponge.test_sponge()).to_bytes()
        lc.set_trace_mode(0)
        p_host = lc.prove(assign, lb.PoseidonSponge.test_sponge()).to_bytes()
        assert p_dev == p_host


def test_device_trace_failure_modes(gpu_ctx):
    c = lb.ArithmeticCircuit()
    x, y = c.new_variable(), c.new_variable()
    s = c.add(x, y)
    dangling = c.mul(x, x)                                        # feeds no output
    out = c.add(s, c.constant(5))
    lc = lb.LigeroCircuit(gpu_ctx, c, [out])
    with pytest.raises(lb.LigeroB200Error, match="truly depends"):
        lc.witness_matrix_device([(x, 1), (y, 2)])
    with pytest.raises(lb.LigeroB200Error, match="truly depends"):
        lc.witness_matrix([(x, 1), (y, 2)])
    c2 = lb.ArithmeticCircuit()
    x, y = c2.new_variable(), c2.new_variable()
    out = c2.add(c2.mul(x, y), c2.constant(3))
    lc2 = lb.LigeroCircuit(gpu_ctx, c2, [out])
    with pytest.raises(lb.LigeroB200Error, match="Uninitialised variable"):
        lc2.witness_matrix_device([(x, 1)])
    with pytest.raises(lb.LigeroB200Error, match="non-variable"):
        lc2.witness_matrix_device([(x, 1), (out, 2)])
    # a repeated index keeps its last value (arithmetic_circuit/mod.rs:345-347)
    a = dev_matrix(lc2, [(x, 7), (y, 2), (x, 4)])
    assert np.array_equal(a, lc2.witness_matrix([(x, 4), (y, 2)]))


def test_synthetic_2p20_gate_circuit_proves_and_verifies(gpu_ctx):
    """BASELINE config 3: synthetic 2^20-gate random Add/Mul circuit, prove on one B200 (device trace, no matrix
    upload), verify; a proof for a perturbed witness is rejected (src/ligero/tests.rs:160-170 at scale)."""
    gates = 1 << 20
    circ, out, assign = lb.ArithmeticCircuit.synthetic(gates, 2024)
    lc = lb.LigeroCircuit(gpu_ctx, circ, [out])
    assert (lc.sol_len, lc.m, lc.k, lc.n, lc.t) == (gates + 4, 1025, 2048, 16384, 156)
    info = lc.trace_info()
    assert info["on_device"] and info["levels"] < 200
    proof = lc.prove(assign, lb.PoseidonSponge.test_sponge())
    assert lc.verify(proof, lb.PoseidonSponge.test_sponge())
    blob = proof.to_bytes()
    lc.set_trace_mode(0)                                          # host evaluator + upload: same bytes
    assert lc.prove(assign, lb.PoseidonSponge.test_sponge()).to_bytes() == blob
    lc.set_trace_mode(-1)
    bad = [(assign[0][0], (assign[0][1] + 1) % P), assign[1]]
    assert not lc.verify(lc.prove(bad, lb.PoseidonSponge.test_sponge()), lb.PoseidonSponge.test_sponge())
