"""lg_circuit_synthetic (host side, no GPU): the seeded Add/Mul circuit of the synthetic configurations obeys SURVEY
8(d)'s rules, and the oracle accepts it (LigeroCircuit::new does not panic, prove -> verify)."""
import pytest

from ligero_b200 import ArithmeticCircuit
from oracle import ligero_oracle as O


def to_oracle(circ: ArithmeticCircuit):
    oc = O.ArithmeticCircuit()
    for i in range(circ.num_nodes()):
        nd = circ.node(i)
        if nd[0] == "const":
            assert oc.constant(nd[1]) == i
        elif nd[0] == "var":
            assert oc.new_variable() == i
        elif nd[0] == "add":
            assert oc.add(nd[1], nd[2]) == i
        else:
            assert oc.mul(nd[1], nd[2]) == i
    return oc


@pytest.mark.parametrize("gates,seed", [(4, 1), (5, 2), (37, 3), (1000, 5), (4099, 7)])
def test_synthetic_circuit_rules(gates, seed):
    circ, out, assign = ArithmeticCircuit.synthetic(gates, seed)
    n = circ.num_nodes()
    assert circ.num_gates() == gates
    assert circ.num_constants() == 2 and circ.num_variables() == 2
    assert n == gates + 4 and out == n - 1
    assert circ.evaluate_node(assign, out) == 1
    depth = [0] * n
    used = [False] * n
    consts = set()
    for i in range(n):
        nd = circ.node(i)
        if nd[0] == "const":
            consts.add(i)
        if nd[0] in ("add", "mul"):
            assert not (nd[1] in consts and nd[2] in consts)          # never two constant operands
            depth[i] = 1 + max(depth[nd[1]], depth[nd[2]])
            used[nd[1]] = used[nd[2]] = True
    assert circ.node(out)[0] == "add"
    # every variable and gate feeds the output (constants are always assigned: arithmetic_circuit/mod.rs:333-343)
    assert all(used[i] for i in range(n) if i != out and i not in consts)
    assert depth[out] <= 12 * max(4, gates).bit_length()              # O(log gates)
    # same seed, same circuit; another seed, another assignment
    c2, out2, assign2 = ArithmeticCircuit.synthetic(gates, seed)
    assert assign2 == assign and [c2.node(i) for i in range(n)] == [circ.node(i) for i in range(n)]
    assert ArithmeticCircuit.synthetic(gates, seed + 1)[2] != assign


def test_oracle_accepts_synthetic_circuit():
    circ, out, assign = ArithmeticCircuit.synthetic(300, 11)
    oc = to_oracle(circ)
    lc = O.LigeroCircuit(oc, [out])
    assert lc.sol_len == 300 + 4
    proof = lc.prove(assign, O.PoseidonSponge(O.test_sponge_config()))
    assert lc.verify(proof, O.PoseidonSponge(O.test_sponge_config()))
