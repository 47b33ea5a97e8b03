"""The product's ArithmeticCircuit (C++ host driver behind the C ABI; no GPU needed) against the reference's own circuit
tests (src/arithmetic_circuit/tests.rs) and, node by node, against the oracle's ArithmeticCircuit."""
import random

import pytest

import ligero_b200 as lb
from oracle import ligero_oracle as O

P = O.P


def test_add_constants():                         # tests.rs:104-111
    c = lb.ArithmeticCircuit()
    c.add(c.constant(1), c.constant(2))
    assert c.evaluate([]) == 3


def test_mul_constants():                         # tests.rs:113-121
    c = lb.ArithmeticCircuit()
    c.mul(c.constant(6), c.constant(2))
    assert c.evaluate([]) == 12


def test_pow_constants():                         # tests.rs:125-131
    c = lb.ArithmeticCircuit()
    c.pow(c.constant(2), 5)
    assert c.evaluate([]) == 32


def test_add_and_mul_variables():                 # tests.rs:133-155
    c = lb.ArithmeticCircuit()
    a, b = c.new_variables(2)
    c.add(a, b)
    assert c.evaluate([(a, 2), (b, 3)]) == 5
    c = lb.ArithmeticCircuit()
    a, b = c.new_variables(2)
    c.mul(a, b)
    assert c.evaluate([(a, 2), (b, 3)]) == 6


def test_pow_variable():                          # tests.rs:157-163
    c = lb.ArithmeticCircuit()
    a = c.new_variable()
    c.pow(a, 4)
    assert c.evaluate([(a, 2)]) == 16


def test_indicator():                             # tests.rs:165-174 (x^(r-1) = 1 for x != 0)
    c = lb.ArithmeticCircuit()
    a = c.new_variable()
    c.indicator(a)
    assert c.evaluate([(a, random.Random(4).randrange(1, P))]) == 1
    assert c.evaluate([(a, 0)]) == 0


def test_fibonacci():                             # tests.rs:244-271
    c = lb.ArithmeticCircuit()
    f0, f1 = c.new_variable(), c.new_variable()
    first, second = f0, f1
    for _ in range(3, 50):
        nxt = c.add(first, second)
        first, second = second, nxt
    assert c.evaluate_node([(f0, 1), (f1, 1)], 42 - 1) == 267914296
    assert c.evaluate_node([(f0, 5), (f1, 8)], 42 - 5) == 267914296


def test_fibonacci_with_const():                  # tests.rs:273-292
    c = lb.ArithmeticCircuit()
    f0 = c.constant(1)
    f1 = c.new_variable()
    first, second = f0, f1
    for _ in range(3, 50):
        nxt = c.add(first, second)
        first, second = second, nxt
    assert c.evaluate_node([(f1, 1)], 42 - 1) == 267914296


def test_lemniscate_and_determinant_circuits():   # tests.rs:52-101, 308-347
    c = lb.ArithmeticCircuit()
    one = c.constant(1)
    x, y = c.new_variable(), c.new_variable()
    a, b = c.constant(120), c.constant(80)
    x2, y2 = c.mul(x, x), c.mul(y, y)
    ax2, by2 = c.mul(a, x2), c.mul(b, y2)
    m = c.minus(ax2)
    s = c.add(x2, y2)
    t = c.add(by2, m)
    sq = c.mul(s, s)
    c.add_nodes([sq, t, one])
    assert c.evaluate([(x, 8), (y, 4)]) == 1
    oc = O.generate_lemniscate_circuit()
    assert c.num_nodes() == oc.num_nodes() and c.num_constants() == oc.num_constants()
    c = lb.ArithmeticCircuit()
    one = c.constant(1)
    v = c.new_variables(9)
    det = c.new_variable()
    s1 = c.add_nodes([c.mul_nodes([v[0], v[4], v[8]]), c.mul_nodes([v[1], v[5], v[6]]), c.mul_nodes([v[2], v[3], v[7]])])
    s2 = c.add_nodes([c.mul_nodes([v[2], v[4], v[6]]), c.mul_nodes([v[1], v[3], v[8]]), c.mul_nodes([v[0], v[5], v[7]])])
    c.add_nodes([s1, c.minus(s2), c.minus(det), one])
    vals = [2, 0, P - 1, 3, 5, 2, P - 4, 1, 4, 13]
    assert c.evaluate(list(zip(v + [det], vals))) == 1
    assert c.evaluate(list(zip(v + [det], vals[:-1] + [14]))) != 1


def test_constants_are_deduplicated_and_labels_unique():   # mod.rs:76-84, 92-100
    c = lb.ArithmeticCircuit()
    a, b = c.constant(7), c.constant(7)
    assert a == b and c.num_constants() == 1
    x = c.new_variable_with_label("x")
    assert c.get_variable("x") == x
    with pytest.raises(lb.LigeroB200Error):
        c.new_variable_with_label("x")
    with pytest.raises(lb.LigeroB200Error):
        c.get_variable("nope")
    with pytest.raises(lb.LigeroB200Error):
        c.add(x, 99)
    with pytest.raises(lb.LigeroB200Error):
        c.evaluate_node([(a, 1)], x)              # value supplied for a non-variable node
    with pytest.raises(lb.LigeroB200Error):
        c.evaluate_node([], c.add(x, a))          # uninitialised variable


def test_random_builder_programs_equal_the_oracle():
    """the same random sequence of builder calls on both sides gives the same node list and the same values"""
    rnd = random.Random(21)
    for trial in range(20):
        c, oc = lb.ArithmeticCircuit(), O.ArithmeticCircuit()
        nodes, vars_ = [], []
        for step in range(60):
            kind = rnd.choice(["const", "var", "add", "mul", "pow", "minus", "scalar"]) if nodes else "var"
            if kind == "const":
                v = rnd.choice([0, 1, 2, P - 1, rnd.randrange(P)])
                i, j = c.constant(v), oc.constant(v)
            elif kind == "var":
                i, j = c.new_variable(), oc.new_variable()
                vars_.append(i)
            elif kind in ("add", "mul"):
                a, b = rnd.choice(nodes), rnd.choice(nodes)
                i, j = (c.add(a, b), oc.add(a, b)) if kind == "add" else (c.mul(a, b), oc.mul(a, b))
            elif kind == "pow":
                a, e = rnd.choice(nodes), rnd.randrange(1, 12)
                i, j = c.pow(a, e), oc.pow(a, e)
            elif kind == "minus":
                a = rnd.choice(nodes)
                i, j = c.minus(a), oc.minus(a)
            else:
                ln = rnd.randrange(1, 4)
                l, r = [rnd.choice(nodes) for _ in range(ln)], [rnd.choice(nodes) for _ in range(ln)]
                i, j = c.scalar_product(l, r), oc.scalar_product(l, r)
            assert i == j
            nodes.append(i)
        assert c.num_nodes() == oc.num_nodes() and c.num_gates() == oc.num_gates()
        for idx in range(c.num_nodes()):
            nd, on = c.node(idx), oc.nodes[idx]
            if nd[0] == "const":
                assert on[0] == O.CONST and on[1] == nd[1]
            elif nd[0] == "var":
                assert on[0] == O.VAR
            else:
                assert on[0] == (O.ADD if nd[0] == "add" else O.MUL) and (on[1], on[2]) == (nd[1], nd[2])
        assign = [(v, rnd.randrange(P)) for v in vars_]
        outs = [nodes[-1], rnd.choice(nodes)]
        assert c.evaluate_multioutput(assign, outs) == oc.evaluate_multioutput(assign, outs)


def _same_nodes(c, oc):
    assert c.num_nodes() == oc.num_nodes() and c.num_constants() == oc.num_constants() and c.num_gates() == oc.num_gates()
    for idx in range(c.num_nodes()):
        nd, on = c.node(idx), oc.nodes[idx]
        if nd[0] == "const":
            assert on[0] == O.CONST and on[1] == nd[1], idx
        elif nd[0] == "var":
            assert on[0] == O.VAR, idx
        else:
            assert on[0] == (O.ADD if nd[0] == "add" else O.MUL) and (on[1], on[2]) == (nd[1], nd[2]), idx


def test_from_constraint_system_equals_the_oracle():
    """from_constraint_system (src/arithmetic_circuit/mod.rs:455-520) on the circom fixtures, the cube system of the
    reference's tests (15 nodes, tests.rs:239) and the hand-built repeated-squaring system: same circuit as the oracle's"""
    from tests.golden_util import load_r1cs
    from tests.util import repeated_squaring_r1cs
    systems = []
    for name in ("multiplication", "poseidon"):
        a, b, c, nw, wit = load_r1cs(name)
        systems.append((a, b, c, nw, wit))
    systems.append(([[(P - 1, 1)], [(1, 1)]], [[(1, 1)], [(1, 2)]], [[(P - 1, 2)], [(27, 0)]], 3, [1, 3, 9]))
    systems.append(repeated_squaring_r1cs(10, 3))
    for a, b, c, nw, wit in systems:
        circ, outs = lb.ArithmeticCircuit.from_constraint_system(a, b, c, nw)
        oc, oouts = O.ArithmeticCircuit.from_constraint_system(a, b, c, nw)
        assert outs == oouts
        _same_nodes(circ, oc)
        va = list(enumerate(wit))[1:]
        assert circ.evaluate_multioutput(va, outs) == oc.evaluate_multioutput(va, oouts) == [1] * len(outs)
    circ, _ = lb.ArithmeticCircuit.from_constraint_system(*systems[2][:4])
    assert circ.num_nodes() == 15
    with pytest.raises(lb.LigeroB200Error):                      # an empty row: add_nodes(empty).unwrap() panics (148-153)
        lb.ArithmeticCircuit.from_constraint_system([[]], [[(1, 1)]], [[(1, 1)]], 2)


def test_non_canonical_limbs_are_refused_at_the_abi():
    """raw Montgomery limbs >= r from a C caller (ark_bn254::Fr can never hold them) are LG_ERR_INVALID, not a
    silently different circuit: a non-canonical 1 would miss the constant table and shift the witness layout."""
    import numpy as np
    from ctypes import byref, c_size_t
    from ligero_b200 import _lib
    from ligero_b200.backend import _ptr, fr_to_limbs
    import ligero_b200 as lb
    lib = _lib.load()
    c = lb.ArithmeticCircuit()
    r_limbs = np.array([[0x43e1f593f0000001, 0x2833e84879b97091, 0xb85045b68181585d, 0x30644e72e131a029]], dtype=np.uint64)
    carry = 0                                                                    # the Montgomery 1 shifted by r: same residue
    vals = []
    for a, b in zip(fr_to_limbs([1])[0].tolist(), r_limbs[0].tolist()):
        t = a + b + carry
        vals.append(t & (2 ** 64 - 1))
        carry = t >> 64
    bad = np.array([vals], dtype=np.uint64)
    idx = c_size_t()
    assert lib.lg_circuit_constant(c.handle, _ptr(bad), byref(idx)) == 1          # LG_ERR_INVALID
    assert lib.lg_circuit_constant(c.handle, _ptr(fr_to_limbs([1])), byref(idx)) == 0
    x = c.new_variable()
    out = c.add(x, idx.value)
    var_idx = np.array([x], dtype=np.uint64)
    outs = np.array([out], dtype=np.uint64)
    res = np.zeros((1, 4), dtype=np.uint64)
    assert lib.lg_circuit_evaluate(c.handle, _ptr(var_idx), _ptr(bad), 1, _ptr(outs), 1, _ptr(res), None) == 1
    assert lib.lg_circuit_evaluate(c.handle, _ptr(var_idx), _ptr(fr_to_limbs([5])), 1, _ptr(outs), 1, _ptr(res), None) == 0
