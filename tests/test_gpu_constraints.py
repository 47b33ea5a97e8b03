"""the constraint matrix of LigeroCircuit::new built on the device (constraints.cu; reference generate_matrices,
src/ligero/mod.rs:296-433) against the host builder: identical CSC arrays, identical proofs; the oracle's A agrees entry by entry."""
import numpy as np
import pytest

import ligero_b200 as lb
from oracle import ligero_oracle as O
from tests.golden_util import load_r1cs
from tests.util import csc_right_block

pytestmark = pytest.mark.gpu
P = O.P


def both(gpu_ctx, circ, outs, monkeypatch):
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("LG_CSC_DEVICE", mode)
        lc = lb.LigeroCircuit(gpu_ctx, circ, outs)
        out[mode] = (lc, lc.constraints_csc())
    return out


def assert_same(a, b):
    for x, y, name in zip(a, b, ("col_ptr", "row_idx", "val_id", "const_table")):
        assert np.array_equal(x, y), name


@pytest.mark.parametrize("log_gates", [6, 12, 17])
def test_device_csc_equals_host_csc_synthetic(gpu_ctx, monkeypatch, log_gates):
    circ, out, assign = lb.ArithmeticCircuit.synthetic(1 << log_gates, 5 + log_gates)
    r = both(gpu_ctx, circ, [out], monkeypatch)
    assert_same(r["0"][1], r["1"][1])
    pa = r["0"][0].prove(assign, lb.PoseidonSponge.test_sponge()).to_bytes()
    pb = r["1"][0].prove(assign, lb.PoseidonSponge.test_sponge()).to_bytes()
    assert pa == pb


def test_device_csc_with_circuit_constants_and_many_outputs(gpu_ctx, monkeypatch):
    """R1CS-compiled circuits carry real constants (value ids >= 2) and one output per constraint: the poseidon fixture"""
    a, b, c, nw, wit = load_r1cs("poseidon")
    circ, outs = lb.ArithmeticCircuit.from_constraint_system(a, b, c, nw)
    r = both(gpu_ctx, circ, outs, monkeypatch)
    assert_same(r["0"][1], r["1"][1])
    assert len(r["1"][1][3]) > 0                      # constants beyond +-1 exist
    # against the oracle's A (mod.rs:296-433 restated in Python): same entries column by column, values compared as field elements
    oc, oouts = O.ArithmeticCircuit.from_constraint_system(a, b, c, nw)
    olc = O.LigeroCircuit(oc, oouts)
    mk = olc.m * olc.k
    ocp, ori, ovi, otab = csc_right_block(olc.a, mk)
    col_ptr, row_idx, val_id, table = r["1"][1]
    assert np.array_equal(col_ptr, ocp)

    def values(vid, tab):
        t = lb.limbs_to_fr(tab) if tab is not None and len(tab) else []
        return [1 if v == 0 else P - 1 if v == 1 else t[v - 2] for v in vid.tolist()]

    def by_column(cp, ri, vals):
        # the product keeps the reference's generation order inside a column (gate by gate), the oracle helper lists a
        # column by ascending row: compare the (row, value) sets
        return [sorted(zip(ri[cp[j]:cp[j + 1]].tolist(), vals[cp[j]:cp[j + 1]])) for j in range(len(cp) - 1)]
    assert by_column(col_ptr, row_idx, values(val_id, table)) == by_column(ocp, ori, values(ovi, otab))
