"""lg_circuit_from_r1cs_bytes (host, no GPU): the iden3 .r1cs v1 container of SURVEY App. C parsed by the product gives the
circuit that from_constraint_system builds from the same matrices (the committed golden arrays), and the oracle's reader
agrees on the container written here; malformed images are refused."""
import os
import struct
import tempfile

import pytest

from ligero_b200 import ArithmeticCircuit, LigeroB200Error
from oracle import ligero_oracle as O
from tests.golden_util import load_r1cs

P = O.P


def write_r1cs(a, b, c, n_wires, prime=P, shuffle_terms=False, duplicate_term=False):
    """rows as [(coeff, wire)] -> iden3 .r1cs v1 bytes (header, constraints, empty wire map)"""
    n = len(a)
    hdr = struct.pack("<I", 32) + prime.to_bytes(32, "little") + struct.pack("<IIII", n_wires, 0, 0, n_wires - 1) + \
        struct.pack("<Q", n_wires) + struct.pack("<I", n)
    body = b""
    for r in range(n):
        for mat in (a, b, c):
            terms = list(mat[r])
            if shuffle_terms:
                terms = terms[::-1]
            if duplicate_term and terms:                       # the same wire twice: the reader must merge them
                coeff, wire = terms[0]
                terms = [((coeff - 5) % prime, wire)] + terms[1:] + [(5, wire)]
            body += struct.pack("<I", len(terms))
            for coeff, wire in terms:
                body += struct.pack("<I", wire) + coeff.to_bytes(32, "little")
    wmap = b"".join(struct.pack("<Q", i) for i in range(n_wires))
    out = b"r1cs" + struct.pack("<II", 1, 3)
    for typ, payload in ((2, body), (1, hdr), (3, wmap)):     # sections in any order
        out += struct.pack("<IQ", typ, len(payload)) + payload
    return out


def nodes_of(circ):
    return [circ.node(i) for i in range(circ.num_nodes())]


@pytest.mark.parametrize("name", ["multiplication", "poseidon"])
@pytest.mark.parametrize("variant", ["plain", "shuffled", "duplicated"])
def test_reader_equals_from_constraint_system(name, variant):
    a, b, c, nw, wit = load_r1cs(name)
    image = write_r1cs(a, b, c, nw, shuffle_terms=variant == "shuffled", duplicate_term=variant == "duplicated")
    circ, outs, n_wires = ArithmeticCircuit.from_r1cs_bytes(image)
    want, want_outs = ArithmeticCircuit.from_constraint_system(a, b, c, nw)
    assert n_wires == nw and outs == want_outs
    assert nodes_of(circ) == nodes_of(want)
    va = list(enumerate(wit))[1:]
    assert all(v == 1 for v in circ.evaluate_multioutput(va, outs))
    # the oracle's reader (validated on the reference's own .r1cs files when the goldens were made) reads the same image
    with tempfile.NamedTemporaryFile(suffix=".r1cs", delete=False) as f:
        f.write(image)
    try:
        oa, ob, oc, onw = O.read_r1cs(f.name)
        assert (oa, ob, oc, onw) == (a, b, c, nw)
        c2, outs2, _ = ArithmeticCircuit.from_r1cs_file(f.name)
        assert nodes_of(c2) == nodes_of(want) and outs2 == want_outs
    finally:
        os.unlink(f.name)


def test_malformed_images_are_refused():
    a, b, c, nw, _ = load_r1cs("multiplication")
    good = write_r1cs(a, b, c, nw)
    for bad in (b"", b"r1cz" + good[4:], good[:40], good[:-7], good[:4] + struct.pack("<I", 2) + good[8:],
                write_r1cs(a, b, c, nw, prime=P + 2)):
        with pytest.raises(LigeroB200Error):
            ArithmeticCircuit.from_r1cs_bytes(bad)
    # a wire id beyond nWires
    with pytest.raises(LigeroB200Error):
        ArithmeticCircuit.from_r1cs_bytes(write_r1cs([[(1, nw + 3)]], [[(1, 1)]], [[(1, 2)]], nw))


def test_header_counts_must_be_justified_by_the_file():
    """a crafted nWires (or nConstraints) that the file cannot back must be refused before it drives an
    allocation (the header field is a u32; 0xF0000000 labelled variables would exhaust memory)."""
    a, b, c, nw, _ = load_r1cs("multiplication")
    good = bytearray(write_r1cs(a, b, c, nw))
    hdr = bytes(good).index(struct.pack("<I", 32) + P.to_bytes(32, "little"))      # start of the header payload
    for field_off, value in ((4 + 32, 0xF0000000), (4 + 32, nw + 1), (4 + 32 + 16 + 8, 0xF0000000)):
        bad = bytearray(good)
        bad[hdr + field_off: hdr + field_off + 4] = struct.pack("<I", value)
        with pytest.raises(LigeroB200Error):
            ArithmeticCircuit.from_r1cs_bytes(bytes(bad))
    # the reference's own fixture (no tampering) still loads when present in this container
    ref = "/root/reference/circom/multiplication.r1cs"
    if os.path.exists(ref):
        data = bytearray(open(ref, "rb").read())
        ArithmeticCircuit.from_r1cs_bytes(bytes(data))
        pos = bytes(data).index(P.to_bytes(32, "little")) + 32
        data[pos: pos + 4] = struct.pack("<I", 0xF0000000)
        with pytest.raises(LigeroB200Error):
            ArithmeticCircuit.from_r1cs_bytes(bytes(data))
