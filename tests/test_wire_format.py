"""Proof wire format on the host (no GPU): lg_proof_deserialize / lg_proof_serialize of the product round-trip the bytes the
oracle's serializer (oracle/wire.py, arkworks-CanonicalSerialize-compatible layout) produces for real proofs; truncated or
non-canonical input is refused instead of crashing."""
import random

import pytest

from ligero_b200 import LigeroB200Error, LigeroProof
from oracle import ligero_oracle as O
from oracle import wire


def oracle_proof(seed=3, gates=40):
    circ, outs, assign = O.synthetic_circuit(gates, seed=seed)
    lc = O.LigeroCircuit(circ, outs)
    return lc, lc.prove(assign, O.PoseidonSponge(O.test_sponge_config()))


def test_round_trip_of_oracle_proofs():
    for seed, gates in ((3, 40), (5, 9), (8, 150)):
        lc, proof = oracle_proof(seed, gates)
        blob = wire.serialize_proof(proof)
        again = LigeroProof.from_bytes(blob).to_bytes()
        assert again == blob
        back = wire.deserialize_proof(again)
        assert lc.verify(back, O.PoseidonSponge(O.test_sponge_config()))


def test_malformed_bytes_are_refused():
    _, proof = oracle_proof()
    blob = wire.serialize_proof(proof)
    rnd = random.Random(1)
    for cut in [0, 1, 31, 32, 40, len(blob) // 2, len(blob) - 1]:
        with pytest.raises(LigeroB200Error):
            LigeroProof.from_bytes(blob[:cut])
    with pytest.raises(LigeroB200Error):
        LigeroProof.from_bytes(blob + b"\x00")                      # trailing bytes
    # layout: root = u64 length + 32 bytes; preenc_u_lc = u64 length at [40, 48), first element at [48, 80)
    # a field element >= r is not canonical
    bad = bytearray(blob)
    bad[48:80] = b"\xff" * 32
    with pytest.raises(LigeroB200Error):
        LigeroProof.from_bytes(bytes(bad))
    # an absurd vector length must not allocate or read out of bounds
    bad = bytearray(blob)
    bad[40:48] = (1 << 60).to_bytes(8, "little")
    with pytest.raises(LigeroB200Error):
        LigeroProof.from_bytes(bytes(bad))
    # random single-byte corruption never crashes: it either parses (a different proof) or is refused
    for _ in range(200):
        t = bytearray(blob)
        t[rnd.randrange(len(t))] ^= 1 << rnd.randrange(8)
        try:
            LigeroProof.from_bytes(bytes(t)).to_bytes()
        except LigeroB200Error:
            pass


def test_columns_of_unequal_length_round_trip_unchanged():
    """The product stores the t opened columns of one opening in a single flat buffer (the device-to-host copy lands in it);
    a deserialised proof whose columns differ in length cannot use that layout -- it must still serialise back byte for byte
    (and can never verify: lg_verify checks every column against 4m rows)."""
    lc, proof = oracle_proof(5, 30)
    cols = proof.interleaved.columns
    cols[1] = cols[1][:-2]                      # one shorter column
    cols[3] = cols[3] + [7, 8, 9]               # one longer column
    blob = wire.serialize_proof(proof)
    assert LigeroProof.from_bytes(blob).to_bytes() == blob
    # all columns shorter by the same amount: the flat layout again
    lc, proof = oracle_proof(5, 30)
    proof.linear.columns[:] = [c[:-1] for c in proof.linear.columns]
    blob = wire.serialize_proof(proof)
    assert LigeroProof.from_bytes(blob).to_bytes() == blob
    # no columns at all
    lc, proof = oracle_proof(5, 30)
    proof.quadratic.columns[:] = []
    proof.quadratic.paths[:] = []
    blob = wire.serialize_proof(proof)
    assert LigeroProof.from_bytes(blob).to_bytes() == blob
