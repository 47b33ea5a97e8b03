"""Host-side Fiat-Shamir pieces of the product that need no GPU: index sampling (lg_expand_indices, src/utils.rs:31-55)
and the Poseidon duplex sponge (lg_sponge_*) against the oracle."""
import random

import numpy as np

import ligero_b200 as lb
from ligero_b200 import _lib
from ligero_b200.backend import _ptr
from oracle import ligero_oracle as O

P = O.P


def expand_indices(seed: bytes, n: int, t: int):
    lib = _lib.load()
    out = np.empty(t, dtype=np.uint64)
    s = np.frombuffer(seed, dtype=np.uint8).copy()
    assert lib.lg_expand_indices(_ptr(s), n, t, _ptr(out)) == 0
    return [int(x) for x in out]


def test_distinct_indices_match_the_oracle():
    rnd = random.Random(2)
    for n, t in [(32, 32), (1024, 156), (64, 40), (64, 32), (64, 33), (16, 3), (65536, 156), (2, 1), (4, 4), (256, 255),
                 (524288, 156), (3, 2), (1000, 600)]:
        seed = bytes(rnd.randrange(256) for _ in range(32))
        got = expand_indices(seed, n, t)
        assert got == O.get_distinct_indices_from_prng(n, t, seed), (n, t)
        assert got == sorted(set(got)) and len(got) == t and all(0 <= j < n for j in got)
    lib = _lib.load()
    out = np.empty(4, dtype=np.uint64)
    assert lib.lg_expand_indices(_ptr(np.zeros(32, dtype=np.uint8)), 3, 4, _ptr(out)) != 0      # t > n


def test_sponge_transcripts_match_the_oracle():
    """random interleavings of absorb(bytes) / absorb(field elements) / squeeze_bytes, including empty absorbs, rate
    boundaries and squeezes longer than one permutation (arkworks duplex rules)"""
    rnd = random.Random(6)
    for trial in range(25):
        a, b = lb.PoseidonSponge.test_sponge(), O.PoseidonSponge(O.test_sponge_config())
        for step in range(12):
            op = rnd.choice(["bytes", "field", "squeeze", "squeeze", "empty"])
            if op == "bytes":
                data = bytes(rnd.randrange(256) for _ in range(rnd.choice([0, 1, 23, 31, 32, 62, 100])))
                a.absorb_bytes(data)
                b.absorb_bytes(data)
            elif op == "field":
                v = [rnd.choice([0, 1, P - 1, rnd.randrange(P)]) for _ in range(rnd.choice([1, 2, 3, 5, 8]))]
                a.absorb_field_elements(v)
                b.absorb_field_elements(v)
            elif op == "empty":
                a.absorb_field_elements([])
                b.absorb_field_elements([])
            else:
                n = rnd.choice([1, 7, 31, 32, 33, 62, 63, 100])
                assert a.squeeze_bytes(n) == b.squeeze_bytes(n), (trial, step, n)
        c = a.clone()
        assert a.squeeze_bytes(32) == b.squeeze_bytes(32) == c.squeeze_bytes(32)


def test_sponge_fast_path_equals_generic_path():
    """The host sponge has a width-3 / alpha-17 fast path (lazy ADX products, host_field.h) next to the generic permutation;
    LG_SPONGE_GENERIC=1 forces the latter.  Same transcript on a long absorb (an odd count, so the duplex ends mid-rate) and
    interleaved squeezes; the default path is compared with the oracle above."""
    import os
    import subprocess
    import sys
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import ligero_b200 as lb\n"
        "s = lb.PoseidonSponge.test_sponge()\n"
        "s.absorb_bytes(bytes(range(32)))\n"
        "out = s.squeeze_bytes(32)\n"
        "s.absorb_field_elements([(i * i * 7919 + 3) %% lb.BN254_R for i in range(1001)])\n"
        "out += s.squeeze_bytes(64)\n"
        "s.absorb_field_elements([lb.BN254_R - 1, 0, 1])\n"
        "print((out + s.squeeze_bytes(32)).hex())\n" % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    outs = []
    for env in ({}, {"LG_SPONGE_GENERIC": "1"}):
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, **env), check=True)
        outs.append(r.stdout.strip())
    assert outs[0] == outs[1] and len(outs[0]) == 256
