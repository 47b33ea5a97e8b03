"""Known-answer tests from public standards for the primitives the oracle restates (SURVEY 4 tier i)."""
import hashlib
import struct

from oracle import cref
from oracle import ligero_oracle as O


def test_bn254_constants():
    assert O.P == 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
    assert O.FR.mont_r == 0x0e0a77c19a07df2f666ea36f7879462e36fc76959f60cd29ac96341c4ffffffb
    root = O.FR.two_adic_root()
    assert root == 19103219067921713944291392827692070036145651957329286315305642004821462161904
    assert pow(root, 1 << 28, O.P) == 1 and pow(root, 1 << 27, O.P) != 1


def test_domain_relations():
    # small_domain.element(c) == large_domain.element(8c); intermediate.element(c) == large.element(4c)
    k = 16
    small, inter, large = O.Domain(k), O.Domain(2 * k), O.Domain(8 * k)
    for c in range(k):
        assert small.element(c) == large.element(8 * c)
    for c in range(2 * k):
        assert inter.element(c) == large.element(4 * c)
    v = list(range(1, k + 1))
    assert small.ifft(small.fft(v)) == v


def test_chacha20_block_zero_key():
    # draft-agl-tls-chacha20poly1305 / RFC 7539 keystream for the all-zero key, counter 0
    blk = struct.pack("<16I", *O.chacha_block([0] * 8, 0, 20))
    assert blk.hex().startswith("76b8e0ada0f13d90405d6ae55386bd28bdd219b8a08ded1aa836efcc8b770dc7"
                                "da41597c5157488d7724e03fb8d84a376a43b8f41518a11cc387b669b2ee6586")


def test_blake2s_rfc7693_abc():
    want = "508c5e8c327c14e2e1a72ba34eeb452f37458b209ed63a294d999b4c86675982"
    assert hashlib.blake2s(b"abc").hexdigest() == want
    assert cref.blake2s(b"abc").hex() == want


def test_sha256_fips_abc():
    want = "ba7816bf8f01cfea414140de5dae2223b00361a396177a9cb410ff61f20015ad"
    assert cref.sha256(b"abc").hex() == want
    two = b"abcdbcdecdefdefgefghfghighijhijkijkljklmklmnlmnomnopnopq"
    assert cref.sha256(two).hex() == "248d6a61d20638b8e5c026930c3e6039a33ce45964ff2167f6ecedd419db06c1"


def test_calculate_t_values():
    # SURVEY A.8 [DERIVED]: lambda=128, distance (n-k+1, n), n = 8k
    for n, want in [(32, 32), (128, 128), (512, 155), (2048, 156), (65536, 156), (1024, 156)]:
        k = n // 8
        assert O.calculate_t(128, (n - k + 1, n), n) == want


def test_gen_range_and_distinct_indices_properties():
    seed = bytes(range(32))
    idx = O.get_distinct_indices_from_prng(1024, 156, seed)
    assert len(idx) == 156 and idx == sorted(set(idx)) and all(0 <= i < 1024 for i in idx)
    assert O.get_distinct_indices_from_prng(32, 32, seed) == list(range(32))      # t == n: every column
    many = O.get_distinct_indices_from_prng(64, 40, seed)                          # complement branch
    assert len(many) == 40 and many == sorted(set(many))


def test_field_rand_is_montgomery_raw():
    seed = bytes([7] * 32)
    rng = O.ChaChaRng(seed, 20)
    limbs = [rng.next_u64() for _ in range(4)]
    limbs[3] &= (1 << 62) - 1
    raw = sum(l << (64 * i) for i, l in enumerate(limbs))
    first = O.get_field_elements_from_prng(1, seed)[0]
    if raw < O.P:
        assert first == O.FR.from_mont(raw)
