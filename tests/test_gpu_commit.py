"""GPU parity: encode + column hash + Merkle root through the C ABI vs the Python oracle."""
import random

import numpy as np
import pytest

from oracle import ligero_oracle as O
from ligero_b200 import fr_to_limbs, limbs_to_fr

pytestmark = pytest.mark.gpu
P = O.P


def oracle_encode_commit(rows, k, rho):
    dk, dn = O.Domain(k), O.Domain(rho * k)
    u = [dn.fft(dk.ifft(r)) for r in rows]
    n = rho * k
    leaves = [O.column_hash([u[i][j] for i in range(len(rows))]) for j in range(n)]
    tree = O.MerkleTree(leaves)
    return u, leaves, tree


def rand_matrix(rnd, R, k, zero_rows=()):
    m = [[rnd.randrange(P) for _ in range(k)] for _ in range(R)]
    for z in zero_rows:
        m[z] = [0] * k
    return m


@pytest.mark.parametrize("R,k,rho", [
    (1, 2, 8), (3, 4, 8), (16, 4, 8), (5, 8, 4), (7, 16, 8), (4, 32, 2), (9, 64, 8), (6, 128, 8),
    (3, 256, 4), (5, 512, 8), (4, 1024, 8), (3, 2048, 8), (2, 4096, 4), (2, 8192, 8), (1, 16384, 4),
])
def test_encode_commit_matches_oracle(gpu_ctx, R, k, rho):
    rnd = random.Random(R * 1000003 + k * 17 + rho)
    zero_rows = (1,) if R > 2 else ()
    msg = rand_matrix(rnd, R, k, zero_rows)
    u, leaves, tree = oracle_encode_commit(msg, k, rho)
    flat = fr_to_limbs([x for row in msg for x in row])
    cm = gpu_ctx.commit(flat, R, k, rho)
    try:
        got = cm.read_rows(0, R)
        for i in range(R):
            assert limbs_to_fr(got[i]) == u[i], f"row {i} of U differs"
        got_leaves = cm.read_leaves()
        for j in range(rho * k):
            assert bytes(got_leaves[j]) == leaves[j], f"leaf {j} differs"
        nodes = cm.read_nodes()
        assert [bytes(x) for x in nodes] == tree.nodes
        assert cm.root == tree.root()
    finally:
        cm.free()


def test_commit_device_input_and_recommit(gpu_ctx):
    import torch
    rnd = random.Random(7)
    R, k, rho = 8, 64, 8
    msg = rand_matrix(rnd, R, k)
    _, _, tree = oracle_encode_commit(msg, k, rho)
    flat = fr_to_limbs([x for row in msg for x in row])
    dev = torch.from_numpy(flat.view(np.int64)).cuda()
    cm = gpu_ctx.commit(dev, R, k, rho)
    assert cm.root == tree.root()
    msg2 = rand_matrix(rnd, R, k)
    _, _, tree2 = oracle_encode_commit(msg2, k, rho)
    assert cm.recommit(fr_to_limbs([x for row in msg2 for x in row])) == tree2.root()
    cm.free()


@pytest.mark.parametrize("R,k", [(2200, 4096), (9000, 1024)])
def test_host_input_pipeline_equals_device_input(gpu_ctx, R, k):
    """Host matrices >= 64 MiB are uploaded in row tiles overlapped with the encoding (two tiles here); the
    result must equal the device-resident path (itself checked against the oracle above) and the C oracle."""
    import torch
    from oracle import cref
    rng = np.random.default_rng(R + k)
    a = rng.integers(0, 2 ** 62, size=(R * k, 4), dtype=np.uint64)
    a[:, 3] &= (1 << 60) - 1
    a.reshape(R, k, 4)[R // 3] = 0                      # an all-zero row inside a tile
    dev = torch.from_numpy(a.view(np.int64)).cuda()
    cm_dev = gpu_ctx.commit(dev, R, k, 8)
    cm_host = gpu_ctx.commit(a, R, k, 8)
    try:
        assert cm_host.root == cm_dev.root
        assert np.array_equal(cm_host.read_rows(R - 2, 2), cm_dev.read_rows(R - 2, 2))
        assert np.array_equal(cm_host.read_leaves(), cm_dev.read_leaves())
        pinned = torch.from_numpy(a.view(np.int64)).pin_memory()
        assert cm_host.recommit(pinned) == cm_dev.root
        if k == 4096:
            sub = 64                                    # C oracle on a row prefix: codeword rows must agree
            ref = cref.commit(a[: sub * k], sub, k, 8, want_u=True)
            assert np.array_equal(cm_host.read_rows(0, 2), ref["u"][:2])
    finally:
        cm_dev.free()
        cm_host.free()


def test_format_switches(gpu_ctx):
    rnd = random.Random(9)
    R, k, rho = 5, 8, 8
    msg = rand_matrix(rnd, R, k)
    dk, dn = O.Domain(k), O.Domain(rho * k)
    u = [dn.fft(dk.ifft(r)) for r in msg]
    flat = fr_to_limbs([x for row in msg for x in row])
    for col_p in (True, False):
        for leaf_p in (True, False):
            fmt = O.Formats(col_len_prefix=col_p, leaf_len_prefix=leaf_p)
            leaves = [O.column_hash([u[i][j] for i in range(R)], O.FR, fmt) for j in range(rho * k)]
            root = O.MerkleTree(leaves, fmt).root()
            gpu_ctx.set_formats(col_p, leaf_p)
            try:
                cm = gpu_ctx.commit(flat, R, k, rho)
                assert cm.root == root
                cm.free()
            finally:
                gpu_ctx.set_formats(True, True)


def test_intt_matches_oracle(gpu_ctx):
    rnd = random.Random(11)
    for rows, size in [(3, 2), (2, 16), (1, 1024), (2, 4096), (1, 32768)]:
        ev = [[rnd.randrange(P) for _ in range(size)] for _ in range(rows)]
        d = O.Domain(size)
        want = [d.ifft(r) for r in ev]
        got = gpu_ctx.intt(fr_to_limbs([x for r in ev for x in r]), rows, size)
        got = limbs_to_fr(got)
        for i in range(rows):
            assert got[i * size:(i + 1) * size] == want[i]


def test_invalid_arguments(gpu_ctx):
    from ligero_b200 import LigeroB200Error
    flat = fr_to_limbs([1, 2, 3])
    with pytest.raises(LigeroB200Error):
        gpu_ctx.commit(flat, 1, 3, 8)      # k not a power of two
    with pytest.raises(LigeroB200Error):
        gpu_ctx.commit(flat, 0, 2, 8)      # no rows
    with pytest.raises(LigeroB200Error):
        gpu_ctx.commit(flat, 1, 2, 3)      # rho_inv not a power of two


@pytest.mark.parametrize("R,k,rho,cuts", [
    (7, 8, 8, [3, 4]),            # odd and even boundaries, single-row tiles
    (12, 64, 8, [1, 2, 5, 11]),
    (9, 128, 4, [4, 5]),
    (16, 2048, 8, [7]),
    (5, 16, 8, [2]),
])
def test_tile_wise_column_hashing_matches_oracle(gpu_ctx, R, k, rho, cuts):
    """lg_matrix_hash_rows / lg_matrix_hash_finish: a column's BLAKE2s stream carried across row tiles with
    arbitrary (odd or even) boundaries gives the leaves and root of the one-shot hash and of the oracle."""
    rnd = random.Random(R * 31 + k)
    msg = rand_matrix(rnd, R, k, (1,) if R > 2 else ())
    _, leaves, tree = oracle_encode_commit(msg, k, rho)
    flat = fr_to_limbs([x for row in msg for x in row])
    cm = gpu_ctx.encode(flat, R, k, rho)
    try:
        for quad_max in (0, 1 << 30):        # thread-per-column tiles, four-lanes-per-column tiles
            gpu_ctx.set_hash_quad_max(quad_max)
            bounds = [0] + list(cuts) + [R]
            for a, b in zip(bounds[:-1], bounds[1:]):
                cm.hash_rows(a, b)
            root = cm.hash_finish()
            assert root == tree.root(), f"quad_max={quad_max}"
            got_leaves = cm.read_leaves()
            assert [bytes(x) for x in got_leaves] == leaves, f"quad_max={quad_max}"
            assert cm.hash() == tree.root()      # the one-shot path on the same matrix
        # the carried state has one format: tiles may alternate between the two kernels
        bounds = [0] + list(cuts) + [R]
        for t, (a, b) in enumerate(zip(bounds[:-1], bounds[1:])):
            gpu_ctx.set_hash_quad_max((1 << 30) if t % 2 == 0 else 0)
            cm.hash_rows(a, b)
        assert cm.hash_finish() == tree.root()
    finally:
        gpu_ctx.set_hash_quad_max(8192)
        cm.free()


@pytest.mark.parametrize("prefix", [(1, 1), (0, 0)])
def test_tile_wise_hashing_without_length_prefixes(gpu_ctx, prefix):
    """the format switches change the block alignment: tiles must still agree with one shot."""
    gpu_ctx.set_formats(*prefix)
    try:
        rnd = random.Random(99)
        R, k, rho = 11, 32, 8
        flat = fr_to_limbs([rnd.randrange(P) for _ in range(R * k)])
        cm = gpu_ctx.encode(flat, R, k, rho)
        gpu_ctx.set_hash_quad_max(0)
        one = cm.hash()
        for quad_max in (0, 1 << 30):
            gpu_ctx.set_hash_quad_max(quad_max)
            for a, b in ((0, 3), (3, 4), (4, 10), (10, 11)):
                cm.hash_rows(a, b)
            assert cm.hash_finish() == one
            assert cm.hash() == one
        cm.free()
    finally:
        gpu_ctx.set_hash_quad_max(8192)
        gpu_ctx.set_formats(1, 1)


@pytest.mark.parametrize("R,k,rho", [(1, 8, 8), (2, 8, 4), (3, 16, 8), (8, 8, 8), (13, 64, 8), (24, 32, 2), (5, 1024, 8)])
@pytest.mark.parametrize("prefix", [True, False])
def test_both_column_hash_kernels_match_oracle(gpu_ctx, R, k, rho, prefix):
    """the thread-per-column kernel and the four-lanes-per-column kernel (few columns: one rank's range on 4-8 GPUs)
    give the oracle's leaves for odd and even row counts, with and without the length prefix."""
    rnd = random.Random(R * 7919 + k + rho)
    msg = rand_matrix(rnd, R, k, (1,) if R > 2 else ())
    dk, dn = O.Domain(k), O.Domain(rho * k)
    u = [dn.fft(dk.ifft(r)) for r in msg]
    fmt = O.Formats(col_len_prefix=prefix, leaf_len_prefix=True)
    leaves = [O.column_hash([u[i][j] for i in range(R)], O.FR, fmt) for j in range(rho * k)]
    flat = fr_to_limbs([x for row in msg for x in row])
    gpu_ctx.set_formats(prefix, True)
    try:
        cm = gpu_ctx.encode(flat, R, k, rho)
        for quad_max in (0, 1 << 30):
            gpu_ctx.set_hash_quad_max(quad_max)
            cm.hash()
            got = cm.read_leaves()
            assert [bytes(x) for x in got] == leaves, f"quad_max={quad_max}"
        cm.free()
    finally:
        gpu_ctx.set_formats(True, True)
        gpu_ctx.set_hash_quad_max(8192)


def test_column_hash_kernels_agree_on_a_tall_matrix(gpu_ctx):
    """4 101 rows x 8 192 columns (the per-rank column range of the 2^20-gate matrix on 2 GPUs): same root either way,
    and the root of the C oracle."""
    from oracle import cref
    R, k, rho = 4101, 1024, 8
    a = np.random.default_rng(5).integers(0, 2 ** 62, size=(R * k, 4), dtype=np.uint64)
    a[:, 3] &= (1 << 60) - 1
    ref = cref.commit(a, R, k, rho)
    cm = gpu_ctx.encode(a, R, k, rho)
    try:
        roots = []
        for quad_max in (0, 1 << 30):
            gpu_ctx.set_hash_quad_max(quad_max)
            roots.append(cm.hash())
        assert roots[0] == roots[1] == ref["root"]
    finally:
        gpu_ctx.set_hash_quad_max(8192)
        cm.free()
