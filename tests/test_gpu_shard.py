"""The multi-GPU layer of the C ABI (lg_shard_*: capi_shard.cu) on ONE GPU, world = 1: the whole machinery -- runs of
rows, mailbox collectives with device-side flags, the block pipeline on the second stream, sharded openings -- is the
same code that runs on 2-8 GPUs (tests/test_gpu_multi.py), with the rank talking to itself.  Results must equal the
single-GPU path, which the other test files pin to the oracle."""
import numpy as np
import pytest

import ligero_b200 as lb
from ligero_b200 import parallel as par
from oracle import cref

pytestmark = pytest.mark.gpu


def rand_matrix(rows, k, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 2 ** 62, size=(rows * k, 4), dtype=np.uint64)
    a[:, 3] &= (1 << 60) - 1
    return a


@pytest.mark.parametrize("m,k,rho", [(3, 8, 8), (5, 64, 8), (86, 128, 8), (33, 2048, 8), (9, 8192, 4), (2, 4096, 2)])
@pytest.mark.parametrize("pipeline,sub", [(1, 1), (0, 1), (1, 2), (2, 1), (2, 3)])
def test_shard_commit_world1_equals_oracle_root(gpu_ctx, m, k, rho, pipeline, sub):
    import torch
    full = rand_matrix(4 * m, k, 100 + m)
    want = cref.commit(full, 4 * m, k, rho)["root"]
    sc = par.ShardedCommitter(gpu_ctx, m, k, rho, 0, 1, pipeline, sub)
    try:
        ids = par.local_row_ids(m, 1, 0, sub)
        assert sorted(ids) == list(range(4 * m))
        local = np.ascontiguousarray(full.reshape(4 * m, k, 4)[ids]).reshape(-1, 4)
        dev = torch.from_numpy(local.view(np.int64)).cuda()
        assert sc.commit(dev) == want
        assert sc.commit(dev) == want                 # steady state: epochs advance, buffers are reused
        assert sc.commit(local) == want               # host input
        for _ in range(3):
            sc.commit_async(dev)                      # back-to-back asynchronous steps, one read-back
        assert sc.root() == want
    finally:
        sc.close()


@pytest.mark.parametrize("pipeline", [0, 1])
def test_two_stream_encoding_world1(gpu_ctx, pipeline, monkeypatch):
    """odd runs on a second stream (the multi-GPU default from 4 GPUs): same root, flags raised out of order are fine"""
    import torch
    monkeypatch.setenv("LG_SHARD_TWO_STREAM", "1")
    for (m, k, rho, sub) in [(33, 2048, 8, 1), (9, 8192, 4, 2), (40, 4096, 8, 3)]:
        full = rand_matrix(4 * m, k, 7 + m)
        want = cref.commit(full, 4 * m, k, rho)["root"]
        sc = par.ShardedCommitter(gpu_ctx, m, k, rho, 0, 1, pipeline, sub)
        try:
            local = np.ascontiguousarray(full.reshape(4 * m, k, 4)[sc.row_ids]).reshape(-1, 4)
            dev = torch.from_numpy(local.view(np.int64)).cuda()
            for _ in range(4):
                sc.commit_async(dev)
            assert sc.root() == want
            assert sc.commit(local) == want           # host input falls back to one stream
        finally:
            sc.close()


def test_shard_layout_matches_python_mirror(gpu_ctx):
    from ctypes import byref, c_size_t
    for (m, world, sub) in [(7, 1, 1), (7, 1, 3), (10, 1, 4)]:
        sc = par.ShardedCommitter(gpu_ctx, m, 16, 8, 0, world, None, sub)
        try:
            rl, nr = c_size_t(), c_size_t()
            base = np.zeros(4 * sub, dtype=np.uint64)
            cnt = np.zeros(4 * sub, dtype=np.uint64)
            assert gpu_ctx.lib.lg_shard_layout(sc.handle, byref(rl), byref(nr), base.ctypes.data, cnt.ctypes.data) == 0
            ids = [int(b) + i for b, c in zip(base[: nr.value], cnt[: nr.value]) for i in range(int(c))]
            assert ids == par.local_row_ids(m, world, 0, sub) and rl.value == len(ids)
        finally:
            sc.close()


@pytest.mark.parametrize("log_gates,sub", [(6, 1), (10, 1), (12, 2)])
def test_shard_prove_world1_equals_single_gpu_proof(gpu_ctx, log_gates, sub):
    circ, out, assign = lb.ArithmeticCircuit.synthetic(1 << log_gates, 7 + log_gates)
    L = lb.LigeroCircuit(gpu_ctx, circ, [out])
    ref = L.prove(assign, lb.PoseidonSponge.test_sponge()).to_bytes()
    sp = par.ShardedProver(gpu_ctx, L, 0, 1, sub)
    try:
        pre = L.witness_matrix(assign)
        s1, s2 = lb.PoseidonSponge.test_sponge(), lb.PoseidonSponge.test_sponge()
        p1 = sp.prove_matrix(sp.local_rows(pre), s1)
        assert p1.to_bytes() == ref
        p2 = sp.prove(assign, s2)                     # replicated device trace + local-row gather
        assert p2.to_bytes() == ref
        assert s1.squeeze_bytes(32) == s2.squeeze_bytes(32)        # the sponge is advanced exactly as by lg_prove
        assert L.verify(p2, lb.PoseidonSponge.test_sponge())
    finally:
        sp.close()


def test_mgpu_single_process_one_device(gpu_ctx):
    """lg_mgpu_* with one device: the one-call entry point a Rust LigeroCircuit::prove binds (host threads + peer access
    are exercised with more devices in tests/test_gpu_multi.py)."""
    from ctypes import byref, c_void_p
    from ligero_b200.backend import _ptr, fr_to_limbs
    lib = gpu_ctx.lib
    g = c_void_p()
    dev = np.array([0], dtype=np.int32)
    assert lib.lg_mgpu_create(dev.ctypes.data, 1, byref(g)) == 0
    try:
        m, k, rho = 9, 256, 8
        full = rand_matrix(4 * m, k, 5)
        root = np.zeros(32, dtype=np.uint8)
        assert lib.lg_mgpu_commit(g, _ptr(full), 4 * m, k, rho, _ptr(root)) == 0, lib.lg_mgpu_last_error(g)
        assert bytes(root) == cref.commit(full, 4 * m, k, rho)["root"]
        circ, out, assign = lb.ArithmeticCircuit.synthetic(1 << 10, 3)
        ref = lb.LigeroCircuit(gpu_ctx, circ, [out]).prove(assign, lb.PoseidonSponge.test_sponge()).to_bytes()
        ml = c_void_p()
        outs = np.array([out], dtype=np.uint64)
        assert lib.lg_mgpu_ligero_new(g, circ.handle, _ptr(outs), 1, lb.DEFAULT_SECURITY_LEVEL, byref(ml)) == 0, lib.lg_mgpu_last_error(g)
        idx = np.array([i for i, _ in assign], dtype=np.uint64)
        vals = fr_to_limbs([v for _, v in assign])
        sponge = lb.PoseidonSponge.test_sponge()
        h = c_void_p()
        assert lib.lg_mgpu_prove(ml, _ptr(idx), _ptr(vals), len(idx), 1, sponge.handle, byref(h)) == 0, lib.lg_mgpu_last_error(g)
        assert lb.LigeroProof(h).to_bytes() == ref
        lib.lg_mgpu_ligero_free(ml)
    finally:
        lib.lg_mgpu_destroy(g)
