"""world_size-2 (and 4) gloo tests of the multi-GPU host logic on CPU: row partition, pack -> exchange ->
unpack index math, subtree roots and the top of the tree.  The per-shard arithmetic that the GPUs do is
stood in for by the oracle here (this is a test of the plumbing, which is shared with the NCCL path)."""
import os
import random
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ligero_b200 import fr_to_limbs, limbs_to_fr
from ligero_b200 import parallel as par
from oracle import ligero_oracle as O

P = O.P


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, m, k, rho, seed, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rnd = random.Random(seed)
        msg = [[rnd.randrange(P) for _ in range(k)] for _ in range(4 * m)]      # same matrix on every rank
        dk, dn = O.Domain(k), O.Domain(rho * k)
        ids = par.local_row_ids(m, world, rank)
        # stand-in for the GPU encode of the local rows, in the plane layout [rho][rows_g][k]
        u_local = [dn.fft(dk.ifft(msg[i])) for i in ids]
        planes = np.zeros((rho, max(len(ids), 1), k, 4), dtype=np.uint64)
        for li, row in enumerate(u_local):
            limbs = fr_to_limbs(row).reshape(k, rho, 4)         # row[rho*c + s] -> [c][s]
            planes[:, li] = limbs.transpose(1, 0, 2)
        u_rows = torch.from_numpy(planes.view(np.int64))
        kg = k // world
        send = par.pack_for_exchange(u_rows[:, : len(ids)], rho, len(ids), k, world)
        recv = [torch.empty((rho, 4 * (b - a), kg, 4), dtype=torch.int64) for a, b in par.block_slices(m, world)]
        par.exchange(send, recv)
        u_cols = torch.empty((rho, 4 * m, kg, 4), dtype=torch.int64)
        par.unpack_after_exchange(recv, u_cols, m, world, rho, kg)
        cols = u_cols.numpy().view(np.uint64)                    # [rho][4m][kg][4]
        # column shard -> leaves of the contiguous range [rank*n/G, (rank+1)*n/G)
        leaves = []
        for c in range(kg):
            for s in range(rho):
                col = limbs_to_fr(cols[s, :, c])
                leaves.append(O.column_hash(col))
        sub = O.MerkleTree(leaves).root()
        gathered = [None] * world
        dist.all_gather_object(gathered, sub)
        root = par.combine_subtree_roots(gathered)
        # ground truth: the unsharded oracle
        u = [dn.fft(dk.ifft(r)) for r in msg]
        want = O.MerkleTree([O.column_hash([u[i][j] for i in range(4 * m)]) for j in range(rho * k)]).root()
        q.put((rank, root == want))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,m,k,rho", [(2, 3, 8, 8), (2, 2, 4, 4), (4, 5, 16, 8)])
def test_sharded_commit_plumbing_matches_unsharded_root(world, m, k, rho):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, m, k, rho, 11, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in results), results


def test_partition_covers_every_row_once():
    for m in (1, 2, 5, 86, 4097):
        for world in (1, 2, 4, 8):
            ids = sorted(i for r in range(world) for i in par.local_row_ids(m, world, r))
            assert ids == list(range(4 * m))
            sizes = [b - a for a, b in par.block_slices(m, world)]
            assert max(sizes) - min(sizes) <= 1


def test_sub_block_runs_cover_every_row_once_and_keep_block_order():
    """lg_shard_layout's mirror (local_row_ids with sub-blocks): the 4 * sub pipeline steps of all ranks together cover the
    rows of one block of global rows each, in row order -- what lets the column owner hash step j as soon as every rank has
    delivered its j-th run (src/ligero/mod.rs:536-542: one BLAKE2s stream per column, rows in order)."""
    for m in (1, 3, 86, 513, 4097):
        for world in (1, 2, 8):
            for sub in (1, 2, 3, 4):
                if sub > m:
                    continue
                per_rank = [par.local_row_ids(m, world, r, sub) for r in range(world)]
                assert sorted(i for ids in per_rank for i in ids) == list(range(4 * m))
                # step boundaries: block b, sub-block u covers [b*m + s0, b*m + s1)
                bounds = [(b * m + s0, b * m + s1) for b in range(4) for (s0, s1) in par.block_slices(m, sub)]
                pos = [0] * world
                for (lo, hi) in bounds:
                    got = []
                    for r in range(world):
                        a, e = par.block_slices(hi - lo, world)[r]
                        run = per_rank[r][pos[r]:pos[r] + (e - a)]
                        pos[r] += e - a
                        assert run == list(range(lo + a, lo + e))       # a run of consecutive global rows
                        got += run
                    assert got == list(range(lo, hi))
                assert all(pos[r] == len(per_rank[r]) for r in range(world))


def test_top_of_tree():
    import hashlib
    roots = [bytes([i]) * 32 for i in range(4)]
    l0 = hashlib.sha256(roots[0] + roots[1]).digest()
    l1 = hashlib.sha256(roots[2] + roots[3]).digest()
    assert par.combine_subtree_roots(roots) == hashlib.sha256(l0 + l1).digest()
    assert par.combine_subtree_roots(roots[:1]) == roots[0]


# ---- host logic of the multi-GPU prover (ShardedProver): column ownership, path assembly, slice gathers ----------
def test_sharded_openings_paths_match_unsharded_tree():
    """Paths assembled from (path inside the owner's subtree) + (siblings over the gathered subtree roots) are the
    oracle's MerkleTree paths of the whole tree, for every leaf."""
    rnd = random.Random(5)
    for world, n in ((2, 8), (4, 16), (8, 64), (2, 4), (4, 8)):
        leaves = [bytes(rnd.randrange(256) for _ in range(32)) for _ in range(n)]
        whole = O.MerkleTree(leaves)
        per = n // world
        subs = [O.MerkleTree(leaves[h * per:(h + 1) * per]) for h in range(world)]
        top = par.top_auth_paths([s.root() for s in subs])
        assert par.combine_subtree_roots([s.root() for s in subs]) == whole.root()
        where = par.split_openings(list(range(n)), n, world)
        for j in range(n):
            h, jl = where[j]
            assert (h, jl) == (j // per, j % per)
            local = subs[h].generate_proof(jl)
            want = whole.generate_proof(j)
            assert local.leaf_sibling_hash == want.leaf_sibling_hash
            assert top[h] + local.auth_path == want.auth_path


def _gather_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        part = torch.arange(6, dtype=torch.int64).view(3, 2) + 100 * rank      # this rank's column slice
        full = par.gather_concat(part)
        want = torch.cat([torch.arange(6, dtype=torch.int64).view(3, 2) + 100 * r for r in range(world)])
        # the "sum as gather" used for the opened columns: exactly one rank contributes each row
        rows = torch.zeros((2 * world, 4), dtype=torch.int64)
        rows[2 * rank: 2 * rank + 2] = torch.tensor([[-1, 2 ** 62, rank, 7]] * 2, dtype=torch.int64)
        dist.all_reduce(rows)
        ok_rows = all(int(rows[2 * r, 2]) == r and int(rows[2 * r, 0]) == -1 for r in range(world))
        q.put((rank, bool(torch.equal(full, want)) and ok_rows))
    finally:
        dist.destroy_process_group()


def test_gather_of_column_slices_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in results), results
