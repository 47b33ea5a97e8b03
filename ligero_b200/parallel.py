"""Multi-GPU encode + commit: one process per GPU, torch.distributed (NCCL over NVLink) for plumbing.

Sharding (north_star):
  1. rows: rank g owns the rows {b*m + i : b in X,Y,Z,W ; i in its slice of [0, m)} -- a 4*m_g-row
     "mini witness matrix" -- and Reed-Solomon-encodes them with no communication;
  2. ONE exchange (all-to-all over NVLink): rank h receives, for every row, the codeword columns whose
     message index c lies in its contiguous range [h*k/G, (h+1)*k/G), in all rho_inv coset planes;
     these are exactly the leaves [h*n/G, (h+1)*n/G) of the Merkle tree;
  3. columns: rank h hashes its column range and builds its Merkle SUBTREE (n/G leaves);
  4. NCCL all-gather of the G subtree roots (32 bytes each); the top log2(G) levels are computed
     redundantly on every rank, so the root equals the single-GPU root (the tree is positional).
The kernels are the single-GPU ones: a column shard is just a plane-layout matrix with k/G columns.

The pure index/host logic (partition, pack/unpack, top-of-tree) is backend-agnostic and is covered on
CPU by world_size-2 gloo tests (tests/test_parallel_gloo.py).
"""
from __future__ import annotations

import hashlib
import math
import os
import time
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


# --------------------------------------------------------------------------------------------
# host logic (no device required)
# --------------------------------------------------------------------------------------------
def block_slices(m: int, world: int) -> List[Tuple[int, int]]:
    """[i0, i1) of each rank inside one m-row block (first m % world ranks get one extra row)."""
    base, extra = divmod(m, world)
    out, start = [], 0
    for g in range(world):
        sz = base + (1 if g < extra else 0)
        out.append((start, start + sz))
        start += sz
    return out


def local_row_ids(m: int, world: int, rank: int) -> List[int]:
    """global row indices (into the 4m-row matrix) owned by `rank`, in local order [X_g; Y_g; Z_g; W_g]."""
    i0, i1 = block_slices(m, world)[rank]
    return [b * m + i for b in range(4) for i in range(i0, i1)]


def pack_for_exchange(u_rows: torch.Tensor, rho: int, rows_g: int, k: int, world: int) -> List[torch.Tensor]:
    """u_rows: [rho, rows_g, k, 4] plane layout of the local row shard -> per destination h the contiguous
    block [rho, rows_g, k/G, 4] of message-index range h."""
    kg = k // world
    v = u_rows.view(rho, rows_g, world, kg, 4)
    return [v[:, :, h].contiguous() for h in range(world)]


def unpack_after_exchange(recv: Sequence[torch.Tensor], out: torch.Tensor, m: int, world: int, rho: int, kg: int) -> None:
    """recv[g]: [rho, 4*m_g, kg, 4] from source rank g -> out: [rho, 4m, kg, 4] in GLOBAL row order."""
    o = out.view(rho, 4, m, kg, 4)
    for g, (i0, i1) in enumerate(block_slices(m, world)):
        mg = i1 - i0
        if mg == 0:
            continue
        o[:, :, i0:i1] = recv[g].view(rho, 4, mg, kg, 4)


def combine_subtree_roots(roots: Sequence[bytes]) -> bytes:
    """top log2(G) levels of the SHA-256 tree over the G subtree roots (inner-node format: H(L || R))."""
    level = list(roots)
    assert len(level) & (len(level) - 1) == 0
    while len(level) > 1:
        level = [hashlib.sha256(level[2 * i] + level[2 * i + 1]).digest() for i in range(len(level) // 2)]
    return level[0]


def top_auth_paths(roots: Sequence[bytes]) -> List[List[bytes]]:
    """For each of the G subtrees: the sibling digests on the way from its root to the tree root, listed from
    just below the root downwards (the order of ark_crypto_primitives::Path::auth_path)."""
    g = len(roots)
    assert g & (g - 1) == 0
    levels = [list(roots)]
    while len(levels[-1]) > 1:
        lv = levels[-1]
        levels.append([hashlib.sha256(lv[2 * i] + lv[2 * i + 1]).digest() for i in range(len(lv) // 2)])
    paths = []
    for h in range(g):
        sib, pos = [], h
        for lv in levels[:-1]:           # bottom (subtree roots) upwards
            sib.append(lv[pos ^ 1])
            pos >>= 1
        paths.append(sib[::-1])          # root side first
    return paths


def split_openings(idx: Sequence[int], n: int, world: int) -> List[Tuple[int, int]]:
    """(owner rank, leaf index inside the owner's subtree) of every opened column: rank h owns the contiguous
    leaves [h*n/G, (h+1)*n/G) (= message indices [h*k/G, (h+1)*k/G) in all coset planes)."""
    per = n // world
    return [(int(j) // per, int(j) % per) for j in idx]


def gather_concat(part: torch.Tensor) -> torch.Tensor:
    """concatenation over ranks (rank order) of equally shaped tensors: the all-gather of column slices"""
    world = dist.get_world_size()
    outs = [torch.empty_like(part) for _ in range(world)]
    dist.all_gather(outs, part.contiguous())
    return torch.cat(outs, dim=0)


def exchange(send: List[torch.Tensor], recv: List[torch.Tensor]) -> None:
    """all-to-all of per-destination blocks.  NCCL: one grouped all_to_all over NVLink; gloo (CPU tests):
    emulated with broadcasts, because gloo has no all_to_all."""
    if dist.get_backend() == "nccl":
        dist.all_to_all(recv, send)
        return
    world, rank = dist.get_world_size(), dist.get_rank()
    for src in range(world):
        for dst in range(world):
            if src == dst:
                if rank == src:
                    recv[src].copy_(send[dst])
                continue
            if rank == src:
                dist.send(send[dst], dst)
            elif rank == dst:
                dist.recv(recv[src], src)


# --------------------------------------------------------------------------------------------
# device path
# --------------------------------------------------------------------------------------------
def open_peer_shards(ctx, rows: int, kg: int, rho: int, rank: int, world: int):
    """One column-shard matrix (rows x kg, rho planes) per rank, each mapped into every peer with CUDA IPC.
    Returns (this rank's CommittedMatrix, ctypes array of the `world` shard base pointers, peer pointers to close)."""
    from ctypes import byref, c_void_p
    import numpy as np
    from .backend import CommittedMatrix, check
    h = c_void_p()
    check(ctx.lib.lg_matrix_create(ctx.handle, rows, kg, rho, byref(h)), ctx.handle, "lg_matrix_create")
    mat = CommittedMatrix(ctx, h, None)
    handle = np.zeros(64, dtype=np.uint8)
    check(ctx.lib.lg_ipc_export(mat.handle, handle.ctypes.data), ctx.handle, "lg_ipc_export")
    handles = [None] * world
    dist.all_gather_object(handles, bytes(handle))
    ptrs, peers = [], []
    for g in range(world):
        if g == rank:
            ptrs.append(int(ctx.lib.lg_matrix_u_dev(mat.handle)))
        else:
            p = c_void_p()
            hb = np.frombuffer(handles[g], dtype=np.uint8).copy()
            check(ctx.lib.lg_ipc_open(ctx.handle, hb.ctypes.data, byref(p)), ctx.handle, "lg_ipc_open")
            ptrs.append(int(p.value))
            peers.append(int(p.value))
    return mat, (c_void_p * world)(*ptrs), peers


class ShardedCommitter:
    """Row-sharded encode -> exchange -> column-sharded hash + subtree -> root all-gather.

    mode "fused" (default): the exchange is fused into the encode kernels -- the last NTT pass stores each
        finished element directly into the owning rank's column shard over NVLink peer memory (CUDA IPC
        mappings exchanged once at construction), so the transfer overlaps the butterflies and there is no
        pack / all-to-all / unpack.  The four row blocks X, Y, Z, W are encoded one after the other; after
        each, one tiny NCCL all-reduce tells every rank that the block has landed everywhere and the column
        owner hashes those rows on its second stream WHILE the next block is encoded and delivered: a column
        hash is a sequential chain per column (`pipeline=False` hashes after the last block instead; the
        default picks by world size from measurements, see __init__).
    mode "nccl": encode locally, then pack -> NCCL all_to_all -> unpack (the plain-library baseline, and the
        plumbing the gloo tests cover).
    """

    def __init__(self, ctx, m: int, k: int, rho: int, rank: int, world: int, mode: str = "fused", pipeline=None):
        assert world & (world - 1) == 0 and k % world == 0 and (rho * k // world) >= 2
        assert mode in ("fused", "nccl")
        self.ctx, self.m, self.k, self.rho, self.rank, self.world, self.mode = ctx, m, k, rho, rank, world, mode
        if pipeline is None:
            # measured on 8 x B200 (2^24-gate shape): the block pipeline wins at 2 GPUs (78.6 vs 81.9 ms) and loses
            # at 4 and 8 (44.3 vs 41.7, 31.6 vs 26.2 ms), where the hash is a pure latency chain that the encoder's
            # warps on the same SM slow down more than the overlap gains
            pipeline = world <= 2
        self.pipeline = bool(pipeline) and mode == "fused"
        self.slices = block_slices(m, world)
        i0, i1 = self.slices[rank]
        self.i0, self.m_g = i0, i1 - i0
        self.rows_g = 4 * self.m_g
        self.kg = k // world
        dev = torch.device("cuda", ctx.device)
        self.dev = dev
        self.stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
        self.roots = torch.empty((world, 32), dtype=torch.uint8, device=dev)
        self.my_root = torch.empty(32, dtype=torch.uint8, device=dev)
        self.flag = torch.zeros(1, dtype=torch.int32, device=dev)
        self._peer_ptrs = []
        if mode == "nccl":
            self.u_rows = torch.empty((rho, max(self.rows_g, 1), k, 4), dtype=torch.int64, device=dev)
            self.u_cols = torch.empty((rho, 4 * m, self.kg, 4), dtype=torch.int64, device=dev)
            self.recv = [torch.empty((rho, 4 * (b - a), self.kg, 4), dtype=torch.int64, device=dev) for a, b in self.slices]
            self.mat_rows = ctx.wrap(self.u_rows, max(self.rows_g, 1), k, rho) if self.rows_g else None
            self.mat_cols = ctx.wrap(self.u_cols, 4 * m, self.kg, rho)
        else:
            self.mat_cols, self.shard_ptrs, self._peer_ptrs = open_peer_shards(ctx, 4 * m, self.kg, rho, rank, world)
            # local intermediate of the coset planes, only for rows longer than one CTA tile (k > 1024)
            self.scratch = (torch.empty(((rho - 1) * max(self.rows_g, 1) * k, 4), dtype=torch.int64, device=dev)
                            if k > 1024 else None)
            dist.barrier()

    def close(self):
        if self._peer_ptrs:
            self.ctx.sync()
            dist.barrier()
            for p in self._peer_ptrs:
                self.ctx.lib.lg_ipc_close(self.ctx.handle, p)
            self._peer_ptrs = []
            dist.barrier()
        if getattr(self, "mat_cols", None) is not None:
            self.mat_cols.free()

    def commit_async(self, msg_local, marks=None) -> None:
        """Enqueue everything on the context stream; the root lands in self.roots (device).
        `marks`: optional list that receives (label, torch.cuda.Event) pairs for per-phase timing."""
        from .backend import _ptr, check

        def mark(label):
            if marks is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record(self.stream)
                marks.append((label, e))
        with torch.cuda.stream(self.stream):
            mark("start")
            if self.mode == "nccl":
                if self.mat_rows is not None:
                    self.mat_rows.encode(msg_local)
                mark("encode")
                send = pack_for_exchange(self.u_rows[:, : self.rows_g], self.rho, self.rows_g, self.k, self.world)
                mark("pack")
                exchange(send, self.recv)
                mark("exchange")
                unpack_after_exchange(self.recv, self.u_cols, self.m, self.world, self.rho, self.kg)
                mark("unpack")
            elif self.pipeline:
                lib, scratch = self.ctx.lib, (_ptr(self.scratch) if self.scratch is not None else None)
                base = int(_ptr(msg_local).value or 0)
                for b in range(4):                 # local rows are [X_g; Y_g; Z_g; W_g], 32 bytes per element
                    off = b * self.m_g * self.k * 32
                    check(lib.lg_encode_sharded_rows(self.ctx.handle, base + off if self.m_g else None, self.m_g,
                                                     b * self.m + self.i0, 4 * self.m, self.k, self.rho,
                                                     self.shard_ptrs, self.world, scratch, 1),
                          self.ctx.handle, "lg_encode_sharded_rows")
                    dist.all_reduce(self.flag)      # block b has landed on every rank ...
                    check(lib.lg_matrix_hash_rows(self.mat_cols.handle, b * self.m, (b + 1) * self.m),
                          self.ctx.handle, "lg_matrix_hash_rows")   # ... and is hashed behind the next block
                mark("encode+scatter over NVLink (column hashing of earlier blocks overlapped)")
                check(lib.lg_matrix_hash_finish(self.mat_cols.handle, None), self.ctx.handle, "lg_matrix_hash_finish")
                mark("hash tail+subtree")
            else:
                check(self.ctx.lib.lg_encode_sharded(self.ctx.handle, _ptr(msg_local), self.m_g, self.k, self.rho,
                                                     self.shard_ptrs, self.world, self.m, self.i0,
                                                     _ptr(self.scratch) if self.scratch is not None else None, 1),
                      self.ctx.handle, "lg_encode_sharded")
                mark("encode+scatter over NVLink")
                dist.all_reduce(self.flag)      # every rank's stores have landed before anyone hashes
                mark("rank barrier")
            if not self.pipeline:
                self.mat_cols.hash_async()
                mark("hash+subtree")
            # subtree root = node 0 of the local tree (device -> device, stays on the stream)
            self._copy_root()
            dist.all_gather_into_tensor(self.roots.view(-1), self.my_root)
            mark("root all-gather")

    def _copy_root(self):
        # subtree root = node 0 of the library-owned node array: a stream-ordered D2D copy through torch
        if getattr(self, "_root_view", None) is None:
            ptr = int(self.ctx.lib.lg_matrix_nodes_dev(self.mat_cols.handle) or 0)
            self._root_view = torch.as_tensor(_DevBytes(ptr, 32, self.ctx.device), device=self.my_root.device)
        self.my_root.copy_(self._root_view)

    def root(self) -> bytes:
        """Synchronise and fold the gathered subtree roots into the tree root (host, log2(G) hashes)."""
        self.ctx.sync()
        torch.cuda.current_stream().synchronize()
        r = self.roots.cpu().numpy()
        return combine_subtree_roots([bytes(r[g]) for g in range(self.world)])

    def commit(self, msg_local) -> bytes:
        self.commit_async(msg_local)
        return self.root()

    # ---------------------------------------------------------------------------------------
    @staticmethod
    def bench(ctx, R: int, k: int, rho: int, args, rank: int, world: int) -> dict:
        """bench.py's N > 1 path: strong scaling of one R x k encode+commit over `world` GPUs."""
        mode = os.environ.get("LG_MGPU_MODE", "fused")
        pipeline = {"0": False, "1": True}.get(os.environ.get("LG_MGPU_PIPELINE", ""), None)
        m = R // 4
        sc = ShardedCommitter(ctx, m, k, rho, rank, world, mode, pipeline)
        dev = torch.device("cuda", ctx.device)
        g = torch.Generator(device=dev)
        g.manual_seed(20240 + rank)
        msg = torch.randint(0, 2 ** 62, (max(sc.rows_g, 1) * k, 4), dtype=torch.int64, device=dev, generator=g)
        msg[:, 3] &= (1 << 60) - 1
        for _ in range(args.warmup):
            sc.commit_async(msg)
        root0 = sc.root()
        launches0 = ctx.launches
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(sc.stream)
        for _ in range(args.steps):
            sc.commit_async(msg)
        e1.record(sc.stream)
        torch.cuda.synchronize()
        dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        launches = ctx.launches - launches0
        assert sc.root() == root0
        ms_per_step = float(ms.item()) / args.steps
        marks = []
        ctx.set_timing(True)
        ctx.phase_ms()
        sc.commit_async(msg, marks)     # one extra, untimed step with per-phase and per-kernel events
        torch.cuda.synchronize()
        kernel_ms = {p: v[0] / max(1, v[1]) for p, v in ctx.phase_ms().items() if v[1]}
        ctx.set_timing(False)
        phase_ms = {marks[i][0]: marks[i - 1][1].elapsed_time(marks[i][1]) for i in range(1, len(marks))}
        value = R * k / (ms_per_step * 1e-3)
        # end to end: pinned host shard -> device, root back on the host, every step
        host = torch.empty_like(msg, device="cpu").pin_memory()
        host.copy_(msg)
        sc.commit(host)
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            r = sc.commit(host)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        assert r == root0
        e2e_ms = float(dt.item()) * 1e3 / args.steps
        sc.close()
        par = (f"rows/{world} encode with the column exchange fused into the last NTT pass (NVLink peer stores) -> "
               f"column-range/{world} hash + subtree -> NCCL root all-gather") if mode == "fused" else \
              f"rows/{world} encode -> NCCL all-to-all -> column-range/{world} hash + subtree -> root all-gather"
        return {
            "metric": "fr_elems_per_s_encode_commit", "value": value, "unit": "Fr elems/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u32x8 Montgomery (BN254 Fr)", "data": "synthetic",
            "config": {"rows": R, "k": k, "n": rho * k, "rho_inv": rho, "parallelism": par,
                       "l2_policy": "inputs larger than L2"},
            "e2e": {"value": R * k / (e2e_ms * 1e-3), "unit": "Fr elems/s", "h2d_bytes_per_step": int(msg.numel() * 8) * world,
                    "d2h_bytes_per_step": 32 * world, "ms_per_step": e2e_ms,
                    "api": "ShardedCommitter.commit(host pinned row shard) -> root on host, per rank"},
            "gpu_launches": int(launches), "root": root0.hex(), "phase_ms_rank0": phase_ms,
            "kernel_ms_per_launch_rank0": kernel_ms, "rows_per_rank": sc.rows_g,
            "hash_pipeline": bool(sc.pipeline),
        }


class ShardedProver:
    """LigeroCircuit::prove over G GPUs: the commitment is ShardedCommitter's; afterwards every rank
    holds ALL rows of its column range, so the three tests are column-parallel with no partial sums to reduce:

      Test-Interleaved   r^T U_pre on the rank's columns                      -> all-gather of k/G values
      Test-Linear        r_a = r^T A (replicated, O(nnz)); its 4m rows are extended to the odd points of the 2k
                         domain row-sharded, with the same fused NVLink scatter as the witness (rho = 2, Montgomery
                         form kept); q on the rank's 2k/G points               -> all-gather, one inverse NTT
      Test-Quadratic     q on the rank's 2k/G points                           -> all-gather, one inverse NTT
      openings           the owner of a column returns it with the path inside its subtree; the top log2(G)
                         siblings come from the all-gathered subtree roots

    Every rank runs the same Fiat-Shamir transcript on the gathered values, so all ranks end with the same proof,
    byte-identical to the single-GPU (and the reference's) proof for the same witness and sponge."""

    def __init__(self, ctx, ligero, rank: int, world: int):
        from ctypes import byref, c_void_p
        from .backend import check
        self.ctx, self.L, self.rank, self.world = ctx, ligero, rank, world
        self.m, self.k, self.n, self.t = ligero.m, ligero.k, ligero.n, ligero.t
        self.rho = self.n // self.k
        self.committer = ShardedCommitter(ctx, self.m, self.k, self.rho, rank, world, "fused")
        c = self.committer
        self.rows, self.kg = 4 * self.m, c.kg
        self.rhat, self.rhat_ptrs, self._rhat_peers = open_peer_shards(ctx, self.rows, self.kg, 2, rank, world)
        self.rhat_scratch = (torch.empty((max(c.rows_g, 1) * self.k, 4), dtype=torch.int64, device=c.dev)
                             if self.k > 1024 else None)
        self.row_ids = torch.tensor(local_row_ids(self.m, world, rank), dtype=torch.int64, device=c.dev)
        a = c_void_p()
        check(ctx.lib.lg_ligero_constraints(ligero.handle, byref(a)), ctx.handle, "lg_ligero_constraints")
        self.a = a
        dist.barrier()

    def close(self):
        self.ctx.sync()
        dist.barrier()
        for p in self._rhat_peers:
            self.ctx.lib.lg_ipc_close(self.ctx.handle, p)
        self._rhat_peers = []
        dist.barrier()
        self.rhat.free()
        self.committer.close()

    def local_rows(self, preenc_u):
        """this rank's rows [X_g; Y_g; Z_g; W_g] of a full 4m x k pre-encoding matrix (numpy or torch, [4mk, 4])"""
        full = torch.as_tensor(preenc_u.view("int64") if hasattr(preenc_u, "ctypes") else preenc_u)
        loc = full.view(self.rows, self.k, 4)[self.row_ids.to(full.device)].reshape(-1, 4).contiguous()
        return loc.to(self.committer.dev)

    def prove(self, var_assignment, sponge):
        """LigeroCircuit::prove over G GPUs from the variable assignment alone: every rank runs the (cheap, replicated)
        evaluation trace on its own GPU (lg_ligero_witness_matrix_dev), keeps its rows of [X;Y;Z;W] and goes on with
        prove_matrix -- no host trace, no matrix upload."""
        pre = self.L.witness_matrix_device(var_assignment)
        return self.prove_matrix(self.local_rows(pre), sponge)

    # -- collectives over host-sized results -------------------------------------------------------
    def _gather(self, part_np):
        import numpy as np
        t = torch.from_numpy(np.ascontiguousarray(part_np).view(np.int64)).to(self.committer.dev)
        return gather_concat(t).cpu().numpy().view(np.uint64)

    def _open(self, sponge, subtree_roots):
        import numpy as np
        dev, world, rank = self.committer.dev, self.world, self.rank
        idx = self.ctx.expand_indices(sponge.squeeze_bytes(32), self.n, self.t)
        where = split_openings(idx, self.n, world)
        mine = [q for q, (h, _) in enumerate(where) if h == rank]
        depth_local = (self.n // world).bit_length() - 2
        cols_all = torch.zeros((self.t, self.rows, 4), dtype=torch.int64, device=dev)
        sib_all = torch.zeros((self.t, 4), dtype=torch.int64, device=dev)
        auth_all = torch.zeros((self.t, max(depth_local, 0) * 4 + 1), dtype=torch.int64, device=dev)
        if mine:
            cols, sib, auth = self.committer.mat_cols.open(np.array([where[q][1] for q in mine], dtype=np.uint64))
            sel = torch.tensor(mine, dtype=torch.int64, device=dev)
            cols_all[sel] = torch.from_numpy(cols.view(np.int64)).to(dev)
            sib_all[sel] = torch.from_numpy(sib.view(np.int64).reshape(len(mine), 4)).to(dev)
            if depth_local > 0:
                auth_all[sel, : depth_local * 4] = torch.from_numpy(auth.view(np.int64).reshape(len(mine), depth_local * 4)).to(dev)
        # exactly one rank contributes a non-zero row per opened column: the sum is a gather
        for tns in (cols_all, sib_all, auth_all):
            dist.all_reduce(tns)
        top = top_auth_paths(subtree_roots)
        cols_np = cols_all.cpu().numpy().view(np.uint64)
        sib_np = sib_all.cpu().numpy().view(np.uint8).reshape(self.t, 32)
        auth_loc = auth_all[:, : max(depth_local, 0) * 4].cpu().numpy().view(np.uint8).reshape(self.t, max(depth_local, 0), 32)
        depth = self.n.bit_length() - 2
        auth_np = np.zeros((self.t, depth, 32), dtype=np.uint8)
        ntop = depth - max(depth_local, 0)
        for q, (h, _) in enumerate(where):
            for d in range(ntop):
                auth_np[q, d] = np.frombuffer(top[h][d], dtype=np.uint8)
            auth_np[q, ntop:] = auth_loc[q]
        return cols_np, np.ascontiguousarray(idx, dtype=np.uint64), sib_np, auth_np

    def prove_matrix(self, local_rows, sponge):
        """local_rows: this rank's rows of the pre-encoding matrix (see local_rows()); returns a LigeroProof."""
        import numpy as np
        from ctypes import byref, c_size_t, c_void_p
        from .api import LigeroProof
        from .backend import _ptr, check
        ctx, lib, c = self.ctx, self.ctx.lib, self.committer
        root = c.commit(local_rows)                                            # mod.rs:521-551
        r_host = c.roots.cpu().numpy()
        subtree_roots = [bytes(r_host[g]) for g in range(self.world)]
        sponge.absorb_bytes(root)                                              # 560
        # Test-Interleaved (646-669)
        r = ctx.expand_fr(sponge.squeeze_bytes(32), self.rows)
        lc = self._gather(c.mat_cols.row_combine(r))                           # k x 4
        sponge.absorb_fr(lc)
        opened = [self._open(sponge, subtree_roots)]
        # Test-Linear-Constraints (712-747)
        seed = np.frombuffer(sponge.squeeze_bytes(32), dtype=np.uint8).copy()
        r_a = torch.empty((self.rows * self.k, 4), dtype=torch.int64, device=c.dev)
        check(lib.lg_linear_ra(ctx.handle, self.a, _ptr(seed), _ptr(r_a)), ctx.handle, "lg_linear_ra")
        loc = r_a.view(self.rows, self.k, 4)[self.row_ids].reshape(-1, 4).contiguous() if c.m_g else None
        with torch.cuda.stream(c.stream):
            check(lib.lg_encode_sharded(ctx.handle, _ptr(loc) if loc is not None else None, c.m_g, self.k, 2, self.rhat_ptrs,
                                        self.world, self.m, c.i0,
                                        _ptr(self.rhat_scratch) if self.rhat_scratch is not None else None, 0),
                  ctx.handle, "lg_encode_sharded(r_a)")
            dist.all_reduce(c.flag)        # every rank's rows of r-hat have landed
        ctx.sync()
        base = int(lib.lg_matrix_u_dev(self.rhat.handle))
        ev = np.empty((2 * self.kg, 4), dtype=np.uint64)
        check(lib.lg_linear_evals(c.mat_cols.handle, c_void_p(base), c_void_p(base + self.rows * self.kg * 32), _ptr(ev)),
              ctx.handle, "lg_linear_evals")
        lin = self._poly(self._gather(ev))
        sponge.absorb_fr(lin)
        opened.append(self._open(sponge, subtree_roots))
        # Test-Quadratic-Constraints (832-859)
        rq = ctx.expand_fr(sponge.squeeze_bytes(32), self.m)
        check(lib.lg_quadratic_evals(c.mat_cols.handle, _ptr(rq), _ptr(ev)), ctx.handle, "lg_quadratic_evals")
        quad = self._poly(self._gather(ev))
        sponge.absorb_fr(quad)
        opened.append(self._open(sponge, subtree_roots))
        # LigeroProof
        depth = self.n.bit_length() - 2
        arr = lambda j: (c_void_p * 3)(*[_ptr(o[j]).value for o in opened])
        rootb = np.frombuffer(root, dtype=np.uint8).copy()
        h = c_void_p()
        check(lib.lg_proof_assemble(_ptr(rootb), _ptr(lc), self.k, _ptr(lin), len(lin), _ptr(quad), len(quad), self.t, self.rows,
                                    depth, arr(0), arr(1), arr(2), arr(3), byref(h)), ctx.handle, "lg_proof_assemble")
        return LigeroProof(h)

    def _poly(self, evals_np):
        import numpy as np
        from ctypes import byref, c_size_t
        from .backend import _ptr, check
        out = np.empty((2 * self.k, 4), dtype=np.uint64)
        n = c_size_t()
        check(self.ctx.lib.lg_poly_from_evals(self.ctx.handle, _ptr(np.ascontiguousarray(evals_np)), 2 * self.k, _ptr(out), byref(n)),
              self.ctx.handle, "lg_poly_from_evals")
        return np.ascontiguousarray(out[: n.value])


class _DevBytes:
    """__cuda_array_interface__ view of raw device memory (library-owned), for torch interop."""

    def __init__(self, ptr: int, nbytes: int, device: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}
