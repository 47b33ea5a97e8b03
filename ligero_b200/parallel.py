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
            from ctypes import byref, c_void_p
            import numpy as np
            from .backend import CommittedMatrix, check
            h = c_void_p()
            check(ctx.lib.lg_matrix_create(ctx.handle, 4 * m, self.kg, rho, byref(h)), ctx.handle, "lg_matrix_create")
            self.mat_cols = CommittedMatrix(ctx, h, None)
            handle = np.zeros(64, dtype=np.uint8)
            check(ctx.lib.lg_ipc_export(self.mat_cols.handle, handle.ctypes.data), ctx.handle, "lg_ipc_export")
            handles = [None] * world
            dist.all_gather_object(handles, bytes(handle))
            ptrs = []
            for g in range(world):
                if g == rank:
                    ptrs.append(int(ctx.lib.lg_matrix_u_dev(self.mat_cols.handle)))
                else:
                    p = c_void_p()
                    hb = np.frombuffer(handles[g], dtype=np.uint8).copy()
                    check(ctx.lib.lg_ipc_open(ctx.handle, hb.ctypes.data, byref(p)), ctx.handle, "lg_ipc_open")
                    ptrs.append(int(p.value))
                    self._peer_ptrs.append(int(p.value))
            self.shard_ptrs = (c_void_p * world)(*ptrs)
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
                                                     self.shard_ptrs, self.world, scratch),
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
                                                     _ptr(self.scratch) if self.scratch is not None else None),
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


class _DevBytes:
    """__cuda_array_interface__ view of raw device memory (library-owned), for torch interop."""

    def __init__(self, ptr: int, nbytes: int, device: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}
