"""Multi-GPU encode + commit + prove, one process per GPU: a veneer over the C ABI's lg_shard_* (capi_shard.cu).
torch.distributed (NCCL) is used for rendezvous only -- exchanging the CUDA IPC handles once, aligning the ranks and
reducing the timing; the data path (codeword scatter, subtree roots, test evaluations, authentication paths) runs
inside the library over NVLink peer memory with device-side flags.

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


def local_row_ids(m: int, world: int, rank: int, sub_blocks: int = 1) -> List[int]:
    """global row indices (into the 4m-row matrix) owned by `rank`, in local order: for every block b in X, Y, Z, W and
    every sub-block u (m rows split into `sub_blocks` runs like block_slices) the rank's slice of that sub-block.
    With sub_blocks = 1 this is [X_g; Y_g; Z_g; W_g].  Mirrors lg_shard_layout."""
    out = []
    for b in range(4):
        for (s0, s1) in block_slices(m, sub_blocks):
            a, e = block_slices(s1 - s0, world)[rank]
            out.extend(b * m + s0 + i for i in range(a, e))
    return out


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
# device path: veneer over lg_shard_* (one lg_shard per rank)
# --------------------------------------------------------------------------------------------
def _make_shard(ctx, m: int, k: int, rho: int, rank: int, world: int, t_max: int, sub_blocks: int):
    """lg_shard_create + IPC handle exchange (the only use of torch.distributed besides barriers) + lg_shard_connect"""
    from ctypes import byref, c_void_p
    import numpy as np
    from .backend import check
    h = c_void_p()
    check(ctx.lib.lg_shard_create(ctx.handle, m, k, rho, rank, world, t_max, sub_blocks, byref(h)), ctx.handle, "lg_shard_create")
    mine = np.zeros(192, dtype=np.uint8)
    check(ctx.lib.lg_shard_handles(h, mine.ctypes.data), ctx.handle, "lg_shard_handles")
    if world > 1:
        handles = [None] * world
        dist.all_gather_object(handles, bytes(mine))
        blob = np.frombuffer(b"".join(handles), dtype=np.uint8).copy()
    else:
        blob = mine
    check(ctx.lib.lg_shard_connect(h, blob.ctypes.data), ctx.handle, "lg_shard_connect")
    if world > 1:
        dist.barrier()
    return h


class ShardedCommitter:
    """Row-sharded encode -> NVLink scatter fused into the encode kernels -> column-sharded hash + subtree -> root.
    `pipeline` (default on): the X, Y, Z, W row blocks (times `sub_blocks`) are encoded one after the other and the
    column owner hashes each block, on a higher-priority stream, while the next one is encoded and delivered."""

    def __init__(self, ctx, m: int, k: int, rho: int, rank: int, world: int, pipeline=None, sub_blocks: int = 0, t_max: int = 0):
        assert world & (world - 1) == 0 and k % world == 0 and (rho * k // world) >= 2
        from ctypes import byref, c_size_t
        from .backend import CommittedMatrix, check
        self.ctx, self.m, self.k, self.rho, self.rank, self.world = ctx, m, k, rho, rank, world
        self.handle = _make_shard(ctx, m, k, rho, rank, world, t_max, sub_blocks)     # sub_blocks = 0: the library's choice
        nr = c_size_t()
        check(ctx.lib.lg_shard_layout(self.handle, None, byref(nr), None, None), ctx.handle, "lg_shard_layout")
        self.sub_blocks = sub_blocks = nr.value // 4
        self.row_ids = local_row_ids(m, world, rank, sub_blocks)
        self.rows_g, self.kg = len(self.row_ids), k // world
        self.pipeline = pipeline
        if pipeline is not None:
            self.set_pipeline(pipeline)
        self.dev = torch.device("cuda", ctx.device)
        self.stream = torch.cuda.ExternalStream(ctx.stream, device=self.dev)
        self.mat_cols = CommittedMatrix(ctx, ctx.lib.lg_shard_matrix(self.handle), None)
        self.mat_cols.free = lambda: None            # borrowed: owned by the shard
        self.subtree_roots = None
        ctx._adopt(self)

    def free(self):
        self.close()

    def set_pipeline(self, enabled):
        """0 / False: off; 1 / True: eager (hash beside the next block's shared-memory kernel); 2: deferred (hash beside the
        next block's NVLink-bound last pass); None: the library default by world size.  Same setting on every rank."""
        from .backend import check
        v = -1 if enabled is None else int(enabled)
        check(self.ctx.lib.lg_shard_set_pipeline(self.handle, v), self.ctx.handle, "lg_shard_set_pipeline")

    def close(self):
        if getattr(self, "handle", None) and self.ctx.handle:
            self.ctx.sync()
            if self.world > 1:
                dist.barrier()                           # nobody unmaps a shard a peer may still be writing
            self.mat_cols.handle = None
            self.ctx.lib.lg_shard_free(self.handle)
            self.handle = None
            if self.world > 1:
                dist.barrier()

    def commit_async(self, msg_local) -> None:
        from .backend import _ptr, check
        check(self.ctx.lib.lg_shard_commit_async(self.handle, _ptr(msg_local) if self.rows_g else None), self.ctx.handle,
              "lg_shard_commit_async")

    def root(self) -> bytes:
        import numpy as np
        from .backend import _ptr, check
        root = np.zeros(32, dtype=np.uint8)
        sub = np.zeros((self.world, 32), dtype=np.uint8)
        check(self.ctx.lib.lg_shard_root(self.handle, _ptr(root), _ptr(sub)), self.ctx.handle, "lg_shard_root")
        self.subtree_roots = [bytes(sub[g]) for g in range(self.world)]
        return bytes(root)

    def commit(self, msg_local) -> bytes:
        self.commit_async(msg_local)
        return self.root()

    # ---------------------------------------------------------------------------------------
    @staticmethod
    def bench(ctx, R: int, k: int, rho: int, args, rank: int, world: int, seed: int = 20240, expected_root=None) -> dict:
        """bench.py's N > 1 path: strong scaling of one R x k encode+commit over `world` GPUs.  Every rank generates its
        own rows of the SAME seeded matrix (ligero_b200.synthetic), so the root is the one N = 1 and the CPU oracle get."""
        from .synthetic import matrix_rows_torch
        pipeline = {"0": 0, "1": 1, "2": 2}.get(os.environ.get("LG_MGPU_PIPELINE", ""), None)
        sub = int(os.environ.get("LG_MGPU_SUB", "0"))
        m = R // 4
        sc = ShardedCommitter(ctx, m, k, rho, rank, world, pipeline, sub)
        dev = sc.dev
        msg = matrix_rows_torch(seed, sc.row_ids, k, dev) if sc.rows_g else torch.zeros((k, 4), dtype=torch.int64, device=dev)
        torch.cuda.synchronize()             # the library's stream does not wait for torch's
        for _ in range(args.warmup):
            sc.commit_async(msg)
        root0 = sc.root()
        if expected_root is not None:
            assert root0.hex() == expected_root, f"root {root0.hex()} differs from the oracle-pinned {expected_root}"
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
        # one extra, untimed step with per-kernel events (phase marks on the context stream; the hash runs on its own)
        ctx.set_timing(True)
        ctx.phase_ms()
        sc.commit_async(msg)
        torch.cuda.synchronize()
        phases = ctx.phase_ms()
        kernel_ms = {p: v[0] for p, v in phases.items() if v[1]}
        kernel_launches = {p: v[1] for p, v in phases.items() if v[1]}
        ctx.set_timing(False)
        value = R * k / (ms_per_step * 1e-3)
        # end to end: pinned host shard -> device, root back on the host, every step (LG_BENCH_SKIP_E2E=1: tuning sweeps)
        if os.environ.get("LG_BENCH_SKIP_E2E") == "1":
            sc.close()
            return {"metric": "fr_elems_per_s_encode_commit", "value": value, "unit": "Fr elems/s", "n_gpus": world,
                    "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                    "scaling": "strong", "config": {"rows": R, "k": k, "n": rho * k, "rho_inv": rho, "seed": seed},
                    "e2e": None, "root": root0.hex(), "kernel_ms_rank0": kernel_ms, "gpu_launches": int(launches)}
        host = torch.empty_like(msg, device="cpu").pin_memory()
        host.copy_(msg)
        torch.cuda.synchronize()
        if pipeline is None and "LG_SHARD_PIPELINE" not in os.environ:
            sc.set_pipeline(int(os.environ.get("LG_MGPU_E2E_PIPELINE", "1")))   # host input: the upload paces the encoder
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
        rows_g, sub = sc.rows_g, sc.sub_blocks
        sc.close()
        par = (f"rows/{world} encode (4x{sub} row blocks) with the column exchange fused into the encode kernels (NVLink peer "
               f"stores) -> column-range/{world} BLAKE2s of each block behind the encoding of the next -> subtree -> root "
               f"exchange through peer mailboxes (device flags)")
        return {
            "metric": "fr_elems_per_s_encode_commit", "value": value, "unit": "Fr elems/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u32x8 Montgomery (BN254 Fr)", "data": "synthetic",
            "config": {"rows": R, "k": k, "n": rho * k, "rho_inv": rho, "parallelism": par,
                       "l2_policy": "inputs larger than L2", "seed": seed},
            "e2e": {"value": R * k / (e2e_ms * 1e-3), "unit": "Fr elems/s", "h2d_bytes_per_step": int(msg.numel() * 8) * world,
                    "d2h_bytes_per_step": 32 * world, "ms_per_step": e2e_ms,
                    "api": "lg_shard_commit(host pinned row shard) -> root on host, per rank"},
            "gpu_launches": int(launches), "root": root0.hex(),
            "root_check": ("equals the CPU-oracle root of the whole matrix (tests/golden/full_size_root.json)"
                           if expected_root is not None else "no pinned root for this shape"),
            "kernel_ms_rank0": kernel_ms, "kernel_launches_rank0": kernel_launches, "rows_per_rank": rows_g,
            "hash_pipeline": (int(pipeline) if pipeline is not None else
                              (int(os.environ["LG_SHARD_PIPELINE"]) if "LG_SHARD_PIPELINE" in os.environ else (1 if world >= 8 else 0))),
        }


class ShardedProver:
    """LigeroCircuit::prove over G GPUs -- lg_shard_prove / lg_shard_prove_matrix: the commitment is
    ShardedCommitter's; afterwards every rank holds ALL rows of its column range, so the three tests are column-parallel
    with no partial sums to reduce; the owner of an opened column serves it over NVLink together with the path inside
    its subtree.  Every rank runs the same Fiat-Shamir transcript, so all ranks end with the same proof, byte-identical
    to the single-GPU prover's (tests/test_gpu_multi.py) and through it to the oracle's."""

    def __init__(self, ctx, ligero, rank: int, world: int, sub_blocks: int = 0):
        self.ctx, self.L, self.rank, self.world = ctx, ligero, rank, world
        self.m, self.k, self.n, self.t = ligero.m, ligero.k, ligero.n, ligero.t
        self.rho = self.n // self.k
        self.committer = ShardedCommitter(ctx, self.m, self.k, self.rho, rank, world, None, sub_blocks, self.t)
        self.rows = 4 * self.m
        self.row_ids = torch.tensor(self.committer.row_ids, dtype=torch.int64, device=self.committer.dev)

    def close(self):
        self.committer.close()

    def local_rows(self, preenc_u):
        """this rank's rows (lg_shard_layout order) of a full 4m x k pre-encoding matrix (numpy or torch, [4mk, 4])"""
        full = torch.as_tensor(preenc_u.view("int64") if hasattr(preenc_u, "ctypes") else preenc_u)
        loc = full.view(self.rows, self.k, 4)[self.row_ids.to(full.device)].reshape(-1, 4).contiguous()
        loc = loc.to(self.committer.dev)
        torch.cuda.synchronize(self.committer.dev)   # torch's stream produced it; the library reads it on its own
        return loc

    def prove(self, var_assignment, sponge):
        import numpy as np
        from ctypes import byref, c_void_p
        from .api import LigeroProof
        from .backend import _ptr, check, fr_to_limbs
        idx = np.array([i for i, _ in var_assignment], dtype=np.uint64)
        vals = fr_to_limbs([v for _, v in var_assignment])
        h = c_void_p()
        check(self.ctx.lib.lg_shard_prove(self.committer.handle, self.L.handle, _ptr(idx), _ptr(vals), len(idx), 1, sponge.handle,
                                          byref(h)), self.ctx.handle, "lg_shard_prove")
        return LigeroProof(h)

    def prove_matrix(self, local_rows, sponge):
        from ctypes import byref, c_void_p
        from .api import LigeroProof
        from .backend import _ptr, check
        h = c_void_p()
        check(self.ctx.lib.lg_shard_prove_matrix(self.committer.handle, self.L.handle,
                                                 _ptr(local_rows) if self.committer.rows_g else None, sponge.handle, byref(h)),
              self.ctx.handle, "lg_shard_prove_matrix")
        return LigeroProof(h)
