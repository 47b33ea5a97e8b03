"""torchrun --nproc-per-node G scripts/mgpu_check.py : sharded commit root == single-GPU root (rank 0)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from ligero_b200 import Context
from ligero_b200 import parallel as par

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
ctx = Context(lr)
ok = True
for (m, k, rho) in [(3, 8, 8), (5, 64, 8), (86, 128, 8), (33, 2048, 8), (9, 8192, 4), (7, 16384, 4)]:
    if k // world < 2:
        continue                      # a column shard needs at least 2 message columns
    rng = np.random.default_rng(1234 + m)
    full = rng.integers(0, 2 ** 62, size=(4 * m * k, 4), dtype=np.uint64)
    full[:, 3] &= (1 << 60) - 1
    ids = par.local_row_ids(m, world, rank)
    local = np.ascontiguousarray(full.reshape(4 * m, k, 4)[ids]).reshape(-1, 4)
    dev = torch.from_numpy(local.view(np.int64)).cuda() if len(ids) else torch.zeros((k, 4), dtype=torch.int64, device="cuda")
    roots = {}
    for mode, pipe in (("fused", True), ("fused", False), ("nccl", False)):
        sc = par.ShardedCommitter(ctx, m, k, rho, rank, world, mode, pipe)
        r1 = sc.commit(dev)
        r2 = sc.commit(dev)
        roots[(mode, pipe)] = (r1, r2)
        sc.close()
    if rank == 0:
        cm = ctx.commit(full, 4 * m, k, rho)
        same = all(r == cm.root for pair in roots.values() for r in pair)
        print(f"m={m} k={k} rho={rho} world={world}: sharded roots (fused+pipelined hash, fused, nccl) {'==' if same else '!='} single-GPU root", flush=True)
        ok &= same
        cm.free()
    dist.barrier()
# host (pinned) row shards: the upload is tiled and overlapped with the encoding (lg_encode_sharded[_rows] on host memory)
for (m, k, rho) in [(1025, 2048, 8), (600, 8192, 8)]:
    rng = np.random.default_rng(99 + m)
    ids = par.local_row_ids(m, world, rank)
    local = rng.integers(0, 2 ** 62, size=(len(ids) * k, 4), dtype=np.uint64)
    local[:, 3] &= (1 << 60) - 1
    host = torch.from_numpy(local.view(np.int64)).pin_memory()
    dev = host.cuda()
    same = True
    for mode, pipe in (("fused", True), ("fused", False)):
        sc = par.ShardedCommitter(ctx, m, k, rho, rank, world, mode, pipe)
        r_dev = sc.commit(dev)
        r_host = sc.commit(host)
        r_host2 = sc.commit(host)
        same &= r_dev == r_host == r_host2
        sc.close()
    flag = torch.tensor([1 if same else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"m={m} k={k} world={world}: host-input (tiled upload) roots {'==' if int(flag.item()) else '!='} device-input roots", flush=True)
        ok &= bool(int(flag.item()))
if rank == 0:
    print("MGPU_OK" if ok else "MGPU_FAIL", flush=True)
dist.destroy_process_group()
