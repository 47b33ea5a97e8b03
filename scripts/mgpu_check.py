"""torchrun --nproc-per-node G scripts/mgpu_check.py : sharded commit root (lg_shard_*, one process per GPU) == the CPU
oracle's root of the whole matrix, for device and host (pinned / pageable) row shards, with and without the block
pipeline and with sub-blocks.  Prints MGPU_OK / MGPU_FAIL on rank 0 (tests/test_gpu_multi.py greps it)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from ligero_b200 import Context
from ligero_b200 import parallel as par
from oracle import cref

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
ctx = Context(lr)
ok = True
for (m, k, rho) in [(3, 8, 8), (5, 64, 8), (86, 128, 8), (33, 2048, 8), (9, 8192, 4), (7, 16384, 4), (1025, 2048, 8)]:
    if k // world < 2:
        continue                      # a column shard needs at least 2 message columns
    rng = np.random.default_rng(1234 + m)
    full = rng.integers(0, 2 ** 62, size=(4 * m * k, 4), dtype=np.uint64)
    full[:, 3] &= (1 << 60) - 1
    want = cref.commit(full, 4 * m, k, rho)["root"] if rank == 0 else None
    got = []
    for pipe, sub in ((1, 1), (0, 1), (2, 1), (2, 3)):
        sc = par.ShardedCommitter(ctx, m, k, rho, rank, world, pipe, sub)
        local = np.ascontiguousarray(full.reshape(4 * m, k, 4)[sc.row_ids]).reshape(-1, 4) if sc.rows_g else np.zeros((k, 4), np.uint64)
        dev = torch.from_numpy(local.view(np.int64)).cuda()
        host = torch.from_numpy(local.view(np.int64)).pin_memory()
        got += [sc.commit(dev), sc.commit(dev), sc.commit(host), sc.commit(local)]
        for _ in range(3):
            sc.commit_async(dev)
        got.append(sc.root())
        sc.close()
    if rank == 0:
        same = all(r == want for r in got)
        print(f"m={m} k={k} rho={rho} world={world}: {len(got)} sharded roots (eager / plain / deferred / deferred with 3 sub-blocks; device, pinned "
              f"and pageable input) {'==' if same else '!='} oracle root", flush=True)
        ok &= same
    dist.barrier()
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MGPU_OK" if int(flag.item()) else "MGPU_FAIL", flush=True)
dist.destroy_process_group()
