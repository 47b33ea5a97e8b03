"""One-GPU emulation of ONE RANK OF EIGHT at the 2^24-gate shape through the real multi-GPU code path (lg_shard_*, world = 1):
m = 4097, k = 1024, rho = 8 gives the rank's element count (16 388 x 1 024 = 2 049 x 8 192) and exactly its column hash
(16 388 rows x 8 192 columns: one sequential BLAKE2s chain of 8 195 compressions per column), without NVLink traffic and
without the strided passes.  Compares the block pipeline modes: python scripts/overlap_probe2.py [sub_blocks]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ligero_b200 import Context
from ligero_b200 import parallel as par
from ligero_b200.synthetic import matrix_rows_torch

sub = int(sys.argv[1]) if len(sys.argv) > 1 else 1
m, k, rho = 4097, 1024, 8
ctx = Context(0)
res = {}
for mode in (0, 1, 2):
    sc = par.ShardedCommitter(ctx, m, k, rho, 0, 1, mode, sub)
    msg = matrix_rows_torch(7, sc.row_ids, k, "cuda")
    for _ in range(3):
        sc.commit_async(msg)
    root = sc.root()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(sc.stream)
    for _ in range(8):
        sc.commit_async(msg)
    e1.record(sc.stream)
    torch.cuda.synchronize()
    res[mode] = (e0.elapsed_time(e1) / 8, root.hex()[:12])
    sc.close()
print(f"groups while hashing = {os.environ.get('LG_SHARD_GROUPS', '2')}, sub_blocks = {sub}: " +
      ", ".join(f"{name} {res[md][0]:.2f} ms" for md, name in ((0, 'encode then hash'), (1, 'eager pipeline'), (2, 'deferred pipeline'))) +
      f"; roots equal: {len({r for _, r in res.values()}) == 1}", flush=True)
