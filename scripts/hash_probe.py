"""On-GPU probe: the two column-hash kernels (thread per column / four lanes per column) at the column counts one rank
sees on 8, 4, 2 and 1 GPUs of the 2^24-gate matrix (16 388 rows).  Prints ms per whole-matrix hash (+ tree)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ligero_b200 import Context
from ligero_b200.backend import check

ctx = Context(0)
st = torch.cuda.ExternalStream(ctx.stream)
R = int(sys.argv[1]) if len(sys.argv) > 1 else 16388
for k in [int(x) for x in os.environ.get('LG_PROBE_K', '1024,2048,4096,8192').split(',')]:
    # the hash only reads U: any bytes will do (plane layout, rho_inv = 8)
    u = torch.randint(0, 2 ** 62, (8 * R * k, 4), dtype=torch.int64, device="cuda")
    cm = ctx.wrap(u, R, k, 8)
    res = {}
    for name, quad_max in (("thread", 0), ("quad", 1 << 30), ("default", 8192)):
        ctx.set_hash_quad_max(quad_max)
        roots, best = [], 1e9
        for rep in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            check(ctx.lib.lg_matrix_hash(cm.handle, None), ctx.handle)
            e1.record(st)
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        res[name] = (best, cm.hash())
    same = res["thread"][1] == res["quad"][1] == res["default"][1]
    print(f"rows={R} columns={8 * k}: thread-per-column {res['thread'][0]:.3f} ms, four-lanes-per-column "
          f"{res['quad'][0]:.3f} ms, default selection "
          f"{res['default'][0]:.3f} ms, roots equal: {same}", flush=True)
    cm.free()
    del u
ctx.set_hash_quad_max(8192)
