"""Quick on-GPU probe: integer peaks + encode/hash timings at the two synthetic-circuit shapes."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ligero_b200 import Context
from ligero_b200.backend import _ptr, check

ctx = Context(0)
print("int peaks", ctx.int_peak(100.0), flush=True)
st = torch.cuda.ExternalStream(ctx.stream)
shapes = [(344, 128, 8), (4100, 2048, 8), (16388, 8192, 8)]
if len(sys.argv) > 1:
    shapes = [tuple(int(x) for x in a.split("x")) for a in sys.argv[1:]]
for (R, k, rho) in shapes:
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    msg = torch.randint(0, 2**62, (R * k, 4), dtype=torch.int64, device="cuda", generator=g)
    msg[:, 3] &= (1 << 60) - 1
    torch.cuda.synchronize()
    cm = ctx.commit(msg, R, k, rho)
    for rep in range(3):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record(st)
        check(ctx.lib.lg_recommit(cm.handle, _ptr(msg), None), ctx.handle)
        e[1].record(st)
        check(ctx.lib.lg_matrix_hash(cm.handle, None), ctx.handle)
        e[2].record(st)
        torch.cuda.synchronize()
        tot, h = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
        print(f"{R}x{k} rho={rho}: encode+commit {tot:.3f} ms (hash+tree {h:.3f} ms, encode {tot-h:.3f} ms) "
              f"-> {R*k/tot/1e3:.1f} M Fr/s", flush=True)
    cm.free()
    del msg
