"""One-GPU emulation of one rank's work at 8 GPUs (2^24-gate matrix): encoding its 4 x 513 rows while hashing the
16 388 x 8 192-column range block by block on the second stream, against doing the two one after the other.
(No NVLink traffic here: this isolates how much the two kernels slow each other down when they share the SMs.)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ligero_b200 import Context

ctx = Context(0)
m, k, kg = 4097, 8192, int(sys.argv[1]) if len(sys.argv) > 1 else 1024
mg = 513
g = torch.Generator(device="cuda"); g.manual_seed(3)
msgs = []
for b in range(4):
    t = torch.randint(0, 2 ** 62, (mg * k, 4), dtype=torch.int64, device="cuda", generator=g)
    t[:, 3] &= (1 << 60) - 1
    msgs.append(t)
enc = [ctx.encode(msgs[b], mg, k, 8) for b in range(4)]
u = torch.randint(0, 2 ** 62, (8 * 4 * m * kg, 4), dtype=torch.int64, device="cuda", generator=g)
col = ctx.wrap(u, 4 * m, kg, 8)

def serial():
    for b in range(4):
        enc[b].encode(msgs[b])
    return col.hash()

def pipelined():
    for b in range(4):
        enc[b].encode(msgs[b])
        col.hash_rows(b * m, (b + 1) * m)
    return col.hash_finish()

def encode_only():
    for b in range(4):
        enc[b].encode(msgs[b])
    ctx.sync()

for quad_max, name in ((0, "thread-per-column tiles"), (1 << 30, "four-lane tiles")):
    ctx.set_hash_quad_max(quad_max)
    res = {}
    for fn in (encode_only, serial, pipelined):
        best, out = 1e9, None
        for rep in range(4):
            ctx.sync(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = fn()
            ctx.sync()
            best = min(best, (time.perf_counter() - t0) * 1e3)
        res[fn.__name__] = (best, out)
    print(f"{8 * kg} columns, {name}: encode only {res['encode_only'][0]:.2f} ms, encode then hash {res['serial'][0]:.2f} ms, "
          f"hash pipelined behind the encode {res['pipelined'][0]:.2f} ms, roots equal: {res['serial'][1] == res['pipelined'][1]}",
          flush=True)
