"""BASELINE config 5 / SURVEY 8(d): isolated RS-encode + Merkle-commit microbenchmark, rho = 1/4, matrix entries from the
repo's own ChaCha -> Fr expander seeded with sha256("ligero-b200/microbench/R/n").  Prints ms per encode+commit (matrix
resident in HBM, best of 3 after one warm-up), message elements per second and the first bytes of the root.

    python scripts/microbench.py [RxN ...]        default: a diagonal and two corners of the 2^8..2^14 x 2^10..2^16 grid"""
import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from ligero_b200 import Context
from ligero_b200.backend import _ptr, check

ctx = Context(0)
st = torch.cuda.ExternalStream(ctx.stream)
points = [(1 << 8, 1 << 10), (1 << 10, 1 << 12), (1 << 12, 1 << 14), (1 << 14, 1 << 16), (1 << 14, 1 << 10), (1 << 8, 1 << 16)]
if len(sys.argv) > 1:
    points = [tuple(int(x) for x in a.split("x")) for a in sys.argv[1:]]
for R, n in points:
    k = n // 4
    seed = np.frombuffer(hashlib.sha256(f"ligero-b200/microbench/{R}/{n}".encode()).digest(), dtype=np.uint8).copy()
    msg = torch.empty((R * k, 4), dtype=torch.int64, device="cuda")
    check(ctx.lib.lg_expand_fr(ctx.handle, _ptr(seed), R * k, _ptr(msg)), ctx.handle, "lg_expand_fr")
    cm = ctx.commit(msg, R, k, 4)
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        check(ctx.lib.lg_recommit(cm.handle, _ptr(msg), None), ctx.handle, "lg_recommit")
        e1.record(st)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    assert cm.hash() == cm.root
    print(f"R=2^{R.bit_length() - 1} n=2^{n.bit_length() - 1} (k={k}, rho_inv=4): {best:.3f} ms per encode+commit, "
          f"{R * k / best / 1e3:.1f} M Fr/s, root {cm.root[:8].hex()}", flush=True)
    cm.free()
    del msg
    torch.cuda.empty_cache()
