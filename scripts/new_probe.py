"""LigeroCircuit::new at 2^24 gates: constraint matrix on the device vs on the host (LG_DEBUG_TIMING=1 prints the stages)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ligero_b200 as lb
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 24
ctx = lb.Context(0)
circ, out, assign = lb.ArithmeticCircuit.synthetic(1 << lg, 2024)
for mode in ("1", "1", "0"):
    os.environ["LG_CSC_DEVICE"] = mode
    t = time.perf_counter()
    lc = lb.LigeroCircuit(ctx, circ, [out])
    print(f"2^{lg} gates, LG_CSC_DEVICE={mode}: LigeroCircuit::new {time.perf_counter() - t:.2f} s", flush=True)
    del lc
