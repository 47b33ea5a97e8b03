"""Kernel-by-kernel view of one warm LigeroCircuit::prove (run under ncu with --profile-from-start off; the profiled
range is the SECOND proof, so no allocation or table generation is inside it):
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off \
      --csv --log-file gpurun_out/prove_launches.csv python scripts/prove_probe2.py 24"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ligero_b200 as lb

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
ctx = lb.Context(0)
circ, out, assign = lb.ArithmeticCircuit.synthetic(1 << lg, 2024)
lc = lb.LigeroCircuit(ctx, circ, [out])
lc.prove(assign, lb.PoseidonSponge.test_sponge())
ctx.sync()
torch.cuda.profiler.start()
t = time.perf_counter()
proof = lc.prove(assign, lb.PoseidonSponge.test_sponge())
dt = time.perf_counter() - t
torch.cuda.profiler.stop()
print(f"2^{lg} gates: prove {dt * 1e3:.1f} ms (under the profiler when run with ncu); phases {lc.prove_ms()}", flush=True)
