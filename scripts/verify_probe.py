"""On-GPU probe: phases of LigeroCircuit::verify (LG_VERIFY_TIMING=1 makes lg_verify print them).  python scripts/verify_probe.py [log2_gates]"""
import os, sys, time
os.environ["LG_VERIFY_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ligero_b200 as lb

ctx = lb.Context(0)
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 24
circ, out, assign = lb.ArithmeticCircuit.synthetic(1 << lg, 2024)
lc = lb.LigeroCircuit(ctx, circ, [out])
proof = lc.prove(assign, lb.PoseidonSponge.test_sponge())
for rep in range(int(os.environ.get("LG_VERIFY_REPS", "2"))):
    t = time.perf_counter()
    ok = lc.verify(proof, lb.PoseidonSponge.test_sponge())
    print(f"verify 2^{lg} gates: {ok} in {(time.perf_counter() - t) * 1e3:.1f} ms", flush=True)
