"""On-GPU probe: whole LigeroCircuit::prove / verify on seeded synthetic circuits (BASELINE configs 3 and 4), with the
evaluation trace on the device and on the host.  python scripts/prove_probe.py [log2_gates ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ligero_b200 as lb

ctx = lb.Context(0)
for lg in [int(a) for a in sys.argv[1:]] or [16, 20]:
    t0 = time.perf_counter()
    circ, out, assign = lb.ArithmeticCircuit.synthetic(1 << lg, 2024)
    t1 = time.perf_counter()
    lc = lb.LigeroCircuit(ctx, circ, [out])
    t2 = time.perf_counter()
    info = lc.trace_info()
    print(f"2^{lg} gates: m={lc.m} k={lc.k} n={lc.n} t={lc.t}; circuit {t1 - t0:.2f} s, LigeroCircuit::new {t2 - t1:.2f} s; "
          f"trace: {info['levels']} levels, {info['launches']} launches, on_device={info['on_device']}", flush=True)
    blobs = {}
    for mode, name in ((1, "device trace"), (0, "host trace")):
        lc.set_trace_mode(mode)
        best = 1e9
        for rep in range(3):
            ctx.sync()
            t = time.perf_counter()
            proof = lc.prove(assign, lb.PoseidonSponge.test_sponge())
            best = min(best, time.perf_counter() - t)
        blobs[name] = proof.to_bytes()
        host_phases = lc.prove_ms()
        ctx.set_timing(True)
        ctx.phase_ms()
        lc.prove(assign, lb.PoseidonSponge.test_sponge())
        dev = ctx.phase_ms()
        ctx.set_timing(False)
        print("  device phases (ms): " + ", ".join(f"{k} {v[0]:.2f}" for k, v in dev.items() if v[1]), flush=True)
        print(f"  prove ({name}): {best * 1e3:.1f} ms; phases of the last one: "
              + ", ".join(f"{k} {v:.1f}" for k, v in host_phases.items()) + f"; proof {len(blobs[name])} bytes", flush=True)
    t = time.perf_counter()
    ok = lc.verify(proof, lb.PoseidonSponge.test_sponge())
    print(f"  verify: {ok} in {(time.perf_counter() - t) * 1e3:.1f} ms; proofs equal: {blobs['device trace'] == blobs['host trace']}", flush=True)
