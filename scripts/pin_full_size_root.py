"""Pin the Merkle root of the FULL 2^24-gate encode+commit (16388 x 8192 -> x 65536, rho_inv = 8) with the CPU oracle.

    python scripts/pin_full_size_root.py [--log-gates 24] [--seed 20240]

Runs oracle/ligero_ref.c's ref_commit (the reference's schedule: per-row iFFT_k + zero-padded FFT_n, BLAKE2s per column,
SHA-256 tree) over the whole synthetic matrix of ligero_b200/synthetic.py and writes tests/golden/full_size_root.json.
Needs ~37 GiB of host memory and a few minutes of CPU time; bench.py (every N) and tests/test_gpu_full_size.py compare
the GPU root with the pinned value."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bench import shape_for_gates, RHO_INV
from ligero_b200.synthetic import matrix_rows_np
from oracle import cref

ap = argparse.ArgumentParser()
ap.add_argument("--log-gates", type=int, default=24)
ap.add_argument("--seed", type=int, default=20240)
args = ap.parse_args()
R, k, n, m = shape_for_gates(args.log_gates)
t0 = time.time()
a = np.empty((R * k, 4), dtype=np.uint64)
for r0 in range(0, R, 512):
    r1 = min(R, r0 + 512)
    a[r0 * k:r1 * k] = matrix_rows_np(args.seed, range(r0, r1), k)
t1 = time.time()
res = cref.commit(a, R, k, RHO_INV)
t2 = time.time()
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "full_size_root.json")
pins = json.load(open(path)) if os.path.exists(path) else {}
pins[f"2^{args.log_gates}/seed{args.seed}"] = {
    "rows": R, "k": k, "rho_inv": RHO_INV, "seed": args.seed, "root": res["root"].hex(),
    "by": "oracle/ligero_ref.c ref_commit over ligero_b200.synthetic.matrix_rows_np (scripts/pin_full_size_root.py)",
    "threads": res["threads"], "secs_encode": round(res["secs_encode"], 1), "secs_hash": round(res["secs_hash"], 1)}
json.dump(pins, open(path, "w"), indent=1, sort_keys=True)
print(json.dumps(pins[f"2^{args.log_gates}/seed{args.seed}"]), f"gen {t1-t0:.0f}s commit {t2-t1:.0f}s")
