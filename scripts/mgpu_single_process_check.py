"""python scripts/mgpu_single_process_check.py G : ONE process driving G GPUs through lg_mgpu_* (what a Rust
LigeroCircuit::prove would bind): commit root == oracle root, proof bytes == single-GPU proof.  Prints MGPU_SP_OK / _FAIL."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ctypes import byref, c_void_p
import numpy as np
import ligero_b200 as lb
from ligero_b200.backend import _ptr, fr_to_limbs
from oracle import cref

G = int(sys.argv[1]) if len(sys.argv) > 1 else 2
ctx0 = lb.Context(0)
lib = ctx0.lib
g = c_void_p()
assert lib.lg_mgpu_create(None, G, byref(g)) == 0
ok = True
for (m, k, rho) in [(9, 256, 8), (40, 4096, 8)]:
    rng = np.random.default_rng(m)
    full = rng.integers(0, 2 ** 62, size=(4 * m * k, 4), dtype=np.uint64)
    full[:, 3] &= (1 << 60) - 1
    root = np.zeros(32, dtype=np.uint8)
    st = lib.lg_mgpu_commit(g, _ptr(full), 4 * m, k, rho, _ptr(root))
    same = st == 0 and bytes(root) == cref.commit(full, 4 * m, k, rho)["root"]
    print(f"lg_mgpu_commit m={m} k={k} on {G} GPUs: status {st} {lib.lg_mgpu_last_error(g) if st else b''} root {'==' if same else '!='} oracle", flush=True)
    ok &= same
for lg in (10, 14):
    circ, out, assign = lb.ArithmeticCircuit.synthetic(1 << lg, 3)
    L = lb.LigeroCircuit(ctx0, circ, [out])
    if L.k % G:
        continue
    ref = L.prove(assign, lb.PoseidonSponge.test_sponge()).to_bytes()
    ml = c_void_p()
    outs = np.array([out], dtype=np.uint64)
    st = lib.lg_mgpu_ligero_new(g, circ.handle, _ptr(outs), 1, lb.DEFAULT_SECURITY_LEVEL, byref(ml))
    same = False
    if st == 0:
        idx = np.array([i for i, _ in assign], dtype=np.uint64)
        vals = fr_to_limbs([v for _, v in assign])
        for rep in range(2):
            h = c_void_p()
            sponge = lb.PoseidonSponge.test_sponge()          # keep the object alive across the call
            st = lib.lg_mgpu_prove(ml, _ptr(idx), _ptr(vals), len(idx), 1, sponge.handle, byref(h))
            same = st == 0 and lb.LigeroProof(h).to_bytes() == ref
            if not same:
                break
        lib.lg_mgpu_ligero_free(ml)
    print(f"lg_mgpu_prove 2^{lg} gates on {G} GPUs: status {st} {lib.lg_mgpu_last_error(g) if st else b''} proof {'==' if same else '!='} single-GPU proof", flush=True)
    ok &= same
lib.lg_mgpu_destroy(g)
print("MGPU_SP_OK" if ok else "MGPU_SP_FAIL", flush=True)
