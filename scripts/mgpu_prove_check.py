"""torchrun --nproc-per-node G scripts/mgpu_prove_check.py [log2_gates ...]

Multi-GPU prover (lg_shard_prove*, through ligero_b200.parallel.ShardedProver) == single-GPU prover, byte for byte, on
the seeded synthetic Add/Mul circuits (lg_circuit_synthetic); the sharded proof is also put through LigeroCircuit::verify.
(The single-GPU prover is pinned to the oracle by tests/test_gpu_host_driver.py.)  Prints MGPU_PROVE_OK / _FAIL."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import ligero_b200 as lb
from ligero_b200 import parallel as par

def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    ctx = lb.Context(lr)
    ok = True
    for lg in [int(a) for a in sys.argv[1:]] or [6, 10, 14]:
        circ, out, assignment = lb.ArithmeticCircuit.synthetic(1 << lg, 100 + lg)
        L = lb.LigeroCircuit(ctx, circ, [out], lb.DEFAULT_SECURITY_LEVEL)
        if L.k % world or (L.n // world) < 2:
            if rank == 0:
                print(f"2^{lg} gates: k={L.k} too small for {world} ranks, skipped", flush=True)
            continue
        pre = L.witness_matrix(assignment)                       # every rank builds the same witness
        sp = par.ShardedProver(ctx, L, rank, world)
        local = sp.local_rows(pre)
        t0 = time.perf_counter()
        proof = sp.prove_matrix(local, lb.PoseidonSponge.test_sponge())
        dt = time.perf_counter() - t0
        blob = proof.to_bytes()
        accepted = L.verify(proof, lb.PoseidonSponge.test_sponge())
        # the same from the assignment alone: replicated device trace, local rows, sharded commit and tests
        blob_dev = sp.prove(assignment, lb.PoseidonSponge.test_sponge()).to_bytes()
        accepted = accepted and blob_dev == blob
        same = True
        if rank == 0:
            t0 = time.perf_counter()
            ref = L.prove_matrix(pre, lb.PoseidonSponge.test_sponge()).to_bytes()
            d1 = time.perf_counter() - t0
            same = ref == blob
            print(f"2^{lg} gates (m={L.m} k={L.k} n={L.n} t={L.t}) world={world}: sharded proof "
                  f"{'==' if same else '!='} single-GPU proof ({len(blob)} bytes), verify={accepted}, "
                  f"{dt * 1e3:.1f} ms sharded vs {d1 * 1e3:.1f} ms single", flush=True)
        ok &= same and accepted
        sp.close()
        dist.barrier()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MGPU_PROVE_OK" if int(flag.item()) else "MGPU_PROVE_FAIL", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
