"""Multi-GPU parity, driver-visible: on a box with >= 2 GPUs these run the sharded commit / prover as one process per GPU
(torchrun, NCCL rendezvous) and as one process driving all GPUs (lg_mgpu_*), and compare roots with the CPU oracle and
proofs with the single-GPU prover byte for byte.  Skipped on a one-GPU box (the same code is covered there with
world = 1 by tests/test_gpu_shard.py; the host logic by the gloo tests)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def gpu_count():
    import torch
    return torch.cuda.device_count()


def run(cmd, timeout=600):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    return subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)


def torchrun(n, script, *args, port=29611):
    return run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                "--master-port", str(port), os.path.join("scripts", script), *args])


@pytest.mark.parametrize("n", [2, 4, 8])
def test_sharded_commit_roots_equal_oracle(n):
    if gpu_count() < n:
        pytest.skip(f"needs {n} GPUs")
    r = torchrun(n, "mgpu_check.py", port=29611 + n)
    assert "MGPU_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("n", [2, 8])
def test_sharded_proofs_equal_single_gpu_proofs(n):
    if gpu_count() < n:
        pytest.skip(f"needs {n} GPUs")
    r = torchrun(n, "mgpu_prove_check.py", "10", "14", "18", port=29631 + n)
    assert "MGPU_PROVE_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("n", [2, 4])
def test_single_process_mgpu(n):
    if gpu_count() < n:
        pytest.skip(f"needs {n} GPUs")
    r = run([sys.executable, os.path.join("scripts", "mgpu_single_process_check.py"), str(n)])
    assert "MGPU_SP_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def build_c_host(tmp_path):
    exe = str(tmp_path / "mgpu_prove")
    libdir = os.path.join(ROOT, "ligero_b200")
    r = run(["gcc", "-std=c11", "-O2", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "mgpu_prove.c"), "-o", exe,
             "-L", libdir, "-lligero_b200", f"-Wl,-rpath,{libdir}"])
    assert r.returncode == 0, r.stderr
    return exe


@pytest.mark.parametrize("n", [1, 2])
def test_c_host_drives_the_multi_gpu_prover(n, tmp_path):
    """tests/c/mgpu_prove.c: a C program (no Python, no torch) proves over n GPUs through lg_mgpu_* and gets the single-GPU
    proof bytes; n = 1 runs on any GPU box, n = 2 needs two."""
    if gpu_count() < n:
        pytest.skip(f"needs {n} GPUs")
    r = run([build_c_host(tmp_path), str(n), "14"])
    assert "C_MGPU_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("log_gates", [10, 17])
def test_c_host_new_prove_verify(log_gates, tmp_path):
    """tests/c/new_prove_check.c on one GPU: LigeroCircuit::new through the C ABI (2^17 gates: constraint matrix built on the
    device, level schedule from the threaded host passes of circuit_host.h; 2^10: the host builders), the proof with the trace
    evaluated on the device equals the proof with the host evaluator byte for byte, and verifies (src/ligero/mod.rs:147-228, 435-455)."""
    exe = str(tmp_path / "new_prove_check")
    libdir = os.path.join(ROOT, "ligero_b200")
    r = run(["gcc", "-std=c11", "-O2", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "new_prove_check.c"), "-o", exe,
             "-L", libdir, "-lligero_b200", f"-Wl,-rpath,{libdir}"])
    assert r.returncode == 0, r.stderr
    r = run([exe, str(log_gates)])
    assert "C_NEW_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
