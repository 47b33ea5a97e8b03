"""Host check of the Fiat-Shamir side's field arithmetic (ligero_b200/csrc/host_field.h) against Python integers: the
portable Montgomery product, the MULX variant picked at run time on CPUs with BMI2, additions and the S-box powers."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
RINV = pow(1 << 256, -1, R)


def test_host_field_against_python_integers(tmp_path):
    exe = os.path.join(str(tmp_path), "host_field_test")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "host", "host_field_test.cpp")], check=True)
    out = subprocess.run([exe, "4000", "11"], check=True, capture_output=True, text=True).stdout.splitlines()
    assert out[0].startswith("mulx ")
    cur, n = {}, 0
    for line in out[1:] + ["a 0"]:
        k, v = line.split()
        if k == "a" and cur:
            a, b = cur["a"], cur["b"]
            assert a < R and b < R
            want = a * b * RINV % R                      # Montgomery product of the raw limbs
            assert cur["mp"] == want and cur["md"] == want
            if "mx" in cur:
                assert cur["mx"] == want
            assert cur["add"] == (a + b) % R and cur["sub"] == (a - b) % R
            # pow_u64 works on Montgomery residues: (aR^-1)^e R
            x = a * RINV % R
            for e in (17, 5, 6):
                assert cur[f"p{e}"] == pow(x, e, R) * (1 << 256) % R
            n += 1
            cur = {}
        cur[k] = int(v, 16)
    assert n == 4000


def test_circuit_host_passes_against_plain_loops(tmp_path):
    """ligero_b200/csrc/circuit_host.h (the threaded host passes of LigeroCircuit::new: formatted node copy, compact node arrays,
    the constant-constant gate check, slot map, reachability, level schedule) against single-threaded statements of the same arrays,
    on random, deep, chain-shaped and gate-less circuits at 1, 2, 3 and 8 threads (src/ligero/mod.rs:179-194, 230-271, 476-478)."""
    exe = os.path.join(str(tmp_path), "circuit_host_test")
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-o", exe, os.path.join(ROOT, "tests", "host", "circuit_host_test.cpp")], check=True)
    for seed in ("1", "20240"):
        out = subprocess.run([exe, seed], check=True, capture_output=True, text=True).stdout
        assert out.startswith("CIRCUIT_HOST_OK 160"), out
