"""bench.py's host-side contract (no GPU): workload shapes are the reference's (src/ligero/mod.rs:171-175, 275-294),
the committed ncu summary is readable, and the reference arm prints one JSON line with the agreed keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_shapes_of_the_synthetic_configurations():
    import bench
    assert bench.shape_for_gates(20) == (4100, 2048, 16384, 1025)
    assert bench.shape_for_gates(24) == (16388, 8192, 65536, 4097)
    assert bench.METRIC == "fr_elems_per_s_encode_commit" and bench.RHO_INV == 8


def test_committed_ncu_summary_feeds_the_roofline_traffic():
    import bench
    t = bench.ncu_traffic("ntt_persist_kernel", 16388, 8192)
    assert t is not None and 30e9 < t < 45e9          # algorithmic 34.4 GB per launch
    assert bench.ncu_traffic("ntt_persist_kernel", 4100, 2048) is None
    # captures of kernels the loaded library does not contain, or of earlier rounds, are refused
    assert bench.ncu_traffic("no_such_kernel", 16388, 8192) is None
    assert bench.ncu_traffic("ntt_local_kernel", 16388, 8192) is None      # only round-1 captures exist for it


def test_micro_workload_shapes_and_work_counts():
    import argparse
    import bench
    ns = argparse.Namespace(workload="micro", rows=1 << 14, n=1 << 16, rho_inv=0, log_gates=24)
    assert bench.resolve_shape(ns) == (16384, 16384, 65536, 4096, 4, None)
    ns = argparse.Namespace(workload="circuit", rows=0, n=0, rho_inv=0, log_gates=20)
    assert bench.resolve_shape(ns) == (4100, 2048, 16384, 1025, 8, 20)
    # executed products never exceed the algorithmic count by more than the unit twiddles the persistent kernel keeps
    for (R, k, rho) in [(16388, 8192, 8), (4100, 2048, 8), (16384, 16384, 4), (344, 128, 8)]:
        q = k.bit_length() - 1
        w_mul = R * ((k // 2) * q + (rho - 1) * ((k // 2) * q + k))
        assert 0.85 * w_mul <= bench.executed_products(R, k, rho) <= w_mul
    assert bench.pinned_root(24) == "73f425e9a839a6eb3b624ce34e2ca26e3e9b36681e8c71ad307ebbd349cdc55c"


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--log-gates", "12",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, check=True).stdout
    lines = [l for l in out.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "fr_elems_per_s_encode_commit" and d["unit"] == "Fr elems/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "Fr elems/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # under torchrun only rank 0 runs the arm
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--log-gates", "12", "--gpus", "2"],
                         capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0 and not [l for l in out.stdout.splitlines() if l.startswith("{")]
