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
    t = bench.ncu_traffic("ntt_local_kernel", 16388, 8192)
    assert t is not None and 30e9 < t < 45e9          # algorithmic 34.4 GB per launch
    assert bench.ncu_traffic("ntt_local_kernel", 4100, 2048) is None


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
