"""Host check of the lazy-reduction NTT arithmetic (ligero_b200/csrc/fr_lazy.cuh).

The header's carry-chain PTX statements each have a plain-C emulation generated next to them
(scripts/gen_fr_shoup.py), so g++ compiles the very algorithm the GPU runs.  Python integers are the
reference: quotient multipliers, the [0, 2r) product range, butterfly values mod r and their range invariants.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617


def _run(n, seed, tmp_path):
    exe = os.path.join(str(tmp_path), "fr_lazy_host_test")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "host", "fr_lazy_host_test.cpp")],
                   check=True)
    out = subprocess.run([exe, str(n), str(seed)], check=True, capture_output=True, text=True).stdout
    recs, cur = [], {}
    for line in out.splitlines():
        k, v = line.split()
        if k == "w" and cur:
            recs.append(cur)
            cur = {}
        cur[k] = int(v, 16)
    recs.append(cur)
    return recs


def test_generated_body_is_current():
    incs = [os.path.join(ROOT, "ligero_b200", "csrc", n) for n in ("fr_shoup_body_q.inc", "fr_shoup_body_t.inc")]
    before = [open(i).read() for i in incs]
    subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "gen_fr_shoup.py")], check=True, capture_output=True)
    assert [open(i).read() for i in incs] == before, "fr_shoup_body_*.inc are stale: run scripts/gen_fr_shoup.py"


def test_shoup_product_and_lazy_butterflies(tmp_path):
    recs = _run(3000, 7, tmp_path)
    assert len(recs) == 3000
    D224, D240 = 1 << 224, 1 << 240
    for c in recs:
        w, p, y, t = c["w"], c["p"], c["y"], c["t"]
        assert w < R and p == (w << 256) // R
        assert y < (1 << 256) - (1 << 227)
        assert t % R == (y * w) % R and 0 <= t < 2 * R
        X, Y = c["X"], c["Y"]
        # forward butterfly
        assert c["ditX"] % R == (X + w * Y) % R and c["ditX"] < 4 * R + D224
        assert c["ditY"] % R == (X - w * Y) % R and c["ditY"] < 4 * R + D224
        assert c["nX"] == (X + w * Y) % R and c["nY"] == (X - w * Y) % R
        assert c["dit1X"] % R == (X + Y) % R and c["dit1X"] < 4 * R + 2 * D224
        assert c["dit1Y"] % R == (X - Y) % R and c["dit1Y"] < 4 * R + D224
        # inverse butterfly (inputs < 2r + 2^224)
        fX, fY = c["fX"], c["fY"]
        assert fX % R == X % R and fX < 2 * R + D224 and fY < 2 * R + D224
        assert c["difX"] % R == (fX + fY) % R and c["difX"] < 2 * R + D240
        assert c["difY"] % R == ((fX - fY) * w) % R and c["difY"] < 2 * R
        assert c["dif1X"] % R == (fX + fY) % R and c["dif1X"] < 2 * R + D240
        assert c["dif1Y"] % R == (fX - fY) % R and c["dif1Y"] < 2 * R + D224
