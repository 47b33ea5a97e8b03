"""Load the committed fixtures (tests/golden) without touching /root/reference."""
import json
import os

import numpy as np

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def expected():
    return json.load(open(os.path.join(HERE, "expected.json")))


def _int(limbs):
    return int(limbs[0]) | (int(limbs[1]) << 64) | (int(limbs[2]) << 128) | (int(limbs[3]) << 192)


def load_r1cs(name):
    """-> (a_rows, b_rows, c_rows, n_wires, witness ints) with rows as [(coeff, column)]"""
    d = np.load(os.path.join(HERE, name + "_r1cs.npz"))
    mats = []
    for m in "abc":
        ptr, col, val = d[m + "_ptr"], d[m + "_col"], d[m + "_val"]
        rows = []
        for r in range(len(ptr) - 1):
            rows.append([(_int(val[e]), int(col[e])) for e in range(int(ptr[r]), int(ptr[r + 1]))])
        mats.append(rows)
    wit = [_int(w) for w in d["witness"]] if "witness" in d else None
    return mats[0], mats[1], mats[2], int(d["n_wires"][0]), wit
