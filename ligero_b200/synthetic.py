"""Synthetic pre-encoding matrices for the benchmarks and the full-size parity tests.

Every element is a function of (seed, global row, column) only, so a rank that owns any subset of the rows of the
R x k matrix generates exactly its rows and the committed root is the same at 1, 2, 4 and 8 GPUs -- and equal to the
root the CPU oracle computes for the whole matrix (tests/golden/full_size_root.json, scripts/pin_full_size_root.py).

Element (row i, column c), limb l in 0..3 (little endian, the Montgomery-limb layout of ark_bn254::Fr):
    z = splitmix64(seed * 0x9E3779B97F4A7C15 + 4 * (i * k + c) + l)
    limb l = z for l < 3;  limb 3 = (z >> 1) mod R_TOP,  R_TOP = top limb of r
so every element is < r (uniform over [0, R_TOP * 2^192), which misses only the top 2^-61 sliver of [0, r)).
The numpy version is the definition; the torch version is the same arithmetic in wrapping int64.
"""
from __future__ import annotations

import numpy as np

R_TOP = 0x30644E72E131A029          # most significant 64-bit limb of the BN254 scalar modulus r
_G = 0x9E3779B97F4A7C15
_M1 = 0xBF58476D1CE4E5B9
_M2 = 0x94D049BB133111EB
_MASK = (1 << 64) - 1


def _splitmix_np(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = x + np.uint64(_G)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(_M1)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(_M2)
        return z ^ (z >> np.uint64(31))


def matrix_rows_np(seed: int, row_ids, k: int) -> np.ndarray:
    """uint64[len(row_ids) * k, 4]: the rows `row_ids` (global indices) of the synthetic matrix with k columns."""
    rows = np.asarray(row_ids, dtype=np.uint64).reshape(-1, 1, 1)
    cols = np.arange(k, dtype=np.uint64).reshape(1, -1, 1)
    limb = np.arange(4, dtype=np.uint64).reshape(1, 1, 4)
    with np.errstate(over="ignore"):
        base = np.uint64((seed * _G) & _MASK)
        idx = base + np.uint64(4) * (rows * np.uint64(k) + cols) + limb
    z = _splitmix_np(idx)
    z[:, :, 3] = (z[:, :, 3] >> np.uint64(1)) % np.uint64(R_TOP)
    return np.ascontiguousarray(z.reshape(-1, 4))


def _s64(x: int) -> int:
    """the int64 with the same bits as the uint64 x"""
    x &= _MASK
    return x - (1 << 64) if x >= (1 << 63) else x


def matrix_rows_torch(seed: int, row_ids, k: int, device, chunk_rows: int = 1024):
    """int64[len(row_ids) * k, 4] on `device`: the same bits as matrix_rows_np (wrapping int64 arithmetic)."""
    import torch

    def lsr(z, n):          # logical shift right on int64
        return (z >> n) & ((1 << (64 - n)) - 1)

    row_ids = torch.as_tensor(list(row_ids) if not hasattr(row_ids, "shape") else row_ids, dtype=torch.int64, device=device)
    out = torch.empty((row_ids.numel() * k, 4), dtype=torch.int64, device=device)
    cols = torch.arange(k, dtype=torch.int64, device=device).view(1, -1, 1)
    limb = torch.arange(4, dtype=torch.int64, device=device).view(1, 1, 4)
    base = _s64(seed * _G)
    for r0 in range(0, row_ids.numel(), chunk_rows):
        rows = row_ids[r0:r0 + chunk_rows].view(-1, 1, 1)
        z = base + 4 * (rows * k + cols) + limb
        z = z + _s64(_G)
        z = (z ^ lsr(z, 30)) * _s64(_M1)
        z = (z ^ lsr(z, 27)) * _s64(_M2)
        z = z ^ lsr(z, 31)
        z[:, :, 3] = lsr(z[:, :, 3], 1) % R_TOP
        out[r0 * k:(r0 + rows.shape[0]) * k] = z.reshape(-1, 4)
    if out.is_cuda:
        # the library launches on its own non-blocking stream, which does not wait for torch's: hand over a finished matrix
        torch.cuda.synchronize(out.device)
    return out
