"""Deterministic synthetic stereo pairs for benchmarks and full-size tests.

Integer-valued float32 images in [0, 255]: a 3x3-box-smoothed uniform texture; the right image is the
left one warped by a piecewise-constant (``block`` x ``block``) disparity field g in [-(D-1), 0], so that
left(r, c) == right(r, c + g(r, c)) inside a block.  ``row0`` lets a rank generate its own row tile of a
taller image consistently (texture and field depend on absolute rows through per-row-block seeds).
"""
from __future__ import annotations

import numpy as np


def synthetic_pair(H: int, W: int, D: int, seed: int = 20240607, block: int = 64):
    rng = np.random.default_rng(seed)
    tex = rng.integers(0, 256, (H + 2, W + D + 2)).astype(np.int64)
    sm = sum(tex[dy: dy + H, dx: dx + W + D] for dy in range(3) for dx in range(3)) // 9
    g = -rng.integers(0, D, ((H + block - 1) // block, 1))      # one disparity per band of `block` rows
    gfull = np.kron(g, np.ones((block, W), dtype=np.int64))[:H, :W]
    cols = np.arange(W)[None, :]
    left = sm[:, D: D + W]
    right = np.take_along_axis(sm, np.clip(D + cols - gfull, 0, W + D - 1), axis=1)
    return left.astype(np.float32), right.astype(np.float32), gfull.astype(np.float32)
