"""Seeded synthetic inputs for the CompaCT hot path.

The reference ships no dataset (its DICOM files live on the author's disk, `scripts/demo.py:46`),
so parity tests and `bench.py` use these generators.  Every image is inside the reference's
lossless domain (12-bit modular FULL delta, `core.py:323` / `core.py:503`): `synth_ct` clips to
[0, 2047], which is the simple sufficient condition of SURVEY §8d.
"""
from __future__ import annotations

import numpy as np


def synth_ct(seed: int, n: int = 512) -> np.ndarray:
    """One synthetic n×n CT-like slice, uint16, 12-bit container / 11-bit dynamic range.

    FOV disc (0 outside), air 25±2.5, body ellipse ≈1000 (±60 low-frequency, σ≈28 noise),
    two lungs ≈180 with 22 % vessel speckle (+U[200,900]) so the meshing path is exercised,
    spine ≈1900, ~10 rib blobs ≈1700; clipped to [0, 2047].
    """
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:n, 0:n].astype(np.float32)
    c = (n - 1) / 2.0
    u = (x - c) / (n / 2.0)
    v = (y - c) / (n / 2.0)
    img = np.zeros((n, n), dtype=np.float32)

    fov = (u * u + v * v) <= 0.98 ** 2
    img[fov] = 25.0
    img += fov * rng.normal(0.0, 2.5, size=(n, n)).astype(np.float32)

    # body ellipse with a slightly random pose
    bx, by = rng.uniform(-0.04, 0.04, size=2)
    ba, bb = rng.uniform(0.70, 0.80), rng.uniform(0.52, 0.62)
    body = ((u - bx) / ba) ** 2 + ((v - by) / bb) ** 2 <= 1.0
    ph = rng.uniform(0, 2 * np.pi, size=2)
    lowf = 60.0 * np.sin(3.1 * u + ph[0]) * np.cos(2.3 * v + ph[1])
    tissue = 1000.0 + lowf + rng.normal(0.0, 28.0, size=(n, n))
    img = np.where(body, tissue, img)

    # lungs with high-contrast vessel speckle
    for sx in (-1.0, 1.0):
        lx = bx + sx * rng.uniform(0.30, 0.36)
        ly = by + rng.uniform(-0.06, 0.02)
        la, lb = rng.uniform(0.20, 0.26), rng.uniform(0.30, 0.38)
        lung = (((u - lx) / la) ** 2 + ((v - ly) / lb) ** 2 <= 1.0) & body
        lv = 180.0 + rng.normal(0.0, 20.0, size=(n, n))
        speck = rng.random(size=(n, n)) < 0.22
        lv = lv + speck * rng.uniform(200.0, 900.0, size=(n, n))
        img = np.where(lung, lv, img)

    # spine
    sp = ((u - bx) / 0.09) ** 2 + ((v - (by + 0.36)) / 0.10) ** 2 <= 1.0
    img = np.where(sp & body, 1900.0 + rng.normal(0.0, 35.0, size=(n, n)), img)

    # rib blobs on an ellipse just inside the body wall
    for k in range(10):
        ang = rng.uniform(0, 2 * np.pi)
        rx = bx + 0.92 * ba * np.cos(ang)
        ry = by + 0.92 * bb * np.sin(ang)
        rr = rng.uniform(0.025, 0.045)
        rib = ((u - rx) ** 2 + (v - ry) ** 2 <= rr * rr) & body
        img = np.where(rib, 1700.0 + rng.normal(0.0, 30.0, size=(n, n)), img)

    return np.clip(np.rint(img), 0, 2047).astype(np.uint16)


def synth_ct_batch(seeds, n: int = 512) -> np.ndarray:
    """Stack of `synth_ct(seed, n)` slices, shape (B, n, n) uint16."""
    seeds = list(seeds)
    out = np.empty((len(seeds), n, n), dtype=np.uint16)
    for i, s in enumerate(seeds):
        out[i] = synth_ct(int(s), n)
    return out


def synth_radiograph(seed: int, n: int = 2048) -> np.ndarray:
    """n×n 16-bit "radiograph" (BASELINE.json configs[3]).

    `14·(x+y)` ramp (starts at 0 where the Hilbert curve starts) + `1500·(1+sin(x/97)·sin(y/131))`
    faded in over the first 400 px + N(0,40) noise faded in over the first 50 px; every step
    along any traversal stays far inside the ±2047 FULL-delta window.
    """
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:n, 0:n].astype(np.float64)
    s = 14.0 * 2048.0 / n
    ramp = s * (x + y)
    r = x + y
    fade_tex = np.clip(r / (400.0 * n / 2048.0), 0.0, 1.0)
    fade_noise = np.clip(r / (50.0 * n / 2048.0), 0.0, 1.0)
    tex = 1500.0 * (1.0 + np.sin(x / 97.0) * np.sin(y / 131.0)) * fade_tex
    noise = rng.normal(0.0, 40.0, size=(n, n)) * fade_noise
    return np.clip(np.rint(ramp + tex + noise), 0, 65535).astype(np.uint16)


def random_walk_image(seed: int, h: int, w: int, step: int = 90, hi: int = 2047) -> np.ndarray:
    """Small adversarial image: noisy values in [0, hi] with many |Δ|>64 steps (forces
    difficult blocks, JUMPs and FULL ops) but always inside the lossless domain."""
    rng = np.random.default_rng(seed)
    base = rng.integers(0, hi + 1, size=(h, w))
    smooth = rng.integers(0, 2, size=(h, w)).astype(bool)
    flat = (rng.integers(0, hi + 1) + rng.integers(-step, step + 1, size=(h, w))).clip(0, hi)
    return np.where(smooth, flat, base).astype(np.uint16)
