"""Seeded synthetic inputs for the CompaCT hot path.

The reference ships no dataset (its DICOM files live on the author's disk, `scripts/demo.py:46`),
so parity tests and `bench.py` use these generators.  Every image is inside the reference's
lossless domain (12-bit modular FULL delta, `core.py:323` / `core.py:503`): `synth_ct` clips to
[0, 2047], which is the simple sufficient condition of SURVEY §8d.
"""
from __future__ import annotations

import numpy as np


_GRID_CACHE: dict = {}


def _grid(n: int):
    g = _GRID_CACHE.get(n)
    if g is None:
        y, x = np.mgrid[0:n, 0:n].astype(np.float32)
        c = (n - 1) / 2.0
        g = _GRID_CACHE[n] = ((x - c) / (n / 2.0), (y - c) / (n / 2.0))
    return g


def synth_ct(seed: int, n: int = 512) -> np.ndarray:
    """One synthetic n×n CT-like slice, uint16, 12-bit container / 11-bit dynamic range.

    FOV disc (0 outside), air 25±2.5, body ellipse ≈1000 (±60 low-frequency, σ≈28 noise),
    two lungs ≈180 with 22 % vessel speckle (+U[200,900]) so the meshing path is exercised,
    spine ≈1900, ~10 rib blobs ≈1700; clipped to [0, 2047].
    """
    rng = np.random.default_rng(seed)
    u, v = _grid(n)
    z = rng.standard_normal((n, n), dtype=np.float32)  # one noise field, scaled per tissue
    r1 = rng.random((n, n), dtype=np.float32)          # speckle mask
    r2 = rng.random((n, n), dtype=np.float32)          # speckle amplitude
    img = np.zeros((n, n), dtype=np.float32)

    fov = (u * u + v * v) <= 0.98 ** 2
    np.putmask(img, fov, 25.0 + 2.5 * z)

    bx, by = rng.uniform(-0.04, 0.04, size=2)
    ba, bb = rng.uniform(0.70, 0.80), rng.uniform(0.52, 0.62)
    body = ((u - bx) / ba) ** 2 + ((v - by) / bb) ** 2 <= 1.0
    ph = rng.uniform(0, 2 * np.pi, size=2)
    lowf = 60.0 * np.sin(3.1 * u + np.float32(ph[0])) * np.cos(2.3 * v + np.float32(ph[1]))
    np.putmask(img, body, 1000.0 + lowf + 28.0 * z)

    for sx in (-1.0, 1.0):
        lx = bx + sx * rng.uniform(0.30, 0.36)
        ly = by + rng.uniform(-0.06, 0.02)
        la, lb = rng.uniform(0.20, 0.26), rng.uniform(0.30, 0.38)
        lung = (((u - lx) / la) ** 2 + ((v - ly) / lb) ** 2 <= 1.0) & body
        lv = 180.0 + 20.0 * z + (r1 < 0.22) * (200.0 + 700.0 * r2)
        np.putmask(img, lung, lv)

    sp = (((u - bx) / 0.09) ** 2 + ((v - (by + 0.36)) / 0.10) ** 2 <= 1.0) & body
    np.putmask(img, sp, 1900.0 + 35.0 * z)

    for k in range(10):
        ang = rng.uniform(0, 2 * np.pi)
        rx = bx + 0.92 * ba * np.cos(ang)
        ry = by + 0.92 * bb * np.sin(ang)
        rr = rng.uniform(0.025, 0.045)
        rib = ((u - rx) ** 2 + (v - ry) ** 2 <= rr * rr) & body
        np.putmask(img, rib, 1700.0 + 30.0 * z)

    return np.clip(np.rint(img), 0, 2047).astype(np.uint16)


def synth_ct_batch(seeds, n: int = 512, workers: int = 0) -> np.ndarray:
    """Stack of `synth_ct(seed, n)` slices, shape (B, n, n) uint16 (threads: numpy releases the GIL)."""
    seeds = [int(s) for s in seeds]
    out = np.empty((len(seeds), n, n), dtype=np.uint16)
    _grid(n)
    if workers <= 1 or len(seeds) < 4:
        for i, s in enumerate(seeds):
            out[i] = synth_ct(s, n)
        return out
    from concurrent.futures import ThreadPoolExecutor

    def work(i):
        out[i] = synth_ct(seeds[i], n)
    with ThreadPoolExecutor(workers) as ex:
        list(ex.map(work, range(len(seeds))))
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
