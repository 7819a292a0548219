"""Synthetic rectified stereo pairs: band-limited texture,
piecewise-planar ground-truth disparity, right view rendered by warping.

There is no network and the reference ships no images (its input.txt names
Middlebury files that are not in the repo), so every test / bench input comes
from here.  Output: 8-bit BGR interleaved images, exactly the CV_8UC3 layout
the reference's imread() hands to PreSSPC / CSPatchMatch (main.cc:68-69).
"""
from __future__ import annotations

import numpy as np

SEED_BASE = 0xC5B20000


def _texture(rng: np.random.Generator, w: int, h: int, n_waves: int = 24) -> np.ndarray:
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.zeros((h, w, 3), np.float64)
    for c in range(3):
        acc = np.zeros((h, w), np.float64)
        for _ in range(n_waves):
            fx, fy = rng.uniform(-0.35, 0.35, 2)
            ph = rng.uniform(0, 2 * np.pi)
            amp = rng.uniform(6.0, 22.0)
            acc += amp * np.sin(fx * xx + fy * yy + ph)
        img[..., c] = 128.0 + acc
    img += rng.integers(-8, 9, size=(h, w, 3))
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def _gt_disparity(rng: np.random.Generator, w: int, h: int, max_dis: int, n_planes: int = 6) -> np.ndarray:
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    lo, hi = 1.0, float(max_dis - 1)
    disp = np.full((h, w), rng.uniform(lo + 0.1 * (hi - lo), lo + 0.3 * (hi - lo)))
    for _ in range(n_planes):
        cx, cy = rng.uniform(0, w), rng.uniform(0, h)
        rx, ry = rng.uniform(0.08, 0.3) * w, rng.uniform(0.1, 0.35) * h
        a, b = rng.uniform(-0.3, 0.3, 2) * min(1.0, 0.25 * (hi - lo) / max(rx, ry))
        d0 = rng.uniform(lo + 0.3 * (hi - lo), lo + 0.9 * (hi - lo))
        plane = d0 + a * (xx - cx) + b * (yy - cy)
        mask = ((xx - cx) / rx) ** 2 + ((yy - cy) / ry) ** 2 <= 1.0
        disp = np.where(mask & (plane > disp), plane, disp)
    return np.clip(disp, lo, hi)


def make_pair(w: int, h: int, max_dis: int, seed: int = SEED_BASE):
    """Return (left_bgr, right_bgr, gt_left_disparity): uint8 HxWx3 (BGR), f64 HxW."""
    rng = np.random.default_rng(seed)
    left = _texture(rng, w, h)
    gt = _gt_disparity(rng, w, h, max_dis)
    fill = _texture(rng, w, h)  # fresh texture for disocclusions
    # right-view disparity by forward warping with a z-buffer (nearest wins)
    dr = np.full((h, w), -1.0)
    xs = np.arange(w)
    for y in range(h):
        xr = np.rint(xs - gt[y]).astype(np.int64)
        ok = (xr >= 0) & (xr < w)
        order = np.argsort(gt[y][ok], kind="stable")  # ascending: larger disparity written last
        dr[y][xr[ok][order]] = gt[y][ok][order]
    right = np.empty_like(left)
    for y in range(h):
        src = xs + dr[y]
        x0 = np.clip(np.floor(src).astype(np.int64), 0, w - 1)
        x1 = np.clip(x0 + 1, 0, w - 1)
        f = (src - np.floor(src))[:, None]
        row = (1 - f) * left[y][x0].astype(np.float64) + f * left[y][x1].astype(np.float64)
        hole = (dr[y] < 0) | (src > w - 1)
        row = np.where(hole[:, None], fill[y].astype(np.float64), row)
        right[y] = np.clip(np.rint(row), 0, 255).astype(np.uint8)
    return np.ascontiguousarray(left), np.ascontiguousarray(right), gt
