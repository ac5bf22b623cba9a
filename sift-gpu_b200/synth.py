"""Synthetic frames for the benchmark and the parity tests: recipe S of SURVEY.md section 8(d).

A flat 128 grey frame with Gaussian blobs of random scale/sign plus mild noise, float32 in 0..255 --
the input contract of SIFT_NCL (reference src/main.cpp:84-85: gray, convertTo CV_32FC1, no scaling).
Pure numpy; deterministic per seed.
"""
from __future__ import annotations

import numpy as np


def recipe_s(width: int, height: int, seed: int = 1234, blobs_per_1080p: int = 6000) -> np.ndarray:
    rng = np.random.default_rng(seed)
    n = max(1, int(round(blobs_per_1080p * (width * height) / (1920 * 1080))))
    # draw order: all sigma, all cx, all cy, all amplitudes, all signs, then the noise field
    sigma = np.exp(rng.uniform(np.log(1.2), np.log(12.0), n))
    cx = rng.uniform(0, width, n)
    cy = rng.uniform(0, height, n)
    amp = rng.uniform(40.0, 110.0, n)
    sign = rng.integers(0, 2, n) * 2 - 1
    img = np.full((height, width), 128.0, dtype=np.float32)
    for s, x, y, a, sg in zip(sigma, cx, cy, amp, sign):
        rad = int(np.ceil(3 * s))
        x0, x1 = max(0, int(x) - rad), min(width, int(x) + rad + 1)
        y0, y1 = max(0, int(y) - rad), min(height, int(y) + rad + 1)
        if x0 >= x1 or y0 >= y1:
            continue
        xs = np.arange(x0, x1, dtype=np.float32) - np.float32(x)
        ys = np.arange(y0, y1, dtype=np.float32) - np.float32(y)
        g = np.exp(-(ys[:, None] ** 2 + xs[None, :] ** 2) / np.float32(2 * s * s))
        img[y0:y1, x0:x1] += np.float32(a * sg) * g.astype(np.float32)
    img += rng.normal(0.0, 2.0, (height, width)).astype(np.float32)
    return np.clip(img, 0.0, 255.0).astype(np.float32)


def upsample2x(img: np.ndarray) -> np.ndarray:
    """cv::resize(img, 2x, INTER_LINEAR) restated in numpy float32 (half-pixel centres, edge replicate; horizontal then
    vertical pass, mul and add rounded separately) -- the checker for the GPU upsample front end (BASELINE config 3)."""
    img = np.asarray(img, dtype=np.float32)

    def taps(n):
        d = np.arange(2 * n)
        f = ((d + 0.5) * 0.5 - 0.5).astype(np.float32)
        s0 = np.floor(f).astype(np.int64)
        w = (f - s0).astype(np.float32)
        lo = s0 < 0
        s0[lo], w[lo] = 0, 0.0
        hi = s0 >= n - 1
        s0[hi], w[hi] = n - 1, 0.0
        return s0, np.minimum(s0 + 1, n - 1), (np.float32(1) - w).astype(np.float32), w

    x0, x1, a0, a1 = taps(img.shape[1])
    y0, y1, b0, b1 = taps(img.shape[0])
    h = (img[:, x0] * a0[None, :]).astype(np.float32) + (img[:, x1] * a1[None, :]).astype(np.float32)
    return ((h[y0] * b0[:, None]).astype(np.float32) + (h[y1] * b1[:, None]).astype(np.float32)).astype(np.float32)
