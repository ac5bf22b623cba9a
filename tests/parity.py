"""Shared parity helpers for the GPU tests (tolerances are BASELINE.json's north star, written out here).

keypoints : one-to-one match on (octave, layer) with |dx|,|dy| <= 0.01 px and |dangle| <= 1 deg; recall / precision >= 0.99
descriptor: L2 <= 1e-3 after normalisation -- except that the reference quantises to uchar INSIDE the float pipeline
            (src/sift.cpp:709) and then takes sqrt(q/sum q): one +-1 LSB flip moves a component by >= 1.2e-3.
            So a pair may exceed 1e-3 only if every differing quantised integer differs by exactly 1 AND the oracle's
            pre-quantisation value sat within QUANT_EDGE of a rounding boundary.
"""
from collections import defaultdict

import numpy as np

POS_TOL = 0.01
ANG_TOL = 1.0
DESC_TOL = 1e-3
QUANT_EDGE = 0.05


def match_keypoints(a, b, pos_tol=POS_TOL, ang_tol=ANG_TOL):
    """Greedy one-to-one match of structured keypoint arrays; returns list of (ia, ib, dpos, dang)."""
    buckets = defaultdict(list)
    for j in range(len(b)):
        buckets[(int(b["octave"][j]) & 0xFFFF, int(round(float(b["x"][j]))), int(round(float(b["y"][j]))))].append(j)
    used, pairs = set(), []
    for i in range(len(a)):
        o = int(a["octave"][i]) & 0xFFFF
        ax, ay, aa = float(a["x"][i]), float(a["y"][i]), float(a["angle"][i])
        cx, cy = int(round(ax)), int(round(ay))
        best = None
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for j in buckets.get((o, cx + dx, cy + dy), ()):
                    if j in used:
                        continue
                    dpos = max(abs(ax - float(b["x"][j])), abs(ay - float(b["y"][j])))
                    dang = abs(aa - float(b["angle"][j]))
                    dang = min(dang, 360.0 - dang)
                    if dpos <= pos_tol and dang <= ang_tol and (best is None or (dpos, dang) < best[1:]):
                        best = (j, dpos, dang)
        if best is not None:
            used.add(best[0])
            pairs.append((i, best[0], best[1], best[2]))
    return pairs


def recall_precision(pairs, n_a, n_b):
    return len(pairs) / max(1, n_b), len(pairs) / max(1, n_a)


def quantised_ints(desc_row, sum_hint):
    """Recover the uchar vector q from out = sqrt(q/sum q): try integer sums near the hint, keep the most integral."""
    sq = desc_row.astype(np.float64) ** 2
    best = None
    for s in range(max(1, sum_hint - 40), sum_hint + 41):
        v = sq * s
        err = np.abs(v - np.round(v)).max()
        if best is None or err < best[0]:
            best = (err, np.round(v).astype(np.int64))
    return best[1], best[0]


def descriptor_report(desc_gpu, desc_ref, prequant_ref=None):
    """Returns (frac_within_tol, n_explained_flips, n_unexplained, max_dist)."""
    dist = np.linalg.norm(desc_gpu.astype(np.float64) - desc_ref.astype(np.float64), axis=1)
    bad = np.nonzero(dist > DESC_TOL)[0]
    explained = unexplained = 0
    for i in bad:
        if prequant_ref is None:
            unexplained += 1
            continue
        q_ref = np.clip(np.rint(prequant_ref[i].astype(np.float64)), 0, 255).astype(np.int64)
        q_gpu, err = quantised_ints(desc_gpu[i], int(q_ref.sum()))
        diff = np.nonzero(q_gpu != q_ref)[0]
        frac = prequant_ref[i].astype(np.float64) - np.floor(prequant_ref[i].astype(np.float64))
        near_edge = np.abs(frac[diff] - 0.5) <= QUANT_EDGE
        if err < 1e-3 and len(diff) > 0 and np.all(np.abs(q_gpu[diff] - q_ref[diff]) == 1) and np.all(near_edge):
            explained += 1
        else:
            unexplained += 1
    return float(np.mean(dist <= DESC_TOL)) if len(dist) else 1.0, explained, unexplained, float(dist.max()) if len(dist) else 0.0


def scan_keys(kps, rows, cols):
    """(octave, layer) of each keypoint -- the coarse part of the reference's output order (o asc, layer ... )."""
    o = kps["octave"] & 255
    return o
