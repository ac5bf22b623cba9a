"""The reference's DRIVER (src/main.cpp) on the GPU: readImage's resize + cvtColor, the matcher, the RANSAC homography consumer
(SURVEY 8f-1, f-3, f-4), and the UNMODIFIED main.cpp itself compiled against include/sift.hpp and run as `./sift <scene> <object>`.
Run with -m gpu on a B200."""
import os
import re
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _resize_linear_u8_restated(src, dw, dh):
    """OpenCV's fixed-point bilinear for 8-bit images (resize.cpp, INTER_LINEAR), restated in numpy; bit-identical to cv2.resize of the
    4.13 wheel on data/scene.jpg -> 960x960 (checked when the fixtures were generated, and again below where cv2 is importable)."""
    sh, sw = src.shape[:2]
    s = src.reshape(sh, sw, -1).astype(np.int64)

    def sat_short(v):
        return int(max(-32768, min(32767, np.rint(np.float32(v) * np.float32(2048)))))

    def table(dn, sn, horizontal):
        scale = 1.0 / (float(dn) / sn)
        idx, co = np.zeros(dn, np.int64), np.zeros((dn, 2), np.int64)
        for d in range(dn):
            f = np.float32((d + 0.5) * scale - 0.5)
            i = int(np.floor(f))
            f = np.float32(f - np.float32(i))
            if horizontal:
                if i < 0:
                    f, i = np.float32(0), 0
                if i >= sn - 1:
                    f, i = np.float32(0), sn - 1
            idx[d], co[d] = i, (sat_short(np.float32(1) - f), sat_short(f))
        return idx, co

    xi, xa = table(dw, sw, True)
    yi, ya = table(dh, sh, False)
    x1 = np.minimum(xi + 1, sw - 1)
    hp = s[:, xi, :] * xa[:, 0][None, :, None] + s[:, x1, :] * xa[:, 1][None, :, None]
    y0, y1 = np.clip(yi, 0, sh - 1), np.clip(yi + 1, 0, sh - 1)
    out = (((ya[:, 0][:, None, None] * (hp[y0] >> 4)) >> 16) + ((ya[:, 1][:, None, None] * (hp[y1] >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8).reshape((dh, dw) + src.shape[2:])


def test_resize_and_gray_front_end(sift):
    """readImage (src/main.cpp:79-87): resize(img, img, Size(960,960)) then cvtColor(COLOR_RGB2GRAY) on BGR bytes."""
    rng = np.random.default_rng(3)
    for (h, w, c), (dh, dw) in [((1280, 2048, 3), (960, 960)), ((37, 53, 3), (96, 71)), ((200, 300, 1), (100, 150)), ((64, 64, 3), (64, 64))]:
        img = rng.integers(0, 256, size=(h, w, c) if c > 1 else (h, w), dtype=np.uint8)
        got = sift.resize_linear_u8(img, dh, dw)
        assert np.array_equal(got, _resize_linear_u8_restated(img, dw, dh)), (h, w, c, dh, dw)
        try:
            import cv2

            assert np.array_equal(got, cv2.resize(img, (dw, dh))), ("cv2", h, w, c, dh, dw)
        except ImportError:
            pass
    bgr = rng.integers(0, 256, size=(120, 200, 3), dtype=np.uint8)
    want = ((bgr[..., 0].astype(np.int64) * 9798 + bgr[..., 1].astype(np.int64) * 19235 + bgr[..., 2].astype(np.int64) * 3735 + 16384) >> 15).astype(np.uint8)
    assert np.array_equal(sift.rgb2gray_u8(bgr), want)


def _project(H, pts):
    p = np.c_[pts, np.ones(len(pts))] @ H.T
    return p[:, :2] / p[:, 2:3]


def test_homography_consumer(sift, golden):
    """findHomography(obj, scene, RANSAC) + perspectiveTransform (src/main.cpp:44-62).  OpenCV's random sequence is not reproduced, so
    the bar is the same consensus set and the same refit: on the known-answer case the inlier masks agree (Jaccard >= 0.98) and the
    projected corners of the 2448 x 2448 object agree with cv2's to 0.25 px (and with the true map to 1 px)."""
    z = golden("homography_query_scene")
    H, mask = sift.find_homography(z["syn_obj"], z["syn_scene"], 3.0)
    assert H is not None and abs(H[2, 2] - 1) < 1e-12
    ref = z["syn_mask"].astype(bool)
    assert (mask & ref).sum() / (mask | ref).sum() >= 0.98, (int(mask.sum()), int(ref.sum()))
    corners = np.array([[0, 0], [2448, 0], [2448, 2448], [0, 2448]], dtype=np.float64)
    got = _project(H, corners)
    assert np.abs(got - z["syn_corners"]).max() <= 0.25, np.abs(got - z["syn_corners"]).max()
    assert np.abs(got - _project(z["syn_H_true"], corners)).max() <= 1.0
    # deterministic: same inputs, same answer
    H2, mask2 = sift.find_homography(z["syn_obj"], z["syn_scene"], 3.0)
    assert np.array_equal(H, H2) and np.array_equal(mask, mask2)
    # the driver's own image pair: the reference's matcher leaves ~10 true pairs among 118, cv2 reports 11 inliers; the consensus this
    # library finds may not be smaller than OpenCV's by more than a point or two (both look at 2000 random samples)
    H, mask = sift.find_homography(z["obj"], z["scene"], 3.0)
    assert H is not None and mask.sum() >= z["mask"].sum() - 2, (int(mask.sum()), int(z["mask"].sum()))
    # degenerate inputs: fewer than 4 points, all points collinear
    assert sift.find_homography(z["obj"][:3], z["scene"][:3])[0] is None
    line = np.stack([np.arange(10.0), 2 * np.arange(10.0)], axis=1).astype(np.float32)
    assert sift.find_homography(line, line)[0] is None


def _write_pgm(path, gray):
    with open(path, "wb") as f:
        f.write(b"P5\n%d %d\n255\n" % (gray.shape[1], gray.shape[0]))
        f.write(np.ascontiguousarray(gray, dtype=np.uint8).tobytes())


def test_unmodified_reference_main_cpp(ge, golden, tmp_path):
    """the reference's own src/main.cpp, UNMODIFIED, compiled against include/sift.hpp + libsift_dropin.so (sift-gpu_b200/host/Makefile,
    target `driver`) and run as the reference is run: ./sift <scene> <object>.  Inputs: the fixtures' gray images written as PGM (imread
    replicates gray to BGR, the driver's resize to 960x960 is then the identity and RGB2GRAY of equal channels returns the gray value, so
    SIFT_NCL sees exactly the fixture images).  Checked: the three timer lines per SIFT_NCL call (src/sift.cpp:70,80,88), the keypoint
    counts of the reference fixtures, the number of ratio-test survivors against the fixture's (L1, 0.86), a homography was produced."""
    exe = os.path.join(ge.PKG_DIR, "sift_driver")
    if not os.path.exists(exe):
        pytest.fail("sift-gpu_b200/sift_driver missing: build it where /root/reference exists (python -c 'import __graft_entry__ as g; g.build()')")
    scene, query = str(tmp_path / "scene.pgm"), str(tmp_path / "query.pgm")
    _write_pgm(scene, golden("scene_960")["gray"])
    _write_pgm(query, golden("query_2448")["gray"])
    r = subprocess.run([exe, scene, query], capture_output=True, text=True, timeout=300)
    out = r.stdout
    assert r.returncode == 0, (r.returncode, out[-2000:], r.stderr[-2000:])
    for line in ("pyramid construction time:", "keypoint localization time:", "descriptor extraction time:"):
        assert out.count(line) == 2, (line, out)
    m = re.search(r"drawMatches: (\d+) keypoints vs (\d+) keypoints, (\d+) good matches", out)
    assert m, out
    n_query, n_scene, n_good = map(int, m.groups())
    z = golden("match_query_scene")
    assert n_query == len(z["query_kp"]) and n_scene == len(golden("scene_960")["keypoints"])
    # the GPU's descriptors differ from the reference's by rare +-1 LSB quantisation flips, which can move a near-tie across the 0.86 ratio
    assert abs(n_good - int(z["good_n2"].sum())) <= 6, (n_good, int(z["good_n2"].sum()))
    assert "findHomography:" in out and "perspectiveTransform:" in out and 'imshow("Keypoints")' in out
    # usage message and exit code of the reference (src/main.cpp:12-13)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 255 and "Usage: ./sift <scene> <object>" in r.stdout
