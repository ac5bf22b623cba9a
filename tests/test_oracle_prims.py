"""Pin the restated OpenCV primitives (oracle/oracle_prims.h) against the cv2 wheel where cv2 exposes them.
OpenCV is the reference's third-party dependency (makefile:28-29, OpenCV 4.0.x); cv2 here is 4.13, so these pins
are tolerance checks of the published algorithms, not bit-exact claims."""
import ctypes as C

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")


@pytest.fixture(scope="module")
def L(oracle):
    lib = oracle.lib()
    lib.oracle_prim_fast_atan2.restype = C.c_float
    lib.oracle_prim_fast_atan2.argtypes = [C.c_float, C.c_float]
    lib.oracle_prim_cv_round.argtypes = [C.c_double]
    lib.oracle_prim_cv_floor.argtypes = [C.c_float]
    lib.oracle_prim_saturate_u8.argtypes = [C.c_float]
    lib.oracle_prim_exp.restype = C.c_float
    lib.oracle_prim_exp.argtypes = [C.c_float]
    lib.oracle_prim_magnitude.restype = C.c_float
    lib.oracle_prim_magnitude.argtypes = [C.c_float, C.c_float]
    return lib


def test_fast_atan2_matches_cv2_phase(L):
    rng = np.random.default_rng(0)
    x = rng.normal(0, 10, 5000).astype(np.float32)
    y = rng.normal(0, 10, 5000).astype(np.float32)
    x[:4] = [0, 1, -1, 0]
    y[:4] = [0, 0, 0, -1]
    ours = np.array([L.oracle_prim_fast_atan2(float(b), float(a)) for a, b in zip(x, y)], dtype=np.float32)
    theirs = cv2.phase(x.reshape(1, -1), y.reshape(1, -1), angleInDegrees=True).ravel()
    d = np.abs(ours - theirs)
    d = np.minimum(d, 360 - d)
    assert d.max() < 1e-3  # same polynomial; cv2 4.13's SIMD path differs by a few float ulps at 360 deg
    exact = np.degrees(np.arctan2(y.astype(np.float64), x.astype(np.float64))) % 360
    e = np.abs(ours - exact)
    assert np.minimum(e, 360 - e).max() < 0.012  # the polynomial's own model error


def test_rounding_is_half_to_even(L):
    for v, want in [(0.5, 0), (1.5, 2), (2.5, 2), (-0.5, 0), (-1.5, -2), (2.4999, 2), (254.5, 254), (255.5, 256)]:
        assert L.oracle_prim_cv_round(v) == want
    vals = np.array([0.5, 1.5, 2.5, 3.5, 254.5, 255.5, 300.0, -3.0, 100.49], dtype=np.float32).reshape(1, -1)
    cv = cv2.add(vals, 0, dtype=cv2.CV_8U).ravel()  # saturate_cast<uchar>(float)
    ours = [L.oracle_prim_saturate_u8(float(v)) for v in vals.ravel()]
    assert ours == cv.tolist()
    assert [L.oracle_prim_cv_floor(v) for v in (1.9, -0.1, -1.0, 3.0)] == [1, -1, -1, 3]


def test_exp_and_magnitude_close_to_cv2(L):
    rng = np.random.default_rng(1)
    w = (-rng.random(2000) * 12).astype(np.float32)
    ours = np.array([L.oracle_prim_exp(float(v)) for v in w], dtype=np.float32)
    assert np.allclose(ours, cv2.exp(w.reshape(1, -1)).ravel(), rtol=2e-6, atol=0)
    x = rng.normal(0, 30, 2000).astype(np.float32)
    y = rng.normal(0, 30, 2000).astype(np.float32)
    ours = np.array([L.oracle_prim_magnitude(float(a), float(b)) for a, b in zip(x, y)], dtype=np.float32)
    assert np.allclose(ours, cv2.magnitude(x.reshape(1, -1), y.reshape(1, -1)).ravel(), rtol=3e-7)


def test_solve3_matches_cv2_solve(L):
    rng = np.random.default_rng(2)
    for _ in range(50):
        a = rng.normal(0, 1, (3, 3)).astype(np.float32)
        a = (a + a.T).astype(np.float32)
        b = rng.normal(0, 1, 3).astype(np.float32)
        x = np.zeros(3, dtype=np.float32)
        L.oracle_prim_solve3(a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p))
        ok, ref = cv2.solve(a.astype(np.float64), b.astype(np.float64).reshape(3, 1), flags=cv2.DECOMP_LU)
        assert ok and np.allclose(x, ref.ravel(), rtol=2e-3, atol=2e-4)
    z = np.zeros((3, 3), dtype=np.float32)
    x = np.ones(3, dtype=np.float32)
    L.oracle_prim_solve3(z.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p))
    assert x.tolist() == [0, 0, 0]  # singular -> zeros (Matx::solve returns Matx::zeros() when the solve fails)


def test_nearest_half_resize_is_src_2y_2x():
    for h, w in [(135, 240), (67, 120), (10, 7), (64, 64)]:
        src = np.arange(h * w, dtype=np.float32).reshape(h, w)
        dst = cv2.resize(src, (w // 2, h // 2), interpolation=cv2.INTER_NEAREST)
        assert np.array_equal(dst, src[: 2 * (h // 2): 2, : 2 * (w // 2): 2])


def test_blur_restatement_matches_cv2_filter2d(oracle):
    """Independent check of the kernel/mask restatement: cv2.filter2D with the same taps on the masked source."""
    rng = np.random.default_rng(3)
    img = (rng.random((40, 56)) * 255).astype(np.float32)
    for sigma in (1.6, 2.771281):
        s = np.float32(sigma)
        w = int(np.floor(np.float32(3) * s))
        den = np.float64(np.float32(2) * s * s)
        i = np.arange(-w, w + 1)
        k = (1.0 / (2 * 3.14159265359 * np.float64(s) * np.float64(s)) * np.exp(-(i[:, None] ** 2 + i[None, :] ** 2) / den) * 8192).astype(np.float32)
        masked = img.copy()
        masked[-1, :] = 0
        masked[:, -1] = 0
        want = cv2.filter2D(masked.astype(np.float64), cv2.CV_64F, k.astype(np.float64), borderType=cv2.BORDER_CONSTANT) / 8192
        got = oracle.f32().gaussian_blur(img, sigma)
        assert np.abs(got - want).max() < 2e-4
