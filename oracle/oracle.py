"""ctypes access to the CPU oracle (oracle/liboracle.so) and, when built, to oracle/_ref/libsift_ref.so
(the unmodified reference src/sift.cpp compiled against third_party/cvshim).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_PATH = os.path.join(HERE, "_ref", "libsift_ref.so")

KP_DTYPE = np.dtype(
    [("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")]
)
assert KP_DTYPE.itemsize == 28

N_SCALES = 5
NORM_L1, NORM_L2 = 2, 4


def build(force: bool = False) -> None:
    """Compile liboracle.so (and oracle/_ref when /root/reference exists) with oracle/Makefile."""
    if force or not os.path.exists(LIB_PATH):
        subprocess.run(["make", "-C", HERE, LIB_PATH], check=True, capture_output=True)
    if os.path.exists("/root/reference/src/sift.cpp") and (force or not os.path.exists(REF_PATH)):
        subprocess.run(["make", "-C", HERE, "ref"], check=True, capture_output=True)


def octave_dims(rows: int, cols: int, n_octaves: int = 5):
    out = []
    for _ in range(n_octaves):
        out.append((rows, cols))
        rows, cols = rows // 2, cols // 2
    return out


def packed_size(rows: int, cols: int, n_octaves: int, per_octave: int) -> int:
    return sum(r * c for r, c in octave_dims(rows, cols, n_octaves)) * per_octave


def unpack(packed: np.ndarray, rows: int, cols: int, n_octaves: int, per_octave: int):
    """Split a packed pyramid into a list of 2-D views in reference index order (o*per_octave+i)."""
    out, off = [], 0
    for r, c in octave_dims(rows, cols, n_octaves):
        for _ in range(per_octave):
            out.append(packed[off : off + r * c].reshape(r, c))
            off += r * c
    return out


def _fp(a):
    return a.ctypes.data_as(C.c_void_p)


class _Oracle:
    """One precision instance (prefix oracle32_ / oracle64_)."""

    def __init__(self, lib, prefix: str, real):
        self.lib, self.p, self.real = lib, prefix, real

    def _f(self, name):
        return getattr(self.lib, self.p + name)

    def _blur(self, fn, src, sigma):
        src = np.ascontiguousarray(src, dtype=self.real)
        dst = np.empty_like(src)
        f = self._f(fn)
        f.restype = None
        f(_fp(src), C.c_int(src.shape[0]), C.c_int(src.shape[1]), C.c_double(sigma), _fp(dst))
        return dst

    def gaussian_blur(self, src, sigma):
        return self._blur("gaussian_blur", src, sigma)

    def gaussian_blur_naive(self, src, sigma):
        return self._blur("gaussian_blur_naive", src, sigma)

    def gaussian_blur_1d(self, src, sigma):
        return self._blur("gaussian_blur_1d", src, sigma)

    def build_gaussian_pyramid(self, img, n_octaves=5):
        img = np.ascontiguousarray(img, dtype=self.real)
        rows, cols = img.shape
        g = np.empty(packed_size(rows, cols, n_octaves, 5), dtype=self.real)
        f = self._f("build_gaussian_pyramid")
        f.restype = None
        f(_fp(img), C.c_int(rows), C.c_int(cols), C.c_int(n_octaves), _fp(g))
        return g

    def build_dog_pyramid(self, gpyr, rows, cols, n_octaves=5):
        gpyr = np.ascontiguousarray(gpyr, dtype=self.real)
        d = np.empty(packed_size(rows, cols, n_octaves, 4), dtype=self.real)
        f = self._f("build_dog_pyramid")
        f.restype = None
        f(_fp(gpyr), C.c_int(rows), C.c_int(cols), C.c_int(n_octaves), _fp(d))
        return d

    def find_scale_space_extrema(self, gpyr, dogpyr, rows, cols, n_octaves=5, cap=1 << 18, debug=False):
        gpyr = np.ascontiguousarray(gpyr, dtype=self.real)
        dogpyr = np.ascontiguousarray(dogpyr, dtype=self.real)
        kps = np.zeros(cap, dtype=KP_DTYPE)
        n = C.c_int(0)
        f = self._f("find_scale_space_extrema")
        f.restype = C.c_int
        if debug:
            cand = np.zeros((cap, 4), dtype=np.int32)
            refd = np.zeros((cap, 4), dtype=np.int32)
            hists = np.zeros((cap, 36), dtype=self.real)
            nc, nr = C.c_int(0), C.c_int(0)
            rc = f(_fp(gpyr), _fp(dogpyr), C.c_int(rows), C.c_int(cols), C.c_int(n_octaves), _fp(kps), C.c_int(cap), C.byref(n),
                   _fp(cand), C.c_int(cap), C.byref(nc), _fp(refd), C.c_int(cap), C.byref(nr), _fp(hists))
            if rc:
                raise RuntimeError(f"oracle extrema rc={rc}")
            return kps[: n.value].copy(), cand[: nc.value].copy(), refd[: nr.value].copy(), hists[: nr.value].copy()
        rc = f(_fp(gpyr), _fp(dogpyr), C.c_int(rows), C.c_int(cols), C.c_int(n_octaves), _fp(kps), C.c_int(cap), C.byref(n),
               None, C.c_int(0), None, None, C.c_int(0), None, None)
        if rc:
            raise RuntimeError(f"oracle extrema rc={rc}")
        return kps[: n.value].copy()

    def cal_descriptor(self, gpyr, rows, cols, kps, first_octave=0, want_prequant=False):
        gpyr = np.ascontiguousarray(gpyr, dtype=self.real)
        kps = np.ascontiguousarray(kps, dtype=KP_DTYPE)
        n = len(kps)
        desc = np.zeros((n, 128), dtype=np.float32)
        pq = np.zeros((n, 128), dtype=np.float32) if want_prequant else None
        f = self._f("cal_descriptor")
        f.restype = C.c_int
        rc = f(_fp(gpyr), C.c_int(rows), C.c_int(cols), _fp(kps), C.c_int(n), _fp(desc), _fp(pq) if want_prequant else None,
               C.c_int(first_octave))
        if rc:
            raise RuntimeError(f"oracle cal_descriptor rc={rc} (CV_Assert in the reference, src/sift.cpp:744)")
        return (desc, pq) if want_prequant else desc

    def sift_ncl(self, img, cap=1 << 18, want_pyramids=False, want_prequant=False):
        """SIFT_NCL (src/sift.cpp:59-91): returns (keypoints, descriptors[, gpyr, dogpyr][, prequant])."""
        img = np.ascontiguousarray(img, dtype=np.float32)
        rows, cols = img.shape
        kps = np.zeros(cap, dtype=KP_DTYPE)
        desc = np.zeros((cap, 128), dtype=np.float32)
        n = C.c_int(0)
        g = np.empty(packed_size(rows, cols, 5, 5), dtype=np.float32) if want_pyramids else None
        d = np.empty(packed_size(rows, cols, 5, 4), dtype=np.float32) if want_pyramids else None
        pq = np.zeros((cap, 128), dtype=np.float32) if want_prequant else None
        f = self._f("sift_ncl")
        f.restype = C.c_int
        rc = f(_fp(img), C.c_int(rows), C.c_int(cols), _fp(kps), _fp(desc), C.c_int(cap), C.byref(n),
               _fp(g) if want_pyramids else None, _fp(d) if want_pyramids else None, _fp(pq) if want_prequant else None)
        if rc:
            raise RuntimeError(f"oracle sift_ncl rc={rc}")
        out = [kps[: n.value].copy(), desc[: n.value].copy()]
        if want_pyramids:
            out += [g, d]
        if want_prequant:
            out.append(pq[: n.value].copy())
        return tuple(out)


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
    return _lib


def f32() -> _Oracle:
    return _Oracle(lib(), "oracle32_", np.float32)


def f64() -> _Oracle:
    return _Oracle(lib(), "oracle64_", np.float64)


def set_threads(n: int) -> None:
    lib().oracle_set_threads(C.c_int(int(n)))


def match_knn2(q, t, norm=NORM_L1, ratio=0.86):
    """knnMatch(k=2) + ratio test of src/main.cpp:25-40.  Returns (idx[nq,2], dist[nq,2], good[nq])."""
    q = np.ascontiguousarray(q, dtype=np.float32)
    t = np.ascontiguousarray(t, dtype=np.float32)
    nq, nt = len(q), len(t)
    idx = np.zeros((nq, 2), dtype=np.int32)
    dist = np.zeros((nq, 2), dtype=np.float32)
    good = np.zeros(nq, dtype=np.uint8)
    f = lib().oracle_match_knn2
    f.restype = C.c_int
    rc = f(_fp(q), C.c_int(nq), _fp(t), C.c_int(nt), C.c_int(norm), C.c_double(ratio), _fp(idx), _fp(dist), _fp(good))
    if rc:
        raise RuntimeError(f"oracle match rc={rc}")
    return idx, dist, good.astype(bool)


# ------------------------------------------------------------------------------------------------
# oracle/_ref: the unmodified reference
# ------------------------------------------------------------------------------------------------
def have_ref() -> bool:
    return os.path.exists(REF_PATH)


class _Ref:
    def __init__(self):
        self.lib = C.CDLL(REF_PATH)
        self.lib.ref_set_quiet(C.c_int(1))

    def omp_max_threads(self) -> int:
        return int(self.lib.ref_omp_max_threads())

    def set_threads(self, n: int) -> None:
        self.lib.ref_set_threads(C.c_int(int(n)))

    def gaussian_blur(self, src, sigma, one_d=False):
        src = np.ascontiguousarray(src, dtype=np.float32)
        dst = np.empty_like(src)
        f = self.lib.ref_gaussian_blur_1d if one_d else self.lib.ref_gaussian_blur
        f.restype = None
        f(_fp(src), C.c_int(src.shape[0]), C.c_int(src.shape[1]), C.c_double(sigma), _fp(dst))
        return dst

    def build_gaussian_pyramid(self, img, n_octaves=5):
        img = np.ascontiguousarray(img, dtype=np.float32)
        rows, cols = img.shape
        g = np.empty(packed_size(rows, cols, n_octaves, 5), dtype=np.float32)
        self.lib.ref_build_gaussian_pyramid.restype = None
        self.lib.ref_build_gaussian_pyramid(_fp(img), C.c_int(rows), C.c_int(cols), C.c_int(n_octaves), _fp(g))
        return g

    def build_dog_pyramid(self, gpyr, rows, cols, n_octaves=5):
        gpyr = np.ascontiguousarray(gpyr, dtype=np.float32)
        d = np.empty(packed_size(rows, cols, n_octaves, 4), dtype=np.float32)
        self.lib.ref_build_dog_pyramid.restype = None
        self.lib.ref_build_dog_pyramid(_fp(gpyr), C.c_int(rows), C.c_int(cols), C.c_int(n_octaves), _fp(d))
        return d

    def find_scale_space_extrema(self, gpyr, dogpyr, rows, cols, n_octaves=5, cap=1 << 18):
        gpyr = np.ascontiguousarray(gpyr, dtype=np.float32)
        dogpyr = np.ascontiguousarray(dogpyr, dtype=np.float32)
        kps = np.zeros(cap, dtype=KP_DTYPE)
        n = C.c_int(0)
        rc = self.lib.ref_find_scale_space_extrema(_fp(gpyr), _fp(dogpyr), C.c_int(rows), C.c_int(cols), C.c_int(n_octaves), _fp(kps),
                                                   C.c_int(cap), C.byref(n))
        if rc:
            raise RuntimeError(f"ref extrema rc={rc}")
        return kps[: n.value].copy()

    def cal_descriptor(self, gpyr, rows, cols, kps, n_octaves=5, first_octave=0):
        gpyr = np.ascontiguousarray(gpyr, dtype=np.float32)
        kps = np.ascontiguousarray(kps, dtype=KP_DTYPE)
        desc = np.zeros((len(kps), 128), dtype=np.float32)
        rc = self.lib.ref_cal_descriptor(_fp(gpyr), C.c_int(rows), C.c_int(cols), C.c_int(n_octaves), _fp(kps), C.c_int(len(kps)),
                                         _fp(desc), C.c_int(first_octave))
        if rc:
            raise RuntimeError(f"ref cal_descriptor rc={rc}")
        return desc

    def sift_ncl(self, img, cap=1 << 18):
        img = np.ascontiguousarray(img, dtype=np.float32)
        rows, cols = img.shape
        kps = np.zeros(cap, dtype=KP_DTYPE)
        desc = np.zeros((cap, 128), dtype=np.float32)
        n = C.c_int(0)
        rc = self.lib.ref_sift_ncl(_fp(img), C.c_int(rows), C.c_int(cols), _fp(kps), _fp(desc), C.c_int(cap), C.byref(n))
        if rc:
            raise RuntimeError(f"ref sift_ncl rc={rc}")
        return kps[: n.value].copy(), desc[: n.value].copy()


def ref() -> _Ref:
    global _ref
    if _ref is None:
        if not have_ref():
            raise FileNotFoundError(REF_PATH)
        _ref = _Ref()
    return _ref
