"""sift_gpu_b200 -- host-side mirror of the reference's include/sift.hpp interface over the C ABI.

The product is libsiftb200.so (hand-written sm_100a kernels behind include/sift_b200.h).  This module is the
thin Python binding used by the tests and the benchmark: same entry-point names and argument meaning as the
reference (SIFT_NCL, Gaussian_Blur, Gaussian_Blur_1D, buildGaussianPyramid, buildDoGPyramid,
findScaleSpaceExtrema, calDescriptor; include/sift.hpp:36-67), numpy arrays in place of cv::Mat, a structured
array with cv::KeyPoint's 28-byte layout in place of std::vector<KeyPoint>.

There is no CPU fallback: if the library is missing or no CUDA device is present, calls raise.
The directory is named sift-gpu_b200 (not importable by name); load it with __graft_entry__.load_package().
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SIFT_B200_LIB") or os.path.join(_HERE, "libsiftb200.so")  # env override: A/B builds

KP_DTYPE = np.dtype(
    [("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")]
)
NORM_L1, NORM_L2 = 2, 4
OK, ERR_CAPACITY, ERR_ARG, ERR_CUDA, ERR_TOO_SMALL, ERR_ASSERT = range(6)

# every symbol include/sift_b200.h declares (tests check the library exports all of them)
ABI_SYMBOLS = [
    "sift_b200_last_error", "sift_b200_version", "sift_b200_create", "sift_b200_destroy",
    "sift_b200_detect_describe", "sift_b200_detect_describe_batch_dev", "sift_b200_detect_describe_batch_host",
    "sift_b200_detect_describe_batch_host_u8", "sift_b200_detect_describe_batch_dev_u8", "sift_b200_rgb2gray_u8_dev", "sift_b200_upsample2x_dev", "sift_b200_detect_describe_up2", "sift_b200_gaussian_blur", "sift_b200_gaussian_blur_1d",
    "sift_b200_build_gaussian_pyramid", "sift_b200_build_dog_pyramid", "sift_b200_find_scale_space_extrema",
    "sift_b200_cal_descriptor", "sift_b200_match_knn2", "sift_b200_match_knn2_ex", "sift_b200_match_knn2_dev", "sift_b200_set_exact_pyramid", "sift_b200_launch_count", "sift_b200_set_stage_timing",
    "sift_b200_get_stage_ms", "sift_b200_chunk_plan", "sift_b200_resize_linear_u8", "sift_b200_rgb2gray_u8", "sift_b200_find_homography",
]


class SiftError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"sift_b200 error {code}: {msg}")
        self.code = code


_lib = None


def lib() -> C.CDLL:
    """Load libsiftb200.so (fails loudly when it has not been built: there is no other implementation)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
        _lib = C.CDLL(LIB_PATH)
        _lib.sift_b200_last_error.restype = C.c_char_p
        _lib.sift_b200_version.restype = C.c_char_p
        _lib.sift_b200_launch_count.restype = C.c_longlong
    return _lib


def chunk_plan(n_frames: int, max_batch: int, taper: bool = True):
    """Chunk schedule of the host-batch entry points (host logic only: works without a GPU)."""
    buf = (C.c_int * (n_frames + 16))()
    n = lib().sift_b200_chunk_plan(n_frames, max_batch, int(taper), buf, len(buf))
    if n < 0:
        raise ValueError("bad chunk_plan argument")
    return list(buf[:n])


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def octave_dims(rows: int, cols: int, n_octaves: int = 5):
    out = []
    for _ in range(n_octaves):
        out.append((rows, cols))
        rows, cols = rows // 2, cols // 2
    return out


def packed_size(rows: int, cols: int, n_octaves: int, per_octave: int) -> int:
    return sum(r * c for r, c in octave_dims(rows, cols, n_octaves)) * per_octave


def unpack(packed, rows, cols, n_octaves, per_octave):
    """Packed pyramid -> list of 2-D views in the reference's vector index order."""
    out, off = [], 0
    for r, c in octave_dims(rows, cols, n_octaves):
        for _ in range(per_octave):
            out.append(packed[off: off + r * c].reshape(r, c))
            off += r * c
    return out


class Sift:
    """Device workspace (SiftB200 handle).  One per GPU; not thread-safe."""

    def __init__(self, max_rows: int, max_cols: int, max_batch: int = 1, max_kp_per_frame: int = 16384, device: int = 0):
        self._h = C.c_void_p()
        self.max_rows, self.max_cols, self.max_batch, self.cap, self.device = max_rows, max_cols, max_batch, max_kp_per_frame, device
        self._check(lib().sift_b200_create(C.byref(self._h), max_rows, max_cols, max_batch, max_kp_per_frame, device))

    def _check(self, rc: int, allow=()):
        if rc != OK and rc not in allow:
            raise SiftError(rc, lib().sift_b200_last_error().decode())
        return rc

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib().sift_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- whole path ----
    def detect_describe(self, image: np.ndarray, cap: int | None = None):
        """SIFT_NCL on one host image (H x W float32, 0..255).  Returns (keypoints, descriptors)."""
        img = np.asarray(image, dtype=np.float32)
        if img.ndim != 2:
            raise ValueError("image must be 2-D (CV_32FC1)")
        if img.strides[1] != 4:
            img = np.ascontiguousarray(img)
        cap = cap or self.cap
        kps = np.zeros(cap, dtype=KP_DTYPE)
        desc = np.zeros((cap, 128), dtype=np.float32)
        n = C.c_int(0)
        self._check(lib().sift_b200_detect_describe(self._h, _p(img), img.shape[0], img.shape[1], C.c_size_t(img.strides[0]), _p(kps), _p(desc), cap,
                                                    C.byref(n)))
        return kps[: n.value].copy(), desc[: n.value].copy()

    def detect_describe_batch_dev(self, d_imgs, d_kp, d_desc, d_counts, cap: int, stream: int = 0):
        """Device-resident batch.  d_imgs: torch float32 (or uint8) CUDA tensor [N,H,W]; outputs: torch CUDA tensors
        d_kp uint8 [N,cap,28], d_desc float32 [N,cap,128], d_counts int32 [N].  Asynchronous on `stream`."""
        n, rows, cols = d_imgs.shape
        if not d_imgs.is_contiguous():
            raise ValueError("frames must be dense")
        fn = lib().sift_b200_detect_describe_batch_dev_u8 if d_imgs.dtype.itemsize == 1 else lib().sift_b200_detect_describe_batch_dev
        self._check(fn(self._h, C.c_void_p(d_imgs.data_ptr()), n, rows, cols, C.c_void_p(d_kp.data_ptr()), C.c_void_p(d_desc.data_ptr()),
                       C.c_void_p(d_counts.data_ptr()), cap, C.c_void_p(stream)))

    def detect_describe_batch_host(self, imgs: np.ndarray, kp_out: np.ndarray, desc_out: np.ndarray, counts_out: np.ndarray, cap: int):
        """Host batch (ideally pinned buffers): copies in, computes, copies keypoints/descriptors/counts out."""
        n, rows, cols = imgs.shape
        return self._check(lib().sift_b200_detect_describe_batch_host(self._h, _p(imgs), n, rows, cols, _p(kp_out), _p(desc_out), _p(counts_out), cap),
                           allow=(ERR_CAPACITY,))

    def detect_describe_batch_host_ptr(self, imgs_ptr: int, n: int, rows: int, cols: int, kp_ptr: int, desc_ptr: int, counts_ptr: int, cap: int):
        return self._check(lib().sift_b200_detect_describe_batch_host(self._h, C.c_void_p(imgs_ptr), n, rows, cols, C.c_void_p(kp_ptr), C.c_void_p(desc_ptr),
                                                                      C.c_void_p(counts_ptr), cap), allow=(ERR_CAPACITY,))

    def rgb2gray_u8_dev(self, d_bgr, d_gray, stream: int = 0):
        """src/main.cpp:84 colour front end on device tensors: uint8 [N,H,W,3] -> uint8 [N,H,W]."""
        n, rows, cols, _ = d_bgr.shape
        self._check(lib().sift_b200_rgb2gray_u8_dev(self._h, C.c_void_p(d_bgr.data_ptr()), n, rows, cols, C.c_void_p(d_gray.data_ptr()), C.c_void_p(stream)))

    def upsample2x_dev(self, d_src, d_dst, stream: int = 0):
        """Config 3 front end on device tensors: float32 [N,H,W] -> float32 [N,2H,2W] (cv::resize INTER_LINEAR semantics), asynchronous."""
        n, rows, cols = d_src.shape
        assert tuple(d_dst.shape) == (n, 2 * rows, 2 * cols) and d_src.is_contiguous() and d_dst.is_contiguous()
        self._check(lib().sift_b200_upsample2x_dev(self._h, C.c_void_p(d_src.data_ptr()), n, rows, cols, C.c_void_p(d_dst.data_ptr()), C.c_void_p(stream)))

    def detect_describe_batch_host_u8_ptr(self, imgs_ptr: int, n: int, rows: int, cols: int, kp_ptr: int, desc_ptr: int, counts_ptr: int, cap: int):
        return self._check(lib().sift_b200_detect_describe_batch_host_u8(self._h, C.c_void_p(imgs_ptr), n, rows, cols, C.c_void_p(kp_ptr), C.c_void_p(desc_ptr),
                                                                         C.c_void_p(counts_ptr), cap), allow=(ERR_CAPACITY,))

    def detect_describe_up2(self, image: np.ndarray, cap: int | None = None, want_upsampled: bool = False):
        """Config 3: 2x bilinear upsample (cv::resize INTER_LINEAR semantics) then SIFT_NCL; coordinates in upsampled pixels."""
        img = np.ascontiguousarray(image, dtype=np.float32)
        cap = cap or self.cap
        kps = np.zeros(cap, dtype=KP_DTYPE)
        desc = np.zeros((cap, 128), dtype=np.float32)
        up = np.empty((2 * img.shape[0], 2 * img.shape[1]), dtype=np.float32) if want_upsampled else None
        n = C.c_int(0)
        self._check(lib().sift_b200_detect_describe_up2(self._h, _p(img), img.shape[0], img.shape[1], _p(kps), _p(desc), cap, C.byref(n), _p(up)))
        out = (kps[: n.value].copy(), desc[: n.value].copy())
        return out + (up,) if want_upsampled else out

    # ---- sub-modules ----
    def gaussian_blur(self, src, sigma: float, one_d: bool = False):
        src = np.ascontiguousarray(src, dtype=np.float32)
        dst = np.empty_like(src)
        fn = lib().sift_b200_gaussian_blur_1d if one_d else lib().sift_b200_gaussian_blur
        self._check(fn(self._h, _p(src), src.shape[0], src.shape[1], C.c_double(sigma), _p(dst)))
        return dst

    def build_gaussian_pyramid(self, image, n_octaves: int = 5):
        img = np.ascontiguousarray(image, dtype=np.float32)
        rows, cols = img.shape
        g = np.empty(packed_size(rows, cols, n_octaves, 5), dtype=np.float32)
        self._check(lib().sift_b200_build_gaussian_pyramid(self._h, _p(img), rows, cols, n_octaves, _p(g)))
        return g

    def build_dog_pyramid(self, gpyr, rows, cols, n_octaves: int = 5):
        gpyr = np.ascontiguousarray(gpyr, dtype=np.float32)
        d = np.empty(packed_size(rows, cols, n_octaves, 4), dtype=np.float32)
        self._check(lib().sift_b200_build_dog_pyramid(self._h, _p(gpyr), rows, cols, n_octaves, _p(d)))
        return d

    def find_scale_space_extrema(self, gpyr, dogpyr, rows, cols, n_octaves: int = 5, cap: int | None = None):
        gpyr = np.ascontiguousarray(gpyr, dtype=np.float32)
        dogpyr = np.ascontiguousarray(dogpyr, dtype=np.float32)
        cap = cap or self.cap
        kps = np.zeros(cap, dtype=KP_DTYPE)
        n = C.c_int(0)
        self._check(lib().sift_b200_find_scale_space_extrema(self._h, _p(gpyr), _p(dogpyr), rows, cols, n_octaves, _p(kps), cap, C.byref(n)))
        return kps[: n.value].copy()

    def cal_descriptor(self, gpyr, rows, cols, kps, n_octaves: int = 5, first_octave: int = 0):
        gpyr = np.ascontiguousarray(gpyr, dtype=np.float32)
        kps = np.ascontiguousarray(kps, dtype=KP_DTYPE)
        desc = np.zeros((len(kps), 128), dtype=np.float32)
        self._check(lib().sift_b200_cal_descriptor(self._h, _p(gpyr), rows, cols, n_octaves, _p(kps), len(kps), _p(desc), first_octave))
        return desc

    # ---- the driver's readImage and homography consumer (src/main.cpp:79-87, 44-62) ----
    def resize_linear_u8(self, img: np.ndarray, drows: int, dcols: int):
        """cv::resize(img, Size(dcols, drows)) with INTER_LINEAR on uint8 pixels (H x W or H x W x C), bit-identical to cv2.resize."""
        a = np.ascontiguousarray(img, dtype=np.uint8)
        cn = 1 if a.ndim == 2 else a.shape[2]
        out = np.zeros((drows, dcols) if a.ndim == 2 else (drows, dcols, cn), dtype=np.uint8)
        self._check(lib().sift_b200_resize_linear_u8(self._h, _p(a), a.shape[0], a.shape[1], cn, _p(out), drows, dcols))
        return out

    def rgb2gray_u8(self, bgr: np.ndarray):
        """cvtColor(COLOR_RGB2GRAY) applied to BGR bytes as the driver does (src/main.cpp:84)."""
        a = np.ascontiguousarray(bgr, dtype=np.uint8)
        assert a.ndim == 3 and a.shape[2] == 3
        out = np.zeros(a.shape[:2], dtype=np.uint8)
        self._check(lib().sift_b200_rgb2gray_u8(self._h, _p(a), a.shape[0], a.shape[1], _p(out)))
        return out

    def find_homography(self, src_xy, dst_xy, ransac_thresh: float = 3.0, max_iters: int = 2000):
        """findHomography(src, dst, RANSAC).  Returns (H 3x3 float64 or None, inlier mask)."""
        s = np.ascontiguousarray(src_xy, dtype=np.float32).reshape(-1, 2)
        d = np.ascontiguousarray(dst_xy, dtype=np.float32).reshape(-1, 2)
        assert len(s) == len(d)
        H = np.zeros(9, dtype=np.float64)
        mask = np.zeros(len(s), dtype=np.uint8)
        n_in = C.c_int(0)
        rc = self._check(lib().sift_b200_find_homography(self._h, _p(s), _p(d), len(s), C.c_double(ransac_thresh), max_iters, _p(H), _p(mask), C.byref(n_in)),
                         allow=(ERR_TOO_SMALL,))
        if rc == ERR_TOO_SMALL:
            return None, mask.astype(bool)
        return H.reshape(3, 3), mask.astype(bool)

    def match_knn2(self, query, train, norm: int = NORM_L1, ratio: float = 0.86, tensor_cores: bool = False, timing: bool = False):
        """BFMatcher(norm).knnMatch(k=2) + ratio test (src/main.cpp:25-40).  tensor_cores=True (NORM_L2 only) takes the tcgen05
        shortlist + exact re-rank kernels; timing=True also returns the CUDA-event time of the kernels in ms."""
        q = np.ascontiguousarray(query, dtype=np.float32)
        t = np.ascontiguousarray(train, dtype=np.float32)
        idx = np.zeros((len(q), 2), dtype=np.int32)
        dist = np.zeros((len(q), 2), dtype=np.float32)
        good = np.zeros(len(q), dtype=np.uint8)
        ms = C.c_float(0.0)
        self._check(lib().sift_b200_match_knn2_ex(self._h, _p(q), len(q), _p(t), len(t), norm, C.c_double(ratio), _p(idx), _p(dist), _p(good),
                                                  int(tensor_cores), C.byref(ms) if timing else None))
        if timing:
            return idx, dist, good.astype(bool), float(ms.value)
        return idx, dist, good.astype(bool)

    def match_knn2_dev(self, d_query, d_train, d_idx, d_dist, norm: int = NORM_L1, tensor_cores: bool = False, stream: int = 0):
        """Device-resident matcher: torch CUDA tensors [nq,128] / [nt,128] float32 in, [nq,2] int32 / float32 out, asynchronous."""
        self._check(lib().sift_b200_match_knn2_dev(self._h, C.c_void_p(d_query.data_ptr()), d_query.shape[0], C.c_void_p(d_train.data_ptr()),
                                                   d_train.shape[0], norm, C.c_void_p(d_idx.data_ptr()), C.c_void_p(d_dist.data_ptr()),
                                                   int(tensor_cores), C.c_void_p(stream)))

    # ---- introspection ----
    def launch_count(self) -> int:
        return int(lib().sift_b200_launch_count(self._h))

    def set_exact_pyramid(self, on: bool):
        """Replay the reference's non-separable Gaussian_Blur loop bit for bit (src/sift.cpp:110-153); ~4x slower, for validation."""
        self._check(lib().sift_b200_set_exact_pyramid(self._h, int(on)))

    def set_stage_timing(self, on: bool):
        self._check(lib().sift_b200_set_stage_timing(self._h, int(on)))

    def stage_ms(self):
        ms = (C.c_float * 8)()
        self._check(lib().sift_b200_get_stage_ms(self._h, ms))
        return list(ms)


# ---- reference-named free functions (include/sift.hpp:36-67), one shared default handle -----------------------
_default: Sift | None = None


def _handle(rows: int, cols: int) -> Sift:
    global _default
    if _default is None or rows > _default.max_rows or cols > _default.max_cols:
        if _default is not None:
            _default.close()
        _default = Sift(max(rows, 64), max(cols, 64), max_batch=1, max_kp_per_frame=1 << 16)
    return _default


def SIFT_NCL(image):
    """include/sift.hpp:41-43 -- returns (keypoints, descriptors) instead of filling out-parameters."""
    img = np.asarray(image)
    return _handle(*img.shape).detect_describe(img)


def Gaussian_Blur(src, sigma):
    src = np.asarray(src)
    return _handle(*src.shape).gaussian_blur(src, sigma)


def Gaussian_Blur_1D(src, sigma):
    src = np.asarray(src)
    return _handle(*src.shape).gaussian_blur(src, sigma, one_d=True)


def buildGaussianPyramid(image, nOctaves):
    img = np.asarray(image)
    return _handle(*img.shape).build_gaussian_pyramid(img, nOctaves)


def buildDoGPyramid(gpyr, rows, cols, nOctaves):
    return _handle(rows, cols).build_dog_pyramid(gpyr, rows, cols, nOctaves)


def findScaleSpaceExtrema(gpyr, dogpyr, rows, cols, nOctaves):
    return _handle(rows, cols).find_scale_space_extrema(gpyr, dogpyr, rows, cols, nOctaves)


def calDescriptor(gpyr, rows, cols, keypoints, nOctaves=5, firstOctave=0):
    return _handle(rows, cols).cal_descriptor(gpyr, rows, cols, keypoints, nOctaves, firstOctave)
