"""The C-ABI library: loads, exports every symbol include/sift_b200.h declares, and fails loudly without a GPU.
No compute calls here (CPU-only container)."""
import ctypes as C
import os
import re

import numpy as np
import pytest


def _declared_symbols(root):
    txt = open(os.path.join(root, "include", "sift_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(sift_b200_\w+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(ge, pkg):
    ge.build()
    lib = pkg.lib()
    declared = _declared_symbols(ge.ROOT)
    assert len(declared) >= 18
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert sorted(pkg.ABI_SYMBOLS) == declared  # the Python binding covers the whole header


def test_keypoint_layout_is_cv_keypoint(pkg):
    assert pkg.KP_DTYPE.itemsize == 28
    assert [pkg.KP_DTYPE.fields[n][1] for n in pkg.KP_DTYPE.names] == [0, 4, 8, 12, 16, 20, 24]


def test_sass_is_sm100a(ge):
    out = os.popen(f"cuobjdump -lelf {ge.PKG_DIR}/libsiftb200.so 2>/dev/null").read()
    assert "sm_100a" in out


def test_no_cpu_fallback(pkg):
    """Without a CUDA device the library must refuse to work rather than compute on the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(pkg.SiftError) as e:
        pkg.Sift(64, 64)
    assert e.value.code == pkg.ERR_CUDA
    h = C.c_void_p()
    assert pkg.lib().sift_b200_create(C.byref(h), 8, 8, 1, 16, 0) == pkg.ERR_ARG  # argument check precedes device probing


def test_image_size_limit_is_checked_before_anything_else(pkg):
    """Extrema candidates are packed as o<<27 | layer<<26 | row<<13 | col (detect.cu), so rows and cols must stay below 8192: the limit
    is an argument error at create time (and again per call), never a silent wrap of the 13-bit fields."""
    h = C.c_void_p()
    lib = pkg.lib()
    for rows, cols in ((8192, 64), (64, 8192), (1 << 20, 1 << 20), (15, 64), (64, 15)):
        assert lib.sift_b200_create(C.byref(h), rows, cols, 1, 16, 0) == pkg.ERR_ARG, (rows, cols)
    assert b"8192" in lib.sift_b200_last_error()
    assert lib.sift_b200_create(C.byref(h), 8191, 8191, 0, 16, 0) == pkg.ERR_ARG  # batch < 1
    assert lib.sift_b200_create(C.byref(h), 8191, 8191, 1, 0, 0) == pkg.ERR_ARG  # capacity < 1
    import torch

    if not torch.cuda.is_available():  # the largest admissible size passes the argument check and only then fails for want of a device
        assert lib.sift_b200_create(C.byref(h), 8191, 8191, 1, 16, 0) == pkg.ERR_CUDA


def test_product_never_imports_the_oracle(ge):
    """oracle/ is test infrastructure: nothing under sift-gpu_b200/ or include/ -- sources AND build recipes -- may mention it, and the
    OpenCV stand-in the product builds against (third_party/cvshim) may reach oracle headers only behind CVSHIM_ORACLE_PRIMS, a macro no
    product recipe defines."""
    bad = []
    for base in (ge.PKG_DIR, os.path.join(ge.ROOT, "include")):
        for dp, _, fs in os.walk(base):
            for f in fs:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp", ".mk")) or f == "Makefile":
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if f == "Makefile" or f.endswith(".mk"):  # recipes: no include / library path, no macro that opens the oracle door
                        txt = "\n".join(l for l in txt.splitlines() if not l.lstrip().startswith("#"))
                        if re.search(r"oracle|CVSHIM_ORACLE_PRIMS", txt):
                            bad.append(os.path.join(dp, f))
                    elif re.search(r"liboracle|oracle\.py|sift_oracle|load_oracle|#include\s+[\"<]oracle|CVSHIM_ORACLE_PRIMS", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad
    shim = os.path.join(ge.ROOT, "third_party", "cvshim")
    for dp, _, fs in os.walk(shim):
        for f in fs:
            txt = open(os.path.join(dp, f), errors="ignore").read()
            for m in re.finditer(r'#include\s+"(oracle[^"]*)"', txt):
                head = txt[: m.start()]
                assert head.count("#ifdef CVSHIM_ORACLE_PRIMS") > head.count("#endif") - head.count("#ifndef") - head.count("#if "), (f, m.group(1))
