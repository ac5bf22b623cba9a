"""GPU parity tests: the CUDA path, called through the C ABI (ctypes), against the CPU oracle on the same inputs and
against the golden fixtures produced by the unmodified reference.  Run with -m gpu on a B200."""
import numpy as np
import pytest

import parity

pytestmark = pytest.mark.gpu

SIZES = [(320, 240, 5), (417, 303, 9), (64, 48, 2), (960, 540, 11)]


def _oracle_all(oracle, img):
    return oracle.f32().sift_ncl(img, want_pyramids=True, want_prequant=True)


@pytest.mark.parametrize("w,h,seed", SIZES)
def test_pyramid_and_dog(sift, pkg, oracle, synth, w, h, seed):
    """buildGaussianPyramid + buildDoGPyramid (src/sift.cpp:229-283).  Separable fp32 FMA blur vs the reference's
    sequential 2-D fp32 sum: tolerance 2e-3 absolute on 0..255 data (observed <= 7.3e-4; the reference's own fp32
    rounding noise over 1369 taps is ~2e-4)."""
    img = synth.recipe_s(w, h, seed=seed, blobs_per_1080p=12000)
    _, _, og, od, _ = _oracle_all(oracle, img)
    g = sift.build_gaussian_pyramid(img)
    assert g.shape == og.shape
    assert np.abs(g - og).max() <= 2e-3
    d = sift.build_dog_pyramid(g, h, w)
    assert np.abs(d - od).max() <= 2e-3
    # the DoG stage alone is an exact float subtraction: bit-exact on identical inputs
    assert np.array_equal(sift.build_dog_pyramid(og, h, w), od)
    # NEAREST decimation (src/sift.cpp:253-254): octave o+1 base == octave o scale 2 at [2y][2x], exactly
    lv = pkg.unpack(g, h, w, 5, 5)
    for o in range(4):
        nb = lv[(o + 1) * 5]
        assert np.array_equal(nb, lv[o * 5 + 2][: 2 * nb.shape[0]: 2, : 2 * nb.shape[1]: 2])


@pytest.mark.parametrize("w,h,seed", SIZES)
def test_extrema_stage_on_oracle_pyramids(sift, oracle, synth, w, h, seed):
    """findScaleSpaceExtrema (src/sift.cpp:547-577) fed the ORACLE's pyramids: same candidates, same refinement
    arithmetic (--fmad=false), same output order.  Integer / position / response fields bit-exact; size within 1 ulp
    (exp2 vs powf); angle within 1e-2 deg (histogram summation order differs)."""
    img = synth.recipe_s(w, h, seed=seed, blobs_per_1080p=12000)
    okp, _, og, od, _ = _oracle_all(oracle, img)
    kp = sift.find_scale_space_extrema(og, od, h, w)
    assert len(kp) == len(okp)
    for f in ("x", "y", "response", "octave", "class_id"):
        assert np.array_equal(kp[f], okp[f]), f
    assert np.allclose(kp["size"], okp["size"], rtol=2e-7, atol=0)
    da = np.abs(kp["angle"] - okp["angle"])
    assert np.minimum(da, 360 - da).max() <= 1e-2


@pytest.mark.parametrize("w,h,seed", SIZES)
def test_descriptor_stage_on_oracle_inputs(sift, oracle, synth, w, h, seed):
    """calDescriptor (src/sift.cpp:733-753) fed the oracle's gpyr + keypoints.  Tolerance: L2 <= 1e-3; rows beyond it
    must be explained +-1 LSB quantisation flips (tests/parity.py)."""
    img = synth.recipe_s(w, h, seed=seed, blobs_per_1080p=12000)
    okp, odesc, og, _, opq = _oracle_all(oracle, img)
    desc = sift.cal_descriptor(og, h, w, okp)
    frac, explained, unexplained, mx = parity.descriptor_report(desc, odesc, opq)
    assert unexplained == 0 and frac >= 0.995, (frac, explained, unexplained, mx)
    assert np.allclose(np.linalg.norm(desc, axis=1), 1.0, atol=1e-5)


# Descriptor gates for the default (separable-blur) pipeline, whole path.  North star: L2 <= 1e-3.  The reference quantises to uchar
# inside the float pipeline (src/sift.cpp:709), so the ~1e-4 rounding difference between the separable blur and the reference's
# 1369-term sequential sum flips one +-1 LSB in a few per cent of rows, and one flip moves a component by >= 1.2e-3.
# Gates: (a) the fraction of rows within 1e-3 may not fall below the observed floor (MIN_FRAC: regressions trip; observed 0.955-0.993
# over the eight test images, tools/parity_report.py), (b) EVERY row beyond 1e-3 must be explained -- either a quantisation flip (all
# differing quantised integers differ by exactly 1 and the reference's pre-quantisation value sat within parity.QUANT_EDGE of a
# rounding boundary) or a keypoint whose interpolated orientation differs by a fraction of a degree (inside the 1 deg tolerance)
# and whose descriptor is right once the descriptor stage is fed the reference's keypoint record (parity.classify_unexplained)
# -- i.e. unexplained == 0, with the reference's own pre-quantisation vectors (fixtures *_prequant.npz, legitimate because the C
# port is bit-identical to oracle/_ref on those images).
MIN_FRAC = 0.95


def _end_to_end(sift, img, okp, odesc, opq, min_frac=MIN_FRAC):
    assert opq is not None, "every end-to-end check classifies the rows beyond 1e-3: pre-quantisation vectors are mandatory"
    kp, desc = sift.detect_describe(img)
    pairs = parity.match_keypoints(kp, okp)
    rec, prec = parity.recall_precision(pairs, len(kp), len(okp))
    assert rec >= 0.99 and prec >= 0.99, (rec, prec, len(kp), len(okp))
    pi = np.array([p[0] for p in pairs]); pj = np.array([p[1] for p in pairs])
    rows = []
    frac, explained, unexplained, mx = parity.descriptor_report(desc[pi], odesc[pj], opq[pj], rows_out=rows)
    assert frac >= min_frac, (frac, explained, unexplained, mx)
    by_keypoint, unexplained = parity.classify_unexplained(sift, img, kp[pi], okp[pj], odesc[pj], opq[pj], rows)
    assert unexplained == 0, (frac, explained, by_keypoint, unexplained, mx)
    assert by_keypoint <= max(2, len(pairs) // 100), (frac, explained, by_keypoint, mx)  # observed <= 0.3 % of rows
    # output order = reference scan order (src/sift.cpp:556-557,487,491,525): matched pairs are index-aligned
    if len(kp) == len(okp) and len(pairs) == len(kp):
        assert all(i == j for i, j, _, _ in pairs)
    return kp, desc


@pytest.mark.parametrize("w,h,seed", SIZES)
def test_end_to_end_vs_oracle(sift, oracle, synth, w, h, seed):
    """SIFT_NCL (src/sift.cpp:59-91) end to end.  Keypoints: recall/precision >= 0.99 at <= 0.01 px, <= 1 deg.
    Descriptors: >= MIN_FRAC within 1e-3 outright, every other row an explained +-1 LSB quantisation flip (unexplained == 0)."""
    img = synth.recipe_s(w, h, seed=seed, blobs_per_1080p=12000)
    okp, odesc, _, _, opq = _oracle_all(oracle, img)
    _end_to_end(sift, img, okp, odesc, opq)


def test_scene_960_against_reference_fixture(sift, golden):
    """BASELINE config 1: data/scene.jpg as src/main.cpp feeds it; expected values from the UNMODIFIED reference."""
    z = golden("scene_960")
    kp, _ = _end_to_end(sift, z["gray"].astype(np.float32), z["keypoints"], z["descriptors"], golden("scene_960_prequant")["prequant"])
    assert abs(len(kp) - 486) <= 5


def test_scene_native_against_reference_fixture(sift, golden):
    """BASELINE config 1 at native size: data/scene.jpg 2048x1280 without the driver's resize; unmodified-reference outputs."""
    z = golden("scene_native_2048x1280")
    kp, _ = _end_to_end(sift, z["gray"].astype(np.float32), z["keypoints"], z["descriptors"], z["prequant"])
    assert abs(len(kp) - 1364) <= 10


def test_query_2448_against_reference_fixture(sift, golden):
    """BASELINE config 5 input: data/query.jpg native 2448x2448; keypoints/descriptors from the unmodified reference."""
    q = golden("query_2448")["gray"].astype(np.float32)
    z = golden("match_query_scene")
    _end_to_end(sift, q, z["query_kp"], z["query_desc"], golden("query_2448_prequant")["prequant"])


@pytest.mark.parametrize("name", ["synth_160x120", "synth_odd_211x173"])
def test_small_fixtures_all_stages(sift, golden, name):
    z = golden(name)
    img = z["image"]
    h, w = img.shape
    assert np.abs(sift.build_gaussian_pyramid(img) - z["gpyr"]).max() <= 2e-3
    assert np.array_equal(sift.build_dog_pyramid(z["gpyr"], h, w), z["dogpyr"])
    kp = sift.find_scale_space_extrema(z["gpyr"], z["dogpyr"], h, w)
    assert len(kp) == len(z["keypoints"]) and np.array_equal(kp["octave"], z["keypoints"]["octave"])
    assert np.array_equal(kp["x"], z["keypoints"]["x"]) and np.array_equal(kp["y"], z["keypoints"]["y"])
    # Gaussian_Blur_1D (src/sift.cpp:170-217) is restated with separately rounded mul/add: bit-exact
    assert np.array_equal(sift.gaussian_blur(img, 1.6, one_d=True), z["blur1d_sigma1p6"])


@pytest.mark.parametrize("sigma", [0.9, 1.6, 1.612452, 2.771281, 4.233202, 6.196774, 9.0])
def test_gaussian_blur_entry_point(sift, oracle, sigma):
    """Gaussian_Blur (include/sift.hpp:47) for arbitrary sigma, including the rows-1 / cols-1 zero quirk (:116)."""
    rng = np.random.default_rng(int(sigma * 100))
    img = (rng.random((61, 83)) * 255).astype(np.float32)
    got, want = sift.gaussian_blur(img, sigma), oracle.f32().gaussian_blur(img, sigma)
    assert np.abs(got - want).max() <= 2e-3
    assert np.array_equal(sift.gaussian_blur(img, sigma, one_d=True), oracle.f32().gaussian_blur_1d(img, sigma))


def test_matcher_identical_indices(sift, pkg, oracle, golden):
    """knnMatch(k=2) + ratio 0.86 (src/main.cpp:25-40): indices identical to the fixture (cross-checked vs cv2)."""
    z = golden("match_query_scene")
    for norm in (pkg.NORM_L1, pkg.NORM_L2):
        idx, dist, good = sift.match_knn2(z["query_desc"], z["scene_desc"], norm, 0.86)
        assert np.array_equal(idx, z[f"idx_n{norm}"])
        assert np.allclose(dist, z[f"dist_n{norm}"], rtol=1e-6)
        assert np.array_equal(good, z[f"good_n{norm}"])
    # ties -> lowest train index; fewer than two train rows -> (-1, inf), never good
    rng = np.random.default_rng(1)
    t = rng.random((6, 128)).astype(np.float32)
    t[4] = t[1]
    idx, dist, good = sift.match_knn2(t[[1]], t)
    assert idx.tolist() == [[1, 4]] and dist[0, 0] == 0 and good[0]
    idx, dist, good = sift.match_knn2(t[[1]], t[:1])
    assert idx.tolist() == [[0, -1]] and np.isinf(dist[0, 1]) and not good[0]
    q = rng.random((300, 128)).astype(np.float32)
    tr = rng.random((257, 128)).astype(np.float32)
    for norm in (pkg.NORM_L1, pkg.NORM_L2):
        gi, gd, gg = sift.match_knn2(q, tr, norm)
        oi, od, og = oracle.match_knn2(q, tr, norm)
        assert np.array_equal(gi, oi) and np.array_equal(gg, og) and np.allclose(gd, od, rtol=1e-6)


def test_tensor_core_matcher_identical_indices(sift, pkg, oracle, golden):
    """tcgen05 L2 matcher (match_tc.cu: split-bf16 MMAs -> per train split the {min, second min} of the 4 best 32-row chunks ->
    exact fp64 re-rank of the entries whose error interval can still reach the top two; exhaustive exact match of any query
    whose shortlist is not provably complete): indices,
    distances and ratio flags identical to the fixture, to the exact kernel and to the oracle -- ragged tiles, duplicates
    (ties -> lowest train index), near-duplicates inside the bf16 error band, several train splits."""
    z = golden("match_query_scene")
    idx, dist, good = sift.match_knn2(z["query_desc"], z["scene_desc"], pkg.NORM_L2, 0.86, tensor_cores=True)
    assert np.array_equal(idx, z["idx_n4"]) and np.array_equal(good, z["good_n4"]) and np.allclose(dist, z["dist_n4"], rtol=1e-6)
    rng = np.random.default_rng(11)

    def rootsift_like(n):
        d = rng.gamma(0.6, 1.0, size=(n, 128)).astype(np.float32)
        d /= d.sum(1, keepdims=True)
        return np.sqrt(d).astype(np.float32)

    for nq, nt in [(1, 4), (3, 2), (5, 7), (128, 128), (129, 127), (300, 1000), (2000, 3000)]:
        q, t = rootsift_like(nq), rootsift_like(nt)
        if nt >= 100:
            t[3] = q[0]; t[97] = q[0]                                               # exact duplicates: tie order
            t[50] = q[1]; t[51] = q[1] + 2e-6 * rng.standard_normal(128).astype(np.float32)  # closer than the bf16-split error
            t[60:66] = q[2] + 1e-5 * rng.standard_normal((6, 128)).astype(np.float32)       # a cluster wider than the shortlist
        gi, gd, gg = sift.match_knn2(q, t, pkg.NORM_L2, 0.86, tensor_cores=True)
        ei, ed, eg = sift.match_knn2(q, t, pkg.NORM_L2, 0.86)
        oi, od, og = oracle.match_knn2(q, t, pkg.NORM_L2, 0.86)
        # t[60:66]: six rows within the bf16-split error of each other -- more than a chunk's shortlist holds; the matcher detects the
        # incomplete shortlist and matches that query exhaustively, so there is no carve-out: every query must agree
        assert np.array_equal(gi, ei) and np.array_equal(gd, ed) and np.array_equal(gg, eg)
        assert np.array_equal(gi, oi) and np.array_equal(gg, og)
    # other value ranges: raw (unnormalised) SIFT-like integers 0..255, signed data, tiny magnitudes, zero rows -- the error bound and
    # the integer sort keys of the tensor path scale with |q|^2 + |t|^2, not with an absolute constant
    for make in (lambda n: rng.integers(0, 256, size=(n, 128)).astype(np.float32),
                 lambda n: rng.standard_normal((n, 128)).astype(np.float32) * 30.0,
                 lambda n: (rng.random((n, 128)) * 1e-6).astype(np.float32)):
        q, t = make(700), make(900)
        t[17] = q[3]; t[400] = q[3]; q[9] = 0.0; t[5] = 0.0
        gi, gd, gg = sift.match_knn2(q, t, pkg.NORM_L2, 0.86, tensor_cores=True)
        ei, ed, eg = sift.match_knn2(q, t, pkg.NORM_L2, 0.86)
        assert np.array_equal(gi, ei) and np.array_equal(gd, ed) and np.array_equal(gg, eg)
    with pytest.raises(pkg.SiftError):
        sift.match_knn2(q, t, pkg.NORM_L1, 0.86, tensor_cores=True)


def test_config5_query_vs_scene_end_to_end(sift, pkg, golden):
    """BASELINE config 5 as src/main.cpp runs it: scene (1st argument, resized 960x960) and query (native 2448x2448) through
    detect+describe, then knnMatch(query, scene, 2) + ratio 0.86 (:23-40), everything on the GPU.  The matcher is index-exact on
    identical inputs (test_matcher_identical_indices); here its inputs are the GPU's own descriptors, which differ from the
    reference's by rare +-1 LSB quantisation flips, so a few near-tie rows may pick another neighbour: >= 97 % of the best
    matches and >= 95 % of the ratio-test survivors must coincide with the reference fixture (observed ~99 %)."""
    z = golden("match_query_scene")
    scene = golden("scene_960")
    kq, dq = sift.detect_describe(golden("query_2448")["gray"].astype(np.float32))
    ks, ds = sift.detect_describe(scene["gray"].astype(np.float32))
    assert len(kq) == len(z["query_kp"]) and len(ks) == len(scene["keypoints"])  # same keypoints, same order
    for norm in (pkg.NORM_L1, pkg.NORM_L2):
        idx, dist, good = sift.match_knn2(dq, ds, norm, 0.86)
        same_best = np.mean(idx[:, 0] == z[f"idx_n{norm}"][:, 0])
        ref_good = z[f"good_n{norm}"]
        jacc = np.sum(good & ref_good) / max(1, np.sum(good | ref_good))
        assert same_best >= 0.97 and jacc >= 0.95, (norm, same_best, jacc)
        # accepted matches that both agree on point to the same scene keypoint
        both = good & ref_good
        assert np.mean(idx[both, 0] == z[f"idx_n{norm}"][both, 0]) >= 0.99


def test_exact_pyramid_mode_is_bit_identical(sift, pkg, oracle, golden):
    """sift_b200_set_exact_pyramid: the Gaussian pyramid replays the reference's non-separable loop (src/sift.cpp:110-153) in its
    own summation order -> every level equals the oracle's (= the compiled reference's) BIT FOR BIT, and with that pyramid the
    whole path reproduces the reference's keypoint positions exactly (angles to 1e-3 deg) and its descriptors to the stated 1e-3:
    every row of the synthetic fixtures, >= 99.5 % of data/scene.jpg's 486 (observed: one row at 1.8e-3, a single uchar flip)."""
    sift.set_exact_pyramid(True)
    try:
        for name in ("synth_160x120", "synth_odd_211x173"):
            z = golden(name)
            img = z["image"].astype(np.float32)
            got = sift.build_gaussian_pyramid(img, 5)
            assert np.array_equal(got, z["gpyr"]), name                               # committed fixture (= compiled reference)
            assert np.array_equal(got, oracle.f32().build_gaussian_pyramid(img, 5)), name  # and the oracle run here
            kp, desc = sift.detect_describe(img)
            okp, odesc = z["keypoints"], z["descriptors"]
            assert len(kp) == len(okp)
            for fld in ("x", "y", "response", "octave"):
                assert np.array_equal(kp[fld], okp[fld]), (name, fld)
            assert np.allclose(kp["size"], okp["size"], rtol=2.5e-7, atol=0), name  # 1 ulp: exp2f here, powf in the reference (:384)
            assert np.abs(kp["angle"] - okp["angle"]).max() <= 1e-3, name
            err = np.linalg.norm(desc - odesc, axis=1)
            assert err.max() <= 1e-3, (name, float(err.max()))
        z = golden("scene_960")
        kp, desc = sift.detect_describe(z["gray"].astype(np.float32))
        assert len(kp) == len(z["keypoints"])
        assert np.array_equal(kp["x"], z["keypoints"]["x"]) and np.array_equal(kp["y"], z["keypoints"]["y"])
        assert np.abs(kp["angle"] - z["keypoints"]["angle"]).max() <= 1e-3
        err = np.linalg.norm(desc - z["descriptors"], axis=1)
        assert (err <= 1e-3).mean() >= 0.995, float((err <= 1e-3).mean())
    finally:
        sift.set_exact_pyramid(False)


def test_config5_device_resident(sift, pkg, golden):
    """Config 5 without leaving the device: query and scene through the batch entry point, their descriptor buffers straight into
    the device matcher (exact L1, exact L2, tensor-core L2) -- same answers as the host entry point on the same descriptors."""
    import torch

    cap = 4096
    st = torch.cuda.current_stream().cuda_stream
    descs = []
    for name in ("query_2448", "scene_960"):
        img = torch.from_numpy(golden(name)["gray"].astype(np.float32)[None]).cuda()
        d_kp = torch.zeros((1, cap, 28), dtype=torch.uint8, device="cuda")
        d_desc = torch.zeros((1, cap, 128), dtype=torch.float32, device="cuda")
        d_cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
        sift.detect_describe_batch_dev(img, d_kp, d_desc, d_cnt, cap, st)
        n = int(d_cnt[0])
        assert 0 < n <= cap
        descs.append(d_desc[0, :n].contiguous())
    dq, ds = descs
    for norm, tc in ((pkg.NORM_L1, False), (pkg.NORM_L2, False), (pkg.NORM_L2, True)):
        d_idx = torch.zeros((len(dq), 2), dtype=torch.int32, device="cuda")
        d_dist = torch.zeros((len(dq), 2), dtype=torch.float32, device="cuda")
        sift.match_knn2_dev(dq, ds, d_idx, d_dist, norm, tensor_cores=tc, stream=st)
        torch.cuda.synchronize()
        idx, dist, good = sift.match_knn2(dq.cpu().numpy(), ds.cpu().numpy(), norm, 0.86)
        assert np.array_equal(d_idx.cpu().numpy(), idx) and np.array_equal(d_dist.cpu().numpy(), dist), (norm, tc)
    assert good.sum() > 50  # the book cover is found in the scene
