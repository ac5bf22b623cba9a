"""The C oracle (oracle/sift_oracle.c) against the golden fixtures produced by oracle/_ref, i.e. by the UNMODIFIED
reference src/sift.cpp (tests/golden/make_golden.py).  Bit-exact: same arithmetic, same order."""
import numpy as np
import pytest


@pytest.mark.parametrize("name", ["synth_160x120", "synth_odd_211x173"])
def test_every_stage_bit_exact(oracle, golden, name):
    z = golden(name)
    img = z["image"]
    rows, cols = img.shape
    o = oracle.f32()
    g = o.build_gaussian_pyramid(img)
    assert np.array_equal(g, z["gpyr"])  # buildGaussianPyramid, src/sift.cpp:229-263
    d = o.build_dog_pyramid(g, rows, cols)
    assert np.array_equal(d, z["dogpyr"])  # buildDoGPyramid, :265-283
    kps = o.find_scale_space_extrema(g, d, rows, cols)
    assert kps.tobytes() == z["keypoints"].tobytes()  # findScaleSpaceExtrema, :547-577 (order included)
    desc = o.cal_descriptor(g, rows, cols, kps)
    assert np.array_equal(desc, z["descriptors"])  # calDescriptor, :733-753
    assert np.array_equal(o.gaussian_blur_1d(img, 1.6), z["blur1d_sigma1p6"])  # Gaussian_Blur_1D, :170-217
    k2, d2 = o.sift_ncl(img)
    assert k2.tobytes() == kps.tobytes() and np.array_equal(d2, desc)  # SIFT_NCL, :59-91


def test_scene_960_config1(oracle, golden):
    """BASELINE config 1 as src/main.cpp feeds it: data/scene.jpg -> 960x960 gray f32."""
    z = golden("scene_960")
    kps, desc = oracle.f32().sift_ncl(z["gray"].astype(np.float32))
    assert len(kps) == 486
    assert kps.tobytes() == z["keypoints"].tobytes()
    assert np.array_equal(desc, z["descriptors"])
    assert np.allclose(np.linalg.norm(desc, axis=1), 1.0, atol=1e-5)  # sqrt(q/sum q): unit L2 rows


@pytest.mark.parametrize("name,desc_of", [("scene_960_prequant", ("scene_960", "descriptors")), ("query_2448_prequant", ("match_query_scene", "query_desc")),
                                          ("scene_native_2048x1280", ("scene_native_2048x1280", "descriptors"))])
def test_prequant_fixtures_reproduce_the_reference_descriptors(golden, name, desc_of):
    """The pre-quantisation vectors come from the C port (the reference does not expose them): they are only legitimate if the
    reference's own tail applied to them -- saturate_cast<uchar> (round half even, clamp), L1 normalise, sqrt (src/sift.cpp:704-721)
    -- gives back the descriptors the UNMODIFIED reference wrote into the fixture, bit for bit."""
    pq = golden(name)["prequant"]
    want = golden(desc_of[0])[desc_of[1]]
    assert pq.shape == want.shape and pq.dtype == np.float32
    q = np.clip(np.rint(pq), 0, 255).astype(np.float32)  # np.rint rounds half to even like cvRound
    nrm1 = q.sum(axis=1, keepdims=True, dtype=np.float32)
    got = np.sqrt(q / np.maximum(nrm1, np.float32(1.1920929e-07)))
    assert np.abs(got - want).max() <= 2e-7  # the reference multiplies by 1/sum where this divides: 1 ulp


def test_scene_native_config1(oracle, golden):
    """BASELINE config 1 at native size (2048x1280, no resize): the port reproduces the unmodified reference's keypoints exactly."""
    z = golden("scene_native_2048x1280")
    kps, desc = oracle.f32().sift_ncl(z["gray"].astype(np.float32))
    assert len(kps) == 1364 and kps.tobytes() == z["keypoints"].tobytes()
    assert np.array_equal(desc, z["descriptors"])


def test_matcher_config5(oracle, golden):
    """knnMatch(query, scene, 2) + ratio 0.86 (src/main.cpp:25-40); the fixture was cross-checked against cv2.BFMatcher."""
    z = golden("match_query_scene")
    for norm in (oracle.NORM_L1, oracle.NORM_L2):
        idx, dist, good = oracle.match_knn2(z["query_desc"], z["scene_desc"], norm, 0.86)
        assert np.array_equal(idx, z[f"idx_n{norm}"])
        assert np.array_equal(dist, z[f"dist_n{norm}"])
        assert np.array_equal(good, z[f"good_n{norm}"])


def test_matcher_ties_and_short_train(oracle):
    rng = np.random.default_rng(1)
    t = rng.random((6, 128)).astype(np.float32)
    t[4] = t[1]  # exact duplicate: ties resolve to the lowest train index
    q = t[[1]].copy()
    idx, dist, good = oracle.match_knn2(q, t)
    assert idx.tolist() == [[1, 4]] and dist[0, 0] == 0 and dist[0, 1] == 0 and good[0]
    idx, dist, good = oracle.match_knn2(q, t[:1])  # fewer than k train rows: no second match, never "good"
    assert idx.tolist() == [[0, -1]] and not good[0]


def test_fp64_twin_agrees_loosely(oracle, golden):
    """The fp64 twin follows the same algorithm; it must find nearly the same keypoints as the fp32 oracle."""
    z = golden("synth_odd_211x173")
    k32, _ = oracle.f32().sift_ncl(z["image"])
    k64, _ = oracle.f64().sift_ncl(z["image"])
    a = {(int(k["octave"]) & 0xFFFF, round(float(k["x"]), 1), round(float(k["y"]), 1)) for k in k32}
    b = {(int(k["octave"]) & 0xFFFF, round(float(k["x"]), 1), round(float(k["y"]), 1)) for k in k64}
    assert len(a & b) >= 0.95 * max(len(a), len(b))
