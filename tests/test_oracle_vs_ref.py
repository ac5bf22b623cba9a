"""The C oracle against oracle/_ref (the unmodified reference compiled against third_party/cvshim), run live on seeded
inputs including the edge cases the reference admits.  Skipped when oracle/_ref was not built (no /root/reference)."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def ref(oracle):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built")
    return oracle.ref()


@pytest.mark.parametrize("shape,seed", [((16, 16), 0), ((33, 47), 1), ((64, 200), 2), ((131, 97), 3), ((240, 320), 4)])
def test_sift_ncl_bit_exact(oracle, ref, synth, shape, seed):
    h, w = shape
    img = synth.recipe_s(w, h, seed=seed, blobs_per_1080p=30000)
    ko, do = oracle.f32().sift_ncl(img)
    kr, dr = ref.sift_ncl(img)
    assert ko.tobytes() == kr.tobytes()
    assert np.array_equal(do, dr)


def test_constant_and_zero_images(oracle, ref):
    for v in (0.0, 128.0):
        img = np.full((48, 64), v, dtype=np.float32)
        ko, do = oracle.f32().sift_ncl(img)
        kr, dr = ref.sift_ncl(img)
        assert len(ko) == len(kr) and ko.tobytes() == kr.tobytes() and np.array_equal(do, dr)


@pytest.mark.parametrize("sigma", [0.5, 1.0, 1.6, 1.612452, 2.771281, 4.233202, 6.196774])
def test_blur_fast_equals_naive_equals_ref(oracle, ref, sigma):
    rng = np.random.default_rng(int(sigma * 1000))
    img = (rng.random((37, 53)) * 255).astype(np.float32)
    o = oracle.f32()
    a, b, c = o.gaussian_blur(img, sigma), o.gaussian_blur_naive(img, sigma), ref.gaussian_blur(img, sigma)
    assert np.array_equal(a, b)
    assert np.array_equal(a, c)
    assert np.array_equal(o.gaussian_blur_1d(img, sigma), ref.gaussian_blur(img, sigma, one_d=True))


def test_stage_functions_bit_exact(oracle, ref, synth):
    img = synth.recipe_s(150, 110, seed=21, blobs_per_1080p=30000)
    h, w = img.shape
    o = oracle.f32()
    g = ref.build_gaussian_pyramid(img)
    assert np.array_equal(o.build_gaussian_pyramid(img), g)
    d = ref.build_dog_pyramid(g, h, w)
    assert np.array_equal(o.build_dog_pyramid(g, h, w), d)
    kr = ref.find_scale_space_extrema(g, d, h, w)
    ko = o.find_scale_space_extrema(g, d, h, w)
    assert len(kr) > 10 and ko.tobytes() == kr.tobytes()
    assert np.array_equal(o.cal_descriptor(g, h, w, ko), ref.cal_descriptor(g, h, w, kr))


def test_multi_threaded_oracle_is_deterministic(oracle, synth):
    img = synth.recipe_s(200, 150, seed=33, blobs_per_1080p=30000)
    oracle.set_threads(1)
    k1, d1 = oracle.f32().sift_ncl(img)
    oracle.set_threads(8)
    k8, d8 = oracle.f32().sift_ncl(img)
    assert k1.tobytes() == k8.tobytes() and np.array_equal(d1, d8)
