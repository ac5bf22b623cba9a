"""Generate the golden fixtures in tests/golden/ from oracle/_ref (the UNMODIFIED reference src/sift.cpp
compiled against oracle/cvshim) -- run in the build container, where /root/reference exists:

    python tests/golden/make_golden.py

Inputs follow the reference driver (src/main.cpp:79-87): imread (BGR u8) -> optional resize to 960x960 (first CLI
argument only, :83) -> cvtColor(COLOR_RGB2GRAY) applied to BGR data -> convertTo(CV_32FC1) without scaling.
cv2 (the 4.13 wheel) does the decoding/colour conversion; the u8 gray image is stored so the tests never need
/root/reference or cv2.  Everything numeric in the fixtures comes from oracle/_ref, not from the C oracle.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge

O = ge.load_oracle()
from importlib import import_module

ge.load_package()
synth = import_module("sift_gpu_b200.synth")
HERE = os.path.dirname(os.path.abspath(__file__))
DATA = "/root/reference/data"


def read_image(name, resized):
    import cv2

    img = cv2.imread(os.path.join(DATA, name))
    assert img is not None
    if resized:
        img = cv2.resize(img, (960, 960))
    return cv2.cvtColor(img, cv2.COLOR_RGB2GRAY)  # u8; convertTo(CV_32FC1) is a plain cast


def main():
    O.build()
    ref = O.ref()
    # 1. small synthetic frames: every intermediate of the reference (gpyr, dogpyr, keypoints, descriptors)
    for tag, (w, h, seed) in {"synth_160x120": (160, 120, 1), "synth_odd_211x173": (211, 173, 2)}.items():
        img = synth.recipe_s(w, h, seed=seed, blobs_per_1080p=20000)
        g = ref.build_gaussian_pyramid(img)
        d = ref.build_dog_pyramid(g, h, w)
        kps = ref.find_scale_space_extrema(g, d, h, w)
        desc = ref.cal_descriptor(g, h, w, kps)
        k2, d2 = ref.sift_ncl(img)
        assert k2.tobytes() == kps.tobytes() and np.array_equal(d2, desc)
        blur1d = ref.gaussian_blur(img, 1.6, one_d=True)
        np.savez_compressed(os.path.join(HERE, tag + ".npz"), image=img, gpyr=g.astype(np.float16 if False else np.float32), dogpyr=d, keypoints=kps,
                            descriptors=desc, blur1d_sigma1p6=blur1d)
        print(tag, "keypoints", len(kps))
    # 2. config 1 / 5: data/scene.jpg as main.cpp feeds it (960x960) and data/query.jpg native
    scene = read_image("scene.jpg", True)
    ks, ds = ref.sift_ncl(scene.astype(np.float32))
    np.savez_compressed(os.path.join(HERE, "scene_960.npz"), gray=scene, keypoints=ks, descriptors=ds)
    print("scene_960 keypoints", len(ks))
    query = read_image("query.jpg", False)
    kq, dq = ref.sift_ncl(query.astype(np.float32))
    print("query native keypoints", len(kq))
    # pre-quantisation descriptor vectors (the float value each component has just before saturate_cast<uchar>, src/sift.cpp:709):
    # the unmodified reference does not expose them, so they come from the C port -- legitimate only because the port's keypoints and
    # descriptors are bit-identical to oracle/_ref's on the very same image, which is asserted here before anything is written.
    port = O.f32()
    for tag, img, k_ref, d_ref in (("scene_960", scene, ks, ds), ("query_2448", query, kq, dq)):
        kp_p, d_p, _, _, pq = port.sift_ncl(img.astype(np.float32), want_pyramids=True, want_prequant=True)
        assert kp_p.tobytes() == k_ref.tobytes() and np.array_equal(d_p, d_ref), tag
        np.savez_compressed(os.path.join(HERE, tag + "_prequant.npz"), prequant=pq.astype(np.float32))
    # config 1, native size: data/scene.jpg 2048x1280 without the driver's resize
    native = read_image("scene.jpg", False)
    kn, dn = ref.sift_ncl(native.astype(np.float32))
    kp_p, d_p, _, _, pq = port.sift_ncl(native.astype(np.float32), want_pyramids=True, want_prequant=True)
    assert kp_p.tobytes() == kn.tobytes() and np.array_equal(d_p, dn)
    np.savez_compressed(os.path.join(HERE, "scene_native_2048x1280.npz"), gray=native, keypoints=kn, descriptors=dn, prequant=pq.astype(np.float32))
    print("scene native keypoints", len(kn))
    # knnMatch(query descriptors, scene descriptors) per main.cpp:25-27, cross-checked against cv2.BFMatcher below
    import cv2

    out = {}
    for norm, cvn in ((O.NORM_L1, cv2.NORM_L1), (O.NORM_L2, cv2.NORM_L2)):
        idx, dist, good = O.match_knn2(dq, ds, norm, 0.86)
        m = cv2.BFMatcher(cvn).knnMatch(dq, ds, 2)
        cv_idx = np.array([[a.trainIdx, b.trainIdx] for a, b in m], dtype=np.int32)
        cv_dist = np.array([[a.distance, b.distance] for a, b in m], dtype=np.float32)
        assert np.array_equal(cv_idx, idx), "oracle matcher disagrees with cv2.BFMatcher"
        assert np.allclose(cv_dist, dist, rtol=1e-5)
        out[f"idx_n{norm}"], out[f"dist_n{norm}"], out[f"good_n{norm}"] = idx, dist, good
        print("norm", norm, "good matches", int(good.sum()))
    np.savez_compressed(os.path.join(HERE, "match_query_scene.npz"), query_desc=dq, scene_desc=ds, query_kp=kq, **out)
    np.savez_compressed(os.path.join(HERE, "query_2448.npz"), gray=query)


def homography_fixture():
    """The driver's homography consumer (src/main.cpp:44-62) on the committed matcher fixture: obj = query keypoints of the L1 ratio-test
    survivors, scene = their matches; cv2.findHomography(obj, scene, cv2.RANSAC) (default threshold 3) is the expected value."""
    import cv2

    z = np.load(os.path.join(HERE, "match_query_scene.npz"))
    sc = np.load(os.path.join(HERE, "scene_960.npz"))
    good = z["good_n2"]
    q = z["query_kp"][good]
    t = sc["keypoints"][z["idx_n2"][good, 0]]
    obj = np.stack([q["x"], q["y"]], axis=1).astype(np.float32)
    scene = np.stack([t["x"], t["y"]], axis=1).astype(np.float32)
    H, mask = cv2.findHomography(obj, scene, cv2.RANSAC)
    corners = np.array([[0, 0], [2448, 0], [2448, 2448], [0, 2448]], dtype=np.float32).reshape(-1, 1, 2)
    proj = cv2.perspectiveTransform(corners, H).reshape(-1, 2)
    out = dict(obj=obj, scene=scene, H=H, mask=mask.ravel().astype(np.uint8), corners=proj)
    print("homography fixture (driver data):", len(obj), "matches,", int(mask.sum()), "inliers; corners", proj.round(2).tolist())
    # The driver's own image pair gives cv2 a consensus of ~10 of 118 matches (the reference's matcher finds few true pairs), which pins
    # nothing.  Second case with a known answer: 400 points under a real perspective map, 0.5 px noise, 40 % gross outliers.
    rng = np.random.default_rng(77)
    Ht = np.array([[0.36, -0.05, 120.0], [0.04, 0.33, 90.0], [2.0e-5, -1.5e-5, 1.0]])
    a = rng.uniform(0, 2448, size=(400, 2))
    w = a @ Ht[2, :2] + Ht[2, 2]
    b = (a @ Ht[:2, :2].T + Ht[:2, 2]) / w[:, None] + rng.normal(0, 0.5, size=(400, 2))
    bad = rng.random(400) < 0.4
    b[bad] = rng.uniform(0, 960, size=(int(bad.sum()), 2))
    a32, b32 = a.astype(np.float32), b.astype(np.float32)
    H2, m2 = cv2.findHomography(a32, b32, cv2.RANSAC)
    p2 = cv2.perspectiveTransform(corners, H2).reshape(-1, 2)
    out.update(syn_obj=a32, syn_scene=b32, syn_H=H2, syn_mask=m2.ravel().astype(np.uint8), syn_corners=p2, syn_H_true=Ht)
    print("homography fixture (synthetic):", int(m2.sum()), "inliers of 400 (", int((~bad).sum()), "true ); corners", p2.round(2).tolist())
    np.savez_compressed(os.path.join(HERE, "homography_query_scene.npz"), **out)


if __name__ == "__main__":
    if "--only-homography" in sys.argv:
        homography_fixture()
    else:
        main()
        homography_fixture()
