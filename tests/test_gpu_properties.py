"""GPU tests at BASELINE.json's full sizes through size-independent properties, the batch entry points, and the
edge cases / error behaviour of the C ABI.  Run with -m gpu on a B200."""
import numpy as np
import pytest

import parity

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def frame1080(synth):
    return synth.recipe_s(1920, 1080, seed=1234)


def _scan_key(kps):
    """Reference output order (src/sift.cpp:556-557,487-491): octave ascending is checkable from the record alone."""
    return (kps["octave"] & 255).astype(np.int64)


def test_1080p_properties(sift, frame1080):
    kp, desc = sift.detect_describe(frame1080)
    assert 2500 <= len(kp) <= 4500  # SURVEY: ~3.5 k keypoints on the seed-1234 frame
    assert np.all(np.diff(_scan_key(kp)) >= 0)  # octaves appear in scan order
    layer = (kp["octave"] >> 8) & 255
    assert set(np.unique(layer)) <= {1, 2}  # nOctaveLayers = 2: extrema only on DoG layers 1..2
    assert np.all((kp["angle"] >= 0) & (kp["angle"] < 360)) and np.all(kp["class_id"] == -1)
    assert np.all(kp["response"] * 2 >= 0.04)  # contrast test, src/sift.cpp:365
    assert np.allclose(np.linalg.norm(desc, axis=1), 1.0, atol=1e-5) and desc.min() >= 0 and desc.max() <= 1
    # idempotence / determinism: the path has no order-dependent float atomics
    kp2, desc2 = sift.detect_describe(frame1080)
    assert kp.tobytes() == kp2.tobytes() and np.array_equal(desc, desc2)


def test_1080p_vs_oracle(sift, oracle, frame1080):
    """BASELINE config 2 at full size: the oracle needs ~1 s for this frame with all cores."""
    okp, odesc, _, _, opq = oracle.f32().sift_ncl(frame1080, want_pyramids=True, want_prequant=True)
    kp, desc = sift.detect_describe(frame1080)
    r = parity.full_report(sift, frame1080, kp, desc, okp, odesc, opq)
    assert r["kp_recall"] >= 0.99 and r["kp_precision"] >= 0.99, r
    assert r["frac_within_1e-3"] >= 0.95 and r["unexplained"] == 0 and r["explained_by_keypoint"] <= r["matched"] // 100, r


def test_batch_dev_equals_single_frames(pkg, synth):
    """Config 4 in miniature: a device-resident batch (chunked internally) gives, per frame, exactly what the
    one-image entry point gives; frames do not interact."""
    import torch

    frames = np.stack([synth.recipe_s(640, 360, seed=100 + k) for k in range(5)])
    cap = 4096
    s = pkg.Sift(360, 640, max_batch=2, max_kp_per_frame=cap)  # 5 frames in chunks of 2 -> ragged last chunk
    d = torch.from_numpy(frames).cuda()
    d_kp = torch.zeros((5, cap, 28), dtype=torch.uint8, device="cuda")
    d_desc = torch.zeros((5, cap, 128), dtype=torch.float32, device="cuda")
    d_cnt = torch.zeros(5, dtype=torch.int32, device="cuda")
    s.detect_describe_batch_dev(d, d_kp, d_desc, d_cnt, cap, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    cnt = d_cnt.cpu().numpy()
    for f in range(5):
        kp, desc = s.detect_describe(frames[f])
        assert cnt[f] == len(kp) > 50
        got_kp = d_kp[f, : cnt[f]].cpu().numpy().view(pkg.KP_DTYPE).ravel()
        assert got_kp.tobytes() == kp.tobytes()
        assert np.array_equal(d_desc[f, : cnt[f]].cpu().numpy(), desc)
    # host-buffer batch entry point: same answers again
    h_kp = np.zeros((5, cap), dtype=pkg.KP_DTYPE); h_desc = np.zeros((5, cap, 128), dtype=np.float32); h_cnt = np.zeros(5, dtype=np.int32)
    assert s.detect_describe_batch_host(frames, h_kp, h_desc, h_cnt, cap) == pkg.OK
    assert np.array_equal(h_cnt, cnt)
    for f in range(5):
        assert np.array_equal(h_desc[f, : cnt[f]], d_desc[f, : cnt[f]].cpu().numpy())
    # u8 front end (src/main.cpp:84-85: gray u8 -> float32 unscaled) == float path on the same integer-valued frames
    u8 = np.clip(np.rint(frames), 0, 255).astype(np.uint8)
    d8 = torch.from_numpy(u8).cuda()
    d_cnt8 = torch.zeros(5, dtype=torch.int32, device="cuda")
    d_desc8 = torch.zeros_like(d_desc)
    s.detect_describe_batch_dev(d8, d_kp, d_desc8, d_cnt8, cap, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    kpf, descf = s.detect_describe(u8[0].astype(np.float32))
    assert int(d_cnt8[0]) == len(kpf) and np.array_equal(d_desc8[0, : len(kpf)].cpu().numpy(), descf)
    # ... and the uint8 host-batch entry point
    import ctypes as C
    h_cnt8 = np.zeros(5, dtype=np.int32)
    rc = pkg.lib().sift_b200_detect_describe_batch_host_u8(s._h, u8.ctypes.data_as(C.c_void_p), 5, 360, 640, h_kp.ctypes.data_as(C.c_void_p),
                                                           h_desc.ctypes.data_as(C.c_void_p), h_cnt8.ctypes.data_as(C.c_void_p), cap)
    assert rc == pkg.OK and np.array_equal(h_cnt8, d_cnt8.cpu().numpy()) and np.array_equal(h_desc[0, : len(kpf)], descf)
    s.close()


def test_host_batch_tapered_schedule(pkg, synth):
    """The host batch pipeline (three input buffers, two compute lanes, short chunks at both ends of a call: api.cu chunk_plan)
    returns for every frame exactly what the device-resident call returns: 13 frames with max_batch 4 -> chunks 1,2,3,4,2,1;
    6 frames (below the taper threshold) -> 2,4; and a capacity overflow in one frame is reported without disturbing the others."""
    import torch

    n = 13
    frames = np.stack([synth.recipe_s(320, 200, seed=300 + k) for k in range(n)])
    cap = 2048
    s = pkg.Sift(200, 320, max_batch=4, max_kp_per_frame=cap)
    d = torch.from_numpy(frames).cuda()
    d_kp = torch.zeros((n, cap, 28), dtype=torch.uint8, device="cuda")
    d_desc = torch.zeros((n, cap, 128), dtype=torch.float32, device="cuda")
    d_cnt = torch.zeros(n, dtype=torch.int32, device="cuda")
    s.detect_describe_batch_dev(d, d_kp, d_desc, d_cnt, cap, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    cnt = d_cnt.cpu().numpy()
    assert cnt.min() > 20
    for m in (n, 6, 1):
        h_kp = np.zeros((m, cap), dtype=pkg.KP_DTYPE); h_desc = np.zeros((m, cap, 128), dtype=np.float32); h_cnt = np.zeros(m, dtype=np.int32)
        assert s.detect_describe_batch_host(frames[:m], h_kp, h_desc, h_cnt, cap) == pkg.OK
        assert np.array_equal(h_cnt, cnt[:m])
        for f in range(m):
            assert h_kp[f, : cnt[f]].tobytes() == d_kp[f, : cnt[f]].cpu().numpy().tobytes()
            assert np.array_equal(h_desc[f, : cnt[f]], d_desc[f, : cnt[f]].cpu().numpy())
    small = int(cnt.min()) - 1  # at least one frame overflows this capacity
    h_kp = np.zeros((n, small), dtype=pkg.KP_DTYPE); h_desc = np.zeros((n, small, 128), dtype=np.float32); h_cnt = np.zeros(n, dtype=np.int32)
    assert s.detect_describe_batch_host(frames, h_kp, h_desc, h_cnt, small) == pkg.ERR_CAPACITY
    assert np.array_equal(h_cnt, cnt)  # true counts
    for f in range(n):
        assert np.array_equal(h_desc[f], d_desc[f, :small].cpu().numpy())  # the first `small` keypoints of every frame
    s.close()


def test_edge_cases_and_errors(pkg, oracle, synth):
    s = pkg.Sift(256, 256, max_batch=1, max_kp_per_frame=2048)
    # smallest image the reference admits (16 px: octave 4 is 1x1) -> no keypoints, like the oracle
    img = synth.recipe_s(16, 16, seed=1)
    kp, desc = s.detect_describe(img)
    assert len(kp) == len(oracle.f32().sift_ncl(img)[0]) == 0 and desc.shape == (0, 128)
    # constant image: no extrema
    assert len(s.detect_describe(np.full((64, 80), 128.0, np.float32))[0]) == 0
    # below 16 px the reference throws inside cv::resize (src/sift.cpp:254) -> explicit status here
    with pytest.raises(pkg.SiftError) as e:
        s.detect_describe(np.zeros((12, 40), np.float32))
    assert e.value.code == pkg.ERR_TOO_SMALL
    with pytest.raises(pkg.SiftError) as e:
        s.detect_describe(np.zeros((300, 300), np.float32))  # larger than the handle's workspace
    assert e.value.code == pkg.ERR_ARG
    # non-continuous input (row stride > cols*4) is honoured, unlike the reference which assumes continuous (:111)
    big = synth.recipe_s(256, 200, seed=3, blobs_per_1080p=20000)
    view = big[:, :180]
    k1, d1 = s.detect_describe(view)
    k2, d2 = s.detect_describe(np.ascontiguousarray(view))
    assert len(k1) > 20 and k1.tobytes() == k2.tobytes() and np.array_equal(d1, d2)
    # capacity overflow: status CAPACITY, true count reported, outputs = prefix of the full result
    full_k, full_d = s.detect_describe(big)
    cap = len(full_k) // 2
    import ctypes as C
    kps = np.zeros(cap, dtype=pkg.KP_DTYPE); desc = np.zeros((cap, 128), np.float32); n = C.c_int(0)
    rc = pkg.lib().sift_b200_detect_describe(s._h, big.ctypes.data_as(C.c_void_p), big.shape[0], big.shape[1], C.c_size_t(0),
                                             kps.ctypes.data_as(C.c_void_p), desc.ctypes.data_as(C.c_void_p), cap, C.byref(n))
    assert rc == pkg.ERR_CAPACITY and n.value == len(full_k)
    assert kps.tobytes() == full_k[:cap].tobytes() and np.array_equal(desc, full_d[:cap])
    # calDescriptor's CV_Assert (src/sift.cpp:744): layer > nOctaveLayers+2 -> ASSERT status
    g = s.build_gaussian_pyramid(big)
    bad = full_k[:1].copy()
    bad["octave"] = 0 | (7 << 8)
    with pytest.raises(pkg.SiftError) as e:
        s.cal_descriptor(g, big.shape[0], big.shape[1], bad)
    assert e.value.code == pkg.ERR_ASSERT
    assert s.cal_descriptor(g, big.shape[0], big.shape[1], full_k[:0]).shape == (0, 128)  # empty keypoint list
    # caller-supplied keypoints whose window misses the image entirely (no sample passes 0 < r < rows-1, 0 < c < cols-1, :621): the
    # reference leaves the histogram empty and returns an all-zero row; so does the oracle, so must the kernel (no stale shared memory)
    off = full_k[:3].copy()
    off["x"] = [-500.0, big.shape[1] + 400.0, 10.0]
    off["y"] = [20.0, big.shape[0] + 300.0, -700.0]
    got = s.cal_descriptor(g, big.shape[0], big.shape[1], np.concatenate([full_k[:5], off, full_k[5:9]]))
    want = oracle.f32().cal_descriptor(g, big.shape[0], big.shape[1], np.concatenate([full_k[:5], off, full_k[5:9]]))
    assert np.array_equal(got[5:8], np.zeros((3, 128), np.float32)) and np.array_equal(want[5:8], got[5:8])
    assert (np.linalg.norm(got - want, axis=1) <= 1e-3).mean() >= 0.8  # the in-image rows around them are untouched by the empty ones
    s.close()


def test_two_host_threads_two_handles(pkg, synth):
    """The threading contract of include/sift_b200.h: different handles are independent, so two host threads, each driving its own handle
    on its own half of a batch (what bench.py's e2e does), must return exactly what one call over the whole batch returns."""
    import threading

    frames = np.stack([synth.recipe_s(480, 270, seed=300 + k) for k in range(12)])
    cap = 2048
    one = pkg.Sift(270, 480, max_batch=2, max_kp_per_frame=cap)
    w_kp = np.zeros((12, cap), dtype=pkg.KP_DTYPE); w_desc = np.zeros((12, cap, 128), dtype=np.float32); w_cnt = np.zeros(12, dtype=np.int32)
    assert one.detect_describe_batch_host(frames, w_kp, w_desc, w_cnt, cap) == pkg.OK
    one.close()
    assert w_cnt.min() > 20
    handles = [pkg.Sift(270, 480, max_batch=2, max_kp_per_frame=cap) for _ in range(2)]
    g_kp = np.zeros_like(w_kp); g_desc = np.zeros_like(w_desc); g_cnt = np.zeros_like(w_cnt)
    errs = []

    def run(h, lo, hi):
        try:
            for _ in range(3):  # repeated calls: the two pipelines drift against each other
                assert h.detect_describe_batch_host(frames[lo:hi], g_kp[lo:hi], g_desc[lo:hi], g_cnt[lo:hi], cap) == pkg.OK
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=run, args=(handles[0], 0, 6)), threading.Thread(target=run, args=(handles[1], 6, 12))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for h in handles:
        h.close()
    assert not errs, errs
    assert np.array_equal(g_cnt, w_cnt)
    for f in range(12):
        assert g_kp[f, : w_cnt[f]].tobytes() == w_kp[f, : w_cnt[f]].tobytes()
        assert np.array_equal(g_desc[f, : w_cnt[f]], w_desc[f, : w_cnt[f]])


def test_internal_list_overflow_contract(pkg, synth):
    """A frame with more refined extrema than the handle's max_kp_per_frame: the true keypoint count is unknown, so the call reports
    max_kp + 1 with ERR_CAPACITY, writes only the records the kernels produced (all of them valid, in reference order) and leaves the
    rest of the caller's buffers untouched (include/sift_b200.h)."""
    import ctypes as C

    img = synth.recipe_s(640, 360, seed=100)
    big = pkg.Sift(360, 640, max_batch=1, max_kp_per_frame=4096)
    kp_all, desc_all = big.detect_describe(img)
    big.close()
    assert len(kp_all) > 200
    cap = 64
    s = pkg.Sift(360, 640, max_batch=1, max_kp_per_frame=cap)
    kps = np.zeros(cap + 8, dtype=pkg.KP_DTYPE)
    desc = np.full((cap + 8, 128), -7.0, dtype=np.float32)
    n = C.c_int(0)
    rc = pkg.lib().sift_b200_detect_describe(s._h, img.ctypes.data_as(C.c_void_p), 360, 640, C.c_size_t(640 * 4), kps.ctypes.data_as(C.c_void_p),
                                             desc.ctypes.data_as(C.c_void_p), cap, C.byref(n))
    assert rc == pkg.ERR_CAPACITY and n.value == cap + 1
    written = int((desc[:, 0] != -7.0).sum())
    assert 0 < written <= cap and np.all(desc[written:] == -7.0)
    # every written row is a real descriptor of this image (unit norm) belonging to one of its keypoints
    assert np.allclose(np.linalg.norm(desc[:written], axis=1), 1.0, atol=1e-5)
    all_xy = {(float(k["x"]), float(k["y"])) for k in kp_all}
    assert all((float(k["x"]), float(k["y"])) in all_xy for k in kps[:written])
    s.close()


def test_launches_are_counted(pkg, synth):
    s = pkg.Sift(128, 128, max_batch=1, max_kp_per_frame=1024)
    n0 = s.launch_count()
    s.detect_describe(synth.recipe_s(128, 128, seed=4))
    assert s.launch_count() - n0 == 13  # base blur, 5 octaves, gradient maps, extrema scan, refine, orientation, order+scan, descriptor prep, descriptors
    s.close()


def test_upsample_front_end_and_config3(pkg, oracle, synth):
    """BASELINE config 3: synthetic 3840x2160 frame (recipe S, 24 000 blobs), 2x bilinear upsample to 7680x4320, then the
    unchanged pipeline.  The reference has no upsample path (src/sift.cpp:219-227 ignores doubleSize), so the checker is
    the numpy restatement of cv::resize(INTER_LINEAR) (tests/test_host_logic.py pins it to cv2) followed by the CPU oracle.
    Keypoint coordinates are in upsampled pixels."""
    s = pkg.Sift(4320, 7680, max_batch=1, max_kp_per_frame=1 << 17)
    small = synth.recipe_s(300, 200, seed=8, blobs_per_1080p=20000)
    kp, desc, up = s.detect_describe_up2(small, want_upsampled=True)
    want_up = synth.upsample2x(small)
    assert np.array_equal(up, want_up)  # separately rounded mul/add: bit-exact against the restatement
    okp, odesc = oracle.f32().sift_ncl(want_up)
    pairs = parity.match_keypoints(kp, okp)
    rec, prec = parity.recall_precision(pairs, len(kp), len(okp))
    assert len(okp) > 100 and rec >= 0.99 and prec >= 0.99
    # full size
    frame = synth.recipe_s(3840, 2160, seed=1234)
    kp, desc, up = s.detect_describe_up2(frame, want_upsampled=True)
    want_up = synth.upsample2x(frame)
    assert up.shape == (4320, 7680) and np.array_equal(up, want_up)
    okp, odesc, _, _, opq = oracle.f32().sift_ncl(want_up, want_pyramids=True, want_prequant=True)
    pairs = parity.match_keypoints(kp, okp)
    rec, prec = parity.recall_precision(pairs, len(kp), len(okp))
    assert len(okp) > 5000 and rec >= 0.99 and prec >= 0.99, (len(kp), len(okp), rec, prec)
    r = parity.full_report(s, want_up, kp, desc, okp, odesc, opq)
    assert r["frac_within_1e-3"] >= 0.95 and r["unexplained"] == 0 and r["explained_by_keypoint"] <= r["matched"] // 100, r
    s.close()


def test_driver_colour_front_end(pkg, synth):
    """src/main.cpp:84 applies COLOR_RGB2GRAY to BGR bytes; the GPU front end reproduces cv2's fixed point
    exactly and feeds the u8 pipeline."""
    import torch

    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    base = np.clip(np.rint(synth.recipe_s(320, 200, seed=6, blobs_per_1080p=20000)), 0, 255).astype(np.uint8)
    bgr = np.stack([base, np.roll(base, 3, 1), rng.integers(0, 256, base.shape, dtype=np.uint8)], axis=-1)[None]
    want = cv2.cvtColor(bgr[0], cv2.COLOR_RGB2GRAY)
    s = pkg.Sift(200, 320, max_batch=1, max_kp_per_frame=4096)
    d_bgr = torch.from_numpy(bgr).cuda()
    d_gray = torch.zeros((1, 200, 320), dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    s.rgb2gray_u8_dev(d_bgr, d_gray, st)
    torch.cuda.synchronize()
    assert np.array_equal(d_gray[0].cpu().numpy(), want)
    cap = 4096
    d_kp = torch.zeros((1, cap, 28), dtype=torch.uint8, device="cuda"); d_desc = torch.zeros((1, cap, 128), device="cuda")
    d_cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    s.detect_describe_batch_dev(d_gray, d_kp, d_desc, d_cnt, cap, st)
    torch.cuda.synchronize()
    kp, desc = s.detect_describe(want.astype(np.float32))  # gray.convertTo(CV_32FC1): plain cast (src/main.cpp:85)
    assert int(d_cnt[0]) == len(kp) > 20 and np.array_equal(d_desc[0, : len(kp)].cpu().numpy(), desc)
    s.close()
