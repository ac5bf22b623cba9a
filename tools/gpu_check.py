"""Developer probe (not a test, not the benchmark): stage-by-stage GPU-vs-oracle report + stage timings.
Run on a GPU box:  python tools/gpu_check.py [--big]  > gpurun_out/check.log
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge

pkg = ge.load_package()
O = ge.load_oracle()
from importlib import import_module

synth = import_module("sift_gpu_b200.synth")


def match_kps(a, b):
    """Greedy one-to-one match on (octave field low 16 bits, position within 0.01 px, angle within 1 deg)."""
    from collections import defaultdict

    buckets = defaultdict(list)
    for j, k in enumerate(b):
        buckets[(int(k["octave"]) & 0xFFFF, int(round(float(k["x"]))), int(round(float(k["y"]))))].append(j)
    used = set()
    pairs = []
    for i, k in enumerate(a):
        o = int(k["octave"]) & 0xFFFF
        cx, cy = int(round(float(k["x"]))), int(round(float(k["y"])))
        best = None
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for j in buckets.get((o, cx + dx, cy + dy), ()):
                    if j in used:
                        continue
                    q = b[j]
                    dpos = max(abs(float(k["x"]) - float(q["x"])), abs(float(k["y"]) - float(q["y"])))
                    dang = abs(float(k["angle"]) - float(q["angle"]))
                    dang = min(dang, 360 - dang)
                    if dpos <= 0.01 and dang <= 1.0 and (best is None or dpos < best[1]):
                        best = (j, dpos, dang)
        if best:
            used.add(best[0])
            pairs.append((i, best[0], best[1], best[2]))
    return pairs


def report(name, img, s):
    rows, cols = img.shape
    o32 = O.f32()
    O.set_threads(os.cpu_count())
    t = time.time()
    okp, odesc, og, od, opq = o32.sift_ncl(img, want_pyramids=True, want_prequant=True)
    t_or = time.time() - t
    g = s.build_gaussian_pyramid(img)
    d = s.build_dog_pyramid(g, rows, cols)
    print(f"[{name}] {cols}x{rows} oracle {t_or:.2f}s  N_oracle={len(okp)}")
    gl, ogl = pkg.unpack(g, rows, cols, 5, 5), O.unpack(og, rows, cols, 5, 5)
    for i, (a, b) in enumerate(zip(gl, ogl)):
        if i % 5 in (0, 2, 4):
            print(f"   gpyr[{i}] max|diff| {np.abs(a - b).max():.3e}  rms {np.sqrt(np.mean((a - b) ** 2)):.3e}")
    dl, odl = pkg.unpack(d, rows, cols, 5, 4), O.unpack(od, rows, cols, 5, 4)
    print("   dog max|diff| per level:", " ".join(f"{np.abs(a - b).max():.1e}" for a, b in zip(dl, odl)))
    # stage parity on ORACLE inputs
    kp_stage = s.find_scale_space_extrema(og, od, rows, cols)
    same = len(kp_stage) == len(okp) and kp_stage.tobytes() == okp.tobytes()
    print(f"   extrema stage on oracle pyramids: N={len(kp_stage)} bit-identical={same}")
    if not same and len(kp_stage) == len(okp):
        for f in okp.dtype.names:
            dd = np.abs(kp_stage[f].astype(np.float64) - okp[f].astype(np.float64)).max()
            print(f"      field {f}: max diff {dd:.3e}")
    desc_stage = s.cal_descriptor(og, rows, cols, okp)
    dist = np.linalg.norm(desc_stage - odesc, axis=1)
    print(f"   descriptor stage on oracle inputs: L2 median {np.median(dist):.2e} p99 {np.percentile(dist, 99):.2e} max {dist.max():.2e} "
          f"frac<=1e-3 {np.mean(dist <= 1e-3):.4f} bit-identical rows {np.mean(np.all(desc_stage == odesc, axis=1)):.4f}")
    # end to end
    kp, desc = s.detect_describe(img)
    pairs = match_kps(kp, okp)
    rec, prec = len(pairs) / max(1, len(okp)), len(pairs) / max(1, len(kp))
    print(f"   end-to-end: N_gpu={len(kp)} recall {rec:.4f} precision {prec:.4f}  order-identical={len(kp) == len(okp) and all(i == j for i, j, _, _ in pairs)}")
    if pairs:
        pi = np.array([p[0] for p in pairs]); pj = np.array([p[1] for p in pairs])
        dist = np.linalg.norm(desc[pi] - odesc[pj], axis=1)
        print(f"   end-to-end descriptors: L2 median {np.median(dist):.2e} p90 {np.percentile(dist, 90):.2e} max {dist.max():.2e} frac<=1e-3 {np.mean(dist <= 1e-3):.4f}")
        print(f"   max pos diff {max(p[2] for p in pairs):.2e} px, max angle diff {max(p[3] for p in pairs):.3f} deg")
    # loose match stats (how far are the unmatched ones?)
    return kp, desc, okp, odesc


def main():
    import torch

    big = "--big" in sys.argv
    print("device:", torch.cuda.get_device_name(0))
    s = pkg.Sift(1280, 2048, max_batch=1, max_kp_per_frame=1 << 15)
    report("synth320", synth.recipe_s(320, 240, seed=5), s)
    report("synth_odd", synth.recipe_s(417, 303, seed=9), s)
    report("synth960x540", synth.recipe_s(960, 540, seed=11), s)
    gpath = os.path.join(ge.ROOT, "tests", "golden", "scene_960.npz")
    if os.path.exists(gpath):
        z = np.load(gpath)
        report("scene960", z["gray"].astype(np.float32), s)
    if big:
        report("synth1080p", synth.recipe_s(1920, 1080, seed=1234), s)
    # blur stage API + 1-D
    img = synth.recipe_s(300, 200, seed=3)
    for sg in (1.6, 2.771281, 6.196774, 0.9):
        a, b = s.gaussian_blur(img, sg), O.f32().gaussian_blur(img, sg)
        print(f"blur sigma={sg}: max|diff| {np.abs(a - b).max():.3e}")
    a, b = s.gaussian_blur(img, 1.6, one_d=True), O.f32().gaussian_blur_1d(img, 1.6)
    print(f"blur_1d: bit-identical={np.array_equal(a, b)} max|diff| {np.abs(a - b).max():.3e}")
    # matcher
    rng = np.random.default_rng(0)
    q = rng.random((700, 128), dtype=np.float32); t = rng.random((500, 128), dtype=np.float32)
    t[10] = t[3]
    q[5] = t[3]
    for norm in (pkg.NORM_L1, pkg.NORM_L2):
        gi, gd, gg = s.match_knn2(q, t, norm)
        oi, od_, og_ = O.match_knn2(q, t, norm)
        print(f"match norm={norm}: idx identical={np.array_equal(gi, oi)} dist max diff {np.abs(gd - od_).max():.2e} good identical={np.array_equal(gg, og_)}")
    s.close()

    # throughput probe: 1080p batch, device resident
    F, NB, cap = 8, 4, 8192
    frames = np.stack([synth.recipe_s(1920, 1080, seed=1234 + k) for k in range(F)])
    d_imgs = torch.from_numpy(np.concatenate([frames] * NB)).cuda()
    n = d_imgs.shape[0]
    d_kp = torch.zeros((n, cap, 28), dtype=torch.uint8, device="cuda")
    d_desc = torch.zeros((n, cap, 128), dtype=torch.float32, device="cuda")
    d_cnt = torch.zeros(n, dtype=torch.int32, device="cuda")
    s = pkg.Sift(1080, 1920, max_batch=F, max_kp_per_frame=cap)
    s.set_stage_timing(True)
    st = torch.cuda.current_stream().cuda_stream
    for it in range(3):
        s.detect_describe_batch_dev(d_imgs, d_kp, d_desc, d_cnt, cap, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    iters = 5
    for it in range(iters):
        s.detect_describe_batch_dev(d_imgs, d_kp, d_desc, d_cnt, cap, st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"1080p batch of {n} (chunks of {F}): {ms:.3f} ms -> {n / ms * 1e3:.1f} frames/s, counts[:4]={d_cnt[:4].tolist()}")
    sm = s.stage_ms()
    names = ["base", "octaves", "extrema", "orient", "order", "describe", "total"]
    print("stage ms for last chunk of", F, "frames:", {k: round(v, 3) for k, v in zip(names, sm)})
    s.close()


if __name__ == "__main__":
    main()
