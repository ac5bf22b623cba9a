"""GPU check of the exact-pyramid mode: bit equality of the pyramid, end-to-end descriptor agreement, and its throughput."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
import torch

pkg = ge.load_package()
G = os.path.join(os.path.dirname(__file__), "..", "tests", "golden")
s = pkg.Sift(2448, 2448, max_batch=8, max_kp_per_frame=1 << 15)
for mode in (False, True):
    s.set_exact_pyramid(mode)
    print("exact" if mode else "fast")
    for name in ("synth_160x120", "synth_odd_211x173"):
        z = np.load(os.path.join(G, name + ".npz"))
        img = z["image"].astype(np.float32)
        g = s.build_gaussian_pyramid(img, 5)
        kp, desc = s.detect_describe(img)
        err = np.linalg.norm(desc - z["descriptors"], axis=1) if len(kp) == len(z["keypoints"]) else np.array([9.0])
        print(f"  {name}: pyramid equal {np.array_equal(g, z['gpyr'])} max|d| {np.abs(g - z['gpyr']).max():.3g}  n_kp {len(kp)}/{len(z['keypoints'])}"
              f"  desc<=1e-3 {100 * (err <= 1e-3).mean():.2f} %  max {err.max():.3g}")
    for name in ("scene_960", "query_2448"):
        z = np.load(os.path.join(G, name + ".npz"))
        kp, desc = s.detect_describe(z["gray"].astype(np.float32))
        ok = "keypoints" in z and len(kp) == len(z["keypoints"])
        if ok:
            err = np.linalg.norm(desc - z["descriptors"], axis=1)
            dx = max(np.abs(kp["x"] - z["keypoints"]["x"]).max(), np.abs(kp["y"] - z["keypoints"]["y"]).max())
            da = np.abs(kp["angle"] - z["keypoints"]["angle"]); da = np.minimum(da, 360 - da).max()
            print(f"  {name}: n_kp {len(kp)}  pos max {dx:.3g} px  angle max {da:.3g} deg  desc<=1e-3 {100 * (err <= 1e-3).mean():.3f} %  exact-equal rows {100 * (err == 0).mean():.2f} %  max {err.max():.3g}")
        else:
            print(f"  {name}: n_kp {len(kp)} (fixture keys {list(z.keys())})")
    # throughput, 8 device-resident 1080p frames
    import bench
    fr = torch.from_numpy(bench.make_frames(8)).cuda()
    cap = 6144
    d_kp = torch.zeros((8, cap, 28), dtype=torch.uint8, device="cuda"); d_desc = torch.zeros((8, cap, 128), device="cuda"); d_cnt = torch.zeros(8, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2): s.detect_describe_batch_dev(fr, d_kp, d_desc, d_cnt, cap, st)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3): s.detect_describe_batch_dev(fr, d_kp, d_desc, d_cnt, cap, st)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"  1080p: {24 / dt:.0f} frames/s")
