"""Tiny workload for compute-sanitizer: odd-sized frames through every entry point (whole path, batch, stage APIs, matcher)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
pkg = ge.load_package()
from importlib import import_module
synth = import_module("sift_gpu_b200.synth")
import torch
s = pkg.Sift(200, 260, max_batch=2, max_kp_per_frame=2048)
for (w, h, seed) in [(97, 83, 1), (260, 200, 2), (33, 47, 3), (16, 16, 4)]:
    img = synth.recipe_s(w, h, seed=seed, blobs_per_1080p=30000)
    kp, d = s.detect_describe(img)
    g = s.build_gaussian_pyramid(img)
    dg = s.build_dog_pyramid(g, h, w)
    k2 = s.find_scale_space_extrema(g, dg, h, w)
    d2 = s.cal_descriptor(g, h, w, k2)
    b = s.gaussian_blur(img, 2.0); b1 = s.gaussian_blur(img, 1.6, one_d=True)
    print(w, h, len(kp), len(k2))
frames = np.stack([synth.recipe_s(129, 75, seed=10 + k, blobs_per_1080p=30000) for k in range(5)])
d = torch.from_numpy(frames).cuda()
cap = 1024
d_kp = torch.zeros((5, cap, 28), dtype=torch.uint8, device="cuda"); d_desc = torch.zeros((5, cap, 128), device="cuda"); d_cnt = torch.zeros(5, dtype=torch.int32, device="cuda")
s.detect_describe_batch_dev(d, d_kp, d_desc, d_cnt, cap, torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
h_kp = np.zeros((5, cap), dtype=pkg.KP_DTYPE); h_desc = np.zeros((5, cap, 128), np.float32); h_cnt = np.zeros(5, np.int32)
s.detect_describe_batch_host(frames, h_kp, h_desc, h_cnt, cap)
up = s.detect_describe_up2(frames[0][:60, :90].copy())
q = np.random.default_rng(0).random((70, 128), dtype=np.float32)
s.match_knn2(q, q[:33], pkg.NORM_L1); s.match_knn2(q, q[:33], pkg.NORM_L2)
print("counts", d_cnt.tolist(), h_cnt.tolist())
s.close()
print("SANITIZE RUN DONE")
