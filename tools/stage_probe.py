"""Developer probe: per-kernel CUDA-event times (us per 1080p frame) of one 32-frame chunk, for A/B runs.

    python tools/stage_probe.py [label]            # library from SIFT_B200_LIB or the in-tree build; env knobs apply
Prints one JSON line: {"label", "us_per_frame": {kernel: us}, "total_us", "counts_sum", "desc_checksum"}.
The checksum (sum of all descriptors, float64) lets two builds be compared for identical outputs.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
import bench

pkg = ge.load_package()
import numpy as np
import torch

label = sys.argv[1] if len(sys.argv) > 1 else os.environ.get("SIFT_B200_LIB", "in-tree")
F, cap = int(os.environ.get("PROF_FRAMES", "32")), 6144
frames = bench.make_frames(32)[:F]
d = torch.from_numpy(frames).cuda()
d_kp = torch.zeros((F, cap, 28), dtype=torch.uint8, device="cuda")
d_desc = torch.zeros((F, cap, 128), dtype=torch.float32, device="cuda")
d_cnt = torch.zeros(F, dtype=torch.int32, device="cuda")
s = pkg.Sift(1080, 1920, max_batch=F, max_kp_per_frame=cap)
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    s.detect_describe_batch_dev(d, d_kp, d_desc, d_cnt, cap, st)
torch.cuda.synchronize()
names = ["base_blur", "octave", "gradient", "extrema", "orientation", "order_scan", "describe", "total"]
acc = np.zeros(8)
reps = int(os.environ.get("PROBE_REPS", "5"))
s.set_stage_timing(True)
for _ in range(reps):
    s.detect_describe_batch_dev(d, d_kp, d_desc, d_cnt, cap, st)
    torch.cuda.synchronize()
    acc += np.array(s.stage_ms()[:8])
s.set_stage_timing(False)
us = acc / reps * 1e3 / F
cnt = d_cnt.cpu().numpy()
desc = d_desc.cpu().numpy()
chk = float(sum(desc[f, : cnt[f]].astype(np.float64).sum() for f in range(F)))
print(json.dumps({"label": label, "us_per_frame": {n: round(float(u), 2) for n, u in zip(names, us)}, "counts_sum": int(cnt.sum()),
                  "desc_checksum": chk}))
s.close()
