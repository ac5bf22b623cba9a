"""Fixed workload for ncu captures: one chunk of 32 synthetic 1080p frames (the bench's chunk size) through the whole path, twice."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
import bench

pkg = ge.load_package()
import torch

F, cap = int(os.environ.get("PROF_FRAMES", "32")), 6144
frames = bench.make_frames(32)[:F]
d = torch.from_numpy(frames).cuda()
d_kp = torch.zeros((F, cap, 28), dtype=torch.uint8, device="cuda")
d_desc = torch.zeros((F, cap, 128), dtype=torch.float32, device="cuda")
d_cnt = torch.zeros(F, dtype=torch.int32, device="cuda")
s = pkg.Sift(1080, 1920, max_batch=F, max_kp_per_frame=cap)
for _ in range(2):
    s.detect_describe_batch_dev(d, d_kp, d_desc, d_cnt, cap, torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print("counts", d_cnt[:8].tolist())
s.close()
