"""Single-frame latency breakdown at 1080p: per-stage CUDA-event times of a 1-frame batch, and host-to-host wall time."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
import bench, torch

pkg = ge.load_package()
fr = bench.make_frames(1)
s = pkg.Sift(1080, 1920, max_batch=1, max_kp_per_frame=6144)
d = torch.from_numpy(fr).cuda(); cap = 6144
d_kp = torch.zeros((1, cap, 28), dtype=torch.uint8, device="cuda"); d_desc = torch.zeros((1, cap, 128), device="cuda"); d_cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
s.set_stage_timing(True)
for _ in range(3):
    s.detect_describe_batch_dev(d, d_kp, d_desc, d_cnt, cap, st); torch.cuda.synchronize()
print("stage ms [base, octave, gradient, extrema+refine, orientation, order, describe, total]:", [round(x, 3) for x in s.stage_ms()])
s.set_stage_timing(False)
for name, img in (("pageable", fr[0]), ("pinned", torch.from_numpy(fr[0]).pin_memory().numpy())):
    lat = []
    for _ in range(12):
        t0 = time.perf_counter(); kp, desc = s.detect_describe(img); lat.append(time.perf_counter() - t0)
    print(name, "host-to-host ms:", round(1e3 * float(np.median(lat[2:])), 3), len(kp))
