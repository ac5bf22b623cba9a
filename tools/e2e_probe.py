"""Developer probe: host-buffer (u8) end-to-end frames/s for one (frames per call, chunk) setting; env knobs SIFT_B200_LANES/TAPER apply.
    python tools/e2e_probe.py B chunk [steps]
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
import bench
import numpy as np, torch

pkg = ge.load_package()
B, chunk = int(sys.argv[1]), int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
cap = 6144
bench.bind_near_gpu(0)
uniq = bench.make_frames(32)
frames = np.concatenate([uniq] * ((B + 31) // 32))[:B]
host_u8 = torch.from_numpy(frames.astype(np.uint8)).pin_memory()
h_kp = torch.zeros((B, cap, 28), dtype=torch.uint8).pin_memory()
h_desc = torch.zeros((B, cap, 128), dtype=torch.float32).pin_memory()
h_cnt = torch.zeros(B, dtype=torch.int32).pin_memory()
s = pkg.Sift(1080, 1920, max_batch=chunk, max_kp_per_frame=cap)
fn = lambda: s.detect_describe_batch_host_u8_ptr(host_u8.data_ptr(), B, 1080, 1920, h_kp.data_ptr(), h_desc.data_ptr(), h_cnt.data_ptr(), cap)
for _ in range(2):
    fn()
torch.cuda.synchronize()
ts = []
for _ in range(steps):
    t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
print(json.dumps({"B": B, "chunk": chunk, "lanes": os.environ.get("SIFT_B200_LANES"), "taper": os.environ.get("SIFT_B200_TAPER"),
                  "fps_mean": round(B * steps / sum(ts), 1), "fps_best": round(B / min(ts), 1), "ms": [round(t * 1e3, 2) for t in ts]}))
s.close()
