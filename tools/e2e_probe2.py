"""Developer probe: host-buffer (u8) end-to-end frames/s with T host threads, each driving its own handle and buffers (calls overlap:
one call's pipeline fill/drain runs under the other's steady state).
    python tools/e2e_probe2.py B chunk threads [steps]
"""
import json, os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
import bench
import numpy as np, torch

pkg = ge.load_package()
B, chunk, T = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 6
cap = 6144
bench.bind_near_gpu(0)
uniq = bench.make_frames(32)
frames = np.concatenate([uniq] * ((B + 31) // 32))[:B]
ctx = []
for t in range(T):
    host_u8 = torch.from_numpy(frames.astype(np.uint8)).pin_memory()
    h_kp = torch.zeros((B, cap, 28), dtype=torch.uint8).pin_memory()
    h_desc = torch.zeros((B, cap, 128), dtype=torch.float32).pin_memory()
    h_cnt = torch.zeros(B, dtype=torch.int32).pin_memory()
    s = pkg.Sift(1080, 1920, max_batch=chunk, max_kp_per_frame=cap)
    ctx.append((s, host_u8, h_kp, h_desc, h_cnt))

def run(c, n):
    s, host_u8, h_kp, h_desc, h_cnt = c
    for _ in range(n):
        s.detect_describe_batch_host_u8_ptr(host_u8.data_ptr(), B, 1080, 1920, h_kp.data_ptr(), h_desc.data_ptr(), h_cnt.data_ptr(), cap)

for c in ctx:
    run(c, 2)
torch.cuda.synchronize()
t0 = time.perf_counter()
th = [threading.Thread(target=run, args=(c, steps)) for c in ctx]
for x in th: x.start()
for x in th: x.join()
dt = time.perf_counter() - t0
print(json.dumps({"B": B, "chunk": chunk, "threads": T, "fps": round(B * steps * T / dt, 1), "counts_equal": bool(all((c[4] == ctx[0][4]).all() for c in ctx))}))
for c in ctx: c[0].close()
