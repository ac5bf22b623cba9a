import sys, os, json
sys.path.insert(0, '/root/repo')
import __graft_entry__ as ge, bench, numpy as np, torch
pkg = ge.load_package()
F, cap = 32, 6144
frames = bench.make_frames(32)
for name, arr in (("f32", frames), ("u8", frames.astype(np.uint8))):
    d = torch.from_numpy(arr).cuda()
    d_kp = torch.zeros((F, cap, 28), dtype=torch.uint8, device="cuda"); d_desc = torch.zeros((F, cap, 128), dtype=torch.float32, device="cuda"); d_cnt = torch.zeros(F, dtype=torch.int32, device="cuda")
    s = pkg.Sift(1080, 1920, max_batch=F, max_kp_per_frame=cap)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3): s.detect_describe_batch_dev(d, d_kp, d_desc, d_cnt, cap, st)
    torch.cuda.synchronize(); s.set_stage_timing(True); acc = np.zeros(8)
    for _ in range(5):
        s.detect_describe_batch_dev(d, d_kp, d_desc, d_cnt, cap, st); torch.cuda.synchronize(); acc += np.array(s.stage_ms()[:8])
    print(name, [round(float(x), 2) for x in acc / 5 * 1e3 / F]); s.close()
