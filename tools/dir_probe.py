"""Developer probe (torchrun, N ranks): aggregate pinned-host copy rates of a box, per direction and both at once, plus the rate at which
the ranks' host threads can WRITE host memory (what a host-side expansion of compact results would need).
    python -m torch.distributed.run --nproc-per-node N tools/dir_probe.py
"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist, numpy as np
import bench
rank = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
bench.bind_near_gpu(rank)
torch.cuda.set_device(rank)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
MB = 1 << 20
n = 1024 * MB
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda"); d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def sync():
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
def timed(fn, reps=4):
    fn(); sync()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    sync()
    return (time.perf_counter() - t0) / reps
def h2d():
    with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
def both():
    h2d(); d2h()
res = {}
for name, fn, nbytes in (("h2d", h2d, n), ("d2h", d2h, n), ("both", both, 2 * n)):
    t = timed(fn)
    res[name + "_GBps_per_rank"] = round(nbytes / t / 1e9, 1)
# host write rate: numpy fill of a 1 GB float32 buffer by this rank's main thread (single thread)
buf = np.empty(n // 4, dtype=np.float32)
src = np.random.default_rng(0).integers(0, 255, n // 4, dtype=np.uint8)
sync()
t0 = time.perf_counter(); np.sqrt(src, out=buf, dtype=np.float32); t = time.perf_counter() - t0
res["host_sqrt_expand_GBps_per_thread"] = round(n / t / 1e9, 2)
sync()
if world > 1:
    allres = [None] * world
    dist.all_gather_object(allres, res)
else:
    allres = [res]
if rank == 0:
    agg = {k: round(sum(r[k] for r in allres), 1) for k in res}
    print(json.dumps({"ranks": world, "aggregate": agg, "rank0": res}))
if world > 1: dist.destroy_process_group()
