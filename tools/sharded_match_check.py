"""Two or more GPUs of one box (torchrun): the train descriptors are sharded over the ranks, every rank matches all queries against its
slice through the C ABI (exact and tensor-core kernels), the per-shard top-2 are exchanged with one NCCL all_gather and merged; the result
must equal the single-device match of the whole train set.  Prints one JSON line on rank 0.
    python -m torch.distributed.run --nproc-per-node N tools/sharded_match_check.py
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import __graft_entry__ as ge

pkg = ge.load_package()
from importlib import import_module
sh = import_module("sift_gpu_b200.sharded")
rank, world = int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
rng = np.random.default_rng(7)
nq, nt = 4000, 20000
q = rng.random((nq, 128)).astype(np.float32); q /= np.linalg.norm(q, axis=1, keepdims=True)
t = rng.random((nt, 128)).astype(np.float32); t /= np.linalg.norm(t, axis=1, keepdims=True)
t[15000] = t[100]; q[3] = t[100]  # duplicates across shards: the tie must go to the lowest index
s = pkg.Sift(64, 64, max_batch=1, max_kp_per_frame=64, device=rank)
lo, hi = sh.shard_rows(nt, world, rank)
out = {}
for tc in (False, True):
    want = s.match_knn2(q, t, pkg.NORM_L2, 0.86, tensor_cores=tc)
    dist.barrier(); t0 = time.perf_counter()
    idx, d, good = sh.match_knn2_train_sharded(lambda a, b: s.match_knn2(a, b, pkg.NORM_L2, 0.86, tensor_cores=tc), q, t[lo:hi], lo,
                                               ratio=0.86, all_gather=sh.torch_all_gather())
    dt = time.perf_counter() - t0
    ok = bool(np.array_equal(idx, want[0]) and np.array_equal(good, want[2]) and np.allclose(d, want[1], rtol=1e-6, atol=1e-7))
    out["tensor_cores" if tc else "exact"] = {"identical_to_single_device": ok, "tie_row": [int(idx[3, 0]), int(idx[3, 1])], "host_ms": round(dt * 1e3, 2)}
flags = [None] * world
dist.all_gather_object(flags, all(v["identical_to_single_device"] for v in out.values()))
if rank == 0:
    print(json.dumps({"ranks": world, "queries": nq, "train": nt, "all_ranks_identical": all(flags), **out}))
s.close()
dist.destroy_process_group()
