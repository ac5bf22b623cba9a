"""Host-side logic that needs no GPU: recipe-S frames, packed-pyramid bookkeeping, frame sharding across ranks
(world_size-2 gloo), roofline accounting."""
import os
import subprocess
import sys

import numpy as np


def test_recipe_s_is_deterministic_and_in_range(synth):
    a = synth.recipe_s(200, 120, seed=7)
    b = synth.recipe_s(200, 120, seed=7)
    c = synth.recipe_s(200, 120, seed=8)
    assert a.dtype == np.float32 and a.shape == (120, 200)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert a.min() >= 0 and a.max() <= 255


def test_packed_pyramid_bookkeeping(pkg, oracle):
    for rows, cols in [(1080, 1920), (135, 241), (17, 16)]:
        dims = pkg.octave_dims(rows, cols, 5)
        assert dims == oracle.octave_dims(rows, cols, 5)
        assert dims[0] == (rows, cols) and dims[4] == (rows // 16, cols // 16)
        n = pkg.packed_size(rows, cols, 5, 5)
        lv = pkg.unpack(np.arange(n, dtype=np.float32), rows, cols, 5, 5)
        assert len(lv) == 25 and lv[7].shape == dims[1] and sum(x.size for x in lv) == n


def test_algorithmic_bytes_match_survey():
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench

    total, per = bench.algorithmic_bytes_per_frame(3500)
    assert abs(total - 164.86e6) < 0.2e6  # 164.8 MB per 1080p frame at N = 3.5 k
    assert sum(per.values()) == total


def test_shard_range_covers_everything():
    import bench

    for n in (1024, 1000, 7, 1):
        for world in (1, 2, 4, 8):
            got = []
            for r in range(world):
                lo, hi = bench.shard_range(n, world, r)
                got += list(range(lo, hi))
            assert got == list(range(n))


_WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
import bench
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
lo, hi = bench.shard_range(10, w, r)
t = bench.max_over_ranks(1.0 + r)           # the slowest rank defines the step time
n = bench.sum_over_ranks(float(hi - lo))    # every frame is counted exactly once
assert t == float(w) and n == 10.0, (t, n)
dist.barrier()
print("ok", r)
"""


def test_two_rank_reduction_gloo(tmp_path):
    """N>1 path on CPU: two gloo ranks shard frames, take the max step time and the total frame count."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", str(script), root], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2


def test_upsample2x_restatement_matches_cv2(synth):
    """The numpy restatement of cv::resize(2x, INTER_LINEAR) used as the checker for the GPU front end (config 3).
    cv2's SIMD path fuses one multiply-add, so agreement is to 1e-4 on 0..255 data, not bit-exact."""
    import pytest

    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    for shape in [(37, 53), (120, 200), (5, 4), (2, 2)]:
        img = (rng.random(shape) * 255).astype(np.float32)
        got = synth.upsample2x(img)
        want = cv2.resize(img, (2 * shape[1], 2 * shape[0]), interpolation=cv2.INTER_LINEAR)
        assert got.shape == want.shape and np.abs(got - want).max() <= 1e-4


def test_bind_near_gpu_is_harmless_without_nvml():
    """bench.bind_near_gpu pins the process to the GPU-local CPUs before the pinned host buffers are allocated; without a GPU / NVML it
    must change nothing and say so."""
    import os

    import bench

    before = os.sched_getaffinity(0)
    prev, n = bench.bind_near_gpu(0)
    after = os.sched_getaffinity(0)
    if prev is None:
        assert n == 0 and after == before
    else:  # a box with NVML: bound to a non-empty subset, and the caller can restore
        assert 0 < n <= len(before) and after <= before
        os.sched_setaffinity(0, prev)
        assert os.sched_getaffinity(0) == before


def test_chunk_plan_is_a_partition(pkg):
    """Host-batch chunk schedule (api.cu chunk_plan): every chunk in [1, max_batch], chunks sum to n_frames, for every
    (max_batch, n_frames) -- the tapered heads once overshot n_frames (e.g. max_batch 9, n 27) and produced a negative chunk."""
    for mb in list(range(1, 70)) + [128, 256]:
        for n in list(range(0, 6 * mb + 3)) + [1024, 1000]:
            for taper in (True, False):
                plan = pkg.chunk_plan(n, mb, taper)
                assert sum(plan) == n, (mb, n, taper, plan)
                assert all(0 < c <= mb for c in plan), (mb, n, taper, plan)
                if not taper:
                    assert len(plan) == (n + mb - 1) // mb
    # the taper is symmetric and leaves full chunks in the middle
    plan = pkg.chunk_plan(256, 32, True)
    assert plan[:3] == [4, 8, 16] and plan[-3:] == [16, 8, 4] and 32 in plan
