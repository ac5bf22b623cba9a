"""Cross-device matcher batching: per-shard knn-2 + top-2 merge == one match over the whole train set.
CPU tests use the oracle's exhaustive matcher as each shard's local matcher; the GPU test runs the shards through the C ABI."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sharded(pkg):
    from importlib import import_module

    return import_module("sift_gpu_b200.sharded")


def _descriptors(n, seed, dup_from=None):
    rng = np.random.default_rng(seed)
    d = rng.random((n, 128)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    if dup_from is not None:  # exact duplicates: equal distances, the tie must go to the lowest train index
        d[dup_from[1]] = d[dup_from[0]]
    return d


def test_shard_rows_partition(pkg):
    sh = _sharded(pkg)
    for n in (0, 1, 7, 486, 1273):
        for world in (1, 2, 3, 8):
            blocks = [sh.shard_rows(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            assert max(hi - lo for lo, hi in blocks) - min(hi - lo for lo, hi in blocks) <= 1


@pytest.mark.parametrize("norm", [2, 4])
@pytest.mark.parametrize("world", [1, 2, 3, 5])
def test_train_sharded_merge_equals_global_match(pkg, oracle, norm, world):
    sh = _sharded(pkg)
    q = _descriptors(97, 1)
    t = _descriptors(203, 2, dup_from=(10, 150))  # rows 10 and 150 identical, in different shards for world >= 2
    q[5] = t[10]  # distance 0 to both duplicates: best = 10, second = 150
    want_idx, want_dist, want_good = oracle.match_knn2(q, t, norm, 0.86)
    packed = []
    for r in range(world):
        lo, hi = sh.shard_rows(len(t), world, r)
        gathered = []
        sh.match_knn2_train_sharded(lambda a, b: oracle.match_knn2(a, b, norm, 0.86), q, t[lo:hi], lo,
                                    all_gather=lambda arr: gathered.append(arr.copy()) or [arr])
        packed.append(gathered[0])
    idx, dist, good = sh.match_knn2_train_sharded(lambda a, b: oracle.match_knn2(a, b, norm, 0.86), q, t[:0], 0, all_gather=lambda arr: packed)
    assert np.array_equal(idx, want_idx) and np.array_equal(dist, want_dist) and np.array_equal(good, want_good)
    assert tuple(idx[5]) == (10, 150)


def test_merge_handles_tiny_shards(pkg, oracle):
    """Shards with one or zero train rows report -1 / +inf for the missing neighbours; the merge ignores them."""
    sh = _sharded(pkg)
    q = _descriptors(9, 3)
    t = _descriptors(3, 4)
    want_idx, want_dist, _ = oracle.match_knn2(q, t, 4, 0.86)
    parts_d, parts_i = [], []
    for lo, hi in [(0, 1), (1, 1), (1, 3)]:
        if hi - lo >= 2:
            i, d, _ = oracle.match_knn2(q, t[lo:hi], 4, 0.86)
            i = i + lo
        elif hi - lo == 1:
            d1 = np.sqrt(((q.astype(np.float64) - t[lo].astype(np.float64)) ** 2).sum(1)).astype(np.float32)
            d = np.stack([d1, np.full(len(q), np.inf, np.float32)], axis=1)
            i = np.stack([np.full(len(q), lo, np.int32), np.full(len(q), -1, np.int32)], axis=1)
        else:
            d = np.full((len(q), 2), np.inf, np.float32)
            i = np.full((len(q), 2), -1, np.int32)
        parts_d.append(d)
        parts_i.append(i)
    idx, dist = sh.merge_knn2(parts_d, parts_i)
    assert np.array_equal(idx, want_idx) and np.allclose(dist, want_dist, rtol=1e-6)


_WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch.distributed as dist
import __graft_entry__ as ge
ge.load_package()
O = ge.load_oracle()
from importlib import import_module
sh = import_module("sift_gpu_b200.sharded")
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
rng = np.random.default_rng(11)
q = rng.random((64, 128)).astype(np.float32)
t = rng.random((301, 128)).astype(np.float32)
lo, hi = sh.shard_rows(len(t), w, r)
idx, d, good = sh.match_knn2_train_sharded(lambda a, b: O.match_knn2(a, b, 4, 0.86), q, t[lo:hi], lo, all_gather=sh.torch_all_gather())
wi, wd, wg = O.match_knn2(q, t, 4, 0.86)
assert np.array_equal(idx, wi) and np.array_equal(d, wd) and np.array_equal(good, wg)
dist.barrier()
print("ok", r)
"""


def test_train_sharded_two_ranks_gloo(tmp_path):
    """The one real exchange of the matcher path (all_gather of the per-shard top-2) over two gloo ranks on the CPU."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", str(script), ROOT], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2


@pytest.mark.gpu
@pytest.mark.parametrize("tensor_cores", [False, True])
def test_train_sharded_through_the_c_abi(pkg, oracle, tensor_cores):
    """Every shard matched on the GPU through sift_b200_match_knn2_ex (exact and tcgen05 kernels), merged on the host: indices,
    distances and ratio flags identical to the oracle's match over the whole train set, duplicates across shards included."""
    sh = _sharded(pkg)
    q = _descriptors(700, 21)
    t = _descriptors(1900, 22, dup_from=(100, 1500))
    q[3] = t[100]
    want_idx, want_dist, want_good = oracle.match_knn2(q, t, 4, 0.86)
    s = pkg.Sift(64, 64, max_batch=1, max_kp_per_frame=64)
    for world in (2, 4):
        packed = []
        for r in range(world):
            lo, hi = sh.shard_rows(len(t), world, r)
            got = []
            sh.match_knn2_train_sharded(lambda a, b: s.match_knn2(a, b, 4, 0.86, tensor_cores=tensor_cores), q, t[lo:hi], lo,
                                        all_gather=lambda arr: got.append(arr.copy()) or [arr])
            packed.append(got[0])
        idx, dist, good = sh.match_knn2_train_sharded(None, q, t[:0], 0, all_gather=lambda arr: packed)
        assert np.array_equal(idx, want_idx) and np.array_equal(good, want_good)
        assert np.allclose(dist, want_dist, rtol=1e-6, atol=1e-7)
    s.close()
