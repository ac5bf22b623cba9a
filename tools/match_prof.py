"""Fixed matcher workload for ncu: MATCH_N x MATCH_N RootSIFT-like descriptors through the tensor-core path, twice."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge

pkg = ge.load_package()
n = int(os.environ.get("MATCH_N", "20000"))
rng = np.random.default_rng(3)
def rs(n):
    d = rng.gamma(0.6, 1.0, size=(n, 128)).astype(np.float32)
    d /= d.sum(1, keepdims=True)
    return np.sqrt(d).astype(np.float32)
q, t = rs(n), rs(n)
s = pkg.Sift(64, 64, 1, 64)
for _ in range(2):
    idx, dist, good, ms = s.match_knn2(q, t, pkg.NORM_L2, 0.86, tensor_cores=True, timing=True)
    print(f"n={n}: kernels {ms * 1e3:.1f} us, good {int(good.sum())}")
s.close()
