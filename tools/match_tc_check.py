"""GPU check of the tensor-core matcher against the exact kernel and the oracle (run under `timeout`)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge

pkg = ge.load_package()
orc = ge.load_oracle()
s = pkg.Sift(64, 64, 1, 64)
rng = np.random.default_rng(7)


def rootsift_like(n):
    d = rng.gamma(0.6, 1.0, size=(n, 128)).astype(np.float32)
    d /= d.sum(1, keepdims=True)
    return np.sqrt(d).astype(np.float32)


# 1. structural probe: one-hot rows make every dot product a single exact product
q = np.zeros((128, 128), np.float32); t = np.zeros((128, 128), np.float32)
for i in range(128):
    q[i, i] = 1.0
    t[i, (i * 5 + 3) % 128] = 1.0
idx, dist, _ = s.match_knn2(q, t, pkg.NORM_L2, 0.86, tensor_cores=True)
want = np.array([((i - 3) * 77) % 128 for i in range(128)])  # 5*77 = 385 = 1 mod 128
print("one-hot probe best idx ok:", bool((idx[:, 0] == want).all()), "dist0 max", float(dist[:, 0].max()))
if not (idx[:, 0] == want).all():
    print(idx[:16, 0], want[:16])

ok = True
for nq, nt in [(1, 4), (5, 7), (128, 128), (129, 127), (300, 1000), (1358, 1444), (4000, 6000), (20000, 20000)]:
    q = rootsift_like(nq); t = rootsift_like(nt)
    if nq > 10 and nt > 10:  # near-duplicates and exact duplicates exercise the shortlist / tie rules
        t[3] = q[5]; t[9] = q[5]
        t[11] = q[7] + 1e-4 * rng.standard_normal(128).astype(np.float32)
    i0, d0, g0 = s.match_knn2(q, t, pkg.NORM_L2, 0.86)
    i1, d1, g1, ms = s.match_knn2(q, t, pkg.NORM_L2, 0.86, tensor_cores=True, timing=True)
    _, _, _, ms_exact = s.match_knn2(q, t, pkg.NORM_L2, 0.86, timing=True)
    io, do, _ = orc.match_knn2(q, t, pkg.NORM_L2)
    same = bool((i0 == i1).all() and (d0 == d1).all() and (g0 == g1).all())
    same_o = bool((io == i1).all())
    ok &= same and same_o
    print(f"nq={nq} nt={nt}: tc==exact {same}  tc==oracle idx {same_o}  tc {ms*1e3:.1f} us  exact {ms_exact*1e3:.1f} us", flush=True)
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "match_query_scene.npz"))
print("golden keys", list(g.keys()))
print("ALL OK" if ok else "MISMATCH")
