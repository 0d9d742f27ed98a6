"""Where the time of the e2e call dpr_sparse_k_means(host doubles) goes (run on a B200): upload alone, then the whole call."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import depecher_b200 as dp  # noqa: E402

print("host cores", os.cpu_count(), flush=True)
eng = dp.Engine(device=0)
n, d, k = 10_000_000, 40, 30
ds = eng.synthetic(n, d, 24, 2.5, 20261021, scale=0.0)
Xh = np.asfortranarray(eng.read(ds, 0, n))
ds.close()
reg = 1.0 * n * np.sqrt(d) / 1450.0
for rep in range(3):
    t0 = time.perf_counter()
    h = eng.dataset(Xh)
    eng.synchronize()
    t1 = time.perf_counter()
    h.close()
    print(f"upload of {Xh.nbytes/1e9:.1f} GB host doubles: {t1-t0:.4f} s  ({Xh.nbytes/1e9/(t1-t0):.1f} GB/s of doubles)", flush=True)
for rep in range(3):
    t0 = time.perf_counter()
    res = eng.sparse_k_means(Xh, k, reg, True, seed_off_set=1)
    t1 = time.perf_counter()
    print(f"sparse_k_means: {t1-t0:.4f} s, {res['n_iter']} iterations", flush=True)
h = eng.dataset(Xh)
eng.synchronize()
t0 = time.perf_counter()
mu0, _ = eng.initialize_mu(h, k, seed=1, fit_id=0)
eng.synchronize()
print(f"k-means++ seeding on the resident matrix: {time.perf_counter()-t0:.4f} s", flush=True)
h.close()
for rep in range(3):
    t0 = time.perf_counter()
    hs, lcs = eng.dataset_scaled(Xh, False, "default", True)
    eng.synchronize()
    print(f"dataset_scaled (upload of doubles + device pre-scaling): {time.perf_counter()-t0:.4f} s", flush=True)
    hs.close()
h = eng.dataset(Xh)
eng.synchronize()
for want in (True, False, True, False):
    t0 = time.perf_counter()
    res = eng.sparse_k_means(h, k, reg, True, seed_off_set=1, want_labels=want)
    eng.synchronize()
    print(f"sparse_k_means on the resident matrix (labels to the host: {want}): {time.perf_counter()-t0:.4f} s, {res['n_iter']} iterations", flush=True)
h.close()
