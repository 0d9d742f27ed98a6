"""The e2e call of bench.py in isolation (run on a B200): synthetic resident matrix kept alive, read back as the host matrix, a warm-up on a
1M-row slice, then the timed dpr_sparse_k_means calls.  DPR_TIMING=1 prints the phases."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import depecher_b200 as dp  # noqa: E402

eng = dp.Engine(device=0)
n, d, k = 10_000_000, 40, 30
ds = eng.synthetic(n, d, 24, 2.5, 20261021, scale=0.0)
Xh = np.asfortranarray(eng.read(ds, 0, n))
eng.synchronize()
reg = 1.0 * n * np.sqrt(d) / 1450.0
eng.sparse_k_means(np.asfortranarray(Xh[:1_000_000]), k, 1.0 * 1_000_000 * np.sqrt(d) / 1450.0, True, seed_off_set=1)
for rep in range(4):
    t0 = time.perf_counter()
    res = eng.sparse_k_means(Xh, k, reg, True, seed_off_set=1)
    print(f"sparse_k_means: {time.perf_counter()-t0:.4f} s, {res['n_iter']} iterations", flush=True)
