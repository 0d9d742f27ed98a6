"""Scoring throughput (run on a B200): M centre sets against the same 10M x 40 cells in ONE launch, as the penalty search does
(Clusterer.cpp:370-371).  DPR_TC_SETS=1 in the environment keeps every set on its own pass over the cells (no bundles)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import depecher_b200 as dp  # noqa: E402

eng = dp.Engine(device=0)
n, d, k, m = 10_000_000, 40, 30, 16
ds = eng.synthetic(n, d, 24, 2.5, 20261021)
reg = 1.0 * n * np.sqrt(d) / 1450.0
mus = []
for f in range(m):
    mu0, _ = eng.initialize_mu(ds, k, seed=1, fit_id=f)
    mus.append(mu0)
eng.allocate_many(ds, mus, False, want_labels=False, want_ari=False)
eng.synchronize()
eng.pass_timing(True)
t0 = time.perf_counter()
eng.allocate_many(ds, mus, False, want_labels=False, want_ari=False)
eng.synchronize()
dt = time.perf_counter() - t0
ms, cnt = eng.pass_timing(False)
print(f"sets kept together: {os.environ.get('DPR_TC_SETS', 'default')}   {m} sets x {n} cells: pass kernel {ms:.3f} ms = {ms/m:.4f} ms per set   (call {dt*1e3:.2f} ms)", flush=True)
