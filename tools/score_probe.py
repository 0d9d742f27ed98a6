"""Scoring throughput (run on a B200): M centre sets against the same 10M x 40 cells in ONE launch, as the penalty search does
(Clusterer.cpp:370-371).  DPR_SCORE_OLD=1 keeps the round-1 bundled instance of tc_pass_kernel; DPR_LIB=<path> loads another build
of the library (e.g. the -DDPR_SC_PROFILE one that prints per-phase cycle counts of one CTA)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import depecher_b200 as dp  # noqa: E402
from depecher_b200 import _capi  # noqa: E402

eng = dp.Engine(device=0, path=os.environ.get("DPR_LIB", _capi.LIB_PATH))
n, d, k, m = int(os.environ.get("N", 10_000_000)), int(os.environ.get("D", 40)), int(os.environ.get("K", 30)), int(os.environ.get("M", 16))
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
print(f"{m} sets x {n} cells x {d} markers, K={k}: pass kernel {ms:.3f} ms = {ms/m:.4f} ms per set   (call {dt*1e3:.2f} ms)", flush=True)
