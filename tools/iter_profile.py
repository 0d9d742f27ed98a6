"""Per-iteration cost of a fresh fit (run on a B200): fused-pass ms and labels changed for iterations 0..n-1."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import depecher_b200 as dp  # noqa: E402

from depecher_b200 import _capi  # noqa: E402
eng = dp.Engine(device=0, path=os.environ.get("DPR_LIB", _capi.LIB_PATH))   # DPR_LIB=<path>: another build of the library (A/B on one box)
n, d, k = 10_000_000, 40, 30
ds = eng.synthetic(n, d, 24, 2.5, 20261021)
reg = 1.0 * n * np.sqrt(d) / 1450.0
mu0, _ = eng.initialize_mu(ds, k, seed=1, fit_id=0)
eng.lloyd_iterations(ds, mu0, reg, True, -1, 3)
prev = 0.0
for it in range(1, 26):
    eng.pass_timing(True)
    eng.recheck_count(True)
    mu, ch = eng.lloyd_iterations(ds, mu0, reg, True, -1, it)
    ms, cnt = eng.pass_timing(False)
    rc = eng.recheck_count(True)
    print(f"iterations 0..{it-1}: pass total {ms:7.3f} ms  -> iteration {it-1}: {ms - prev:6.3f} ms   changed {ch:9d} ({100.0*ch/n:5.2f} %)  alive {int((np.abs(mu).sum(1)>0).sum())}  exact rechecks (all {it} iterations) {rc}", flush=True)
    prev = ms
