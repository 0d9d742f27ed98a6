"""Where a Lloyd pass spends its time (run on a B200): the same centres (iteration 8 of a fresh fit, 30 clusters alive, ~1.4 % of
the cells still moving) through the assignment-only pass, the assignment pass that also compares with the previous labels,
and the Lloyd pass."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import depecher_b200 as dp  # noqa: E402

eng = dp.Engine(device=0)
n, d, k = 10_000_000, 40, 30
ds = eng.synthetic(n, d, 24, 2.5, 20261021)
reg = 1.0 * n * np.sqrt(d) / 1450.0
mu0, _ = eng.initialize_mu(ds, k, seed=1, fit_id=0)
mu8, ch = eng.lloyd_iterations(ds, mu0, reg, True, -1, 8)


def timed(fn, label):
    fn()
    eng.recheck_count(True)
    eng.pass_timing(True)
    fn()
    ms, cnt = eng.pass_timing(False)
    print(f"{label:46s} {ms / max(cnt, 1):7.4f} ms/pass   exact re-decisions per pass {eng.recheck_count(True) // max(cnt, 1)}", flush=True)


timed(lambda: eng.assign_iterations(ds, mu8, False, 5), "assign, centres of iteration 8")
timed(lambda: eng.assign_iterations(ds, mu8, True, 5), "assign + compare with previous labels")
timed(lambda: eng.assign_iterations(ds, mu0, False, 5), "assign, k-means++ centres")
for it in (9, 10):
    eng.pass_timing(True)
    eng.recheck_count(True)
    m, ch = eng.lloyd_iterations(ds, mu0, reg, True, -1, it)
    ms, cnt = eng.pass_timing(False)
    print(f"fresh fit, {it} iterations: {ms:7.3f} ms in {cnt} passes, last changed {ch}, re-decisions {eng.recheck_count(True)}", flush=True)
