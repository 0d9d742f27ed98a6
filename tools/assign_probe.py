"""Assignment pass on a resident matrix (run on a B200): ms per pass, algorithmic GB/s (4 d + 4 bytes per cell), fraction of the HBM
roofline.  DPR_SCORE_OLD=1 keeps assignments on tc_pass_kernel (the round-1 kernel).   N=..., D=..., K=... in the environment."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import depecher_b200 as dp  # noqa: E402

eng = dp.Engine(device=0)
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
for n, d, k in [(10_000_000, 40, 30), (10_000_000, 40, 8), (10_000_000, 15, 20), (1_000_000, 15, 20), (10_000_000, 50, 30), (2_000_000, 50, 30)]:
    ds = eng.synthetic(n, d, 12, 2.5, 20261021)
    mu, _ = eng.initialize_mu(ds, k, seed=1)
    eng.assign_iterations(ds, mu, True, 3)
    eng.pass_timing(True)
    eng.assign_iterations(ds, mu, True, 20)
    ms, cnt = eng.pass_timing(False)
    ms /= cnt
    gbs = n * (4.0 * d + 4) / (ms * 1e-3) / 1e9
    print(f"{n:>9} x {d:>2}, K={k:>2}: {ms:.4f} ms per pass, {gbs:7.0f} GB/s = {100 * gbs / peak:5.1f} % of the measured HBM copy bandwidth", flush=True)
    ds.close()
