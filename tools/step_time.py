"""The bench's timed region on its own (run on a B200): `steps` Lloyd iterations of a fresh fit on 10M x 40, K = 30, repeated; wall
clock around the call (it ends with a synchronising read of the centres) and the pass kernel's own time.
    python tools/step_time.py [steps] [repeats]          DPR_NO_ROWS_COPY=1: delta gathers from the tile-major matrix (A/B)"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import depecher_b200 as dp  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
from depecher_b200 import _capi  # noqa: E402
eng = dp.Engine(device=0, path=os.environ.get("DPR_LIB", _capi.LIB_PATH))   # DPR_LIB=<path>: another build of the library (A/B on one box)
n, d, k = 10_000_000, 40, 30
ds = eng.synthetic(n, d, 24, 2.5, 20261021)
reg = 1.0 * n * np.sqrt(d) / 1450.0
mu0, _ = eng.initialize_mu(ds, k, seed=20261021, fit_id=0)
eng.lloyd_iterations(ds, mu0, reg, True, -1, 5)
for _ in range(3):
    eng.lloyd_iterations(ds, mu0, reg, True, -1, 30)
out = []
for r in range(reps):
    eng.synchronize()
    eng.pass_timing(True)
    t0 = time.perf_counter()
    mu, ch = eng.lloyd_iterations(ds, mu0, reg, True, -1, steps)
    dt = time.perf_counter() - t0
    ms, cnt = eng.pass_timing(False)
    out.append((dt * 1e3 / steps, ms / max(cnt, 1)))
print(f"{steps} steps x {reps}: ms/step (wall) " + " ".join(f"{a:.4f}" for a, _ in out) + "   pass ms/launch " + " ".join(f"{b:.4f}" for _, b in out) +
      f"   changed last {ch}", flush=True)
