"""The launches of a few Lloyd iterations of a fresh fit (10M x 40, K = 30), for `ncu --metrics gpu__time_duration.sum`:
which kernels a step consists of and what each costs."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import depecher_b200 as dp  # noqa: E402

from depecher_b200 import _capi  # noqa: E402
eng = dp.Engine(device=0, path=os.environ.get("DPR_LIB", _capi.LIB_PATH))
n, d, k = 10_000_000, 40, 30
ds = eng.synthetic(n, d, 24, 2.5, 20261021)
reg = 1.0 * n * np.sqrt(d) / 1450.0
mu0, _ = eng.initialize_mu(ds, k, seed=1, fit_id=0)
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 6
mu, ch = eng.lloyd_iterations(ds, mu0, reg, True, -1, iters)
print("changed in the last iteration:", ch)
