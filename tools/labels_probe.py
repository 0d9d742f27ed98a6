"""dAllocate with the labels copied to the host (run on a B200): call time on 10M and 100M cells x 40 markers."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import depecher_b200 as dp  # noqa: E402

eng = dp.Engine(device=0)
rng = np.random.default_rng(1)
mu = rng.normal(size=(24, 40)) * 2.5
for n in (10_000_000, 100_000_000):
    ds = eng.synthetic(n, 40, 24, 2.5, 20261021)
    for rep in range(4):
        eng.synchronize()
        t0 = time.perf_counter()
        lab = eng.allocate_points(ds, mu, 1, base=1)["i"]
        t1 = time.perf_counter()
        print(f"n = {n}: allocate_points with {lab.nbytes/1e6:.0f} MB of labels to the host {1e3*(t1-t0):.2f} ms", flush=True)
        del lab
    ds.close()
