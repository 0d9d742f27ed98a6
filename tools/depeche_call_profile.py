"""Where the wall time of depeche() on a 10M x 40 host matrix goes (run on a B200): cProfile of the second call."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import depecher_b200 as dp  # noqa: E402
from depecher_b200 import rapi  # noqa: E402

eng = dp.Engine(device=0)
n, d = 10_000_000, 40
ds = eng.synthetic(n, d, 24, 2.5, 20261021, scale=0.0)
Xh = np.asfortranarray(eng.read(ds, 0, n))
ds.close()
import warnings
warnings.simplefilter("ignore")
for rep in range(2):
    t0 = time.perf_counter()
    r = rapi.depeche(Xh, nCores=10, engine=eng, seed=5)
    print("depeche() wall", time.perf_counter() - t0, "clusters", r["clusterCenters"].shape[0], flush=True)
pr = cProfile.Profile()
pr.enable()
r = rapi.depeche(Xh, nCores=10, engine=eng, seed=5)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
