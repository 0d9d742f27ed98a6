"""depeche(testData[, 2:15]) on the GPU (BASELINE configs[0]) with the library's phase times (run on a B200): DPR_TIMING=1 python tools/depeche_testdata_probe.py"""
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import depecher_b200 as dp  # noqa: E402
from depecher_b200 import rapi  # noqa: E402

eng = dp.Engine(device=0)
z = np.load(os.path.join(ROOT, "tests", "golden", "testdata.npz"))
X = z["X_1e4"].astype(np.float64) / 1e4
warnings.simplefilter("ignore")
rapi.depeche(X, nCores=10, engine=eng, seed=3)
for rep in range(3):
    t0 = time.perf_counter()
    out = rapi.depeche(X, nCores=10, engine=eng, seed=4 + rep)
    print(f"depeche(testData): {time.perf_counter() - t0:.4f} s, {out['penaltyOptList']['workerIterations']} worker iterations, {out['clusterCenters'].shape[0]} clusters", flush=True)
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
rapi.depeche(X, nCores=10, engine=eng, seed=4)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
