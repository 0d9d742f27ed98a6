import os, sys, time, warnings
import numpy as np
sys.path.insert(0, "/root/repo")
import depecher_b200 as dp
from depecher_b200 import rapi
eng = dp.Engine(device=0)
n, d, k = 10_000_000, 40, 30
ds = eng.synthetic(n, d, 24, 2.5, 20261021, scale=1.0)
reg = 1.0 * n * np.sqrt(d) / 1450.0
mu0, _ = eng.initialize_mu(ds, k, seed=20261021, fit_id=0)
eng.lloyd_iterations(ds, mu0, reg, True, -1, 20)
Xh = np.asfortranarray(eng.read(ds, 0, n))
eng.synchronize()
for _ in range(2):
    eng.sparse_k_means(Xh, k, reg, True, seed_off_set=1)
warnings.simplefilter("ignore")
rapi.depeche(np.asfortranarray(Xh[:200_000]), nCores=10, engine=eng, seed=20261021)
for rep in range(3):
    t0 = time.perf_counter()
    dep = rapi.depeche(Xh, nCores=10, engine=eng, seed=20261022)
    print("depeche()", time.perf_counter() - t0, dep["penaltyOptList"]["workerIterations"], flush=True)
