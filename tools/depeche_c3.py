"""BASELINE configs[2]: the depeche() hot path on synthetic 10M cells x 40 markers with the full default penalty sweep.
Pre-scaling is skipped (the generator already divides by the global sd; SURVEY f-3 keeps it on the host): the timed
part is depechePenaltyOpt (all bootstrap fits + full-data scorings) + depecheClusterCenterSelection + the final
full-data dAllocate, with X resident in HBM.   python tools/depeche_c3.py [--n 10000000] [--d 40] [--ncores 10]"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import depecher_b200 as dp  # noqa: E402
from depecher_b200 import rapi  # noqa: E402


def run(eng, n, d, k, ncores, groups, amp, seed, verbose=True):
    ds = eng.synthetic(n, d, groups, amp, seed)
    eng.synchronize()
    t0 = time.perf_counter()
    launches0 = eng.kernel_launches()
    opt = rapi.depechePenaltyOpt(ds, k=k, maxIter=100, minARIImprovement=0.01, sampleSize=10000, penalties=2.0 ** np.arange(0, 5.5, 0.5),
                                 optimARI=0.95, nCores=ncores, engine=eng, seed=seed, disableWarnings=True, verbose=verbose)
    t1 = time.perf_counter()
    rows = np.random.default_rng(seed).integers(0, n, 10000)
    sel = eng.gather(ds, rows)
    r = rapi.depecheClusterCenterSelection(opt["bestClusterCenters"], sel, k, ncores, engine=eng)
    sel.close()
    t2 = time.perf_counter()
    full = np.zeros((r["clusterCenters"].shape[0], d))
    full[:, r["cols"]] = r["clusterCenters"]
    labels = rapi.dAllocate(ds, {"clusterCenters": full, "logCenterSd": False}, engine=eng)
    t3 = time.perf_counter()
    out = {"n": n, "d": d, "k": k, "nCores": ncores, "bestPenalty": opt["bestPenalty"], "workerIterations": opt["workerIterations"],
           "clusters": int(r["clusterCenters"].shape[0]), "penalty_opt_s": t1 - t0, "selection_s": t2 - t1, "final_allocate_s": t3 - t2,
           "total_s": t3 - t0, "kernel_launches": eng.kernel_launches() - launches0,
           "ARI": [round(float(v), 3) for v in opt["meanOptimDf"]["ARI"]], "nClust": [round(float(v), 2) for v in opt["meanOptimDf"]["nClust"]],
           "cluster_sizes": np.bincount(labels)[1:].tolist()}
    ds.close()
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=10_000_000)
    ap.add_argument("--d", type=int, default=40)
    ap.add_argument("--k", type=int, default=30)
    ap.add_argument("--ncores", type=int, default=10)
    ap.add_argument("--groups", type=int, default=24)
    a = ap.parse_args()
    eng = dp.Engine(device=0)
    import json
    print(json.dumps(run(eng, a.n, a.d, a.k, a.ncores, a.groups, 2.5, 20261021)))
