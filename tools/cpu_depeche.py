"""CPU wall-time of depeche() on the reference's own engine — the CPU arm of BASELINE configs[0] (bench.py runs this in a
subprocess; TEST / BASELINE INFRASTRUCTURE, the product never imports it).

The engine is oracle/_ref: /root/reference/src/Clusterer.cpp compiled where it lies (the oracle restatement when _ref was not
built).  The callers are the same Python mirror of the R code the GPU arm uses (depecher_b200/rapi.py), and the fan-out is the
reference's: `nCores = min(10, floor(0.875 * cores))` worker PROCESSES (R/depeche.R:169-174), each running
grid_search(dataMat, k, usedPenalties, 1, sampleSize, i) per set (R/depechePenaltyOpt.R:50-55).  Unlike the reference's SOCK
workers, which receive the matrix again on every call, these workers inherit it once through fork: the figure is a floor.

    python tools/cpu_depeche.py [--data testdata|<rows>x<cols>] [--ncores N]    ->  one JSON line"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from depecher_b200 import rapi  # noqa: E402
from oracle.oracle import Oracle, Ref  # noqa: E402

_X = None        # the data matrix of the current search (set before the pool forks)
_ENG = None


def _engine():
    global _ENG
    if _ENG is None:
        _ENG = Ref() if Ref.available() else Oracle()
    return _ENG


def _worker(args):
    k, reg, boot, seed = args
    eng = _engine()
    if isinstance(eng, Ref):
        eng.set_time(int(time.time()))
        g = eng.grid_search(_X, k, reg, 1, boot, seed_off_set=seed)
    else:
        eng.set_rng(1, seed)
        g = eng.grid_search(_X, k, reg, 1, boot)
    return g["d"], g["n"], [p + p for p in g["c"]]


class WorkersEngine:
    """the four entry points of the reference on its CPU engine, grid_search fanned out over worker processes"""

    def __init__(self, ncores):
        self.ncores, self.pool, self.kind = ncores, None, "reference" if Ref.available() else "port"

    def _pool_for(self, X):
        global _X
        if self.pool is None or _X is not X:
            if self.pool is not None:
                self.pool.terminate()
            _X = X
            self.pool = mp.get_context("fork").Pool(self.ncores)
        return self.pool

    def grid_search_each(self, X, k, reg, iterations, boot, seed_off_set=0):
        out = self._pool_for(X).map(_worker, [(k, np.asarray(reg), boot, seed_off_set * 1000 + i + 1) for i in range(int(iterations))])
        z, m, c = [o[0] for o in out], [o[1] for o in out], [o[2] for o in out]
        return {"d": z, "z": z, "n": m, "m": m, "c": c}

    def sparse_k_means(self, X, k, reg, no_zero, seed_off_set=0, want_labels=True):
        eng = _engine()
        if isinstance(eng, Ref):
            r = eng.sparse_k_means(X, k, reg, bool(no_zero), seed_off_set)
        else:
            eng.set_rng(1, seed_off_set)
            r = eng.sparse_k_means(X, k, reg, bool(no_zero))
        return {"i": r["i"], "c": r["c"], "n": r["n"]}

    def allocate_points(self, X, mu, no_zero):
        return {"i": _engine().allocate_points(X, mu, bool(no_zero))}

    def rand_index(self, a, b, k):
        eng = _engine()
        return eng.rand_index(a, b, k) if isinstance(eng, Ref) else eng.rand_index_exact(a, b, k)

    def close(self):
        if self.pool is not None:
            self.pool.terminate()
            self.pool = None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--data", default="testdata")
    ap.add_argument("--ncores", type=int, default=0)
    a = ap.parse_args()
    ncores = a.ncores or rapi.default_ncores()
    if a.data == "testdata":
        z = np.load(os.path.join(ROOT, "tests", "golden", "testdata.npz"))
        X = z["X_1e4"].astype(np.float64) / 1e4
        what = "depeche(testData[, 2:15]): 20000 cells x 14 markers, K = 30, default penalty grid, sampleSize 10000"
    else:
        rows, cols = (int(v) for v in a.data.split("x"))
        rng = np.random.default_rng(1)
        cent = rng.choice([-2.5, 0.0, 2.5], size=(12, cols))
        X = cent[rng.integers(0, 12, rows)] + rng.normal(size=(rows, cols))
        what = f"depeche(synthetic {rows} x {cols})"
    eng = WorkersEngine(ncores)
    t0 = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = rapi.depeche(X, nCores=ncores, engine=eng, seed=int(time.time()) & 0xffff)
    dt = time.perf_counter() - t0
    eng.close()
    print(json.dumps({"seconds": dt, "what": what, "kind": eng.kind, "workers": ncores, "host_cores": os.cpu_count(),
                      "workerIterations": out["penaltyOptList"]["workerIterations"], "bestPenalty": out["penaltyOptList"]["bestPenalty"],
                      "clusters": int(out["clusterCenters"].shape[0])}), flush=True)


if __name__ == "__main__":
    main()
