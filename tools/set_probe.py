"""One set of the penalty search with NC workers on one GPU (run on a B200; DPR_TIMING=1 prints the phase times): what one rank of a grid
split gets at 8 GPUs is NC = 1-2.     NC=2 DPR_TIMING=1 python tools/set_probe.py"""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import depecher_b200 as dp
from depecher_b200 import rapi
eng = dp.Engine(device=0)
ds = eng.synthetic(10_000_000, 40, 24, 2.5, 20261022)
nc = int(os.environ.get("NC", 1))
for rep in range(int(os.environ.get("REPS", 3))):
    eng.synchronize(); t0 = time.perf_counter()
    r = eng.grid_search_each(ds, [30], (2.0 ** np.arange(0, 5.5, 0.5)) * 10000 * np.sqrt(40) / 1450.0, nc, 10000, 20261022 + rep)
    print(json.dumps({"nCores": nc, "set_s": time.perf_counter() - t0}), flush=True)
