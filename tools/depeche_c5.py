"""BASELINE configs[4]: synthetic scRNA-seq shape, 2M cells x 50 PCs, the (worker x penalty) grid of every penalty-search
set split over the ranks (each rank holds X; no data-path collective), exact ARI scoring.  Launch with torchrun:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/depeche_c5.py"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import depecher_b200 as dp  # noqa: E402
from depecher_b200 import rapi  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = dp.Engine(device=local)
n, d, k, workers = 2_000_000, 50, 30, 24
ds = eng.synthetic(n, d, 16, 2.5, 20261023)          # the same matrix on every rank (same seed, same rows)
eng.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
opt = rapi.depechePenaltyOpt(ds, k=k, maxIter=100, minARIImprovement=0.01, sampleSize=10000, penalties=2.0 ** np.arange(0, 5.5, 0.5),
                             optimARI=0.95, nCores=workers, engine=eng, seed=7, disableWarnings=True, dist=dist if world > 1 else None)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
dt = time.perf_counter() - t0
if rank == 0:
    print(json.dumps({"config": "2M x 50, K=30, 11 penalties x 24 workers per set", "n_gpus": world, "seconds": dt, "bestPenalty": opt["bestPenalty"],
                      "workerIterations": opt["workerIterations"], "ARI": [round(float(v), 3) for v in opt["meanOptimDf"]["ARI"]],
                      "nClust": [round(float(v), 2) for v in opt["meanOptimDf"]["nClust"]]}))
ds.close()
if world > 1:
    dist.destroy_process_group()
