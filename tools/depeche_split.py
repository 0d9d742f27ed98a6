"""The penalty search of depeche() on 10M x 40 with the problems of every set split over the ranks (run under torchrun on N GPUs;
DPR_TIMING=1 prints every rank's per-phase times of each set):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29513 tools/depeche_split.py"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import depecher_b200 as dp  # noqa: E402
from depecher_b200 import rapi, shard  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
eng = dp.Engine(device=local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    shard.init_comm(eng, dist, device=torch.device("cuda", local))
n, d, k = int(os.environ.get("N", 10_000_000)), 40, 30
ds = eng.synthetic(n, d, 24, 2.5, 20261022)
for rep in range(2):
    eng.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    opt = rapi.depechePenaltyOpt(ds, k=k, maxIter=100, minARIImprovement=0.01, sampleSize=10000, penalties=2.0 ** np.arange(0, 5.5, 0.5), optimARI=0.95,
                                 nCores=10, engine=eng, seed=20261022, disableWarnings=True, dist=dist if world > 1 else None)
    dt = time.perf_counter() - t0
    if rank == 0:
        print(json.dumps({"n_gpus": world, "penalty_opt_s": dt, "workerIterations": opt["workerIterations"], "bestPenalty": opt["bestPenalty"]}), flush=True)
ds.close()
if world > 1:
    eng.comm_destroy()
    dist.destroy_process_group()
