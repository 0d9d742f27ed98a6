"""Row-sharded fit on N GPUs (run under torchrun on a multi-GPU box): the peer-memory exchange, the NCCL allreduce and a single GPU
holding all rows must agree — identical labels and iteration counts, centres to rounding (the three add the ranks' tables in different orders).
The exit code reflects the peer-memory path against the NCCL path; the one-GPU comparison is printed.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29555 tools/multirank_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import depecher_b200 as dp  # noqa: E402
from depecher_b200 import shard  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = dp.Engine(device=local)
n_local, d, k = 400_000, 40, 30
reg = 1.0 * n_local * world * np.sqrt(d) / 1450.0
# one common scale for every shard (scale = 0 lets the generator divide by the sd of the rows it makes): rank 0's, broadcast
probe = eng.synthetic(n_local, d, 24, 2.5, 77, row0=0, scale=0.0)
sc = torch.tensor([probe.scale], dtype=torch.float64).cuda()
probe.close()
dist.broadcast(sc, src=0)
scale = float(sc.item())
ds = eng.synthetic(n_local, d, 24, 2.5, 77, row0=rank * n_local, scale=scale)
mu0, _ = eng.initialize_mu(ds, k, seed=3, fit_id=0)
t = torch.from_numpy(np.ascontiguousarray(mu0)).cuda()
dist.broadcast(t, src=0)
mu0 = t.cpu().numpy()
res = {}
for mode in ("peer", "nccl"):
    shard.init_comm(eng, dist, device=torch.device("cuda", local), peer=(mode == "peer"))
    res[mode] = eng.find_centers_from(ds, mu0, reg, True)
    eng.comm_destroy()
ok = np.array_equal(res["peer"]["i"], res["nccl"]["i"]) and res["peer"]["n_iter"] == res["nccl"]["n_iter"]
err = float(np.abs(res["peer"]["c"] - res["nccl"]["c"]).max())
# every rank must hold the same centres (bitwise) on the peer path
c = torch.from_numpy(np.ascontiguousarray(res["peer"]["c"])).cuda()
c0 = c.clone()
dist.broadcast(c0, src=0)
same = bool(torch.equal(c, c0))
line = f"rank {rank}: peer vs nccl labels/iterations equal {ok}, max |centre diff| {err:.3e}, centres identical to rank 0's: {same}, iterations {res['peer']['n_iter']}"
if rank == 0:   # the same rows on one GPU
    full = eng.synthetic(n_local * world, d, 24, 2.5, 77, row0=0, scale=scale)
    one = eng.find_centers_from(full, mu0, reg, True)
    lab_ok = np.array_equal(one["i"][:n_local], res["peer"]["i"])
    line += f" | one GPU with all rows: labels of shard 0 equal {lab_ok}, iterations {one['n_iter']}, max |centre diff| {float(np.abs(one['c'] - res['peer']['c']).max()):.3e}"
print(line, flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok and same and err < 1e-10 else 1)
