"""Multi-GPU host plumbing: one process per GPU, torch.distributed for the control plane.

Two ways the path shards:
  * row shard   — rank r owns cells [lo, hi); every Lloyd iteration the library allreduces the K x (D+1)
                  partial sums/counts (+ the changed-label count) with NCCL and every rank applies the
                  identical centre update.  dAllocate of a row shard needs no communication at all.
  * grid split  — the (worker, penalty) fits of one penalty-search set are independent
                  (src/Clusterer.cpp:349-384); ranks take them round-robin and only the small results
                  are gathered.
"""
from __future__ import annotations

import os

import numpy as np


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """contiguous, near-equal row ranges covering [0, n)"""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def grid_split(n_problems: int, rank: int, world: int) -> np.ndarray:
    """problem indices of this rank (round-robin: neighbouring penalties cost about the same)"""
    return np.arange(rank, n_problems, world)


def soft_threshold_update(sums: np.ndarray, counts: np.ndarray, reg: float) -> np.ndarray:
    """mu = min((S + reg/2)/n, max((S - reg/2)/n, 0)) with an empty cluster -> exact zero row
    (reevaluate_centers, src/Clusterer.cpp:103-107; the same rule finalize_kernel applies on the device)"""
    with np.errstate(divide="ignore", invalid="ignore"):
        lo = (sums - reg / 2) / counts[:, None]
        hi = (sums + reg / 2) / counts[:, None]
    m0 = np.where(lo > 0, lo, 0.0)
    return np.where(hi < m0, hi, m0)


def init_comm(engine, dist, device=None, peer=True) -> None:
    """rank 0 makes the NCCL unique id, torch.distributed broadcasts it, every rank joins the communicator"""
    rank, world = dist.get_rank(), dist.get_world_size()
    ids = [engine.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0, device=device)
    engine.comm_init(rank, world, ids[0])
    # ranks of one node: mailboxes in peer memory for the per-iteration exchange (falls back to ncclAllReduce when unavailable)
    if peer and 1 < world <= 8 and os.environ.get("DPR_NO_PEER_EXCHANGE") is None and hasattr(engine, "comm_peer_handle"):
        ok, err = True, ""
        try:
            mine = engine.comm_peer_handle()
        except Exception as exc:   # noqa: BLE001
            mine, ok, err = None, False, repr(exc)
        handles = [None] * world
        dist.all_gather_object(handles, mine)
        if ok and all(h is not None for h in handles):
            try:
                engine.comm_peer_attach(handles)
            except Exception as exc:   # noqa: BLE001
                ok, err = False, repr(exc)
        else:
            ok = False
        # all ranks or none: a rank on the NCCL path cannot talk to one on the mailbox path
        oks = [None] * world
        dist.all_gather_object(oks, ok)
        if not all(oks):
            engine.comm_peer_detach()
            if rank == 0:
                import warnings
                warnings.warn(f"peer-memory exchange unavailable, using NCCL ({err or 'another rank failed'})")


def allreduce_partials(dist, sums: np.ndarray, counts: np.ndarray):
    """CPU/gloo twin of the device allreduce (tests): element-wise sum over ranks of the partial tables"""
    import torch
    t = torch.from_numpy(np.concatenate([sums.ravel(), counts.astype(np.float64)]))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    out = t.numpy()
    return out[:sums.size].reshape(sums.shape), np.rint(out[sums.size:]).astype(np.int64)
