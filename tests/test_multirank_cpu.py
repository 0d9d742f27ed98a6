"""world_size-2 gloo test of the row-shard path's host logic: shards cover the rows, per-rank partial sums and
counts allreduce to the full-data values, and every rank derives the same centres as the oracle on all rows."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from depecher_b200 import shard


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(11)          # same data on every rank; each keeps only its rows
        n, d, k, reg = 1003, 7, 6, 9.5
        X = rng.normal(size=(n, d))
        idx = rng.integers(0, k, n)
        idx[idx == 2] = 3                        # an empty cluster
        lo, hi = shard.shard_range(n, rank, world)
        sums = np.zeros((k, d))
        counts = np.zeros(k, np.int64)
        np.add.at(sums, idx[lo:hi], X[lo:hi])
        np.add.at(counts, idx[lo:hi], 1)
        S, C = shard.allreduce_partials(dist, sums, counts)
        mu = shard.soft_threshold_update(S, C, reg)
        ids = [b"x" * 128 if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)   # the unique-id plumbing init_comm uses
        q.put((rank, lo, hi, mu, C, ids[0], shard.grid_split(23, rank, world)))
    finally:
        dist.destroy_process_group()


def test_row_shard_two_ranks(oracle):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == 1003
    rng = np.random.default_rng(11)
    X = rng.normal(size=(1003, 7))
    idx = rng.integers(0, 6, 1003)
    idx[idx == 2] = 3
    want, wcnt = oracle.reevaluate_centers(X, idx, 6, 9.5)
    for _rank, _lo, _hi, mu, C, uid, _split in res:
        assert np.array_equal(C, wcnt)
        assert (mu[2] == 0).all()
        np.testing.assert_allclose(mu, want, rtol=1e-12, atol=1e-14)
        assert np.array_equal(mu == 0, want == 0)
        assert uid == b"x" * 128
    assert sorted(np.concatenate([r[6] for r in res]).tolist()) == list(range(23))


def test_shard_ranges_cover_rows():
    for n in (1, 7, 64, 1000, 10_000_001):
        for world in (1, 2, 3, 8):
            edges = [shard.shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1


def _grid_worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from depecher_b200 import rapi
        from oracle_engine import OracleEngine
        x = rapi.generateBimodalData(observations=200, dataCols=10, rng=np.random.default_rng(6))
        xs = x["samples"] / np.std(x["samples"], ddof=1)
        r = rapi.depechePenaltyOpt(xs, k=12, maxIter=20, minARIImprovement=0.01, sampleSize=100, penalties=[0, 4, 16, 64], optimARI=0.95, nCores=4,
                                   engine=OracleEngine(), seed=2, disableWarnings=True, dist=dist)
        q.put((rank, r["bestPenalty"], r["meanOptimDf"]["ARI"].tolist(), r["meanOptimDf"]["nClust"].tolist(), len(r["bestClusterCenters"])))
    finally:
        dist.destroy_process_group()


def test_grid_split_two_ranks_same_state_machine():
    """penalty/restart grid split over 2 ranks: every rank ends with the identical search result"""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_grid_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1:] == res[1][1:]
    best, ari, ncl, ncent = res[0][1:]
    pos = [0, 4, 16, 64].index(best)
    assert ncl[pos] == 2 and ari[pos] == 1 and ncent >= 2 * 20


class _FakeEngine:
    """records what shard.init_comm does to an engine; `fail_attach_on` makes one rank's mailbox attach fail"""

    def __init__(self, rank, fail_attach_on=None, fail_handle_on=None):
        self.rank, self.fail_attach_on, self.fail_handle_on = rank, fail_attach_on, fail_handle_on
        self.log = []

    def comm_unique_id(self):
        return b"u" * 128

    def comm_init(self, rank, world, uid):
        self.log.append(("init", rank, world, uid))

    def comm_peer_handle(self):
        if self.rank == self.fail_handle_on:
            raise RuntimeError("no IPC here")
        return bytes([self.rank]) * 64

    def comm_peer_attach(self, handles):
        if self.rank == self.fail_attach_on:
            raise RuntimeError("peer access denied")
        self.log.append(("attach", tuple(handles)))

    def comm_peer_detach(self):
        self.log.append(("detach",))


def _init_comm_worker(rank, world, port, q, fail_attach_on, fail_handle_on):
    import warnings
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        eng = _FakeEngine(rank, fail_attach_on, fail_handle_on)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            shard.init_comm(eng, dist)
        q.put((rank, eng.log))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("fail_attach_on,fail_handle_on", [(None, None), (1, None), (None, 0)])
def test_init_comm_peer_mailboxes_all_ranks_or_none(fail_attach_on, fail_handle_on):
    """shard.init_comm: every rank joins the communicator with rank 0's id; the peer-memory mailboxes are attached on ALL ranks (handles in
    rank order) or, if any rank fails, detached on all — a rank on the NCCL path cannot talk to one on the mailbox path"""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_init_comm_worker, args=(r, world, port, q, fail_attach_on, fail_handle_on)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, log in res:
        assert log[0] == ("init", rank, world, b"u" * 128)
        kinds = [e[0] for e in log[1:]]
        if fail_attach_on is None and fail_handle_on is None:
            assert kinds == ["attach"] and log[1][1] == (bytes([0]) * 64, bytes([1]) * 64)
        else:
            assert kinds[-1] == "detach" and kinds.count("detach") == 1
