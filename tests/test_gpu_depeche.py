"""GPU end-to-end: the reference's hot-path testthat properties and its stored depeche(testData[, 2:15]) result,
through the Python mirror of the R callers over the CUDA library."""
import numpy as np
import pytest

from depecher_b200 import rapi

pytestmark = pytest.mark.gpu


def test_depeche_bimodal(engine):
    """test_depeche.R:4-10"""
    x = rapi.generateBimodalData(observations=100, dataCols=150, rng=np.random.default_rng(1))
    out = rapi.depeche(x["samples"], maxIter=8, nCores=2, engine=engine, seed=3)
    cv = out["clusterVector"]
    assert cv.max() == 2
    assert np.array_equal(x["ids"], cv) or np.array_equal(x["ids"] % 2 + 1, cv)


def test_depeche_sparsity(engine):
    """test_depeche.R:14-26 with the reference's own sizes: 500 cells x 100 markers, 5 modes, K = 3k = 90"""
    x = rapi.generateSparseData(observations=500, rng=np.random.default_rng(2))
    out = rapi.depeche(x["samples"], maxIter=8, nCores=2, center=False, engine=engine, seed=5)
    binTruth = x["centers"] == 0
    binGot = np.ones((out["clusterCenters"].shape[0], 100), bool)
    binGot[:, out["columns"]] = out["clusterCenters"] == 0
    for i in range(1, 6):
        row = np.flatnonzero(binGot.sum(axis=1) == i)
        assert len(row) == 1 and np.array_equal(binTruth[i - 1], binGot[row[0]])


def test_dallocate(engine):
    """test_dAllocate.R:4-11 and :16-23"""
    rng = np.random.default_rng(3)
    x = rapi.generateBimodalData(observations=20, dataCols=150, rng=rng)
    centers = np.vstack([rng.random(150) + 50, rng.random(150) - 50])
    assert np.array_equal(rapi.dAllocate(x["samples"], {"clusterCenters": centers, "logCenterSd": False}, engine=engine), x["ids"])
    dep = rapi.depeche(x["samples"], center=False, nCores=2, engine=engine, seed=1)
    alloc = rapi.dAllocate(x["samples"][:, dep["columns"]], {"clusterCenters": dep["clusterCenters"],
                                                             "logCenterSd": [False, False, dep["logCenterSd"][2]]}, engine=engine,
                           columns=np.arange(len(dep["columns"])))
    assert engine.rand_index(dep["clusterVector"], alloc, 100) == 1.0


def test_doptpenalty(engine):
    """test_dOptPenalty.R:4-22"""
    x = rapi.generateBimodalData(observations=200, dataCols=10, rng=np.random.default_rng(6))
    xs = x["samples"] / np.std(x["samples"], ddof=1)
    r = rapi.dOptPenalty(xs, k=30, maxIter=20, minARIImprovement=0.01, sampleSize=100, penalties=[0, 2, 4, 8, 16, 32, 128], optimARI=0.95,
                         nCores=2, engine=engine, seed=2, disableWarnings=True)
    pos = r["bestPenaltyPosition"]
    assert r["meanOptimDf"]["nClust"][pos] == 2 and r["meanOptimDf"]["ARI"][pos] == 1


def test_testdata_against_stored_depeche_result(engine, testdata):
    """BASELINE configs[0]: depeche(testData[, 2:15]) with the default penalty grid.  The stored run
    (data/testDataDepeche.RData) is time-seeded, so the comparison is distributional: same best penalty
    neighbourhood, 8 clusters, ARI curve and cluster-count curve close to the stored ones, same partition."""
    out = rapi.depeche(testdata["X"], nCores=7, engine=engine, seed=11)
    opt = out["penaltyOptList"]
    pen = opt["allPenalties"]["penalty"]
    assert np.allclose(pen, testdata["stored_penalties"])
    ari, ncl = opt["allPenalties"]["ARI"], opt["allPenalties"]["nClust"]
    assert opt["bestPenalty"] in (11.3, 16.0, 22.6)                                 # stored: 16
    hi = np.isin(pen, [11.3, 16.0, 22.6])
    assert np.all(ari[hi] > 0.96) and np.all(np.abs(ncl[hi] - 8) < 0.6)            # stored: .987 .991 .991 / 8.07 8.00 8.00
    assert ari[0] < 0.75 and ncl[0] > 26                                            # stored: .590 / 29.4 at penalty 1
    assert np.all(np.diff(ncl) < 0.8)                                               # cluster count falls with the penalty
    assert out["clusterCenters"].shape[0] == 8
    from sklearn.metrics import adjusted_rand_score
    assert adjusted_rand_score(out["clusterVector"], testdata["stored_clusterVector"]) > 0.93
    sizes = np.bincount(out["clusterVector"])[1:]
    assert np.all(np.diff(sizes) <= 0)                                              # clusters numbered by size (R/depeche.R:236-249)
    stored = np.bincount(testdata["stored_clusterVector"])[1:]
    assert np.all(np.abs(sizes - stored) < 0.05 * 20000)


def test_depeche_full_length_sampling_subset_keeps_row_order(engine):
    """a samplingSubset as long as the data (here a permutation) fits on the permuted rows, but the cluster vector is
    the allocation of the FULL matrix in its original order (dAllocate(inDataFrameScaled, ...), R/depeche.R:228-232)."""
    rng = np.random.default_rng(8)
    n, d = 12000, 6
    cent = np.array([[6.0] * d, [-6.0] * d, [6.0, -6.0] * (d // 2)])
    ids = np.sort(rng.integers(0, 3, n))                   # truth is sorted by row: a permuted result cannot match it
    X = cent[ids] + rng.normal(size=(n, d))
    out = rapi.depeche(X, samplingSubset=rng.permutation(n), nCores=2, maxIter=8, center=False, engine=engine, seed=4)   # (no blob at the origin)
    cv = out["clusterVector"]
    assert cv.shape == (n,)
    # every cluster lies inside one true blob (the blobs are 12 sd apart); labels in the subset's order would mix all three
    table = np.zeros((cv.max() + 1, 3), np.int64)
    np.add.at(table, (cv, ids), 1)
    assert (table.max(axis=1) == table.sum(axis=1)).all()


def test_ranked_allocation_equals_allocation_then_renumbering(engine):
    """dpr_dataset_allocate_points_ranked (allocation + R/depeche.R:236-252 in one call) against dAllocate followed by the host renumbering,
    on sizes on both sides of the chunked copy-out, with equal-sized clusters (ties keep the order of the old numbers) and an empty one"""
    from depecher_b200.rapi import _renumber_by_size
    rng = np.random.default_rng(11)
    for n in (3000, 700_000):
        cent = rng.normal(size=(7, 9)) * 3
        lab = np.repeat(np.arange(6), n // 6)            # six clusters of the same size: all ties
        lab[: n // 12] = 3                               # ... then one larger, one smaller
        X = cent[lab] + 0.05 * rng.normal(size=(len(lab), 9))
        mu = np.vstack([cent, np.full((1, 9), 100.0)])   # the last centre owns nothing
        ds = engine.dataset(X)
        got, order = engine.allocate_points_ranked(ds, mu, 1)
        plain = engine.allocate_points(ds, mu, 1, base=1)["i"]
        ds.close()
        want, worder, _ = _renumber_by_size(plain, mu.shape[0])
        assert np.array_equal(got, want) and np.array_equal(order + 1, worder)


def test_device_prescaling_matches_host(engine, testdata):
    """depecheLogCenterSd on the device (dpr_dataset_create_scaled) against the numpy restatement: same peak centres,
    same sd, scaled cells equal up to the last fp32 bit; also the mean / given-centre / log2 paths."""
    X = testdata["X"]
    want, winfo = rapi.depecheLogCenterSd(X)
    ds, info = engine.dataset_scaled(X)
    got = engine.read(ds, 0, len(X))
    ds.close()
    assert info[0] is False and np.allclose(info[1], winfo[1], rtol=0, atol=1e-12)
    assert abs(info[2] - winfo[2]) <= 1e-12 * winfo[2]
    w32 = want.astype(np.float32).astype(np.float64)
    assert np.max(np.abs(got - w32)) <= 2.5e-7 * np.max(np.abs(w32))
    assert (got == w32).mean() > 0.999
    # a matrix whose doubles are all fp32 values crosses PCIe as floats and is widened on the device (exactly): same result
    Xf = X.astype(np.float32).astype(np.float64)
    want, winfo = rapi.depecheLogCenterSd(Xf)
    ds, info = engine.dataset_scaled(Xf)
    got = engine.read(ds, 0, len(Xf))
    ds.close()
    assert np.allclose(info[1], winfo[1], rtol=0, atol=1e-12) and abs(info[2] - winfo[2]) <= 1e-12 * winfo[2]
    w32 = want.astype(np.float32).astype(np.float64)
    assert np.max(np.abs(got - w32)) <= 2.5e-7 * np.max(np.abs(w32)) and (got == w32).mean() > 0.999
    Xm = Xf.copy()
    Xm[12345, 3] += 1e-9          # one value that is not an fp32 value, late in the matrix: the upload starts over with doubles
    want, winfo = rapi.depecheLogCenterSd(Xm)
    ds, info = engine.dataset_scaled(Xm)
    got = engine.read(ds, 0, len(Xm))
    ds.close()
    assert np.allclose(info[1], winfo[1], rtol=0, atol=1e-12) and abs(info[2] - winfo[2]) <= 1e-12 * winfo[2]
    assert np.allclose(got, want, rtol=3e-7, atol=1e-7)
    for center in ("mean", False, list(np.arange(14.0))):
        want, winfo = rapi.depecheLogCenterSd(X, True, center, True)
        ds, info = engine.dataset_scaled(X, True, center, True)
        got = engine.read(ds, 0, 2000)
        ds.close()
        assert abs(info[2] - winfo[2]) <= 1e-12 * winfo[2]
        assert np.allclose(got, want[:2000], rtol=3e-7, atol=1e-7)
    heavy = np.abs(np.random.default_rng(0).standard_cauchy(size=(5000, 6))) ** 2      # kurtosis >> 100 -> log2(x + 1)
    want, winfo = rapi.depecheLogCenterSd(heavy)
    ds, info = engine.dataset_scaled(heavy)
    got = engine.read(ds, 0, 5000)
    ds.close()
    assert winfo[0] is True and info[0] is True
    assert np.allclose(info[1], winfo[1], atol=1e-9) and np.allclose(got, want, rtol=3e-6, atol=1e-6)


def test_device_prescaling_pinned_by_stored_depeche_result(engine, testdata):
    """the same pin as tests/test_host_logic.py::test_prescaling_pinned_by_stored_depeche_result through the device path: the matrix
    scaled by dpr_dataset_create_scaled (peak centres and sd found on the GPU), allocated onto the stored centres by the tensor-core
    pass — every one of the 20 000 labels of the stored R result"""
    X, cc, stored = testdata["X"], testdata["stored_clusterCenters"], testdata["stored_clusterVector"]
    ds, info = engine.dataset_scaled(X)
    lab = engine.allocate_points(ds, cc / info[2], True, base=1)["i"]
    ds.close()
    assert np.array_equal(lab, stored)


def test_peer_mailbox_setup_and_error_paths(engine):
    """dpr_comm_peer_*: a rank can always make its mailbox and hand out the 64-byte CUDA IPC handle; attaching needs a communicator of
    2..8 ranks and one handle per rank — anything else fails loudly and leaves the library on the NCCL / single-GPU path"""
    h = engine.comm_peer_handle()
    assert isinstance(h, bytes) and len(h) == 64 and any(h)
    with pytest.raises(Exception):
        engine.comm_peer_attach([h])            # world is 1: nothing to attach to
    with pytest.raises(Exception):
        engine.comm_peer_attach([h, h, h])      # not the communicator's size
    engine.comm_peer_detach()
    # the single-GPU path is untouched
    rng = np.random.default_rng(3)
    X = rng.normal(size=(3000, 8))
    res = engine.sparse_k_means(X, 5, 1.0, True, seed_off_set=2)
    assert res["i"].shape == (3000,)


def test_penalty_opt_one_call_equals_stepwise_and_split(engine, testdata):
    """SURVEY f-1: depechePenaltyOpt's loop inside the library (dpr_dataset_penalty_opt) against the same state machine driven from
    outside with one grid_search_each per set (what rapi did in round 1), and against the sets computed as three separate ranges of
    problems (the grid split: a range equals that range of the whole grid bit for bit)."""
    import depecher_b200 as dp
    Xs, _ = rapi.depecheLogCenterSd(testdata["X"])
    k, S, nCores, seed = 30, 5000, 4, 9
    pens = 2.0 ** np.arange(0, 5.5, 0.5)
    with engine.dataset(Xs) as ds:
        a = dp.PenaltyOpt(pens, k, ds.d, S, nCores, 100, 0.01, 0.95)
        engine.penalty_opt(ds, a, seed)
        ra = a.result()
        b = dp.PenaltyOpt(pens, k, ds.d, S, nCores, 100, 0.01, 0.95)
        c = dp.PenaltyOpt(pens, k, ds.d, S, nCores, 100, 0.01, 0.95)
        it = 1
        while b.used()[2]:
            used, regs, _ = b.used()
            res = engine.grid_search_each(ds, [k], regs, nCores, S, seed + 7919 * it)
            ari = np.array([np.array(res["z"][w]).ravel() for w in range(nCores)])
            ncl = np.array([np.array(res["m"][w]).ravel() for w in range(nCores)])
            cen = np.array([[[m for m in res["c"][w][q][:2]] for q in range(len(used))] for w in range(nCores)])
            b.feed(ari, ncl, cen)
            cells = nCores * len(used)                      # the same set as three ranges of its problems
            cuts = [0, cells // 3, cells // 3 + 1, cells]
            parts = [engine.grid_search_range(ds, k, regs, nCores, S, seed + 7919 * it, lo, hi - lo) for lo, hi in zip(cuts[:-1], cuts[1:])]
            ari_s = np.concatenate([p[0] for p in parts]).reshape(nCores, len(used))
            ncl_s = np.concatenate([p[1] for p in parts]).reshape(nCores, len(used))
            cen_s = np.concatenate([p[2] for p in parts]).reshape(nCores, len(used), 2, k, ds.d)
            # labels, and with them ARI and cluster counts, are identical; the centres agree to the rounding of their fp64 sums
            # (the order in which a fit's sums are added follows the launch's partition of the tiles over the SMs)
            assert np.array_equal(ari_s, ari, equal_nan=True) and np.array_equal(ncl_s, ncl)
            assert np.array_equal(cen_s == 0, cen == 0)
            np.testing.assert_allclose(cen_s, cen, rtol=1e-11, atol=1e-13)
            c.feed(ari_s, ncl_s, cen_s)
            it += 1
        rb, rc = b.result(), c.result()
        for r in (rb, rc):
            assert r["bestPenalty"] == ra["bestPenalty"] and r["workerIterations"] == ra["workerIterations"] and r["warn"] == ra["warn"]
            assert np.array_equal(r["meanARI"], ra["meanARI"], equal_nan=True) and np.array_equal(r["meanClust"], ra["meanClust"])
            assert len(r["bestClusterCenters"]) == len(ra["bestClusterCenters"]) == 2 * ra["workerIterations"]
            for x, y in zip(r["bestClusterCenters"], ra["bestClusterCenters"]):
                np.testing.assert_allclose(x, y, rtol=1e-11, atol=1e-13)
        assert all(np.array_equal(x, y) for x, y in zip(rb["bestClusterCenters"], ra["bestClusterCenters"]))   # same launches: bit for bit
        assert ra["bestPenalty"] in (11.3, 16.0, 22.6, 32.0) and ra["workerIterations"] >= 20
        # depecheClusterCenterSelection behind the ABI against the allocate_many route of round 1
        sel = engine.gather(ds, np.random.default_rng(0).integers(0, ds.n, 5000))
        pick, mean = engine.select_solution(sel, ra["bestClusterCenters"])
        masked = [np.where((m.sum(axis=1) == 0)[:, None], 0.0, m) for m in ra["bestClusterCenters"]]
        _, ari = engine.allocate_many(sel, masked, True, want_labels=False, want_ari=True)
        sel.close()
        np.testing.assert_allclose(mean, ari.mean(axis=0), rtol=1e-12)
        assert pick == int(np.flatnonzero(mean == np.nanmax(mean))[0])
        for o in (a, b, c):
            o.close()
