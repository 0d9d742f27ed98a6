"""CPU tests of the host layer: the reference's own hot-path testthat files transcribed
(tests/testthat/test_depeche.R, test_dAllocate.R, test_dOptPenalty.R) and run through the Python mirror of the
R callers (depecher_b200/rapi.py) with the CPU oracle injected as the engine, plus the pre-scaling helpers."""
import numpy as np
import pytest

from depecher_b200 import rapi
from oracle_engine import OracleEngine


@pytest.fixture(scope="module")
def eng(oracle):
    return OracleEngine(oracle)


def test_depeche_expected_output(eng):
    """test_depeche.R:4-10 — two far-apart blobs, 100 x 150: exactly two clusters equal to the truth up to a swap"""
    rng = np.random.default_rng(1)
    x = rapi.generateBimodalData(observations=100, dataCols=150, rng=rng)
    out = rapi.depeche(x["samples"], maxIter=8, nCores=2, createOutput=False, engine=eng, seed=3)
    cv = out["clusterVector"]
    assert cv.max() == 2
    assert np.array_equal(x["ids"], cv) or np.array_equal(x["ids"] % 2 + 1, cv)


def test_depeche_expected_sparsity(eng):
    """test_depeche.R:14-26 — centres come back with exact zeros in exactly the columns the truth has them"""
    rng = np.random.default_rng(2)
    x = rapi.generateSparseData(modeN=5, dataCols=40, observations=250, rng=rng)
    out = rapi.depeche(x["samples"], maxIter=8, createOutput=False, nCores=2, center=False, engine=eng, seed=5)
    cols = out["columns"]
    binTruth = x["centers"] == 0
    binGot = np.ones((out["clusterCenters"].shape[0], x["centers"].shape[1]), bool)
    binGot[:, cols] = out["clusterCenters"] == 0
    assert out["clusterCenters"].shape[0] == 5
    for i in range(1, 6):
        row = np.flatnonzero(binGot.sum(axis=1) == i)
        assert len(row) == 1
        assert np.array_equal(binTruth[i - 1], binGot[row[0]])


def test_dAllocate_expected_output(eng):
    """test_dAllocate.R:4-11"""
    rng = np.random.default_rng(3)
    cols = 150
    x = rapi.generateBimodalData(observations=20, dataCols=cols, rng=rng)
    centers = np.vstack([rng.random(cols) + 50, rng.random(cols) - 50])
    out = rapi.dAllocate(x["samples"], {"clusterCenters": centers, "logCenterSd": False}, engine=eng)
    assert np.array_equal(out, x["ids"])


def test_dAllocate_after_depeche_ari_is_one(eng):
    """test_dAllocate.R:16-23 — rand_index(depeche labels, dAllocate labels, k = 100) == 1 exactly"""
    rng = np.random.default_rng(4)
    x = rapi.generateBimodalData(observations=20, dataCols=150, rng=rng)
    dep = rapi.depeche(x["samples"], center=False, nCores=2, createOutput=False, engine=eng, seed=1)
    model = {"clusterCenters": dep["clusterCenters"], "logCenterSd": dep["logCenterSd"]}
    alloc = rapi.dAllocate(x["samples"][:, dep["columns"]], {"clusterCenters": dep["clusterCenters"],
                                                             "logCenterSd": [False, False, dep["logCenterSd"][2]]}, engine=eng,
                           columns=np.arange(len(dep["columns"])))
    assert eng.rand_index(dep["clusterVector"], alloc, 100) == 1.0
    assert model["logCenterSd"][0] is False


def test_dOptPenalty_expected_output(eng):
    """test_dOptPenalty.R:4-22 — incl. penalty 0: at the best penalty nClust == 2 and ARI == 1"""
    rng = np.random.default_rng(6)
    x = rapi.generateBimodalData(observations=200, dataCols=10, rng=rng)
    xs = x["samples"] / np.std(x["samples"], ddof=1)
    r = rapi.dOptPenalty(xs, k=30, maxIter=20, minARIImprovement=0.01, sampleSize=100, penalties=[0, 2, 4, 8, 16, 32, 128],
                         createOutput=False, optimARI=0.95, nCores=2, engine=eng, seed=2, disableWarnings=True)
    pos = r["bestPenaltyPosition"]
    assert r["meanOptimDf"]["nClust"][pos] == 2
    assert r["meanOptimDf"]["ARI"][pos] == 1
    assert r["workerIterations"] >= 20          # the loop never stops before 20 worker-iterations (:47)


def test_penalty_scaling_and_best_rule(eng):
    """reg = penalty * sampleSize * sqrt(D) / 1450 (R/depechePenaltyOpt.R:22-23) and the median-of-near-optimal rule (:184-192)"""
    calls = []

    class Spy(OracleEngine):
        def grid_search_each(self, X, k, reg, iterations, boot, seed_off_set=0):
            calls.append(np.array(reg))
            nreg = len(reg)
            ari = np.linspace(0.5, 0.99, nreg)[None, :]
            return {"z": [ari.copy() for _ in range(iterations)], "m": [np.full((1, nreg), 5.0) for _ in range(iterations)],
                    "c": [[[np.eye(3, 16), np.eye(3, 16)] * 2 for _ in range(nreg)] for _ in range(iterations)]}

    X = np.zeros((50, 16))
    r = rapi.depechePenaltyOpt(X, k=3, maxIter=20, minARIImprovement=0.01, sampleSize=50, penalties=[1, 2, 4, 8], optimARI=0.95, nCores=4,
                               engine=Spy(eng.o), disableWarnings=True)
    assert np.allclose(calls[0], np.array([1, 2, 4, 8]) * 50 * 4 / 1450)
    assert r["bestPenalty"] == 8.0                     # ARI rises with the penalty: only the last is within 0.05 of the best
    assert len(calls[-1]) <= len(calls[0])             # penalties are pruned, never added


def test_pretty_and_peak_centre():
    assert np.allclose(rapi._r_pretty(0.0, 1.0, 10), np.arange(0, 1.01, 0.1))
    assert np.allclose(rapi._r_pretty(-19.0, 161.5, 400), np.arange(-19.0, 162.0, 0.5))
    rng = np.random.default_rng(0)
    x = np.r_[rng.normal(3.0, 0.2, 5000), rng.normal(9.0, 2.0, 3000)]
    assert abs(rapi._peak_center(x) - 3.0) < 0.1
    Xs, info = rapi.depecheLogCenterSd(np.c_[x, x * 2], False, "default", True)
    assert info[0] is False and abs(np.std(Xs, ddof=1) - 1) < 1e-12 and abs(info[1][0] - 3.0) < 0.1


def test_testdata_scaling_matches_stored_fixture(testdata):
    """the stored depeche(testData[, 2:15]) result was produced without log2 (kurtosis 1.93)"""
    X = testdata["X"]
    flat = X.ravel()
    kurt = len(flat) * np.sum((flat - flat.mean()) ** 4) / np.sum((flat - flat.mean()) ** 2) ** 2
    assert abs(kurt - 1.93) < 0.05
    Xs, info = rapi.depecheLogCenterSd(X)
    assert info[0] is False and Xs.shape == (20000, 14)
    assert abs(10000 * np.sqrt(14) / 1450 - 25.80) < 0.01


def test_prescaling_pinned_by_stored_depeche_result(testdata):
    """f-3 pin: R itself is not in the image, so pretty() / hist() / the peak centres cannot be regenerated here — but the
    stored depeche(testData) result carries them implicitly.  The stored cluster vector is dAllocate() of the data scaled with R's own
    peak centres and sd onto the stored centres (which that package version keeps centred: scaled centre x sd).  Scaling with OUR
    restatement's centres and allocating with the reference's arithmetic must reproduce all 20 000 stored labels; moving any single
    marker's centre to a neighbouring histogram bin (0.5 units) changes at least one label (checked below), so equality pins the peak
    bin of every marker.  (R/depecheLogCenterSd.R:49-63, R/dScaleCoFunction.R:74-88, R/depeche.R:228-232.)"""
    from oracle.oracle import Oracle
    orc = Oracle()
    X, cc, stored = testdata["X"], testdata["stored_clusterCenters"], testdata["stored_clusterVector"]
    _, info = rapi.depecheLogCenterSd(X)
    c, sd = np.asarray(info[1], float), float(info[2])

    def labels(centres):
        return orc.allocate_points(np.ascontiguousarray((X - centres) / sd), cc / sd, True) + 1

    assert info[0] is False
    assert np.array_equal(labels(c), stored)
    for j in range(X.shape[1]):          # the pin has power for every marker, in both directions
        for shift in (0.5, -0.5):
            moved = c.copy()
            moved[j] += shift
            assert not np.array_equal(labels(moved), stored), (j, shift)


def test_renumber_by_size_matches_sort_based_definition():
    """depeche() renumbers the clusters by decreasing size (R/depeche.R:236-252); the counting implementation must equal the definition
    through np.unique + one comparison per cluster, including ties in size (old number decides) and unused cluster numbers"""
    from depecher_b200.rapi import _renumber_by_size
    rng = np.random.default_rng(8)
    for trial in range(20):
        k = int(rng.integers(2, 12))
        present = np.sort(rng.choice(np.arange(1, k + 1), size=int(rng.integers(1, k + 1)), replace=False))
        sizes = rng.integers(1, 6, size=len(present)) * 10          # few distinct sizes: ties are common
        cv = rng.permutation(np.repeat(present, sizes)).astype(np.int64)
        got, order, rank = _renumber_by_size(cv, k)
        labels, counts = np.unique(cv, return_counts=True)
        want_order = labels[np.argsort(-counts, kind="stable")]
        want = np.zeros_like(cv)
        for i, lab in enumerate(want_order, start=1):
            want[cv == lab] = i
        assert np.array_equal(got, want) and np.array_equal(order, want_order)
        assert rank == {int(lab): i for i, lab in enumerate(want_order, start=1)}
        assert got.dtype == cv.dtype and np.array_equal(np.bincount(got)[1:], np.sort(counts)[::-1])
