"""Host-side mirror of the reference's R callers of the hot path (names, arguments and results as in R).

The reference drives its C++ engine from R (`depeche()` -> `depechePenaltyOpt()` -> `grid_search`,
`depecheAllData()` -> `sparse_k_means`, `depecheClusterCenterSelection()` / `dAllocate()` ->
`allocate_points`, `rand_index`).  R is not part of this image, so the same control flow is restated here
in Python over the C ABI; every function cites the R lines it follows.  The state machine is unchanged;
what changes is the fan-out: the reference starts `nCores` SOCK worker processes per set and ships the data
matrix to each (R/depechePenaltyOpt.R:37-38,50-54) — here one call runs all workers' fits as one batch on
the GPU from a matrix that stays resident in HBM.

`engine` is anything with the four entry points of the reference (`grid_search`, `sparse_k_means`,
`allocate_points`, `rand_index`, results keyed like the Rcpp lists).  The default is the CUDA engine
(`depecher_b200.Engine`), which fails loudly without a B200; tests inject a CPU oracle engine to check the
host logic on machines without a GPU.
"""
from __future__ import annotations

import math
import os
import warnings

import numpy as np

__all__ = ["depeche", "dAllocate", "depechePenaltyOpt", "dOptPenalty", "depecheAllData", "depecheClusterCenterSelection",
           "depecheLogCenterSd", "depecheGroundCenters", "generateBimodalData", "generateSparseData", "default_ncores"]


# --------------------------------------------------------------------------------------------------------
# fixtures: R/simulateData.R
# --------------------------------------------------------------------------------------------------------
def generateBimodalData(centers=None, prop=0.3, dataCols=5, observations=10000, rng=None):
    """R/simulateData.R:8-36: two unit-variance blobs at runif+50 / runif-50, `prop` of the rows in the first."""
    rng = rng or np.random.default_rng()
    if centers is None:
        centers = np.vstack([rng.random(dataCols) + 50, rng.random(dataCols) - 50])
    centers = np.asarray(centers, float)
    dataCols = centers.shape[1]
    n1 = int(math.floor(prop * observations))
    ids = np.r_[np.full(n1, 1), np.full(observations - n1, 2)]
    rands = rng.normal(size=(observations, dataCols))
    samples = np.vstack([rands[:n1] + centers[0], rands[n1:] + centers[1]])
    return {"samples": samples, "ids": ids}


def generateSparseData(modeN=5, dataCols=100, observations=10000, rng=None):
    """R/simulateData.R:47-80: modeN blobs at +-50 per coordinate, blob i has i coordinates set to 0."""
    rng = rng or np.random.default_rng()
    if observations % modeN:
        raise ValueError("Observations has to be divisible by modeN")
    per = observations // modeN
    centers = rng.choice([-50.0, 50.0], size=(modeN, dataCols))
    for i in range(modeN):
        centers[i, rng.choice(dataCols, size=i + 1, replace=False)] = 0.0
    samples = np.zeros((observations, dataCols))
    ids = np.zeros(observations, int)
    for i in range(modeN):
        samples[i * per:(i + 1) * per] = rng.normal(size=(per, dataCols)) + centers[i]
        ids[i * per:(i + 1) * per] = i + 1
    return {"samples": samples, "ids": ids, "centers": centers}


# --------------------------------------------------------------------------------------------------------
# pre-scaling: R/depecheLogCenterSd.R, R/dScaleCoFunction.R:74-99
# --------------------------------------------------------------------------------------------------------
def _r_pretty(lo: float, hi: float, n: int, min_n: int = 1):
    """R's pretty.default(c(lo, hi), n, min.n) as hist.default calls it: R_pretty (src/appl/pretty.c) for the 'nice' unit
    (1, 2, 5 x 10^k) and bounds, then seq.int(l, u, length.out = k + 1) for the breakpoints."""
    h, h5, eps = 1.5, 0.5 + 1.5 * 1.5, 1e-10
    dx = hi - lo
    if dx == 0 and hi == 0:
        cell, i_small = 1.0, True
    else:
        cell = max(abs(lo), abs(hi))
        U = 1 + (1 / (1 + h) if h5 >= 1.5 * h + 0.5 else 1.5 / (1 + h5))
        U *= max(1, n) * np.finfo(float).eps
        i_small = dx < cell * U * 3
    if i_small:
        if cell > 10:
            cell = 9 + cell / 10
        cell *= 0.75          # shrink.sml
        if min_n > 1:
            cell /= min_n
    else:
        cell = dx
        if n > 1:
            cell /= n
    cell = max(cell, 20 * np.finfo(float).tiny)
    base = 10.0 ** math.floor(math.log10(cell))
    unit = base
    if 2 * base - cell < h * (cell - unit):
        unit = 2 * base
        if 5 * base - cell < h5 * (cell - unit):
            unit = 5 * base
            if 10 * base - cell < h * (cell - unit):
                unit = 10 * base
    ns, nu = math.floor(lo / unit + eps), math.ceil(hi / unit - eps)
    while ns * unit > lo + eps * unit:
        ns -= 1
    while nu * unit < hi - eps * unit:
        nu += 1
    k = int(0.5 + nu - ns)
    if k < min_n:
        k = min_n - k
        if ns >= 0:
            nu += k // 2
            ns -= k // 2 + k % 2
        else:
            ns -= k // 2
            nu += k // 2 + k % 2
        k = min_n
    l, u = ns * unit, nu * unit
    by = (u - l) / k if k else 0.0
    b = l + np.arange(k + 1) * by
    b[0], b[-1] = l, u
    b[np.abs(b) < 1e-14 * by] = 0.0
    return b


def _peak_center(x: np.ndarray) -> float:
    """dScaleCoFunction.R:74-88: mid of the fullest bin of hist(x, breaks = n/50 (10 below 500 rows)); hist.default's
    right-closed bins with its 1e-7 * median(diff(breaks)) fuzz, lowest value included."""
    n_breaks = 10 if len(x) < 500 else int(len(x) / 50)
    br = _r_pretty(float(x.min()), float(x.max()), n_breaks, 1)
    h = np.diff(br)
    nB = len(br)
    diddle = 1e-7 * (np.median(h) if nB > 5 else (np.ptp(x) if nB <= 3 else max(np.ptp(x), np.median(h))))
    fz = br + diddle
    fz[0] = br[0] - diddle
    idx = np.searchsorted(fz, x, side="left") - 1        # fz[b] < x <= fz[b+1]
    idx[(x >= fz[0]) & (idx < 0)] = 0                     # include.lowest
    ok = (idx >= 0) & (idx < nB - 1)
    counts = np.bincount(idx[ok], minlength=nB - 1)
    b = int(np.argmax(counts))          # match(max(counts), counts): first maximum
    return float((br[b] + br[b + 1]) / 2)


def depecheLogCenterSd(X, log2Off=False, center="default", scale=True):
    """R/depecheLogCenterSd.R:6-115 -> (scaled matrix, [log2 applied, centres | False, sd])."""
    X = np.array(X, dtype=np.float64)
    info = [False, False, 1]
    flat = X.ravel()
    kurt = len(flat) * np.sum((flat - flat.mean()) ** 4) / np.sum((flat - flat.mean()) ** 2) ** 2   # moments::kurtosis
    if not log2Off and kurt > 100:                                                            # :11-45
        if X.min() >= 0:
            X = np.log2(X + 1)
        else:
            X = np.log2(X - X.min(axis=0) + 1)
            X[np.isnan(X)] = 0
        info[0] = True
    d = X.shape[1]
    if isinstance(center, str):
        if center == "peak" or (center == "default" and d <= 100):                             # :49-63
            c = np.array([_peak_center(X[:, j]) for j in range(d)])
        elif center == "mean" or (center == "default" and d > 100):                            # :64-78
            c = X.mean(axis=0)
        else:
            raise ValueError("center must be 'default', 'peak', 'mean', False or a vector")
        X = X - c
        info[1] = c
    elif center is False:                                                                      # :95-98
        pass
    else:                                                                                      # :79-94
        c = np.asarray(center, float)
        if len(c) != d:
            raise ValueError("Mismatch between the number of columns and the length of the centering vector")
        X = X - c
        info[1] = c
    if scale is True:                                                                          # :102-106
        sd = float(np.std(X, ddof=1))
        X = X / sd
        info[2] = sd
    elif scale is not False:
        X = X / float(scale)
        info[2] = float(scale)
    return X, info


def depecheGroundCenters(reduced, logCenterSd):
    """R/depecheGroundCenters.R:7-28: un-scale the centres (x sd, + centre); exact zeros stay zero."""
    reduced = np.asarray(reduced, float)
    out = reduced * logCenterSd[2]
    if logCenterSd[1] is not False:
        out = out + np.asarray(logCenterSd[1], float)[None, :]
        out[reduced == 0] = 0
    return out


def default_ncores() -> int:
    """R/depeche.R:169-174"""
    return max(1, min(10, int(math.floor((os.cpu_count() or 1) * 0.875))))


def _engine(engine):
    if engine is not None:
        return engine
    from . import Engine
    return Engine()


# --------------------------------------------------------------------------------------------------------
# dAllocate: R/dAllocate.R:62-137
# --------------------------------------------------------------------------------------------------------
def _reduce_centres(cc):
    """rows with rowSums != 0, columns with colSums != 0 (R/dAllocate.R:99-103)"""
    rows = np.flatnonzero(cc.sum(axis=1) != 0)
    cols = np.flatnonzero(cc.sum(axis=0) != 0)
    return rows, cols


def dAllocate(inDataFrame, depModel, engine=None, columns=None):
    """Allocate cells to a fitted model; returns 1-based labels into the reduced centre matrix.

    `columns`: when the centre matrix carries column names in R (model from depeche()), the data keeps
    exactly those columns (:112-114); pass their indices.  Otherwise zero-sum columns are dropped (:115-118)."""
    eng = _engine(engine)
    cc = np.asarray(depModel["clusterCenters"], float)
    lcs = depModel.get("logCenterSd", False)
    X = inDataFrame
    if isinstance(lcs, (list, tuple)):
        if lcs[0]:
            raise ValueError("dAllocate does not work well with data that has been internally log-transformed")   # :71-79
        X, _ = depecheLogCenterSd(np.asarray(X, float), True, lcs[1] if lcs[1] is not False else False, lcs[2])  # :81-85
        ccs, _ = depecheLogCenterSd(cc, True, lcs[1] if lcs[1] is not False else False, lcs[2])                  # :86-90
        ccs[cc == 0] = 0
        cc = ccs
    rows, cols = _reduce_centres(cc)
    red = cc[np.ix_(rows, cols)]
    if columns is not None:
        Xr = np.asarray(X)[:, columns] if not hasattr(X, "handle") else X
    elif hasattr(X, "handle"):
        # resident dataset: its columns cannot be dropped, so the centre columns R would drop (colSums == 0, :99-103 — also a column
        # that only sums to zero by cancellation) are zeroed instead; they then add the same x_j^2 to every distance.  Labels equal R's
        # except at fp64 near-ties of the distances (the extra term can change the last bit of a sum).
        Xr = X
        red = cc[rows].copy()
        drop = np.ones(cc.shape[1], bool)
        drop[cols] = False
        red[:, drop] = 0.0
    else:
        Xr = np.asarray(X)[:, cols]
    if red.ndim == 1:
        red = red[None, :]
    if hasattr(eng, "lib"):                                                      # the CUDA engine adds the 1 (:127-134) while it copies the labels out
        return eng.allocate_points(Xr, red, 1, base=1)["i"]
    lab = eng.allocate_points(Xr, red, 1)["i"]                                   # R's integers are 32-bit too: no widening copy of a vector
    lab += 1                                                                     # :127-134 (in place)                  with one entry per cell
    return lab


# --------------------------------------------------------------------------------------------------------
# depechePenaltyOpt (and its legacy twin dOptPenalty): R/depechePenaltyOpt.R:14-265
# --------------------------------------------------------------------------------------------------------
def _grid_search_set(eng, X, k, regs, nCores, sampleSize, seed, dist=None):
    """One set of the search = nCores workers x grid_search(X, k, regs, 1, sampleSize, .) (R/depechePenaltyOpt.R:50-55).
    Single process: one batched call.  With a torch.distributed group (`dist`) the workers are dealt round-robin over the
    ranks (each rank holds X), every rank fits its share, and the small per-worker results are all-gathered, so every rank
    runs the identical state machine afterwards.  No data-path collective: the fits are independent (Clusterer.cpp:349-384)."""
    if dist is None or dist.get_world_size() == 1:
        return eng.grid_search_each(X, [k], regs, nCores, sampleSize, seed)
    from . import shard
    rank, world = dist.get_rank(), dist.get_world_size()
    mine = shard.grid_split(nCores, rank, world)
    part = None
    if len(mine):
        part = eng.grid_search_each(X, [k], regs, len(mine), sampleSize, seed * 1000 + rank)
    gathered = [None] * world
    dist.all_gather_object(gathered, (mine.tolist(), None if part is None else (part["z"], part["m"], part["c"])))
    z, m, c = [None] * nCores, [None] * nCores, [None] * nCores
    for workers, payload in gathered:
        for q, w in enumerate(workers):
            z[w], m[w], c[w] = payload[0][q], payload[1][q], payload[2][q]
    return {"d": z, "z": z, "n": m, "m": m, "c": c}


def depechePenaltyOpt(inDataFrameUsed, k, maxIter, minARIImprovement, sampleSize, penalties, createOutput=False, disableWarnings=False,
                      optimARI=0.95, nCores=None, plotDir=".", engine=None, seed=0, verbose=False, dist=None):
    """R/depechePenaltyOpt.R:14-265.  The state machine (loop rule :47-48, trivial-solution zeroing :75-78, mean / sd per penalty
    :89-100, pruning with the k1 / k2 schedule :123-140, stdChange :145-156, best penalty :184-192, warnings :173-207, outputs
    :242-264) runs inside the library behind the C ABI (dpr_penalty_opt_*, csrc/penalty_opt.cu).  With the CUDA engine and a
    resident matrix the whole search is ONE call (dpr_dataset_penalty_opt: every set one batch of problems on the device; with a
    torch.distributed group the problems of a set are dealt over the ranks).  Any other engine (the CPU oracle of the tests)
    supplies the sets through `grid_search_each` and feeds them to the same state machine."""
    from ._capi import PenaltyOpt
    eng = _engine(engine)
    nCores = nCores or default_ncores()
    X = inDataFrameUsed
    d = X.d if hasattr(X, "handle") else np.asarray(X).shape[1]
    penalties = np.asarray(penalties, float)
    opt = PenaltyOpt(penalties, k, d, sampleSize, nCores, maxIter, minARIImprovement, optimARI, lib=getattr(eng, "lib", None))
    try:
        if hasattr(eng, "penalty_opt"):
            split = dist is not None and dist.get_world_size() > 1
            ds, own = eng._as_ds(X)
            try:
                eng.penalty_opt(ds, opt, seed, split_over_ranks=split)
            finally:
                if own:
                    ds.close()
        else:
            it = 1
            while True:
                usedPos, regs, cont = opt.used()
                if not cont:
                    break
                # nCores workers x grid_search(dataMat, k, usedPenalties, 1, sampleSize, i)  (:50-55) as ONE batched call
                res = _grid_search_set(eng, X, k, regs, nCores, sampleSize, seed + 7919 * it, dist)
                ari = np.array([np.array(res["z"][w]).ravel() for w in range(nCores)])
                ncl = np.array([np.array(res["m"][w]).ravel() for w in range(nCores)])
                cen = np.array([[[np.asarray(m, float) for m in res["c"][w][q][:2]] for q in range(len(usedPos))] for w in range(nCores)])
                opt.feed(ari, ncl, cen)
                if verbose:
                    print(f"Set {it} with {nCores} iterations completed; {len(opt.used()[0])} penalties kept")
                it += 1
        r = opt.result()
    finally:
        opt.close()
    if not disableWarnings:
        if r["warn"] & 1:                                                                        # :173-177
            warnings.warn("The maximum number of iterations was reached before stable optimal solution was found")
        if r["warn"] & 2:                                                                        # :196-207
            warnings.warn("The lowest penalty was the most optimal in the range.")
        if r["warn"] & 4:
            warnings.warn("The highest penalty was the most optimal in the range.")
    return {"bestPenalty": r["bestPenalty"], "bestPenaltyPosition": r["bestPenaltyPosition"],
            "meanOptimDf": {"penalty": np.round(penalties, 1), "ARI": r["meanARI"], "nClust": r["meanClust"]},
            "bestClusterCenters": r["bestClusterCenters"], "workerIterations": r["workerIterations"]}


dOptPenalty = depechePenaltyOpt       # R/dOptPenalty.R is the same function under its old name


# --------------------------------------------------------------------------------------------------------
# depecheAllData: R/depecheAllData.R:14-79   (fewer than 10 000 rows)
# --------------------------------------------------------------------------------------------------------
def depecheAllData(inDataFrameUsed, penalty, k, nCores=None, engine=None, seed=0):
    eng = _engine(engine)
    X = inDataFrameUsed
    n, d = (X.n, X.d) if hasattr(X, "handle") else np.asarray(X).shape
    reg = penalty * (n * math.sqrt(d)) / 1450.0                                               # :15-16
    runs = [eng.sparse_k_means(X, int(round(k * 3)), reg, 1, seed + i) for i in range(1, 22)]  # :30-31
    logMaxLik = np.array([r["n"] for r in runs])
    nClust = np.array([int(((r["c"] != 0).sum(axis=1) != 0).sum()) for r in runs])            # :37-40
    ok = logMaxLik[nClust > 1]
    maxN = ok.max()
    best = runs[int(np.flatnonzero(logMaxLik == maxN)[0])]                                     # :42-44
    cc = best["c"]
    rows, cols = _reduce_centres(cc)                                                           # :56-59
    return {"clusterVector": best["i"].astype(np.int64), "clusterCenters": cc[np.ix_(rows, cols)], "rows": rows, "cols": cols}


# --------------------------------------------------------------------------------------------------------
# depecheClusterCenterSelection: R/depecheClusterCenterSelection.R:7-59   (10 000 rows or more)
# --------------------------------------------------------------------------------------------------------
def depecheClusterCenterSelection(allSolutions, selectionDataSet, k, nCores=None, engine=None):
    eng = _engine(engine)
    if hasattr(eng, "select_solution"):
        # the whole selection behind the C ABI (dpr_dataset_select_solution): every solution's dAllocate (:11-16) and all
        # rand_index pairs (:24-29) in one pass over the selection set
        pick, meanARI = eng.select_solution(selectionDataSet, allSolutions)
        best = np.asarray(allSolutions[pick], float)
        rows, cols = _reduce_centres(best)                                                         # :41-45
        return {"clusterCenters": best[np.ix_(rows, cols)], "rows": rows, "cols": cols, "meanARI": meanARI}
    if hasattr(eng, "allocate_many"):
        # every solution's dAllocate (:11-16) and all rand_index pairs (:24-29) in one pass over the selection set.
        # dAllocate drops zero-sum rows; the labels it returns index the kept rows, 1-based.
        masked = []
        for s in allSolutions:
            m = np.array(s, float)
            m[m.sum(axis=1) == 0] = 0.0
            masked.append(m)
        _, ari = eng.allocate_many(selectionDataSet, masked, True, want_labels=False, want_ari=True)
        meanARI = ari.mean(axis=0)
    else:
        alloc = [dAllocate(selectionDataSet, {"clusterCenters": s, "logCenterSd": False}, engine=eng) for s in allSolutions]
        meanARI = np.array([np.mean([eng.rand_index(a, alloc[i], k) for a in alloc]) for i in range(len(alloc))])
    best = np.asarray(allSolutions[int(np.flatnonzero(meanARI == np.nanmax(meanARI))[0])], float)  # :35-37
    rows, cols = _reduce_centres(best)                                                         # :41-45
    return {"clusterCenters": best[np.ix_(rows, cols)], "rows": rows, "cols": cols, "meanARI": meanARI}


def _renumber_by_size(clusterVector, n_centres):
    """clusters renumbered 1, 2, ... by decreasing size, ties in the order of the old numbers (R/depeche.R:236-252: `sort(table(.), decreasing = TRUE)`
    and a loop over its names).  Labels are small positive integers, so one counting pass and a lookup table do it - no sort of the vector, no
    full-length comparison per cluster.  Returns (new vector, old labels in rank order, {old label: new number})."""
    cv = np.asarray(clusterVector)
    hist = np.bincount(cv, minlength=n_centres + 1)
    labels = np.flatnonzero(hist)
    order = labels[np.argsort(-hist[labels], kind="stable")]
    lut = np.zeros(len(hist), dtype=cv.dtype)
    rank_of_label = {}
    for i, lab in enumerate(order, start=1):
        lut[lab] = i
        rank_of_label[lab] = i
    return lut[cv], order, rank_of_label


# --------------------------------------------------------------------------------------------------------
# depeche: R/depeche.R:127-294
# --------------------------------------------------------------------------------------------------------
def depeche(inDataFrame, samplingSubset=None, penalties=None, sampleSize="default", selectionSampleSize="default", k=30,
            minARIImprovement=0.01, optimARI=0.95, maxIter=100, log2Off=False, center="default", scale=True, nCores="default",
            plotDir=".", createOutput=False, engine=None, seed=0, verbose=False):
    eng = _engine(engine)
    rng = np.random.default_rng(seed)
    X = np.asarray(inDataFrame, float)
    if penalties is None:
        penalties = 2.0 ** np.arange(0, 5.5, 0.5)                                              # :128
    device_scaled = hasattr(eng, "dataset_scaled") and samplingSubset is None
    if device_scaled:   # the whole pre-scaling on the device; the scaled matrix never exists on the host (SURVEY f-3)
        ds_scaled, logCenterSd = eng.dataset_scaled(X, log2Off, center, scale)                  # :141
        Xs = None
        n_rows = X.shape[0]
    else:
        Xs, logCenterSd = depecheLogCenterSd(X, log2Off, center, scale)                         # :141
        n_rows = len(Xs)
    n_subset = n_rows if samplingSubset is None else len(samplingSubset)                       # (no index vector for "all rows": 80 MB at 10M rows)
    if n_subset > 1e6:                                                                          # :150-156
        warnings.warn("This dataset is very large; the reference recommends fitting on a sampling subset and allocating the rest with dAllocate")
    used = None if device_scaled else (Xs if samplingSubset is None else Xs[samplingSubset])
    n_used = n_rows if device_scaled else len(used)
    if sampleSize == "default":
        sampleSize = n_used if n_used <= 10000 else 10000                                       # :161-167
    if nCores == "default":
        nCores = default_ncores()                                                               # :169-174
    use_ds = hasattr(eng, "dataset")
    ranked = None
    out_labels, toucher = None, None
    if device_scaled and n_rows >= 10000 and hasattr(eng, "allocate_points_ranked"):
        # the cluster vector of the result: allocated now and written once by a helper thread while the device searches (every page of a fresh
        # vector faults on its first write: 40 MB at 10M cells), so that the final allocation copies its labels into mapped memory
        import threading
        out_labels = np.empty(n_rows, np.int32)
        toucher = threading.Thread(target=out_labels.fill, args=(0,))
        toucher.start()
    ds_used = ds_scaled if device_scaled else (eng.dataset(used) if use_ds else used)           # resident for the whole search (SURVEY f-2)
    try:
        opt = depechePenaltyOpt(ds_used, k=k, maxIter=maxIter, sampleSize=sampleSize, penalties=penalties, createOutput=createOutput,
                                minARIImprovement=minARIImprovement, optimARI=optimARI, nCores=nCores, plotDir=plotDir, engine=eng,
                                seed=seed, verbose=verbose)                                    # :176-184
        if selectionSampleSize == "default":
            selectionSampleSize = sampleSize
        if n_used < 10000:                                                                      # :207-215
            r = depecheAllData(ds_used, penalty=opt["bestPenalty"], k=k, nCores=nCores, engine=eng, seed=seed + 104729)
            clusterVector, reduced, cols = r["clusterVector"], r["clusterCenters"], r["cols"]
        else:
            if n_used <= selectionSampleSize:                                                   # :191-199
                sel = ds_used if device_scaled else used
            elif device_scaled:
                sel = eng.gather(ds_used, rng.integers(0, n_used, selectionSampleSize))
            else:
                sel = used[rng.integers(0, n_used, selectionSampleSize)]
            r = depecheClusterCenterSelection(opt["bestClusterCenters"], sel, k, nCores, engine=eng)   # :222-225
            reduced, cols = r["clusterCenters"], r["cols"]
            if device_scaled and sel is not ds_used:
                sel.close()
            # the final allocation is of the FULL matrix in its original row order (dAllocate(inDataFrameScaled, ...), R/depeche.R:228-232):
            # the resident matrix can serve only when it IS that matrix — not a samplingSubset of the same length (a permutation or a
            # resample), whose rows are in another order
            ds_all = ds_used if (use_ds and samplingSubset is None) else (eng.dataset(Xs[:, cols]) if use_ds else Xs[:, cols])
            full_cc = np.zeros((reduced.shape[0], X.shape[1]))
            full_cc[:, cols] = reduced
            model_cc = full_cc if ds_all is ds_used else reduced
            if hasattr(eng, "allocate_points_ranked") and hasattr(ds_all, "handle"):
                # allocation (:228-232) and renumbering by size (:236-252) in one call: sizes counted on the device, the new numbers
                # applied while the labels are copied out (no pass over a host vector with one entry per cell)
                rows_kept, cols_kept = _reduce_centres(model_cc)
                cc = model_cc[rows_kept].copy()
                drop = np.ones(cc.shape[1], bool)
                drop[cols_kept] = False
                cc[:, drop] = 0.0                                                               # (as dAllocate does for a resident matrix)
                if toucher is not None:
                    toucher.join()
                    toucher = None
                newVec, order0 = eng.allocate_points_ranked(ds_all, cc, 1, out=out_labels if ds_all is ds_used else None)
                ranked = (newVec, order0 + 1, {int(l) + 1: i for i, l in enumerate(order0, start=1)})
            else:
                clusterVector = dAllocate(ds_all, {"clusterCenters": model_cc, "logCenterSd": False}, engine=eng)   # :228-232
            if use_ds and ds_all is not ds_used:
                ds_all.close()
    finally:
        if toucher is not None:
            toucher.join()
        if use_ds:
            ds_used.close()
    # clusters renumbered by size (:236-252)
    newVec, order, rank_of_label = ranked if ranked is not None else _renumber_by_size(clusterVector, int(reduced.shape[0]))
    sortedLabels = np.sort(order)
    if len(sortedLabels) == reduced.shape[0]:
        reduced = reduced[np.argsort([rank_of_label[l] for l in sortedLabels])]
    ground = depecheGroundCenters(reduced, [logCenterSd[0], (np.asarray(logCenterSd[1])[cols] if logCenterSd[1] is not False else False),
                                            logCenterSd[2]])                                   # :262-263
    essence = [list(np.asarray(cols)[np.flatnonzero(reduced[i] != 0)]) for i in range(reduced.shape[0])]   # :266-271
    return {"clusterVector": newVec, "clusterCenters": ground, "scaledClusterCenters": reduced, "columns": np.asarray(cols),
            "essenceElementList": essence, "penaltyOptList": {"bestPenalty": opt["bestPenalty"], "allPenalties": opt["meanOptimDf"],
                                                              "workerIterations": opt["workerIterations"]},
            "logCenterSd": logCenterSd}
