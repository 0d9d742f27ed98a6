"""Independent Python restatement of depechePenaltyOpt's state machine (R/depechePenaltyOpt.R:14-265) — TEST INFRASTRUCTURE.

The product runs the state machine inside the library (depecher_b200/csrc/penalty_opt.cu behind dpr_penalty_opt_*); this file is the
round-1 line-by-line Python mirror of the R code, kept as the checker the C++ is compared with on identical sets
(tests/test_penalty_opt_cabi.py).  `run_set(used_positions, real_penalties_of_them, set_number)` supplies one set's results as
(ari[nCores][n_used], nclust[nCores][n_used], centres[nCores][n_used] -> [ret1, ret2])."""
import math
import warnings

import numpy as np


def _sd(v):
    v = np.asarray(v, float)
    return float(np.std(v, ddof=1)) if len(v) > 1 else float("nan")


def penalty_opt_py(run_set, d, k, maxIter, minARIImprovement, sampleSize, penalties, optimARI=0.95, nCores=1, disableWarnings=True):
    penalties = np.asarray(penalties, float)
    penaltyConstant = sampleSize * math.sqrt(d) / 1450.0                                       # :22
    realPenalties = penalties * penaltyConstant
    roundPenalties = np.round(penalties, 1)
    it, lastStd, stdChange, iterTimesNCores = 1, 1.0, 1.0, 1
    usedPos = np.arange(len(penalties))
    full = []          # one entry per worker-iteration: {penalty position -> (ARI, nClust, [ret1 centres, ret2 centres])}
    trace = []
    while iterTimesNCores < 20 or (iterTimesNCores < maxIter and stdChange >= minARIImprovement):     # :47-48
        ari_set, ncl_set, cen_set = run_set(usedPos, realPenalties[usedPos], it)
        for w in range(nCores):
            ari = np.array(ari_set[w], float).ravel()
            ncl = np.array(ncl_set[w], float).ravel()
            ari = np.where(ncl == 1, 0.0, ari)                                                 # :75-78
            full.append({int(p): (ari[q], ncl[q], cen_set[w][q][:2]) for q, p in enumerate(usedPos)})
        ariVecs = [[e[p][0] for e in full if p in e] for p in range(len(penalties))]           # :89-97 (keyed by position)
        meanARI = np.array([np.mean(v) if v else np.nan for v in ariVecs])
        stdOpt = np.array([_sd(v) if v else np.nan for v in ariVecs])
        maxPos = int(np.nanargmax(meanARI))                                                    # :103
        meanMinus2StdMax = meanARI[maxPos] - 2 * stdOpt[maxPos]
        meanPlus2StdAll = meanARI + 2 * stdOpt
        localIter = min(it, 3)
        k1, k2 = (2, 1, 0)[localIter - 1], (4, 3, 2)[localIter - 1]                            # :123-125
        with np.errstate(invalid="ignore"):
            keep = (meanPlus2StdAll >= meanMinus2StdMax - k1 * stdOpt[maxPos]) | (meanARI >= np.nanmax(meanARI) - (1 - optimARI) * k2)
        usedPos = np.array([p for p in usedPos if keep[p]], dtype=int)                         # :126-140
        if len(usedPos) > 1:                                                                   # :145-154
            stdChange = ((lastStd - stdOpt[maxPos]) / lastStd) / math.sqrt(it * nCores) if lastStd != 0 else 0.0
            if np.isnan(stdChange):
                stdChange = 0.0
        else:
            stdChange = 0.0
        lastStd = stdOpt[maxPos]
        iterTimesNCores = it * nCores
        trace.append(usedPos.copy())
        it += 1
    warn = 0
    if it * nCores >= maxIter and stdChange > minARIImprovement:                                # :173-177
        warn |= 1
    with np.errstate(invalid="ignore"):
        optimal = np.flatnonzero(meanARI >= np.nanmax(meanARI) - (1 - optimARI))               # :184-185
    bestPos = int(optimal[int(round(np.mean(np.arange(1, len(optimal) + 1)))) - 1])            # :192 (R rounds half to even, like Python)
    bestPenalty = float(roundPenalties[bestPos])
    if len(penalties) > 1:                                                                     # :196-207
        if bestPenalty == roundPenalties[0]:
            warn |= 2
        if bestPenalty == roundPenalties[-1]:
            warn |= 4
    meanClust = np.array([np.nanmean([e[p][1] for e in full if p in e]) if any(p in e for e in full) else np.nan
                          for p in range(len(penalties))])                                     # :242-249
    bestCenters = [c for e in full if bestPos in e for c in e[bestPos][2]]                     # :255-258
    return {"bestPenalty": bestPenalty, "bestPenaltyPosition": bestPos, "meanARI": meanARI, "meanClust": meanClust,
            "bestClusterCenters": bestCenters, "workerIterations": (it - 1) * nCores, "warn": warn, "trace": trace}
