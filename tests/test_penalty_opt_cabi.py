"""The penalty state machine behind the C ABI (dpr_penalty_opt_*, depecher_b200/csrc/penalty_opt.cu) against the independent
Python restatement of R/depechePenaltyOpt.R:14-265 (tests/penalty_opt_py.py) on identical, randomly generated sets.  Host code
only: runs without a GPU."""
import numpy as np
import pytest

import depecher_b200 as dp
from penalty_opt_py import penalty_opt_py


def make_sets(rng, n_pen, k, d, nCores, flavour):
    """a deterministic source of sets: results depend on (penalty position, set number, worker) only, so the C and the Python
    state machine see the same numbers whatever order they ask in"""
    base = np.clip(np.sort(rng.uniform(0.2, 1.0, n_pen)) if flavour != "flat" else np.full(n_pen, 0.9), 0, 1)
    if flavour == "peak":
        base = 1.0 - 0.08 * np.abs(np.arange(n_pen) - n_pen // 2)
    noise = {"rising": 0.05, "peak": 0.02, "flat": 0.0, "noisy": 0.3}[flavour]
    seed = int(rng.integers(1 << 30))

    def run_set(used, regs, it):
        ari = np.empty((nCores, len(used)))
        ncl = np.empty((nCores, len(used)))
        cen = np.empty((nCores, len(used), 2, k, d))
        for w in range(nCores):
            for q, p in enumerate(used):
                r = np.random.default_rng([seed, int(p), it, w])
                ari[w, q] = base[p] + noise * r.normal()
                ncl[w, q] = float(r.integers(1, k + 1)) if r.random() < 0.9 else 1.0      # sometimes trivial (:75-78)
                if r.random() < 0.03 and p > 0:
                    ari[w, q] = np.nan                                                  # both partitions trivial: 0/0
                cen[w, q] = r.normal(size=(2, k, d))
        return ari, ncl, cen
    return run_set


@pytest.mark.parametrize("flavour", ["rising", "peak", "flat", "noisy"])
def test_state_machine_equals_python_restatement(flavour):
    rng = np.random.default_rng({"rising": 1, "peak": 2, "flat": 3, "noisy": 4}[flavour])
    for trial in range(40):
        n_pen = int(rng.integers(1, 12))
        k, d = int(rng.integers(2, 6)), int(rng.integers(1, 5))
        nCores = int(rng.integers(1, 11))
        maxIter = int(rng.choice([8, 20, 40, 100]))
        minImp = float(rng.choice([0.01, 0.001, 0.1]))
        optimARI = float(rng.choice([0.95, 0.99, 0.8]))
        sampleSize = int(rng.integers(50, 10001))
        penalties = 2.0 ** np.arange(0, n_pen * 0.5, 0.5)[:n_pen] if rng.random() < 0.7 else np.sort(rng.uniform(0, 40, n_pen))
        run_set = make_sets(rng, n_pen, k, d, nCores, flavour)
        want = penalty_opt_py(run_set, d, k, maxIter, minImp, sampleSize, penalties, optimARI, nCores)
        opt = dp.PenaltyOpt(penalties, k, d, sampleSize, nCores, maxIter, minImp, optimARI)
        trace, it = [], 1
        while True:
            used, regs, cont = opt.used()
            if not cont:
                break
            assert np.allclose(regs, penalties[used] * sampleSize * np.sqrt(d) / 1450.0, rtol=1e-15)      # :22-23
            cont_after = opt.feed(*run_set(used, regs, it))
            trace.append(opt.used()[0])
            assert cont_after == opt.used()[2]
            it += 1
        got = opt.result()
        opt.close()
        assert len(trace) == len(want["trace"]) and all(np.array_equal(a, b) for a, b in zip(trace, want["trace"])), (flavour, trial)
        assert got["workerIterations"] == want["workerIterations"] and got["warn"] == want["warn"]
        assert got["bestPenaltyPosition"] == want["bestPenaltyPosition"] and got["bestPenalty"] == want["bestPenalty"]
        np.testing.assert_allclose(got["meanARI"], want["meanARI"], rtol=1e-13, atol=1e-15, equal_nan=True)
        np.testing.assert_allclose(got["meanClust"], want["meanClust"], rtol=1e-13, equal_nan=True)
        assert len(got["bestClusterCenters"]) == len(want["bestClusterCenters"])
        for a, b in zip(got["bestClusterCenters"], want["bestClusterCenters"]):
            assert np.array_equal(a, b)


def test_loop_rule_and_best_penalty_rule():
    """at least 20 worker-iterations (:47); all-equal ARI 1 keeps every penalty and picks the median position, the even one of the two
    central positions when their number is even (R's round half to even, :192); exact mean of ones is 1 (test_dOptPenalty.R:19-22)"""
    for n_pen, want_pos in [(7, 3), (4, 1), (6, 3), (2, 1), (1, 0)]:    # seq 1..n: mean 4 -> 4th; 2.5 -> 2nd; 3.5 -> 4th; 1.5 -> 2nd
        opt = dp.PenaltyOpt(np.arange(1.0, n_pen + 1), 3, 2, 100, 3, 100, 0.01, 0.95)
        sets = 0
        while opt.used()[2]:
            nu = len(opt.used()[0])
            opt.feed(np.ones((3, nu)), np.full((3, nu), 2.0), np.zeros((3, nu, 2, 3, 2)))
            sets += 1
        r = opt.result()
        assert sets == 7 and r["workerIterations"] == 21          # 3 workers: 7 sets reach 20; sd 0 -> stdChange 0 stops there
        assert r["bestPenaltyPosition"] == want_pos
        assert (r["meanARI"] == 1.0).all() and (r["meanClust"] == 2.0).all()
        assert len(r["bestClusterCenters"]) == 2 * 21
        opt.close()


def test_errors():
    with pytest.raises(dp.DepecherError):
        dp.PenaltyOpt([1.0], 1, 2, 100, 2, 100, 0.01, 0.95)                 # k < 2
    opt = dp.PenaltyOpt([1.0, 2.0], 3, 2, 100, 2, 100, 0.01, 0.95)
    with pytest.raises(dp.DepecherError):
        opt.result()                                                         # nothing fed yet
    with pytest.raises(dp.DepecherError):
        opt.feed(np.ones((2, 3)), np.ones((2, 3)))                           # wrong number of penalties
    opt.close()
