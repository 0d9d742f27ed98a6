"""CPU engine with the reference's four entry points, backed by the oracle — TEST INFRASTRUCTURE.

It lets the host-side mirror of the R callers (depecher_b200/rapi.py) run on a machine without a GPU, so the
penalty state machine and the callers' logic are covered by the CPU test suite.  The product never uses it."""
import numpy as np

from oracle.oracle import RNG_PHILOX, Oracle


class OracleEngine:
    def __init__(self, oracle: Oracle | None = None):
        self.o = oracle or Oracle()
        self.o.set_dead_work(False)

    def allocate_points(self, X, mu, no_zero):
        return {"i": self.o.allocate_points(X, mu, bool(no_zero))}

    def sparse_k_means(self, X, k, reg, no_zero, seed_off_set=0, want_labels=True):
        self.o.set_rng(RNG_PHILOX, seed_off_set)
        r = self.o.sparse_k_means(X, k, reg, bool(no_zero), fit_id=0)
        return {"i": r["i"], "o": r["i"], "c": r["c"], "v": r["c"], "n": r["n"], "m": r["n"], "n_iter": r["n_iter"]}

    def grid_search(self, X, k, reg, iterations, bootstrapSamples, seed_off_set=0):
        self.o.set_rng(RNG_PHILOX, seed_off_set)
        g = self.o.grid_search(X, k, reg, iterations, bootstrapSamples)
        return {"d": g["d"], "z": g["d"], "n": g["n"], "m": g["n"], "c": [p + p for p in g["c"]]}

    def grid_search_each(self, X, k, reg, iterations, bootstrapSamples, seed_off_set=0):
        z, m, c = [], [], []
        for w in range(int(iterations)):
            g = self.grid_search(X, k, reg, 1, bootstrapSamples, seed_off_set * 1000003 + w)
            z.append(g["z"])
            m.append(g["m"])
            c.append(g["c"])
        return {"d": z, "z": z, "n": m, "m": m, "c": c}

    def rand_index(self, a, b, k):
        return self.o.rand_index_exact(np.asarray(a), np.asarray(b), k)
