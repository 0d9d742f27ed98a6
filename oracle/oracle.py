"""ctypes front-end of the CPU parity oracle — TEST INFRASTRUCTURE.

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / reference arm import this
module; ``depecher_b200`` never does.  Two checkers are exposed:

* ``Oracle``  — ``liboracle.so``: the plain-C++ restatement of /root/reference/src/Clusterer.cpp
  (oracle/clusterer_oracle.cpp).
* ``Ref``     — ``_ref/libdepecher_ref.so``: the reference's own Clusterer.cpp compiled against the
  header stand-ins in oracle/eigen_shim (only when it was built; see oracle/Makefile).

All matrices are numpy float64, shape (n, d); they cross the C interface column-major like R's.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
RNG_LIBC, RNG_PHILOX = 0, 1

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)


def build(quiet: bool = True) -> None:
    """make -C oracle (compiles the restatement, and _ref when /root/reference is present)."""
    subprocess.run(["make", "-C", _HERE] + (["-s"] if quiet else []), check=True)


def _f(a) -> np.ndarray:
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


def _p(a, t):
    return a.ctypes.data_as(t)


class Oracle:
    def __init__(self, path: str | None = None):
        path = path or os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        self.lib = C.CDLL(path)
        L = self.lib
        L.oracle_set_rng.argtypes = [C.c_int, C.c_uint64]
        L.oracle_set_rng.restype = None
        L.oracle_srand.argtypes = [C.c_uint]
        L.oracle_srand.restype = None
        L.oracle_set_dead_work.argtypes = [C.c_int]
        L.oracle_set_dead_work.restype = None
        self.set_rng(RNG_PHILOX, 0)

    # -- randomness ------------------------------------------------------------------------
    def set_rng(self, mode: int, seed: int = 0) -> None:
        self.lib.oracle_set_rng(mode, C.c_uint64(seed))

    def srand(self, seed32: int) -> None:
        self.lib.oracle_srand(C.c_uint(seed32 & 0xFFFFFFFF))

    def set_dead_work(self, on: bool) -> None:
        self.lib.oracle_set_dead_work(int(on))

    # -- the four exports --------------------------------------------------------------------
    def allocate_points(self, X, mu, no_zero: bool) -> np.ndarray:
        X, mu = _f(X), _f(mu)
        n, d = X.shape
        out = np.empty(n, np.int32)
        rc = self.lib.oracle_allocate_points(_p(X, _dp), C.c_int64(n), d, _p(mu, _dp), mu.shape[0], int(no_zero), _p(out, _ip))
        assert rc == 0
        return out

    def assignment_margin(self, X, mu, no_zero: bool) -> np.ndarray:
        X, mu = _f(X), _f(mu)
        n, d = X.shape
        out = np.empty(n, np.float64)
        self.lib.oracle_assignment_margin(_p(X, _dp), C.c_int64(n), d, _p(mu, _dp), mu.shape[0], int(no_zero), _p(out, _dp))
        return out

    def sparse_k_means(self, X, k: int, reg: float, no_zero: bool, fit_id: int = 0):
        X = _f(X)
        n, d = X.shape
        idx = np.empty(n, np.int32)
        cen = np.empty((k, d), np.float64, order="F")
        norm = C.c_double()
        nit = C.c_int32()
        rc = self.lib.oracle_sparse_k_means(_p(X, _dp), C.c_int64(n), d, C.c_uint32(k), C.c_double(reg), int(no_zero),
                                            C.c_uint32(fit_id), _p(idx, _ip), _p(cen, _dp), C.byref(norm), C.byref(nit))
        assert rc == 0
        return {"i": idx, "c": np.array(cen), "n": norm.value, "n_iter": nit.value}

    def grid_search(self, X, k, reg, iterations: int, boot_n: int):
        X = _f(X)
        n, d = X.shape
        k = _i32(np.atleast_1d(k))
        reg = np.ascontiguousarray(np.atleast_1d(reg), dtype=np.float64)
        nk, nreg = len(k), len(reg)
        ari = np.empty((nk, nreg), np.float64, order="F")
        ncl = np.empty((nk, nreg), np.float64, order="F")
        tot = int(iterations * nreg * 2 * d * int(k.sum()))
        cen = np.empty(tot, np.float64)
        rc = self.lib.oracle_grid_search(_p(X, _dp), C.c_int64(n), d, _p(k, _ip), nk, _p(reg, _dp), nreg,
                                         C.c_uint32(iterations), C.c_uint32(boot_n), _p(ari, _dp), _p(ncl, _dp), _p(cen, _dp))
        assert rc == 0
        return {"d": np.array(ari), "n": np.array(ncl), "c": _unpack_centers(cen, k, nreg, iterations, d)}

    def rand_index(self, a, b, k: int, pair_id: int = 0) -> float:
        a, b = _i32(a), _i32(b)
        out = C.c_double()
        rc = self.lib.oracle_rand_index(_p(a, _ip), _p(b, _ip), C.c_int64(len(a)), C.c_uint32(k), C.c_uint32(pair_id), C.byref(out))
        if rc != 0:
            raise ValueError("oracle_rand_index: bad labels")
        return out.value

    def rand_index_exact(self, a, b, k: int) -> float:
        a, b = _i32(a), _i32(b)
        out = C.c_double()
        rc = self.lib.oracle_rand_index_exact(_p(a, _ip), _p(b, _ip), C.c_int64(len(a)), C.c_uint32(k), C.byref(out))
        if rc != 0:
            raise ValueError("oracle_rand_index_exact: bad labels")
        return out.value

    # -- stages --------------------------------------------------------------------------------
    def find_centers_from(self, X, init_mu, reg: float, no_zero: bool, max_iter: int = 1000):
        X, mu0 = _f(X), _f(init_mu)
        n, d = X.shape
        k = mu0.shape[0]
        idx = np.empty(n, np.int32)
        cen = np.empty((k, d), np.float64, order="F")
        norm = C.c_double()
        nit = C.c_int32()
        rc = self.lib.oracle_find_centers_from(_p(X, _dp), C.c_int64(n), d, _p(mu0, _dp), C.c_uint32(k), C.c_double(reg),
                                               int(no_zero), max_iter, _p(idx, _ip), _p(cen, _dp), C.byref(norm), C.byref(nit))
        assert rc == 0
        return {"i": idx, "c": np.array(cen), "n": norm.value, "n_iter": nit.value}

    def reevaluate_centers(self, X, idx, k: int, reg: float):
        X, idx = _f(X), _i32(idx)
        n, d = X.shape
        cen = np.empty((k, d), np.float64, order="F")
        cnt = np.empty(k, np.int64)
        self.lib.oracle_reevaluate_centers(_p(X, _dp), C.c_int64(n), d, _p(idx, _ip), C.c_uint32(k), C.c_double(reg), _p(cen, _dp), _p(cnt, _lp))
        return np.array(cen), cnt

    def initialize_mu(self, X, k: int, fit_id: int = 0):
        X = _f(X)
        n, d = X.shape
        mu = np.empty((k, d), np.float64, order="F")
        rows = np.empty(k, np.int64)
        rc = self.lib.oracle_initialize_mu(_p(X, _dp), C.c_int64(n), d, C.c_uint32(k), C.c_uint32(fit_id), _p(mu, _dp), _p(rows, _lp))
        assert rc == 0
        return np.array(mu), rows

    def bootstrap_rows(self, n: int, boot_n: int, fit_id: int = 0) -> np.ndarray:
        rows = np.empty(boot_n, np.int64)
        self.lib.oracle_bootstrap_rows(C.c_int64(n), C.c_uint32(boot_n), C.c_uint32(fit_id), _p(rows, _lp))
        return rows

    def cluster_norm(self, X, mu, idx, reg: float) -> float:
        X, mu, idx = _f(X), _f(mu), _i32(idx)
        out = C.c_double()
        self.lib.oracle_cluster_norm(_p(X, _dp), C.c_int64(X.shape[0]), X.shape[1], _p(mu, _dp), C.c_uint32(mu.shape[0]), _p(idx, _ip),
                                     C.c_double(reg), C.byref(out))
        return out.value

    def n_used_clusters(self, idx, k: int) -> int:
        idx = _i32(idx)
        out = C.c_uint32()
        self.lib.oracle_n_used_clusters(_p(idx, _ip), C.c_int64(len(idx)), C.c_uint32(k), C.byref(out))
        return out.value

    def time_lloyd(self, X, mu, reg: float, no_zero: bool, iters: int) -> float:
        X, mu = _f(X), _f(mu)
        out = C.c_double()
        self.lib.oracle_time_lloyd(_p(X, _dp), C.c_int64(X.shape[0]), X.shape[1], _p(mu, _dp), C.c_uint32(mu.shape[0]), C.c_double(reg),
                                   int(no_zero), iters, C.byref(out))
        return out.value


def _unpack_centers(flat, k, nreg, iterations, d):
    """[(it, j, l)] -> [ret1 centres, ret2 centres] in the reference's loop order (Clusterer.cpp:349-352)."""
    out, off = [], 0
    for _it in range(iterations):
        for kj in k:
            for _l in range(nreg):
                pair = []
                for _w in range(2):
                    pair.append(np.array(flat[off:off + kj * d].reshape((kj, d), order="F")))
                    off += kj * d
                out.append(pair)
    return out


class Ref:
    """The reference's own engine (oracle/_ref).  ``available()`` is False where it was not built."""

    PATH = os.path.join(_HERE, "_ref", "libdepecher_ref.so")

    @classmethod
    def available(cls) -> bool:
        return os.path.exists(cls.PATH)

    def __init__(self):
        self.lib = C.CDLL(self.PATH)
        self.lib.ref_set_time.argtypes = [C.c_long]
        self.lib.ref_set_time.restype = None

    def set_time(self, t: int) -> None:
        """value of time(0) seen by Clusterer(): the stream is srand(t + seed_off_set)."""
        self.lib.ref_set_time(C.c_long(t))

    def allocate_points(self, X, mu, no_zero):
        X, mu = _f(X), _f(mu)
        n, d = X.shape
        out = np.empty(n, np.int32)
        self.lib.ref_allocate_points(_p(X, _dp), C.c_int64(n), d, _p(mu, _dp), mu.shape[0], int(no_zero), _p(out, _ip))
        return out

    def sparse_k_means(self, X, k, reg, no_zero, seed_off_set=0):
        X = _f(X)
        n, d = X.shape
        idx = np.empty(n, np.int32)
        cen = np.empty((k, d), np.float64, order="F")
        norm = C.c_double()
        self.lib.ref_sparse_k_means(_p(X, _dp), C.c_int64(n), d, C.c_uint32(k), C.c_double(reg), int(no_zero), C.c_ulong(seed_off_set),
                                    _p(idx, _ip), _p(cen, _dp), C.byref(norm))
        return {"i": idx, "c": np.array(cen), "n": norm.value}

    def grid_search(self, X, k, reg, iterations, boot_n, seed_off_set=0):
        X = _f(X)
        n, d = X.shape
        k = _i32(np.atleast_1d(k))
        reg = np.ascontiguousarray(np.atleast_1d(reg), dtype=np.float64)
        nk, nreg = len(k), len(reg)
        ari = np.empty((nk, nreg), np.float64, order="F")
        ncl = np.empty((nk, nreg), np.float64, order="F")
        cen = np.empty(int(iterations * nreg * 2 * d * int(k.sum())), np.float64)
        self.lib.ref_grid_search(_p(X, _dp), C.c_int64(n), d, _p(k, _ip), nk, _p(reg, _dp), nreg, C.c_uint32(iterations), C.c_uint32(boot_n),
                                 C.c_ulong(seed_off_set), _p(ari, _dp), _p(ncl, _dp), _p(cen, _dp))
        return {"d": np.array(ari), "n": np.array(ncl), "c": _unpack_centers(cen, k, nreg, iterations, d)}

    def rand_index(self, a, b, k, seed_off_set=0):
        a, b = _i32(a), _i32(b)
        out = C.c_double()
        self.lib.ref_rand_index(_p(a, _ip), _p(b, _ip), C.c_int64(len(a)), C.c_uint32(k), C.c_ulong(seed_off_set), C.byref(out))
        return out.value

    def reevaluate_centers(self, X, idx, k, reg):
        X, idx = _f(X), _i32(idx)
        n, d = X.shape
        cen = np.empty((k, d), np.float64, order="F")
        self.lib.ref_reevaluate_centers(_p(X, _dp), C.c_int64(n), d, _p(idx, _ip), C.c_uint32(k), C.c_double(reg), _p(cen, _dp))
        return np.array(cen)

    def initialize_mu(self, X, k, seed_off_set=0):
        X = _f(X)
        n, d = X.shape
        mu = np.empty((k, d), np.float64, order="F")
        self.lib.ref_initialize_mu(_p(X, _dp), C.c_int64(n), d, C.c_uint32(k), C.c_ulong(seed_off_set), _p(mu, _dp))
        return np.array(mu)

    def bootstrap_data(self, X, boot_n, seed_off_set=0):
        X = _f(X)
        n, d = X.shape
        out = np.empty((boot_n, d), np.float64, order="F")
        self.lib.ref_bootstrap_data(_p(X, _dp), C.c_int64(n), d, C.c_uint32(boot_n), C.c_ulong(seed_off_set), _p(out, _dp))
        return np.array(out)

    def cluster_norm(self, X, mu, idx, reg):
        X, mu, idx = _f(X), _f(mu), _i32(idx)
        out = C.c_double()
        self.lib.ref_cluster_norm(_p(X, _dp), C.c_int64(X.shape[0]), X.shape[1], _p(mu, _dp), C.c_uint32(mu.shape[0]), _p(idx, _ip),
                                  C.c_double(reg), C.byref(out))
        return out.value

    def n_used_clusters(self, idx, k):
        idx = _i32(idx)
        out = C.c_uint32()
        self.lib.ref_n_used_clusters(_p(idx, _ip), C.c_int64(len(idx)), C.c_uint32(k), C.byref(out))
        return out.value

    def time_lloyd(self, X, mu, reg, no_zero, iters):
        X, mu = _f(X), _f(mu)
        out = C.c_double()
        self.lib.ref_time_lloyd(_p(X, _dp), C.c_int64(X.shape[0]), X.shape[1], _p(mu, _dp), C.c_uint32(mu.shape[0]), C.c_double(reg),
                                int(no_zero), iters, C.byref(out))
        return out.value
