"""ctypes binding of the C ABI in include/depecher_b200.h (libdepecher_b200.so).

This is the same boundary an Rcpp layer binds: plain pointers and sizes, column-major
doubles as R stores them, 0-based int32 labels.  There is no CPU fallback: every compute call fails with
``DeviceError`` when the library cannot find a B200, and ``import`` fails loudly when the shared library
was not built (``python -c "import __graft_entry__ as g; g.build()"``).
"""
from __future__ import annotations

import ctypes as C
import os
import re

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libdepecher_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "depecher_b200.h")

ARI_EXACT, ARI_MONTE_CARLO = 0, 1


class DepecherError(RuntimeError):
    pass


class DeviceError(DepecherError):
    pass


_dp, _fp = C.POINTER(C.c_double), C.POINTER(C.c_float)
_ip, _lp = C.POINTER(C.c_int32), C.POINTER(C.c_int64)


def declared_symbols(header: str = HEADER_PATH) -> list[str]:
    """every function the public header declares (used by the CPU test that the library exports them all)"""
    txt = open(header).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(dpr_[a-z0-9_]+)\s*\(", txt)))


def load_library(path: str = LIB_PATH) -> C.CDLL:
    if not os.path.exists(path):
        raise DepecherError(f"{path} is missing: build it first (make -C depecher_b200/csrc, or __graft_entry__.build())")
    lib = C.CDLL(path)
    lib.dpr_last_error.restype = C.c_char_p
    return lib


def _f64(a) -> np.ndarray:
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


def _ptr(a, t):
    return a.ctypes.data_as(t)


class Dataset:
    """X resident in HBM as column-major fp32 (dpr_dataset_t)."""

    def __init__(self, engine: "Engine", handle, n: int, d: int):
        self.engine, self.handle, self.n, self.d = engine, handle, n, d

    def close(self) -> None:
        if self.handle is not None:
            self.engine.lib.dpr_dataset_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class PenaltyOpt:
    """depechePenaltyOpt's state machine (dpr_penalty_opt_*; R/depechePenaltyOpt.R:14-265).  Pure host code inside the library:
    usable without a device, with any producer of the sets (`used()` -> results of one set -> `feed()`)."""

    def __init__(self, penalties, k: int, d: int, sampleSize: int, nCores: int, maxIter: int, minARIImprovement: float, optimARI: float,
                 lib: C.CDLL | None = None):
        self.lib = lib or load_library()
        self.penalties = np.ascontiguousarray(np.atleast_1d(penalties), dtype=np.float64)
        self.k, self.d, self.nCores = int(k), int(d), int(nCores)
        self.handle = C.c_void_p()
        self._chk(self.lib.dpr_penalty_opt_create(_ptr(self.penalties, _dp), C.c_int32(len(self.penalties)), C.c_int32(k), C.c_int32(d),
                                                  C.c_uint32(sampleSize), C.c_int32(nCores), C.c_int32(maxIter), C.c_double(minARIImprovement),
                                                  C.c_double(optimARI), C.byref(self.handle)))

    def _chk(self, rc: int) -> None:
        if rc != 0:
            raise DepecherError(f"[{rc}] {(self.lib.dpr_last_error() or b'').decode()}")

    def close(self) -> None:
        if self.handle:
            self.lib.dpr_penalty_opt_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def used(self):
        """(positions of the penalties still in use, their scaled values for grid_search, loop continues?)"""
        n, cont = C.c_int32(), C.c_int32()
        pos = np.empty(len(self.penalties), np.int32)
        reg = np.empty(len(self.penalties), np.float64)
        self._chk(self.lib.dpr_penalty_opt_used(self.handle, C.byref(n), _ptr(pos, _ip), _ptr(reg, _dp), C.byref(cont)))
        return pos[:n.value].copy(), reg[:n.value].copy(), bool(cont.value)

    def feed(self, ari_each, nclust_each, centres=None) -> bool:
        """one set: (nCores, n_used) ARI and cluster-count arrays, centres (nCores, n_used, 2, k, d) or None; returns `continue`"""
        a = np.ascontiguousarray(ari_each, dtype=np.float64)
        m = np.ascontiguousarray(nclust_each, dtype=np.float64)
        cont = C.c_int32()
        n_used = len(self.used()[0])
        if a.shape != (self.nCores, n_used) or m.shape != (self.nCores, n_used):
            raise DepecherError(f"feed: expected {(self.nCores, n_used)} ARI / cluster-count arrays, got {a.shape} and {m.shape}")
        cp = None
        if centres is not None:
            if np.shape(centres) != (self.nCores, n_used, 2, self.k, self.d):
                raise DepecherError(f"feed: expected centres of shape {(self.nCores, n_used, 2, self.k, self.d)}, got {np.shape(centres)}")
            cen = np.ascontiguousarray(np.asarray(centres, dtype=np.float64).transpose(0, 1, 2, 4, 3))   # k x d col-major per matrix
            cp = _ptr(cen, _dp)
        self._chk(self.lib.dpr_penalty_opt_feed(self.handle, _ptr(a, _dp), _ptr(m, _dp), cp, C.byref(cont)))
        return bool(cont.value)

    def result(self) -> dict:
        n_pen = len(self.penalties)
        best, pos, wi, ns, warn = C.c_double(), C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        ari, ncl = np.empty(n_pen), np.empty(n_pen)
        self._chk(self.lib.dpr_penalty_opt_result(self.handle, C.byref(best), C.byref(pos), _ptr(ari, _dp), _ptr(ncl, _dp), C.byref(wi), C.byref(ns),
                                                  C.byref(warn)))
        sol = np.empty((ns.value, self.d, self.k), np.float64)
        if ns.value:
            self._chk(self.lib.dpr_penalty_opt_solutions(self.handle, _ptr(sol, _dp)))
        return {"bestPenalty": best.value, "bestPenaltyPosition": pos.value, "meanARI": ari, "meanClust": ncl, "workerIterations": wi.value,
                "bestClusterCenters": [np.array(m.T) for m in sol], "warn": warn.value}


class Engine:
    """Thin object wrapper over the C ABI; one per process (the library keeps one CUDA context)."""

    def __init__(self, device: int | None = None, path: str = LIB_PATH):
        self.lib = load_library(path)
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        self._chk(self.lib.dpr_set_device(C.c_int32(device)))
        self.device = device

    # -- plumbing ------------------------------------------------------------------------------------
    def _chk(self, rc: int) -> None:
        if rc == 0:
            return
        msg = (self.lib.dpr_last_error() or b"").decode()
        if rc == -2:
            raise DeviceError(msg)
        raise DepecherError(f"[{rc}] {msg}")

    def device_count(self) -> int:
        out = C.c_int32()
        self.lib.dpr_device_count(C.byref(out))
        return out.value

    def set_stream(self, cuda_stream_ptr: int | None) -> None:
        self._chk(self.lib.dpr_set_stream(C.c_void_p(cuda_stream_ptr or 0)))

    def synchronize(self) -> None:
        self._chk(self.lib.dpr_synchronize())

    def recheck_count(self, reset: bool = True) -> int:
        """Cells decided by the exact fp64 path since the last reset (diagnostics)."""
        out = C.c_int64(0)
        self._chk(self.lib.dpr_recheck_count(C.byref(out), C.c_int32(1 if reset else 0)))
        return out.value

    def set_pass_kernel(self, which: int) -> None:
        """0 = automatic (tensor-core pass when the shape fits), 1 = always the FFMA pass."""
        self._chk(self.lib.dpr_set_pass_kernel(C.c_int32(which)))

    def set_ari_mode(self, mode: int) -> None:
        self._chk(self.lib.dpr_set_ari_mode(C.c_int32(mode)))

    def pass_timing(self, enable: bool):
        """(summed ms, launches) of the fused pass since the last call; enable/disable further timing"""
        ms, cnt = C.c_double(), C.c_int64()
        self._chk(self.lib.dpr_pass_timing(C.c_int32(int(enable)), C.byref(ms), C.byref(cnt)))
        return ms.value, cnt.value

    def kernel_launches(self, reset: bool = False) -> int:
        out = C.c_int64()
        self._chk(self.lib.dpr_kernel_launches(C.byref(out), C.c_int32(int(reset))))
        return out.value

    # -- datasets ------------------------------------------------------------------------------------
    def dataset(self, X) -> Dataset:
        X = np.asarray(X)
        h = C.c_void_p()
        if X.dtype == np.float32:
            Xf = np.asfortranarray(X)
            self._chk(self.lib.dpr_dataset_create_f32(_ptr(Xf, _fp), C.c_int64(X.shape[0]), C.c_int32(X.shape[1]), C.byref(h)))
        else:
            Xf = _f64(X)
            self._chk(self.lib.dpr_dataset_create(_ptr(Xf, _dp), C.c_int64(X.shape[0]), C.c_int32(X.shape[1]), C.byref(h)))
        return Dataset(self, h, X.shape[0], X.shape[1])

    def dataset_scaled(self, X, log2Off=False, center="default", scale=True):
        """depecheLogCenterSd (R/depecheLogCenterSd.R:6-115) on the device -> (resident scaled Dataset, [log2 applied, centres | False, sd])"""
        Xf = _f64(X)
        n, d = Xf.shape
        centres = np.zeros(d, np.float64)
        if isinstance(center, str):
            if center not in ("default", "peak", "mean"):
                raise DepecherError("center must be 'default', 'peak', 'mean', False or a vector")
            mode = 1 if (center == "peak" or (center == "default" and d <= 100)) else 2
        elif center is False:
            mode = 0
        else:
            mode = 3
            centres[:] = np.asarray(center, np.float64)
        sd = C.c_double(1.0)
        if scale is True:
            smode = 1
        elif scale is False:
            smode = 0
        else:
            smode, sd = 2, C.c_double(float(scale))
        applied, h = C.c_int32(), C.c_void_p()
        self._chk(self.lib.dpr_dataset_create_scaled(_ptr(Xf, _dp), C.c_int64(n), C.c_int32(d), C.c_int32(int(bool(log2Off))), C.c_int32(mode),
                                                     _ptr(centres, _dp), C.c_int32(smode), C.byref(sd), C.byref(applied), C.byref(h)))
        info = [bool(applied.value), (centres if mode != 0 else False), (sd.value if smode != 0 else 1)]
        return Dataset(self, h, n, d), info

    def synthetic(self, n: int, d: int, groups: int, amplitude: float, seed: int, row0: int = 0, scale: float = 0.0) -> Dataset:
        h = C.c_void_p()
        used = C.c_double()
        self._chk(self.lib.dpr_dataset_create_synthetic(C.c_int64(n), C.c_int32(d), C.c_int32(groups), C.c_double(amplitude),
                                                        C.c_uint64(seed), C.c_int64(row0), C.c_double(scale), C.byref(used), C.byref(h)))
        ds = Dataset(self, h, n, d)
        ds.scale = used.value
        return ds

    def read(self, ds: Dataset, row0: int, rows: int) -> np.ndarray:
        out = np.empty((rows, ds.d), np.float64, order="F")
        self._chk(self.lib.dpr_dataset_read(ds.handle, C.c_int64(row0), C.c_int64(rows), _ptr(out, _dp)))
        return out

    def gather(self, ds: Dataset, rows) -> Dataset:
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        h = C.c_void_p()
        self._chk(self.lib.dpr_dataset_gather(ds.handle, _ptr(rows, _lp), C.c_int64(len(rows)), C.byref(h)))
        return Dataset(self, h, len(rows), ds.d)

    def _as_ds(self, X):
        if isinstance(X, Dataset):
            return X, False
        return self.dataset(X), True

    # -- the four entry points (names and results as the reference's Rcpp exports) ----------------------
    def allocate_points(self, X, mu, no_zero, base: int = 0) -> dict:
        """allocate_points(X, mu, no_zero) -> {'i': labels}   (src/InterfaceUtils.cpp:134-148); `base` is added to every label"""
        mu = _f64(mu)
        k = mu.shape[0]
        if isinstance(X, Dataset):
            idx = np.empty(X.n, np.int32)
            self._chk(self.lib.dpr_dataset_allocate_points_base(X.handle, _ptr(mu, _dp), C.c_int32(k), C.c_int32(int(no_zero)), C.c_int32(base),
                                                                _ptr(idx, _ip)))
        else:
            Xf = _f64(X)
            idx = np.empty(Xf.shape[0], np.int32)
            self._chk(self.lib.dpr_allocate_points(_ptr(Xf, _dp), C.c_int64(Xf.shape[0]), C.c_int32(Xf.shape[1]), _ptr(mu, _dp), C.c_int32(k),
                                                   C.c_int32(int(no_zero)), _ptr(idx, _ip)))
            if base:
                idx += base
        return {"i": idx}

    def allocate_points_ranked(self, ds: "Dataset", mu, no_zero, out=None):
        """depeche()'s final step on a resident matrix: labels renumbered 1, 2, ... by decreasing cluster size, and the old 0-based labels in
        rank order (R/depeche.R:228-252).  out: an int32 vector of ds.n entries to write the labels into (depeche() hands over one whose
        pages it has touched while the search ran)"""
        mu = _f64(mu)
        k = mu.shape[0]
        idx = out if (out is not None and out.dtype == np.int32 and out.shape == (ds.n,) and out.flags.c_contiguous) else np.empty(ds.n, np.int32)
        order = np.empty(k, np.int32)
        self._chk(self.lib.dpr_dataset_allocate_points_ranked(ds.handle, _ptr(mu, _dp), C.c_int32(k), C.c_int32(int(no_zero)), _ptr(idx, _ip), _ptr(order, _ip)))
        return idx, order[order >= 0]

    def sparse_k_means(self, X, k, reg, no_zero, seed_off_set=0, want_labels=True) -> dict:
        """sparse_k_means(X, k, reg, no_zero, seed) -> {'i','o','c','v','n','m'}   (src/InterfaceUtils.cpp:66-94)"""
        nit, norm = C.c_int32(), C.c_double()
        if isinstance(X, Dataset):
            n, d = X.n, X.d
            idx = np.empty(n, np.int32) if want_labels else None
            cen = np.empty((k, d), np.float64, order="F")
            self._chk(self.lib.dpr_dataset_sparse_k_means(X.handle, C.c_uint32(k), C.c_double(reg), C.c_int32(int(no_zero)), C.c_uint64(seed_off_set),
                                                          _ptr(idx, _ip) if want_labels else None, _ptr(cen, _dp), C.byref(norm), C.byref(nit)))
        else:
            Xf = _f64(X)
            n, d = Xf.shape
            idx = np.empty(n, np.int32)
            cen = np.empty((k, d), np.float64, order="F")
            self._chk(self.lib.dpr_sparse_k_means(_ptr(Xf, _dp), C.c_int64(n), C.c_int32(d), C.c_uint32(k), C.c_double(reg), C.c_int32(int(no_zero)),
                                                  C.c_uint64(seed_off_set), _ptr(idx, _ip), _ptr(cen, _dp), C.byref(norm), C.byref(nit)))
        cen = np.array(cen)
        return {"i": idx, "o": idx, "c": cen, "v": cen, "n": norm.value, "m": norm.value, "n_iter": nit.value}

    def grid_search(self, X, k, reg, iterations, bootstrapSamples, seed_off_set=0) -> dict:
        """grid_search(X, k, reg, iterations, bootstrapSamples, seed) -> {'d','z','n','m','c'}   (src/InterfaceUtils.cpp:99-131)"""
        k = _i32(np.atleast_1d(k))
        reg = np.ascontiguousarray(np.atleast_1d(reg), dtype=np.float64)
        nk, nreg = len(k), len(reg)
        ari = np.empty((nk, nreg), np.float64, order="F")
        ncl = np.empty((nk, nreg), np.float64, order="F")
        ds, own = self._as_ds(X)
        try:
            cen = np.empty(int(iterations) * nreg * 2 * ds.d * int(k.sum()), np.float64)
            self._chk(self.lib.dpr_dataset_grid_search(ds.handle, _ptr(k, _ip), C.c_int32(nk), _ptr(reg, _dp), C.c_int32(nreg), C.c_uint32(iterations),
                                                       C.c_uint32(bootstrapSamples), C.c_uint64(seed_off_set), _ptr(ari, _dp), _ptr(ncl, _dp),
                                                       _ptr(cen, _dp)))
            d = ds.d
        finally:
            if own:
                ds.close()
        centers, off = [], 0
        for _it in range(int(iterations)):
            for kj in k:
                for _l in range(nreg):
                    mats = []
                    for _w in range(2):
                        mats.append(np.array(cen[off:off + kj * d].reshape((kj, d), order="F")))
                        off += kj * d
                    centers.append(mats + mats)  # entries [3],[4] of the reference's list repeat [1],[2] (Clusterer.cpp:363-368)
        ari, ncl = np.array(ari), np.array(ncl)
        return {"d": ari, "z": ari, "n": ncl, "m": ncl, "c": centers}

    def grid_search_each(self, X, k, reg, iterations, bootstrapSamples, seed_off_set=0) -> dict:
        """`iterations` workers' grid_search(X, k, reg, 1, bootstrapSamples, .) results as one batch (R/depechePenaltyOpt.R:50-55):
        z, m: lists (one per worker) of nk x nreg matrices; c: per worker, per (k, penalty) the list [ret1, ret2, ret1, ret2]."""
        k = _i32(np.atleast_1d(k))
        reg = np.ascontiguousarray(np.atleast_1d(reg), dtype=np.float64)
        nk, nreg, iterations = len(k), len(reg), int(iterations)
        ari = np.empty((iterations, nreg, nk), np.float64)
        ncl = np.empty((iterations, nreg, nk), np.float64)
        ds, own = self._as_ds(X)
        try:
            cen = np.empty(iterations * nreg * 2 * ds.d * int(k.sum()), np.float64)
            self._chk(self.lib.dpr_dataset_grid_search_each(ds.handle, _ptr(k, _ip), C.c_int32(nk), _ptr(reg, _dp), C.c_int32(nreg),
                                                            C.c_uint32(iterations), C.c_uint32(bootstrapSamples), C.c_uint64(seed_off_set),
                                                            _ptr(ari, _dp), _ptr(ncl, _dp), _ptr(cen, _dp)))
            d = ds.d
        finally:
            if own:
                ds.close()
        z = [ari[w].T.copy() for w in range(iterations)]
        m = [ncl[w].T.copy() for w in range(iterations)]
        c, off = [], 0
        for _w in range(iterations):
            per = []
            for kj in k:
                for _l in range(nreg):
                    mats = []
                    for _q in range(2):
                        mats.append(np.array(cen[off:off + kj * d].reshape((kj, d), order="F")))
                        off += kj * d
                    per.append(mats + mats)
            c.append(per)
        return {"d": z, "z": z, "n": m, "m": m, "c": c}

    def grid_search_range(self, ds: Dataset, k: int, reg, iterations: int, bootstrapSamples: int, seed_off_set: int, first: int, count: int):
        """problems [first, first + count) of the (worker, penalty) grid of one k -> (ari[count], nclust[count], centres[count, 2, k, d])"""
        reg = np.ascontiguousarray(np.atleast_1d(reg), dtype=np.float64)
        kk = _i32([k])
        ari, ncl = np.empty(count), np.empty(count)
        cen = np.empty((count, 2, ds.d, k), np.float64)
        self._chk(self.lib.dpr_dataset_grid_search_range(ds.handle, _ptr(kk, _ip), C.c_int32(1), _ptr(reg, _dp), C.c_int32(len(reg)), C.c_uint32(iterations),
                                                         C.c_uint32(bootstrapSamples), C.c_uint64(seed_off_set), C.c_int32(first), C.c_int32(count),
                                                         _ptr(ari, _dp), _ptr(ncl, _dp), _ptr(cen, _dp)))
        return ari, ncl, cen.transpose(0, 1, 3, 2).copy()

    def penalty_opt(self, ds: Dataset, opt: PenaltyOpt, seed: int = 0, split_over_ranks: bool = False) -> None:
        """the whole adaptive search of depechePenaltyOpt on a resident dataset in ONE call (dpr_dataset_penalty_opt)"""
        self._chk(self.lib.dpr_dataset_penalty_opt(ds.handle, opt.handle, C.c_uint64(seed), C.c_int32(int(split_over_ranks))))

    def select_solution(self, selection_set, solutions):
        """depecheClusterCenterSelection.R:11-37 -> (index of the most central solution, mean ARI of each)"""
        sols = [np.asarray(m, dtype=np.float64) for m in solutions]
        m, k = len(sols), sols[0].shape[0]
        flat = np.concatenate([np.asfortranarray(mm).ravel(order="F") for mm in sols])
        ds, own = self._as_ds(selection_set)
        try:
            best = C.c_int32()
            mean = np.empty(m, np.float64)
            self._chk(self.lib.dpr_dataset_select_solution(ds.handle, _ptr(flat, _dp), C.c_int32(m), C.c_int32(k), C.byref(best), _ptr(mean, _dp)))
        finally:
            if own:
                ds.close()
        return best.value, mean

    def rand_index(self, inds1, inds2, k) -> float:
        """rand_index(inds1, inds2, k) -> double   (src/InterfaceUtils.cpp:151-161)"""
        a, b = _i32(inds1), _i32(inds2)
        if len(a) != len(b):
            raise DepecherError("rand_index: label vectors differ in length")
        out = C.c_double()
        self._chk(self.lib.dpr_rand_index(_ptr(a, _ip), _ptr(b, _ip), C.c_int64(len(a)), C.c_uint32(k), C.byref(out)))
        return out.value

    # -- stage-level hooks -----------------------------------------------------------------------------
    def find_centers_from(self, X, init_mu, reg, no_zero, max_iter=1000) -> dict:
        mu0 = _f64(init_mu)
        k = mu0.shape[0]
        ds, own = self._as_ds(X)
        try:
            idx = np.empty(ds.n, np.int32)
            cen = np.empty((k, ds.d), np.float64, order="F")
            nit, norm = C.c_int32(), C.c_double()
            self._chk(self.lib.dpr_dataset_find_centers_from(ds.handle, _ptr(mu0, _dp), C.c_uint32(k), C.c_double(reg), C.c_int32(int(no_zero)),
                                                             C.c_int32(max_iter), _ptr(idx, _ip), _ptr(cen, _dp), C.byref(norm), C.byref(nit)))
        finally:
            if own:
                ds.close()
        return {"i": idx, "c": np.array(cen), "n": norm.value, "n_iter": nit.value}

    def reevaluate_centers(self, X, idx, k, reg):
        idx = _i32(idx)
        ds, own = self._as_ds(X)
        try:
            cen = np.empty((k, ds.d), np.float64, order="F")
            cnt = np.empty(k, np.int64)
            self._chk(self.lib.dpr_dataset_reevaluate_centers(ds.handle, _ptr(idx, _ip), C.c_uint32(k), C.c_double(reg), _ptr(cen, _dp), _ptr(cnt, _lp)))
        finally:
            if own:
                ds.close()
        return np.array(cen), cnt

    def initialize_mu(self, X, k, seed=0, fit_id=0):
        ds, own = self._as_ds(X)
        try:
            mu = np.empty((k, ds.d), np.float64, order="F")
            rows = np.empty(k, np.int64)
            self._chk(self.lib.dpr_dataset_initialize_mu(ds.handle, C.c_uint32(k), C.c_uint64(seed), C.c_uint32(fit_id), _ptr(mu, _dp), _ptr(rows, _lp)))
        finally:
            if own:
                ds.close()
        return np.array(mu), rows

    def bootstrap_rows(self, n, boot_n, seed=0, fit_id=0) -> np.ndarray:
        rows = np.empty(boot_n, np.int64)
        self._chk(self.lib.dpr_bootstrap_rows(C.c_int64(n), C.c_uint32(boot_n), C.c_uint64(seed), C.c_uint32(fit_id), _ptr(rows, _lp)))
        return rows

    def cluster_norm(self, X, mu, idx, reg) -> float:
        mu, idx = _f64(mu), _i32(idx)
        ds, own = self._as_ds(X)
        try:
            out = C.c_double()
            self._chk(self.lib.dpr_dataset_cluster_norm(ds.handle, _ptr(mu, _dp), C.c_uint32(mu.shape[0]), _ptr(idx, _ip), C.c_double(reg), C.byref(out)))
        finally:
            if own:
                ds.close()
        return out.value

    def n_used_clusters(self, idx, k) -> int:
        idx = _i32(idx)
        out = C.c_uint32()
        self._chk(self.lib.dpr_n_used_clusters(_ptr(idx, _ip), C.c_int64(len(idx)), C.c_uint32(k), C.byref(out)))
        return out.value

    def allocate_many(self, X, mus, no_zero, want_labels=True, want_ari=True):
        """labels of X under each of the m centre sets + the m x m all-pairs ARI (depecheClusterCenterSelection.R:11-29)"""
        mus = [np.asarray(m, dtype=np.float64) for m in mus]
        m, k = len(mus), mus[0].shape[0]
        flat = np.concatenate([np.asfortranarray(mm).ravel(order="F") for mm in mus])
        ds, own = self._as_ds(X)
        try:
            lab = np.empty((m, ds.n), np.int32) if want_labels else None
            ari = np.empty((m, m), np.float64, order="F") if want_ari else None
            self._chk(self.lib.dpr_dataset_allocate_many(ds.handle, _ptr(flat, _dp), C.c_int32(m), C.c_int32(k), C.c_int32(int(no_zero)),
                                                         _ptr(lab, _ip) if want_labels else None, _ptr(ari, _dp) if want_ari else None))
        finally:
            if own:
                ds.close()
        return lab, (np.array(ari) if want_ari else None)

    # -- device-resident loops (bench) -------------------------------------------------------------------
    def lloyd_iterations(self, ds: Dataset, mu, reg, no_zero, ramp_i, iters):
        mu = _f64(np.array(mu, dtype=np.float64))
        changed = C.c_int64()
        self._chk(self.lib.dpr_dataset_lloyd_iterations(ds.handle, _ptr(mu, _dp), C.c_uint32(mu.shape[0]), C.c_double(reg), C.c_int32(int(no_zero)),
                                                        C.c_int32(ramp_i), C.c_int32(iters), C.byref(changed)))
        return np.array(mu), changed.value

    def assign_iterations(self, ds: Dataset, mu, no_zero, iters) -> None:
        mu = _f64(mu)
        self._chk(self.lib.dpr_dataset_assign_iterations(ds.handle, _ptr(mu, _dp), C.c_uint32(mu.shape[0]), C.c_int32(int(no_zero)), C.c_int32(iters)))

    # -- multi-GPU --------------------------------------------------------------------------------------------
    def comm_unique_id(self, nccl_path: str | None = None) -> bytes:
        buf = C.create_string_buffer(128)
        self._chk(self.lib.dpr_comm_unique_id(buf, nccl_path.encode() if nccl_path else None))
        return buf.raw

    def comm_init(self, rank: int, world: int, unique_id: bytes | None, nccl_path: str | None = None) -> None:
        self._chk(self.lib.dpr_comm_init(C.c_int32(rank), C.c_int32(world), unique_id, nccl_path.encode() if nccl_path else None))

    def comm_peer_handle(self) -> bytes:
        """this rank's mailbox for the peer-memory exchange: its 64-byte CUDA IPC handle"""
        buf = C.create_string_buffer(64)
        self._chk(self.lib.dpr_comm_peer_handle(buf))
        return buf.raw

    def comm_peer_attach(self, handles: list[bytes]) -> None:
        """handles of all ranks in rank order: the per-iteration exchange then goes through peer memory instead of ncclAllReduce"""
        blob = b"".join(handles)
        self._chk(self.lib.dpr_comm_peer_attach(blob, C.c_int32(len(handles))))

    def comm_peer_detach(self) -> None:
        self._chk(self.lib.dpr_comm_peer_detach())

    def comm_destroy(self) -> None:
        self._chk(self.lib.dpr_comm_destroy())
