"""Generates the committed fixtures under tests/golden/.  Run in the BUILD container only:

    python tests/golden/make_golden.py

It needs /root/reference (the bundled RData fixtures) and oracle/_ref/libdepecher_ref.so (the
reference's own src/Clusterer.cpp compiled by oracle/Makefile).  Nothing at test/bench time reads
/root/reference: the GPU box only sees the .npz files this script wrote.

  testdata.npz      data/testData.RData columns 2..15 (the input of the reference's documented
                    depeche(testData[, 2:15]) run) + what data/testDataDepeche.RData stores of its
                    result (ARI / nClust per penalty, bestPenalty, cluster sizes, centres).
  ref_vectors.npz   outputs of the reference engine itself on small seeded inputs, for every
                    function on the hot path, under libc rand() seeded as srand(T + offset).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle.oracle import Ref  # noqa: E402
from rdata import matrix, read_rdata, unwrap  # noqa: E402

REF = "/root/reference"
T0 = 20261019  # the value time(0) is pinned to


def make_testdata():
    df = read_rdata(f"{REF}/data/testData.RData")["testData"]
    cols = [np.asarray(unwrap(c), dtype=np.float64) for c in df.value]
    X = np.stack(cols[1:15], 1)
    names = df.names[1:15]
    # the file holds short decimals: store them as exact integers of 1e-4 units when that is lossless
    q = np.rint(X * 1e4)
    assert np.array_equal(q / 1e4, X), "testData is not 4-decimal"
    res = read_rdata(f"{REF}/data/testDataDepeche.RData")["testDataDepeche"]
    cv = np.asarray(unwrap(res["clusterVector"]), np.int32)
    cc = matrix(res["clusterCenters"])
    pol = res["penaltyOptList"]
    best = float(np.asarray(unwrap(pol[0]["bestPenalty"]))[0])
    opt = pol[1]
    ari = np.asarray(unwrap(opt["ARI"]), np.float64)
    ncl = np.asarray(unwrap(opt["nClust"]), np.float64)
    pens = np.array([float(s) for s in unwrap(opt.attrs["row.names"])])
    np.savez_compressed(os.path.join(HERE, "testdata.npz"), X_1e4=q.astype(np.int32), names=np.array(names),
                        ids=np.asarray(unwrap(cols[0]), np.int32), label=np.asarray(unwrap(cols[15]), np.int32),
                        stored_clusterVector=cv, stored_clusterCenters=cc, stored_bestPenalty=best,
                        stored_penalties=pens, stored_ARI=ari, stored_nClust=ncl)
    print("testdata.npz", X.shape, "best", best, "sizes", np.bincount(cv)[1:])


def blobs(rng, n, d, g, amp=3.0, f32=True):
    cent = rng.choice([-amp, 0.0, amp], size=(g, d))
    X = cent[rng.integers(0, g, n)] + rng.normal(size=(n, d))
    X /= X.std()
    return X.astype(np.float32).astype(np.float64) if f32 else X


def make_ref_vectors():
    assert Ref.available(), "build oracle/_ref first (make -C oracle)"
    r = Ref()
    rng = np.random.default_rng(20261019)
    out = {"T0": T0}

    # --- allocate_points: plain, no_zero with dead rows, < 2 active, exact tie, arbitrary doubles
    X = blobs(rng, 3000, 14, 6)
    mu = blobs(rng, 9, 14, 9)
    mu[[2, 5]] = 0.0
    out["ap_X"], out["ap_mu"] = X, mu
    out["ap_idx"] = r.allocate_points(X, mu, False)
    out["ap_idx_nz"] = r.allocate_points(X, mu, True)
    one = np.zeros_like(mu)
    one[4] = mu[4]
    out["ap_mu_one"] = one
    out["ap_idx_one_nz"] = r.allocate_points(X, one, True)   # one active row -> all zeros (Clusterer.cpp:50-53)
    out["ap_idx_one"] = r.allocate_points(X, one, False)     # 8 identical zero rows + 1: first zero row wins ties
    Xd = blobs(rng, 2000, 7, 4, f32=False)
    mud = blobs(rng, 5, 7, 5, f32=False)
    out["apd_X"], out["apd_mu"], out["apd_idx"] = Xd, mud, r.allocate_points(Xd, mud, False)

    # --- reevaluate_centers incl. an empty cluster, reg = 0 and reg > 0
    idx = rng.integers(0, 7, 3000).astype(np.int32)
    idx[idx == 3] = 4  # cluster 3 empty
    out["rc_idx"] = idx
    out["rc_mu_reg0"] = r.reevaluate_centers(X, idx, 7, 0.0)
    out["rc_mu_reg"] = r.reevaluate_centers(X, idx, 7, 37.5)

    # --- cluster_norm, n_used_clusters
    mu7 = out["rc_mu_reg"]
    out["cn_val"] = r.cluster_norm(X, mu7, idx, 37.5)
    out["nu_val"] = r.n_used_clusters(idx, 7)

    # --- initialize_mu / bootstrap_data / sparse_k_means / grid_search / rand_index under srand(T0 + off)
    r.set_time(T0)
    out["im_mu"] = r.initialize_mu(X, 8, seed_off_set=5)
    r.set_time(T0)
    out["bs_X"] = r.bootstrap_data(X, 500, seed_off_set=6)
    for tag, reg in (("p0", 0.0), ("p1", 12.0), ("p2", 90.0)):
        r.set_time(T0)
        s = r.sparse_k_means(X, 10, reg, True, seed_off_set=7)
        out[f"sk_{tag}_idx"], out[f"sk_{tag}_c"], out[f"sk_{tag}_n"] = s["i"], s["c"], s["n"]
    out["sk_regs"] = np.array([0.0, 12.0, 90.0])
    r.set_time(T0)
    regs = np.array([0.0, 6.0, 25.0, 80.0])
    g = r.grid_search(X, [10], regs, 2, 800, seed_off_set=8)
    out["gs_regs"], out["gs_d"], out["gs_n"] = regs, g["d"], g["n"]
    out["gs_c"] = np.stack([np.stack(p) for p in g["c"]])
    a = rng.integers(0, 6, 5000).astype(np.int32)
    b = np.where(rng.random(5000) < 0.8, a, rng.integers(0, 6, 5000)).astype(np.int32)
    out["ri_a"], out["ri_b"] = a, b
    r.set_time(T0)
    out["ri_val"] = r.rand_index(a, b, 6, seed_off_set=9)
    r.set_time(T0)
    out["ri_same"] = r.rand_index(a + 1, a + 1, 100, seed_off_set=9)   # test_dAllocate.R:20 shape: 1-based, k = 100
    np.savez_compressed(os.path.join(HERE, "ref_vectors.npz"), **out)
    print("ref_vectors.npz", {k: np.shape(v) for k, v in out.items()})


if __name__ == "__main__":
    make_testdata()
    make_ref_vectors()
