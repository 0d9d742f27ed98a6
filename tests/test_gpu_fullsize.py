"""GPU tests at BASELINE.json's full sizes, through size-independent properties (the oracle cannot finish these
sizes in seconds): a fit's labels are the nearest centres of its own centres (idempotence), its centres are the
soft-threshold update of its own labels, assignment is deterministic, and a random sample of cells of the 100M
allocation equals the oracle's decision bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_c2_sparse_k_means_1m_x_15(engine, oracle):
    """configs[1]: 1M cells x 15 markers, K = 20, one penalty"""
    n, d, k, pen = 1_000_000, 15, 20, 8.0
    reg = pen * n * np.sqrt(d) / 1450.0
    with engine.synthetic(n, d, 12, 3.0, 20261020) as ds:
        fit = engine.sparse_k_means(ds, k, reg, True, seed_off_set=3)
        assert 22 <= fit["n_iter"] <= 1000                       # Clusterer.cpp:247: never fewer than 22 assignments
        lab = engine.allocate_points(ds, fit["c"], True)["i"]
        assert np.array_equal(lab, fit["i"])                     # converged: labels are a fixed point of the assignment
        again, cnt = engine.reevaluate_centers(ds, fit["i"], k, reg)
        np.testing.assert_allclose(again, fit["c"], rtol=1e-9, atol=1e-12)   # ... and the centres of the update
        assert cnt.sum() == n and np.array_equal(cnt, np.bincount(fit["i"], minlength=k))
        alive = np.abs(fit["c"]).sum(1) > 0
        assert 2 <= alive.sum() <= k and not np.isin(fit["i"], np.flatnonzero(~alive)).any()   # dead centres own no cell
        # a sample against the oracle (bit-exact labels)
        rows = np.random.default_rng(0).choice(n, 20000, replace=False)
        rows.sort()
        Xs = np.vstack([engine.read(ds, int(r), 1) for r in rows[:2000]])
        assert np.array_equal(oracle.allocate_points(Xs, fit["c"], True), fit["i"][rows[:2000]])
        assert abs(fit["n"] - engine.cluster_norm(ds, fit["c"], fit["i"], reg)) <= 1e-9 * abs(fit["n"])


def test_c3_lloyd_10m_x_40(engine, oracle):
    """configs[2] shape (the bench workload): 10M x 40, K = 30 — conservation and determinism of the fused iteration, and
    the oracle's word on it: a 20 000-cell sample of the labels (allocate_clusters, Clusterer.cpp:56-66) and the centre update
    of ALL 10M rows (reevaluate_centers, :99-107; one CPU sweep)"""
    n, d, k = 10_000_000, 40, 30
    reg = 1.0 * n * np.sqrt(d) / 1450.0
    with engine.synthetic(n, d, 24, 2.5, 20261021) as ds:
        mu0, rows = engine.initialize_mu(ds, k, seed=5)
        assert len(np.unique(rows)) == k
        mu_a, ch_a = engine.lloyd_iterations(ds, mu0, reg, True, -1, 6)
        mu_b, ch_b = engine.lloyd_iterations(ds, mu0, reg, True, -1, 6)
        assert ch_a == ch_b and np.array_equal(mu_a, mu_b)       # fixed reduction order: bit-reproducible
        lab = engine.allocate_points(ds, mu_a, True)["i"]
        full, cnt = engine.reevaluate_centers(ds, lab, k, reg)   # full scatter-add of 10M rows
        assert cnt.sum() == n
        # delta-accumulated sums (6 iterations) must equal a from-scratch update of the same labels
        mu_c, _ = engine.lloyd_iterations(ds, mu_a, reg, True, 20, 1)
        np.testing.assert_allclose(mu_c, full, rtol=1e-9, atol=1e-12)
        # the oracle on the same cells
        Xh = engine.read(ds, 0, n)
        rows = np.sort(np.random.default_rng(1).choice(n, 20000, replace=False))
        assert np.array_equal(oracle.allocate_points(Xh[rows], mu_a, True), lab[rows])          # bit-exact labels
        want, wcnt = oracle.reevaluate_centers(Xh, lab, k, reg)
        assert np.array_equal(wcnt, cnt)
        assert np.array_equal(want == 0, full == 0)                                               # identical zero pattern
        np.testing.assert_allclose(full, want, rtol=1e-9, atol=1e-12)                             # (north_star asks for 1e-5)
        np.testing.assert_allclose(mu_c, want, rtol=1e-9, atol=1e-12)                             # 6 delta iterations + 1, against a from-scratch CPU update


def test_c4_dallocate_100m_x_40(engine, oracle):
    """configs[3]: dAllocate of 100M cells x 40 markers onto fixed centres, 6 of 24 rows all-zero (no_zero)"""
    n, d, k = 100_000_000, 40, 24
    rng = np.random.default_rng(4)
    mu = rng.choice([-1.2, 0.0, 1.2], size=(k, d))
    mu[[2, 5, 9, 13, 17, 21]] = 0.0
    with engine.synthetic(n, d, 24, 2.5, 20261022) as ds:
        lab = engine.allocate_points(ds, mu, True)["i"]
        assert lab.shape == (n,) and lab.min() >= 0 and lab.max() < k
        assert not np.isin(lab, [2, 5, 9, 13, 17, 21]).any()
        assert np.array_equal(lab, engine.allocate_points(ds, mu, True)["i"])    # deterministic
        for r0 in (0, 49_999_937, n - 4096):
            Xs = engine.read(ds, r0, 4096)
            assert np.array_equal(oracle.allocate_points(Xs, mu, True), lab[r0:r0 + 4096])
        counts = np.bincount(lab, minlength=k)
        assert counts.sum() == n and (counts[[0, 1, 3, 4]] > 0).all()


def test_host_pipelines_over_several_chunks(engine):
    """The host <-> device pipelines (api.cu: staged_pipeline, copy_labels_out) at sizes that take several 32 MB / 4 MB chunks of the ring
    of pinned buffers: what comes back is exactly what went in — fp32(X) for an upload, the same labels through the chunked copy-out as
    through a plain read of a slice, and the device pre-scaling on the fp32 route and on the doubles route (one value that is not an fp32
    value, in the LAST chunk: the upload starts over) agree with the host restatement."""
    from depecher_b200 import rapi
    rng = np.random.default_rng(17)
    n, d = 1_300_000, 20                                          # 26M values: four chunks of 8M
    X = rng.normal(size=(n, d)) * 3.0 + rng.integers(-2, 3, size=(n, 1)) * 4.0
    ds = engine.dataset(X)
    for r0 in (0, 419_999, n - 1000):                             # rows whose values sit in different chunks of the column-major matrix
        got = engine.read(ds, r0, 1000)
        assert np.array_equal(got, X[r0:r0 + 1000].astype(np.float32).astype(np.float64))
    mu = rng.normal(size=(9, d)) * 3.0
    lab = engine.allocate_points(ds, mu, True)["i"]               # 1.3M labels = 5.2 MB: two chunks of the copy-out
    lab1 = engine.allocate_points(ds, mu, True, base=1)["i"]
    assert np.array_equal(lab + 1, lab1)
    with engine.gather(ds, np.arange(n - 3000, n)) as tail:       # the same cells through the small (unchunked) copy-out
        assert np.array_equal(engine.allocate_points(tail, mu, True)["i"], lab[n - 3000:])
    ds.close()
    Xf = X.astype(np.float32).astype(np.float64)
    for M in (Xf, X):                                             # fp32 route; doubles route (nearly every value fails the check: first chunk)
        want, winfo = rapi.depecheLogCenterSd(M)
        dss, info = engine.dataset_scaled(M)
        got = engine.read(dss, n - 2000, 2000)
        dss.close()
        assert np.allclose(info[1], winfo[1], rtol=0, atol=1e-12) and abs(info[2] - winfo[2]) <= 1e-12 * winfo[2]
        assert np.allclose(got, want[n - 2000:], rtol=3e-7, atol=1e-7)
    Xm = Xf.copy()
    Xm[n - 5, d - 1] += 1e-9                                      # the very last chunk holds the one value that is not an fp32 value
    want, winfo = rapi.depecheLogCenterSd(Xm)
    dss, info = engine.dataset_scaled(Xm)
    got = engine.read(dss, n - 2000, 2000)
    dss.close()
    assert np.allclose(info[1], winfo[1], rtol=0, atol=1e-12) and abs(info[2] - winfo[2]) <= 1e-12 * winfo[2]
    assert np.allclose(got, want[n - 2000:], rtol=3e-7, atol=1e-7)
