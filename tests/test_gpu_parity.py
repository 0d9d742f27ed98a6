"""GPU parity: the CUDA library through its C ABI against the CPU oracle on the same seeded inputs.

Bars: labels bit-exact on fp32-representable inputs; on arbitrary doubles only cells whose
oracle top-2 margin is below 1e-6 may differ; centres within 1e-5 relative (abs floor 1e-5 * data sd — in
practice they agree to ~1e-12 because the sums are accumulated in fp64); identical iteration counts.
"""
import numpy as np
import pytest

from oracle.oracle import RNG_PHILOX

pytestmark = pytest.mark.gpu


def blobs(rng, n, d, g, amp=3.0, f32=True):
    cent = rng.choice([-amp, 0.0, amp], size=(g, d))
    X = cent[rng.integers(0, g, n)] + rng.normal(size=(n, d))
    X /= X.std()
    return X.astype(np.float32).astype(np.float64) if f32 else X


def centres_close(a, b, sd=1.0):
    assert a.shape == b.shape
    assert np.array_equal(a == 0, b == 0), "zero pattern differs"
    np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-5 * sd)


# ---------------------------------------------------------------------------------------------- allocate_points
def test_allocate_points_golden(engine, ref_vectors):
    v = ref_vectors
    assert np.array_equal(engine.allocate_points(v["ap_X"], v["ap_mu"], False)["i"], v["ap_idx"])
    assert np.array_equal(engine.allocate_points(v["ap_X"], v["ap_mu"], True)["i"], v["ap_idx_nz"])
    assert np.array_equal(engine.allocate_points(v["ap_X"], v["ap_mu_one"], True)["i"], v["ap_idx_one_nz"])
    assert np.array_equal(engine.allocate_points(v["ap_X"], v["ap_mu_one"], False)["i"], v["ap_idx_one"])


@pytest.mark.parametrize("n,d,k", [(1, 3, 2), (127, 5, 2), (128, 14, 30), (129, 15, 20), (5000, 40, 30), (3333, 50, 30),
                                   (2000, 7, 33), (1500, 20, 90), (700, 100, 12), (64, 150, 2)])
def test_allocate_points_shapes(engine, oracle, n, d, k):
    rng = np.random.default_rng(n * 1000 + d * 10 + k)
    X = blobs(rng, n, d, min(k, 8))
    mu = blobs(rng, k, d, min(k, 8))
    if k > 4:
        mu[[1, k - 2]] = 0.0
    for nz in (False, True):
        got = engine.allocate_points(X, mu, nz)["i"]
        want = oracle.allocate_points(X, mu, nz)
        assert np.array_equal(got, want), f"{(got != want).sum()} of {n} labels differ (no_zero={nz})"


def test_allocate_points_exact_ties_and_duplicates(engine, oracle):
    """identical centres and cells equidistant from two centres: first minimum wins (minCoeff, Clusterer.cpp:64)"""
    rng = np.random.default_rng(5)
    d = 6
    mu = rng.integers(-3, 4, size=(8, d)).astype(np.float64)
    mu[5] = mu[2]                       # duplicate row: the lower index must win
    X = np.concatenate([mu[rng.integers(0, 8, 500)],                     # cells on centres
                        (mu[0] + mu[1])[None, :].repeat(50, 0) / 2.0,    # exactly between two centres
                        rng.integers(-3, 4, size=(1000, d)).astype(np.float64)])
    got = engine.allocate_points(X, mu, False)["i"]
    assert np.array_equal(got, oracle.allocate_points(X, mu, False))
    assert not (got == 5).any()


def test_allocate_points_arbitrary_doubles(engine, oracle):
    """inputs that are NOT fp32-representable: only documented near-ties (oracle margin < 1e-6) may differ"""
    rng = np.random.default_rng(11)
    X = blobs(rng, 20000, 15, 12, f32=False)
    mu = blobs(rng, 20, 15, 12, f32=False)
    got = engine.allocate_points(X, mu, False)["i"]
    want = oracle.allocate_points(X, mu, False)
    diff = np.flatnonzero(got != want)
    margin = oracle.assignment_margin(X, mu, False)
    assert (margin[diff] < 1e-6).all(), f"{len(diff)} differ, worst margin {margin[diff].max() if len(diff) else 0}"


def test_allocate_points_nan_and_errors(engine, oracle):
    import depecher_b200 as dp
    rng = np.random.default_rng(2)
    X = blobs(rng, 300, 4, 3)
    mu = blobs(rng, 5, 4, 3)
    X[7, 2] = np.nan
    assert np.array_equal(engine.allocate_points(X, mu, False)["i"], oracle.allocate_points(X, mu, False))
    with pytest.raises(dp.DepecherError):
        engine.allocate_points(X, np.zeros((0, 4)), False)


# ---------------------------------------------------------------------------------------------- centre update
@pytest.mark.parametrize("n,d,k,reg", [(3000, 14, 7, 0.0), (3000, 14, 7, 37.5), (10000, 40, 30, 120.0), (999, 33, 90, 5.0)])
def test_reevaluate_centers(engine, oracle, n, d, k, reg):
    rng = np.random.default_rng(n + d + k)
    X = blobs(rng, n, d, 6)
    idx = rng.integers(0, k, n).astype(np.int32)
    idx[idx == 3] = 4   # an empty cluster: exact zero row (SURVEY A-6/A-7)
    got, cnt = engine.reevaluate_centers(X, idx, k, reg)
    want, wcnt = oracle.reevaluate_centers(X, idx, k, reg)
    assert np.array_equal(cnt, wcnt)
    assert (got[3] == 0).all()
    centres_close(got, want)
    np.testing.assert_allclose(got, want, rtol=1e-11, atol=1e-13)


def test_golden_reevaluate(engine, ref_vectors):
    v = ref_vectors
    got0, _ = engine.reevaluate_centers(v["ap_X"], v["rc_idx"], 7, 0.0)
    got1, _ = engine.reevaluate_centers(v["ap_X"], v["rc_idx"], 7, 37.5)
    centres_close(got0, v["rc_mu_reg0"])
    centres_close(got1, v["rc_mu_reg"])


# ---------------------------------------------------------------------------------------------- Lloyd loop
@pytest.mark.parametrize("n,d,k,pen", [(3000, 14, 10, 0.0), (3000, 14, 10, 8.0), (10000, 14, 30, 16.0), (5000, 40, 30, 4.0),
                                       (500, 100, 90, 8.0)])
def test_find_centers_from_same_trajectory(engine, oracle, n, d, k, pen):
    """identical initial centres -> identical labels, iteration count and zero pattern; centres to 1e-5"""
    rng = np.random.default_rng(n + 7 * d + k)
    X = blobs(rng, n, d, 6, amp=2.5)
    reg = pen * n * np.sqrt(d) / 1450.0          # R/depecheAllData.R:15-16
    oracle.set_rng(RNG_PHILOX, 31)
    mu0, _ = oracle.initialize_mu(X, k, fit_id=2)
    want = oracle.find_centers_from(X, mu0, reg, True)
    got = engine.find_centers_from(X, mu0, reg, True)
    assert got["n_iter"] == want["n_iter"]
    assert np.array_equal(got["i"], want["i"])
    centres_close(got["c"], want["c"])
    assert abs(got["n"] - want["n"]) <= 1e-9 * abs(want["n"])


def test_kmeanspp_seeding_matches_oracle(engine, oracle):
    """identical rows and centres, from a fit of two 256-cell blocks to one of 274 (several block sums per scanning thread)"""
    rng = np.random.default_rng(3)
    for n, d, k in ((7000, 14, 30), (300, 5, 12), (65536, 9, 20), (70000, 14, 30)):
        X = blobs(rng, n, d, 8)
        oracle.set_rng(RNG_PHILOX, 77)
        for fit in (0, 5):
            want_mu, want_rows = oracle.initialize_mu(X, k, fit_id=fit)
            got_mu, got_rows = engine.initialize_mu(X, k, seed=77, fit_id=fit)
            assert np.array_equal(got_rows, want_rows), (n, d, k, fit)
            assert np.array_equal(got_mu, want_mu)


def test_bootstrap_rows_match_oracle(engine, oracle):
    oracle.set_rng(RNG_PHILOX, 123)
    assert np.array_equal(engine.bootstrap_rows(20000, 10000, seed=123, fit_id=9), oracle.bootstrap_rows(20000, 10000, 9))


def test_sparse_k_means_matches_oracle(engine, oracle):
    rng = np.random.default_rng(8)
    X = blobs(rng, 4000, 15, 7, amp=2.5)
    for pen in (0.0, 4.0, 22.6):
        reg = pen * 4000 * np.sqrt(15) / 1450.0
        oracle.set_rng(RNG_PHILOX, 5)
        want = oracle.sparse_k_means(X, 20, reg, True, fit_id=0)
        got = engine.sparse_k_means(X, 20, reg, True, seed_off_set=5)
        assert got["n_iter"] == want["n_iter"]
        assert np.array_equal(got["i"], want["i"])
        centres_close(got["c"], want["c"])
        assert abs(got["n"] - want["n"]) <= 1e-9 * abs(want["n"])


# ---------------------------------------------------------------------------------------------- ARI
def test_rand_index(engine, oracle, ref_vectors):
    import depecher_b200 as dp
    v = ref_vectors
    a, b = v["ri_a"], v["ri_b"]
    exact = engine.rand_index(a, b, 6)
    assert abs(exact - oracle.rand_index_exact(a, b, 6)) < 1e-12
    # the reference's Monte-Carlo value (golden, libc stream) scatters around it: 4 sigma of 10 000 Bernoulli pairs
    assert abs(exact - float(v["ri_val"])) < 0.03
    assert engine.rand_index(a + 1, a + 1, 100) == 1.0           # test_dAllocate.R:20
    assert np.isnan(engine.rand_index(np.zeros(40, int), np.zeros(40, int), 3))
    engine.set_ari_mode(dp.ARI_MONTE_CARLO)
    try:
        oracle.set_rng(RNG_PHILOX, 0)
        mc = engine.rand_index(a, b, 6)
        assert abs(mc - oracle.rand_index(a, b, 6, pair_id=0)) < 1e-9   # same Philox pairs -> same estimate
        assert engine.rand_index(a, a, 6) == 1.0
    finally:
        engine.set_ari_mode(dp.ARI_EXACT)
    with pytest.raises(dp.DepecherError):
        engine.rand_index(a + 10, b, 6)


def test_n_used_and_cluster_norm(engine, oracle, ref_vectors):
    v = ref_vectors
    assert engine.n_used_clusters(v["rc_idx"], 7) == int(v["nu_val"])
    got = engine.cluster_norm(v["ap_X"], v["rc_mu_reg"], v["rc_idx"], 37.5)
    assert abs(got - float(v["cn_val"])) <= 1e-10 * abs(float(v["cn_val"]))


# ---------------------------------------------------------------------------------------------- grid_search
def test_grid_search_matches_oracle(engine, oracle):
    rng = np.random.default_rng(21)
    X = blobs(rng, 6000, 14, 6, amp=2.5)
    regs = np.array([0.0, 2.0, 8.0, 32.0]) * 1500 * np.sqrt(14) / 1450.0
    oracle.set_rng(RNG_PHILOX, 99)
    oracle.set_dead_work(False)
    want = oracle.grid_search(X, [12], regs, 2, 1500)
    got = engine.grid_search(X, [12], regs, 2, 1500, seed_off_set=99)
    assert np.array_equal(got["n"], want["n"])
    for p, q in zip(got["c"], want["c"]):
        centres_close(p[0], q[0])
        centres_close(p[1], q[1])
    # exact ARI vs the oracle's Monte-Carlo estimate of the same pairs of partitions: 2 iterations averaged
    assert np.all(np.abs(got["d"] - want["d"]) < 0.04)


def test_allocate_many_matches_single(engine, oracle):
    rng = np.random.default_rng(4)
    X = blobs(rng, 5000, 14, 6)
    mus = [blobs(rng, 9, 14, 6) for _ in range(5)]
    mus[2][[3, 4]] = 0
    lab, ari = engine.allocate_many(X, mus, True)
    for s, mu in enumerate(mus):
        assert np.array_equal(lab[s], oracle.allocate_points(X, mu, True))
    for i in range(5):
        for j in range(5):
            assert abs(ari[i, j] - oracle.rand_index_exact(lab[j] + 1, lab[i] + 1, 9)) < 1e-12
    assert np.allclose(np.diag(ari), 1.0)


@pytest.mark.parametrize("n,d,k,m", [(70001, 40, 30, 7), (1000, 15, 20, 4), (129, 5, 3, 9)])
def test_bundled_centre_sets_share_one_pass(engine, oracle, n, d, k, m, monkeypatch):
    """m centre sets against the same cells: the tensor-core pass keeps up to 4 of them resident and multiplies every landed tile
    with each (Problem::bundle).  Labels must equal the oracle's for every set — with a set of dead rows, a trivial set (one
    live row: every label 0, Clusterer.cpp:50-53), duplicate rows — and must not depend on how many sets ride together."""
    rng = np.random.default_rng(n + m)
    X = blobs(rng, n, d, 6)
    mus = [blobs(rng, k, d, 6) for _ in range(m)]
    mus[1][1:] = 0.0                       # trivial under no_zero
    mus[2][[0, k - 1]] = 0.0               # dead rows
    mus[3][k - 1] = mus[3][0]              # duplicate: the lower index wins
    for nz in (True, False):
        want = [oracle.allocate_points(X, mu, nz) for mu in mus]
        for sets in ("4", "3", "1"):
            monkeypatch.setenv("DPR_TC_SETS", sets)
            lab, _ = engine.allocate_many(X, mus, nz, want_ari=False)
            for s in range(m):
                assert np.array_equal(lab[s], want[s]), f"set {s}, {sets} sets per pass, no_zero={nz}: {(lab[s] != want[s]).sum()} labels differ"


# ---------------------------------------------------------------------------------------------- edges of the supported range
@pytest.mark.parametrize("n,d,k", [(63, 1, 2), (64, 1, 3), (65, 2, 2), (1000, 3, 240), (4097, 64, 31), (300, 200, 7)])
def test_edge_shapes(engine, oracle, n, d, k):
    rng = np.random.default_rng(n + d + k)
    X = blobs(rng, n, d, 3)
    mu = blobs(rng, k, d, 3)
    assert np.array_equal(engine.allocate_points(X, mu, False)["i"], oracle.allocate_points(X, mu, False))
    idx = rng.integers(0, k, n).astype(np.int32)
    got, cnt = engine.reevaluate_centers(X, idx, k, 1.5)
    want, wcnt = oracle.reevaluate_centers(X, idx, k, 1.5)
    assert np.array_equal(cnt, wcnt)
    centres_close(got, want)


def test_unsupported_shapes_and_bad_arguments_fail_loudly(engine):
    import depecher_b200 as dp
    rng = np.random.default_rng(0)
    X = rng.normal(size=(200, 4))
    with pytest.raises(dp.DepecherError, match="not supported"):
        engine.allocate_points(X, rng.normal(size=(481, 4)), False)          # more than 480 centres
    with pytest.raises(dp.DepecherError, match="not supported"):
        engine.allocate_points(rng.normal(size=(10, 1000)), rng.normal(size=(3, 1000)), False)   # a slot would not fit in shared memory
    with pytest.raises(dp.DepecherError):
        engine.sparse_k_means(X, 1, 1.0, True)                               # k >= 2 (Clusterer.cpp:157-162 writes two rows)
    with pytest.raises(dp.DepecherError):
        engine.reevaluate_centers(X, np.full(200, 9, np.int32), 5, 0.0)      # label outside [0, k)
    with pytest.raises(dp.DepecherError):
        engine.grid_search(X, [1], [1.0], 1, 50)
    # all-identical cells: every k-means++ weight is zero after the first pick; the reference then picks row 0 again
    same = np.ones((100, 3))
    mu, rows = engine.initialize_mu(same, 4, seed=1)
    assert np.array_equal(mu, np.ones((4, 3))) and (rows[1:] == 0).all()


def test_many_centres_k300(engine, oracle):
    """depecheAllData fits K = 3k centres (R/depecheAllData.R:30-31): depeche(k = 100) on fewer than 10 000 rows needs K = 300 —
    beyond the tensor-core pass (K <= 128), on the FFMA pass with 10 chunks of centres (VERDICT r1 missing-6)"""
    rng = np.random.default_rng(300)
    n, d, k = 4000, 12, 300
    X = blobs(rng, n, d, 9, amp=2.5)
    mu = X[rng.choice(n, k, replace=False)].copy()
    mu[[7, 150, 299]] = 0.0
    for nz in (False, True):
        assert np.array_equal(engine.allocate_points(X, mu, nz)["i"], oracle.allocate_points(X, mu, nz))
    reg = 2.0 * n * np.sqrt(d) / 1450.0
    a = engine.find_centers_from(X, mu, reg, True)
    b = oracle.find_centers_from(X, mu, reg, True)
    assert a["n_iter"] == b["n_iter"] and np.array_equal(a["i"], b["i"])
    centres_close(a["c"], b["c"])
    oracle.set_rng(RNG_PHILOX, 5)
    want = oracle.sparse_k_means(X, k, reg, True, fit_id=0)
    got = engine.sparse_k_means(X, k, reg, True, seed_off_set=5)
    assert got["n_iter"] == want["n_iter"] and np.array_equal(got["i"], want["i"])
    assert abs(got["n"] - want["n"]) <= 1e-9 * abs(want["n"])
    assert engine.rand_index(got["i"] + 1, want["i"] + 1, k) == 1.0


def test_random_shapes_fuzz(engine, oracle):
    """60 random problems: sizes, scales, zero rows, duplicated centres, cells sitting on centres — labels bit-exact, and for a
    subset the whole Lloyd trajectory (iteration count, labels, zero pattern) equal to the oracle's."""
    rng = np.random.default_rng(20261020)
    for trial in range(60):
        n, d, k = int(rng.integers(1, 3000)), int(rng.integers(1, 70)), int(rng.integers(2, 70))
        scale = float(10.0 ** rng.uniform(-3, 3))
        g = int(rng.integers(1, 8))
        X = (blobs(rng, n, d, g) * scale).astype(np.float32).astype(np.float64)
        mu = (blobs(rng, k, d, g) * scale).astype(np.float32).astype(np.float64)
        if k > 3:
            mu[rng.integers(0, k, 2)] = 0.0
            mu[int(rng.integers(0, k))] = mu[int(rng.integers(0, k))]
        if n > 10:
            X[rng.integers(0, n, 3)] = mu[rng.integers(0, k, 3)]
        nz = bool(rng.integers(0, 2))
        got = engine.allocate_points(X, mu, nz)["i"]
        want = oracle.allocate_points(X, mu, nz)
        assert np.array_equal(got, want), f"trial {trial}: n={n} d={d} k={k} no_zero={nz}: {(got != want).sum()} labels differ"
        if trial % 6 == 0 and n >= 2 * k:
            Xu = X / scale
            reg = float(rng.choice([0.0, 2.0, 16.0])) * n * np.sqrt(d) / 1450.0
            mu0 = Xu[rng.choice(n, k, replace=False)]
            a = engine.find_centers_from(Xu, mu0, reg, True)
            b = oracle.find_centers_from(Xu, mu0, reg, True)
            assert a["n_iter"] == b["n_iter"] and np.array_equal(a["i"], b["i"]), f"trial {trial}: trajectories differ"
            centres_close(a["c"], b["c"])


# ---------------------------------------------------------------------------------------------- the two pass kernels
@pytest.mark.parametrize("n,d,k", [(5000, 40, 30), (3001, 15, 20), (2500, 50, 31), (4000, 8, 100), (1000, 3, 2), (6000, 64, 16)])
def test_tensor_core_pass_and_ffma_pass_agree(engine, oracle, n, d, k):
    """the tcgen05 pass (default where the shape fits tensor memory) and the FP32-pipe pass give the oracle's labels, the same
    Lloyd trajectory and centres equal to ~1e-12 (both accumulate in fp64, in different but fixed orders)"""
    rng = np.random.default_rng(n + 13 * d + k)
    X = blobs(rng, n, d, 6, amp=2.5)
    mu = X[rng.choice(n, k, replace=False)].copy()
    reg = 4.0 * n * np.sqrt(d) / 1450.0
    want = oracle.allocate_points(X, mu, True)
    out = {}
    try:
        for which in (0, 1):   # DPR_PASS_AUTO, DPR_PASS_FFMA
            engine.set_pass_kernel(which)
            assert np.array_equal(engine.allocate_points(X, mu, True)["i"], want)
            out[which] = engine.find_centers_from(X, mu, reg, True)
    finally:
        engine.set_pass_kernel(0)
    assert out[0]["n_iter"] == out[1]["n_iter"]
    assert np.array_equal(out[0]["i"], out[1]["i"])
    np.testing.assert_allclose(out[0]["c"], out[1]["c"], rtol=1e-10, atol=1e-12)
    b = oracle.find_centers_from(X, mu, reg, True)
    assert out[0]["n_iter"] == b["n_iter"] and np.array_equal(out[0]["i"], b["i"])
    centres_close(out[0]["c"], b["c"])


@pytest.mark.parametrize("scale", [2.0 ** -40, 2.0 ** -12, 3.0e-3, 1.0, 777.0, 2.0 ** 15, 2.0 ** 40])
def test_tensor_core_pass_scales(engine, oracle, scale):
    """cells and centres far outside fp16's range: the power-of-two scaling in front of the fp16 split keeps labels bit-exact;
    a few cells far from the bulk (tiles whose norm bound is tiny next to the matrix maximum) exercise the absolute error term"""
    rng = np.random.default_rng(int(np.log2(scale) * 7) + 1000)
    n, d, k = 4000, 24, 12
    X = blobs(rng, n, d, 5)
    X[:130] *= 1e-3          # a whole tile of tiny cells
    X[rng.integers(0, n, 4)] *= 60.0
    X = (X * scale).astype(np.float32).astype(np.float64)
    mu = X[rng.choice(n, k, replace=False)].copy()
    mu[0] = 0.0
    for nz in (False, True):
        got = engine.allocate_points(X, mu, nz)["i"]
        want = oracle.allocate_points(X, mu, nz)
        assert np.array_equal(got, want), f"scale {scale}: {(got != want).sum()} labels differ"


def test_recheck_counter(engine):
    rng = np.random.default_rng(5)
    X = blobs(rng, 3000, 10, 4)
    mu = X[:6].copy()
    mu[3] = mu[2]            # duplicate centre: every cell nearest to it is an exact tie -> decided by the fp64 path
    engine.recheck_count(True)
    engine.allocate_points(X, mu, False)
    first = engine.recheck_count(True)
    assert first > 0
    assert engine.recheck_count(True) == 0


def test_lloyd_is_bit_reproducible(engine):
    """the per-cluster sums are accumulated in fixed orders (group tables in shared memory, private per-warp tables updated by
    reductions of one lane per address, fixed combine order): repeated fits give bitwise identical centres and labels"""
    rng = np.random.default_rng(77)
    n, d, k = 300_000, 40, 30
    X = blobs(rng, n, d, 12, amp=2.0)
    mu0 = X[rng.choice(n, k, replace=False)].copy()
    reg = 2.0 * n * np.sqrt(d) / 1450.0
    runs = [engine.find_centers_from(X, mu0, reg, True) for _ in range(3)]
    for r in runs[1:]:
        assert r["n_iter"] == runs[0]["n_iter"]
        assert np.array_equal(r["i"], runs[0]["i"])
        assert np.array_equal(r["c"].view(np.uint64), runs[0]["c"].view(np.uint64)), "centres differ bitwise between runs"


# ---------------------------------------------------------------------------------------------- near ties at the error bound
def near_tie_cells(rng, mu, n, eps_lo=1e-9, eps_hi=1e-4):
    """cells on the bisector plane of two centres, pushed off it by +-eps * |cb - ca| along the line that joins them, eps
    log-uniform in [eps_lo, eps_hi]: the two best distances then differ by about 2 eps relative — from far inside the
    doubtful band of the fp32-grade scores (tau ~ 1e-6) to far outside it.  A small component inside the bisector plane
    keeps the cells distinct; everything is rounded to fp32 (which moves a cell by ~6e-8 relative: the margins end up
    spread densely around the bound rather than on a lattice)."""
    k, d = mu.shape
    a = rng.integers(0, k, n)
    b = (a + 1 + rng.integers(0, k - 1, n)) % k
    ca, cb = mu[a], mu[b]
    u = cb - ca
    un = np.linalg.norm(u, axis=1, keepdims=True)
    noise = rng.normal(size=(n, d)) * 0.05
    noise -= (noise * u).sum(1, keepdims=True) / un ** 2 * u          # stays in the bisector plane
    eps = np.exp(rng.uniform(np.log(eps_lo), np.log(eps_hi), n)) * rng.choice([-1.0, 1.0], n)
    X = (ca + cb) / 2 + noise + eps[:, None] * u
    return X.astype(np.float32).astype(np.float64)


@pytest.mark.parametrize("d,k", [(15, 20), (40, 30), (50, 30), (40, 8)])
def test_near_ties_at_the_error_bound(engine, oracle, d, k):
    """the tensor-core accumulation allowance inside tau is measured, not proved — so construct cells whose
    top-2 margin straddles tau and demand zero mislabels (allocate_clusters' first minimum of the fp64 sqrt distances,
    Clusterer.cpp:56-66), on the single-set pass, the bundled scoring pass and the FFMA pass; the exact path must have been
    exercised (dpr_recheck_count) without swallowing everything."""
    rng = np.random.default_rng(1000 * d + k)
    mu = (rng.normal(size=(k, d)) * 1.5).astype(np.float32).astype(np.float64)
    mu[rng.random((k, d)) < 0.3] = 0.0                                       # sparse centres, like fitted ones
    n = 200_000
    X = near_tie_cells(rng, mu, n)
    want = oracle.allocate_points(X, mu, False)
    margin = oracle.assignment_margin(X, mu, False)
    assert (margin < 1e-7).sum() > n // 20 and (margin > 1e-5).sum() > n // 20    # both sides of the bound are populated
    with engine.dataset(X) as ds:
        engine.recheck_count(reset=True)
        got = engine.allocate_points(ds, mu, False)["i"]
        redone = engine.recheck_count(reset=True)
        assert np.array_equal(got, want), f"{(got != want).sum()} mislabels, worst oracle margin {margin[got != want].max():.3e}"
        assert n // 50 < redone < n, f"{redone} of {n} cells took the exact path"
        # the same cells through a bundle of centre sets (the scoring instance of the pass): every member must agree too
        mus = [mu, mu[::-1].copy(), np.roll(mu, 3, axis=0), mu * 1.0]
        lab, _ = engine.allocate_many(ds, mus, False, want_labels=True, want_ari=False)
        for m, l in zip(mus, lab):
            w = oracle.allocate_points(X, m, False)
            assert np.array_equal(l, w), f"bundle member: {(l != w).sum()} mislabels"
        engine.set_pass_kernel(1)   # DPR_PASS_FFMA
        try:
            assert np.array_equal(engine.allocate_points(ds, mu, False)["i"], want)
        finally:
            engine.set_pass_kernel(0)


def test_near_ties_scaled_operands(engine, oracle):
    """the same construction with cells and centres far outside fp16's range (the pass scales both by powers of two) and with a
    handful of tiny cells among large ones (the absolute part of the split error, tc_abs1 / tc_abs2)"""
    rng = np.random.default_rng(77)
    d, k, n = 40, 30, 100_000
    for scale in (2.0 ** 20, 2.0 ** -20):
        mu = (rng.normal(size=(k, d)) * 1.5).astype(np.float32).astype(np.float64) * scale
        X = near_tie_cells(rng, mu, n)
        X[::1000] *= 2.0 ** -12
        got = engine.allocate_points(X, mu, False)["i"]
        want = oracle.allocate_points(X, mu, False)
        assert np.array_equal(got, want), f"scale {scale}: {(got != want).sum()} mislabels"
