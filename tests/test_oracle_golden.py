"""The oracle restatement against (1) golden vectors produced by the reference's own engine
(tests/golden/ref_vectors.npz, made by make_golden.py from oracle/_ref) and (2) the live _ref
build where it is present.  Under the same libc rand() stream everything must be bit-identical."""
import numpy as np
import pytest

from oracle.oracle import RNG_LIBC, RNG_PHILOX, Oracle, Ref


@pytest.fixture()
def libc(oracle):
    oracle.set_rng(RNG_LIBC)
    yield oracle
    oracle.set_rng(RNG_PHILOX, 0)


def test_allocate_points_golden(oracle, ref_vectors):
    v = ref_vectors
    assert np.array_equal(oracle.allocate_points(v["ap_X"], v["ap_mu"], False), v["ap_idx"])
    assert np.array_equal(oracle.allocate_points(v["ap_X"], v["ap_mu"], True), v["ap_idx_nz"])
    # dead rows are never chosen with no_zero, and act as one origin cluster (lowest index) without
    assert not np.isin(v["ap_idx_nz"], [2, 5]).any()
    assert not (v["ap_idx"] == 5).any()
    # one active row: all labels 0 even though the active row is 4 (Clusterer.cpp:50-53)
    assert np.array_equal(oracle.allocate_points(v["ap_X"], v["ap_mu_one"], True), v["ap_idx_one_nz"])
    assert (v["ap_idx_one_nz"] == 0).all()
    assert np.array_equal(oracle.allocate_points(v["ap_X"], v["ap_mu_one"], False), v["ap_idx_one"])
    assert set(np.unique(v["ap_idx_one"])) <= {0, 4}
    assert np.array_equal(oracle.allocate_points(v["apd_X"], v["apd_mu"], False), v["apd_idx"])


def test_reevaluate_centers_golden(oracle, ref_vectors):
    v = ref_vectors
    mu0, cnt = oracle.reevaluate_centers(v["ap_X"], v["rc_idx"], 7, 0.0)
    mu1, _ = oracle.reevaluate_centers(v["ap_X"], v["rc_idx"], 7, 37.5)
    assert np.array_equal(mu0, v["rc_mu_reg0"]) and np.array_equal(mu1, v["rc_mu_reg"])
    assert cnt[3] == 0 and (mu0[3] == 0).all() and (mu1[3] == 0).all()  # empty cluster -> exact zero row
    assert (mu1 == 0).sum() > (mu0 == 0).sum()  # the L1 threshold zeroes coefficients


def test_norm_and_used_golden(oracle, ref_vectors):
    v = ref_vectors
    assert oracle.cluster_norm(v["ap_X"], v["rc_mu_reg"], v["rc_idx"], 37.5) == float(v["cn_val"])
    assert oracle.n_used_clusters(v["rc_idx"], 7) == int(v["nu_val"]) == 6


def test_seeded_stages_golden(libc, ref_vectors):
    v = ref_vectors
    T0 = int(v["T0"])
    libc.srand(T0 + 5)
    mu, rows = libc.initialize_mu(v["ap_X"], 8)
    assert np.array_equal(mu, v["im_mu"]) and np.array_equal(v["ap_X"][rows], mu)
    libc.srand(T0 + 6)
    rows = libc.bootstrap_rows(3000, 500)
    assert np.array_equal(v["ap_X"][rows], v["bs_X"])


def test_sparse_k_means_golden(libc, ref_vectors):
    v = ref_vectors
    for tag, reg in zip(("p0", "p1", "p2"), v["sk_regs"]):
        libc.srand(int(v["T0"]) + 7)
        s = libc.sparse_k_means(v["ap_X"], 10, float(reg), True)
        assert np.array_equal(s["i"], v[f"sk_{tag}_idx"])
        assert np.array_equal(s["c"], v[f"sk_{tag}_c"])
        assert s["n"] == float(v[f"sk_{tag}_n"])
        assert s["n_iter"] >= 22  # at least 22 assignments (Clusterer.cpp:247)


def test_grid_search_golden(libc, ref_vectors):
    v = ref_vectors
    libc.srand(int(v["T0"]) + 8)
    g = libc.grid_search(v["ap_X"], [10], v["gs_regs"], 2, 800)
    assert np.array_equal(g["d"], v["gs_d"]) and np.array_equal(g["n"], v["gs_n"])
    assert np.array_equal(np.stack([np.stack(p) for p in g["c"]]), v["gs_c"])


def test_rand_index_golden(libc, ref_vectors):
    v = ref_vectors
    libc.srand(int(v["T0"]) + 9)
    assert libc.rand_index(v["ri_a"], v["ri_b"], 6) == float(v["ri_val"])
    libc.srand(int(v["T0"]) + 9)
    assert libc.rand_index(v["ri_a"] + 1, v["ri_a"] + 1, 100) == float(v["ri_same"]) == 1.0
    # the Monte-Carlo value scatters around the contingency-table value
    exact = libc.rand_index_exact(v["ri_a"], v["ri_b"], 6)
    assert abs(exact - float(v["ri_val"])) < 0.03


def test_exact_ari_matches_sklearn(oracle):
    from sklearn.metrics import adjusted_rand_score
    rng = np.random.default_rng(3)
    a = rng.integers(0, 9, 4000)
    b = np.where(rng.random(4000) < 0.6, a, rng.integers(0, 9, 4000))
    assert abs(oracle.rand_index_exact(a, b, 9) - adjusted_rand_score(a, b)) < 1e-12
    assert oracle.rand_index_exact(a, a, 9) == 1.0
    assert np.isnan(oracle.rand_index_exact(np.zeros(50, int), np.zeros(50, int), 3))  # both trivial: 0/0


@pytest.mark.skipif(not Ref.available(), reason="oracle/_ref not built here")
def test_restatement_equals_reference_live(libc):
    """Fresh random problems: the reference's Clusterer.cpp and the restatement, same rand() stream."""
    ref = Ref()
    rng = np.random.default_rng(77)
    for trial in range(3):
        n, d, k = int(rng.integers(300, 1500)), int(rng.integers(2, 20)), int(rng.integers(2, 15))
        cent = rng.normal(size=(4, d)) * 3
        X = cent[rng.integers(0, 4, n)] + rng.normal(size=(n, d))
        X /= X.std()
        reg = float(rng.choice([0.0, 0.02, 0.1]) * n)
        T = 1000 + trial
        ref.set_time(T)
        a = ref.sparse_k_means(X, k, reg, True, seed_off_set=trial)
        libc.srand(T + trial)
        b = libc.sparse_k_means(X, k, reg, True)
        assert np.array_equal(a["i"], b["i"]) and np.array_equal(a["c"], b["c"]) and a["n"] == b["n"]
        ref.set_time(T)
        ga = ref.grid_search(X, [k], [reg, reg * 3 + 1], 1, n // 2, seed_off_set=trial)
        libc.srand(T + trial)
        gb = libc.grid_search(X, [k], [reg, reg * 3 + 1], 1, n // 2)
        assert np.array_equal(ga["d"], gb["d"], equal_nan=True) and np.array_equal(ga["n"], gb["n"])


def test_philox_streams_are_addressed(oracle):
    """Philox mode: a fit is a pure function of (seed, fit id) — order of calls does not matter."""
    rng = np.random.default_rng(5)
    X = rng.normal(size=(500, 6))
    oracle.set_rng(RNG_PHILOX, 99)
    a1 = oracle.sparse_k_means(X, 6, 3.0, True, fit_id=4)
    oracle.sparse_k_means(X, 6, 3.0, True, fit_id=1)
    a2 = oracle.sparse_k_means(X, 6, 3.0, True, fit_id=4)
    assert np.array_equal(a1["i"], a2["i"]) and np.array_equal(a1["c"], a2["c"])
    oracle.set_rng(RNG_PHILOX, 100)
    b = oracle.sparse_k_means(X, 6, 3.0, True, fit_id=4)
    assert not np.array_equal(a1["c"], b["c"])
    rows = oracle.bootstrap_rows(500, 1000, 3)
    assert rows.min() >= 0 and rows.max() < 500 and len(np.unique(rows)) > 350
    oracle.set_rng(RNG_PHILOX, 0)
