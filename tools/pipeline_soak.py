"""Soak test of the host <-> device pipelines (run on a B200): many uploads, scaled uploads and label copy-outs of varied sizes, each checked."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import depecher_b200 as dp  # noqa: E402

eng = dp.Engine(device=0)
rng = np.random.default_rng(3)
t0 = time.perf_counter()
bad = 0
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 60):
    n = int(rng.choice([1, 130, 70_000, 600_000, 2_100_000, 3_000_017]))
    d = int(rng.choice([1, 3, 15, 40]))
    X = rng.normal(size=(n, d)).astype(np.float32).astype(np.float64) if it % 3 else rng.normal(size=(n, d))
    ds = eng.dataset(X)
    r0 = max(0, n - 500)
    got = eng.read(ds, r0, n - r0)
    if not np.array_equal(got, X[r0:].astype(np.float32).astype(np.float64)):
        bad += 1
        print("upload mismatch", n, d, flush=True)
    mu = rng.normal(size=(5, d))
    lab = eng.allocate_points(ds, mu, False, base=1)["i"]
    d2 = ((X[r0:, None, :] - mu[None, :, :]) ** 2).sum(-1)
    margin = np.sort(d2, axis=1)
    sure = (margin[:, 1] - margin[:, 0]) > 1e-4 * (1 + margin[:, 1])
    if not np.array_equal(lab[r0:][sure], d2.argmin(1)[sure] + 1):
        bad += 1
        print("label mismatch", n, d, flush=True)
    ds.close()
    if n >= 1000 and it % 4 == 0:
        dss, info = eng.dataset_scaled(X, True, "mean", True)
        dss.close()
print(f"soak: {bad} mismatches, {time.perf_counter() - t0:.1f} s", flush=True)
