"""Upload of a 10M x 40 host double matrix (dpr_dataset_create): wall time of repeated calls (run on a B200).
DPR_COPY_THREADS=<n> overrides the number of host threads that narrow the doubles."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import depecher_b200 as dp  # noqa: E402

eng = dp.Engine(device=0)
n, d = 10_000_000, 40
ds = eng.synthetic(n, d, 24, 2.5, 20261021, scale=0.0)
Xh = np.asfortranarray(eng.read(ds, 0, n))
ds.close()
ts = []
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 9):
    t0 = time.perf_counter()
    h = eng.dataset(Xh)
    eng.synchronize()
    ts.append(time.perf_counter() - t0)
    h.close()
ts = np.array(ts[1:])
print(f"threads {os.environ.get('DPR_COPY_THREADS', 'default')}: upload median {np.median(ts)*1e3:.1f} ms, best {ts.min()*1e3:.1f}, all {np.round(ts*1e3,1).tolist()}", flush=True)
