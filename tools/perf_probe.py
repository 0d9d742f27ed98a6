"""Kernel-level timing probe (run on a B200): fused pass in assignment and Lloyd mode over a few shapes.
    python tools/perf_probe.py [--n 10000000] [--shapes 40x30,15x20,...] [--iters 10]
Prints one line per shape/mode: mean launch ms, GB/s (algorithmic bytes), TFMA/s, cell*centre/s."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import depecher_b200 as dp  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=10_000_000)
    ap.add_argument("--shapes", default="40x30,40x8,15x20,50x30")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--modes", default="assign,lloyd")
    args = ap.parse_args()
    from depecher_b200 import _capi
    eng = dp.Engine(device=0, path=os.environ.get("DPR_LIB", _capi.LIB_PATH))   # DPR_LIB=<path>: another build of the library (A/B on one box)
    for shp in args.shapes.split(","):
        d, k = (int(v) for v in shp.split("x"))
        ds = eng.synthetic(args.n, d, 24, 2.5, 7)
        head = eng.read(ds, 0, 4096)
        mu = np.array(head[np.random.default_rng(3).choice(4096, k, replace=False)])
        for mode in args.modes.split(","):
            if mode == "assign":
                eng.assign_iterations(ds, mu, False, 3)
                eng.pass_timing(True)
                eng.assign_iterations(ds, mu, False, args.iters)
                bytes_cell = 4 * d + 4
            else:
                m2, _ = eng.lloyd_iterations(ds, mu, 0.0, True, 20, 3)
                eng.pass_timing(True)
                m2, ch = eng.lloyd_iterations(ds, m2, 0.0, True, 20, args.iters)
                bytes_cell = 4 * d + 8
            ms, cnt = eng.pass_timing(False)
            ms /= max(cnt, 1)
            print(f"d={d:3d} k={k:3d} {mode:6s}: {ms:8.4f} ms/launch  {args.n * bytes_cell / ms / 1e6:8.1f} GB/s  "
                  f"{args.n * k * d / ms / 1e9:7.2f} TFMA/s  {args.n * k / ms / 1e6:8.2f} G cell*centre/s", flush=True)
        ds.close()


if __name__ == "__main__":
    main()
