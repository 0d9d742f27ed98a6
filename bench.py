#!/usr/bin/env python
"""bench.py — headline benchmark of the penalized-k-means hot path (BASELINE.json: cell-assignments/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE.json configs[2] — synthetic 10M cells x 40 markers, K = 30, the shape
the metric is quoted on; it fits one B200 (1.6 GB as fp32).  A step = ONE fused Lloyd iteration over the
whole matrix: nearest-centre assignment of every cell to the K centres + per-cluster sums/counts + the
soft-threshold centre update (allocate_clusters + reevaluate_centers, /root/reference/src/Clusterer.cpp:34-122).
  value     cells x K x steps / device time, X resident in HBM (CUDA events on the launching stream)
  e2e       the same metric through the drop-in C-ABI call dpr_sparse_k_means() with HOST double buffers:
            H2D of the R-layout matrix, k-means++ seeding, the whole Lloyd loop and the D2H of the labels
            are inside the timed region
  roofline  tc_pass_kernel (the tcgen05 pass over the cells): algorithmic bytes (4*D + 8 per cell: the fp32 row, the
            previous label read and the new label written) / mean launch duration, against the measured HBM copy bandwidth
  cpu_baseline  the reference's own Clusterer.cpp (oracle/_ref; the oracle port when absent) timed on a
            bounded sample on the host cores, one single-threaded engine per worker like the reference's
            SOCK workers (R/depeche.R:169-174)
N > 1: rows are sharded (every rank holds its own 10M x 40 shard: weak scaling) and each iteration
exchanges the K x (D+1) partial sums/counts between the ranks — inside the reduce + finalize launch, through peer-memory
mailboxes over NVLink (ncclAllReduce when they cannot be set up); value = all ranks' cells x K x steps / max time.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_CELLS, D, K, GROUPS, AMP, SEED = 10_000_000, 40, 30, 24, 2.5, 20261019 + 2
PENALTY = 1.0
METRIC = "cell_assignments_per_s"
UNIT = "cell*centre/s"


def reg_for(n: int, d: int, penalty: float) -> float:
    return penalty * n * np.sqrt(d) / 1450.0   # R/depecheAllData.R:15-16


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_engine():
    from oracle.oracle import Oracle, Ref
    if Ref.available():
        return Ref(), "reference"
    return Oracle(), "port"


def cpu_lloyd_rate(sample: np.ndarray, mu: np.ndarray, reg: float, iters: int, workers: int):
    """`workers` independent single-threaded engines, each running `iters` Lloyd iterations on its own copy of
    `sample` (ctypes releases the GIL).  Returns (cell*centre/s over all workers, seconds, kind)."""
    eng, kind = cpu_engine()
    secs = [0.0] * workers

    def run(w):
        secs[w] = eng.time_lloyd(sample, mu, reg, True, iters)

    th = [threading.Thread(target=run, args=(w,)) for w in range(workers)]
    t0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    wall = time.perf_counter() - t0
    work = float(sample.shape[0]) * mu.shape[0] * iters * workers
    return work / max(max(secs), 1e-9), wall, kind


def host_workers() -> int:
    n = os.cpu_count() or 1
    return max(1, min(10, int(n * 0.875)))     # R/depeche.R:169-174


def synth_host(n: int, d: int, groups: int, amp: float, seed: int) -> np.ndarray:
    """host-side sample of the same mixture shape (used only when no GPU is present, i.e. the reference arm)"""
    rng = np.random.default_rng(seed)
    cent = rng.choice([-amp, 0.0, amp], size=(groups, d))
    for g in range(groups):
        cent[g, rng.choice(d, size=(g * d + 2 * groups - 1) // (2 * groups), replace=False)] = 0.0
    X = cent[rng.integers(0, groups, n)] + rng.normal(size=(n, d))
    X /= X.std()
    return X.astype(np.float32).astype(np.float64)


def run_reference(args, rank: int, world: int) -> None:
    if rank != 0:
        return
    workers = host_workers()
    rows = 60_000
    X = synth_host(rows, D, GROUPS, AMP, SEED)
    mu = X[np.random.default_rng(1).choice(rows, K, replace=False)]
    reg = reg_for(rows, D, PENALTY)
    eng, kind = cpu_engine()
    for _ in range(min(args.warmup, 1)):
        cpu_lloyd_rate(X, mu, reg, 1, workers)
    t0 = time.perf_counter()
    rate, _, kind = cpu_lloyd_rate(X, mu, reg, args.steps, workers)
    wall = time.perf_counter() - t0
    sample = f"{workers} single-threaded engines x {rows} cells x {D} markers, K={K}, {args.steps} Lloyd iterations each"
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": wall * 1e3 / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": f"synthetic {N_CELLS} cells x {D} markers, K={K}, one fused Lloyd iteration per step (bounded CPU sample)"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": workers, "kind": kind, "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ GPU arm
def run_ours(args, rank: int, world: int, local_rank: int) -> None:
    import torch
    import torch.distributed as dist

    import depecher_b200 as dp

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    eng = dp.Engine(device=local_rank)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        from depecher_b200 import shard
        shard.init_comm(eng, dist, device=torch.device("cuda", local_rank))

    n_local = N_CELLS
    ds = eng.synthetic(n_local, D, GROUPS, AMP, SEED, row0=rank * n_local, scale=0.0)
    reg = reg_for(n_local * world, D, PENALTY)
    # initial centres: k-means++ on rank 0's shard (initialize_mu), identical on every rank; untimed like the data upload
    mu0, _ = eng.initialize_mu(ds, K, seed=SEED, fit_id=0)
    if world > 1:
        t = torch.from_numpy(np.ascontiguousarray(mu0)).cuda()
        dist.broadcast(t, src=0)
        mu0 = t.cpu().numpy()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # A step = one fused Lloyd iteration of a FRESH fit (iterations 0, 1, 2, ... of find_centers: all labels change in
    # iteration 0, the penalty ramps up over the first 20, dying clusters drop out) - not a converged steady state.
    sampler = ClockSampler(local_rank)   # samples from the warm-up on, so that a short timed region still has clock readings under load
    sampler.start()
    eng.lloyd_iterations(ds, mu0, reg, True, -1, max(args.warmup, 3))             # warm-up (>= 3 steps)
    for _ in range(3):                                                            # keep the GPU busy ~100 ms for the sampler
        eng.lloyd_iterations(ds, mu0, reg, True, -1, 30)
    eng.pass_timing(True)
    launches0 = eng.kernel_launches(reset=True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    mu, changed = eng.lloyd_iterations(ds, mu0, reg, True, -1, args.steps)
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    launches = eng.kernel_launches()
    pass_ms, pass_cnt = eng.pass_timing(False)
    k_active1 = int((np.abs(mu).sum(1) > 0).sum())
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    k_eff = k_active1          # clusters only ever die during a fit: the count at the end is the smallest of the run
    value = float(n_local) * world * k_eff * args.steps / (ms * 1e-3)
    # the same loop continued in its steady state (ramp complete, a handful of labels still moving)
    steady = None
    if rank == 0 and world == 1:
        eng.pass_timing(True)
        t0s, t1s = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0s.record()
        mu_s, _ = eng.lloyd_iterations(ds, mu, reg, True, 20, 10)
        t1s.record()
        torch.cuda.synchronize()
        sp_ms, sp_cnt = eng.pass_timing(False)
        k_s = int((np.abs(mu_s).sum(1) > 0).sum())
        steady = {"ms_per_step": t0s.elapsed_time(t1s) / 10, "fused_pass_ms": sp_ms / max(sp_cnt, 1), "k_active": k_s,
                  "value": float(n_local) * k_s * 10 / (t0s.elapsed_time(t1s) * 1e-3)}

    line = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak, which = (peaks.get("hbm_gbs"), "measured (MEASURED_PEAKS.json)") if peaks.get("hbm_gbs") else (6650.0, "fallback (B200_PROFILING.md)")
        bytes_per_launch = float(n_local) * (4 * D + 8)
        launch_ms = pass_ms / max(pass_cnt, 1)
        achieved = bytes_per_launch / (launch_ms * 1e-3) / 1e9
        fma_per_launch = float(n_local) * k_eff * D
        traffic = None   # dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed `ncu --set full` capture
        try:
            for ln in open(os.path.join(ROOT, "profiles", "r01_tc_pass_bench_ncu.csv")):
                f = ln.strip().split(",")
                if f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[f[1]]
                    traffic = (traffic or 0.0) + float(np.mean([float(v) for v in f[2:]])) * scale
        except Exception:
            traffic = None
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                    "peak_source": which, "kernel": "tc_pass_kernel", "launch_ms": launch_ms, "launches_timed": pass_cnt,
                    "fma_per_s": fma_per_launch / (launch_ms * 1e-3),
                    "note": "distance products on tcgen05 (fp16-split operands, fp32 accumulate, exact fp64 re-decision of near ties); the launches timed "
                            "include the full-sums first iteration and the iterations in which 1-6 % of the cells change cluster"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"synthetic {n_local} cells x {D} markers per GPU, K={K} ({k_eff} still active at the end), penalty {PENALTY}, "
                                   f"one fused Lloyd iteration per step: iterations 0..{args.steps - 1} of a fresh fit from k-means++ centres",
                       "cells_per_gpu": n_local, "markers": D, "k": K, "k_active": k_eff, "parallelism": f"row-shard x{world}" if world > 1 else "single GPU",
                       "l2": "inputs (1.6 GB per GPU) exceed the 126 MB L2", "labels_changed_last_step": int(changed)},
            "clocks": clocks, "gpu_launches": int(launches), "roofline": roofline, "steady_state": steady,
        }
    # ---- e2e through the C ABI with host buffers, and the CPU baseline: rank 0 at N=1 only (bounded)
    if rank == 0 and world == 1:
        Xh = np.asfortranarray(eng.read(ds, 0, n_local))            # the R-side matrix: column-major doubles on the host
        eng.synchronize()
        # warm-up of the host path (pinned staging buffers, thread start-up) on a 1M-row slice, then the timed call
        eng.sparse_k_means(np.asfortranarray(Xh[:1_000_000]), K, reg_for(1_000_000, D, PENALTY), True, seed_off_set=1)
        t0 = time.perf_counter()
        res = eng.sparse_k_means(Xh, K, reg, True, seed_off_set=1)
        dt = time.perf_counter() - t0
        k_e2e = int((np.abs(res["c"]).sum(1) > 0).sum())
        line["e2e"] = {"value": float(n_local) * K * res["n_iter"] / dt, "unit": UNIT, "h2d_bytes_per_step": int(Xh.nbytes),
                       "d2h_bytes_per_step": int(res["i"].nbytes + res["c"].nbytes), "seconds": dt, "lloyd_iterations": res["n_iter"],
                       "k_active_final": k_e2e, "call": "dpr_sparse_k_means(host double*, ...) - H2D, k-means++ seeding, Lloyd loop, D2H inside"}
        # the other half of the metric: depeche() itself on the 10M x 40 host matrix - upload, device-side pre-scaling
        # (kurtosis test, histogram-peak centring, global sd), full default penalty sweep, solution selection, final allocation
        try:
            from depecher_b200 import rapi
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                rapi.depeche(np.asfortranarray(Xh[:200_000]), nCores=10, engine=eng, seed=SEED)   # warm-up (kernel modules, staging buffers) on a slice, like the e2e call above
                t0 = time.perf_counter()
                dep = rapi.depeche(Xh, nCores=10, engine=eng, seed=SEED + 1)
                first = time.perf_counter() - t0      # includes growing the device memory pool by the 3.2 GB of raw doubles + the resident matrix
                t0 = time.perf_counter()
                dep = rapi.depeche(Xh, nCores=10, engine=eng, seed=SEED + 1)
            line["depeche_call_10Mx40"] = {"seconds": time.perf_counter() - t0, "first_call_seconds": first, "bestPenalty": dep["penaltyOptList"]["bestPenalty"],
                                           "clusters": int(dep["clusterCenters"].shape[0]),
                                           "workerIterations": dep["penaltyOptList"]["workerIterations"],
                                           "call": "depecher_b200.rapi.depeche(host double matrix, nCores=10): everything inside"}
        except Exception as exc:
            line["depeche_call_10Mx40"] = {"error": repr(exc)}
        workers = host_workers()
        rows = 60_000
        sample = np.array(Xh[:rows])
        del Xh
        mu_cpu = sample[np.random.default_rng(1).choice(rows, K, replace=False)]
        rate, wall, kind = cpu_lloyd_rate(sample, mu_cpu, reg_for(rows, D, PENALTY), 3, workers)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": workers, "kind": kind,
                                "sample": f"{workers} single-threaded engines x {rows} cells x {D} markers, K={K}, 3 Lloyd iterations each ({wall:.1f} s)"}
        # the other half of the metric: depeche() on 10M x 40 with the full default penalty sweep (X resident, pre-scaling excluded)
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import depeche_c3
            ds.close()
            dres = depeche_c3.run(eng, N_CELLS, D, K, 10, GROUPS, AMP, SEED + 1, verbose=False)
            line["depeche_10Mx40"] = {kk: dres[kk] for kk in ("total_s", "penalty_opt_s", "selection_s", "final_allocate_s", "workerIterations",
                                                              "bestPenalty", "clusters", "kernel_launches")}
        except Exception as exc:   # the headline line must still print
            line["depeche_10Mx40"] = {"error": repr(exc)}
    if world > 1:
        # e2e at N > 1: every rank hands its row shard to the library as HOST doubles (R layout) and all ranks run one row-sharded fit
        # from the common initial centres: upload, Lloyd loop with the NCCL allreduce per iteration, labels back - all inside
        Xh = np.asfortranarray(eng.read(ds, 0, n_local))
        eng.synchronize()
        eng.find_centers_from(np.asfortranarray(Xh[:1_000_000]), mu0, reg, True, max_iter=3)      # warm-up of the host path (collective: every rank)
        barrier()
        t0 = time.perf_counter()
        res = eng.find_centers_from(Xh, mu0, reg, True)
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        if rank == 0:
            line["e2e"] = {"value": float(n_local) * world * K * res["n_iter"] / dt, "unit": UNIT, "h2d_bytes_per_step": int(Xh.nbytes) * world,
                           "d2h_bytes_per_step": int(res["i"].nbytes + res["c"].nbytes) * world, "seconds": dt, "lloyd_iterations": res["n_iter"],
                           "call": "dpr_dataset_create(host double*) + dpr_dataset_find_centers_from(...) on every rank's row shard - H2D, row-sharded "
                                   "Lloyd loop (per-iteration exchange of the K x (D+1) tables through peer memory, ncclAllReduce as fallback), D2H inside; max over ranks"}
        del Xh
    if rank == 0:
        print(json.dumps(line), flush=True)
    ds.close()
    if world > 1:
        eng.comm_destroy()
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
