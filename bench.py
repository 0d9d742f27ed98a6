#!/usr/bin/env python
"""bench.py — headline benchmark of the penalized-k-means hot path (BASELINE.json: cell-assignments/s + depeche() wall-time).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE.json configs[2] — synthetic 10M cells x 40 markers, K = 30, the shape the metric is
quoted on; it fits one B200 (1.6 GB as fp32).  A step = ONE fused Lloyd iteration over the whole matrix: nearest-centre
assignment of every cell to the K centres + per-cluster sums/counts + the soft-threshold centre update
(allocate_clusters + reevaluate_centers, /root/reference/src/Clusterer.cpp:34-122).
  value     cells x K_active x steps / device time, X resident in HBM (CUDA events on the launching stream); the K steps are
            timed `REPEATS` times, the line carries the median repeat (and all of them under `repeats`)
  e2e       the same metric (same K_active convention) through the drop-in C-ABI call dpr_sparse_k_means() with HOST double
            buffers: H2D of the R-layout matrix, k-means++ seeding, the whole Lloyd loop and the D2H of the labels are timed
            (one untimed call of the same size first, then three timed calls: the median, all three under `seconds_all`)
  roofline  tc_pass_kernel (the tcgen05 pass over the cells): algorithmic bytes (4*D + 8 per cell: the fp32 row, the previous
            label read and the new label written) / mean launch duration, against the measured HBM copy bandwidth
  cpu_baseline  the reference's own Clusterer.cpp (oracle/_ref; the oracle port when absent) on a 1M-row slice of the workload,
            one single-threaded engine per worker like the reference's SOCK workers (R/depeche.R:169-174)
  configs   the other BASELINE configs at N = 1: [0] depeche(testData) GPU vs the reference's CPU engine, [1] 1M x 15, [3] 100M x 40
            dAllocate, [4] 2M x 50 — each with its kernel time, algorithmic GB/s and fraction of the HBM roofline
  depeche_10Mx40  the other half of the metric: the depeche() hot path on the workload with the full default penalty sweep
N > 1: rows are sharded (every rank holds its own 10M x 40 shard: weak scaling) and each iteration exchanges the K x (D+1)
partial sums/counts between the ranks inside the reduce + finalize launch, through peer-memory mailboxes over NVLink
(ncclAllReduce when they cannot be set up); value = all ranks' cells x K x steps / max time.  The line then also carries
`parity` (row-sharded fit == NCCL path == one GPU holding all rows), `strong_scaling` (10M cells in total) and
`depeche_10Mx40` with the (worker x penalty) grid of every set split over the ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_CELLS, D, K, GROUPS, AMP, SEED = 10_000_000, 40, 30, 24, 2.5, 20261019 + 2
PENALTY = 1.0
REPEATS = 5
CPU_ROWS, CPU_ITERS = 1_000_000, 10
METRIC = "cell_assignments_per_s"
UNIT = "cell*centre/s"
TRAFFIC_CSV = os.path.join("profiles", "r02_tc_pass_bench_ncu.csv")


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


def cpu_baseline_block(sample: np.ndarray, iters: int) -> dict:
    """the CPU figure both arms report: the reference's engine, `host_workers()` single-threaded engines, `iters` Lloyd iterations
    each on the same 1M-row slice of the workload, K = 30 initial centres drawn from the slice"""
    workers = host_workers()
    rows = sample.shape[0]
    mu = sample[np.random.default_rng(1).choice(rows, K, replace=False)]
    rate, wall, kind = cpu_lloyd_rate(sample, mu, reg_for(rows, D, PENALTY), iters, workers)
    return {"value": rate, "unit": UNIT, "cores": workers, "kind": kind, "seconds": wall,
            "sample": f"{workers} single-threaded engines x {rows} cells x {D} markers, K={K}, {iters} Lloyd iterations each"}


def run_reference(args, rank: int, world: int) -> None:
    if rank != 0:
        return
    X = synth_host(CPU_ROWS, D, GROUPS, AMP, SEED)
    mu = X[np.random.default_rng(1).choice(CPU_ROWS, K, replace=False)]
    workers = host_workers()
    if args.warmup > 0:
        cpu_lloyd_rate(X[:50_000].copy(), mu, reg_for(50_000, D, PENALTY), 1, workers)
    t0 = time.perf_counter()
    cb = cpu_baseline_block(X, args.steps)
    wall = time.perf_counter() - t0
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": wall * 1e3 / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": f"synthetic {N_CELLS} cells x {D} markers, K={K}, one fused Lloyd iteration per step (bounded CPU sample: a {CPU_ROWS}-row slice per worker)"},
        "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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


def hbm_peak():
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    return (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if peaks.get("hbm_gbs") else (6650.0, "fallback (B200_PROFILING.md)")


def committed_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the pass kernel, from the committed `ncu --set full` capture of this
    command (NOT measured in this run: a run under a profiler is never a bench run)"""
    traffic = None
    try:
        for ln in open(os.path.join(ROOT, TRAFFIC_CSV)):
            f = ln.strip().split(",")
            if f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[f[1]]
                traffic = (traffic or 0.0) + float(np.mean([float(v) for v in f[2:]])) * scale
    except Exception:
        traffic = None
    return traffic


# ------------------------------------------------------------------------------------------------ pieces of the GPU arm
def time_device(torch, fn, repeats=1):
    """device time of fn() in ms (CUDA events on the current stream, a synchronize on both sides), best and median of `repeats`"""
    ms = []
    for _ in range(repeats):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.median(ms)), float(np.min(ms))


def other_configs(eng, torch, peak: float) -> dict:
    """BASELINE configs[1], [3], [4] at one GPU: kernel time of the pass, algorithmic GB/s (4*D + 4 B per cell for an assignment,
    4*D + 8 for a Lloyd pass) and the fraction of the measured HBM copy bandwidth; plus the call a user makes."""
    out = {}

    def assign_block(ds, mu, nz, n, d, iters):
        eng.assign_iterations(ds, mu, nz, 2)
        eng.pass_timing(True)
        eng.assign_iterations(ds, mu, nz, iters)
        ms, cnt = eng.pass_timing(False)
        ms /= max(cnt, 1)
        gbs = n * (4.0 * d + 4.0) / (ms * 1e-3) / 1e9
        return {"assign_pass_ms": ms, "assign_GBps": gbs, "assign_frac_of_hbm": gbs / peak, "bytes_per_cell": 4 * d + 4}

    # ---- configs[1]: 1M x 15, K = 20, one penalty
    try:
        n, d, k = 1_000_000, 15, 20
        reg = reg_for(n, d, 8.0)
        with eng.synthetic(n, d, 12, 3.0, 20261020) as ds:
            eng.sparse_k_means(ds, k, reg, True, seed_off_set=3, want_labels=False)          # warm-up
            t0 = time.perf_counter()
            fit = eng.sparse_k_means(ds, k, reg, True, seed_off_set=3, want_labels=False)
            eng.synchronize()
            dt = time.perf_counter() - t0
            k_act = int((np.abs(fit["c"]).sum(1) > 0).sum())
            blk = {"workload": "synthetic 1M cells x 15 markers, K=20, penalty 8: sparse_k_means (k-means++ + Lloyd loop), X resident", "seconds": dt,
                   "lloyd_iterations": fit["n_iter"], "k_active": k_act, "value": float(n) * k_act * fit["n_iter"] / dt, "unit": UNIT}
            blk.update(assign_block(ds, fit["c"], True, n, d, 20))
            Xh = np.asfortranarray(eng.read(ds, 0, n))
            eng.sparse_k_means(Xh, k, reg, True, seed_off_set=3)
            t0 = time.perf_counter()
            r = eng.sparse_k_means(Xh, k, reg, True, seed_off_set=3)
            blk["e2e_seconds"] = time.perf_counter() - t0
            blk["e2e_value"] = float(n) * k_act * r["n_iter"] / blk["e2e_seconds"]
        out["1Mx15"] = blk
    except Exception as exc:   # the headline line must still print
        out["1Mx15"] = {"error": repr(exc)}
    # ---- configs[4]: 2M x 50 PCs, K = 30
    try:
        n, d, k = 2_000_000, 50, 30
        reg = reg_for(n, d, 2.0)
        with eng.synthetic(n, d, 16, 2.5, 20261023) as ds:
            mu0, _ = eng.initialize_mu(ds, k, seed=7)
            eng.lloyd_iterations(ds, mu0, reg, True, -1, 5)
            eng.pass_timing(True)
            ms_med, ms_best = time_device(torch, lambda: eng.lloyd_iterations(ds, mu0, reg, True, -1, 20), 3)
            pms, pcnt = eng.pass_timing(False)
            pms /= max(pcnt, 1)
            mu, _ = eng.lloyd_iterations(ds, mu0, reg, True, -1, 20)
            k_act = int((np.abs(mu).sum(1) > 0).sum())
            gbs = n * (4.0 * d + 8.0) / (pms * 1e-3) / 1e9
            blk = {"workload": "synthetic 2M cells x 50 markers, K=30: 20 fused Lloyd iterations of a fresh fit; the penalty search with 24 workers x 11 penalties per set",
                   "lloyd_ms_per_step": ms_med / 20, "lloyd_ms_per_step_best": ms_best / 20, "pass_ms": pms, "pass_GBps": gbs, "pass_frac_of_hbm": gbs / peak,
                   "k_active": k_act, "value": float(n) * k_act * 20 / (ms_med * 1e-3), "unit": UNIT}
            blk.update(assign_block(ds, mu, True, n, d, 20))
            from depecher_b200 import rapi
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                t0 = time.perf_counter()
                opt = rapi.depechePenaltyOpt(ds, k=k, maxIter=100, minARIImprovement=0.01, sampleSize=10000, penalties=2.0 ** np.arange(0, 5.5, 0.5),
                                             optimARI=0.95, nCores=24, engine=eng, seed=7, disableWarnings=True)
                blk["penalty_search_s"] = time.perf_counter() - t0
            blk["penalty_search_workerIterations"] = opt["workerIterations"]
            blk["bestPenalty"] = opt["bestPenalty"]
        out["2Mx50"] = blk
    except Exception as exc:
        out["2Mx50"] = {"error": repr(exc)}
    # ---- configs[3]: dAllocate of 100M x 40 onto fixed centres (6 of 24 rows all-zero)
    try:
        n, d, k = 100_000_000, 40, 24
        rng = np.random.default_rng(4)
        mu = rng.choice([-1.2, 0.0, 1.2], size=(k, d))
        mu[[2, 5, 9, 13, 17, 21]] = 0.0
        with eng.synthetic(n, d, 24, 2.5, 20261022) as ds:
            blk = {"workload": "dAllocate of synthetic 100M cells x 40 markers onto 24 fixed centres (6 all-zero rows, no_zero = TRUE): pure assignment"}
            blk.update(assign_block(ds, mu, True, n, d, 5))
            blk["value"] = float(n) * 18 / (blk["assign_pass_ms"] * 1e-3)
            blk["unit"] = UNIT
            t0 = time.perf_counter()
            lab = eng.allocate_points(ds, mu, True)["i"]
            blk["call_seconds_with_labels_to_host"] = time.perf_counter() - t0
            blk["d2h_bytes"] = int(lab.nbytes)
            del lab
        out["100Mx40_dAllocate"] = blk
    except Exception as exc:
        out["100Mx40_dAllocate"] = {"error": repr(exc)}
    return out


def config0_testdata(eng) -> dict:
    """BASELINE configs[0]: depeche(testData[, 2:15]) with the default penalty grid — the reference's only published timing
    (vignettes/DepecheR_test.Rmd:85-90: 26 s on an 8-core machine).  GPU: rapi.depeche() over the CUDA library; CPU: the same callers
    over the reference's own Clusterer.cpp with the reference's worker-process fan-out (tools/cpu_depeche.py, a subprocess)."""
    blk = {"workload": "depeche(testData[, 2:15]): 20000 cells x 14 markers, K=30, default penalty grid, sampleSize 10000"}
    try:
        from depecher_b200 import rapi
        z = np.load(os.path.join(ROOT, "tests", "golden", "testdata.npz"))
        X = z["X_1e4"].astype(np.float64) / 1e4
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            rapi.depeche(X, nCores=host_workers(), engine=eng, seed=3)
            t0 = time.perf_counter()
            out = rapi.depeche(X, nCores=host_workers(), engine=eng, seed=4)
            blk["gpu_seconds"] = time.perf_counter() - t0
        blk["gpu_bestPenalty"] = out["penaltyOptList"]["bestPenalty"]
        blk["gpu_clusters"] = int(out["clusterCenters"].shape[0])
        blk["gpu_workerIterations"] = out["penaltyOptList"]["workerIterations"]
    except Exception as exc:
        blk["gpu_error"] = repr(exc)
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "cpu_depeche.py")], capture_output=True, text=True, timeout=600)
        cpu = json.loads(r.stdout.strip().splitlines()[-1])
        blk["cpu_baseline_depeche_testdata_s"] = cpu["seconds"]
        blk["cpu"] = cpu
        if "gpu_seconds" in blk:
            blk["speedup"] = cpu["seconds"] / blk["gpu_seconds"]
    except Exception as exc:
        blk["cpu_error"] = repr(exc)
    return blk


def depeche_hot_path(eng, dist, n, d, k, ncores, groups, amp, seed, split: bool) -> dict:
    """the depeche() hot path on a resident matrix: depechePenaltyOpt (whole loop inside the library; with `split` the problems of each
    set dealt over the ranks, every rank holding the same matrix), depecheClusterCenterSelection and the final full-data dAllocate"""
    from depecher_b200 import rapi
    ds = eng.synthetic(n, d, groups, amp, seed)
    eng.synchronize()
    if dist is not None:
        dist.barrier()
    launches0 = eng.kernel_launches()
    t0 = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        opt = rapi.depechePenaltyOpt(ds, k=k, maxIter=100, minARIImprovement=0.01, sampleSize=10000, penalties=2.0 ** np.arange(0, 5.5, 0.5),
                                     optimARI=0.95, nCores=ncores, engine=eng, seed=seed, disableWarnings=True, dist=dist if split else None)
    t1 = time.perf_counter()
    # The search is the part that splits.  What follows it — choosing among the stored solutions on a 10 000-row sample and ONE
    # assignment pass over the matrix, whose cluster vector the caller wants once — is done by rank 0 alone; the other ranks wait
    # (the time reported is the maximum over the ranks, i.e. rank 0's).
    if dist is not None and split and dist.get_rank() != 0:
        dist.barrier()
        t3 = time.perf_counter()
        ds.close()
        return {"total_s": t3 - t0, "penalty_opt_s": t1 - t0}
    rows = np.random.default_rng(seed).integers(0, n, 10000)
    sel = eng.gather(ds, rows)
    r = rapi.depecheClusterCenterSelection(opt["bestClusterCenters"], sel, k, ncores, engine=eng)
    sel.close()
    t2 = time.perf_counter()
    full = np.zeros((r["clusterCenters"].shape[0], d))
    full[:, r["cols"]] = r["clusterCenters"]
    labels = rapi.dAllocate(ds, {"clusterCenters": full, "logCenterSd": False}, engine=eng)
    t3 = time.perf_counter()
    if dist is not None and split:
        dist.barrier()
    out = {"total_s": t3 - t0, "penalty_opt_s": t1 - t0, "selection_s": t2 - t1, "final_allocate_s": t3 - t2,
           "workerIterations": opt["workerIterations"], "bestPenalty": opt["bestPenalty"], "clusters": int(r["clusterCenters"].shape[0]),
           "kernel_launches": eng.kernel_launches() - launches0, "cluster_sizes_head": np.bincount(labels)[1:6].tolist(),
           "ARI": [round(float(v), 4) for v in opt["meanOptimDf"]["ARI"]]}
    ds.close()
    return out


def multirank_parity(eng, dist, torch, local_rank: int, rank: int, world: int) -> dict:
    """row-sharded fit on N GPUs: the peer-memory exchange, the NCCL allreduce and ONE GPU holding all the rows must agree — identical
    labels and iteration counts, centres to rounding (the three add the ranks' tables in different orders), bitwise identical centres
    on every rank (reevaluate_centers' sums, Clusterer.cpp:99-107, over row shards)"""
    from depecher_b200 import shard
    n_local, d, k = 400_000, D, K
    reg = reg_for(n_local * world, d, PENALTY)
    dev = torch.device("cuda", local_rank)
    probe = eng.synthetic(n_local, d, GROUPS, AMP, 77, row0=0, scale=0.0)
    sc = torch.tensor([probe.scale], dtype=torch.float64, device=dev)
    probe.close()
    dist.broadcast(sc, src=0)
    scale = float(sc.item())
    ds = eng.synthetic(n_local, d, GROUPS, AMP, 77, row0=rank * n_local, scale=scale)
    mu0, _ = eng.initialize_mu(ds, k, seed=3, fit_id=0)
    t = torch.from_numpy(np.ascontiguousarray(mu0)).to(dev)
    dist.broadcast(t, src=0)
    mu0 = t.cpu().numpy()
    res = {}
    eng.comm_destroy()
    for mode in ("peer", "nccl"):
        shard.init_comm(eng, dist, device=dev, peer=(mode == "peer"))
        res[mode] = eng.find_centers_from(ds, mu0, reg, True)
        eng.comm_destroy()
    labels_equal = bool(np.array_equal(res["peer"]["i"], res["nccl"]["i"]) and res["peer"]["n_iter"] == res["nccl"]["n_iter"])
    diff_nccl = float(np.abs(res["peer"]["c"] - res["nccl"]["c"]).max())
    c = torch.from_numpy(np.ascontiguousarray(res["peer"]["c"])).to(dev)
    c0 = c.clone()
    dist.broadcast(c0, src=0)
    same = torch.tensor([1 if torch.equal(c, c0) else 0, 1 if labels_equal else 0], device=dev)
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    out = {"labels_equal_nccl": bool(same[1].item()), "centres_bitwise_across_ranks": bool(same[0].item()), "max_centre_diff_peer_vs_nccl": diff_nccl,
           "iterations": int(res["peer"]["n_iter"]), "cells_per_rank": n_local}
    if rank == 0:   # the same rows on one GPU (no communicator at this point: a plain single-GPU fit)
        full = eng.synthetic(n_local * world, d, GROUPS, AMP, 77, row0=0, scale=scale)
        one = eng.find_centers_from(full, mu0, reg, True)
        full.close()
        out["shard0_labels_equal_single_gpu"] = bool(np.array_equal(one["i"][:n_local], res["peer"]["i"]))
        out["single_gpu_iterations"] = int(one["n_iter"])
        out["max_centre_diff"] = float(np.abs(one["c"] - res["peer"]["c"]).max())
    dist.barrier()
    shard.init_comm(eng, dist, device=dev)     # back to the bench's communicator (peer mailboxes when they can be set up)
    ds.close()
    return out


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
    parity = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        from depecher_b200 import shard
        shard.init_comm(eng, dist, device=torch.device("cuda", local_rank))
        try:
            parity = multirank_parity(eng, dist, torch, local_rank, rank, world)
        except Exception as exc:   # noqa: BLE001
            parity = {"error": repr(exc)}

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
    warm = max(args.warmup, 3)
    sampler = ClockSampler(local_rank)   # samples from the warm-up on, so that a short timed region still has clock readings under load
    sampler.start()
    eng.lloyd_iterations(ds, mu0, reg, True, -1, warm)                            # warm-up (>= 3 steps)
    for _ in range(3):                                                            # keep the GPU busy ~100 ms for the sampler
        eng.lloyd_iterations(ds, mu0, reg, True, -1, 30)
    eng.pass_timing(True)
    launches0 = eng.kernel_launches(reset=True)
    rep_ms, mu, changed = [], None, 0
    for _ in range(REPEATS):           # every repeat: EXACTLY args.steps steps of a fresh fit, barrier + synchronize on both sides
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        mu, changed = eng.lloyd_iterations(ds, mu0, reg, True, -1, args.steps)
        e1.record()
        barrier()
        ms_r = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms_r], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)      # max over ranks
            ms_r = float(t.item())
        rep_ms.append(ms_r)
    clocks = sampler.stop()
    launches = eng.kernel_launches() // REPEATS
    pass_ms, pass_cnt = eng.pass_timing(False)
    ms = float(np.median(rep_ms))
    k_eff = int((np.abs(mu).sum(1) > 0).sum())   # clusters only ever die during a fit: the count at the end is the smallest of the run
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

    peak, which = hbm_peak()
    line = None
    if rank == 0:
        bytes_per_launch = float(n_local) * (4 * D + 8)
        launch_ms = pass_ms / max(pass_cnt, 1)
        achieved = bytes_per_launch / (launch_ms * 1e-3) / 1e9
        traffic = committed_traffic()
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                    "traffic_source": f"{TRAFFIC_CSV}: ncu --set full capture of this command, committed (not measured in this run)" if traffic else None,
                    "peak_source": which, "kernel": "tc_pass_kernel", "launch_ms": launch_ms, "launches_timed": pass_cnt,
                    "algorithmic_bytes_per_launch": bytes_per_launch, "fma_per_s": float(n_local) * k_eff * D / (launch_ms * 1e-3),
                    "step_frac": (bytes_per_launch / (ms / args.steps * 1e-3) / 1e9) / peak,
                    "note": "distance products on tcgen05 (fp16-split operands, fp32 accumulate, exact fp64 re-decision of near ties); the launches timed "
                            "include the full-sums first iteration and the iterations in which 1-6 % of the cells change cluster; step_frac is the same "
                            "bytes over the whole step (pass + sums of the moved cells + reduce/finalize)"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"synthetic {n_local} cells x {D} markers per GPU, K={K} ({k_eff} still active at the end), penalty {PENALTY}, "
                                   f"one fused Lloyd iteration per step: iterations 0..{args.steps - 1} of a fresh fit from k-means++ centres",
                       "cells_per_gpu": n_local, "markers": D, "k": K, "k_active": k_eff, "parallelism": f"row-shard x{world}" if world > 1 else "single GPU",
                       "l2": "inputs (1.6 GB per GPU) exceed the 126 MB L2", "labels_changed_last_step": int(changed)},
            "repeats": {"n": REPEATS, "ms_per_step_median": ms / args.steps, "ms_per_step_best": float(np.min(rep_ms)) / args.steps,
                        "ms_per_step_all": [m / args.steps for m in rep_ms], "headline": "median"},
            "clocks": clocks, "gpu_launches": int(launches), "roofline": roofline, "steady_state": steady,
        }
        if parity is not None:
            line["parity"] = parity
    # ---- e2e through the C ABI with host buffers, and the CPU baseline: rank 0 at N=1 only (bounded)
    if rank == 0 and world == 1:
        Xh = np.asfortranarray(eng.read(ds, 0, n_local))            # the R-side matrix: column-major doubles on the host
        eng.synchronize()
        # one untimed call first (pinned staging buffers, kernel modules, and the device memory pool grown to the size of this matrix and
        # its work areas: the first call of a process pays 50-80 ms for that), then three timed calls; the median is the line's value
        eng.sparse_k_means(Xh, K, reg, True, seed_off_set=1)
        e2e_all = []
        for _ in range(3):
            t0 = time.perf_counter()
            res = eng.sparse_k_means(Xh, K, reg, True, seed_off_set=1)
            e2e_all.append(time.perf_counter() - t0)
        dt = float(np.median(e2e_all))
        k_e2e = int((np.abs(res["c"]).sum(1) > 0).sum())
        line["e2e"] = {"value": float(n_local) * k_e2e * res["n_iter"] / dt, "unit": UNIT, "h2d_bytes_per_step": int(Xh.nbytes),
                       "d2h_bytes_per_step": int(res["i"].nbytes + res["c"].nbytes), "host_bytes_per_step": int(Xh.nbytes),
                       "pcie_bytes_per_step": int(n_local) * D * 4 + int(res["i"].nbytes + res["c"].nbytes),
                       "seconds": dt, "seconds_all": e2e_all, "lloyd_iterations": res["n_iter"], "k_active_final": k_e2e, "value_counting_k30": float(n_local) * K * res["n_iter"] / dt,
                       "call": "dpr_sparse_k_means(host double*, ...) - H2D, k-means++ seeding, Lloyd loop, D2H inside; counted with the centres still "
                               "active at the end, like `value`; host threads narrow the doubles to fp32 before they cross PCIe (pcie_bytes_per_step)"}
        # the other half of the metric: depeche() itself on the 10M x 40 host matrix - upload, device-side pre-scaling
        # (kurtosis test, histogram-peak centring, global sd), full default penalty sweep, solution selection, final allocation
        try:
            from depecher_b200 import rapi
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
        sample = np.array(Xh[:CPU_ROWS])
        del Xh
        line["cpu_baseline"] = cpu_baseline_block(sample, CPU_ITERS)
        del sample
    ds.close()
    # ---- depeche() hot path on 10M x 40 (X resident, pre-scaling excluded): at N > 1 every rank holds the matrix and the problems of a set are split
    try:
        # (one untimed run first: kernel modules, the device memory pool and, at N > 1, NCCL's first allreduce are set up by it)
        depeche_hot_path(eng, dist if world > 1 else None, N_CELLS, D, K, 10, GROUPS, AMP, SEED + 1, split=world > 1)
        dres = depeche_hot_path(eng, dist if world > 1 else None, N_CELLS, D, K, 10, GROUPS, AMP, SEED + 1, split=world > 1)
        if world > 1:
            t = torch.tensor([dres["total_s"], dres["penalty_opt_s"]], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dres["total_s"], dres["penalty_opt_s"] = float(t[0].item()), float(t[1].item())
            dres["parallelism"] = (f"grid split x{world}: the (worker x penalty) problems of every set dealt over the ranks, X replicated, one small allreduce per "
                                   "set; solution selection and the final assignment pass on rank 0")
        if rank == 0:
            line["depeche_10Mx40"] = dres
    except Exception as exc:   # the headline line must still print
        if rank == 0:
            line["depeche_10Mx40"] = {"error": repr(exc)}
    if world > 1:
        # strong scaling: the SAME 10M cells in total, row-sharded
        try:
            n_s = N_CELLS // world
            dss = eng.synthetic(n_s, D, GROUPS, AMP, SEED, row0=rank * n_s, scale=1.0)
            reg_s = reg_for(n_s * world, D, PENALTY)
            eng.lloyd_iterations(dss, mu0, reg_s, True, -1, warm)
            reps = []
            for _ in range(REPEATS):
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                mu_ss, _ = eng.lloyd_iterations(dss, mu0, reg_s, True, -1, args.steps)
                e1.record()
                barrier()
                t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                reps.append(float(t.item()))
            dss.close()
            if rank == 0:
                k_ss = int((np.abs(mu_ss).sum(1) > 0).sum())
                m = float(np.median(reps))
                line["strong_scaling"] = {"cells_total": n_s * world, "cells_per_gpu": n_s, "ms_per_step": m / args.steps, "ms_per_step_best": float(np.min(reps)) / args.steps,
                                          "k_active": k_ss, "value": float(n_s) * world * k_ss * args.steps / (m * 1e-3), "unit": UNIT}
        except Exception as exc:
            if rank == 0:
                line["strong_scaling"] = {"error": repr(exc)}
        # e2e at N > 1: every rank hands its row shard to the library as HOST doubles (R layout) and all ranks run one row-sharded fit
        # from the common initial centres: upload, Lloyd loop with the exchange per iteration, labels back - all inside
        ds = eng.synthetic(n_local, D, GROUPS, AMP, SEED, row0=rank * n_local, scale=0.0)
        Xh = np.asfortranarray(eng.read(ds, 0, n_local))
        ds.close()
        eng.synchronize()
        eng.find_centers_from(Xh, mu0, reg, True, max_iter=3)      # one untimed call of the same size (collective: every rank): staging buffers, memory pool
        e2e_all = []
        for _ in range(3):
            barrier()
            t0 = time.perf_counter()
            res = eng.find_centers_from(Xh, mu0, reg, True)
            t = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_all.append(float(t.item()))
        dt = float(np.median(e2e_all))
        if rank == 0:
            k_e2e = int((np.abs(res["c"]).sum(1) > 0).sum())
            line["e2e"] = {"value": float(n_local) * world * k_e2e * res["n_iter"] / dt, "unit": UNIT, "h2d_bytes_per_step": int(Xh.nbytes) * world,
                           "d2h_bytes_per_step": int(res["i"].nbytes + res["c"].nbytes) * world, "host_bytes_per_step": int(Xh.nbytes) * world,
                           "pcie_bytes_per_step": (int(n_local) * D * 4 + int(res["i"].nbytes + res["c"].nbytes)) * world,
                           "seconds": dt, "seconds_all": e2e_all, "lloyd_iterations": res["n_iter"], "k_active_final": k_e2e,
                           "call": "dpr_dataset_create(host double*) + dpr_dataset_find_centers_from(...) on every rank's row shard - H2D, row-sharded "
                                   "Lloyd loop (per-iteration exchange of the K x (D+1) tables through peer memory, ncclAllReduce as fallback), D2H inside; max over ranks"}
        del Xh
    if rank == 0 and world == 1:
        line["configs"] = other_configs(eng, torch, peak)
        line["configs"]["testData"] = config0_testdata(eng)
    if rank == 0:
        print(json.dumps(line), flush=True)
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
