#!/usr/bin/env python
"""Bench of the GeneClust gene-grouping hot path on B200 (contract: see the task prompt / DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]

A "step" is one full GeneClust-fast gene selection (`scGeneClust(adata, version='fast', n_var_clusters=200)`:
Pearson residuals -> gene PCA (16 randomized-SVD passes) -> mini-batch k-means -> closeness -> selection) over one
synthetic negative-binomial count matrix.  N = 1 runs BASELINE.json config c3 (50 000 cells x 20 000 genes on one
B200); N > 1 is weak scaling in cells (50 000 * N cells x 20 000 genes), genes sharded across the ranks.
`value` = matrix entries (cells x genes) selected-over per second with the counts already resident in HBM;
`e2e` = the same through the public entry point from pinned HOST counts (H2D of the counts and D2H of the
embedding, labels and closeness inside the timed region).  The same line carries the GeneClust-ps redundancy
stage (gene-pairs/s, BASELINE.json's second metric) under "ps_redundancy".
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CELLS_PER_GPU, GENES, K_CLUSTERS, SEED = 50_000, 20_000, 200, 0
PS_GENES = 4_000                       # config c4: ~20 % of 20 000 genes -> 7 998 000 pairs
CPU_SAMPLE_CELLS = 8_000               # bounded CPU sample of the same matrix
CPU_SAMPLE_PAIRS = 4_000


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--cells", type=int, default=None, help="cells per GPU (default: config c3)")
    ap.add_argument("--genes", type=int, default=GENES)
    ap.add_argument("--clusters", type=int, default=K_CLUSTERS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ps", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-resident arm (config c5 would pin 7.5 GB per rank)")
    ap.add_argument("--cpu-sample-cells", type=int, default=CPU_SAMPLE_CELLS)
    return ap.parse_args()


# -------------------------------------------------------------------------------------------------------------
# clocks / throttle sampling during the timed region
# -------------------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index: int, period: float = 0.02):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None
        self.period = period

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake_slowdown": 0x80, "sync_boost": 0x10, "applications_clocks_setting": 0x2}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# -------------------------------------------------------------------------------------------------------------
# CPU reference arm: the oracle (same sklearn/numpy calls the reference makes) on the host cores
# -------------------------------------------------------------------------------------------------------------
def _pairs_worker(args):
    from oracle import stages
    E, pairs, seed = args
    return stages.gene_redundancy(E, seed, pairs)


def cpu_fast_sample(X_sample: np.ndarray, n_clusters: int):
    """One GeneClust-fast selection on the CPU (oracle/stages.py).  Returns (seconds, stage timings)."""
    from oracle import stages
    names = np.array([f"g{i}" for i in range(X_sample.shape[1])], dtype=object)
    tm = {}
    t0 = time.perf_counter()
    stages.geneclust_fast(X_sample, names, n_clusters, SEED, timings=tm)
    return time.perf_counter() - t0, tm


def cpu_pairs_sample(E: np.ndarray, n_pairs: int, workers: int):
    """KSG redundancy of a fixed random sample of gene pairs with a process pool like the reference
    (scGeneClust/tl/information.py:83-84).  Returns pairs/s."""
    from multiprocessing import get_context
    rs = np.random.RandomState(1)
    pairs = np.sort(np.stack([rs.randint(0, len(E), n_pairs), rs.randint(0, len(E), n_pairs)], 1), 1)
    pairs = pairs[pairs[:, 0] != pairs[:, 1]]
    chunks = np.array_split(pairs, workers * 4)
    with get_context("fork").Pool(workers) as pool:
        pool.map(_pairs_worker, [(E, c[:2], SEED) for c in chunks[:workers]])      # warm the workers
        t0 = time.perf_counter()
        pool.map(_pairs_worker, [(E, c, SEED) for c in chunks])
        dt = time.perf_counter() - t0
    return len(pairs) / dt


def synth_host_sample(params, n_cells: int, n_genes: int):
    """Host copy of cells [0, n_cells) of the synthetic matrix: generated on the device when there is one (fast),
    by the numpy twin otherwise."""
    try:
        import torch
        if torch.cuda.is_available():
            from scgeneclust_b200.synth import DeviceSynth
            return DeviceSynth(params).counts(0, n_cells, 0, n_genes).to_numpy()
    except Exception:  # noqa: BLE001
        pass
    from oracle.synth import synth_counts
    return np.concatenate([synth_counts(params, c0, min(500, n_cells - c0), 0, n_genes)
                           for c0 in range(0, n_cells, 500)])


def run_reference(args, rank: int):
    if rank != 0:
        return
    from scgeneclust_b200.synth import synth_params
    params = synth_params(args.genes, SEED)
    n_cells = min(args.cpu_sample_cells, (args.cells or CELLS_PER_GPU) * args.gpus)
    X = synth_host_sample(params, n_cells, args.genes)
    times, tm = [], {}
    for i in range(args.warmup + args.steps):
        dt, tm = cpu_fast_sample(X, args.clusters)
        if i >= args.warmup:
            times.append(dt)
    sec = float(np.mean(times))
    value = n_cells * args.genes / sec
    cores = os.cpu_count()
    line = {
        "impl": "reference", "metric": "geneclust_fast_selection_throughput", "value": value,
        "unit": "matrix-entries/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": "matrix-entries/s", "cores": cores, "kind": "port",
                         "sample": f"cells[0:{n_cells}] x {args.genes} genes of the same synthetic matrix, full "
                                   f"GeneClust-fast pipeline (numpy Pearson residuals + sklearn PCA/MiniBatchKMeans/"
                                   f"IsolationForest), library threads unrestricted",
                         "stage_s": {k: round(v, 3) for k, v in tm.items()}},
        "e2e": {"value": value, "unit": "matrix-entries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit_line(line)


def workload_config(args, n_gpus):
    cells = (args.cells or CELLS_PER_GPU) * n_gpus
    if cells == 1_000_000 and args.genes == 30_000:
        which = "BASELINE.json config c5"
    elif (args.cells or CELLS_PER_GPU) == CELLS_PER_GPU and args.genes == GENES:
        which = "BASELINE.json config c3 at N=1; cells scale with N"
    else:
        which = "custom size"
    return {"workload": f"GeneClust-fast, synthetic negative-binomial counts {cells} cells x {args.genes} genes, "
                        f"n_var_clusters={args.clusters} ({which})",
            "cells": cells, "genes": args.genes, "n_var_clusters": args.clusters,
            "parallelism": f"genes sharded over {n_gpus} GPU(s)", "cache": f"inputs ({cells * args.genes * 2 / n_gpus / 1e9:.1f} GB/GPU) larger than L2",
            "random_state": SEED}


# -------------------------------------------------------------------------------------------------------------
# native arm
# -------------------------------------------------------------------------------------------------------------
def run_native(args, rank: int, world: int, local_rank: int):
    import torch
    import torch.distributed as dist

    from scgeneclust_b200 import AnnDataLite, scGeneClust
    from scgeneclust_b200._device import HostShard
    from scgeneclust_b200._lib import lib
    from scgeneclust_b200.synth import DeviceSynth, synth_params
    import ctypes

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    L = lib()
    cells = (args.cells or CELLS_PER_GPU) * world
    G = args.genes
    params = synth_params(G, SEED)
    comm = None
    g_lo, g_hi = 0, G
    if world > 1:
        from scgeneclust_b200.dist import GeneShardComm
        comm = GeneShardComm(G)
        g_lo, g_hi = comm.gene_lo, comm.gene_hi
    synth = DeviceSynth(params, dev)
    counts = synth.counts(0, cells, g_lo, g_hi - g_lo)
    counts.comm = comm
    names = np.array([f"g{i}" for i in range(G)], dtype=object)
    host_counts = None
    if not args.no_e2e:
        host_counts = torch.empty((cells, g_hi - g_lo), dtype=torch.int16).pin_memory()
        host_counts.copy_(counts.data[:, :g_hi - g_lo])
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        return scGeneClust(AnnDataLite(counts, var_names=names), n_var_clusters=args.clusters, version='fast',
                           verbosity=0, random_state=SEED)

    def step_e2e():
        arr = host_counts.numpy().view(np.uint16)
        X = HostShard(arr, comm) if comm is not None else arr
        return scGeneClust(AnnDataLite(X, var_names=names), n_var_clusters=args.clusters, version='fast',
                           verbosity=0, random_state=SEED)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        out = None
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = max(e0.elapsed_time(e1), 0.0)
        t = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return out, float(t[0]), float(t[1])

    for _ in range(args.warmup):
        sel = step_resident()
    L.profile_enable(1)
    n0 = L.launch_count()
    with ClockSampler(local_rank) as clk:
        sel, dev_ms, wall_ms = timed(step_resident, args.steps)
    launches = L.launch_count() - n0
    L.profile_enable(0)
    tot, cnt = ctypes.c_double(), ctypes.c_longlong()
    L.profile_collect(0, ctypes.byref(tot), ctypes.byref(cnt))
    ms_per_step = max(dev_ms, wall_ms) / args.steps          # the step has host syncs: take the larger clock
    value = cells * G / (ms_per_step * 1e-3)

    # roofline of the dominant kernel: one randomized-SVD pass re-reads the uint16 counts once
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    pass_bytes = cells * (g_hi - g_lo) * 2                   # algorithmic: 2 B per matrix entry per pass
    avg_ms = tot.value / max(cnt.value, 1)
    achieved = pass_bytes / (avg_ms * 1e-3) / 1e9 if cnt.value else None
    # DRAM traffic of the same kernel from the committed ncu capture (only valid for the workload it was taken on)
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r1_rsvd_tc_v3_traffic.json")))
        if tj["workload"] == {"cells": cells, "genes": G, "n_gpus": world}:
            traffic = tj["mean_bytes_per_launch"]
    except Exception:  # noqa: BLE001
        pass
    roofline = {"kernel": "rsvd_apply_tc_kernel (one Y = M Q / M^T Q pass over the implicit Pearson-residual matrix)",
                "bound": "hbm", "achieved": achieved, "peak": peak_gbs,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                "unit": "GB/s", "frac": (achieved / peak_gbs) if achieved else None, "traffic": traffic,
                "traffic_source": "profiles/r1_rsvd_tc_v3_traffic.json (ncu dram read+write bytes per launch)" if traffic else None,
                "launches": int(cnt.value), "avg_launch_ms": avg_ms, "algorithmic_bytes_per_launch": pass_bytes,
                "share_of_step": (tot.value / args.steps) / ms_per_step if ms_per_step else None}

    # end to end from pinned host counts through the public entry point
    if args.no_e2e:
        e_ms, e2e = float("nan"), None
    else:
        for _ in range(min(2, args.warmup)):
            step_e2e()
        _, e_dev_ms, e_wall_ms = timed(step_e2e, args.steps)
        e_ms = max(e_dev_ms, e_wall_ms) / args.steps
        h2d = int(host_counts.numel() * 2) * world
        d2h = int(G * 50 * 4 + G * 4 + G * 4)
        e2e = {"value": cells * G / (e_ms * 1e-3), "unit": "matrix-entries/s", "ms_per_step": e_ms,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h}

    line = {
        "metric": "geneclust_fast_selection_throughput", "value": value, "unit": "matrix-entries/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "end_to_end_selection_s": None if args.no_e2e else e_ms * 1e-3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args, world),
        "clocks": clk.summary(), "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
        "selected_genes": int(len(sel)),
    }

    # GeneClust-ps redundancy stage (second half of BASELINE.json's metric): KSG MI of all pairs of 4000 genes
    if not args.no_ps:
        from scgeneclust_b200.tl.information import gene_redundancy
        st_emb = step_embedding(scGeneClust, AnnDataLite, counts, names, args)
        E = st_emb[:PS_GENES].contiguous()
        n_pairs_total = E.shape[0] * (E.shape[0] - 1) // 2
        lo, hi = (0, n_pairs_total) if comm is None else comm.split_range(n_pairs_total)
        for _ in range(2):
            gene_redundancy(E, SEED, lo, hi, dense=False)
        L.profile_enable(1)
        _, p_dev_ms, p_wall_ms = timed(lambda: gene_redundancy(E, SEED, lo, hi, dense=False), 3)
        L.profile_enable(0)
        L.profile_collect(1, ctypes.byref(tot), ctypes.byref(cnt))
        p_ms = max(p_dev_ms, p_wall_ms) / 3
        n = E.shape[1]
        # roofline of the KSG pair kernel: fp64 / integer issue bound; denominator = the DFMA issue rate measured on
        # this GPU right now (MEASURED_PEAKS.json has no fp64 figure); algorithmic work = 8 n^2 fp64-class ops per pair
        peak64 = ctypes.c_double()
        probe_buf = torch.zeros(1, dtype=torch.float64, device=dev)
        L.fp64_probe(ctypes.byref(peak64), probe_buf.data_ptr(), torch.cuda.current_stream().cuda_stream)
        k_ms = tot.value / max(cnt.value, 1)
        ops = 8.0 * n * n * (hi - lo) / (k_ms * 1e-3)
        line["ps_redundancy"] = {
            "roofline": {"kernel": "mi_cc_pairs_kernel", "bound": "fp64-issue", "achieved": ops / 1e12,
                         "peak": peak64.value / 1e12, "unit": "T fp64-class op/s", "frac": ops / peak64.value,
                         "peak_source": "gc_fp64_probe (DFMA lane-ops/s measured in this run)",
                         "algorithmic_ops_per_pair": 8.0 * n * n},
            "metric": "gene-pairs/s", "value": n_pairs_total / (p_ms * 1e-3), "pairs": n_pairs_total,
            "genes": int(E.shape[0]), "samples_per_gene": int(n), "ms": p_ms,
            "kernel_ms": tot.value / max(cnt.value, 1),
            "fp64_class_ops_per_s": 8.0 * n * n * (hi - lo) / (tot.value / max(cnt.value, 1) * 1e-3),
            "config": "BASELINE.json config c4 size: 4000 relevant genes x 50 PCs, k=3, pairs split over ranks"}

    # GeneClust-ps through the public entry point at config c4 size (single GPU only: the ps path keeps the dense
    # residuals of the high-confidence cells on one device): 10 000 cells x 20 000 genes from HOST counts, the
    # generator's cell types as pseudo cell types, a deterministic 60 % of the cells marked high-confidence
    if not args.no_ps and world == 1:
        from scgeneclust_b200.synth import synth_cell_types
        c4_cells = 10_000
        Xc4 = synth.counts(0, c4_cells).to_numpy()
        keep = (Xc4 > 0).sum(0) >= 3                                 # the reference loader's min_cells=3 filter (_utils.py:35)
        Xc4 = np.ascontiguousarray(Xc4[:, keep])
        names4 = names[keep]
        types4 = np.array([f"type{t}" for t in synth_cell_types(params, 0, c4_cells)])
        hc4 = np.random.RandomState(SEED).rand(c4_cells) < 0.6

        def step_ps():
            ad = AnnDataLite(Xc4, var_names=names4)
            ad.obs['cluster'] = types4
            ad.obs['highly_confident'] = hc4
            return scGeneClust(ad, version='ps', n_obs_clusters=8, verbosity=0, random_state=SEED)
        step_ps()
        sel_ps, ps_dev_ms, ps_wall_ms = timed(step_ps, 3)
        line["ps_end_to_end"] = {
            "seconds": max(ps_dev_ms, ps_wall_ms) / 3 * 1e-3, "selected_genes": int(len(sel_ps)),
            "config": f"BASELINE.json config c4: synthetic {c4_cells} cells x {int(keep.sum())} genes (of {G}, min_cells=3), "
                      f"{int(hc4.sum())} high-confidence cells in 8 given pseudo cell types, relevant_gene_pct=20, "
                      f"host counts -> selected genes through scGeneClust(version='ps')"}

    if rank == 0 and not args.no_cpu_baseline:
        n_cells = min(args.cpu_sample_cells, cells)
        Xs = synth_host_sample(params, n_cells, G)
        sec, tm = cpu_fast_sample(Xs, args.clusters)
        line["cpu_baseline"] = {
            "value": n_cells * G / sec, "unit": "matrix-entries/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"cells[0:{n_cells}] x {G} genes of the same synthetic matrix, one full GeneClust-fast "
                      f"selection with the oracle (numpy residuals + sklearn PCA/MiniBatchKMeans/IsolationForest)",
            "seconds": sec, "stage_s": {k: round(v, 3) for k, v in tm.items()}}
        if not args.no_ps:
            workers = max(1, os.cpu_count() - 1)
            pps = cpu_pairs_sample(E.cpu().numpy(), CPU_SAMPLE_PAIRS, workers)
            line["cpu_baseline"]["ps_pairs_per_s"] = pps
            line["cpu_baseline"]["ps_sample"] = (f"{CPU_SAMPLE_PAIRS} random gene pairs, sklearn mutual_info_regression "
                                                 f"per pair in a {workers}-process pool (as the reference)")
    if rank == 0:
        emit_line(line)


def step_embedding(scGeneClust, AnnDataLite, counts, names, args):
    """Device embedding (genes x 50) of one fast run - the input of the ps redundancy stage."""
    from scgeneclust_b200 import pp
    from scgeneclust_b200.pp._preprocessing import get_state
    ad = AnnDataLite(counts, var_names=names)
    pp.normalize(ad, 'sc')
    pp.reduce_dim(ad, 'fast', SEED)
    return get_state(ad).varm_pca


class _CleanStdout:
    """stdout carries exactly one JSON line: native libraries (NCCL's version banner) write to file descriptor 1 too, so
    fd 1 is pointed at stderr for the duration of the run and the line is written to the saved descriptor."""

    def __enter__(self):
        sys.stdout.flush()
        self._saved = os.dup(1)
        os.dup2(2, 1)
        self._out = os.fdopen(self._saved, "w")
        return self

    def emit(self, text: str):
        self._out.write(text + "\n")
        self._out.flush()

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self._saved, 1)
        return False


_STDOUT = None


def emit_line(line: dict):
    text = json.dumps(line)
    if _STDOUT is not None:
        _STDOUT.emit(text)
    else:
        print(text, flush=True)


def main():
    global _STDOUT
    with _CleanStdout() as _STDOUT:
        _main()
    _STDOUT = None


def _main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (native arm) needs a CUDA device: the hot path has no CPU fallback")
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_native(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
