#!/usr/bin/env python
"""Bench of the GeneClust gene-grouping hot path on B200 (contract: see the task prompt / DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]

A "step" is one full GeneClust-fast gene selection (`scGeneClust(adata, version='fast', n_var_clusters=200)`:
Pearson residuals -> gene PCA (16 randomized-SVD passes) -> mini-batch k-means -> closeness -> selection) over one
synthetic negative-binomial count matrix.  N = 1 runs BASELINE.json config c3 (50 000 cells x 20 000 genes on one
B200); N > 1 is weak scaling in cells (50 000 * N cells x 20 000 genes), genes sharded across the ranks.

`value`  = matrix entries (cells x genes) selected-over per second with the uint16 counts already resident in HBM;
`e2e`    = the same through the public entry point from a float32 scipy CSR matrix in HOST memory (what `adata.X`
           holds after a scanpy load): H2D of values / indices / indptr and D2H of the embedding, labels and
           closeness inside the timed region.  `e2e_uint16` / `e2e_f32_dense`: the same from a dense host array.
The same line carries the GeneClust-ps redundancy stage (gene-pairs/s, BASELINE.json's second metric) under
"ps_redundancy", GeneClust-ps through the entry point under "ps_end_to_end", and for N > 1 a "sharded_parity"
block (gene-sharded result vs the single-GPU result on the same matrix, computed outside the timed region).

`--impl reference` times the CPU oracle (oracle/stages.py: the same numpy / scikit-learn calls the reference makes) on
the host cores over the FULL config-c3 matrix per step (N = 1); it imports nothing of the product package.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
import zlib

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

_THREAD_VARS = ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "NUMEXPR_NUM_THREADS")
_INHERITED_THREAD_ENV = {k: os.environ[k] for k in _THREAD_VARS if k in os.environ}
if "--impl" in sys.argv and "reference" in sys.argv:
    # the CPU arm uses every host core, as the reference does (library threads unrestricted);
    # torchrun exports OMP_NUM_THREADS=1 to its workers - undo that BEFORE numpy / OpenBLAS / OpenMP initialise
    for _k in _THREAD_VARS:
        os.environ.pop(_k, None)

import numpy as np  # noqa: E402

CELLS_PER_GPU, GENES, K_CLUSTERS, SEED = 50_000, 20_000, 200, 0
PS_GENES = 4_000                       # config c4: ~20 % of 20 000 genes -> 7 998 000 pairs
CPU_SAMPLE_PAIRS = 4_000
SUBSAMPLE_CELLS = 20_000               # CPU oracle on cells[0:20 000] of the same stream
REF_BUDGET_S = 1500.0                  # wall-clock valve of the CPU arm (the driver's limit is 1800 s per run)
METRIC, UNIT = "geneclust_fast_selection_throughput", "matrix-entries/s"
# what the tensor cores consume in the dominant kernel: every fp32 operand is split into bf16 hi + bf16 lo (16 mantissa
# bits together), three bf16 products per term, fp32 accumulation in TMEM
DTYPE = "bf16x2-split operands (16 mantissa bits), f32 accumulate"


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
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-resident arms (config c5 would pin 7.5 GB per rank)")
    ap.add_argument("--no-parity", action="store_true", help="skip the sharded-vs-single-GPU parity block (N > 1)")
    ap.add_argument("--verify-subsample", action="store_true",
                    help=f"also run the CPU oracle on cells[0:{SUBSAMPLE_CELLS}] of the same stream and compare the "
                         f"selected-gene list with the GPU run on that block (BASELINE.md section 3; default at config c5)")
    ap.add_argument("--cpu-sample-cells", type=int, default=None,
                    help="cells of the CPU arms (default: the full matrix at N = 1, config c3 size at N > 1)")
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
# CPU arms: the oracle (same sklearn/numpy calls the reference makes) on the host cores.  Input comes from the oracle's
# own generator (oracle/synth.py + oracle/synth_c.c); nothing of the product package or its .so is touched.
# -------------------------------------------------------------------------------------------------------------
def host_threads() -> dict:
    """Thread-pool sizes the numeric libraries will really use in this process."""
    out = {"cpu_count": os.cpu_count(), "inherited_env": _INHERITED_THREAD_ENV or None}
    try:
        from threadpoolctl import threadpool_info
        out["pools"] = sorted({(p.get("user_api"), p.get("num_threads")) for p in threadpool_info()})
        out["threads"] = max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:  # noqa: BLE001
        out["threads"] = os.cpu_count()
    return out


def oracle_counts(n_cells: int, n_genes: int) -> np.ndarray:
    """cells [0, n_cells) of the synthetic matrix (data seed SEED) from the oracle's C generator."""
    from oracle import synth as osy
    return osy.synth_counts_c(osy.synth_params(n_genes, SEED), 0, n_cells, 0, n_genes)


def cpu_fast(X: np.ndarray, n_clusters: int):
    """One GeneClust-fast selection on the CPU (oracle/stages.py).  Returns (seconds, stage timings, result)."""
    from oracle import stages
    names = np.array([f"g{i}" for i in range(X.shape[1])], dtype=object)
    tm = {}
    t0 = time.perf_counter()
    res = stages.geneclust_fast(X, names, n_clusters, SEED, timings=tm)
    return time.perf_counter() - t0, tm, res


def _pairs_worker(args):
    from oracle import stages
    E, pairs, seed = args
    return stages.gene_redundancy(E, seed, pairs)


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


def cpu_sample_cells(args, n_gpus: int) -> int:
    total = (args.cells or CELLS_PER_GPU) * n_gpus
    if args.cpu_sample_cells:
        return min(args.cpu_sample_cells, total)
    # the full matrix when it is config c3 (or smaller); at N > 1 / config c5 sizes the CPU arm would need > 100 GB of
    # float32 temporaries, so it runs on the first CELLS_PER_GPU cells and the workload is labelled as sampled
    return min(total, CELLS_PER_GPU)


def run_reference(args, rank: int):
    if rank != 0:
        return
    total_cells = (args.cells or CELLS_PER_GPU) * args.gpus
    n_cells = cpu_sample_cells(args, args.gpus)
    sampled = n_cells < total_cells
    t_start = time.perf_counter()
    X = oracle_counts(n_cells, args.genes)
    threads = host_threads()
    # a warm-up step costs as much as a timed one on the CPU (tens of seconds): one is enough to page the libraries in
    warm = min(args.warmup, 1)
    times, tm, truncated = [], {}, False
    for i in range(warm + args.steps):
        dt, tm, _ = cpu_fast(X, args.clusters)
        if i >= warm:
            times.append(dt)
        if time.perf_counter() - t_start + dt > REF_BUDGET_S and i + 1 < warm + args.steps and times:
            truncated = True                    # the next step would cross the wall-clock valve: stop, and say so
            break
    sec = float(np.mean(times))
    value = n_cells * args.genes / sec
    cfg = workload_config(args, args.gpus)
    if sampled:
        cfg["workload"] = (f"SAMPLED: cells[0:{n_cells}] of the {total_cells} cells x {args.genes} genes matrix "
                           f"(the CPU arm at full size needs > 100 GB of float32 temporaries); " + cfg["workload"])
        cfg["cpu_cells"] = n_cells
    line = {
        "impl": "reference", "metric": METRIC, "value": value,
        "unit": UNIT, "n_gpus": args.gpus, "steps": len(times), "warmup": warm,
        "steps_requested": args.steps, "warmup_requested": args.warmup, "truncated_by_time_budget": truncated,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads["threads"], "kind": "port",
                         "sample": (f"{'cells[0:%d] of the' % n_cells if sampled else 'the FULL'} {total_cells} x "
                                    f"{args.genes} synthetic matrix per step, whole GeneClust-fast pipeline of "
                                    f"oracle/stages.py (numpy Pearson residuals + sklearn PCA / MiniBatchKMeans / "
                                    f"IsolationForest), library threads unrestricted; input from oracle/synth_c.c"),
                         "host_threads": threads, "stage_s": {k: round(v, 3) for k, v in tm.items()}},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        # evidence that nothing of the product ran in this arm: the product's shared library is not even mapped
        "product_so_mapped": any("libgeneclust_b200" in ln for ln in open("/proc/self/maps")),
    }
    emit_line(line)


def workload_config(args, n_gpus):
    cells = (args.cells or CELLS_PER_GPU) * n_gpus
    if cells == 1_000_000 and args.genes == 30_000:
        which = "BASELINE.json config c5"
    elif (args.cells or CELLS_PER_GPU) == CELLS_PER_GPU and args.genes == GENES:
        which = "BASELINE.json config c3" if n_gpus == 1 else "BASELINE.json config c3 per GPU; cells scale with N"
    else:
        which = "custom size"
    return {"workload": f"GeneClust-fast, synthetic negative-binomial counts {cells} cells x {args.genes} genes, "
                        f"n_var_clusters={args.clusters} ({which})",
            "cells": cells, "genes": args.genes, "n_var_clusters": args.clusters,
            "parallelism": f"genes sharded over {n_gpus} GPU(s)",
            "cache": f"inputs ({cells * args.genes * 2 / n_gpus / 1e9:.1f} GB/GPU) larger than L2",
            "random_state": SEED}


# -------------------------------------------------------------------------------------------------------------
# native arm
# -------------------------------------------------------------------------------------------------------------
def host_csr_from_device(counts, torch, rows_per_chunk: int = 4096):
    """float32 scipy CSR (values, int32 indices, int64 indptr in PINNED host memory) of a DeviceCounts block - the
    host-side input of the e2e arm, built once outside the timed region."""
    import scipy.sparse as sp
    C_, G_ = counts.C, counts.G
    nnz_rows = torch.zeros(C_, dtype=torch.int64, device=counts.device)
    for r0 in range(0, C_, rows_per_chunk):
        nnz_rows[r0:r0 + rows_per_chunk] = (counts.data[r0:r0 + rows_per_chunk, :G_] != 0).sum(1)
    indptr = torch.zeros(C_ + 1, dtype=torch.int64)
    indptr[1:] = torch.cumsum(nnz_rows.cpu(), 0)
    nnz = int(indptr[-1])
    data = torch.empty(nnz, dtype=torch.float32).pin_memory()
    indices = torch.empty(nnz, dtype=torch.int32).pin_memory()
    indptr = indptr.pin_memory()
    for r0 in range(0, C_, rows_per_chunk):
        blk = counts.data[r0:r0 + rows_per_chunk, :G_]
        nz = torch.nonzero(blk)                             # row-major order = CSR order
        lo, hi = int(indptr[r0]), int(indptr[min(C_, r0 + rows_per_chunk)])
        vals = blk[nz[:, 0], nz[:, 1]].to(torch.int32) & 0xFFFF
        data[lo:hi].copy_(vals.to(torch.float32))
        indices[lo:hi].copy_(nz[:, 1].to(torch.int32))
    torch.cuda.synchronize()
    m = sp.csr_matrix((data.numpy(), indices.numpy(), indptr.numpy()), shape=(C_, G_), copy=False)
    m._pinned_keepalive = (data, indices, indptr)
    return m, int(data.numel() * 4 + indices.numel() * 4 + indptr.numel() * 8)


def fast_stages(AnnDataLite, X, names, n_clusters):
    """One GeneClust-fast selection through the stage functions, keeping the intermediate slots (used outside the timed
    region for the parity blocks; `return_info=True` would materialise the dense residual matrix)."""
    from scgeneclust_b200 import pp, tl
    ad = AnnDataLite(X, var_names=names)
    pp.normalize(ad, 'sc')
    pp.reduce_dim(ad, 'fast', SEED)
    tl.cluster_genes(ad, None, 'fast', n_gene_clusters=n_clusters, random_state=SEED)
    sel = tl.select_from_clusters(ad, 'fast', True, SEED)
    return dict(X_pca=np.asarray(ad.varm['X_pca']), labels=ad.var['cluster'].to_numpy(),
                closeness=ad.var['closeness'].to_numpy(), selected=sel)


def compare_runs(a: dict, b: dict, names, n_clusters: int) -> dict:
    """Link-wise parity of run `a` (GPU) against run `b` (BASELINE.md section 3 parity gates; oracle.stages.fast_parity):
    PCA scores to float tolerance, the k-means / closeness / selection tail EXACT against scikit-learn on a's own
    embedding, and the end-to-end label / list agreement as a reported figure."""
    from oracle import stages
    return stages.fast_parity(a, b, names, n_clusters, SEED)


def run_native(args, rank: int, world: int, local_rank: int):
    import ctypes

    import torch
    import torch.distributed as dist

    from scgeneclust_b200 import AnnDataLite, scGeneClust
    from scgeneclust_b200._device import HostShard
    from scgeneclust_b200._lib import lib
    from scgeneclust_b200.synth import DeviceSynth, synth_params

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
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def select(X):
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
        sel = select(counts)
    L.profile_enable(1)
    n0 = L.launch_count()
    with ClockSampler(local_rank) as clk:
        sel, dev_ms, wall_ms = timed(lambda: select(counts), args.steps)
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
    traffic, traffic_src = None, None
    for name in ("r2_rsvd_tc_traffic.json", "r1_rsvd_tc_v3_traffic.json"):
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", name)))
            if tj["workload"] == {"cells": cells, "genes": G, "n_gpus": world}:
                traffic, traffic_src = tj["mean_bytes_per_launch"], f"profiles/{name} (ncu dram read+write bytes per launch)"
                break
        except Exception:  # noqa: BLE001
            pass
    roofline = {"kernel": "rsvd_apply_ts_kernel (one Y = M Q / M^T Q pass over the implicit Pearson-residual matrix)",
                "bound": "hbm", "achieved": achieved, "peak": peak_gbs,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                "unit": "GB/s", "frac": (achieved / peak_gbs) if achieved else None, "traffic": traffic,
                "traffic_source": traffic_src,
                "launches": int(cnt.value), "avg_launch_ms": avg_ms, "algorithmic_bytes_per_launch": pass_bytes,
                "share_of_step": (tot.value / args.steps) / ms_per_step if ms_per_step else None,
                "whole_step": {"algorithmic_bytes": (2 * 7 + 2 + 1) * pass_bytes,
                               "note": "16 passes + 1 count-sum pass over the uint16 counts",
                               "achieved_gbs": (2 * 7 + 2 + 1) * pass_bytes / (ms_per_step * 1e-3) / 1e9}}

    # ---- end to end from HOST memory through the public entry point ------------------------------------------
    e2e = e2e_u16 = e2e_f32 = None
    e_ms = float("nan")
    if not args.no_e2e:
        def wrap(arr):
            return HostShard(arr, comm) if comm is not None else arr

        def time_e2e(X, h2d):
            for _ in range(min(2, args.warmup)):
                select(X)
            _, d_ms, w_ms = timed(lambda: select(X), args.steps)
            ms = max(d_ms, w_ms) / args.steps
            return {"value": cells * G / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
                    "h2d_bytes_per_step": int(h2d) * world, "d2h_bytes_per_step": int(G * 50 * 4 + G * 4 + G * 4)}

        # (a) the headline: float32 CSR, what an AnnData loaded by scanpy holds
        csr, csr_bytes = host_csr_from_device(counts, torch)
        e2e = time_e2e(wrap(csr), csr_bytes)
        e2e["input"] = (f"scipy CSR float32 values + int32 indices + int64 indptr in pinned host memory, "
                        f"{csr.nnz} stored entries per rank ({100.0 * csr.nnz / (cells * (g_hi - g_lo)):.1f} % dense)")
        e_ms = e2e["ms_per_step"]
        del csr
        # (b) dense uint16 host counts (the smallest dense form; last round's e2e input)
        host_u16 = torch.empty((cells, g_hi - g_lo), dtype=torch.int16).pin_memory()
        host_u16.copy_(counts.data[:, :g_hi - g_lo])
        torch.cuda.synchronize()
        e2e_u16 = time_e2e(wrap(host_u16.numpy().view(np.uint16)), host_u16.numel() * 2)
        e2e_u16["input"] = "dense uint16 counts in pinned host memory"
        del host_u16
        # (c) dense float32 host counts (AnnData.X as a dense float32 ndarray)
        if world == 1:
            host_f32 = torch.empty((cells, G), dtype=torch.float32).pin_memory()
            for r0 in range(0, cells, 8192):
                host_f32[r0:r0 + 8192].copy_((counts.data[r0:r0 + 8192, :G].to(torch.int32) & 0xFFFF).to(torch.float32))
            torch.cuda.synchronize()
            e2e_f32 = time_e2e(host_f32.numpy(), host_f32.numel() * 4)
            e2e_f32["input"] = "dense float32 counts in pinned host memory"
            del host_f32

    line = {
        "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "end_to_end_selection_s": None if args.no_e2e else e_ms * 1e-3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": DTYPE, "data": "synthetic", "config": workload_config(args, world),
        "clocks": clk.summary(), "e2e": e2e, "e2e_uint16": e2e_u16, "e2e_f32_dense": e2e_f32,
        "gpu_launches": int(launches), "roofline": roofline, "selected_genes": int(len(sel)),
        "parity_note": "a1 (scanpy Pearson residuals) is restated from the published algorithm: parity unpinned",
    }

    # ---- gene-sharded vs single-GPU on the same matrix (outside the timed region) ---------------------------
    if world > 1 and not args.no_parity:
        sharded = fast_stages(AnnDataLite, counts, names, args.clusters)
        if rank == 0:
            full = synth.counts(0, cells, 0, G)
            single = fast_stages(AnnDataLite, full, names, args.clusters)
            del full
            line["sharded_parity"] = dict(compare_runs(sharded, single, names, args.clusters),
                                          vs=f"single-GPU run of the same {cells} x {G} matrix on rank 0")
        barrier()

    # ---- subsample verification against the CPU oracle ------------------------------
    if args.verify_subsample or (cells == 1_000_000 and G == 30_000):
        n_sub = min(SUBSAMPLE_CELLS, cells)
        sub = synth.counts(0, n_sub, g_lo, g_hi - g_lo)
        sub.comm = comm
        from oracle import synth as osy
        twin = osy.synth_counts_c(osy.synth_params(G, SEED), 0, n_sub, g_lo, g_hi - g_lo)
        same = torch.tensor([int(np.array_equal(sub.to_numpy(), twin))], device=dev)
        if world > 1:
            dist.all_reduce(same, op=dist.ReduceOp.MIN)
        gpu_sub = fast_stages(AnnDataLite, sub, names, args.clusters)
        if rank == 0:
            Xs = twin if world == 1 else osy.synth_counts_c(osy.synth_params(G, SEED), 0, n_sub, 0, G)
            sec, tm, cpu_sub = cpu_fast(Xs, args.clusters)
            line["subsample_check"] = dict(
                compare_runs(gpu_sub, cpu_sub, names, args.clusters), cells=n_sub, genes=G, gpus=world, cpu_seconds=sec,
                input_identical=bool(same.item()),
                what=f"GPU ({world} rank(s), genes sharded) vs CPU oracle on cells[0:{n_sub}] of the same counter-based "
                     f"stream; the device-generated block is compared bit for bit with the oracle generator's")
        barrier()

    # ---- GeneClust-ps redundancy stage (second half of BASELINE.json's metric): KSG MI of all pairs of 4000 genes
    E = None
    if not args.no_ps:
        from scgeneclust_b200.tl.information import gene_redundancy
        E = step_embedding(AnnDataLite, counts, names)[:PS_GENES].contiguous()
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
            "genes": int(E.shape[0]), "samples_per_gene": int(n), "ms": p_ms, "kernel_ms": k_ms,
            "fp64_class_ops_per_s": ops,
            "config": "BASELINE.json config c4 size: 4000 relevant genes x 50 PCs, k=3, pairs split over ranks"}

    # ---- GeneClust-ps through the public entry point at config c4 size: 10 000 cells x 20 000 genes from HOST counts,
    # the generator's cell types as pseudo cell types, a deterministic 60 % of the cells marked high-confidence; with
    # N > 1 the same call runs gene-sharded (ps_sharded_entry)
    if not args.no_ps:
        ps = run_ps_entry(args, synth, params, names, comm, timed, AnnDataLite, scGeneClust)
        if ps is not None:
            line["ps_end_to_end"] = ps

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_cells = cpu_sample_cells(args, world)
        Xs = oracle_counts(n_cells, G)
        threads = host_threads()
        sec, tm, cpu_res = cpu_fast(Xs, args.clusters)
        del Xs
        line["cpu_baseline"] = {
            "value": n_cells * G / sec, "unit": UNIT, "cores": threads["threads"], "kind": "port",
            "sample": (f"ONE step over {'the FULL' if n_cells == cells else 'cells[0:%d] of the' % n_cells} {cells} x {G} "
                       f"synthetic matrix with the oracle (numpy residuals + sklearn PCA/MiniBatchKMeans/IsolationForest)"),
            "seconds": sec, "stage_s": {k: round(v, 3) for k, v in tm.items()}, "host_threads": threads}
        if n_cells == cells:
            # the same full-size CPU result doubles as a parity check of the benchmarked GPU run
            gpu_res = fast_stages(AnnDataLite, counts, names, args.clusters)
            line["cpu_parity_full_size"] = compare_runs(gpu_res, cpu_res, names, args.clusters)
        if E is not None:
            workers = max(1, os.cpu_count() - 1)
            pps = cpu_pairs_sample(E.cpu().numpy(), CPU_SAMPLE_PAIRS, workers)
            line["cpu_baseline"]["ps_pairs_per_s"] = pps
            line["cpu_baseline"]["ps_sample"] = (f"{CPU_SAMPLE_PAIRS} random gene pairs, sklearn mutual_info_regression "
                                                 f"per pair in a {workers}-process pool (as the reference)")
    if rank == 0:
        emit_line(line)


def run_ps_entry(args, synth, params, names, comm, timed, AnnDataLite, scGeneClust):
    """`scGeneClust(version='ps')` at config c4 size from host counts.  Returns the report block (rank 0 semantics: every
    rank returns the same numbers)."""
    from scgeneclust_b200.synth import synth_cell_types
    c4_cells = 10_000
    G = args.genes
    Xc4 = synth.counts(0, c4_cells).to_numpy()
    keep = (Xc4 > 0).sum(0) >= 3                                 # the reference loader's min_cells=3 filter (_utils.py:35)
    Xc4 = np.ascontiguousarray(Xc4[:, keep])
    names4 = names[keep]
    types4 = np.array([f"type{t}" for t in synth_cell_types(params, 0, c4_cells)])
    hc4 = np.random.RandomState(SEED).rand(c4_cells) < 0.6
    world = 1 if comm is None else comm.world

    def make_adata():
        if world > 1:
            from scgeneclust_b200._device import HostShard
            from scgeneclust_b200.dist import GeneShardComm
            c4comm = GeneShardComm(int(keep.sum()))
            X = HostShard(np.ascontiguousarray(Xc4[:, c4comm.gene_lo:c4comm.gene_hi]), c4comm)
        else:
            X = Xc4
        ad = AnnDataLite(X, var_names=names4)
        ad.obs['cluster'] = types4
        ad.obs['highly_confident'] = hc4
        return ad

    def step_ps():
        return scGeneClust(make_adata(), version='ps', n_obs_clusters=8, verbosity=0, random_state=SEED)

    try:
        step_ps()
    except NotImplementedError as exc:
        return {"unavailable": str(exc)} if world > 1 else None
    sel_ps, ps_dev_ms, ps_wall_ms = timed(step_ps, 3)
    return {
        "seconds": max(ps_dev_ms, ps_wall_ms) / 3 * 1e-3, "selected_genes": int(len(sel_ps)), "n_gpus": world,
        "selected_crc32": zlib.crc32(",".join(str(s) for s in sel_ps).encode()),
        "selected_first": [str(s) for s in sel_ps[:5]],
        "config": f"BASELINE.json config c4: synthetic {c4_cells} cells x {int(keep.sum())} genes (of {G}, min_cells=3), "
                  f"{int(hc4.sum())} high-confidence cells in 8 given pseudo cell types, relevant_gene_pct=20, "
                  f"host counts -> selected genes through scGeneClust(version='ps')"
                  + (f", genes sharded over {world} ranks" if world > 1 else "")}


def step_embedding(AnnDataLite, counts, names):
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
