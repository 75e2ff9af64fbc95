"""Per-kernel GPU time (torch profiler / CUPTI, rank 0) + stage wall clock of the gene-sharded GeneClust-fast step under
torchrun (developer tool; also runs with one process).   torchrun ... scripts/profile_dist.py [cells_per_gpu] [genes] [out]"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from scgeneclust_b200 import AnnDataLite, pp, tl
from scgeneclust_b200.synth import DeviceSynth, synth_params

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cells_per_gpu = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
G = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
out_path = sys.argv[3] if len(sys.argv) > 3 else None
C, K = cells_per_gpu * world, 200
comm = None
g_lo, ng = 0, G
if world > 1:
    from scgeneclust_b200.dist import GeneShardComm
    comm = GeneShardComm(G)
    g_lo, ng = comm.gene_lo, comm.local_genes
counts = DeviceSynth(synth_params(G, 0), torch.device("cuda", local)).counts(0, C, g_lo, ng)
counts.comm = comm
names = np.array([f"g{i}" for i in range(G)], dtype=object)


def run(timing=None):
    def tick(name, t0):
        torch.cuda.synchronize()
        if timing is not None:
            timing[name] = timing.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
        return time.perf_counter()
    t = time.perf_counter()
    ad = AnnDataLite(counts, var_names=names)
    pp.start_omega(ad, 0)
    pp.normalize(ad, 'sc')
    t = tick("normalize", t)
    pp.reduce_dim(ad, 'fast', 0)
    t = tick("reduce_dim", t)
    tl.cluster_genes(ad, None, 'fast', n_gene_clusters=K, random_state=0)
    t = tick("cluster_genes", t)
    sel = tl.select_from_clusters(ad, 'fast', True, 0)
    t = tick("select", t)
    return sel


for _ in range(3):
    run()
tm = {}
for _ in range(5):
    run(tm)
lines = [f"world {world}, {C} cells x {G} genes: stage wall ms (rank {rank}) " + str({k: round(v / 5, 2) for k, v in tm.items()})]
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        run()
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages():
    dt = getattr(e, "device_time_total", None)
    if dt is None:
        dt = getattr(e, "cuda_time_total", 0)
    if dt > 0 and e.device_type is not None and "cuda" in str(e.device_type).lower():
        rows.append((dt / 3 / 1e3, e.count // 3, e.key[:90]))
rows.sort(reverse=True)
lines.append(f"GPU kernels per step (rank {rank}): total {sum(r[0] for r in rows):.3f} ms")
for ms, cnt, key in rows[:40]:
    lines.append(f"  {ms:9.3f} ms {cnt:5d} x  {key}")
text = "\n".join(lines)
if rank == 0:
    print(text, flush=True)
    if out_path:
        open(out_path, "w").write(text + "\n")
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
