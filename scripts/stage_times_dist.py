"""Stage breakdown of one gene-sharded GeneClust-fast selection under torchrun (developer tool)."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from scgeneclust_b200 import AnnDataLite, pp, tl
from scgeneclust_b200.dist import GeneShardComm
from scgeneclust_b200.synth import DeviceSynth, synth_params

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
C, G, K = 50000 * world, 20000, 200
comm = GeneShardComm(G)
counts = DeviceSynth(synth_params(G, 0), torch.device("cuda", local)).counts(0, C, comm.gene_lo, comm.local_genes)
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
    t = tick("adata", t)
    pp.normalize(ad, 'sc')
    t = tick("normalize", t)
    pp.reduce_dim(ad, 'fast', 0)
    t = tick("reduce_dim", t)
    tl.cluster_genes(ad, None, 'fast', n_gene_clusters=K, random_state=0)
    t = tick("cluster_genes", t)
    sel = tl.select_from_clusters(ad, 'fast', True, 0)
    t = tick("select", t)
    return sel


for _ in range(2):
    run()
tm = {}
for _ in range(3):
    run(tm)
if rank == 0:
    print({k: round(v / 3, 2) for k, v in tm.items()})
pr = cProfile.Profile()
pr.enable()
run()
torch.cuda.synchronize()
pr.disable()
if rank == 0:
    pstats.Stats(pr).sort_stats("tottime").print_stats(18)
dist.destroy_process_group()
