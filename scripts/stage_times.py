"""Stage-level wall-clock breakdown of one GeneClust-fast selection on the device (developer tool)."""
import cProfile
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from scgeneclust_b200 import AnnDataLite, pp, tl
from scgeneclust_b200.synth import DeviceSynth, synth_params

C, G, K = int(sys.argv[1]) if len(sys.argv) > 1 else 50000, 20000, 200
counts = DeviceSynth(synth_params(G, 0)).counts(0, C)
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
print({k: round(v / 3, 2) for k, v in tm.items()})
pr = cProfile.Profile()
pr.enable()
run()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
