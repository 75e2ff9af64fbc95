"""Stage-level wall-clock breakdown of one GeneClust-ps run at config c4 size (developer tool): synthetic 10 000 cells x
20 000 genes, generator cell types as pseudo cell types, 60 % of the cells marked high-confidence."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from scgeneclust_b200.synth import synth_cell_types as cell_types
from scgeneclust_b200 import AnnDataLite, pp, tl
from scgeneclust_b200.synth import DeviceSynth, synth_params
from scgeneclust_b200.tl.cluster import generate_gene_clusters
from scgeneclust_b200.tl.confidence import find_high_confidence_cells
from scgeneclust_b200.tl.information import compute_gene_complementarity, compute_gene_redundancy, find_relevant_genes

C, G = int(sys.argv[1]) if len(sys.argv) > 1 else 10000, 20000
params = synth_params(G, 0)
counts = DeviceSynth(params).counts(0, C)
X = counts.to_numpy()
keep = (X > 0).sum(0) >= 3
X = np.ascontiguousarray(X[:, keep])
names = np.array([f"g{i}" for i in np.flatnonzero(keep)], dtype=object)
types = cell_types(params, 0, C)
hc = np.random.RandomState(0).rand(C) < 0.6


def run():
    tm = {}

    def tick(name, t0):
        torch.cuda.synchronize()
        tm[name] = (time.perf_counter() - t0) * 1e3
        return time.perf_counter()
    t = time.perf_counter()
    ad = AnnDataLite(X, var_names=names)
    ad.obs['cluster'] = np.array([f"type{k}" for k in types])
    ad.obs['highly_confident'] = hc
    t = tick("adata", t)
    pp.normalize(ad, 'sc', materialize=True)
    t = tick("normalize(+dense residuals to host)", t)
    pp.reduce_dim(ad, 'ps', 0)
    t = tick("reduce_dim (gene + cell PCA)", t)
    find_high_confidence_cells(ad, 8)
    t = tick("confidence (given labels)", t)
    find_relevant_genes(ad, 20, 1, 0)
    t = tick("relevance", t)
    compute_gene_redundancy(ad, 1, 0)
    t = tick("redundancy", t)
    compute_gene_complementarity(ad, 1, 0)
    t = tick("mst+complementarity", t)
    generate_gene_clusters(ad)
    t = tick("hdbscan tail", t)
    sel = tl.select_from_clusters(ad, 'ps', True, 0)
    t = tick("select", t)
    return tm, ad, sel


run()
tm, ad, sel = run()
print({k: round(v, 1) for k, v in tm.items()}, "total", round(sum(tm.values()), 1), "ms;", ad.n_obs, "hc cells,", ad.n_vars,
      "relevant genes,", len(sel), "selected")
