"""Fine-grained wall-clock breakdown of the cluster / select stages on a real embedding (developer tool)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from scgeneclust_b200 import AnnDataLite, pp, tl
from scgeneclust_b200.pp._preprocessing import get_state
from scgeneclust_b200.synth import DeviceSynth, synth_params
from scgeneclust_b200.tl import cluster as C

Cn, G, K = int(sys.argv[1]) if len(sys.argv) > 1 else 50000, 20000, 200
counts = DeviceSynth(synth_params(G, 0)).counts(0, Cn)
names = np.array([f"g{i}" for i in range(G)], dtype=object)
ad = AnnDataLite(counts, var_names=names)
pp.normalize(ad, 'sc')
pp.reduce_dim(ad, 'fast', 0)
X = get_state(ad).varm_pca


def sync():
    torch.cuda.synchronize()
    return time.perf_counter()


for rep in range(3):
    tm = {}
    km = C.DeviceMiniBatchKMeans(n_clusters=K, batch_size=max(1024, G // 10), random_state=0)
    rs = np.random.RandomState(0)
    batch = min(km.batch_size, G)
    t0 = sync()
    rs.randint(0, G, 3 * batch)
    centers = km._init_centroids(X, rs, 3 * batch)
    t1 = sync()
    tm['kpp_init'] = (t1 - t0) * 1e3
    t0 = sync()
    labels = km.fit_predict(X)
    t1 = sync()
    tm['fit_predict_total'] = (t1 - t0) * 1e3
    tm['n_steps'] = km.n_steps_
    t0 = sync()
    clo = C.compute_gene_closeness(X, labels, km.cluster_centers_)
    t1 = sync()
    tm['closeness'] = (t1 - t0) * 1e3
    t0 = sync()
    ad.var['cluster'] = labels.cpu().numpy()
    ad.var['closeness'] = clo.cpu().numpy()
    t1 = sync()
    tm['d2h_var'] = (t1 - t0) * 1e3
    t0 = sync()
    sel = tl.select_from_clusters(ad, 'fast', True, 0)
    t1 = sync()
    tm['select'] = (t1 - t0) * 1e3
    cl = ad.var['cluster'].to_numpy()
    tm['singletons'] = int((np.bincount(cl) == 1).sum())
    print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in tm.items()}, flush=True)
import cProfile
import pstats
pr = cProfile.Profile()
pr.enable()
tl.cluster_genes(ad, None, 'fast', n_gene_clusters=K, random_state=0)
sel = tl.select_from_clusters(ad, 'fast', True, 0)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(25)
