"""cProfile of the public entry point on config c3 (developer tool): device-resident and from pinned host counts."""
import cProfile
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from scgeneclust_b200 import AnnDataLite, scGeneClust
from scgeneclust_b200.synth import DeviceSynth, synth_params

C, G, K = int(sys.argv[1]) if len(sys.argv) > 1 else 50000, 20000, 200
counts = DeviceSynth(synth_params(G, 0)).counts(0, C)
names = np.array([f"g{i}" for i in range(G)], dtype=object)
host = torch.empty((C, G), dtype=torch.int16).pin_memory()
host.copy_(counts.data[:, :G])
torch.cuda.synchronize()


def resident():
    return scGeneClust(AnnDataLite(counts, var_names=names), n_var_clusters=K, version='fast', verbosity=0)


def e2e():
    return scGeneClust(AnnDataLite(host.numpy().view(np.uint16), var_names=names), n_var_clusters=K, version='fast',
                       verbosity=0)


for name, fn in (("resident", resident), ("e2e", e2e)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    print(f"{name}: {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms per call")
    pr = cProfile.Profile()
    pr.enable()
    fn()
    torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(25)
