"""Host-side profile of one GeneClust-fast selection with the counts resident on the device (where does the wall time
go that the GPU kernels do not account for?).  python scripts/host_profile.py [cells genes]"""
import cProfile
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import scgeneclust_b200 as gc                                    # noqa: E402
from scgeneclust_b200.synth import DeviceSynth, synth_params      # noqa: E402
from scgeneclust_b200._anndata_lite import AnnDataLite             # noqa: E402

C = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
G = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
counts = DeviceSynth(synth_params(G, 0), torch.device("cuda")).counts(0, C, 0, G)
names = np.array([f"g{i}" for i in range(G)], dtype=object)


def step():
    ad = AnnDataLite(counts, var_names=names)
    return gc.scGeneClust(ad, version='fast', n_var_clusters=200, verbosity=0)


for _ in range(3):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    step()
torch.cuda.synchronize()
print(f"step {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms")
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(35)
st.sort_stats("cumulative").print_stats(45)
