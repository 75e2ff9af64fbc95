"""Gene-PCA accuracy vs the slab-length cap of the pass kernel (developer tool; the cap is read from the environment once
per process: GC_RSVD_MAX_STAGES=n python scripts/pca_accuracy.py <case> [save.npz | compare.npz])
   case c3   : 50 000 x 20 000 vs tests/golden/c3_fast_oracle.npz (CPU oracle)
   case sub  : 20 000 x 20 000 vs the CPU oracle (computed once, cached in /tmp)
   case big  : 400 000 x 20 000 - saves the embedding to / compares it with the given .npz (cap-vs-cap agreement)"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from scgeneclust_b200._device import ResidualOperator, randomized_pca
from scgeneclust_b200.synth import DeviceSynth, synth_params

case = sys.argv[1]
cap = os.environ.get("GC_RSVD_MAX_STAGES", "default")
C, G = {"c3": (50000, 20000), "sub": (20000, 20000), "big": (400000, 20000)}[case]
counts = DeviceSynth(synth_params(G, 0)).counts(0, C)
op = ResidualOperator.from_counts(counts)
for _ in range(2):
    emb = randomized_pca(op, 'genes', 0)
torch.cuda.synchronize()
t0 = time.perf_counter()
emb = randomized_pca(op, 'genes', 0)
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) * 1e3
emb = emb.cpu().numpy()


def rel(a, b):
    r = np.linalg.norm(a - b, axis=0) / np.linalg.norm(b, axis=0)
    return f"max {r.max():.2e} first10 {r[:10].max():.2e} median {np.median(r):.2e}"


if case == "c3":
    ref = np.load("tests/golden/c3_fast_oracle.npz")["X_pca"]
    print(f"cap {cap}: c3 PCA {ms:.2f} ms, vs CPU oracle: {rel(emb, ref)}", flush=True)
elif case == "sub":
    cache = "/tmp/sub_cpu_pca.npy"
    if not os.path.exists(cache):
        from oracle import stages
        from oracle import synth as osy
        X = osy.synth_counts_c(osy.synth_params(G, 0), 0, C, 0, G)
        np.save(cache, stages.gene_pca(stages.pearson_residuals(X.astype(np.float32)), 0))
    print(f"cap {cap}: 20k x 20k PCA {ms:.2f} ms, vs CPU oracle: {rel(emb, np.load(cache))}", flush=True)
else:
    path = sys.argv[2]
    if os.path.exists(path):
        print(f"cap {cap}: 400k x 20k PCA {ms:.2f} ms, vs {path}: {rel(emb, np.load(path))}", flush=True)
    else:
        np.save(path, emb)
        print(f"cap {cap}: 400k x 20k PCA {ms:.2f} ms, saved to {path}", flush=True)
