"""Writes tests/golden/c3_fast_oracle.npz: the CPU oracle's GeneClust-fast result at BASELINE.json config c3
(synth(50 000, 20 000, seed 0), n_var_clusters = 200, random_state = 0).

    python tests/golden/make_c3_oracle.py          # ~50 s and ~22 GB of host memory on 8 cores

The fixture pins the full-size GPU run (tests/test_gpu_fullsize.py) without a 50 s CPU run on the GPU box.  It is the
ORACLE's output (oracle/stages.py: numpy Pearson residuals + the installed scikit-learn 1.9.0 PCA / MiniBatchKMeans /
IsolationForest, i.e. the calls the reference makes), not the reference package's: scanpy / anndata are absent, so the
reference's `pp.normalize` / `pp.reduce_dim` cannot run here (a1 is "parity unpinned", see oracle/__init__.py).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import stages  # noqa: E402
from oracle import synth as osy  # noqa: E402

C, G, K, SEED = 50_000, 20_000, 200, 0
X = osy.synth_counts_c(osy.synth_params(G, SEED), 0, C, 0, G)
names = np.array([f"g{i}" for i in range(G)], dtype=object)
tm = {}
res = stages.geneclust_fast(X, names, K, SEED, timings=tm)
print(tm)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "c3_fast_oracle.npz"),
                    X_pca=res['X_pca'].astype(np.float32), labels=res['labels'].astype(np.int32),
                    closeness=res['closeness'].astype(np.float32), selected=res['selected'].astype(str),
                    cells=C, genes=G, n_var_clusters=K, seed=SEED)
