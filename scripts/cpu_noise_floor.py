"""How closely does the CPU reference agree with ITSELF?  (evidence for the parity bars of tests/test_gpu_subsample.py)

    python scripts/cpu_noise_floor.py          # ~2 min on 8 cores

1. The oracle's GeneClust-fast on synth(20 000, 20 000) run with two BLAS thread counts: PCA scores differ by ~5e-4 in
   the trailing components (1e-5 in the first ten) - float32 GEMM summation order alone.
2. scikit-learn's MiniBatchKMeans on the oracle's embedding perturbed by a relative 1e-7 ... 1e-4: from 1e-5 on, some
   of the ~1 400 k-means++ draws fall on another sample and the clustering of the (mostly unstructured) synthetic genes
   changes completely (ARI ~0.1).  Identical labels between two float32 PCA implementations are therefore a coin flip,
   whichever implementations they are.
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
C, G, K = 20_000, 20_000, 200
_shape = [a for a in sys.argv[1:] if a.isdigit()]
if len(_shape) >= 2:                                   # python scripts/cpu_noise_floor.py 20000 30000
    C, G = int(_shape[0]), int(_shape[1])

if len(sys.argv) > 1 and sys.argv[1] == "--worker":
    from oracle import stages
    from oracle import synth as osy
    X = osy.synth_counts_c(osy.synth_params(G, 0), 0, C, 0, G)
    names = np.array([f"g{i}" for i in range(G)], dtype=object)
    r = stages.geneclust_fast(X, names, K, 0)
    # and the clustering of the OTHER run's embedding with this run's BLAS thread count (argv[5], optional): is the tail
    # alone thread-count independent?
    extra = {}
    if len(sys.argv) > 5 and os.path.exists(sys.argv[5]):
        E_other = np.load(sys.argv[5])['X_pca']
        extra['labels_on_other'] = stages.kmeans_genes(E_other, K, 0)[0]
    np.savez(sys.argv[2], X_pca=r['X_pca'], labels=r['labels'], selected=r['selected'].astype(str), **extra)
    sys.exit(0)

from sklearn.metrics import adjusted_rand_score  # noqa: E402

from oracle import stages  # noqa: E402

out = {}
for threads in ("8", "3"):
    path = f"/tmp/cpu_noise_{threads}.npz"
    env = dict(os.environ, OMP_NUM_THREADS=threads, OPENBLAS_NUM_THREADS=threads)
    subprocess.run([sys.executable, __file__, "--worker", path, str(C), str(G), "/tmp/cpu_noise_8.npz"], check=True, env=env)
    out[threads] = np.load(path)
a, b = out["8"], out["3"]
rel = np.linalg.norm(a['X_pca'] - b['X_pca'], axis=0) / np.linalg.norm(b['X_pca'], axis=0)
print(f"CPU 8 vs 3 threads: PCA rel err max {rel.max():.2e}, first 10 PCs {rel[:10].max():.2e}, labels equal "
      f"{np.array_equal(a['labels'], b['labels'])}, ARI {adjusted_rand_score(a['labels'], b['labels']):.3f}")
print(f"selected lists equal {np.array_equal(a['selected'], b['selected'])} ({len(a['selected'])} / {len(b['selected'])} genes)")
if 'labels_on_other' in b.files:
    print(f"k-means with 3 BLAS threads on the 8-thread run's embedding: labels identical to the 8-thread run's "
          f"{np.array_equal(b['labels_on_other'], a['labels'])}, ARI {adjusted_rand_score(b['labels_on_other'], a['labels']):.4f}")
E, lab0 = a['X_pca'], a['labels']
rs = np.random.RandomState(1)
for eps in (1e-7, 1e-6, 1e-5, 1e-4):
    res = []
    for _ in range(4):
        lab, _, _ = stages.kmeans_genes((E * (1 + eps * rs.standard_normal(E.shape))).astype(np.float32), K, 0)
        res.append((bool(np.array_equal(lab, lab0)), round(float(adjusted_rand_score(lab, lab0)), 3)))
    print(f"sklearn k-means on the embedding perturbed by {eps:g}: (labels identical, ARI) = {res}")
