"""GeneClust-ps through the entry point on gene-sharded input (N ranks) vs the single-GPU run of the same matrix
(acceptance check under torchrun; wrapped by tests/test_gpu_dist.py): identical selected-gene list.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
        scripts/check_sharded_ps.py [cells] [genes] [--backend nccl|gloo] [--same-device]"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from scgeneclust_b200 import AnnDataLite, scGeneClust
from scgeneclust_b200._device import HostShard
from scgeneclust_b200.dist import GeneShardComm
from scgeneclust_b200.synth import DeviceSynth, synth_cell_types, synth_params

argv = [a for a in sys.argv[1:] if not a.startswith("--")]
backend = sys.argv[sys.argv.index("--backend") + 1] if "--backend" in sys.argv else "nccl"
if "--backend" in sys.argv:
    argv.remove(backend)
same_device = "--same-device" in sys.argv
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
local = 0 if same_device else local
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if backend == "nccl":
    dist.init_process_group("nccl", device_id=dev)
else:
    dist.init_process_group("gloo")
C, G = int(argv[0]) if len(argv) > 0 else 3000, int(argv[1]) if len(argv) > 1 else 2500
params = synth_params(G, 0)
full = DeviceSynth(params, dev).counts(0, C).to_numpy()
keep = (full > 0).sum(0) >= 3
full = np.ascontiguousarray(full[:, keep])
names = np.array([f"g{i}" for i in np.flatnonzero(keep)], dtype=object)
types = np.array([f"type{t}" for t in synth_cell_types(params, 0, C)])
hc = np.random.RandomState(0).rand(C) < 0.6
comm = GeneShardComm(full.shape[1])


def adata(X):
    ad = AnnDataLite(X, var_names=names)
    ad.obs['cluster'] = types
    ad.obs['highly_confident'] = hc
    return ad


sel_sharded = scGeneClust(adata(HostShard(np.ascontiguousarray(full[:, comm.gene_lo:comm.gene_hi]), comm)), version='ps',
                          n_obs_clusters=8, verbosity=0)
# the same with return_info (dense varp['redundancy'] assembled from the ranks' pair ranges, residuals gathered)
info, sel_info = scGeneClust(adata(HostShard(np.ascontiguousarray(full[:, comm.gene_lo:comm.gene_hi]), comm)),
                             version='ps', n_obs_clusters=8, return_info=True, verbosity=0)
ok = True
if rank == 0:
    sel_single = scGeneClust(adata(full), version='ps', n_obs_clusters=8, verbosity=0)
    same = np.array_equal(sel_sharded, sel_single)
    same_info = np.array_equal(sel_info, sel_single)
    R = info.varp['redundancy']
    print(f"world {world} ({backend}): GeneClust-ps selected list identical to single GPU: {same} ({len(sel_single)} genes); "
          f"with return_info: {same_info}, redundancy {R.shape} symmetric {np.array_equal(R, R.T)}, X {info.X.shape}", flush=True)
    ok = bool(same and same_info and np.array_equal(R, R.T))
flag = torch.tensor([int(ok)], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) else 1)
