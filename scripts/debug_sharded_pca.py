"""Where does a gene-sharded PCA go wrong?  Runs the sharded GeneClust-fast PCA with the Cholesky status checked after every
panel normalisation (rank 0 prints the Gram diagonal range).  Under torchrun; `--same-device` = all ranks on cuda:0, gloo.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
        scripts/debug_sharded_pca.py 1000000 30000 --same-device"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from scgeneclust_b200 import AnnDataLite, pp          # noqa: E402
from scgeneclust_b200 import _device                   # noqa: E402
from scgeneclust_b200.dist import GeneShardComm         # noqa: E402
from scgeneclust_b200.synth import DeviceSynth, synth_params    # noqa: E402

argv = [a for a in sys.argv[1:] if not a.startswith("--")]
same = "--same-device" in sys.argv
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
local = 0 if same else local
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("gloo" if same else "nccl", **({} if same else {"device_id": dev}))
C, G = int(argv[0]), int(argv[1])
comm = GeneShardComm(G)
shard = DeviceSynth(synth_params(G, 0), dev).counts(0, C, comm.gene_lo, comm.local_genes)
shard.comm = comm
calls = [0]
orig = _device.PanelAlgebra.chol_qr


def checked(self, P, q, out, comm=None, replicated=False):
    r = orig(self, P, q, out, comm, replicated)
    calls[0] += 1
    d = torch.diagonal(self.gram)[:q]
    bad = int(self.status.item())
    finite = bool(torch.isfinite(P).all().item())
    if bad or not finite:
        print(f"[rank {rank}] chol_qr #{calls[0]} rows {P.shape[0]} replicated={replicated} status {bad} finite_in {finite} "
              f"gram diag min {d.min().item():.3e} max {d.max().item():.3e} out finite {bool(torch.isfinite(out).all().item())}",
              flush=True)
    return r


_device.PanelAlgebra.chol_qr = checked
reps = int(argv[2]) if len(argv) > 2 else 1
names = np.array([f"g{i}" for i in range(G)], dtype=object)
ref = None
for rep in range(reps):
    calls[0] = 0
    ad = AnnDataLite(shard, var_names=names)
    pp.start_omega(ad, 0)
    pp.normalize(ad, 'sc')
    try:
        pp.reduce_dim(ad, 'fast', 0)
        E = np.asarray(ad.varm['X_pca'])
        if ref is None:
            ref = E.copy()
        same = bool(np.array_equal(E, ref))
        if rank == 0 or not same:
            print(f"[rank {rank}] rep {rep}: PCA ok, finite {bool(np.isfinite(E).all())}, identical to rep 0: {same}", flush=True)
    except Exception as e:            # noqa: BLE001
        print(f"[rank {rank}] rep {rep}: PCA failed: {e}", flush=True)
dist.barrier()
dist.destroy_process_group()
