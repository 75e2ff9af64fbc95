"""Gene-sharded GeneClust-fast on synth(cells, genes) under torchrun; rank 0 saves the embedding, labels, closeness and
selected list (npz) for offline comparison with scikit-learn's tail on the same embedding.
    torchrun --nproc-per-node 8 scripts/dump_sharded_fast.py 20000 30000 out.npz [--same-device]"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from scgeneclust_b200 import AnnDataLite, pp, tl          # noqa: E402
from scgeneclust_b200.dist import GeneShardComm             # noqa: E402
from scgeneclust_b200.synth import DeviceSynth, synth_params    # noqa: E402

argv = [a for a in sys.argv[1:] if not a.startswith("--")]
same = "--same-device" in sys.argv
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
local = 0 if same else local
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("gloo" if same else "nccl", **({} if same else {"device_id": dev}))
C, G, out = int(argv[0]), int(argv[1]), argv[2]
K = int(argv[3]) if len(argv) > 3 else 200
comm = GeneShardComm(G)
shard = DeviceSynth(synth_params(G, 0), dev).counts(0, C, comm.gene_lo, comm.local_genes)
shard.comm = comm
names = np.array([f"g{i}" for i in range(G)], dtype=object)
ad = AnnDataLite(shard, var_names=names)
pp.normalize(ad, 'sc')
pp.reduce_dim(ad, 'fast', 0)
tl.cluster_genes(ad, None, 'fast', n_gene_clusters=K, random_state=0)
sel = tl.select_from_clusters(ad, 'fast', True, 0)
st = pp.get_state(ad)
if rank == 0:
    np.savez(out, X_pca=np.asarray(ad.varm['X_pca']), labels=ad.var['cluster'].to_numpy(),
             closeness=ad.var['closeness'].to_numpy(), selected=np.asarray(sel).astype(str),
             centers=st.info['kmeans']['centers'].cpu().numpy(), n_steps=st.info['kmeans']['n_steps'])
    print("saved", out, flush=True)
dist.barrier()
dist.destroy_process_group()
