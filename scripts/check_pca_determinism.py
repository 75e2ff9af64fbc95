"""Run-to-run determinism of the (optionally gene-sharded) GeneClust-fast PCA: every raw pass output and every panel
normalisation is checksummed on the device (no extra synchronisation), the checksum lists of the repetitions are compared
at the end and the first call that differs is named.  Stand-alone (one GPU) or under torchrun; `--same-device` = all ranks
on cuda:0 with gloo.

    python scripts/check_pca_determinism.py 1000000 3776 10
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
        scripts/check_pca_determinism.py 1000000 7552 8"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from scgeneclust_b200 import AnnDataLite, pp          # noqa: E402
from scgeneclust_b200 import _device                   # noqa: E402
from scgeneclust_b200.synth import DeviceSynth, synth_params    # noqa: E402

argv = [a for a in sys.argv[1:] if not a.startswith("--")]
same = "--same-device" in sys.argv
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
local = 0 if same else local
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
C, G = int(argv[0]), int(argv[1])
reps = int(argv[2]) if len(argv) > 2 else 4
comm = None
if world > 1:
    from scgeneclust_b200.dist import GeneShardComm
    dist.init_process_group("gloo" if same else "nccl", **({} if same else {"device_id": dev}))
    comm = GeneShardComm(G)
    shard = DeviceSynth(synth_params(G, 0), dev).counts(0, C, comm.gene_lo, comm.local_genes)
    shard.comm = comm
else:
    shard = DeviceSynth(synth_params(G, 0), dev).counts(0, C, 0, G)

MAXC = 256
sums = torch.zeros((reps, MAXC), dtype=torch.float64, device=dev)
labels = []
state = {"rep": 0, "i": 0}


def note(name, t):
    i = state["i"]
    if i < MAXC:
        sums[state["rep"], i] = torch.sum(t, dtype=torch.float64)
        if state["rep"] == 0:
            labels.append(f"{name} {tuple(t.shape)}")
    state["i"] = i + 1


orig_apply = _device.ResidualOperator.apply
orig_chol = _device.PanelAlgebra.chol_qr
orig_gram = _device.PanelAlgebra.gram_of


def apply(self, trans, Q, *a, **k):
    note(f"pass{trans} input", Q)
    out = orig_apply(self, trans, Q, *a, **k)
    note(f"pass{trans} output", out)
    return out


def gram_of(self, P, comm=None, replicated=False):
    g = orig_gram(self, P, comm, replicated)
    note("gram", g)
    return g


def chol_qr(self, P, q, out, comm=None, replicated=False):
    r = orig_chol(self, P, q, out, comm, replicated)
    note("rinv", self.rinv)
    note("chol_qr out", out)
    return r


_device.ResidualOperator.apply = apply
_device.PanelAlgebra.chol_qr = chol_qr
_device.PanelAlgebra.gram_of = gram_of
names = np.array([f"g{i}" for i in range(G)], dtype=object)
embs = []
for rep in range(reps):
    state["rep"], state["i"] = rep, 0
    ad = AnnDataLite(shard, var_names=names)
    pp.start_omega(ad, 0)
    pp.normalize(ad, 'sc')
    st = pp.get_state(ad)
    if st.omega is not None:
        torch.cuda.current_stream().wait_stream(st.omega.stream)
        note("omega", st.omega.panel)
    try:
        pp.reduce_dim(ad, 'fast', 0)
        embs.append(np.asarray(ad.varm['X_pca']).copy())
    except Exception as e:            # noqa: BLE001
        print(f"[rank {rank}] rep {rep}: PCA failed: {e}", flush=True)
        embs.append(None)
torch.cuda.synchronize()
S = sums.cpu().numpy()
n = min(state["i"], MAXC)
for rep in range(1, reps):
    diff = [i for i in range(n) if S[rep, i] != S[0, i] and not (np.isnan(S[rep, i]) and np.isnan(S[0, i]))]
    same_e = embs[rep] is not None and embs[0] is not None and np.array_equal(embs[rep], embs[0])
    if diff or not same_e:
        i = diff[0] if diff else -1
        what = labels[i] if 0 <= i < len(labels) else "?"
        print(f"[rank {rank}] rep {rep} DIFFERS from rep 0: embedding identical {same_e}; first differing call #{i} ({what}): "
              f"{S[rep, i]!r} vs {S[0, i]!r}; {len(diff)} of {n} checksums differ", flush=True)
    elif rank == 0:
        print(f"[rank {rank}] rep {rep}: identical ({n} checksums)", flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
