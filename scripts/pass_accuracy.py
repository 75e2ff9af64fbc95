"""Accuracy of the randomized-SVD pass kernels against a float64 product with the materialised residual matrix
(developer tool).   [GC_RSVD_SPLIT=n] python scripts/pass_accuracy.py [cells] [genes]"""
import os
import sys

import torch

sys.path.insert(0, ".")
from scgeneclust_b200._device import Q_LD, ResidualOperator
from scgeneclust_b200.synth import DeviceSynth, synth_params

C_, G_ = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (50000, 20000)
counts = DeviceSynth(synth_params(G_, 0)).counts(0, C_)
op = ResidualOperator.from_counts(counts)
dev = counts.device
torch.manual_seed(0)
Q = {0: torch.zeros(G_, Q_LD, device=dev), 1: torch.zeros(C_, Q_LD, device=dev)}
for t in (0, 1):
    Q[t][:, :60] = torch.randn(Q[t].shape[0], 60, device=dev)
Z = op.materialize()
ref = {}
rows = 5000
acc0 = torch.empty(C_, Q_LD, dtype=torch.float64, device=dev)
acc1 = torch.zeros(G_, Q_LD, dtype=torch.float64, device=dev)
for r0 in range(0, C_, rows):
    blk = Z[r0:r0 + rows].double()
    acc0[r0:r0 + rows] = blk @ Q[0].double()
    acc1 += blk.t() @ Q[1][r0:r0 + rows].double()
ref = {0: acc0, 1: acc1}
del Z
for kernel in ("ts", "ss", "simt"):
    out = {}
    for trans in (0, 1):
        y = op.apply(trans, Q[trans], kernel=kernel).double()
        d = y - ref[trans]
        out[f"fro{trans}"] = float(d.norm() / ref[trans].norm())
        out[f"max{trans}"] = float(d.abs().max() / ref[trans].abs().max())
        out[f"bias{trans}"] = float((d * ref[trans].sign()).sum() / ref[trans].abs().sum())   # > 0: magnitudes too large
    print(os.environ.get("GC_RSVD_SPLIT", "auto"), kernel, (C_, G_), {k: f"{v:.2e}" for k, v in out.items()}, flush=True)
