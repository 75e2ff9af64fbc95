"""Time the two directions of the randomized-SVD pass kernel separately (developer tool).
    GENECLUST_B200_LIB=<variant .so> python scripts/pass_times.py [cells] [genes]"""
import os
import sys

import torch

sys.path.insert(0, ".")
from scgeneclust_b200._device import Q_LD, ResidualOperator
from scgeneclust_b200._lib import lib
from scgeneclust_b200.synth import DeviceSynth, synth_params

C_, G_ = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (50000, 20000)
counts = DeviceSynth(synth_params(G_, 0)).counts(0, C_)
op = ResidualOperator.from_counts(counts)
dev = counts.device
torch.manual_seed(0)
Q = {0: torch.randn(G_, Q_LD, device=dev), 1: torch.randn(C_, Q_LD, device=dev)}
out = {}
for trans in (0, 1):
    for _ in range(3):
        y = op.apply(trans, Q[trans])
    lib().profile_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        y = op.apply(trans, Q[trans])
    e1.record()
    torch.cuda.synchronize()
    out[trans] = e0.elapsed_time(e1) / 20
    out[f"chk{trans}"] = float(y.double().abs().sum().item())
print(os.environ.get("GENECLUST_B200_LIB", "in-tree"), {k: (round(v, 4) if isinstance(k, int) else v) for k, v in out.items()})
