"""Time the two directions of the randomized-SVD pass for each tensor-core kernel (developer tool).
    [GENECLUST_B200_LIB=<variant .so>] python scripts/pass_times.py [cells] [genes] [kernels, e.g. ts,ss]"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, ".")
from scgeneclust_b200._device import Q_LD, ResidualOperator
from scgeneclust_b200._lib import lib
from scgeneclust_b200.synth import DeviceSynth, synth_params

C_, G_ = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (50000, 20000)
kernels = sys.argv[3].split(",") if len(sys.argv) > 3 else ["ts", "ss"]
counts = DeviceSynth(synth_params(G_, 0)).counts(0, C_)
op = ResidualOperator.from_counts(counts)
dev = counts.device
torch.manual_seed(0)
Q = {0: torch.randn(G_, Q_LD, device=dev), 1: torch.randn(C_, Q_LD, device=dev)}
L = lib()
for kernel in kernels:
    out = {}
    ref = {}
    for trans in (0, 1):
        for _ in range(3):
            y = op.apply(trans, Q[trans], kernel=kernel)
        tot, cnt = ctypes.c_double(), ctypes.c_longlong()
        L.profile_enable(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            y = op.apply(trans, Q[trans], kernel=kernel)
        e1.record()
        torch.cuda.synchronize()
        L.profile_enable(0)
        L.profile_collect(0, ctypes.byref(tot), ctypes.byref(cnt))
        out[f"apply{trans}_ms"] = round(e0.elapsed_time(e1) / 20, 4)
        out[f"kernel{trans}_ms"] = round(tot.value / max(cnt.value, 1), 4)
        out[f"chk{trans}"] = float(y.double().abs().sum().item())
        ys = op.apply(trans, Q[trans], kernel="simt")
        out[f"vs_simt{trans}"] = float((ys - y).abs().max() / ys.abs().max())
    gbs = 2.0 * C_ * G_ / 1e9
    out["frac_hbm_6547"] = round(gbs / (0.5 * (out["kernel0_ms"] + out["kernel1_ms"]) * 1e-3) / 6547.2, 3)
    print(os.environ.get("GENECLUST_B200_LIB", "in-tree"), kernel, (C_, G_), out, flush=True)
