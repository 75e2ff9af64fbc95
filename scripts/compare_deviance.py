"""Bitwise comparison of gc_binomial_deviance between two builds of the library (developer tool):
    python scripts/compare_deviance.py dump out.npz   (honours GENECLUST_B200_LIB)
    python scripts/compare_deviance.py diff a.npz b.npz"""
import sys
import time

import numpy as np

if sys.argv[1] == "diff":
    a, b = np.load(sys.argv[2]), np.load(sys.argv[3])
    ok = True
    for k in a.files:
        if k.startswith("ms_"):
            print(k, round(float(a[k]), 3), "->", round(float(b[k]), 3))
            continue
        same = np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32))
        ok &= same
        print(k, "identical" if same else "DIFFERENT")
    sys.exit(0 if ok else 1)

import torch

sys.path.insert(0, ".")
from scgeneclust_b200.tl.selection import device_deviance

out = {}
rs = np.random.RandomState(3)
SHAPES = ((50000, 3, 0), (100000, 6, 0), (1000, 1, 0), (777, 40, 0), (5000, 300, 0), (4099, 5, 1), (400000, 6, 0))
if len(sys.argv) > 3 and sys.argv[3] == "big":
    SHAPES = ((50000, 3, 0), (400000, 6, 0))
for C, m, off in SHAPES:
    X = np.abs(rs.standard_normal((C, m))).astype(np.float32) * rs.choice([0.0, 1.0, 3.0], size=(C, m)).astype(np.float32)
    flat = torch.zeros(C * m + 8, dtype=torch.float32, device="cuda")
    Xd = flat[off:off + C * m].view(C, m)
    Xd.copy_(torch.from_numpy(X))
    device_deviance(Xd)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    d = device_deviance(Xd)
    out[f"ms_{C}x{m}"] = (time.perf_counter() - t0) * 1e3
    out[f"dev_{C}x{m}_{off}"] = d
np.savez(sys.argv[2], **out)
print({k: round(float(v), 3) for k, v in out.items() if k.startswith("ms_")})
