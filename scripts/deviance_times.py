"""Time of the singleton rescue's binomial deviance on residual-like columns.  python scripts/deviance_times.py [C m]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from scgeneclust_b200.tl.selection import device_deviance      # noqa: E402

C = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 7
rs = np.random.RandomState(0)
X = (rs.standard_normal((C, m)) * rs.uniform(0.5, 3, size=m) + rs.uniform(-0.2, 0.4, size=m)).astype(np.float32)
X[rs.rand(C, m) < 0.3] = np.float32(-0.31)
Xd = torch.from_numpy(X).cuda()
for _ in range(3):
    device_deviance(Xd)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    device_deviance(Xd)
e1.record()
torch.cuda.synchronize()
print(f"deviance C={C} m={m}: {e0.elapsed_time(e1) / 10:.3f} ms per call")
from torch.profiler import ProfilerActivity, profile     # noqa: E402
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    device_deviance(Xd)
    torch.cuda.synchronize()
for ev in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:10]:
    print(f"  {ev.device_time_total / 1e3:8.3f} ms {ev.count:3d} x {ev.key[:90]}")
