"""NCCL timing of the cell-space panel exchange (developer tool, under torchrun): all-reduce vs reduce-scatter +
all-gather of a (cells x 64) float32 panel."""
import os
import sys

import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for cells in (int(a) for a in (sys.argv[1:] or ["400000", "1000000"])):
    chunk = -(-cells // world)
    full = torch.randn(chunk * world, 64, device="cuda")
    part = torch.empty(chunk, 64, device="cuda")
    res = {}
    for name, fn in (("all_reduce", lambda: dist.all_reduce(full)),
                     ("reduce_scatter", lambda: dist.reduce_scatter_tensor(part, full)),
                     ("all_gather", lambda: dist.all_gather_into_tensor(full, part))):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 20], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[name] = round(float(t), 4)
    if rank == 0:
        mb = full.numel() * 4 / 1e6
        print(f"world {world}, panel {cells} x 64 fp32 ({mb:.0f} MB): ms {res}; all_reduce bus GB/s "
              f"{2 * (world - 1) / world * mb / 1e3 / (res['all_reduce'] * 1e-3):.0f}", flush=True)
dist.destroy_process_group()
