"""What the two collectives of one sharded pair cost on this box (NCCL through torch.distributed, same NCCL build the library
dlopens): an all-gather of the match slots and an 8-byte max all-reduce, timed with CUDA events after a device-side barrier
(a preceding all-reduce), 200 repetitions.  torchrun --nproc-per-node N scripts/comm_probe.py"""
import os

import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
slot = (100000 // world + 2) * 16
send = torch.zeros(slot, dtype=torch.uint8, device="cuda")
recv = torch.zeros(slot * world, dtype=torch.uint8, device="cuda")
word = torch.zeros(1, dtype=torch.int64, device="cuda")
sync = torch.zeros(1, dtype=torch.int32, device="cuda")
ev = lambda: torch.cuda.Event(enable_timing=True)
for name, fn in (("all_gather %d B per rank" % slot, lambda: dist.all_gather_into_tensor(recv, send)),
                 ("all_reduce max 8 B", lambda: dist.all_reduce(word, op=dist.ReduceOp.MAX))):
    ts = []
    for it in range(220):
        dist.all_reduce(sync)                     # ranks aligned on the device
        a, b = ev(), ev()
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        if it >= 20:
            ts.append(a.elapsed_time(b) * 1e3)
    t = torch.tensor([sorted(ts)[len(ts) // 2], max(ts)], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("N=%d %s: median %.1f us, max %.1f us (max over ranks)" % (world, name, t[0].item(), t[1].item()), flush=True)
dist.destroy_process_group()
