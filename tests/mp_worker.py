"""One rank of the multi-process parity check (launched by tests/test_gpu_multi.py through torch.distributed.run):
the library's own clique (erp_comm_init; peer windows mapped with cudaIpc, or NCCL with ERP_B200_PEER=0) splits ONE
pair; rank 0 writes what it got so that the parent can compare it with the single-GPU result and the oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import erp_match_eightpoint_test_b200 as erp  # noqa: E402
from test_gpu_multi import scene_pair  # noqa: E402

out_path = sys.argv[1]
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")                       # plumbing only: the 128-byte id travels over gloo
box = [erp.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(box, src=0)
ctx = erp.Context(local)
ctx.comm_init(world, rank, box[0])
assert ctx.comm_size == world and ctx.comm_rank == rank
res = {}
for tag, (nq, nt, H, cross, seed) in {"tc": (6001, 7003, 60001, False, 41), "cross": (6001, 7003, 60001, True, 41), "simt": (1501, 1999, 2049, False, 43)}.items():
    q, t, left, right = scene_pair(nq, nt, 64, 4096, 2048, seed)
    for rep in range(3):                              # direct call, captured call, graph replay: all three must agree
        m, r = ctx.pair_pose_dist(q, t, left, right, 4096, 2048, ratio=0.3, cross_check=cross, seed=7, H=H)
        key = (m.tobytes(), r["packed"], r["mask"].tobytes(), r["E_refit"].tobytes())
        if rep == 0:
            first = key
        assert key == first, (tag, rep)
    part = ctx.knn2_match_dist(q, t, 0.3, cross)
    lo, hi = erp.shard_range(nq, rank, world)
    assert ((part["queryIdx"] >= lo) & (part["queryIdx"] < hi)).all()
    parts = [None] * world
    dist.all_gather_object(parts, part)
    if rank == 0:
        res[tag + "_matches"] = m
        res[tag + "_packed"] = np.uint64(r["packed"])
        res[tag + "_mask"] = r["mask"]
        res[tag + "_E_refit"] = r["E_refit"]
        res[tag + "_parts"] = np.concatenate(parts)
if rank == 0:
    np.savez(out_path, **res)
ctx.close()
dist.destroy_process_group()
