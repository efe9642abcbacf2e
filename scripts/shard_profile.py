"""The work of ONE rank of an N-way split of cfg3 on one GPU (the ncu target for the strong-scaling shape):
    python scripts/shard_profile.py [N=8] [REPS=3]
match: nq/N query rows x all train rows (erp_knn2_match_dev); pose: H/N hypotheses over ALL ~50k correspondences.
Prints device times of the two parts; N = 1 is the whole pair through erp_pair_pose_dev."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import bench
import erp_match_eightpoint_test_b200 as erp

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg = bench.WORKLOADS["cfg3"]
pair = bench.make_pair(cfg, 0xE8B0 + 3)
dev = torch.device("cuda", 0)
ctx = erp.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
nq, nt, dim = cfg["nq"], cfg["nt"], cfg["dim"]
lo, hi = erp.shard_range(nq, 0, N)
d_q, d_t = torch.from_numpy(pair["q"]).to(dev), torch.from_numpy(pair["t"]).to(dev)
d_left, d_right = torch.from_numpy(pair["left"]).to(dev), torch.from_numpy(pair["right"]).to(dev)
d_m = torch.empty((nq, 4), dtype=torch.int32, device=dev)
d_n = torch.zeros(1, dtype=torch.int32, device=dev)
d_mask = torch.empty(nq, dtype=torch.uint8, device=dev)
d_res = torch.zeros(216, dtype=torch.uint8, device=dev)
# the full match list (what the all-gather delivers on every rank)
ctx.knn2_match_dev(d_q, nq, d_t, nt, dim, 0.3, False, d_m, d_n)
ctx.synchronize()
m = int(d_n.item())
d_full = d_m[:m].clone()
d_l3, d_r3 = torch.empty((m, 3), dtype=torch.float64, device=dev), torch.empty((m, 3), dtype=torch.float64, device=dev)
d_l4, d_r4 = torch.empty((m, 4), dtype=torch.float32, device=dev), torch.empty((m, 4), dtype=torch.float32, device=dev)
d_packed = torch.zeros(1, dtype=torch.int64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
H = cfg["hyps"] // N
ev = lambda: torch.cuda.Event(enable_timing=True)
for it in range(reps):
    with torch.cuda.stream(stream):
        flush.zero_()
        e0, e1, e2 = ev(), ev(), ev()
        e0.record(stream)
        ctx.knn2_match_dev(d_q[lo:hi], hi - lo, d_t, nt, dim, 0.3, False, d_m, d_n)
        e1.record(stream)
        ctx.gather_bearings_dev(d_full, m, d_left, d_right, 8, 0, cfg["W"], cfg["H"], d_l3, d_r3, d_l4, d_r4)
        ctx.ransac_local_dev(d_l3, d_r3, d_l4, d_r4, m, 1, 0, H, 8, 0, 0.002, d_packed)
        stream.synchronize()
        res = ctx.ransac_finish_dev(d_l3, d_r3, d_l4, d_r4, m, 1, int(d_packed.item()), 8, 0, 0.002)
        e2.record(stream)
    torch.cuda.synchronize()
    print(f"N={N}: match {e0.elapsed_time(e1):.3f} ms (kernel {ctx.last_knn_kernel_ms():.3f}), pose {e1.elapsed_time(e2):.3f} ms "
          f"(score kernels {ctx.last_score_kernel_ms()[0]:.3f}), inliers {res['count']}")
ctx.close()
