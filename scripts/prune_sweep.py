"""Sweep of the pass-B extent of the tensor-core best-hypothesis search (exact for every setting: only the split of the
work between pass B and pass C moves).  One process per setting (the knobs are read once)."""
import os
import subprocess
import sys

if len(sys.argv) > 1:
    import numpy as np
    import torch
    sys.path.insert(0, ".")
    import erp_match_eightpoint_test_b200 as erp
    import oracle as O
    from erp_match_eightpoint_test_b200 import synth
    H, m = int(sys.argv[1]), 50000
    kp = synth.keypoint_pair(m, 8192, 4096, seed=5)
    l, r = O.bearings(kp["left_xy"], 8192, 4096), O.bearings(kp["right_xy"], 8192, 4096)
    ctx = erp.Context(0)
    dev = torch.device("cuda", 0)
    d_l3, d_r3 = torch.from_numpy(l).to(dev), torch.from_numpy(r).to(dev)
    d_l4 = torch.empty((m, 4), dtype=torch.float32, device=dev); d_r4 = torch.empty_like(d_l4)
    ctx.pack_float4_dev(d_l3, m, d_l4); ctx.pack_float4_dev(d_r3, m, d_r4)
    d_packed = torch.zeros(1, dtype=torch.int64, device=dev)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    ts = []
    for it in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ctx.ransac_local_dev(d_l3, d_r3, d_l4, d_r4, m, 1, 0, H, 8, 0, 0.002, d_packed)
        e1.record(stream)
        ctx.synchronize()
        ts.append(e0.elapsed_time(e1))
    print("div %s tiles %s H %d: %.3f ms  packed %x  %s" % (os.environ.get("ERP_B200_PRUNE_DIV"), os.environ.get("ERP_B200_PRUNE_TILES"), H,
                                                           sorted(ts[1:])[2], int(d_packed.item()), ctx.last_score_stats()), flush=True)
else:
    for H in ("1000000", "125000"):
        for div, tiles in (("8", "2"), ("16", "2"), ("32", "1"), ("64", "1"), ("1000000", "1"), ("1000000", "0")):
            subprocess.run([sys.executable, __file__, H], env=dict(os.environ, ERP_B200_PRUNE_DIV=div, ERP_B200_PRUNE_TILES=tiles))
