"""Times scoring of H hypotheses x m correspondences with both engines.  usage: score_profile.py H M [REPS]"""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
import erp_match_eightpoint_test_b200 as erp
import oracle as O
from erp_match_eightpoint_test_b200 import binding, synth

H, m = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
ctx = erp.Context(0)
kp = synth.keypoint_pair(m, 8192, 4096, seed=5)
l, r = O.bearings(kp["left_xy"], 8192, 4096), O.bearings(kp["right_xy"], 8192, 4096)
dev = torch.device("cuda", 0)
d_l3, d_r3 = torch.from_numpy(l).to(dev), torch.from_numpy(r).to(dev)
d_l4 = torch.empty((m, 4), dtype=torch.float32, device=dev); d_r4 = torch.empty_like(d_l4)
ctx.pack_float4_dev(d_l3, m, d_l4); ctx.pack_float4_dev(d_r3, m, d_r4)
d_packed = torch.zeros(1, dtype=torch.int64, device=dev)
out = {}
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
for name, eng in (("simt", binding.ENGINE_EXACT_SIMT), ("tc", binding.ENGINE_TCGEN05)):
    ctx.set_engine(eng)
    for it in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ctx.ransac_local_dev(d_l3, d_r3, d_l4, d_r4, m, 1, 0, H, 8, 0, 0.002, d_packed)
        e1.record(stream)
        ctx.synchronize()
        print(f"{name}: {e0.elapsed_time(e1):.3f} ms  {H / e0.elapsed_time(e1) / 1e3:.2f} M hyps/s", flush=True)
    out[name] = int(d_packed.item())
print("packed equal:", out["simt"] == out["tc"], hex(out["tc"]), "count", out["tc"] >> 32)
ctx.close()
