"""cfg3-sized RANSAC (1M hypotheses x 50k correspondences) for the three residuals, SIMT engine vs tensor-core search:
python scripts/metric_profile.py [H=1000000] [M=50000]"""
import sys
import numpy as np
sys.path.insert(0, ".")
import erp_match_eightpoint_test_b200 as erp
import oracle as O
from erp_match_eightpoint_test_b200 import binding, synth
import time

H = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
kp = synth.keypoint_pair(m, 8192, 4096, seed=5)
l, r = O.bearings(kp["left_xy"], 8192, 4096), O.bearings(kp["right_xy"], 8192, 4096)
ctx = erp.Context(0)
for metric in (0, 1, 2):
    out = {}
    for name, eng in (("simt", binding.ENGINE_EXACT_SIMT), ("tc", binding.ENGINE_TCGEN05)):
        ctx.set_engine(eng)
        for it in range(2):
            t0 = time.perf_counter()
            res = ctx.ransac(l, r, seed=1, hyp_offset=0, H=H, metric=metric, tau=0.002)
            dt = time.perf_counter() - t0
        out[name] = (res["packed"], dt)
        print(f"metric {metric} {name}: {dt * 1e3:.2f} ms per call (host buffers), count {res['count']}", ctx.last_score_stats() if name == "tc" else "")
    assert out["simt"][0] == out["tc"][0]
ctx.close()
