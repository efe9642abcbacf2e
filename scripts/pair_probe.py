"""The whole pair through ONE host-buffer call (erp_pair_pose) at cfg3 size, pinned and pageable sources."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import erp_match_eightpoint_test_b200 as erp
from erp_match_eightpoint_test_b200 import synth
import bench
cfg = bench.WORKLOADS["cfg3"]
pair = bench.make_pair(cfg, 0xE8B0 + 3)
ctx = erp.Context(0)
def pin(a): return torch.from_numpy(a).pin_memory().numpy()
for tag, f in (("pinned", pin), ("pageable", lambda a: a)):
    q, t, l, r = f(pair["q"]), f(pair["t"]), f(pair["left"]), f(pair["right"])
    ts = []
    for it in range(10):
        torch.cuda.synchronize()
        a = time.perf_counter()
        m, res = ctx.pair_pose(q, t, l, r, cfg["W"], cfg["H"], ratio=0.3, seed=1, H=cfg["hyps"])
        ts.append(time.perf_counter() - a)
    print("%s: pair in one call %.3f ms (median of 8), %d matches, %d inliers, stages %s" % (tag, 1e3 * sorted(ts[2:])[4], len(m), res["count"], ctx.last_stage_ms()), flush=True)
