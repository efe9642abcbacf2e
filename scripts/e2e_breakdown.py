"""Host-side timing of the e2e (host-buffer C-ABI) calls of one cfg3 step.  usage: e2e_breakdown.py [REPS]"""
import sys, time, faulthandler
faulthandler.dump_traceback_later(120, exit=True)
import numpy as np
sys.path.insert(0, ".")
import erp_match_eightpoint_test_b200 as erp
from erp_match_eightpoint_test_b200 import synth

import bench
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
cfg = bench.WORKLOADS["cfg3"]
W, H = cfg["W"], cfg["H"]
pair = bench.make_pair(cfg, 0xE8B0 + 3)
import torch
q = torch.from_numpy(pair["q"]).pin_memory().numpy()
t = torch.from_numpy(pair["t"]).pin_memory().numpy()
left, right = pair["left"], pair["right"]
ctx = erp.Context(0)
def tm(f, *a, **k):
    t0 = time.perf_counter(); r = f(*a, **k); return r, 1e3 * (time.perf_counter() - t0)
for it in range(reps):
    mt, t_match = tm(ctx.knn2_match, q, t, 0.3, False)
    (lxy, rxy), t_gather = tm(lambda: (left[mt["queryIdx"]], right[mt["trainIdx"]]))
    l3, t_b1 = tm(ctx.bearings, lxy, W, H)
    r3, t_b2 = tm(ctx.bearings, rxy, W, H)
    res, t_r = tm(ctx.ransac, l3, r3, 1, 0, 1000000, 8, 0, 0.002)
    print(f"matches {len(mt)}  knn2_match {t_match:.2f} ms | gather {t_gather:.2f} | bearings {t_b1:.2f} + {t_b2:.2f} | ransac {t_r:.2f} (keys {sorted(res.keys())})")
ctx.close()
