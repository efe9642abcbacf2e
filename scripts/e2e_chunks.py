"""Host-buffer match call (erp_knn2_match; pinned and pageable source) at cfg3 size for ERP_B200_HOST_CHUNKS = 1, 2, 3, 4, 6:
the query set is uploaded in chunks on a copy stream while the previous chunk is searched.  One process per setting (the
knob is read once)."""
import os
import subprocess
import sys

if len(sys.argv) > 1:
    import time
    import numpy as np
    import torch
    sys.path.insert(0, ".")
    import erp_match_eightpoint_test_b200 as erp
    from erp_match_eightpoint_test_b200 import synth
    q, t, _ = synth.descriptor_pair(100000, 100000, 64, seed=0xE8B0 + 3)
    pq, pt = torch.from_numpy(q).pin_memory().numpy(), torch.from_numpy(t).pin_memory().numpy()
    ctx = erp.Context(0)
    for tag, (a_q, a_t) in {"pinned": (pq, pt), "pageable": (q, t)}.items():
        ts = []
        for it in range(12):
            torch.cuda.synchronize()
            a = time.perf_counter()
            m = ctx.knn2_match(a_q, a_t, 0.3, False)
            ts.append(time.perf_counter() - a)
        print("chunks %s, %s source: %.3f ms (median of 10), %d matches" % (sys.argv[1], tag, 1e3 * sorted(ts[2:])[5], len(m)), flush=True)
else:
    for c in ("1", "2", "3", "4", "6"):
        subprocess.run([sys.executable, __file__, c], env=dict(os.environ, ERP_B200_HOST_CHUNKS=c))
