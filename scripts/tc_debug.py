"""First-light check of the tcgen05 engine against the exact engine (run on a B200)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import erp_match_eightpoint_test_b200 as erp
from erp_match_eightpoint_test_b200 import binding, synth

ctx = erp.Context(0)
for (nq, nt, dim) in [(128, 256, 64), (300, 700, 64), (1000, 3000, 64), (1000, 3000, 128), (1000, 3000, 32), (20000, 20000, 64)]:
    q, t, _ = synth.descriptor_pair(nq, nt, dim, seed=nq + nt)
    ctx.set_engine(binding.ENGINE_EXACT_SIMT)
    eidx, edist = ctx.knn2_raw(q, t)
    for eng in (binding.ENGINE_TCGEN05, binding.ENGINE_TCGEN05_1X):
        ctx.set_engine(eng)
        t0 = time.perf_counter()
        idx, dist = ctx.knn2_raw(q, t)
        dt = time.perf_counter() - t0
        st = ctx.last_knn_stats()
        ms = ctx.last_knn_kernel_ms()
        bad = np.nonzero((idx != eidx).any(1))[0]
        print(f"engine {eng} nq={nq} nt={nt} dim={dim}: mismatched rows {len(bad)}, rescanned {st['rescanned']}, deviation {st['deviation']:.3e}, "
              f"segs {st['chunks']}, units/cta {st['items']}, kernel {ms:.3f} ms, call {dt*1e3:.1f} ms", flush=True)
        if len(bad):
            print("  first bad rows", bad[:8], idx[bad[:4]], eidx[bad[:4]])
ctx.close()
print("done")
