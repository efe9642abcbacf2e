"""cfg5-style measurement (BASELINE.json configs[4]): a sequence of ERP frame pairs, 20k keypoints each, end to end through
ONE C-ABI call per pair (erp_pair_pose: descriptors + keypoints in, match records + pose out).

    python scripts/video_pairs.py [PAIRS=64] [THREADS=2] [NQ=20000] [HYPS=10000]

THREADS host threads each own an erp context on cuda:0 (a context is single threaded; ctypes releases the GIL during the
call), so the uploads of one pair overlap the kernels of another.  Every distinct pair is checked against the first
result of the same pair id; prints pairs/s for 1 thread and for THREADS threads."""
import sys
import threading
import time

import numpy as np

sys.path.insert(0, ".")
import bench
import erp_match_eightpoint_test_b200 as erp

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 64
threads = int(sys.argv[2]) if len(sys.argv) > 2 else 2
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 20000
hyps = int(sys.argv[4]) if len(sys.argv) > 4 else 10000
cfg = dict(bench.WORKLOADS["cfg2"], nq=nq, nt=nq, hyps=hyps)
distinct = [bench.make_pair(cfg, 0xE8B0 + 5 + i) for i in range(4)]          # four different frame pairs, cycled


def run(n_threads):
    ctxs = [erp.Context(0) for _ in range(n_threads)]
    out = [None] * pairs

    def work(tid):
        c = ctxs[tid]
        for i in range(tid, pairs, n_threads):
            p = distinct[i % len(distinct)]
            m, r = c.pair_pose(p["q"], p["t"], p["left"], p["right"], cfg["W"], cfg["H"], ratio=0.3, seed=1, H=hyps)
            out[i] = (len(m), r["packed"], int(r["count"]))

    for warm in (True, False):
        ts = [threading.Thread(target=work, args=(t,)) for t in range(n_threads)]
        t0 = time.perf_counter()
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        dt = time.perf_counter() - t0
    for c in ctxs:
        c.close()
    for i in range(pairs):
        assert out[i] == out[i % len(distinct)], (i, out[i], out[i % len(distinct)])
    return dt, out[:len(distinct)]


t1, ref = run(1)
print(f"1 thread : {pairs / t1:8.1f} pairs/s  ({1e3 * t1 / pairs:.3f} ms per pair)  results {ref}")
if threads > 1:
    tn, got = run(threads)
    assert got == ref
    print(f"{threads} threads: {pairs / tn:8.1f} pairs/s  ({1e3 * tn / pairs:.3f} ms per pair)")
