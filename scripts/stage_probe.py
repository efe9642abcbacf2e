import os, sys, time, subprocess
if len(sys.argv) > 1:
    import numpy as np, torch
    sys.path.insert(0, ".")
    import erp_match_eightpoint_test_b200 as erp
    from erp_match_eightpoint_test_b200 import synth
    q, t, _ = synth.descriptor_pair(100000, 100000, 64, seed=0xE8B0 + 3)
    ctx = erp.Context(0)
    ts = []
    for it in range(12):
        if it == 11: os.environ["ERP_B200_STAGE_TRACE"] = "1"
        torch.cuda.synchronize()
        a = time.perf_counter()
        m = ctx.knn2_match(q, t, 0.3, False)
        ts.append(time.perf_counter() - a)
    print("threads %s chunks %s pageable: %.3f ms (median), last %.3f" % (sys.argv[1], sys.argv[2], 1e3 * sorted(ts[2:])[5], 1e3*ts[-1]), flush=True)
    pq, pt = torch.from_numpy(q).pin_memory().numpy(), torch.from_numpy(t).pin_memory().numpy()
    os.environ.pop("ERP_B200_STAGE_TRACE")
    ts = []
    for it in range(8):
        if it == 7: os.environ["ERP_B200_STAGE_TRACE"] = "1"
        torch.cuda.synchronize()
        a = time.perf_counter()
        m = ctx.knn2_match(pq, pt, 0.3, False)
        ts.append(time.perf_counter() - a)
    print("threads %s chunks %s pinned: %.3f ms (median), last %.3f" % (sys.argv[1], sys.argv[2], 1e3 * sorted(ts[2:])[3], 1e3*ts[-1]), flush=True)
    # raw host memcpy rate of one thread
    buf = np.empty_like(q)
    a = time.perf_counter(); np.copyto(buf, q); b = time.perf_counter()
    print("  numpy copy of 25.6 MB: %.3f ms" % (1e3*(b-a)))
else:
    for th, ch in (("4", "0"), ("4", "3"), ("8", "0"), ("2", "0")):          # chunks 0: the library's own cut
        subprocess.run([sys.executable, __file__, th, ch], env=dict(os.environ, ERP_B200_STAGE_THREADS=th, ERP_B200_HOST_CHUNKS=ch))
