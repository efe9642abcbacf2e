"""One tcgen05-engine call on synthetic descriptors (the ncu target).  usage: tc_profile.py NQ NT DIM [REPS] [ENGINE 2=3xTF32 | 3=1xTF32]"""
import sys
import numpy as np
sys.path.insert(0, ".")
import erp_match_eightpoint_test_b200 as erp
from erp_match_eightpoint_test_b200 import binding, synth

nq, nt, dim = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
engine = int(sys.argv[5]) if len(sys.argv) > 5 else binding.ENGINE_TCGEN05_1X
ctx = erp.Context(0)
ctx.set_engine(engine)
q, t, _ = synth.descriptor_pair(nq, nt, dim, seed=11)
for _ in range(reps):
    idx, dist = ctx.knn2_raw(q, t)
    st = ctx.last_knn_stats()
    print(f"kernel {ctx.last_knn_kernel_ms():.3f} ms  {nq * nt / ctx.last_knn_kernel_ms() / 1e9:.1f} G evals/s  rescanned {st['rescanned']} dev {st['deviation']:.2e}")
ctx.close()
