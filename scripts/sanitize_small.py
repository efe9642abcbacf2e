"""Small end-to-end pass for compute-sanitizer memcheck (all engines, sizes that finish under the tool)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import erp_match_eightpoint_test_b200 as erp
from erp_match_eightpoint_test_b200 import binding, synth

ctx = erp.Context(0)
q, t, planted = synth.descriptor_pair(700, 1300, 64, seed=3)
ref = None
for eng in (binding.ENGINE_EXACT_SIMT, binding.ENGINE_TCGEN05, binding.ENGINE_TCGEN05_1X):
    ctx.set_engine(eng)
    m = ctx.knn2_match(q, t, 0.3, True)
    ref = m if ref is None else ref
    assert m.tobytes() == ref.tobytes()
q2, t2, _ = synth.descriptor_pair(300, 400, 128, seed=4)
ctx.set_engine(binding.ENGINE_TCGEN05_1X); a = ctx.knn2_raw(q2, t2)
ctx.set_engine(binding.ENGINE_TCGEN05); b = ctx.knn2_raw(q2, t2)
assert np.array_equal(a[0], b[0])
kp = synth.keypoint_pair(2500, 4096, 2048, seed=5)
l, r = ctx.bearings(kp["left_xy"], 4096, 2048), ctx.bearings(kp["right_xy"], 4096, 2048)
res = {}
for eng in (binding.ENGINE_EXACT_SIMT, binding.ENGINE_TCGEN05):
    ctx.set_engine(eng)
    res[eng] = ctx.ransac(l, r, seed=2, hyp_offset=0, H=3000)
assert res[binding.ENGINE_EXACT_SIMT]["packed"] == res[binding.ENGINE_TCGEN05]["packed"]
ctx.set_engine(binding.ENGINE_AUTO)
R, T = ctx.find(4096, 2048, kp["left_xy"][kp["inlier"]][:400], kp["right_xy"][kp["inlier"]][:400])
im = (np.arange(64 * 128 * 3) % 251).astype(np.uint8).reshape(64, 128, 3)
ctx.rotate_image(im, np.eye(3)[[1, 0, 2]] * [1, 1, -1]); ctx.crop_rotated_image(im, 33.0)
ctx.draw_epipole(kp["E"].T, kp["left_xy"][:7], kp["right_xy"][:7], 4096, 2048, 160, 90)
ctx.close()
print("sanitize_small ok", len(ref), res[binding.ENGINE_TCGEN05]["count"])
