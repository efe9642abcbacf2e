"""CPU oracle binding (TEST INFRASTRUCTURE ONLY).

ctypes wrapper over ``oracle/_build/liberp_oracle.so`` (built from ``erp_oracle.c``,
a plain-C restatement of the reference's hot path; see ``erp_oracle.h`` for the
reference file:line each function follows).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product
(``erp_match_eightpoint_test_b200``) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liberp_oracle.so")

METRIC_ALGEBRAIC, METRIC_SAMPSON, METRIC_ANGULAR = 0, 1, 2

DMATCH = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"), ("distance", "<f4")])


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "erp_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(
        os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "erp_oracle.h"))
    ):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_ransac.restype = C.c_uint64
    return _lib


def _p(a, t=C.c_void_p):
    return a.ctypes.data_as(t) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(int(n))


def num_threads() -> int:
    return lib().orc_num_threads()


def knn2(q, t):
    q, t = _f32(q), _f32(t)
    nq, dim = q.shape
    nt = t.shape[0]
    assert nt >= 2 and t.shape[1] == dim
    idx = np.empty((nq, 2), np.int32)
    dist = np.empty((nq, 2), np.float32)
    d2 = np.empty((nq, 2), np.float64)
    lib().orc_knn2(_p(q), nq, _p(t), nt, dim, _p(idx), _p(dist), _p(d2))
    return idx, dist, d2


def nn1_reverse(q, t):
    q, t = _f32(q), _f32(t)
    best = np.empty(t.shape[0], np.int32)
    d2 = np.empty(t.shape[0], np.float64)
    lib().orc_nn1_reverse(_p(q), q.shape[0], _p(t), t.shape[0], q.shape[1], _p(best), _p(d2))
    return best, d2


def match(q, t, ratio=0.3, cross_check=False):
    q, t = _f32(q), _f32(t)
    out = np.empty(max(q.shape[0], 1), DMATCH)
    n = lib().orc_match(_p(q), q.shape[0], _p(t), t.shape[0], q.shape[1],
                        C.c_float(ratio), int(cross_check), _p(out))
    return out[:n].copy()


def bearings(xy, W, H):
    """xy: (n,2) float32 pixel coordinates (KeyPoint.pt)."""
    xy = _f32(xy)
    out = np.empty((xy.shape[0], 3), np.float64)
    lib().orc_bearings(_p(xy), 8, xy.shape[0], int(W), int(H), _p(out))
    return out


def inv3(A):
    A = _f64(A).reshape(9)
    out = np.empty(9)
    lib().orc_inv3(_p(A), _p(out))
    return out.reshape(3, 3)


def rotate_pixels(rc, R, W, H):
    """erp_rotation::rotate_pixel on (n,2) int32 (row, col) pairs."""
    rc = np.ascontiguousarray(rc, np.int32)
    out = np.empty_like(rc)
    lib().orc_rotate_pixels(_p(rc), rc.shape[0], _p(_f64(R).reshape(9)), int(W), int(H), _p(out))
    return out


def rotate_image(im, R):
    """erp_rotation::rotate_image on an (H, W, 3) uint8 image."""
    im = np.ascontiguousarray(im, np.uint8)
    out = np.empty_like(im)
    lib().orc_rotate_image(_p(im), im.shape[1], im.shape[0], _p(_f64(R).reshape(9)), _p(out))
    return out


def crop_rotated_image(im, pitch_deg):
    im = np.ascontiguousarray(im, np.uint8)
    out = np.empty((im.shape[0] // 4, im.shape[1], 3), np.uint8)
    lib().orc_crop_rotated_image(_p(im), im.shape[1], im.shape[0], C.c_float(pitch_deg), _p(out))
    return out


def rotate_keypoints(xy, pitch_inv_deg, W, H):
    xy = np.array(xy, np.float32, copy=True, order="C")
    lib().orc_rotate_keypoints(_p(xy), 8, xy.shape[0], C.c_float(pitch_inv_deg), int(W), int(H))
    return xy


def draw_epipole(E, left_xy, right_xy, im_w, im_h, out_w, out_h):
    """epipolar_tool::draw_epipole for the given (already selected) correspondences."""
    left_xy, right_xy = _f32(left_xy), _f32(right_xy)
    out = np.empty((out_h, out_w, 3), np.uint8)
    lib().orc_draw_epipole(_p(_f64(E).reshape(9)), _p(left_xy), _p(right_xy), 8, left_xy.shape[0], int(im_w), int(im_h),
                           int(out_w), int(out_h), _p(out))
    return out


def eular2rot(theta):
    R = np.empty(9, np.float64)
    lib().orc_eular2rot(_p(_f64(theta)), _p(R))
    return R.reshape(3, 3)


def rot2eular(R):
    e = np.empty(3, np.float64)
    lib().orc_rot2eular(_p(_f64(R).reshape(9)), _p(e))
    return e


def svd(A):
    A = _f64(A)
    m, n = A.shape
    k = min(m, n)
    w = np.empty(k)
    u = np.empty((m, k))
    vt = np.empty((k, n))
    lib().orc_svd(_p(A), m, n, _p(w), _p(u), _p(vt))
    return w, u, vt


def decompose_essential(E):
    R1, R2, t = np.empty(9), np.empty(9), np.empty(3)
    lib().orc_decompose_essential(_p(_f64(E).reshape(9)), _p(R1), _p(R2), _p(t))
    return R1.reshape(3, 3), R2.reshape(3, 3), t


def eight_point(l3, r3, null_mode=1):
    l3, r3 = _f64(l3), _f64(r3)
    n = l3.shape[0]
    e9, Ec = np.empty(9), np.empty(9)
    R1, R2, T = np.empty(3, np.float32), np.empty(3, np.float32), np.empty(3, np.float32)
    v1, v2 = C.c_int(0), C.c_int(0)
    lib().orc_eight_point(_p(l3), _p(r3), n, int(null_mode), _p(e9), _p(Ec), _p(R1), _p(R2), _p(T),
                          C.byref(v1), C.byref(v2))
    return dict(e=e9.reshape(3, 3), E=Ec.reshape(3, 3), R1=R1, R2=R2, T=T,
                R1_valid=bool(v1.value), R2_valid=bool(v2.value))


def random_array(size, reseed=1):
    out = np.empty(size, np.int32)
    lib().orc_random_array(size, _p(out), reseed)
    return out


def ref_sample_table(M, H=80, S=None, reseed=1):
    S = int(M * 0.25) if S is None else S
    t = np.empty((H, S), np.int32)
    lib().orc_ref_sample_table(M, H, S, _p(t), reseed)
    return t


def philox_samples(seed, hyp_id, M, S=8):
    out = np.empty(S, np.int32)
    lib().orc_philox_samples(C.c_uint64(seed), C.c_uint64(hyp_id), M, S, _p(out))
    return out


def initial_guess(l3, r3, table, null_mode=1):
    l3, r3 = _f64(l3), _f64(r3)
    table = np.ascontiguousarray(table, np.int32)
    H, S = table.shape
    R, T = np.empty(3, np.float32), np.empty(3, np.float32)
    cR, cT = np.empty((2 * H, 3), np.float32), np.empty((2 * H, 3), np.float32)
    nc, ch = C.c_int(0), C.c_int(0)
    rc = lib().orc_initial_guess(_p(l3), _p(r3), l3.shape[0], _p(table), H, S, int(null_mode),
                                 _p(R), _p(T), _p(cR), _p(cT), C.byref(nc), C.byref(ch))
    return dict(rc=rc, R=R, T=T, cand_R=cR[:nc.value].copy(), cand_T=cT[:nc.value].copy(), chosen=ch.value)


def consensus_pick(cand_R):
    cand_R = _f32(cand_R)
    tm = np.empty(cand_R.shape[0], np.float64)
    i = lib().orc_consensus_pick(_p(cand_R), cand_R.shape[0], _p(tm))
    return i, tm


def to4(v3):
    v3 = np.asarray(v3)
    out = np.zeros((v3.shape[0], 4), np.float32)
    out[:, :3] = v3.astype(np.float32)
    return out


def score(E, l3, r3, metric=METRIC_ALGEBRAIC, tau=0.002):
    E = _f64(E).reshape(-1, 9)
    l4, r4 = to4(l3), to4(r3)
    counts = np.empty(E.shape[0], np.int32)
    lib().orc_score(_p(E), E.shape[0], _p(l4), _p(r4), l4.shape[0], int(metric), C.c_float(tau), _p(counts))
    return counts


def inlier_mask(E, l3, r3, metric=METRIC_ALGEBRAIC, tau=0.002):
    l4, r4 = to4(l3), to4(r3)
    mask = np.empty(l4.shape[0], np.uint8)
    lib().orc_inlier_mask(_p(_f64(E).reshape(9)), _p(l4), _p(r4), l4.shape[0], int(metric), C.c_float(tau), _p(mask))
    return mask


def ransac(l3, r3, seed, hyp0, H, S=8, metric=METRIC_ALGEBRAIC, tau=0.002, want_counts=True):
    l3, r3 = _f64(l3), _f64(r3)
    E = np.empty(9)
    counts = np.empty(H, np.int32) if want_counts else None
    packed = lib().orc_ransac(_p(l3), _p(r3), l3.shape[0], C.c_uint64(seed), C.c_uint64(hyp0), H, S,
                              int(metric), C.c_float(tau), _p(E), _p(counts))
    count = packed >> 32
    hyp = 0xFFFFFFFF - (packed & 0xFFFFFFFF)
    return dict(packed=packed, count=int(count), hyp=int(hyp), E=E.reshape(3, 3), counts=counts)
