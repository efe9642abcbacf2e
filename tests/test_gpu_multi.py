"""GPU tests of north_star's split of ONE ERP pair over several B200s: query rows and hypothesis ids per
GPU, match lists all-gathered in rank order, one 8-byte max all-reduce of the packed best model, min all-reduce for
cross-check -- all inside liberp_b200.so (NCCL, no torch).  The bar: the match records (bytes), the packed winner, the
inlier mask and the refit are IDENTICAL for G = 1, 2, 4, 8 devices.  Group sizes beyond the box's device count are
skipped (the 1-GPU box runs G = 1; `gpurun --gpus 8` runs all of them, log under profiles/)."""
import numpy as np
import pytest

import erp_match_eightpoint_test_b200 as erp
import oracle as O
from erp_match_eightpoint_test_b200 import binding, synth

pytestmark = pytest.mark.gpu


def n_devices():
    return erp.lib().erp_device_count()


def scene_pair(nq, nt, dim, W, H, seed):
    q, t, planted = synth.descriptor_pair(nq, nt, dim, seed=seed)
    kp = synth.keypoint_pair(int((planted >= 0).sum()), W, H, seed=seed + 1)
    rng = np.random.default_rng(seed + 2)
    left = (rng.uniform(0, 1, (nq, 2)) * [W, H - 1]).astype(np.float32)
    right = (rng.uniform(0, 1, (nt, 2)) * [W, H - 1]).astype(np.float32)
    qi = np.nonzero(planted >= 0)[0]
    left[qi] = kp["left_xy"]
    right[planted[qi]] = kp["right_xy"]
    return q, t, left, right


@pytest.fixture(scope="module")
def single():
    c = erp.Context(0)
    yield c
    c.close()


def test_shard_range_matches_the_python_helper():
    from erp_match_eightpoint_test_b200 import sharding
    for n in (0, 1, 7, 100, 100003):
        for w in (1, 2, 3, 8):
            for r in range(w):
                assert erp.shard_range(n, r, w) == sharding.shard_range(n, r, w)


@pytest.mark.parametrize("G", [1, 2, 4, 8])
def test_group_pair_pose_is_independent_of_the_group_size(single, G):
    if n_devices() < G:
        pytest.skip(f"{G} devices needed, {n_devices()} visible")
    with erp.Group(range(G)) as grp:
        assert len(grp) == G
        # tensor-core search (H * m >= 3e7) and the SIMT path; ragged sizes: nq, nt, H not multiples of G
        for (nq, nt, H, cross, seed) in [(6001, 7003, 60001, False, 41), (6001, 7003, 60001, True, 41), (1501, 1999, 2049, False, 43)]:
            q, t, left, right = scene_pair(nq, nt, 64, 4096, 2048, seed)
            wm, want = single.pair_pose(q, t, left, right, 4096, 2048, ratio=0.3, cross_check=cross, seed=7, H=H)
            gm, got = grp.pair_pose(q, t, left, right, 4096, 2048, ratio=0.3, cross_check=cross, seed=7, H=H)
            assert gm.tobytes() == wm.tobytes()
            assert got["packed"] == want["packed"] and got["count"] == want["count"]
            assert np.array_equal(got["mask"], want["mask"])
            assert np.array_equal(got["E"], want["E"]) and np.array_equal(got["E_refit"], want["E_refit"])
            assert np.array_equal(got["pose"], want["pose"])
        # the oracle on the last one: records and winner
        om = O.match(q, t, 0.3, False)
        assert gm.tobytes() == om.tobytes()
        l, r = O.bearings(left[om["queryIdx"]], 4096, 2048), O.bearings(right[om["trainIdx"]], 4096, 2048)
        assert got["packed"] == O.ransac(l, r, seed=7, hyp0=0, H=2049)["packed"]


@pytest.mark.parametrize("G", [1, 2, 4, 8])
@pytest.mark.parametrize("dim", [64, 128])
def test_group_match_with_cross_check_min_reduce(single, G, dim):
    """Query-sharded match_two_image: the per-train nearest query is min-reduced over the ranks (d2, then lowest id)."""
    if n_devices() < G:
        pytest.skip(f"{G} devices needed, {n_devices()} visible")
    q, t, _ = synth.descriptor_pair(5003, 4099, dim, seed=51 + dim)
    t[100] = t[7]; q[11] = q[4000]; q[12] = q[4000]                 # ties across shards: lowest index wins on both sides
    with erp.Group(range(G)) as grp:
        for ratio, cross in [(0.3, False), (0.3, True), (-1.0, True), (0.8, True)]:
            want = O.match(q, t, ratio, cross)
            got = grp.knn2_match(q, t, ratio, cross)
            assert got.tobytes() == want.tobytes(), (G, dim, ratio, cross)
            assert single.knn2_match(q, t, ratio, cross).tobytes() == want.tobytes()
        # fewer queries than ranks: empty shards take part in the collectives
        few = grp.knn2_match(q[:3], t, -1.0, True)
        assert few.tobytes() == O.match(q[:3], t, -1.0, True).tobytes()


@pytest.mark.parametrize("peer", ["1", "0"])
@pytest.mark.timeout(600)
def test_one_process_per_gpu_clique_equals_single_gpu(single, tmp_path, peer):
    """The torchrun layout (one process per GPU, erp_comm_init): peer windows mapped with cudaIpc (peer = 1) or NCCL
    collectives (peer = 0).  Rank 0's records, winner, mask and refit equal the single-GPU call bit for bit."""
    import os
    import socket
    import subprocess
    import sys
    G = min(n_devices(), 8)
    if G < 2:
        pytest.skip("2 devices needed")
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    out = tmp_path / "mp.npz"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, ERP_B200_PEER=peer)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={G}", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(root, "tests", "mp_worker.py"), str(out)],
                       capture_output=True, text=True, timeout=540, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    z = np.load(out)
    for tag, (nq, nt, H, cross, seed) in {"tc": (6001, 7003, 60001, False, 41), "cross": (6001, 7003, 60001, True, 41), "simt": (1501, 1999, 2049, False, 43)}.items():
        q, t, left, right = scene_pair(nq, nt, 64, 4096, 2048, seed)
        wm, want = single.pair_pose(q, t, left, right, 4096, 2048, ratio=0.3, cross_check=cross, seed=7, H=H)
        assert z[tag + "_matches"].tobytes() == wm.tobytes() == z[tag + "_parts"].tobytes()
        assert int(z[tag + "_packed"]) == want["packed"]
        assert np.array_equal(z[tag + "_mask"], want["mask"]) and np.array_equal(z[tag + "_E_refit"], want["E_refit"])


def test_dropin_classes_fan_out_over_erp_b200_devices(tmp_path):
    """The C++ classes with $ERP_B200_DEVICES naming several GPUs: feature_matcher::match_two_image splits the query
    rows inside the call (erp_group); the driver's output file equals the one-GPU run byte for byte."""
    import os
    import subprocess
    import test_host_dropin as hd
    if n_devices() < 2:
        pytest.skip("2 devices needed")
    hd._build()
    q, t, lxy, rxy, W, H = hd._scene()
    inp = tmp_path / "in.bin"
    hd._write_input(inp, q, t, lxy, rxy, W, H)
    outs = []
    for devs in ("0", ",".join(str(i) for i in range(min(n_devices(), 4)))):
        out = tmp_path / ("out_%d.bin" % len(devs))
        r = subprocess.run([hd.EXE, str(inp), str(out)], capture_output=True, text=True, timeout=300, env=dict(os.environ, ERP_B200_DEVICES=devs))
        assert r.returncode == 0, r.stderr
        outs.append(open(out, "rb").read())
    assert outs[0] == outs[1]


def test_graph_cache_follows_the_arguments(single):
    """The device-resident pair call is captured into a CUDA graph on its second use and replayed afterwards; a call with
    other arguments (seed, hypotheses, metric, sizes) must never replay a stale graph.  Interleaved calls against the oracle."""
    q, t, left, right = scene_pair(3000, 3500, 64, 4096, 2048, 71)
    om = O.match(q, t, 0.3, False)
    l, r = O.bearings(left[om["queryIdx"]], 4096, 2048), O.bearings(right[om["trainIdx"]], 4096, 2048)
    q2, t2, left2, right2 = scene_pair(2000, 3500, 64, 4096, 2048, 72)
    om2 = O.match(q2, t2, 0.3, False)
    l2, r2 = O.bearings(left2[om2["queryIdx"]], 4096, 2048), O.bearings(right2[om2["trainIdx"]], 4096, 2048)
    want = {}
    for key, (seed, H, metric) in {"a": (5, 3000, 0), "b": (6, 3000, 0), "c": (5, 4000, 1)}.items():
        want[key] = O.ransac(l, r, seed=seed, hyp0=0, H=H, metric=metric)["packed"]
    want["d"] = O.ransac(l2, r2, seed=5, hyp0=0, H=3000)["packed"]
    args = {"a": (q, t, left, right, 5, 3000, 0), "b": (q, t, left, right, 6, 3000, 0), "c": (q, t, left, right, 5, 4000, 1),
            "d": (q2, t2, left2, right2, 5, 3000, 0)}
    for key in "aaabbacccaddaab":
        qq, tt, ll, rr, seed, H, metric = args[key]
        m, res = single.pair_pose(qq, tt, ll, rr, 4096, 2048, ratio=0.3, seed=seed, H=H, metric=metric)
        assert res["packed"] == want[key], key
        assert m.tobytes() == (om2 if key == "d" else om).tobytes()


def test_pair_pose_dev_keeps_the_match_count_on_the_device(single):
    """erp_pair_pose (one call, no host synchronisation between matcher and pose) against the oracle end to end."""
    q, t, left, right = scene_pair(5000, 6000, 64, 4096, 2048, 61)
    for metric in (0, 1, 2):
        m, r = single.pair_pose(q, t, left, right, 4096, 2048, ratio=0.3, seed=9, H=4000, metric=metric)
        om = O.match(q, t, 0.3, False)
        assert m.tobytes() == om.tobytes()
        l, rr = O.bearings(left[om["queryIdx"]], 4096, 2048), O.bearings(right[om["trainIdx"]], 4096, 2048)
        ref = O.ransac(l, rr, seed=9, hyp0=0, H=4000, metric=metric)
        assert r["packed"] == ref["packed"]
        assert np.array_equal(r["mask"], O.inlier_mask(ref["E"], l, rr, metric=metric))
        assert r["n_refit"] == int(r["mask"].sum())
    ms = single.last_stage_ms()
    assert all(x >= 0 for x in ms) and sum(ms) > 0
