"""CPU tests of the multi-GPU host logic: world_size-2 gloo processes run the same
sharding + all-reduce code the NCCL ranks run, with the oracle standing in for the device."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle as O
from erp_match_eightpoint_test_b200 import sharding, synth


def test_shard_range_partitions():
    for n in (0, 1, 7, 100, 100003):
        for w in (1, 2, 3, 8):
            parts = [sharding.shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1


def test_pack_best_order():
    assert sharding.pack_best(10, 5) > sharding.pack_best(9, 0)
    assert sharding.pack_best(10, 5) > sharding.pack_best(10, 6)      # ties -> lowest id
    assert sharding.unpack_best(sharding.pack_best(123, 456789)) == (123, 456789)
    assert sharding.pack_best(2**31 - 1, 0) < 2**63


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # ---- hypothesis sharding + packed MAX all-reduce
        kp = synth.keypoint_pair(600, 4096, 2048, seed=17)
        l, r = O.bearings(kp["left_xy"], 4096, 2048), O.bearings(kp["right_xy"], 4096, 2048)
        H = 401
        lo, hi = sharding.shard_range(H, rank, world)
        local = O.ransac(l, r, seed=3, hyp0=lo, H=hi - lo)
        t = torch.tensor([local["packed"]], dtype=torch.int64)
        sharding.allreduce_best(t, dist)
        # ---- query sharding + cross-check MIN all-reduce
        q, tr, _ = synth.descriptor_pair(301, 257, 64, seed=23)
        qlo, qhi = sharding.shard_range(301, rank, world)
        bq, bd2 = O.nn1_reverse(q[qlo:qhi], tr)
        gd2, gq = sharding.allreduce_cross_check(torch.from_numpy(bd2), torch.from_numpy(bq + qlo), dist)
        idx, dist2, _ = O.knn2(q[qlo:qhi], tr)
        keep = gq.numpy()[idx[:, 0]] == np.arange(qlo, qhi)
        mine = np.stack([np.arange(qlo, qhi)[keep], idx[keep, 0]], 1)
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        if rank == 0:
            out.put((int(t.item()), np.concatenate(gathered), gq.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world_size_2_gloo_matches_single_process():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    packed, cross, gq = out.get(timeout=240)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0

    kp = synth.keypoint_pair(600, 4096, 2048, seed=17)
    l, r = O.bearings(kp["left_xy"], 4096, 2048), O.bearings(kp["right_xy"], 4096, 2048)
    full = O.ransac(l, r, seed=3, hyp0=0, H=401)
    assert packed == full["packed"]                       # same winner regardless of world size

    q, tr, _ = synth.descriptor_pair(301, 257, 64, seed=23)
    want = O.match(q, tr, ratio=-1.0, cross_check=True)
    assert np.array_equal(cross, np.stack([want["queryIdx"], want["trainIdx"]], 1))
    assert np.array_equal(gq, O.nn1_reverse(q, tr)[0])
