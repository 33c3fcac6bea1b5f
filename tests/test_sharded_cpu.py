"""Host-side logic of the multi-GPU matcher under a real 2-process rendezvous (gloo, CPU): shard planning, global track
indices, packed-key ordering and the all-gather plumbing. The per-shard top-k lists come from the oracle here (there is
no GPU); the device merge kernel itself is covered by tests/test_matcher_gpu.py::test_sharded_merge_equals_unsharded."""
import os
import socket

import numpy as np
import pytest

import oracle
from hpfw_b200 import synth
from hpfw_b200.sharded import KEY_NONE, allgather_keys, allreduce_sum, pack_key, plan_shards


def test_plan_shards_contiguous_and_balanced():
    rng = np.random.default_rng(0)
    lens = rng.integers(0, 20000, size=1000)
    for n in (1, 2, 3, 4, 8):
        sh = plan_shards(lens, n)
        assert sh[0][0] == 0 and sh[-1][1] == 1000
        assert all(sh[i][1] == sh[i + 1][0] for i in range(n - 1))
        k = np.minimum(lens, 385)
        work = (lens - k + 1) * np.maximum(k, 1)
        per = np.array([work[a:b].sum() for a, b in sh], dtype=np.float64)
        assert per.max() <= 1.1 * per.mean() + work.max()
    assert plan_shards([5, 5], 4) [-1][1] == 2      # more shards than tracks: trailing shards are empty, none lost
    assert sum(b - a for a, b in plan_shards([5, 5], 4)) == 2


def test_plan_shards_weighted_by_device_speed():
    """hpfw_shard_plan_weighted: shard s gets speeds[s] / sum(speeds) of the matcher work (GPUs of one node run at different
    clocks under their power caps; the all-gather waits for the slowest rank)."""
    lens = np.full(10000, 14411)
    sh = plan_shards(lens, 4, speeds=[1.0, 1.0, 1.0, 1.0])
    assert [b - a for a, b in sh] == [2500] * 4
    sh = plan_shards(lens, 4, speeds=[1.02, 1.0, 0.98, 1.0])
    sizes = [b - a for a, b in sh]
    assert sum(sizes) == 10000 and sh[0][0] == 0 and sh[-1][1] == 10000
    assert sizes[0] == 2550 and sizes[2] == 2450 and sizes[1] == sizes[3] == 2500
    rng = np.random.default_rng(2)
    rl = rng.integers(0, 20000, size=3000)
    sp = [1.0, 2.0, 1.0]
    sh = plan_shards(rl, 3, speeds=sp)
    k = np.minimum(rl, 385)
    work = (rl - k + 1) * np.maximum(k, 1)
    per = np.array([work[a:b].sum() for a, b in sh], dtype=np.float64)
    assert abs(per[1] / per.sum() - 0.5) < 0.01 and abs(per[0] / per.sum() - 0.25) < 0.01


def test_pack_key_orders_like_the_reference_scan():
    # smaller distance first, then earlier track, then lower offset (storage.h:50-60)
    keys = [pack_key(5, 3, 7), pack_key(5, 3, 6), pack_key(5, 2, 900), pack_key(4, 1000, 1 << 19), pack_key(6, 0, 0)]
    assert sorted(keys) == [keys[3], keys[2], keys[1], keys[0], keys[4]]
    # largest legal key (HPFW_MAX_TRACKS, HPFW_MAX_TRACK_WORDS bound track and offset below 2^20 - 1) stays below KEY_NONE
    assert pack_key((1 << 24) - 1, (1 << 20) - 2, (1 << 20) - 2) < int(KEY_NONE)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, seed, topk, ret):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(seed)
        lens = rng.integers(1, 600, size=23)
        words, offs = synth.synth_hashprint_db(seed, len(lens), lens)
        qw, qo, _ = synth.synth_hashprint_queries(seed + 1, words, offs, 9, 40)
        a, b = plan_shards(lens, world, query_words=40)[rank]
        local = np.full((9, topk), KEY_NONE, dtype=np.uint64)
        for q in range(9):
            tr, d, o = oracle.find_topk(words[offs[a]:offs[b]], offs[a:b + 1] - offs[a], qw[qo[q]:qo[q + 1]], topk)
            for r in range(topk):
                if tr[r] >= 0:
                    local[q, r] = pack_key(int(d[r]), int(tr[r]) + a, int(o[r]))      # global track index
        allk = allgather_keys(torch.from_numpy(local.view(np.int64)))
        assert tuple(allk.shape) == (world, 9, topk)
        merged = np.sort(allk.numpy().view(np.uint64).transpose(1, 0, 2).reshape(9, -1), axis=1)[:, :topk]
        # unsharded oracle
        exp = np.full((9, topk), KEY_NONE, dtype=np.uint64)
        for q in range(9):
            tr, d, o = oracle.find_topk(words, offs, qw[qo[q]:qo[q + 1]], topk)
            for r in range(topk):
                if tr[r] >= 0:
                    exp[q, r] = pack_key(int(d[r]), int(tr[r]), int(o[r]))
        ret[rank] = bool(np.array_equal(merged, exp))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_allgather_merge_equals_unsharded_gloo(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    mgr = ctx.Manager()
    ret = mgr.dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 5, 6, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(ret.get(r) for r in range(world)), dict(ret)


def _np_cov(spec):
    """calc_cov(calc_frames(S)^T) in float64 (hashprint_handle.h:79-102), as in tests/test_learn_gpu.py."""
    nf = spec.shape[0] - 19
    X = np.empty((nf, 2420), dtype=np.float64)
    for b in range(121):
        for c in range(20):
            X[:, b * 20 + c] = spec[c:c + nf, b]
    X -= X.mean(axis=0)
    return X.T @ X / (nf - 1)


def _cov_worker(rank, world, port, ret):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # every rank accumulates the covariance of ITS tracks (the reference's calc_cov restated in numpy); the all-reduced
        # sum must be the accumulator one process would have built over all tracks (index-time filter learning, §8(e))
        rng = np.random.default_rng(3)
        specs = [rng.standard_normal((40 + 7 * i, 121)).astype(np.float32) for i in range(5)]
        mine = [s for i, s in enumerate(specs) if i % world == rank]
        acc = np.zeros((2420, 2420), dtype=np.float64)
        for s in mine:
            acc += _np_cov(s)
        t = torch.from_numpy(acc.astype(np.float32))
        allreduce_sum(t)
        full = np.zeros((2420, 2420), dtype=np.float64)
        for s in specs:
            full += _np_cov(s)
        err = np.abs(t.numpy().astype(np.float64) - full).max() / np.abs(full).max()
        ret[rank] = bool(err < 1e-5)
    finally:
        dist.destroy_process_group()


def test_covariance_allreduce_equals_single_process_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_cov_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    assert all(ret.get(r) for r in range(2)), dict(ret)
