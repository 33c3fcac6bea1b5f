"""The database sharded over GPUs (hpfw_shard_*, hpfw_b200/csrc/shard.cu): the merged result must be bit-identical to one
GPU holding the whole database (reference semantics: db::MemoryStorage::find, storage.h:27-64; tie rules storage.h:50-60).

Runs on whatever the box has: the single-process ("local") mode with every visible GPU (one GPU exercises the same code
without collectives); the rank-per-GPU mode under torchrun needs >= 2 GPUs and is skipped otherwise."""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle
from hpfw_b200 import MemoryStorage, synth
from hpfw_b200._lib import Match, check

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("n_dev", [1, 2, 4, 8])
def test_local_shards_equal_one_gpu(ctx, n_dev):
    if n_dev > _gpus():
        pytest.skip(f"needs {n_dev} GPUs")
    L = ctx._lib
    rng = np.random.default_rng(3)
    lens = rng.integers(1, 3000, size=53)
    lens[7] = 0
    words, offs = synth.synth_hashprint_db(21, len(lens), lens)
    kk = np.array([1, 40, 143, 385, 385, 900])[rng.integers(0, 6, size=160)]
    qw, qo, _ = synth.synth_hashprint_queries(22, words, offs, len(kk), kk)
    topk = 7
    whole = MemoryStorage(ctx).build_packed(words, offs).find_topk_packed(qw, qo, topk)
    h = C.c_void_p()
    check(L.hpfw_shard_create_local(None, n_dev, C.byref(h)))
    try:
        assert L.hpfw_shard_world(h) == n_dev
        check(L.hpfw_shard_build(h, words.ctypes.data_as(C.c_void_p), offs.ctypes.data_as(C.c_void_p), len(lens), 385))
        out = np.zeros((len(kk), topk), dtype=whole.dtype)
        check(L.hpfw_shard_find_topk(h, qw.ctypes.data_as(C.c_void_p), qo.ctypes.data_as(C.c_void_p), len(kk), topk,
                                     out.ctypes.data_as(C.POINTER(Match))))
        for f in ("track", "cnt", "offset"):
            assert np.array_equal(out[f], whole[f]), f
        # a second batch through the same buffers, and a single find()
        out2 = np.zeros((1, 1), dtype=whole.dtype)
        one_q = np.array([0, qo[1]], dtype=np.int64)
        check(L.hpfw_shard_find_topk(h, qw.ctypes.data_as(C.c_void_p), one_q.ctypes.data_as(C.c_void_p), 1, 1,
                                     out2.ctypes.data_as(C.POINTER(Match))))
        assert out2["track"][0, 0] == whole["track"][0, 0] and out2["cnt"][0, 0] == whole["cnt"][0, 0]
    finally:
        L.hpfw_shard_destroy(h)
    tr, d, o = oracle.find_topk_batch(words, offs, qw, qo, topk, 8)
    assert np.array_equal(whole["track"], tr) and np.array_equal(whole["cnt"], d) and np.array_equal(whole["offset"], o)


def test_shard_plan_matches_contract():
    """hpfw_shard_plan is host-only: contiguous, complete, balanced (also covered without a GPU in test_sharded_cpu.py)."""
    from hpfw_b200.sharded import plan_shards
    lens = np.random.default_rng(1).integers(0, 20000, size=500)
    sh = plan_shards(lens, 8)
    assert sh[0][0] == 0 and sh[-1][1] == 500 and all(sh[i][1] == sh[i + 1][0] for i in range(7))


def test_rank_mode_two_gpus_bit_identical():
    """One process per GPU under torchrun (the bench's N > 1 layout): merged keys byte-identical to one GPU, the same on
    every rank; ShardedLiveSongIdentification (index with the covariance all-reduce + search) gives the single-GPU records."""
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611",
                        os.path.join(ROOT, "scripts", "sharded_liveid_check.py")], capture_output=True, text=True, timeout=600)
    recs = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert p.returncode == 0 and recs, p.stderr[-2000:]
    r = json.loads(recs[-1])
    assert r["world"] == 2 and r["keys_bit_identical"] and r["same_on_every_rank"], r
    assert r["records_equal"] and r["top1_ok"] and r["subspace_err"] < 1e-3, r
