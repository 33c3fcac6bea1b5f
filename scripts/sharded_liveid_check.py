"""The rank-per-GPU sharded path (hpfw_shard_* through hpfw_b200/sharded.py) under torchrun on N GPUs against the single-GPU
path computed on rank 0: (1) byte-identical top-10 keys on a ragged synthetic hashprint DB; (2) ShardedLiveSongIdentification:
same learned filters (up to fp32 summation order of the all-reduce), same top-k records. tests/test_shard_gpu.py runs it on 2
GPUs when the box has them.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/sharded_liveid_check.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import hpfw_b200
from hpfw_b200 import HashprintExtractor, MemoryStorage, synth
from hpfw_b200.sharded import ShardedLiveSongIdentification


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = hpfw_b200.Context(local)
    out = {"world": world}
    # ---- (1) hashprint level, bit-exact: ragged synthetic DB sharded over the ranks (hpfw_shard_*: local match, in-place
    # ncclAllGather, merge kernel inside the library) against ONE GPU matching the whole DB: identical [Q][10] keys,
    # i.e. distances, tracks, offsets and the full top-10 order
    from hpfw_b200.sharded import ShardedMemoryStorage, plan_shards
    rng = np.random.default_rng(5)
    lens = rng.integers(1, 4000, size=67)
    lens[5] = 0
    words, offs = synth.synth_hashprint_db(41, len(lens), lens)
    kk = np.array([1, 63, 143, 385, 385, 385, 777, 1514])[rng.integers(0, 8, size=300)]
    qw, qo, _ = synth.synth_hashprint_queries(42, words, offs, len(kk), kk)
    sst = ShardedMemoryStorage(ctx, rank, world)
    a, b = plan_shards(lens, world)[rank]
    sst.build_local(words[offs[a]:offs[b]], offs[a:b + 1] - offs[a], track_base=a)
    dq = torch.from_numpy(qw.view(np.int64)).to(f"cuda:{local}")
    keys = sst.search_device(dq, qo, 10).clone()
    torch.cuda.synchronize()
    if rank == 0:
        one = MemoryStorage(ctx).build_packed(words, offs)
        k1 = torch.empty_like(keys)
        one.match_device(dq.data_ptr(), qo, 10, k1.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        out["keys_bit_identical"] = bool(torch.equal(k1, keys))
        out["queries"] = int(len(kk))
    if world > 1:
        # every rank holds the same merged keys
        allk = [torch.empty_like(keys) for _ in range(world)]
        dist.all_gather(allk, keys)
        out["same_on_every_rank"] = bool(all(torch.equal(allk[0], x) for x in allk))
    # ---- (2) audio level: index() with filter learning across ranks + search()
    sr = 44100
    tracks = [synth.synth_track(300 + i, 10.0 + 2.0 * (i % 3), sr) for i in range(9)]
    queries, truth = [], []
    for i in (0, 3, 4, 8, 6):
        q, _ = synth.synth_query(tracks[i], 900 + i, 6.0, sr, max_semitones=0.2)
        queries.append(q)
        truth.append(i)
    lid = ShardedLiveSongIdentification(ctx, rank, world).index(tracks)
    res = lid.search(queries, topk=3)
    out.update({"top1": [int(x) for x in res["track"][:, 0]], "truth": truth})
    if rank == 0:
        # single-GPU path over all tracks: own filter learning, own DB
        ex = HashprintExtractor(ctx)
        ex.cov_reset()
        specs = [ex.spectrogram(t) for t in tracks]
        for sp in specs:
            ex.cov_add_spectrogram(sp)
        f1, _ = ex.calc_filters()
        st = MemoryStorage(ctx).build([(str(i), ex.hashprint_from_spectrogram(sp)) for i, sp in enumerate(specs)])
        ref = st.find_topk_packed(*hpfw_b200.api.pack([ex.calc_hashprint(q) for q in queries]), 3)
        # filters span the same subspace: compare projectors (signs / near-degenerate pairs may differ)
        a, b = f1.astype(np.float64), lid.filters.astype(np.float64)
        out["filters_max_abs_diff"] = float(np.abs(a - b).max())
        out["subspace_err"] = float(np.linalg.norm(a @ (a.T @ b) - b) / np.linalg.norm(b))
        out["records_equal"] = bool(np.array_equal(ref["track"], res["track"]) and np.array_equal(ref["offset"], res["offset"]))
        out["cnt_max_diff"] = int(np.abs(ref["cnt"].astype(np.int64) - res["cnt"].astype(np.int64)).max())
        out["top1_ok"] = out["top1"] == truth
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
