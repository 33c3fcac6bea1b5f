"""ShardedLiveSongIdentification (hpfw_b200/sharded.py) under torchrun on N GPUs against the single-GPU path computed on
rank 0: same learned filters (up to fp32 summation order of the all-reduce), same top-k records.

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
    sr = 44100
    tracks = [synth.synth_track(300 + i, 10.0 + 2.0 * (i % 3), sr) for i in range(9)]
    queries, truth = [], []
    for i in (0, 3, 4, 8, 6):
        q, _ = synth.synth_query(tracks[i], 900 + i, 6.0, sr, max_semitones=0.2)
        queries.append(q)
        truth.append(i)
    lid = ShardedLiveSongIdentification(ctx, rank, world).index(tracks)
    res = lid.search(queries, topk=3)
    out = {"world": world, "top1": [int(x) for x in res["track"][:, 0]], "truth": truth}
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
