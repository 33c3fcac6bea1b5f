"""BASELINE.json configs[2] and configs[4] on one GPU (configs[1] and configs[3] are bench.py's own legs):

  cfg2  live-id search: 1k-track DB, 1000 x 6 s queries, full-offset Hamming cross-correlation on 1 B200
  cfg4  scale sweep: 100k-track hashprint DB (3-min tracks, 11.5 GB), query length 2-20 s (63 .. 1514 words)

Matcher-only (hashprints generated in HBM, as SURVEY.md §8(d) allows). Checks that travel with the size: every planted
query comes back as top-1 with the planted (track, offset); keys are strictly ascending per query; a random sample of
(query, track) pairs is recomputed with torch integer ops on the device and must equal the reported per-track best
distance when that track is in the top-k. Prints one JSON line per configuration. Run on the GPU box.
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import hpfw_b200
from hpfw_b200 import MemoryStorage, synth
from hpfw_b200.api import decode_keys


def popcount64(x):
    """population count of int64 tensor (as unsigned bits)"""
    m1, m2, m4 = 0x5555555555555555, 0x3333333333333333, 0x0F0F0F0F0F0F0F0F
    x = x - ((x >> 1) & m1)
    x = (x & m2) + ((x >> 2) & m2)
    x = (x + (x >> 4)) & m4
    return ((x * 0x0101010101010101) >> 56) & 0xFF


def check_sample(words, offs, qwords, qoffs, rec, n_check, rng):
    """Recompute the best distance / offset of reported (query, track) pairs with torch integer ops."""
    bad = 0
    nq = len(qoffs) - 1
    for _ in range(n_check):
        q = int(rng.integers(0, nq))
        r = int(rng.integers(0, rec.shape[1]))
        tr = int(rec["track"][q, r])
        qw = qwords[int(qoffs[q]):int(qoffs[q + 1])]
        ref = words[int(offs[tr]):int(offs[tr + 1])]
        k = min(len(qw), len(ref))
        win = ref.unfold(0, k, 1)                                   # [n-k+1, k]
        # logical shift for the unsigned popcount: mask after arithmetic shifts
        x = win ^ qw[:k][None, :]
        lo = popcount64(x & 0x7FFFFFFFFFFFFFFF) + ((x >> 63) & 1)
        d = lo.sum(dim=1)
        best = int(d.min().item())
        off = int(torch.nonzero(d == best)[0].item())
        if best != int(rec["cnt"][q, r]) or off != int(rec["offset"][q, r]):
            bad += 1
    return bad


def run(name, n_tracks, track_words, qlens, topk, ctx, dev, n_check):
    words, offs = synth.device_hashprint_db(torch, dev, 11, n_tracks, track_words)
    st = MemoryStorage(ctx).build_device(words.data_ptr(), offs)
    rng = np.random.default_rng(5)
    parts, truths, qo = [], [], [0]
    for i, k in enumerate(qlens):                       # one query per entry, possibly different lengths
        qw, _, tr = synth.device_hashprint_queries(torch, words, offs, 100 + i, 1, int(k))
        parts.append(qw)
        truths.append(tr[0])
        qo.append(qo[-1] + int(k))
    qwords = torch.cat(parts).contiguous()
    qoffs = np.asarray(qo, dtype=np.int64)
    truth = np.stack(truths)
    nq = len(qlens)
    keys = torch.empty(nq * topk, dtype=torch.int64, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    st.match_device(qwords.data_ptr(), qoffs, topk, keys.data_ptr(), s)     # warm-up
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    st.match_device(qwords.data_ptr(), qoffs, topk, keys.data_ptr(), s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    k_np = keys.cpu().numpy().view(np.uint64).reshape(nq, topk)
    rec = decode_keys(k_np)
    top1 = float(np.mean((rec["track"][:, 0] == truth[:, 0]) & (rec["offset"][:, 0] == truth[:, 1])))
    ascending = bool(np.all(k_np[:, 1:] > k_np[:, :-1]))
    bad = check_sample(words, offs, qwords, qoffs, rec, n_check, rng)
    wops = st.word_ops(qoffs)
    line = {"config": name, "tracks": n_tracks, "track_words": track_words, "queries": nq,
            "query_words": [int(min(qlens)), int(max(qlens))], "topk": topk, "ms": ms,
            "queries_per_s": nq / (ms * 1e-3), "gwordops_per_s": wops / (ms * 1e-3) / 1e9,
            "db_gb": n_tracks * track_words * 8 / 1e9, "top1_planted": top1, "keys_ascending": ascending,
            "recomputed_pairs": n_check, "recomputed_mismatch": bad}
    print(json.dumps(line), flush=True)
    del st, words, qwords, keys
    torch.cuda.empty_cache()
    return line


def main():
    dev = torch.device("cuda:0")
    ctx = hpfw_b200.Context(0)
    t0 = time.time()
    # cfg2: 1k tracks x 1000 queries of 6 s
    run("cfg2: 1k-track DB, 1000 x 6 s queries", 1000, 14411, [385] * 1000, 10, ctx, dev, 40)
    # cfg4: 100k tracks, query length 2 .. 20 s
    qlens = [63, 143, 385, 707, 1111, 1514] * int(os.environ.get("SWEEP_PER_LENGTH", "128"))
    run("cfg4: 100k-track DB, 2-20 s queries", 100000, 14411, qlens, 10, ctx, dev, 24)
    print(f"# total {time.time() - t0:.1f} s", file=sys.stderr)
    ctx.close()


if __name__ == "__main__":
    main()
