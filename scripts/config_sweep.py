"""BASELINE.json configs[2] and configs[4] on one GPU; with the argument `sharded` under torchrun, configs[3] (10k-track DB,
ONE batch of 10,000 queries) and configs[4] on N GPUs through the library's NCCL path (configs[1] is bench.py's extraction leg):

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


def run_sharded(name, n_tracks, track_words, qlens_per_rank, topk, ctx, dev, rank, world, n_check):
    """The same over `world` ranks (torchrun, one GPU each): DB sharded by track through the library's own NCCL path
    (hpfw_shard_match_device: local match, in-place ncclAllGather, merge kernel). Every rank plants its queries in ITS shard;
    the query batch is the concatenation over ranks (replicated). Rank 0 recomputes sampled pairs of its own shard."""
    import torch.distributed as dist
    from hpfw_b200.sharded import ShardedMemoryStorage
    lo, hi = rank * n_tracks // world, (rank + 1) * n_tracks // world
    words, offs = synth.device_hashprint_db(torch, dev, 11 + rank, hi - lo, track_words)
    st = ShardedMemoryStorage(ctx, rank, world)
    s = torch.cuda.current_stream().cuda_stream
    st.build_local_device(words.data_ptr(), offs, track_base=lo, stream=s)
    parts, truths, qlens_local = [], [], []
    for i, (k, cnt) in enumerate(qlens_per_rank):        # (query words, how many): one generator call per length
        qw, _, tr = synth.device_hashprint_queries(torch, words, offs, 1000 * rank + 100 + i, int(cnt), int(k))
        parts.append(qw)
        truths.append(tr)
        qlens_local += [int(k)] * int(cnt)
    q_local = torch.cat(parts).contiguous()
    truth_local = np.concatenate(truths)
    truth_local[:, 0] += lo
    nql = len(qlens_local)
    if world > 1:
        q_all = torch.empty(world * q_local.numel(), dtype=torch.int64, device=dev)
        st.allgatherv(q_local.data_ptr(), q_all.data_ptr(), [8 * q_local.numel()] * world, s)
        t_all = torch.empty((world, nql, 2), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(t_all.view(-1), torch.from_numpy(truth_local).to(dev).view(-1))
        truth = t_all.view(-1, 2).cpu().numpy()
    else:
        q_all, truth = q_local, truth_local
    qlens = qlens_local * world
    qoffs = np.zeros(len(qlens) + 1, dtype=np.int64)
    np.cumsum(qlens, out=qoffs[1:])
    nq = len(qlens)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    keys = st.search_device(q_all, qoffs, topk)          # warm-up (routing tables, self-test, scratch)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    keys = st.search_device(q_all, qoffs, topk)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    k_np = keys.cpu().numpy().view(np.uint64).reshape(nq, topk)
    rec = decode_keys(k_np)
    top1 = float(np.mean((rec["track"][:, 0] == truth[:, 0]) & (rec["offset"][:, 0] == truth[:, 1])))
    ascending = bool(np.all(k_np[:, 1:] > k_np[:, :-1]))
    line = None
    if rank == 0:
        # recompute sampled (query, reported track) pairs whose track lives in this rank's shard
        rng = np.random.default_rng(5)
        bad = checked = 0
        for _ in range(n_check * 8):
            if checked >= n_check:
                break
            q, r = int(rng.integers(0, nq)), int(rng.integers(0, topk))
            tr = int(rec["track"][q, r])
            if not (lo <= tr < hi):
                continue
            one = np.zeros((1, 1), dtype=rec.dtype)
            one[0, 0] = rec[q, r]
            one["track"] -= lo
            bad += check_sample(words, offs, q_all[int(qoffs[q]):int(qoffs[q + 1])], np.array([0, qoffs[q + 1] - qoffs[q]]), one, 1,
                                np.random.default_rng(0))
            checked += 1
        wops = float(sum((track_words - min(k, track_words) + 1) * min(k, track_words) for k in qlens)) * n_tracks
        line = {"config": name, "n_gpus": world, "tracks": n_tracks, "track_words": track_words, "queries": nq,
                "query_words": [int(min(qlens)), int(max(qlens))], "topk": topk, "ms": ms,
                "queries_per_s": nq / (ms * 1e-3), "gwordops_per_s": wops / (ms * 1e-3) / 1e9,
                "db_gb": n_tracks * track_words * 8 / 1e9, "db_gb_per_gpu": (hi - lo) * track_words * 8 / 1e9,
                "top1_planted": top1, "keys_ascending": ascending, "recomputed_pairs": checked, "recomputed_mismatch": bad,
                "path": "hpfw_shard_match_device (NCCL inside the library)"}
        print(json.dumps(line), flush=True)
    del st, words, q_all, keys
    torch.cuda.empty_cache()
    return line


def main_sharded():
    """torchrun --nproc-per-node N scripts/config_sweep.py sharded: BASELINE configs[3] (10k-track DB, ONE batch of 10,000
    six-second queries) and configs[4] (100k-track DB, 2-20 s queries) on N GPUs."""
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = hpfw_b200.Context(local)
    per_rank3 = int(os.environ.get("SWEEP_CFG3_QUERIES", "10000")) // world
    run_sharded(f"cfg3: 10k-track DB on {world} GPU(s), ONE batch of {per_rank3 * world} x 6 s queries", 10000, 14411,
                [(385, per_rank3)], 10, ctx, dev, rank, world, 16)
    per_len = max(1, int(os.environ.get("SWEEP_PER_LENGTH", "128")) // world)
    qlens = [(k, per_len) for k in (63, 143, 385, 707, 1111, 1514)]
    run_sharded(f"cfg4: 100k-track DB on {world} GPU(s), 2-20 s queries", 100000, 14411, qlens, 10, ctx, dev, rank, world, 12)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "sharded":
        return main_sharded()
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
