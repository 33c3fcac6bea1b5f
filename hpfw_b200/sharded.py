"""Multi-GPU matcher: the reference database sharded by track across ranks (one process per GPU), queries replicated, one
all-gather of the per-rank top-k keys, one merge kernel.

The reference has no multi-device path (storage.h:29 "TODO: maybe parallelize search"); semantics are those of the single
MemoryStorage: because a packed key (dist<<40 | global_track<<20 | offset) orders exactly like the reference's strict '<'
scan, the merged result is bit-identical for any number of shards.

torch / torch.distributed are plumbing here: device buffers, the NCCL process group and the all-gather call.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Tuple

import numpy as np

from ._lib import check
from .api import Context, MemoryStorage, decode_keys, stream_arg

KEY_OFFSET_BITS = 20
KEY_TRACK_BITS = 20
KEY_DIST_SHIFT = 40
KEY_NONE = np.uint64((1 << 64) - 1)


def pack_key(dist: int, track: int, offset: int) -> int:
    """= the kernels' key layout (include/hpfw_b200.h HPFW_KEY_*)."""
    return (int(dist) << KEY_DIST_SHIFT) | (int(track) << KEY_OFFSET_BITS) | int(offset)


def plan_shards(track_words: Sequence[int], n_shards: int, query_words: int = 385) -> List[Tuple[int, int]]:
    """Contiguous track ranges [begin, end) per shard, balanced by matcher work sum (n_r - k + 1) * k (clamped at n_r >= 1
    offset). Contiguous ranges keep the global track index = DB order, which the tie rule (earliest track) relies on."""
    lens = np.asarray(track_words, dtype=np.int64)
    k = np.minimum(lens, query_words)
    work = (lens - k + 1) * np.maximum(k, 1)
    total = float(work.sum())
    csum = np.concatenate([[0], np.cumsum(work)]).astype(np.float64)
    bounds = [0]
    for s in range(1, n_shards):
        target = total * s / n_shards
        b = int(np.searchsorted(csum, target, side="left"))
        # pick the closer of b-1 / b, never before the previous bound
        if b > 0 and abs(csum[b - 1] - target) <= abs(csum[min(b, len(lens))] - target):
            b -= 1
        bounds.append(min(max(b, bounds[-1]), len(lens)))
    bounds.append(len(lens))
    return [(bounds[i], bounds[i + 1]) for i in range(n_shards)]


def allgather_keys(local, group=None):
    """local: int64 tensor [Q, topk] (packed keys viewed as int64) -> [world, Q, topk] on the same device. One collective
    per query batch: NCCL over NVLink on GPUs, gloo in the CPU tests."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    out = torch.empty((world,) + tuple(local.shape), dtype=local.dtype, device=local.device)
    if world == 1:
        out[0].copy_(local)
    else:
        dist.all_gather_into_tensor(out.view(-1), local.contiguous().view(-1), group=group)
    return out


def allreduce_sum(t, group=None):
    """In-place sum of a tensor over the ranks (NCCL over NVLink on GPUs, gloo in the CPU tests); a no-op for one rank."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def allreduce_covariance(ctx: Context, group=None, stream: int = 0):
    """Index-time filter learning over tracks that are split across the ranks: every rank has added the
    sample covariances of ITS tracks to its context's accumulator (HashprintExtractor.cov_add_spectrogram*, i.e.
    HashprintHandle::calc_cov + the accumulate of ParallelCollector::preprocess, parallel_collector.h:92-97); this sums the
    2420 x 2420 accumulators with ONE all-reduce and writes the sum back, so that calc_filters() on any rank returns the
    filters of the whole collection (the reference's later division by the track count does not change eigenvectors)."""
    import torch
    dev = torch.device("cuda", ctx.device)
    acc = torch.empty(2420 * 2420, dtype=torch.float32, device=dev)
    s = stream or torch.cuda.current_stream(dev).cuda_stream
    check(ctx._lib.hpfw_cov_get_device(ctx.handle, C.c_void_p(acc.data_ptr()), stream_arg(s)))
    allreduce_sum(acc, group)
    check(ctx._lib.hpfw_cov_set_device(ctx.handle, C.c_void_p(acc.data_ptr()), stream_arg(s)))
    return acc


class ShardedMemoryStorage:
    """MemoryStorage semantics over a DB sharded across the ranks of a torch.distributed group (1 rank = 1 GPU)."""

    def __init__(self, ctx: Context, rank: int = 0, world: int = 1, group=None):
        self.ctx, self.rank, self.world, self.group = ctx, rank, world, group
        self.local = MemoryStorage(ctx)
        self.track_base = 0
        self._keys_local = None
        self._keys_all = None
        self._keys_merged = None

    def build_local(self, words: np.ndarray, offsets: np.ndarray, track_base: int) -> "ShardedMemoryStorage":
        """This rank's shard (host buffers); track_base = global index of its first track."""
        self.local.build_packed(words, offsets, track_base=track_base)
        self.track_base = track_base
        return self

    def build_local_device(self, d_words_ptr: int, offsets: np.ndarray, track_base: int, stream: int = 0):
        self.local.build_device(d_words_ptr, offsets, stream=stream, track_base=track_base)
        self.track_base = track_base
        return self

    def search_device(self, d_qwords, qoffs: np.ndarray, topk: int):
        """d_qwords: int64 CUDA tensor of the (replicated) query words. Returns the merged keys, int64 CUDA tensor
        [Q, topk], identical on every rank. Everything is enqueued on torch's current stream; no host sync."""
        import torch
        nq = len(qoffs) - 1
        dev = d_qwords.device
        if self._keys_local is None or tuple(self._keys_local.shape) != (nq, topk):
            self._keys_local = torch.empty((nq, topk), dtype=torch.int64, device=dev)
            self._keys_merged = torch.empty((nq, topk), dtype=torch.int64, device=dev)
        s = torch.cuda.current_stream(dev).cuda_stream
        self.local.match_device(d_qwords.data_ptr(), qoffs, topk, self._keys_local.data_ptr(), s)
        if self.world == 1:
            return self._keys_local
        allk = allgather_keys(self._keys_local, self.group)
        check(self.ctx._lib.hpfw_topk_merge_device(self.ctx.handle, C.c_void_p(allk.data_ptr()), self.world, nq, topk,
                                                   C.c_void_p(self._keys_merged.data_ptr()), stream_arg(s)))
        self._keys_all = allk     # keep alive until the stream has consumed it
        return self._keys_merged

    def search_host(self, qwords_pinned, qoffs: np.ndarray, topk: int, out_pinned=None):
        """End-to-end call with HOST buffers: qwords_pinned = pinned int64 CPU tensor; returns a structured numpy array
        [Q, topk] (track, cnt, offset). Includes the H2D copy of the queries and the D2H copy of the result."""
        import torch
        dq = qwords_pinned.to(f"cuda:{self.ctx.device}", non_blocking=True)
        keys = self.search_device(dq, qoffs, topk)
        if out_pinned is None:
            out_pinned = torch.empty(keys.shape, dtype=torch.int64, pin_memory=True)
        out_pinned.copy_(keys, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return decode_keys(out_pinned.numpy().view(np.uint64))


class ShardedLiveSongIdentification:
    """hpfw::LiveSongIdentification::index()/search() (live_song_id.h:31-54) over several GPUs, one process per GPU.

    index(): the tracks are split into contiguous ranges (DB order = track order, as the tie rule needs); every rank runs
    the CQT of ITS tracks, adds their covariances to its accumulator, ONE all-reduce gives every rank the collection's
    covariance, rank 0's filters (calc_filters) are broadcast, and every rank hashes its own tracks into its shard of the
    database — the hashprints never leave the GPU that computed them. search(): every rank extracts the (short) queries,
    matches them against its shard, one all-gather of the top-k keys, merge. Results equal the single-GPU path's.

    Decoded mono float32 buffers in, like ParallelCollector::calc_hashprint's decoded side; file decoding is the caller's.
    """

    def __init__(self, ctx: Context, rank: int = 0, world: int = 1, group=None):
        from .api import HashprintExtractor
        self.ctx, self.rank, self.world, self.group = ctx, rank, world, group
        self.extractor = HashprintExtractor(ctx)
        self.storage = ShardedMemoryStorage(ctx, rank, world, group)
        self.names: List[str] = []
        self.filters = None

    def index(self, tracks: Sequence[np.ndarray], names: Sequence[str] | None = None, filters: np.ndarray | None = None):
        """tracks: ALL tracks of the collection on every rank (only this rank's range is touched). filters: skip the
        learning step and use these (e.g. cache/filters.cereal) — the reference's search-only mode."""
        import torch
        import torch.distributed as dist
        ex = self.extractor
        n = len(tracks)
        self.names = [str(x) for x in (names if names is not None else range(n))]
        words = [max(ex.words(len(t)), 0) for t in tracks]
        a, b = plan_shards(words, self.world)[self.rank]
        dev = torch.device("cuda", self.ctx.device)
        specs = [ex.spectrogram(np.ascontiguousarray(tracks[i], dtype=np.float32)) for i in range(a, b)]
        if filters is None:
            ex.cov_reset()
            for sp in specs:
                ex.cov_add_spectrogram(sp)
            allreduce_covariance(self.ctx, self.group)
            f = torch.zeros((2420, 64), dtype=torch.float32, device=dev)
            if self.rank == 0:
                f.copy_(torch.from_numpy(ex.calc_filters(install=False)[0]))
            if self.world > 1:
                dist.broadcast(f, src=0, group=self.group)
            filters = f.cpu().numpy()
        self.filters = np.ascontiguousarray(filters, dtype=np.float32)
        ex.set_filters(self.filters)
        hps = [ex.hashprint_from_spectrogram(sp) for sp in specs]
        offs = np.zeros(len(hps) + 1, dtype=np.int64)
        np.cumsum([len(h) for h in hps], out=offs[1:])
        flat = np.concatenate(hps) if hps else np.zeros(0, dtype=np.uint64)
        self.storage.build_local(flat, offs, track_base=a)
        return self

    def search(self, queries: Sequence[np.ndarray], topk: int = 1):
        """queries: decoded query buffers (the same list on every rank). Rank r extracts the hashprints of queries r, r+W,
        r+2W, ...; one all-gather hands every rank all of them (their word counts follow from the sample counts, so every
        rank knows the layout); then the sharded search. Returns a structured array [Q, topk] of (track, cnt, offset) with
        GLOBAL track indices; self.names[track] is the reference's SearchResult::filename."""
        import torch
        import torch.distributed as dist
        ex = self.extractor
        nq = len(queries)
        words = [max(ex.words(len(q)), 0) for q in queries]
        qoffs = np.zeros(nq + 1, dtype=np.int64)
        np.cumsum(words, out=qoffs[1:])
        mine = list(range(self.rank, nq, self.world))
        local = [ex.calc_hashprint(np.ascontiguousarray(queries[i], dtype=np.float32)) for i in mine]
        flat = np.zeros(int(qoffs[-1]), dtype=np.uint64)
        if self.world == 1:
            for i, h in zip(mine, local):
                flat[qoffs[i]:qoffs[i + 1]] = h
        else:
            per_rank = [sum(words[i] for i in range(r, nq, self.world)) for r in range(self.world)]
            pad = max(max(per_rank), 1)
            dev = torch.device("cuda", self.ctx.device)
            buf = np.zeros(pad, dtype=np.uint64)
            if local:
                cat = np.concatenate(local)
                buf[:len(cat)] = cat
            t_local = torch.from_numpy(buf.view(np.int64)).to(dev)
            t_all = torch.empty((self.world, pad), dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(t_all.view(-1), t_local, group=self.group)
            allw = t_all.cpu().numpy().view(np.uint64)
            for r in range(self.world):
                pos = 0
                for i in range(r, nq, self.world):
                    flat[qoffs[i]:qoffs[i + 1]] = allw[r, pos:pos + words[i]]
                    pos += words[i]
        pinned = torch.from_numpy(flat.view(np.int64).copy()).pin_memory()
        return self.storage.search_host(pinned, qoffs, topk)
