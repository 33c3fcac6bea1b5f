"""Multi-GPU matcher: ctypes binding of hpfw_shard_* (include/hpfw_b200.h, hpfw_b200/csrc/shard.cu) in rank mode — one
process per GPU, the reference database sharded by track, queries replicated, ONE in-place ncclAllGather of the per-rank
top-k keys and one merge kernel, all inside the library. NCCL is called by libhpfw_b200.so itself;
torch.distributed only carries the 128-byte NCCL id from rank 0 to the other ranks at start-up (any launcher-side channel
would do) and is not on the data path.

The reference has no multi-device path (storage.h:29 "TODO: maybe parallelize search"); semantics are those of the single
MemoryStorage: because a packed key (dist<<40 | global_track<<20 | offset) orders exactly like the reference's strict '<'
scan, the merged result is bit-identical for any number of shards.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import check
from .api import Context, ExtractionStream, HashprintExtractor, decode_keys, stream_arg, _ptr

KEY_OFFSET_BITS = 20
KEY_TRACK_BITS = 20
KEY_DIST_SHIFT = 40
KEY_NONE = np.uint64((1 << 64) - 1)


def pack_key(dist: int, track: int, offset: int) -> int:
    """= the kernels' key layout (include/hpfw_b200.h HPFW_KEY_*)."""
    return (int(dist) << KEY_DIST_SHIFT) | (int(track) << KEY_OFFSET_BITS) | int(offset)


def plan_shards(track_words: Sequence[int], n_shards: int, query_words: int = 385,
                speeds: Sequence[float] | None = None) -> List[Tuple[int, int]]:
    """Contiguous track ranges [begin, end) per shard, balanced by matcher work sum (n_r - k + 1) * k. Contiguous ranges keep
    the global track index = DB order, which the tie rule (earliest track) relies on. Computed by the library
    (hpfw_shard_plan, host-only: needs no GPU) so that Python and the C++ ShardedMemoryStorage split a DB identically."""
    lens = np.ascontiguousarray(track_words, dtype=np.int64)
    bounds = np.zeros(n_shards + 1, dtype=np.int32)
    if speeds is not None:
        # share of shard s = speeds[s] / sum(speeds): balance by measured device speed (hpfw_shard_plan_weighted)
        sp = np.ascontiguousarray(speeds, dtype=np.float64)
        if len(sp) != n_shards:
            raise ValueError("one speed per shard")
        check(_lib.load().hpfw_shard_plan_weighted(_ptr(lens) if len(lens) else None, len(lens), n_shards, query_words,
                                                   _ptr(sp), _ptr(bounds)))
    else:
        check(_lib.load().hpfw_shard_plan(_ptr(lens) if len(lens) else None, len(lens), n_shards, query_words, _ptr(bounds)))
    return [(int(bounds[i]), int(bounds[i + 1])) for i in range(n_shards)]


# ---- CPU-test plumbing (gloo): the same exchange pattern with torch tensors, used by tests/test_sharded_cpu.py only
def allgather_keys(local, group=None):
    """local: int64 tensor [Q, topk] -> [world, Q, topk]. The product path does this with ncclAllGather inside
    hpfw_shard_match_device; this torch version exists for the world-size-2 gloo tests on hosts without a GPU."""
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
    """In-place sum over the ranks (gloo in the CPU tests); the product path is hpfw_shard_allreduce_cov."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def exchange_nccl_id(rank: int, world: int, device=None, group=None) -> bytes:
    """Rank 0 creates the NCCL unique id (hpfw_shard_unique_id); torch.distributed carries its 128 bytes to the others."""
    import torch
    import torch.distributed as dist
    buf = (C.c_ubyte * 128)()
    if world > 1 and rank == 0:
        check(_lib.load().hpfw_shard_unique_id(buf))
    if world == 1:
        return bytes(buf)
    backend = dist.get_backend(group)
    dev = device if (backend == "nccl" and device is not None) else torch.device("cpu")
    t = torch.tensor(list(bytes(buf)), dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=0, group=group)
    return bytes(t.cpu().numpy().tobytes())


class ShardedMemoryStorage:
    """MemoryStorage semantics over a DB sharded across ranks (1 rank = 1 GPU); thin binding of a rank-mode hpfw_shard."""

    def __init__(self, ctx: Context, rank: int = 0, world: int = 1, group=None, nccl_id: bytes | None = None):
        self.ctx, self.rank, self.world, self.group = ctx, rank, world, group
        self._lib = ctx._lib
        if nccl_id is None:
            import torch
            nccl_id = exchange_nccl_id(rank, world, torch.device("cuda", ctx.device), group)
        idbuf = (C.c_ubyte * 128).from_buffer_copy(nccl_id)
        h = C.c_void_p()
        check(self._lib.hpfw_shard_create_rank(ctx.handle, rank, world, idbuf, C.byref(h)))
        self._h = h
        self.track_base = 0
        self._keys = None

    def build_local(self, words: np.ndarray, offsets: np.ndarray, track_base: int) -> "ShardedMemoryStorage":
        """This rank's shard (host buffers); track_base = global index of its first track."""
        words = np.ascontiguousarray(words, dtype=np.uint64)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        check(self._lib.hpfw_shard_build_rank(self._h, _ptr(words) if words.size else None, _ptr(offsets), len(offsets) - 1,
                                              track_base))
        self.track_base = track_base
        return self

    def build_local_device(self, d_words_ptr: int, offsets: np.ndarray, track_base: int, stream: int = 0):
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        check(self._lib.hpfw_shard_build_rank_device(self._h, C.c_void_p(d_words_ptr), _ptr(offsets), len(offsets) - 1,
                                                     track_base, stream_arg(stream)))
        self.track_base = track_base
        return self

    def match_device_ptr(self, d_qwords_ptr: int, qoffs: np.ndarray, topk: int, d_keys_out_ptr: int, stream: int = 0):
        """Everything in HBM and on `stream`: local match, in-place ncclAllGather, merge. Collective over the ranks."""
        qoffs = np.ascontiguousarray(qoffs, dtype=np.int64)
        check(self._lib.hpfw_shard_match_device(self._h, C.c_void_p(d_qwords_ptr), _ptr(qoffs), len(qoffs) - 1, topk,
                                                C.c_void_p(d_keys_out_ptr), stream_arg(stream)))

    def search_device(self, d_qwords, qoffs: np.ndarray, topk: int):
        """d_qwords: int64 CUDA tensor of the (replicated) query words. Returns the merged keys, int64 CUDA tensor
        [Q, topk], identical on every rank. Enqueued on torch's current stream; no host sync. torch only owns the buffers."""
        import torch
        nq = len(qoffs) - 1
        dev = d_qwords.device
        if self._keys is None or tuple(self._keys.shape) != (nq, topk) or self._keys.device != dev:
            self._keys = torch.empty((nq, topk), dtype=torch.int64, device=dev)
        self.match_device_ptr(d_qwords.data_ptr(), qoffs, topk, self._keys.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
        return self._keys

    def search_host(self, qwords_pinned, qoffs: np.ndarray, topk: int, out_pinned=None):
        """End-to-end call with HOST buffers: qwords_pinned = pinned int64 CPU tensor; returns a structured numpy array
        [Q, topk] (track, cnt, offset). Includes the H2D copy of the queries and the D2H copy of the result."""
        import torch
        dev = torch.device("cuda", self.ctx.device)
        with torch.cuda.device(dev):
            dq = qwords_pinned.to(dev, non_blocking=True)
            keys = self.search_device(dq, qoffs, topk)
            if out_pinned is None:
                out_pinned = torch.empty(keys.shape, dtype=torch.int64, pin_memory=True)
            out_pinned.copy_(keys, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()     # the stream of the context's device, whatever the caller's is
        return decode_keys(out_pinned.numpy().view(np.uint64))

    def allreduce_covariance(self, stream: int = 0) -> None:
        """Index-time filter learning over tracks split across the ranks: sums the contexts' 2420 x 2420 covariance
        accumulators in place with ONE ncclAllReduce (parallel_collector.h:94-97's mutex, across GPUs)."""
        check(self._lib.hpfw_shard_allreduce_cov(self._h, stream_arg(stream) if stream else None))

    def broadcast(self, d_ptr: int, nbytes: int, root: int = 0, stream: int = 0) -> None:
        check(self._lib.hpfw_shard_broadcast_device(self._h, C.c_void_p(d_ptr), nbytes, root,
                                                    stream_arg(stream) if stream else None))

    def allgatherv(self, d_send_ptr: int, d_recv_ptr: int, bytes_per_rank: Sequence[int], stream: int = 0) -> None:
        b = np.ascontiguousarray(bytes_per_rank, dtype=np.uint64)
        check(self._lib.hpfw_shard_allgatherv_device(self._h, C.c_void_p(d_send_ptr) if d_send_ptr else None,
                                                     C.c_void_p(d_recv_ptr), _ptr(b), stream_arg(stream) if stream else None))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.hpfw_shard_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def allreduce_covariance(ctx: Context, storage: "ShardedMemoryStorage", stream: int = 0):
    """Kept for callers of the r1 name: the all-reduce now runs inside the library on the accumulator itself."""
    storage.allreduce_covariance(stream)


class ShardedLiveSongIdentification:
    """hpfw::LiveSongIdentification::index()/search() (live_song_id.h:31-54) over several GPUs, one process per GPU.

    index(): the tracks are split into contiguous ranges (DB order = track order, as the tie rule needs); every rank pushes ITS
    tracks through its extraction stream (upload -> CQT -> covariance, spectrograms resident in HBM), ONE ncclAllReduce sums the
    covariance accumulators, rank 0's filters (calc_filters) are broadcast, every rank hashes its resident spectrograms in one
    batched launch and builds its shard device-to-device: between the audio upload and the database nothing visits the host per
    track (the 64 x 2420 filters do, once). search(): rank r extracts queries r, r+W, ... through its stream, one all-gather
    (variable sizes) hands every rank all query hashprints in HBM, then the sharded match. Results equal the single-GPU path's.

    Decoded mono buffers in (float32 or int16), like ParallelCollector::calc_hashprint's decoded side; file decoding is the
    caller's.
    """

    def __init__(self, ctx: Context, rank: int = 0, world: int = 1, group=None, nccl_id: bytes | None = None):
        self.ctx, self.rank, self.world, self.group = ctx, rank, world, group
        self.extractor = HashprintExtractor(ctx)
        self.storage = ShardedMemoryStorage(ctx, rank, world, group, nccl_id)
        self.ixs = ExtractionStream(ctx, slots=6, slot_bytes=16 << 20)
        self.qxs = ExtractionStream(ctx, slots=6, slot_bytes=1 << 20)
        self.names: List[str] = []
        self.filters = None

    def index(self, tracks: Sequence[np.ndarray], names: Sequence[str] | None = None, filters: np.ndarray | None = None):
        """tracks: ALL tracks of the collection on every rank (only this rank's range is touched). filters: skip the
        learning step and use these (e.g. cache/filters.cereal) — the reference's search-only mode."""
        import torch
        ex, xs = self.extractor, self.ixs
        n = len(tracks)
        self.names = [str(x) for x in (names if names is not None else range(n))]
        words = [max(ex.words(len(t)), 0) for t in tracks]
        a, b = plan_shards(words, self.world)[self.rank]
        dev = torch.device("cuda", self.ctx.device)
        xs.reset()
        learn = filters is None
        if learn:
            ex.cov_reset()
        for i in range(a, b):
            xs.submit(tracks[i], cov=learn)
        xs.wait()
        if learn:
            self.storage.allreduce_covariance()
            f = torch.zeros((2420, 64), dtype=torch.float32, device=dev)
            if self.rank == 0:
                f.copy_(torch.from_numpy(ex.calc_filters(install=False)[0]))
            torch.cuda.synchronize(dev)
            self.storage.broadcast(f.data_ptr(), f.numel() * 4, root=0)
            self.ctx.synchronize()
            filters = f.cpu().numpy()
        self.filters = np.ascontiguousarray(filters, dtype=np.float32)
        ex.set_filters(self.filters)
        xs.hash_kept()
        ptr, offs, lens = xs.hashprints_device()
        db_offs = np.zeros(b - a + 1, dtype=np.int64)
        np.cumsum(lens, out=db_offs[1:])
        assert np.array_equal(offs, db_offs[:-1])           # store order = this rank's track order, contiguous
        self.storage.build_local_device(ptr, db_offs, track_base=a)
        xs.drop_kept()
        return self

    def search(self, queries: Sequence[np.ndarray], topk: int = 1):
        """queries: decoded query buffers (the same list on every rank). Returns a structured array [Q, topk] of (track, cnt,
        offset) with GLOBAL track indices; self.names[track] is the reference's SearchResult::filename."""
        import torch
        ex, xs = self.extractor, self.qxs
        nq = len(queries)
        words = np.array([max(ex.words(len(q)), 0) for q in queries], dtype=np.int64)
        # rank-major processing order: rank r owns queries r, r+W, ...; the gathered buffer holds rank 0's queries first
        order = [i for r in range(self.world) for i in range(r, nq, self.world)]
        per_rank = [int(sum(words[i] for i in range(r, nq, self.world))) for r in range(self.world)]
        xs.reset()
        for i in range(self.rank, nq, self.world):
            xs.submit(queries[i])
        xs.hash_kept()
        ptr, offs, lens = xs.hashprints_device()
        dev = torch.device("cuda", self.ctx.device)
        total = int(words.sum())
        allq = torch.empty(max(total, 1), dtype=torch.int64, device=dev)
        torch.cuda.synchronize(dev)
        self.storage.allgatherv(ptr if per_rank[self.rank] else 0, allq.data_ptr(), [8 * w for w in per_rank])
        qoffs = np.zeros(nq + 1, dtype=np.int64)
        np.cumsum(words[order], out=qoffs[1:])
        keys = torch.empty((nq, topk), dtype=torch.int64, device=dev)
        self.storage.match_device_ptr(allq.data_ptr(), qoffs, topk, keys.data_ptr())
        self.ctx.synchronize()
        got = decode_keys(keys.cpu().numpy().view(np.uint64))
        out = np.zeros_like(got)
        out[np.asarray(order)] = got                         # back to the caller's query order
        return out
