"""Host-side mirror of the reference's interfaces for the hot path, over the C ABI.

  MemoryStorage   <- hpfw::db::MemoryStorage<Collector>  (include/hpfw/audioproblems/live-song-id/storage.h:8-93)
  Context         <- one GPU (no equivalent in the CPU-only reference)
Names, argument meaning and error behaviour follow the reference; the filename of a SearchResult is looked up from the
names given to build().
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Iterable, List, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import Match, check

SIZE_MAX = (1 << 64) - 1


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def stream_arg(stream: int):
    """cudaStream_t for the C ABI. The ABI reads NULL as "the context's own stream", so the CUDA legacy default stream
    (handle 0, e.g. torch.cuda.current_stream().cuda_stream when no stream is set) is passed as cudaStreamLegacy (0x1)."""
    return C.c_void_p(int(stream) if stream else 1)


class Context:
    """One CUDA device. There is no CPU fallback: construction fails without a B200-class GPU."""

    def __init__(self, device: int = 0):
        self._lib = _lib.load()
        h = C.c_void_p()
        check(self._lib.hpfw_ctx_create(device, C.byref(h)))
        self._h = h
        self.device = device

    @property
    def handle(self):
        return self._h

    def launch_count(self) -> int:
        return int(self._lib.hpfw_ctx_launch_count(self._h))

    def synchronize(self) -> None:
        check(self._lib.hpfw_ctx_synchronize(self._h))

    def timing_enable(self, on: bool = True) -> None:
        check(self._lib.hpfw_ctx_timing_enable(self._h, 1 if on else 0))

    def timing_read(self, kernel: int, reset: bool = True):
        """(total device ms, launches) of a kernel class (_lib.K_MATCH, ...) since the last reset."""
        ms, n = C.c_double(), C.c_uint64()
        check(self._lib.hpfw_ctx_timing_read(self._h, kernel, C.byref(ms), C.byref(n), 1 if reset else 0))
        return ms.value, int(n.value)

    def microbench_pipes(self):
        out = (C.c_double * 3)()
        clk = C.c_double()
        check(self._lib.hpfw_microbench_pipes(self._h, out, C.byref(clk)))
        return {"popc_per_clk_sm": out[0], "lop3_per_clk_sm": out[1], "wordops_per_clk_sm": out[2],
                "sm_clock_mhz": clk.value}

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.hpfw_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


@dataclass
class SearchResult:
    """= MemoryStorage::SearchResult {filename, cnt, offset} (storage.h:11-15); `track` is the DB index (-1: none)."""
    filename: str
    cnt: int
    offset: int
    track: int = -1


def pack(hashprints: Sequence[np.ndarray]) -> Tuple[np.ndarray, np.ndarray]:
    offs = np.zeros(len(hashprints) + 1, dtype=np.int64)
    for i, h in enumerate(hashprints):
        offs[i + 1] = offs[i] + len(h)
    if len(hashprints) and offs[-1] > 0:
        words = np.ascontiguousarray(np.concatenate([np.asarray(h, dtype=np.uint64) for h in hashprints]))
    else:
        words = np.zeros(0, dtype=np.uint64)
    return words, offs


class MemoryStorage:
    """GPU-resident hashprint database with the reference's build()/find() semantics (storage.h:21-64).

    build() takes the collector's output, a list of (filename, hashprint) pairs in DB order (the reference moves a
    tbb::concurrent_vector<FilenameFingerprintPair> into a std::vector; DB order = arrival order).
    """

    def __init__(self, ctx: Context):
        self.ctx = ctx
        self._lib = ctx._lib
        self._db = None
        self.filenames: List[str] = []
        self.track_base = 0

    def build(self, pairs: Iterable[Tuple[str, np.ndarray]], track_base: int = 0) -> "MemoryStorage":
        pairs = list(pairs)
        self.filenames = [p[0] for p in pairs]
        words, offs = pack([p[1] for p in pairs])
        return self.build_packed(words, offs, self.filenames, track_base)

    def build_packed(self, words: np.ndarray, offsets: np.ndarray, filenames=None, track_base: int = 0):
        self._free()
        words = np.ascontiguousarray(words, dtype=np.uint64)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = len(offsets) - 1
        self.filenames = list(filenames) if filenames is not None else [str(track_base + i) for i in range(n)]
        self.track_base = track_base
        h = C.c_void_p()
        check(self._lib.hpfw_db_build(self.ctx.handle, _ptr(words) if words.size else None, _ptr(offsets), n,
                                      track_base, C.byref(h)))
        self._db = h
        self._offsets = offsets
        return self

    def build_device(self, d_words_ptr: int, offsets: np.ndarray, stream: int = 0, filenames=None, track_base: int = 0):
        """words already in HBM (e.g. a torch tensor's data_ptr())."""
        self._free()
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = len(offsets) - 1
        self.filenames = list(filenames) if filenames is not None else [str(track_base + i) for i in range(n)]
        self.track_base = track_base
        h = C.c_void_p()
        check(self._lib.hpfw_db_build_device(self.ctx.handle, C.c_void_p(d_words_ptr), _ptr(offsets), n, track_base,
                                             stream_arg(stream), C.byref(h)))
        self._db = h
        self._offsets = offsets
        return self

    # ---- reference API ----
    def find(self, hp: np.ndarray) -> SearchResult:
        """MemoryStorage::find (storage.h:27-64): best track, its Hamming distance and alignment offset."""
        hp = np.ascontiguousarray(hp, dtype=np.uint64)
        m = Match()
        check(self._lib.hpfw_db_find(self._require(), _ptr(hp) if hp.size else None, len(hp), C.byref(m)))
        return self._result(m)

    # ---- batched / top-k (notebook semantics, liveid.ipynb:909-927) ----
    def find_topk(self, queries: Sequence[np.ndarray], topk: int = 10) -> List[List[SearchResult]]:
        qwords, qoffs = pack(list(queries))
        nq = len(qoffs) - 1
        out = (Match * (nq * topk))()
        check(self._lib.hpfw_db_find_topk(self._require(), _ptr(qwords) if qwords.size else None, _ptr(qoffs), nq, topk,
                                          out))
        return [[self._result(out[q * topk + r]) for r in range(topk)] for q in range(nq)]

    def find_topk_packed(self, qwords: np.ndarray, qoffs: np.ndarray, topk: int = 10) -> np.ndarray:
        """Same, returning a structured array [nq, topk] with fields track/cnt/offset (no Python object per result)."""
        qwords = np.ascontiguousarray(qwords, dtype=np.uint64)
        qoffs = np.ascontiguousarray(qoffs, dtype=np.int64)
        nq = len(qoffs) - 1
        out = np.zeros((nq, topk), dtype=np.dtype([("track", "<i8"), ("cnt", "<u8"), ("offset", "<i8")]))
        check(self._lib.hpfw_db_find_topk(self._require(), _ptr(qwords) if qwords.size else None, _ptr(qoffs), nq, topk,
                                          out.ctypes.data_as(C.POINTER(Match))))
        return out

    def match_device(self, d_qwords_ptr: int, qoffs: np.ndarray, topk: int, d_keys_out_ptr: int, stream: int = 0):
        """Device path: query words and the key output live in HBM; enqueues on `stream` without synchronising."""
        qoffs = np.ascontiguousarray(qoffs, dtype=np.int64)
        check(self._lib.hpfw_db_match_device(self._require(), C.c_void_p(d_qwords_ptr), _ptr(qoffs), len(qoffs) - 1,
                                             topk, C.c_void_p(d_keys_out_ptr), stream_arg(stream)))

    def word_ops(self, qoffs: np.ndarray) -> float:
        qoffs = np.ascontiguousarray(qoffs, dtype=np.int64)
        return float(self._lib.hpfw_db_word_ops(self._require(), _ptr(qoffs), len(qoffs) - 1))

    @property
    def n_tracks(self) -> int:
        return int(self._lib.hpfw_db_tracks(self._require()))

    def _result(self, m: Match) -> SearchResult:
        if m.track < 0:
            return SearchResult("", SIZE_MAX, 0, -1)   # storage.h:28 initial value
        local = m.track - self.track_base
        name = self.filenames[local] if 0 <= local < len(self.filenames) else str(m.track)
        return SearchResult(name, int(m.cnt), int(m.offset), int(m.track))

    def _require(self):
        if self._db is None:
            raise RuntimeError("MemoryStorage: build() has not been called")
        return self._db

    def _free(self):
        if self._db is not None:
            self._lib.hpfw_db_destroy(self._db)
            self._db = None

    def __del__(self):
        try:
            self._free()
        except Exception:
            pass


def decode_keys(keys: np.ndarray) -> np.ndarray:
    """Packed keys (dist<<40 | track<<20 | offset) -> structured array with track/cnt/offset."""
    L = _lib.load()
    keys = np.ascontiguousarray(keys, dtype=np.uint64)
    out = np.zeros(keys.shape, dtype=np.dtype([("track", "<i8"), ("cnt", "<u8"), ("offset", "<i8")]))
    L.hpfw_keys_decode(_ptr(keys), keys.size, out.ctypes.data_as(C.POINTER(Match)))
    return out


class HashprintExtractor:
    """Stages 1-3 on one GPU: the decoded-buffer side of ParallelCollector (parallel_collector.h:54-59, 115-137)."""

    def __init__(self, ctx: Context):
        self.ctx = ctx
        self._lib = ctx._lib

    def set_filters(self, filters_cm: np.ndarray) -> None:
        """filters_cm: float32 memory of the reference's column-major 64 x 2420 Filters, i.e. numpy shape [2420, 64]."""
        f = np.ascontiguousarray(filters_cm, dtype=np.float32)
        if f.size != 64 * 2420:
            raise ValueError("filters must hold 64 x 2420 floats")
        check(self._lib.hpfw_set_filters(self.ctx.handle, _ptr(f)))

    def cols(self, n_samples: int) -> int:
        return int(self._lib.hpfw_cqt_cols(n_samples))

    def words(self, n_samples: int) -> int:
        return int(self._lib.hpfw_hashprint_words_for_samples(n_samples))

    def spectrogram(self, audio: np.ndarray, magnitude: bool = False) -> np.ndarray:
        """spectrum::CQT::spectrogram on a decoded mono buffer -> float32 [cols, 121] (dB, or linear magnitudes)."""
        a = np.ascontiguousarray(audio, dtype=np.float32)
        out = np.zeros((self.cols(len(a)), 121), dtype=np.float32)
        got = C.c_int()
        fn = self._lib.hpfw_cqt_magnitude if magnitude else self._lib.hpfw_cqt_spectrogram
        check(fn(self.ctx.handle, _ptr(a), len(a), _ptr(out), C.byref(got)))
        return out

    def hashprint_from_spectrogram(self, spec_tm: np.ndarray) -> np.ndarray:
        s = np.ascontiguousarray(spec_tm, dtype=np.float32)
        hp = np.zeros(max(s.shape[0] - 99, 1), dtype=np.uint64)
        n = C.c_int()
        check(self._lib.hpfw_hashprint_from_spectrogram(self.ctx.handle, _ptr(s), s.shape[0], _ptr(hp), C.byref(n)))
        return hp[:n.value]

    def calc_hashprint(self, audio: np.ndarray) -> np.ndarray:
        """ParallelCollector::calc_hashprint on a decoded buffer (host in, host out)."""
        a = np.ascontiguousarray(audio, dtype=np.float32)
        hp = np.zeros(max(self.words(len(a)), 1), dtype=np.uint64)
        n = C.c_int()
        check(self._lib.hpfw_calc_hashprint_audio(self.ctx.handle, _ptr(a), len(a), _ptr(hp), C.byref(n)))
        return hp[:n.value]

    # ---- index-time filter learning (HashprintHandle::calc_cov / calc_filters, hashprint_handle.h:96-112) ----
    def cov_reset(self) -> None:
        check(self._lib.hpfw_cov_reset(self.ctx.handle))

    def cov_add_spectrogram(self, spec_tm: np.ndarray) -> None:
        """accum_cov += calc_cov(calc_frames(spectrogram)^T) (parallel_collector.h:92-97), accumulator resident in HBM."""
        s = np.ascontiguousarray(spec_tm, dtype=np.float32)
        check(self._lib.hpfw_cov_add_spectrogram(self.ctx.handle, _ptr(s), s.shape[0]))

    def cov_get(self) -> np.ndarray:
        out = np.zeros((2420, 2420), dtype=np.float32)
        check(self._lib.hpfw_cov_get(self.ctx.handle, _ptr(out)))
        return out

    def cov_set(self, accum: np.ndarray) -> None:
        a = np.ascontiguousarray(accum, dtype=np.float32)
        if a.shape != (2420, 2420):
            raise ValueError("accum_cov must be 2420 x 2420")
        check(self._lib.hpfw_cov_set(self.ctx.handle, _ptr(a)))

    def calc_filters(self, cov: np.ndarray | None = None, install: bool = True):
        """Top-64 eigenvectors of `cov` (or of the context's accumulator) -> (filters [2420, 64] = memory of the
        column-major 64 x 2420 Filters, eigenvalues[64]). Sign: largest-|component| of each filter positive."""
        f = np.zeros((2420, 64), dtype=np.float32)
        w = np.zeros(64, dtype=np.float32)
        c = None if cov is None else np.ascontiguousarray(cov, dtype=np.float32)
        check(self._lib.hpfw_calc_filters(self.ctx.handle, None if c is None else _ptr(c), _ptr(f), _ptr(w)))
        if install:
            self.set_filters(f)
        return f, w

    def calc_hashprint_pcm16(self, pcm: np.ndarray) -> np.ndarray:
        """calc_hashprint on 16-bit PCM samples (host in, host out): half the PCIe bytes of the float call, same hashprint
        as calc_hashprint(pcm / 32768)."""
        a = np.ascontiguousarray(pcm, dtype=np.int16)
        hp = np.zeros(max(self.words(len(a)), 1), dtype=np.uint64)
        n = C.c_int()
        check(self._lib.hpfw_calc_hashprint_pcm16(self.ctx.handle, _ptr(a), len(a), _ptr(hp), C.byref(n)))
        return hp[:n.value]

    def calc_hashprint_pcm16_batch_device(self, d_pcm_ptr: int, sample_offsets: np.ndarray, d_hp_out_ptr: int,
                                          stream: int = 0) -> None:
        """calc_hashprint_batch_device for int16 PCM already in HBM (converted to float on the device)."""
        so = np.ascontiguousarray(sample_offsets, dtype=np.int64)
        check(self._lib.hpfw_calc_hashprint_pcm16_batch_device(self.ctx.handle, C.c_void_p(d_pcm_ptr), _ptr(so),
                                                               len(so) - 1, C.c_void_p(d_hp_out_ptr), stream_arg(stream)))

    def calc_hashprint_batch_device(self, d_audio_ptr: int, sample_offsets: np.ndarray, d_hp_out_ptr: int,
                                    stream: int = 0) -> None:
        """Many tracks already in HBM (concatenated, even offsets) -> concatenated hashprints in HBM; no host sync."""
        so = np.ascontiguousarray(sample_offsets, dtype=np.int64)
        check(self._lib.hpfw_calc_hashprint_audio_batch_device(self.ctx.handle, C.c_void_p(d_audio_ptr), _ptr(so),
                                                               len(so) - 1, C.c_void_p(d_hp_out_ptr), stream_arg(stream)))


XS_PCM16, XS_COV = 1, 2


class ExtractionStream:
    """hpfw_xs_*: the device-resident pipeline behind ParallelCollector::prepare / LiveSongIdentification::search
    (include/hpfw_b200.h "extraction stream"; hpfw_b200/csrc/xstream.cu). Decoded buffers go into pinned staging slots, every
    submit enqueues H2D -> [int16 -> float] -> CQT (-> covariance accumulate) and keeps the dB spectrogram in HBM; hash_kept()
    runs the batched projection; the hashprints stay on the device."""

    def __init__(self, ctx: Context, slots: int = 8, slot_bytes: int = 1 << 20):
        self.ctx, self._lib = ctx, ctx._lib
        h = C.c_void_p()
        check(self._lib.hpfw_xs_create(ctx.handle, slots, slot_bytes, C.byref(h)))
        self._h = h

    def submit(self, samples: np.ndarray, cov: bool = False) -> int:
        """samples: mono float32 or int16. Returns the track index in the stream (submission order)."""
        a = np.ascontiguousarray(samples)
        if a.dtype == np.int16:
            flags = XS_PCM16
        else:
            a = np.ascontiguousarray(a, dtype=np.float32)
            flags = 0
        slot, host = C.c_int(), C.c_void_p()
        check(self._lib.hpfw_xs_acquire(self._h, a.nbytes, C.byref(slot), C.byref(host)))
        C.memmove(host, a.ctypes.data, a.nbytes)
        track = C.c_int()
        check(self._lib.hpfw_xs_submit(self._h, slot.value, len(a), flags | (XS_COV if cov else 0), C.byref(track)))
        return track.value

    def wait(self) -> None:
        check(self._lib.hpfw_xs_wait(self._h))

    def hash_kept(self) -> None:
        check(self._lib.hpfw_xs_hash_kept(self._h))

    def reset(self) -> None:
        check(self._lib.hpfw_xs_reset(self._h))

    def drop_kept(self) -> None:
        check(self._lib.hpfw_xs_drop_kept(self._h))

    @property
    def n_tracks(self) -> int:
        return int(self._lib.hpfw_xs_tracks(self._h))

    def hashprints_device(self):
        """(device pointer of the hashprint store, offsets[n], lengths[n]) in store order; offset -1 = not hashed."""
        n = self.n_tracks
        ptr = C.c_void_p()
        offs, lens = np.zeros(n, dtype=np.int64), np.zeros(n, dtype=np.int64)
        check(self._lib.hpfw_xs_hashprints_device(self._h, C.byref(ptr), _ptr(offs), _ptr(lens)))
        return int(ptr.value or 0), offs, lens

    def hashprint_host(self, track: int) -> np.ndarray:
        cols, words, res = C.c_int(), C.c_int(), C.c_int()
        check(self._lib.hpfw_xs_track_info(self._h, track, C.byref(cols), C.byref(words), C.byref(res)))
        out = np.zeros(words.value, dtype=np.uint64)
        check(self._lib.hpfw_xs_hashprint_host(self._h, track, _ptr(out)))
        return out

    def fetch_spectrogram(self, track: int) -> np.ndarray:
        cols, words, res = C.c_int(), C.c_int(), C.c_int()
        check(self._lib.hpfw_xs_track_info(self._h, track, C.byref(cols), C.byref(words), C.byref(res)))
        out = np.zeros((cols.value, 121), dtype=np.float32)
        check(self._lib.hpfw_xs_fetch_spectrogram(self._h, track, _ptr(out)))
        return out

    def build_db(self, storage: "MemoryStorage", order: Sequence[int], filenames=None, track_base: int = 0) -> "MemoryStorage":
        """MemoryStorage from hashed tracks, device-to-device; DB index i = stream track order[i]."""
        storage._free()
        o = np.ascontiguousarray(order, dtype=np.int32)
        h = C.c_void_p()
        check(self._lib.hpfw_xs_build_db(self._h, _ptr(o), len(o), track_base, C.byref(h)))
        storage._db = h
        storage.track_base = track_base
        storage.filenames = list(filenames) if filenames is not None else [str(track_base + i) for i in range(len(o))]
        return storage

    def match(self, storage: "MemoryStorage", topk: int) -> np.ndarray:
        """All tracks of the stream as queries (store order) -> structured array [n, topk] (track, cnt, offset)."""
        n = self.n_tracks
        out = np.zeros((n, topk), dtype=np.dtype([("track", "<i8"), ("cnt", "<u8"), ("offset", "<i8")]))
        check(self._lib.hpfw_xs_match(self._h, storage._require(), topk, out.ctypes.data_as(C.POINTER(Match))))
        return out

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.hpfw_xs_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
