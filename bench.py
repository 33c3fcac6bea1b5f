#!/usr/bin/env python
"""bench.py — live-ID queries/s against a 10k-track hashprint DB (BASELINE.json metric), B200 vs the reference CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W]             # this repo's CUDA path
    python bench.py --impl reference [--gpus N] --steps K --warmup W # the reference's MemoryStorage::find on host cores
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ... # N>1: one rank per GPU, DB sharded by track

Workload (config.workload): DB = 10,000 tracks x 14,411 hashprint words (3-min tracks), synthetic iid words;
one STEP = one batch of 128*N queries of 385 words (6 s), each a DB slice with 25 % of its bits flipped, matched at every
alignment offset of every track (Hamming cross-correlation), top-10 per query. The DB is sharded by track over the N
GPUs, queries are replicated, per-rank top-k keys are all-gathered and merged inside the library (hpfw_shard_match_device:
ncclAllGather in place + merge kernel; torch.distributed only bootstraps the NCCL id and the barriers). Per-GPU work per step
is constant in N ("weak"); a fixed-size batch and single-query latency are reported in `strong`.

  value   device-resident: query words already in HBM when the timed region starts
  e2e     through the host API (ShardedMemoryStorage.search_host): pinned host query words -> H2D -> match -> (allgather,
          merge) -> D2H of the top-k records, every step
  roofline  the dominant kernel — match_tc_kernel<1> (exact +-1 GEMM on the tensor cores, fp4 operands, f32 accumulators)
          against the fp4 tensor-pipe ceiling; roofline_popc: the integer-pipe match_kernel on the same step against the
          POPC/LOP3 pipe roof; durations from CUDA events around every launch
  e2e_cpp   the reference's own C++ API timed from WAV files by examples/cpp/bench-liveid.cpp (rank 0): queries/s of
          LiveSongIdentification::search() against the same 10k-track DB and frames/s of index()
  cpu_baseline  the reference's own MemoryStorage::find (oracle/_ref, compiled from /root/reference headers) on all host
          cores over a bounded sample, on rank 0 at N=1 only
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TRACKS = 10_000
TRACK_WORDS = 14_411        # 3 min @ 44.1 kHz
QUERY_WORDS = 385           # 6 s
QUERIES_PER_GPU = 128
TOPK = 10
FLIP = 0.25
AUDIO_TRACKS = 8            # DB tracks whose hashprints come from synthetic audio (end-to-end arm)
METRIC = "live_id_queries_per_sec_vs_10k_track_db"
UNIT = "queries/s"
WORDOPS_PER_CLK_SM = 16     # carry-save matcher: min(64 LOP3 lanes / 4, 16 POPC lanes / 1) per clk per SM
POPC32_PER_CLK_SM = 16      # CUDA programming guide arithmetic-throughput table (population count); checked by the microbenchmark
I8_OPS_PER_CLK_SM = 16384   # tcgen05.mma kind::i8, M=128 N=256 K=32 in 128 clk (B300_MICROARCH.md pacing law): 8192 MAC/clk/SM
F4_OPS_PER_CLK_SM = 32768   # tcgen05.mma kind::mxf4, M=128 N=256 K=64 in 128 clk: 16384 MAC/clk/SM (nominal 9 PFLOP/s fp4 dense)
OPS_PER_WORDOP = 128        # tensor-core matcher: one 64-bit word-op = 64 s8 multiply-adds
NCU_TC_TRAFFIC_RATIO = 243.07 / 233.8   # measured DRAM bytes / algorithmic bytes of match_tc_kernel<1> (profiles/r03d_*)
NCU_CQT_TRAFFIC_PER_TRACK = 268.3e6     # measured DRAM bytes per 3-min track over a whole 48-track batch (profiles/r2q_*, r2n)


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def word_ops_per_query(tracks=TRACKS, n=TRACK_WORDS, k=QUERY_WORDS) -> float:
    return float(tracks) * float(n - k + 1) * float(k)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                                          "50", "-i", str(self.device)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------ reference arm
def host_db(seed: int, tracks: int):
    rng = np.random.default_rng(seed)
    words = rng.integers(0, 1 << 64, size=tracks * TRACK_WORDS, dtype=np.uint64)
    offs = np.arange(tracks + 1, dtype=np.int64) * TRACK_WORDS
    return words, offs


def host_queries(seed: int, words, offs, nq: int):
    rng = np.random.default_rng(seed)
    tracks = len(offs) - 1
    tr = rng.integers(0, tracks, size=nq)
    off = rng.integers(0, TRACK_WORDS - QUERY_WORDS + 1, size=nq)
    q = np.empty(nq * QUERY_WORDS, dtype=np.uint64)
    for i in range(nq):
        sl = words[offs[tr[i]] + off[i]: offs[tr[i]] + off[i] + QUERY_WORDS].copy()
        noise = np.zeros(QUERY_WORDS, dtype=np.uint64)
        for b in range(64):
            noise |= (rng.random(QUERY_WORDS) < FLIP).astype(np.uint64) << np.uint64(b)
        q[i * QUERY_WORDS:(i + 1) * QUERY_WORDS] = sl ^ noise
    qoffs = np.arange(nq + 1, dtype=np.int64) * QUERY_WORDS
    return q, qoffs, np.stack([tr, off], axis=1)


def cpu_find_batch(words, offs, q, qoffs, cores):
    """The reference's MemoryStorage::find over host threads (oracle/_ref) or, if that library is absent, the C port."""
    import oracle
    if oracle.ref_available():
        tr, d, o = oracle.ref_find_batch(words, offs, q, qoffs, cores)
        return "reference", tr, d, o
    tr, d, o = oracle.find_topk_batch(words, offs, q, qoffs, 1, cores)
    return "port", tr[:, 0], d[:, 0], o[:, 0]


def cpu_sample(cores: int, sample_tracks: int, nq: int, seed: int = 7, tracks: int = TRACKS):
    """Time one bounded sample: nq queries x sample_tracks tracks; returns (kind, seconds, queries/s scaled to `tracks`)."""
    words, offs = host_db(seed, sample_tracks)
    q, qoffs, truth = host_queries(seed + 1, words, offs, nq)
    t0 = time.perf_counter()
    kind, tr, d, o = cpu_find_batch(words, offs, q, qoffs, cores)
    dt = time.perf_counter() - t0
    assert np.array_equal(tr, truth[:, 0]) and np.array_equal(o, truth[:, 1]), "CPU baseline returned wrong matches"
    qps = nq / dt * (sample_tracks / tracks)       # find() is linear in the number of reference tracks
    return kind, dt, qps


def cpu_wordops_per_core(qps: float, cores: int, tracks: int = TRACKS) -> float:
    """word-ops (XOR64 + popcount64) per second per host core behind a queries/s figure quoted on a `tracks`-track DB:
    comparable across boxes with different core counts."""
    return qps * word_ops_per_query(tracks) / max(1, cores)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = host_cores()
    sample_tracks = 1000
    nq = max(cores, 8)
    times = []
    kind = "reference"
    for i in range(args.warmup + args.steps):
        kind, dt, qps = cpu_sample(cores, sample_tracks, nq, seed=11 + i, tracks=args.tracks)
        if i >= args.warmup:
            times.append((dt, qps))
    dt = float(np.mean([t[0] for t in times]))
    qps = float(np.mean([t[1] for t in times]))
    sample = (f"each step: {nq} queries x {sample_tracks}-track subset of the DB on {cores} host threads, "
              f"scaled x{sample_tracks}/{args.tracks} (find is linear in tracks); MemoryStorage::find compiled from the "
              f"reference headers with -Ofast -march=native (no MKL involved in this path)")
    line = {
        "impl": "reference", "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": workload_config(args.gpus, args.tracks, args.queries_per_gpu),
        "cpu_baseline": {"value": qps, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                         "wordops_per_s_per_core": cpu_wordops_per_core(qps, cores, args.tracks)},
        "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(n_gpus: int, tracks: int = TRACKS, qpg: int = QUERIES_PER_GPU):
    return {"workload": f"live-id search: {tracks}-track DB x {TRACK_WORDS} words (3-min tracks), "
                        f"{qpg}*N x {QUERY_WORDS}-word (6 s) queries per step, full-offset Hamming "
                        f"cross-correlation, top-{TOPK}",
            "tracks": tracks, "track_words": TRACK_WORDS, "query_words": QUERY_WORDS,
            "queries_per_step": qpg * n_gpus, "topk": TOPK,
            "parallelism": f"db-shard x{n_gpus} (tracks), queries replicated, in-library ncclAllGather of top-k keys + merge kernel",
            "l2": "DB shard per GPU (>=144 MB) exceeds the 126 MB L2; no flush needed"}


# ------------------------------------------------------------------------------------------------------ extraction leg
EXTRACT_SECONDS = 180.0
EXTRACT_SR = 44100


def cpu_extraction_sample(cores: int, n_tracks: int, filters):
    """CPU path for stages 1-3 on `n_tracks` 3-min tracks over `cores` threads: oracle/nsgcq.py (numpy restatement of the
    essentia CQT: the reference's own CQT cannot run here) + the reference's calc_frames / filters*frames /
    calc_fingerprint / fingerprint_to_hashprint (oracle/_ref, Eigen GEBP, no MKL)."""
    import oracle
    from oracle import nsgcq
    from concurrent.futures import ThreadPoolExecutor
    n = int(EXTRACT_SECONDS * EXTRACT_SR)
    rng = np.random.default_rng(3)
    base = (0.1 * rng.standard_normal(n)).astype(np.float32)
    t = np.arange(n) / EXTRACT_SR
    for f in (220.0, 440.0, 554.37, 1318.5):
        base += (0.2 * np.sin(2 * np.pi * f * t)).astype(np.float32)
    kind = "reference" if oracle.ref_available() else "port"

    def one(i):
        spec = nsgcq.spectrogram(np.roll(base, 1000 * i))
        if kind == "reference":
            return len(oracle.ref_hashprint_from_spectrogram(spec, filters)) + 80
        return len(oracle.hashprint_from_spectrogram(spec, filters)) + 80

    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=cores) as ex:
        frames = sum(ex.map(one, range(n_tracks)))
    dt = time.perf_counter() - t0
    return kind, dt, frames / dt


def run_extraction(args, ctx, rank, world, dev, max_over_ranks, barrier, st):
    """BASELINE.json configs[1]: hashprint extraction only, 1000 synthetic 3-min tracks (split over the ranks), CQT + 64-filter
    projection, frames/s. Audio resident in HBM for `value`; `e2e` streams every track from pinned host memory."""
    import torch
    import hpfw_b200
    from hpfw_b200 import _lib
    n = int(EXTRACT_SECONDS * EXTRACT_SR)
    per_rank = max(1, args.extract_tracks // world)
    g = np.load(os.path.join(ROOT, "tests", "golden", "hashprint.npz"))
    filters = np.ascontiguousarray(g["filters"])
    ex = hpfw_b200.HashprintExtractor(ctx)
    ex.set_filters(filters)
    cols, words = ex.cols(n), ex.words(n)
    frames = cols - 19
    gen = torch.Generator(device=dev)
    gen.manual_seed(777 + rank)
    tt = torch.arange(n, device=dev, dtype=torch.float32) / EXTRACT_SR
    nbase = 16
    base = torch.empty((nbase, n), dtype=torch.float32, device=dev)
    for b in range(nbase):
        x = 0.05 * torch.randn(n, device=dev, generator=gen)
        for _ in range(6):
            semis = int(torch.randint(0, 49, (1,), device=dev, generator=gen).item())
            f0 = 130.81 * 2.0 ** (semis / 12.0)
            rate = float(torch.rand(1, device=dev, generator=gen).item()) * 2.0 + 0.5
            x += 0.15 * torch.sin(2 * np.pi * f0 * tt) * (0.5 + 0.5 * torch.sin(2 * np.pi * rate * tt))
        base[b] = x
    audio = torch.empty((per_rank, n), dtype=torch.float32, device=dev)
    for i in range(per_rank):
        audio[i] = torch.roll(base[i % nbase], shifts=9973 * (i // nbase))
    del tt
    hp = torch.empty(per_rank * words, dtype=torch.int64, device=dev)
    offs = np.arange(per_rank + 1, dtype=np.int64) * n
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        ex.calc_hashprint_batch_device(audio.data_ptr(), offs, hp.data_ptr(), stream)

    step()
    barrier()
    ctx.timing_read(_lib.K_CQT, reset=True)
    ctx.timing_read(_lib.K_PROJECT, reset=True)
    ctx.timing_enable(True)
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(1, args.steps)
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / reps
    launches = (ctx.launch_count() - l0) // reps
    cq_ms, cq_n = ctx.timing_read(_lib.K_CQT, reset=True)
    pj_ms, pj_n = ctx.timing_read(_lib.K_PROJECT, reset=True)
    ctx.timing_enable(False)
    value = per_rank * world * frames / (ms * 1e-3)

    # e2e: pinned host audio -> H2D (copy stream, double-buffered) -> CQT+projection -> D2H of the hashprint, every track
    npin = 4
    pin = [torch.empty(n, dtype=torch.float32, pin_memory=True) for _ in range(npin)]
    for k in range(npin):
        pin[k].copy_(audio[k % per_rank])
    stage = [torch.empty(n, dtype=torch.float32, device=dev) for _ in range(2)]
    hp_dev = [torch.empty(words, dtype=torch.int64, device=dev) for _ in range(2)]
    hp_host = [torch.empty(words, dtype=torch.int64, pin_memory=True) for _ in range(2)]
    copy_s, comp_s = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]
    e2e_tracks = min(per_rank, 200)
    import ctypes as C
    from hpfw_b200.api import stream_arg
    from hpfw_b200._lib import check
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_tracks):
        k = i & 1
        with torch.cuda.stream(copy_s):
            copy_s.wait_event(freed[k])
            stage[k].copy_(pin[i % npin], non_blocking=True)
            ready[k].record(copy_s)
        with torch.cuda.stream(comp_s):
            comp_s.wait_event(ready[k])
            check(ctx._lib.hpfw_calc_hashprint_audio_device(ctx.handle, C.c_void_p(stage[k].data_ptr()), n,
                                                            C.c_void_p(hp_dev[k].data_ptr()), stream_arg(comp_s.cuda_stream)))
            freed[k].record(comp_s)
            hp_host[k].copy_(hp_dev[k], non_blocking=True)
    torch.cuda.synchronize()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = e2e_tracks * world * frames / e2e_s

    # the same with 16-bit PCM in pinned host memory (what a WAV file / decoder delivers; converted to float on the device)
    pin16 = [torch.empty(n, dtype=torch.int16, pin_memory=True) for _ in range(npin)]
    for k in range(npin):
        pin16[k].copy_((audio[k % per_rank] * 16384.0).clamp_(-32768, 32767).to(torch.int16))
    stage16 = [torch.empty(n, dtype=torch.int16, device=dev) for _ in range(2)]
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_tracks):
        k = i & 1
        with torch.cuda.stream(copy_s):
            copy_s.wait_event(freed[k])
            stage16[k].copy_(pin16[i % npin], non_blocking=True)
            ready[k].record(copy_s)
        with torch.cuda.stream(comp_s):
            comp_s.wait_event(ready[k])
            check(ctx._lib.hpfw_calc_hashprint_pcm16_device(ctx.handle, C.c_void_p(stage16[k].data_ptr()), n,
                                                            C.c_void_p(hp_dev[k].data_ptr()), stream_arg(comp_s.cuda_stream)))
            freed[k].record(comp_s)
            hp_host[k].copy_(hp_dev[k], non_blocking=True)
    torch.cuda.synchronize()
    barrier()
    e2e16_s = max_over_ranks(time.perf_counter() - t0)
    e2e16_value = e2e_tracks * world * frames / e2e16_s
    del pin16, stage16

    # host -> device copy rate of this box (the bound of the e2e extraction legs above)
    pin_big = torch.empty(256 << 20, dtype=torch.uint8, pin_memory=True)
    dev_big = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    best_copy = None
    for _ in range(5):
        e0.record()
        dev_big.copy_(pin_big, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1)
        best_copy = t if best_copy is None else min(best_copy, t)
    pcie_gbs = (256 << 20) / (best_copy * 1e-3) / 1e9 if best_copy else None
    del pin_big, dev_big
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            pk = json.load(f)
        hbm, sm_max, src = float(pk["hbm_gbs"]), float(pk["sm_max_mhz"]), "measured (MEASURED_PEAKS.json)"
    except (OSError, KeyError, ValueError):
        hbm, sm_max, src = 6650.0, 1965.0, "fallback (B200_PROFILING.md)"
    # tracks of a batch run concurrently on 4 streams, so the summed kernel durations overlap: the CQT stage time per track is
    # the step time minus the (serial) projection launches
    try:
        f16_peak = float(pk["bf16_tflops_sustained"])
        f16_src = ("the measured sustained bf16 cuBLAS rate (MEASURED_PEAKS.json; fp16 and bf16 share the tensor rate). The kernel "
                   "itself is bound by shared-memory operand reads: N = 64 filters means 6 KB of operands per 32-clk MMA, 192 "
                   "B/clk against the 128 B/clk an SM delivers (67 % of the tensor rate at best); the pre-pass that writes the fp16 "
                   "difference matrix (HBM-bound, ~2 us per track) is inside this time")
    except (NameError, KeyError, ValueError):
        f16_peak, f16_src = 1590.0, "the fallback bf16 rate (B200_PROFILING.md)"
    cq_per_track = max(1e-6, (ms * reps - pj_ms) / max(1, per_rank * reps))
    cq_kernel_sum_per_track = cq_ms / max(1, per_rank * reps)
    cqt_bytes = 4.0 * n + 4.0 * 121 * cols
    out = {
        "metric": "hashprint_frames_per_sec", "value": value, "unit": "frames/s",
        "workload": f"hashprint extraction only: {per_rank * world} synthetic 3-min tracks @44.1 kHz "
                    f"({per_rank} per GPU), CQT + 64-filter projection + pack; audio resident in HBM",
        "ms_per_track": ms / per_rank, "frames_per_track": frames, "gpu_launches_per_step": launches,
        "e2e": {"value": e2e_value, "unit": "frames/s", "tracks": e2e_tracks * world,
                "h2d_bytes_per_track": 4 * n, "d2h_bytes_per_track": 8 * words,
                "note": "pinned host audio, H2D on a copy stream double-buffered against compute, hashprint D2H per track",
                "pcm16": {"value": e2e16_value, "unit": "frames/s", "h2d_bytes_per_track": 2 * n,
                          "note": "the same from 16-bit PCM in pinned host memory (hpfw_calc_hashprint_pcm16_device: "
                                  "sample / 32768 on the device)"}},
        "roofline": {"bound": "hbm", "kernel": "CQT (7 kernels per track, cqt.cu)", "achieved": cqt_bytes / (cq_per_track * 1e-3) / 1e9,
                     "peak": hbm, "unit": "GB/s", "frac": cqt_bytes / (cq_per_track * 1e-3) / 1e9 / hbm,
                     "traffic": NCU_CQT_TRAFFIC_PER_TRACK,
                     "traffic_note": "per track: dram__bytes_read+write of a whole 48-track batch (8 lanes) from ncu --replay-mode "
                                     "app-range, profiles/r2q_cqt_batch_range_lanes8.csv: 150 MB read + 118 MB written; two FFT "
                                     "passes + three chirp-z passes, and 8 lanes of scratch (720 MB) cannot stay in the 126 MB L2; "
                                     "one track in flight: 134 MB; the six kernels captured one by one with cold caches: 144 MB "
                                     "(profiles/r01y_cqt_kernels_ncu_full.md)",
                     "ms_per_track": cq_per_track, "kernel_ms_sum_per_track_overlapped": cq_kernel_sum_per_track,
                     "peak_how": src,
                     "algorithmic_bytes_per_track": cqt_bytes},
        "roofline_e2e": {"bound": "pcie", "what": "extraction.e2e.pcm16: every track crosses PCIe once as 16-bit PCM",
                         "achieved": 2.0 * n * e2e_tracks / e2e16_s / 1e9, "peak": pcie_gbs, "unit": "GB/s",
                         "frac": 2.0 * n * e2e_tracks / e2e16_s / 1e9 / pcie_gbs if pcie_gbs else None,
                         "peak_how": "pinned host -> device copy of 256 MiB measured in this run (cudaMemcpyAsync, best of 5)",
                         "float32": {"achieved": 4.0 * n * e2e_tracks / e2e_s / 1e9,
                                     "frac": 4.0 * n * e2e_tracks / e2e_s / 1e9 / pcie_gbs if pcie_gbs else None}},
        "roofline_projection": {"bound": "tensor", "kernel": "project_tc_multi_kernel<2> (tcgen05 kind::f16, fp16 operands, 2 tiles per CTA share the filter stream, one issuing thread per tile) + tc_delta_kernel pre-pass",
                                "achieved": 2.0 * 64 * 2420 * frames * per_rank * reps / (pj_ms * 1e-3) / 1e12,
                                "peak": f16_peak, "unit": "TFLOP/s",
                                "ms_per_track": pj_ms / max(1, per_rank * reps), "launches": pj_n,
                                "peak_how": f16_src},
    }
    out["roofline_projection"]["frac"] = out["roofline_projection"]["achieved"] / out["roofline_projection"]["peak"]
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        nt = max(4, cores)
        kind, dt, fps = cpu_extraction_sample(cores, nt, filters)
        out["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind,
                               "sample": f"{nt} synthetic 3-min tracks in {dt:.1f} s on {cores} threads: numpy/pocketfft "
                                         f"restatement of the essentia CQT (oracle/nsgcq.py; essentia itself is absent) + the "
                                         f"reference's frames/projection/fingerprint code (Eigen GEBP, no MKL)"}
    # ---- index-time filter learning (row a10: calc_cov per track + calc_filters once), rank 0 only: the covariance is
    # accumulated in HBM per GPU; a multi-GPU index would add one all-reduce of 2420 x 2420 floats (not part of this leg)
    if rank == 0:
        n_idx = min(per_rank, 32)
        spec = torch.empty((n_idx, cols, 121), dtype=torch.float32, device=dev)
        for i in range(n_idx):
            check(ctx._lib.hpfw_cqt_spectrogram_device(ctx.handle, C.c_void_p(audio[i].data_ptr()), n,
                                                       C.c_void_p(spec[i].data_ptr()), stream_arg(stream)))
        torch.cuda.synchronize()

        def cov_pass():
            check(ctx._lib.hpfw_cov_reset(ctx.handle))
            for i in range(n_idx):
                check(ctx._lib.hpfw_cov_add_spectrogram_device(ctx.handle, C.c_void_p(spec[i].data_ptr()), cols,
                                                               stream_arg(stream)))
        cov_pass()
        torch.cuda.synchronize()
        e0.record()
        cov_pass()
        e1.record()
        torch.cuda.synchronize()
        cov_ms = e0.elapsed_time(e1) / n_idx
        f_out = np.zeros((2420, 64), dtype=np.float32)
        ev = np.zeros(64, dtype=np.float32)
        t0 = time.perf_counter()
        check(ctx._lib.hpfw_calc_filters(ctx.handle, None, f_out.ctypes.data_as(C.c_void_p), ev.ctypes.data_as(C.c_void_p)))
        filt_ms = (time.perf_counter() - t0) * 1e3
        ex.set_filters(filters)          # the timed legs above and the CPU sample below use the golden filters
        syrk_flop = 2.0 * 2420 * 2420 * frames
        out["index"] = {
            "what": "index-time filter learning, rows a10: HashprintHandle::calc_cov per track (accumulated in HBM) and "
                    "calc_filters once (top-64 eigenvectors of the 2420 x 2420 covariance)",
            "cov_ms_per_track": cov_ms, "tracks": n_idx,
            "cov_equivalent_tflops": syrk_flop / (cov_ms * 1e-3) / 1e12,
            "cov_note": "the reference forms the 2420 x 2420 x frames SYRK (170 GFLOP per 3-min track); the context window "
                        "makes it twenty 121 x 121 x frames correlations + edge corrections (4.2 GFLOP), learn.cu",
            "calc_filters_ms": filt_ms,
            "calc_filters_note": "block subspace iteration on the GPU + Rayleigh-Ritz on the host, wall clock",
        }
        if world == 1 and not args.no_cpu_baseline:
            import oracle
            if oracle.ref_available():
                R = oracle.ref()
                sp = np.ascontiguousarray(spec[0].cpu().numpy())
                cov = np.zeros((2420, 2420), dtype=np.float32)
                t0 = time.perf_counter()
                R.ref_calc_cov(sp.reshape(-1), cols, cov.reshape(-1))
                cpu_cov_s = time.perf_counter() - t0
                fl = np.zeros((2420, 64), dtype=np.float32)
                t0 = time.perf_counter()
                R.ref_calc_filters(cov.reshape(-1), fl.reshape(-1))
                cpu_filt_s = time.perf_counter() - t0
                out["index"]["cpu_baseline"] = {
                    "kind": "reference", "cores": 1, "cov_ms_per_track": cpu_cov_s * 1e3, "calc_filters_ms": cpu_filt_s * 1e3,
                    "sample": "one 3-min spectrogram through the reference's calc_cov and one calc_filters "
                              "(hashprint_handle.h:96-112, Eigen 3.3.7 without MKL), single thread as in one taskflow worker"}
        del spec
    if world > 1:
        # multi-GPU index: every rank accumulates the covariance of its own tracks; ONE all-reduce of the 2420 x 2420
        # accumulator (NCCL over NVLink, 23 MB) gives every rank the collection's covariance (sharded.allreduce_covariance)
        st.allreduce_covariance(stream)
        barrier()
        e0.record()
        st.allreduce_covariance(stream)
        e1.record()
        barrier()
        ar_ms = max_over_ranks(e0.elapsed_time(e1))
        if rank == 0 and "index" in out:
            out["index"]["cov_allreduce_ms"] = ar_ms
            out["index"]["cov_allreduce_note"] = (f"one in-place ncclAllReduce (sum) of the covariance accumulator over {world} "
                                                  "ranks inside the library (hpfw_shard_allreduce_cov), once per index() call")
    del audio, base, hp
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------------ C++ API leg
def build_cpp_bench():
    exe = os.path.join(ROOT, "examples", "cpp", "bench-liveid")
    src = os.path.join(ROOT, "examples", "cpp", "bench-liveid.cpp")
    libdir = os.path.join(ROOT, "hpfw_b200")
    deps = [src, os.path.join(libdir, "libhpfw_b200.so")]
    for dp, _, fns in os.walk(os.path.join(ROOT, "include")):
        deps += [os.path.join(dp, fn) for fn in fns]
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-I" + os.path.join(ROOT, "include"), src, "-o", exe,
                               "-L" + libdir, "-lhpfw_b200", "-lpthread", "-Wl,-rpath," + libdir])
    return exe


def cpp_index_leg(exe, work, env, dev, world, n_tracks, reps=3):
    """LiveSongIdentification::index() over n_tracks three-minute PCM16 WAV files in `work`/tracks (32 distinct synthetic
    signals; the other names are links to them, so every name is a separate track of the DB)."""
    import torch
    from hpfw_b200 import bench_data
    idir = os.path.join(work, "tracks")
    os.makedirs(idir)
    n = int(EXTRACT_SECONDS * EXTRACT_SR)
    gen = torch.Generator(device=dev)
    gen.manual_seed(4242)
    tt = torch.arange(n, device=dev, dtype=torch.float32) / EXTRACT_SR
    uniq = min(32, n_tracks)
    for b in range(uniq):
        x = 0.05 * torch.randn(n, device=dev, generator=gen)
        for _ in range(6):
            semis = int(torch.randint(0, 49, (1,), device=dev, generator=gen).item())
            f0 = 130.81 * 2.0 ** (semis / 12.0)
            rate = float(torch.rand(1, device=dev, generator=gen).item()) * 2.0 + 0.5
            x += 0.15 * torch.sin(2 * np.pi * f0 * tt) * (0.5 + 0.5 * torch.sin(2 * np.pi * rate * tt))
        pcm = (x * 16384.0).clamp_(-32768, 32767).to(torch.int16).cpu().numpy()
        bench_data.write_wav_pcm16(os.path.join(idir, f"song{b:05d}.wav"), pcm, EXTRACT_SR)
    for i in range(uniq, n_tracks):
        os.symlink(os.path.join(idir, f"song{i % uniq:05d}.wav"), os.path.join(idir, f"song{i:05d}.wav"))
    del tt
    p = subprocess.run([exe, "index-sharded" if world > 1 else "index", "tracks", str(reps)], cwd=work, env=env,
                       capture_output=True, text=True, timeout=900)
    recs = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    if p.returncode != 0 or not recs:
        return {"error": f"rc={p.returncode}", "stderr_tail": p.stderr[-600:]}
    out = json.loads(recs[-1]) | {
        "what": f"LiveSongIdentification::index() over {n_tracks} three-minute PCM16 WAV files on tmpfs: "
                "decode threads -> pinned ring -> H2D -> CQT -> covariance (filters learned, as the reference does) "
                "-> batched projection -> DB built device-to-device; cache/spectros written in the background",
        "host_threads": host_cores()}
    if env.get("HPFW_TRACE"):
        out["trace"] = [ln for ln in p.stderr.splitlines() if "[hpfw trace]" in ln]
    return out


def run_cpp_leg(args, ctx, ex, dev, world, tracks, audio_tracks, audio_hps, h_audio, src_track, filters, e2e_value):
    """The reference's own C++ API from WAV files (rank 0): hpfw::LiveSongIdentification::search() over the 128 query WAVs
    against the same database the Python arm uses (loaded from a MemoryStorage dump), and ::index() over 3-minute PCM16 WAVs.
    Files live on tmpfs; the ranks of an N > 1 run idle at a barrier meanwhile and the binary uses db::ShardedMemoryStorage
    over the N GPUs from ONE process."""
    import shutil
    import tempfile
    import torch
    from hpfw_b200 import bench_data
    exe = build_cpp_bench()
    # tmpfs when it has room (WAVs 0.5 GB + DB dump 1.2 GB + 7 MB of cached spectrogram per indexed track), else the default
    # temporary directory; the index leg shrinks to what fits
    need = lambda n_idx: (2.0 + 0.0075 * n_idx) * 1e9
    base = None
    for cand in ("/dev/shm", tempfile.gettempdir()):
        if os.path.isdir(cand) and os.access(cand, os.W_OK):
            free = shutil.disk_usage(cand).free
            if free > need(64):
                base = cand
                while args.cpp_index_tracks > 64 and free < need(args.cpp_index_tracks):
                    args.cpp_index_tracks //= 2
                break
    if base is None:
        return {"skipped": "no writable directory with 3 GB free for the WAV files"}
    work = tempfile.mkdtemp(prefix="hpfw_cpp_", dir=base)
    out = {"binary": "examples/cpp/bench-liveid (g++ -O2, include/hpfw/*.h over libhpfw_b200.so)", "work_dir": work}
    env = dict(os.environ, HPFW_NUM_GPUS=str(world))
    try:
        # ---- search leg: same DB content as the Python arm (synthetic words; the first tracks carry the audio hashprints)
        os.makedirs(os.path.join(work, "cache", "spectros"))
        os.makedirs(os.path.join(work, "queries"))
        bench_data.write_matrix_cereal(os.path.join(work, "cache", "filters.cereal"), 64, 2420, filters)
        rng = np.random.default_rng(4321)
        names = [f"track{i:05d}" for i in range(tracks)]
        with open(os.path.join(work, "db.cereal"), "wb") as f:
            f.write(np.uint64(tracks).tobytes())
            for i in range(tracks):
                hp = rng.integers(0, 1 << 63, size=TRACK_WORDS, dtype=np.int64).view(np.uint64)
                if i < len(audio_hps):
                    hp[:len(audio_hps[i])] = audio_hps[i]
                b = names[i].encode()
                f.write(np.uint64(len(b)).tobytes() + b + np.uint64(TRACK_WORDS).tobytes())
                f.write(hp.tobytes())
        nqf = h_audio.shape[0]
        expect = []
        for q in range(nqf):
            pcm = (h_audio[q].numpy() * 32768.0 * 0.9).round().clip(-32768, 32767).astype(np.int16)
            fn = f"q{q:04d}_{names[int(src_track[q])]}.wav"
            bench_data.write_wav_pcm16(os.path.join(work, "queries", fn), pcm, 44100)
            expect.append(f"{fn} {names[int(src_track[q])]}")
        with open(os.path.join(work, "expect.txt"), "w") as f:
            f.write("\n".join(expect) + "\n")
        cmd = "search-sharded" if world > 1 else "search"
        p = subprocess.run([exe, cmd, "db.cereal", "queries", str(max(2, args.steps)), "expect.txt"], cwd=work, env=env,
                           capture_output=True, text=True, timeout=900)
        recs = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
        if p.returncode != 0 or not recs:
            out["search"] = {"error": f"rc={p.returncode}", "stderr_tail": p.stderr[-600:]}
        else:
            r = json.loads(recs[-1])
            out["search"] = r | {
                "what": "LiveSongIdentification::search() over 128 six-second PCM16 WAV files (decode threads -> pinned ring "
                        "-> H2D -> CQT -> projection -> match -> printed results), one process, "
                        f"{'db::ShardedMemoryStorage over ' + str(world) + ' GPUs' if world > 1 else 'db::MemoryStorage'}",
                "vs_python_e2e": r["queries_per_s"] / e2e_value if world == 1 else None,
                "vs_python_e2e_note": "same 128 queries per batch, same DB size; N > 1: the C++ batch is 128 queries over N "
                                      "GPUs (strong), the Python step 128*N (weak): not comparable"}
        os.remove(os.path.join(work, "db.cereal"))
        # ---- index leg: 3-minute PCM16 WAVs (32 distinct signals, the other names are links to them)
        if args.cpp_index_tracks > 0:
            shutil.rmtree(os.path.join(work, "cache"))
            out["index"] = cpp_index_leg(exe, work, env, dev, world, args.cpp_index_tracks)
    finally:
        shutil.rmtree(work, ignore_errors=True)
    return out


# ------------------------------------------------------------------------------------------------------ CUDA arm
def run_cuda(args):
    import torch
    import torch.distributed as dist
    import hpfw_b200
    from hpfw_b200 import _lib, synth
    from hpfw_b200.sharded import ShardedMemoryStorage

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} needs torchrun --nproc-per-node {args.gpus} (one rank per GPU)")
        args.gpus = world
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # stdout carries exactly one JSON line: everything any library writes to fd 1 meanwhile (NCCL prints its version banner
    # there) is sent to stderr; the JSON line goes to the saved descriptor at the end
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        # a host-side group for the one long wait (rank 0 runs the C++ leg): an NCCL barrier would park a spinning kernel on
        # every other GPU, and the C++ process uses those GPUs
        cpu_group = dist.new_group(backend="gloo")
    ctx = hpfw_b200.Context(local_rank)

    # ---- synthetic DB shard in HBM + planted queries (setup, untimed)
    n = args.gpus
    tracks = args.tracks
    qpg = args.queries_per_gpu
    st = ShardedMemoryStorage(ctx, rank, world)
    bounds = [r * tracks // n for r in range(n + 1)]
    balance = None
    if world > 1 and args.balance:
        # The GPUs of one node do not run at the same clock under their power caps, and the all-gather of a step waits for the
        # slowest rank. Calibration (setup, untimed): an equal split is matched three times, every rank reports the device
        # time of its match kernel, and the track ranges are re-planned in proportion to the measured speeds
        # (hpfw_shard_plan_weighted). The timed DB below is generated for the re-planned ranges.
        from hpfw_b200.sharded import plan_shards
        clo, chi = bounds[rank], bounds[rank + 1]
        cw, coffs = synth.device_hashprint_db(torch, dev, 4321 + rank, chi - clo, TRACK_WORDS)
        cq, _, _ = synth.device_hashprint_queries(torch, cw, coffs, 55 + rank, qpg * n, QUERY_WORDS, FLIP)
        cqo = np.arange(qpg * n + 1, dtype=np.int64) * QUERY_WORDS
        st.build_local_device(cw.data_ptr(), coffs, track_base=clo, stream=torch.cuda.current_stream().cuda_stream)
        st.search_device(cq, cqo, TOPK)
        torch.cuda.synchronize()
        dist.barrier()
        ctx.timing_read(_lib.K_MATCH_TC, reset=True)
        ctx.timing_enable(True)
        for _ in range(3):
            st.search_device(cq, cqo, TOPK)
        torch.cuda.synchronize()
        t_ms, t_n = ctx.timing_read(_lib.K_MATCH_TC, reset=True)
        ctx.timing_enable(False)
        t_all = torch.zeros(world, dtype=torch.float64, device=dev)
        t_all[rank] = t_ms / max(1, t_n)
        dist.all_reduce(t_all)
        times = t_all.cpu().numpy()
        speeds = 1.0 / np.maximum(times, 1e-6)
        sh = plan_shards([TRACK_WORDS] * tracks, n, QUERY_WORDS, speeds=speeds)
        bounds = [a for a, _ in sh] + [sh[-1][1]]
        balance = {"how": "track ranges proportional to 1 / (match kernel time of an equal split), measured per rank before "
                          "the timed DB is built (hpfw_shard_plan_weighted)",
                   "calibration_ms_per_rank": [float(x) for x in times],
                   "tracks_per_rank": [bounds[r + 1] - bounds[r] for r in range(n)]}
        del cw, cq
        torch.cuda.empty_cache()
    lo, hi = bounds[rank], bounds[rank + 1]
    d_words, offs = synth.device_hashprint_db(torch, dev, 1234 + rank, hi - lo, TRACK_WORDS)
    # audio-derived part of the DB for the end-to-end arm: the first AUDIO_TRACKS tracks of the DB start with the hashprints
    # of synthetic 30 s tracks (extracted on the GPU); the e2e queries are noisy, pitch-shifted 6 s slices of those tracks
    golden = np.load(os.path.join(ROOT, "tests", "golden", "hashprint.npz"))
    ex = hpfw_b200.HashprintExtractor(ctx)
    ex.set_filters(np.ascontiguousarray(golden["filters"]))
    audio_tracks = [synth.synth_track(900 + i, 30.0, 44100) for i in range(AUDIO_TRACKS)]
    audio_hps = []
    if lo == 0:
        for i, a in enumerate(audio_tracks):
            hp_i = ex.calc_hashprint(a)
            audio_hps.append(hp_i)
            d_words[i * TRACK_WORDS: i * TRACK_WORDS + len(hp_i)] = torch.from_numpy(hp_i.view(np.int64)).to(dev)
    d_q_local, _, truth_local = synth.device_hashprint_queries(torch, d_words, offs, 99 + rank, qpg, QUERY_WORDS, FLIP)
    truth_local[:, 0] += lo
    if world > 1:
        d_q = torch.empty((world, qpg * QUERY_WORDS), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(d_q.view(-1), d_q_local.contiguous())
        d_q = d_q.view(-1)
        t_all = torch.empty((world, qpg, 2), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(t_all.view(-1), torch.from_numpy(truth_local).to(dev).view(-1))
        truth = t_all.view(-1, 2).cpu().numpy()
    else:
        d_q, truth = d_q_local, truth_local
    nq = qpg * n
    qoffs = np.arange(nq + 1, dtype=np.int64) * QUERY_WORDS
    stream = torch.cuda.current_stream().cuda_stream
    st.build_local_device(d_words.data_ptr(), offs, track_base=lo, stream=stream)
    del d_words
    h_q = torch.empty(d_q.shape, dtype=torch.int64, pin_memory=True)
    h_q.copy_(d_q)
    h_out = torch.empty((nq, TOPK), dtype=torch.int64, pin_memory=True)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- pipe microbenchmark (rank 0, untimed): pins the POPC roof
    micro = ctx.microbench_pipes() if rank == 0 else None

    # ---- the integer-pipe kernel (matcher.cu) on the same step, timed first: its roofline is reported beside the default's
    from hpfw_b200._lib import check
    popc_leg = None
    if args.match_impl != 0 and not args.no_popc_leg:
        check(ctx._lib.hpfw_set_match_impl(ctx.handle, 0))
        keys = st.search_device(d_q, qoffs, TOPK)
        barrier()
        ctx.timing_read(_lib.K_MATCH, reset=True)
        ctx.timing_enable(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(2):
            keys = st.search_device(d_q, qoffs, TOPK)
        e1.record()
        barrier()
        p_ms = max_over_ranks(e0.elapsed_time(e1)) / 2
        pk_ms, pk_n = ctx.timing_read(_lib.K_MATCH, reset=True)
        ctx.timing_enable(False)
        popc_leg = {"ms_per_step": p_ms, "kernel_ms": pk_ms, "launches": pk_n, "steps": 2,
                    "keys": keys.clone()}
    check(ctx._lib.hpfw_set_match_impl(ctx.handle, args.match_impl))

    # ---- device-resident arm
    for _ in range(args.warmup):
        keys = st.search_device(d_q, qoffs, TOPK)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ctx.timing_read(_lib.K_MATCH, reset=True)
    ctx.timing_read(_lib.K_MATCH_TC, reset=True)
    ctx.timing_read(_lib.K_TOPK, reset=True)
    ctx.timing_enable(True)
    launches0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        keys = st.search_device(d_q, qoffs, TOPK)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    launches = ctx.launch_count() - launches0
    match_ms, match_n = ctx.timing_read(_lib.K_MATCH, reset=True)
    tc_ms, tc_n = ctx.timing_read(_lib.K_MATCH_TC, reset=True)
    topk_ms, topk_n = ctx.timing_read(_lib.K_TOPK, reset=True)
    ctx.timing_enable(False)
    clocks = sampler.stop() if rank == 0 else None
    value = nq / (ms * 1e-3)
    # the two kernels must agree bit for bit on the step that was timed
    impls_equal = None
    if popc_leg is not None:
        impls_equal = bool(torch.equal(popc_leg.pop("keys"), keys))

    # correctness of what was just timed: planted queries must come back at their true (track, offset)
    got = hpfw_b200.api.decode_keys(keys.cpu().numpy().view(np.uint64))
    top1_ok = float(np.mean((got["track"][:, 0] == truth[:, 0]) & (got["offset"][:, 0] == truth[:, 1])))

    # ---- N > 1: the merged result of the sharded match must be byte-identical to ONE GPU matching the whole database
    # (untimed): rank 0 rebuilds every rank's shard from its seed, concatenates them into the un-sharded DB and matches the
    # same queries; the [Q][10] key arrays are compared on the device
    sharded_identical = None
    if world > 1 and not args.no_identity_check:
        if rank == 0:
            full = torch.empty(tracks * TRACK_WORDS, dtype=torch.int64, device=dev)
            for r in range(world):
                rlo, rhi = bounds[r], bounds[r + 1]
                w_r, _ = synth.device_hashprint_db(torch, dev, 1234 + r, rhi - rlo, TRACK_WORDS)
                full[rlo * TRACK_WORDS: rhi * TRACK_WORDS] = w_r
                del w_r
            for i, hp_i in enumerate(audio_hps):
                full[i * TRACK_WORDS: i * TRACK_WORDS + len(hp_i)] = torch.from_numpy(hp_i.view(np.int64)).to(dev)
            one = hpfw_b200.MemoryStorage(ctx).build_device(full.data_ptr(), np.arange(tracks + 1, dtype=np.int64) * TRACK_WORDS,
                                                            stream=stream)
            del full
            keys_one = torch.empty((nq, TOPK), dtype=torch.int64, device=dev)
            one.match_device(d_q.data_ptr(), qoffs, TOPK, keys_one.data_ptr(), stream)
            torch.cuda.synchronize()
            sharded_identical = bool(torch.equal(keys_one, keys))
            del one, keys_one
            torch.cuda.empty_cache()
        barrier()

    # ---- strong-scaling view and latency (the weak-scaling step above keeps per-GPU work constant and hides the fixed costs):
    # a FIXED batch of 128 queries over the N shards, and single find() calls (one query: integer-pipe kernel + all-gather +
    # merge + top-k), p50 / p99 over 30 calls, device time from CUDA events, max over ranks
    strong = None
    if not args.no_strong_leg:
        sq = min(args.strong_queries, nq) if args.strong_queries <= nq else args.strong_queries
        if sq > nq:      # a fixed batch larger than the weak step: the step's queries repeated (same work per query)
            reps_q = (sq + nq - 1) // nq
            d_qs = d_q.repeat(reps_q)[:sq * QUERY_WORDS].contiguous()
        else:
            d_qs = d_q
        q128 = sq * QUERY_WORDS
        qo128 = np.arange(sq + 1, dtype=np.int64) * QUERY_WORDS
        s_reps = 5 if sq <= 1024 else 2
        for _ in range(2 if sq <= 1024 else 1):
            st.search_device(d_qs[:q128], qo128, TOPK)
        barrier()
        ctx.timing_read(_lib.K_MATCH_TC, reset=True)
        ctx.timing_read(_lib.K_TOPK, reset=True)
        ctx.timing_enable(True)
        e0.record()
        for _ in range(s_reps):
            st.search_device(d_qs[:q128], qo128, TOPK)
        e1.record()
        barrier()
        s_ms = max_over_ranks(e0.elapsed_time(e1)) / s_reps
        s_tc_ms, _ = ctx.timing_read(_lib.K_MATCH_TC, reset=True)
        s_topk_ms, _ = ctx.timing_read(_lib.K_TOPK, reset=True)
        s_tc_ms, s_topk_ms = s_tc_ms * 5 / s_reps, s_topk_ms * 5 / s_reps     # reported per step below (/ 5)
        del d_qs
        ctx.timing_enable(False)
        qo1 = np.array([0, QUERY_WORDS], dtype=np.int64)
        lat = []
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(32)]
        for i, (a, b) in enumerate(evs):
            if world > 1:
                dist.barrier()
            a.record()
            st.search_device(d_q[i * QUERY_WORDS:(i + 1) * QUERY_WORDS], qo1, TOPK)
            b.record()
        torch.cuda.synchronize()
        lat = sorted(max_over_ranks(a.elapsed_time(b)) for a, b in evs[2:])
        strong = {"queries_per_step": sq, "ms_per_step": s_ms, "queries_per_s": sq / (s_ms * 1e-3),
                  "match_kernel_ms_this_rank": s_tc_ms / 5, "topk_merge_ms_this_rank": s_topk_ms / 5,
                  "fixed_cost_us_per_step": (s_ms - s_tc_ms / 5) * 1e3,
                  "fixed_cost_note": "step time minus match_tc_kernel on rank 0: query expansion, top-k, all-gather of "
                                     "[128][10] keys, merge, launch gaps",
                  "single_find_ms": {"p50": lat[len(lat) // 2], "p99": lat[-1], "calls": len(lat),
                                     "note": "one 385-word query per call through the sharded path (integer-pipe kernel)"}}

    # ---- end-to-end arm (a13 search(): calc_hashprint + find per query): AUDIO in pinned host memory -> H2D -> CQT ->
    # projection/pack -> (all-gather of the hashprints) -> match -> (all-gather of keys, merge) -> D2H records, every step
    q_samples = int(6.0 * 44100)
    h_audio = torch.empty((qpg, q_samples), dtype=torch.float32, pin_memory=True)
    src_track = np.zeros(qpg, dtype=np.int64)
    for q in range(qpg):
        src_track[q] = (q + rank) % AUDIO_TRACKS
        qa, _ = synth.synth_query(audio_tracks[src_track[q]], 7000 + rank * qpg + q, 6.0, 44100, max_semitones=0.25)
        h_audio[q] = torch.from_numpy(qa)
    q_offs = np.arange(qpg + 1, dtype=np.int64) * q_samples
    q_words = ex.words(q_samples)
    assert q_words == QUERY_WORDS
    d_hp_local = torch.empty(qpg * q_words, dtype=torch.int64, device=dev)
    if world > 1:
        src_all = torch.empty((world, qpg), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(src_all.view(-1), torch.from_numpy(src_track).to(dev))
        src_all = src_all.view(-1).cpu().numpy()
    else:
        src_all = src_track

    d_hp_all_buf = torch.empty(world * qpg * q_words, dtype=torch.int64, device=dev) if world > 1 else None

    # Every step uploads its own query audio from pinned host memory and reads its own result back; the upload of step i + 1
    # runs on a copy stream while step i is being matched (two device buffers), as a serving loop would do with independent
    # batches. All K uploads, extractions, matches and result reads are inside the timed region.
    d_audio_buf = [torch.empty_like(h_audio, device=dev) for _ in range(2)]
    up_ready = [torch.cuda.Event() for _ in range(2)]
    up_free = [torch.cuda.Event() for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream(dev)

    def e2e_upload(k):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(up_free[k])              # the previous extraction that read this buffer has finished
            d_audio_buf[k].copy_(h_audio, non_blocking=True)
            up_ready[k].record(copy_stream)

    def e2e_step(k, prefetch_next):
        if prefetch_next:
            e2e_upload(k ^ 1)
        main_stream.wait_event(up_ready[k])
        ex.calc_hashprint_batch_device(d_audio_buf[k].data_ptr(), q_offs, d_hp_local.data_ptr(), main_stream.cuda_stream)
        up_free[k].record(main_stream)
        if world > 1:
            # every rank extracted its own queries: one all-gather (NCCL, inside the library) hands all ranks all hashprints
            st.allgatherv(d_hp_local.data_ptr(), d_hp_all_buf.data_ptr(), [8 * qpg * q_words] * world, main_stream.cuda_stream)
            d_hp_all = d_hp_all_buf
        else:
            d_hp_all = d_hp_local
        kk = st.search_device(d_hp_all, qoffs, TOPK)
        h_out.copy_(kk, non_blocking=True)
        main_stream.synchronize()
        return hpfw_b200.api.decode_keys(h_out.numpy().view(np.uint64))

    for k in range(2):
        up_free[k].record(main_stream)
    e2e_upload(0)
    e2e_step(0, False)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    e2e_upload(0)
    for i in range(args.steps):
        res = e2e_step(i & 1, i + 1 < args.steps)
    e1.record()
    barrier()
    wall = (time.perf_counter() - t0) / args.steps
    # host work (argument packing, key decoding) sits between the launches: take the larger of device and wall time
    ms_e2e = max(max_over_ranks(e0.elapsed_time(e1)) / args.steps, max_over_ranks(wall * 1e3))
    e2e_ok = float(np.mean(res["track"][:, 0] == src_all))
    e2e_value = nq / (ms_e2e * 1e-3)
    # the same through hashprint-level host buffers (MemoryStorage::find on precomputed query hashprints)
    st.search_host(h_q, qoffs, TOPK, h_out)
    barrier()
    t0 = time.perf_counter()
    res_h = st.search_host(h_q, qoffs, TOPK, h_out)
    barrier()
    ms_hp = max_over_ranks((time.perf_counter() - t0) * 1e3)
    hp_ok = float(np.mean((res_h["track"][:, 0] == truth[:, 0]) & (res_h["offset"][:, 0] == truth[:, 1])))

    extraction = None
    if not args.no_extraction:
        extraction = run_extraction(args, ctx, rank, world, dev, max_over_ranks, barrier, st)

    cpp = None
    if not args.no_cpp:
        barrier()
        if rank == 0:
            try:
                cpp = run_cpp_leg(args, ctx, ex, dev, world, tracks, audio_tracks, audio_hps, h_audio, src_track,
                                  np.ascontiguousarray(golden["filters"]), e2e_value)
            except Exception as e:      # the C++ leg must never take the bench line down with it
                cpp = {"error": f"{type(e).__name__}: {e}"}
        if world > 1:
            dist.barrier(group=cpu_group)   # the other ranks wait on the host, their GPUs stay free for the C++ process
        barrier()

    if world > 1:
        lt = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(lt)
        launches = int(lt.item())

    if rank == 0:
        sm_max = None
        peaks_src = "B200_PROFILING.md nominal clocks.max.sm 1965 MHz"
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                sm_max = float(json.load(f)["sm_max_mhz"])
                peaks_src = "MEASURED_PEAKS.json sm_max_mhz"
        except (OSError, KeyError, ValueError):
            sm_max = 1965.0
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        # Integer-pipe roof of the carry-save formulation: per 64-bit word-op the kernel issues 4 LOP3 on the
        # ALU pipe (64 lanes/clk/SM) and 1 POPC on the XU pipe (16 lanes/clk/SM): both allow 16 word-ops/clk/SM.
        # The plain XOR+POPC formulation (2 POPC per word-op) is capped at 8 word-ops/clk/SM.
        peak = sms * WORDOPS_PER_CLK_SM * sm_max * 1e6 / 1e9                  # Gword-op/s
        plain_peak = sms * (POPC32_PER_CLK_SM / 2.0) * sm_max * 1e6 / 1e9
        if args.match_impl == 0:
            p_match_ms, p_match_n, p_steps, p_step_ms = match_ms, match_n, args.steps, ms
        elif popc_leg is not None:
            p_match_ms, p_match_n, p_steps, p_step_ms = (popc_leg["kernel_ms"], popc_leg["launches"], popc_leg["steps"],
                                                         popc_leg["ms_per_step"])
        else:
            p_match_ms, p_match_n, p_steps, p_step_ms = 0.0, 0, 1, 1.0
        ops_per_launch = word_ops_per_query(hi - lo) * nq / max(1, p_match_n / p_steps)
        avg_ms = p_match_ms / max(1, p_match_n)
        achieved = ops_per_launch / (avg_ms * 1e-3) / 1e9 if p_match_n else 0.0
        roof = {"bound": "int-pipe (alu lop3 + xu popc)", "achieved": achieved, "peak": peak, "unit": "Gwordop/s",
                "frac": achieved / peak,
                "traffic": 8.0 * (hi - lo) * TRACK_WORDS * ((nq // 2 + 15) // 16),
                "traffic_note": "modelled: the DB shard is read once per group of 32 queries (grid.y); the ncu --set full "
                                "capture at 2000 tracks x 128 queries measured 0.949 GB against 0.922 GB modelled "
                                "(profiles/r01b_*). HBM is <0.1 % utilised: the kernel is integer-pipe bound",
                "vs_plain_popc_roof": achieved / plain_peak, "plain_popc_roof": plain_peak,
                "kernel": "match_kernel", "avg_launch_ms": avg_ms, "launches": p_match_n,
                "kernel_share_of_step": p_match_ms / p_steps / p_step_ms,
                "queries_per_s": nq / (p_step_ms * 1e-3),
                "peak_how": f"{sms} SMs x {WORDOPS_PER_CLK_SM} word-ops/clk/SM (4 LOP3 @64 lanes/clk + 1 POPC @16 "
                            f"lanes/clk per word-op, carry-save) x {sm_max:.0f} MHz ({peaks_src}); 1 word-op = XOR64 + "
                            f"popcount64; pipe rates measured by the in-run microbenchmark below",
                "microbench": micro,
                "hbm_view": {"algorithmic_bytes_per_launch": 8.0 * (hi - lo) * TRACK_WORDS + 8.0 * nq * QUERY_WORDS,
                             "note": "compute-bound: AI = k word-ops per 8 B of reference, HBM need is <1 % of peak"}}
        roof_popc = None
        dtype = "u64"
        if args.match_impl != 0:
            # default path: the cross-correlation as an exact int8 GEMM on the tensor cores (match_tc.cu); the integer-pipe
            # kernel's roofline (the same step, timed above) goes beside it
            roof_popc = roof if popc_leg is not None else None
            f4 = args.match_impl == 3 or (args.match_impl == 2 and os.environ.get("HPFW_MATCH_TC_F4", "1") != "0")
            dtype = ("e2m1 x e2m1 -> f32 (exact: +1.0/-1.0 per hashprint bit, unit block scales, sums < 2^24)" if f4 else
                     "s8 x s8 -> s32 (exact; +1/-1 byte per hashprint bit)")
            try:
                with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                    bf16_burst = float(json.load(f)["bf16_tflops"])
            except (OSError, KeyError, ValueError):
                bf16_burst = 1590.0
            ops_clk = F4_OPS_PER_CLK_SM if f4 else I8_OPS_PER_CLK_SM
            tc_peak = sms * ops_clk * sm_max * 1e6 / 1e12                      # TOP/s
            tc_ops_per_launch = OPS_PER_WORDOP * word_ops_per_query(hi - lo) * nq / max(1, tc_n / args.steps)
            tc_avg_ms = tc_ms / max(1, tc_n)
            tc_achieved = tc_ops_per_launch / (tc_avg_ms * 1e-3) / 1e12
            kq = (QUERY_WORDS + 7) // 8 * 8 if f4 else (QUERY_WORDS + 3) // 4 * 4
            lib_x = 4 if f4 else 2
            tc_alg_bytes = 8.0 * (hi - lo) * TRACK_WORDS + 64.0 * nq * kq
            roof = {"bound": "tensor", "achieved": tc_achieved, "peak": tc_peak, "unit": "TFLOP/s" if f4 else "TOP/s",
                    "frac": tc_achieved / tc_peak,
                    "traffic": tc_alg_bytes * NCU_TC_TRAFFIC_RATIO,
                    "traffic_note": "dram__bytes_read+write of one launch in the ncu --set full capture at 2000 tracks x 128 "
                                    "queries (profiles/r03d_match_tc_fp4_ncu_full.md: 243.1 MB against 233.8 MB algorithmic), "
                                    "scaled to this launch's algorithmic bytes; 0.14 % of HBM bandwidth",
                    "kernel": ("match_tc_kernel<1> (tcgen05.mma kind::mxf4.block_scale, M=128 queries x N=2x240 offsets, "
                               "f32 in TMEM)" if f4 else
                               "match_tc_kernel<0> (tcgen05.mma kind::i8, M=128 queries x N=2x256 offsets, s32 in TMEM)"),
                    "avg_launch_ms": tc_avg_ms, "launches": tc_n, "kernel_share_of_step": tc_ms / args.steps / ms,
                    "algorithmic_ops_per_launch": tc_ops_per_launch,
                    "peak_how": f"{sms} SMs x {ops_clk} {'fp4' if f4 else 'int8'} ops/clk/SM ({ops_clk // 2} MAC/clk: one "
                                f"M=128,N=256,K={64 if f4 else 32} tcgen05.mma per 128 clk) x {sm_max:.0f} MHz ({peaks_src}); "
                                f"1 word-op (XOR64+popcount64) = 64 multiply-adds = {OPS_PER_WORDOP} ops. MEASURED_PEAKS.json "
                                f"holds no {'fp4' if f4 else 'int8'} figure: {lib_x}x its bf16 cuBLAS burst rate would be "
                                f"{lib_x * bf16_burst:.0f} TOP/s, which this kernel exceeds, so the pipe ceiling at the maximum "
                                f"SM clock is the denominator (under sw_power_cap the clock is lower: see clocks)",
                    "peak_bf16_cublas_scaled": lib_x * bf16_burst, "frac_of_bf16_cublas_scaled": tc_achieved / (lib_x * bf16_burst),
                    "wordops_per_s_G": tc_achieved * 1e3 / OPS_PER_WORDOP,
                    "vs_popc_pipe_roof": tc_achieved * 1e3 / OPS_PER_WORDOP / plain_peak,
                    "hbm_view": {"algorithmic_bytes_per_launch": 8.0 * (hi - lo) * TRACK_WORDS + 64.0 * nq * kq,
                                 "note": "tensor-bound: a tile reads (480 or 512) + k reference words (expanded in shared "
                                         "memory) for that many x k word-ops per query; the expanded queries (1.6 or 3 MB "
                                         "per group of 128) stream from L2"}}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": dtype,
            "data": "synthetic", "config": workload_config(n, tracks, qpg),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h_audio.numel() * 4 * n),
                    "d2h_bytes_per_step": int(h_out.numel() * 8), "ms_per_step": ms_e2e, "top1_ok": e2e_ok,
                    "what": "6 s query AUDIO (pinned host) -> H2D -> CQT -> projection/pack -> match -> top-k records on "
                            "the host, every step; the upload of step i+1 overlaps the match of step i (two device buffers); "
                            "queries are noisy pitch-shifted slices of audio-derived DB tracks",
                    "hashprint_in": {"value": nq / (ms_hp * 1e-3), "unit": UNIT, "ms_per_step": ms_hp,
                                     "h2d_bytes_per_step": int(h_q.numel() * 8), "top1_ok": hp_ok}},
            "gpu_launches": launches,
            "roofline": roof,
            "match_impl": {0: "integer pipes (matcher.cu)", 1: "tensor cores, int8 operands (match_tc.cu)",
                           2: "tensor cores (fp4 operands unless HPFW_MATCH_TC_F4=0) for groups of 128 queries, integer "
                              "pipes for a small remainder (default)",
                           3: "tensor cores, fp4 operands (match_tc.cu)"}[args.match_impl],
            "top1_ok": top1_ok,
            "topk_ms_per_step": topk_ms / args.steps,
        }
        if sharded_identical is not None:
            line["sharded_bit_identical"] = sharded_identical
        if balance is not None:
            line["shard_balance"] = balance
        if strong is not None:
            line["strong"] = strong
        if cpp is not None:
            line["e2e_cpp"] = cpp
        if roof_popc is not None:
            line["roofline_popc"] = roof_popc
            line["impls_bit_identical"] = impls_equal
        if extraction is not None:
            line["extraction"] = extraction
        if n == 1 and not args.no_cpu_baseline:
            cores = host_cores()
            nqc = max(cores, 8)
            sample_tracks = 2000
            kind, dt, qps = cpu_sample(cores, sample_tracks, nqc, tracks=tracks)
            line["cpu_baseline"] = {
                "value": qps, "unit": UNIT, "cores": cores, "kind": kind,
                "wordops_per_s_per_core": cpu_wordops_per_core(qps, cores, tracks),
                "sample": f"{nqc} queries x {sample_tracks}-track subset ({dt:.1f} s on {cores} host threads), scaled "
                          f"x{sample_tracks}/{tracks} to the {tracks}-track DB; reference MemoryStorage::find, "
                          f"-Ofast -march=native"}
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--tracks", type=int, default=TRACKS)
    ap.add_argument("--queries-per-gpu", type=int, default=QUERIES_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--match-impl", type=int, default=2, choices=[0, 1, 2, 3],
                    help="0 = integer-pipe matcher, 1 = tensor-core matcher (int8), 3 = tensor-core matcher (fp4), "
                         "2 = default routing (hpfw_set_match_impl)")
    ap.add_argument("--no-popc-leg", action="store_true", help="skip timing the integer-pipe kernel beside the default")
    ap.add_argument("--no-extraction", action="store_true", help="skip the secondary hashprint-extraction leg")
    ap.add_argument("--balance", action="store_true",
                    help="N > 1: re-plan the track ranges by each GPU's measured match speed (hpfw_shard_plan_weighted). Off by "
                         "default: measured on 2 GPUs the sustained, power-capped speeds differ less than the calibration's "
                         "short runs suggest and the step time does not change (2,243 vs 2,240 queries/s)")
    ap.add_argument("--no-identity-check", action="store_true", help="N > 1: skip the untimed sharded-vs-one-GPU key comparison")
    ap.add_argument("--no-strong-leg", action="store_true", help="skip the fixed-batch / single-find latency leg")
    ap.add_argument("--strong-queries", type=int, default=128,
                    help="size of the FIXED query batch of the strong-scaling leg (128; BASELINE configs[3] uses 10000)")
    ap.add_argument("--no-cpp", action="store_true", help="skip the C++ API leg (examples/cpp/bench-liveid.cpp, e2e_cpp)")
    ap.add_argument("--only-cpp-index", action="store_true", help="run only the C++ index() leg with HPFW_TRACE=1 (tuning aid)")
    ap.add_argument("--cpp-index-tracks", type=int, default=1024, help="WAV files index() reads in the C++ leg (3-min PCM16)")
    ap.add_argument("--extract-tracks", type=int, default=1000, help="3-min tracks of the extraction leg (all GPUs together)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.only_cpp_index:
        import tempfile
        import shutil
        import torch
        work = tempfile.mkdtemp(prefix="hpfw_cpp_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        try:
            r = cpp_index_leg(build_cpp_bench(), work, dict(os.environ, HPFW_TRACE="1"), torch.device("cuda", 0), 1,
                              args.cpp_index_tracks, reps=max(1, args.steps))
        finally:
            shutil.rmtree(work, ignore_errors=True)
        print(json.dumps(r, indent=1))
        return 0
    return run_cuda(args)


if __name__ == "__main__":
    sys.exit(main())
