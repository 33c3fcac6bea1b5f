"""oracle/ — TEST INFRASTRUCTURE ONLY.

CPU checkers for the hashprint feature-to-match path. Importable only from tests/, __graft_entry__.smoke()/build() and
bench.py's cpu_baseline / --impl reference legs. The product (hpfw_b200/, include/) never imports, links or executes
anything in this directory.

Two libraries, both built by `oracle.build()` (called from __graft_entry__.build()):
  * oracle/libhpfw_oracle.so   — hpfw_oracle.c, our plain-C restatement (kind "port");
  * oracle/_ref/libhpfw_ref.so — the reference's OWN headers compiled from /root/reference through
    oracle/ref_build/ (kind "reference"). Only buildable where /root/reference exists (this container); the built
    .so travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libhpfw_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libhpfw_ref.so")                    # -march=native (reference's own flags)
REF_SO_PORTABLE = os.path.join(HERE, "_ref", "libhpfw_ref_x86_64_v3.so")  # fallback if the GPU box's CPU differs
REFERENCE_ROOT = "/root/reference"

BINS, CTX, LAG, NFILT, FRAME = 121, 20, 80, 64, 2420

_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def _cpu_flags() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return " ".join(sorted(line.split(":", 1)[1].split()))
    except OSError:
        pass
    return ""


def _flags_ok(sidecar: str) -> bool:
    """A -march=native library only runs where every CPU feature of its build host exists."""
    if not os.path.exists(sidecar):
        return False
    built = set(open(sidecar).read().split())
    return built.issubset(set(_cpu_flags().split()))


def _stale(target: str, sources) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(s) and os.path.getmtime(s) > t for s in sources)


def build(verbose: bool = False) -> None:
    """Compile the C restatement; compile the reference headers too when /root/reference is present."""
    src = os.path.join(HERE, "hpfw_oracle.c")
    if _stale(ORACLE_SO, [src]) or not _flags_ok(ORACLE_SO + ".cpuflags"):
        cmd = ["gcc", "-O3", "-march=native", "-std=c11", "-fPIC", "-shared", src, "-o", ORACLE_SO, "-lm", "-lpthread"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
        open(ORACLE_SO + ".cpuflags", "w").write(_cpu_flags())
    rb = os.path.join(HERE, "ref_build")
    if os.path.isdir(os.path.join(REFERENCE_ROOT, "include", "hpfw")):
        srcs = [os.path.join(rb, "ref_driver.cpp"), os.path.join(rb, "eigen_core_shim", "Core"), os.path.join(rb, "Makefile")]
        if _stale(REF_SO, srcs) or _stale(REF_SO_PORTABLE, srcs):
            subprocess.check_call(["make", "-C", rb, "all"] + ([] if verbose else ["-s"]))
            open(REF_SO + ".cpuflags", "w").write(_cpu_flags())


_oracle = None
_ref = None


def lib():
    """ctypes handle of the C restatement (builds it on first use)."""
    global _oracle
    if _oracle is None:
        if _stale(ORACLE_SO, [os.path.join(HERE, "hpfw_oracle.c")]) or not _flags_ok(ORACLE_SO + ".cpuflags"):
            build()
        L = C.CDLL(ORACLE_SO)
        L.orc_amplitude_to_db.argtypes = [_f32p, C.c_int]
        L.orc_calc_frames.argtypes = [_f32p, C.c_int, _f32p]
        L.orc_calc_frames.restype = C.c_int
        L.orc_project_f32.argtypes = [_f32p, C.c_int, _f32p, _f32p]
        L.orc_project_f32.restype = C.c_int
        L.orc_project_f64.argtypes = [_f32p, C.c_int, _f32p, _f64p]
        L.orc_project_f64.restype = C.c_int
        L.orc_fingerprint_pack_f32.argtypes = [_f32p, C.c_int, _u64p]
        L.orc_fingerprint_pack_f32.restype = C.c_int
        L.orc_fingerprint_pack_f64.argtypes = [_f64p, C.c_int, _u64p]
        L.orc_fingerprint_pack_f64.restype = C.c_int
        L.orc_delta_f64.argtypes = [_f64p, C.c_int, _f64p]
        L.orc_delta_f64.restype = C.c_int
        L.orc_hashprint_from_spectrogram.argtypes = [_f32p, C.c_int, _f32p, _u64p]
        L.orc_hashprint_from_spectrogram.restype = C.c_int
        L.orc_find.argtypes = [_u64p, _i64p, C.c_int, _u64p, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_int64)]
        L.orc_find.restype = C.c_int64
        L.orc_per_track_best.argtypes = [_u64p, _i64p, C.c_int, _u64p, C.c_int, _u64p, _i64p]
        L.orc_find_topk.argtypes = [_u64p, _i64p, C.c_int, _u64p, C.c_int, C.c_int, _i64p, _u64p, _i64p]
        L.orc_find_topk_batch.argtypes = [_u64p, _i64p, C.c_int, _u64p, _i64p, C.c_int, C.c_int, _i64p, _u64p, _i64p,
                                          C.c_int]
        _oracle = L
    return _oracle


def ref_available() -> bool:
    return os.path.exists(REF_SO) or os.path.exists(REF_SO_PORTABLE)


def ref_flavour() -> str:
    """Which compiled-reference library this host can run: 'native' (-march=native of the build host) or 'x86-64-v3'."""
    if os.path.exists(REF_SO) and _flags_ok(REF_SO + ".cpuflags"):
        return "native"
    return "x86-64-v3"


def ref():
    """ctypes handle of the compiled reference headers (oracle/_ref); raises if it was never built."""
    global _ref
    if _ref is None:
        path = REF_SO if ref_flavour() == "native" else REF_SO_PORTABLE
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run oracle.build() where /root/reference exists")
        L = C.CDLL(path)
        L.ref_amplitude_to_db.argtypes = [_f32p, C.c_int]
        L.ref_calc_frames.argtypes = [_f32p, C.c_int, _f32p]
        L.ref_calc_frames.restype = C.c_int
        L.ref_project.argtypes = [_f32p, C.c_int, _f32p, _f32p]
        L.ref_project.restype = C.c_int
        L.ref_hashprint_from_spectrogram.argtypes = [_f32p, C.c_int, _f32p, _u64p]
        L.ref_hashprint_from_spectrogram.restype = C.c_int
        L.ref_fingerprint_pack.argtypes = [_f32p, C.c_int, _u64p]
        L.ref_fingerprint_pack.restype = C.c_int
        L.ref_calc_cov.argtypes = [_f32p, C.c_int, _f32p]
        L.ref_calc_cov_generic.argtypes = [_f32p, C.c_int, C.c_int, _f32p]
        L.ref_calc_filters.argtypes = [_f32p, _f32p]
        L.ref_register_spectrogram.argtypes = [C.c_char_p, _f32p, C.c_int]
        L.ref_collector_new.restype = C.c_void_p
        L.ref_collector_del.argtypes = [C.c_void_p]
        L.ref_collector_prepare.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.c_int]
        L.ref_collector_prepare.restype = C.c_void_p
        L.ref_prepared_count.argtypes = [C.c_void_p]
        L.ref_prepared_name.argtypes = [C.c_void_p, C.c_int]
        L.ref_prepared_name.restype = C.c_char_p
        L.ref_prepared_size.argtypes = [C.c_void_p, C.c_int]
        L.ref_prepared_words.argtypes = [C.c_void_p, C.c_int]
        L.ref_prepared_words.restype = C.POINTER(C.c_uint64)
        L.ref_prepared_free.argtypes = [C.c_void_p]
        L.ref_collector_calc_hashprint.argtypes = [C.c_void_p, C.c_char_p, _u64p, C.c_int]
        L.ref_storage_build.argtypes = [_u64p, _i64p, C.c_int]
        L.ref_storage_build.restype = C.c_void_p
        L.ref_storage_del.argtypes = [C.c_void_p]
        L.ref_storage_find.argtypes = [C.c_void_p, _u64p, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_int64)]
        L.ref_storage_find.restype = C.c_int64
        L.ref_storage_find_batch.argtypes = [C.c_void_p, _u64p, _i64p, C.c_int, _i64p, _u64p, _i64p, C.c_int]
        _ref = L
    return _ref


# ------------------------------------------------------------------------------------------------ numpy conveniences
def pack_db(tracks):
    """list of uint64 arrays -> (concatenated words, int64 offsets[R+1])."""
    offs = np.zeros(len(tracks) + 1, dtype=np.int64)
    for i, t in enumerate(tracks):
        offs[i + 1] = offs[i] + len(t)
    words = np.concatenate([np.asarray(t, dtype=np.uint64) for t in tracks]) if len(tracks) and offs[-1] > 0 \
        else np.zeros(0, dtype=np.uint64)
    return np.ascontiguousarray(words), offs


def _nonempty(a, dtype):
    a = np.ascontiguousarray(a, dtype=dtype)
    return a if a.size else np.zeros(1, dtype=dtype)


def find(words, offsets, q):
    """MemoryStorage::find restated -> (track, cnt, offset)."""
    cnt, off = C.c_uint64(), C.c_int64()
    tr = lib().orc_find(_nonempty(words, np.uint64), np.ascontiguousarray(offsets, dtype=np.int64), len(offsets) - 1,
                        _nonempty(q, np.uint64), len(q), C.byref(cnt), C.byref(off))
    return int(tr), int(cnt.value), int(off.value)


def find_topk(words, offsets, q, topk):
    tr = np.empty(topk, np.int64); d = np.empty(topk, np.uint64); o = np.empty(topk, np.int64)
    lib().orc_find_topk(_nonempty(words, np.uint64), np.ascontiguousarray(offsets, dtype=np.int64), len(offsets) - 1,
                        _nonempty(q, np.uint64), len(q), topk, tr, d, o)
    return tr, d, o


def find_topk_batch(words, offsets, qwords, qoffsets, topk, n_threads):
    nq = len(qoffsets) - 1
    tr = np.empty((nq, topk), np.int64); d = np.empty((nq, topk), np.uint64); o = np.empty((nq, topk), np.int64)
    lib().orc_find_topk_batch(_nonempty(words, np.uint64), np.ascontiguousarray(offsets, dtype=np.int64),
                              len(offsets) - 1, _nonempty(qwords, np.uint64),
                              np.ascontiguousarray(qoffsets, dtype=np.int64), nq, topk, tr, d, o, n_threads)
    return tr, d, o


def per_track_best(words, offsets, q):
    r = len(offsets) - 1
    d = np.empty(max(r, 1), np.uint64); o = np.empty(max(r, 1), np.int64)
    lib().orc_per_track_best(_nonempty(words, np.uint64), np.ascontiguousarray(offsets, dtype=np.int64), r,
                             _nonempty(q, np.uint64), len(q), d, o)
    return d[:r], o[:r]


def hashprint_from_spectrogram(spectro_tm, filters_cm):
    """spectro_tm: float32 [cols,121] (time-major = reference's col-major 121 x cols); filters_cm: float32 memory of a
    column-major 64x2420 matrix, i.e. numpy [2420,64]. Returns uint64[cols-99]."""
    s = np.ascontiguousarray(spectro_tm, dtype=np.float32)
    cols = s.shape[0]
    hp = np.zeros(max(cols - 99, 1), dtype=np.uint64)
    n = lib().orc_hashprint_from_spectrogram(s.reshape(-1), cols, np.ascontiguousarray(filters_cm, np.float32).reshape(-1), hp)
    return hp[:n]


def project_f64(spectro_tm, filters_cm):
    s = np.ascontiguousarray(spectro_tm, dtype=np.float32)
    cols = s.shape[0]
    nf = cols - CTX + 1
    y = np.zeros((max(nf, 1), NFILT), dtype=np.float64)
    lib().orc_project_f64(s.reshape(-1), cols, np.ascontiguousarray(filters_cm, np.float32).reshape(-1), y.reshape(-1))
    return y[:max(nf, 0)]


def hashprint_f64(spectro_tm, filters_cm):
    """Rounding-free yardstick: projection and delta in double. Returns (hashprint, |delta| margins [n,64])."""
    y = project_f64(spectro_tm, filters_cm)
    n = y.shape[0] - LAG
    hp = np.zeros(max(n, 1), dtype=np.uint64)
    dl = np.zeros((max(n, 1), NFILT), dtype=np.float64)
    if n > 0:
        lib().orc_fingerprint_pack_f64(y.reshape(-1), y.shape[0], hp)
        lib().orc_delta_f64(y.reshape(-1), y.shape[0], dl.reshape(-1))
    return hp[:max(n, 0)], dl[:max(n, 0)]


def ref_find(words, offsets, q):
    L = ref()
    st = L.ref_storage_build(_nonempty(words, np.uint64), np.ascontiguousarray(offsets, dtype=np.int64), len(offsets) - 1)
    try:
        cnt, off = C.c_uint64(), C.c_int64()
        tr = L.ref_storage_find(st, _nonempty(q, np.uint64), len(q), C.byref(cnt), C.byref(off))
        return int(tr), int(cnt.value), int(off.value)
    finally:
        L.ref_storage_del(st)


def ref_find_batch(words, offsets, qwords, qoffsets, n_threads):
    L = ref()
    nq = len(qoffsets) - 1
    st = L.ref_storage_build(_nonempty(words, np.uint64), np.ascontiguousarray(offsets, dtype=np.int64), len(offsets) - 1)
    try:
        tr = np.empty(nq, np.int64); d = np.empty(nq, np.uint64); o = np.empty(nq, np.int64)
        L.ref_storage_find_batch(st, _nonempty(qwords, np.uint64), np.ascontiguousarray(qoffsets, dtype=np.int64), nq,
                                 tr, d, o, n_threads)
        return tr, d, o
    finally:
        L.ref_storage_del(st)


def ref_hashprint_from_spectrogram(spectro_tm, filters_cm):
    s = np.ascontiguousarray(spectro_tm, dtype=np.float32)
    cols = s.shape[0]
    hp = np.zeros(max(cols - 99, 1), dtype=np.uint64)
    n = ref().ref_hashprint_from_spectrogram(s.reshape(-1), cols, np.ascontiguousarray(filters_cm, np.float32).reshape(-1), hp)
    return hp[:n]
