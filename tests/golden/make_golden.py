"""Generate tests/golden/*.npz from the REFERENCE'S OWN CODE (oracle/_ref/libhpfw_ref.so = /root/reference headers compiled by
oracle/ref_build/Makefile). Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Nothing under /root/reference is read at test time; the tests only read the .npz files written here.
Inputs are deterministic (fixed numpy seeds, hpfw_b200/synth.py generators). The CQT spectrograms that feed the stage-2/3
fixtures come from oracle/nsgcq.py (essentia is absent => that stage is "parity unpinned"); everything downstream of the
spectrogram in these files is produced by the reference headers.
"""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from oracle import nsgcq  # noqa: E402
from hpfw_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def gen_matcher():
    """MemoryStorage::find (storage.h:27-64) on small ragged DBs, incl. the edge cases the semantics imply."""
    cases = {}

    def add(name, words, offs, qwords, qoffs):
        nq = len(qoffs) - 1
        res = np.zeros((nq, 3), dtype=np.int64)
        for i in range(nq):
            tr, cnt, off = oracle.ref_find(words, offs, qwords[qoffs[i]:qoffs[i + 1]])
            # cnt = SIZE_MAX for an empty DB: store as -1
            res[i] = (tr, cnt if cnt < (1 << 63) else -1, off)
        cases[name] = dict(words=words, offs=offs, qwords=qwords, qoffs=qoffs, res=res)

    # A: ragged lengths, noisy sub-sequence queries of several lengths
    lens = np.array([300, 1111, 257, 2049, 96, 2048, 513], dtype=np.int64)
    w, o = synth.synth_hashprint_db(101, len(lens), lens)
    qw, qo, _ = synth.synth_hashprint_queries(102, w, o, 12, np.array([63, 64, 65, 96, 33, 8, 1, 143, 7, 90, 17, 50]))
    add("ragged", w, o, qw, qo)

    # B: tracks shorter than the query (query truncated to the track: storage.h:34-38) and an empty track (distance 0 wins)
    lens = np.array([40, 500, 10, 385], dtype=np.int64)
    w, o = synth.synth_hashprint_db(103, len(lens), lens)
    qw, qo, _ = synth.synth_hashprint_queries(104, w, o, 4, np.array([385, 100, 41, 11]))
    add("short_refs", w, o, qw, qo)
    lens = np.array([120, 0, 200], dtype=np.int64)
    w, o = synth.synth_hashprint_db(105, len(lens), lens)
    qw, qo, _ = synth.synth_hashprint_queries(106, w, o, 2, np.array([50, 20]))
    add("empty_track", w, o, qw, qo)

    # C: exact ties: duplicated tracks and a periodic track (lowest offset, earliest track must win)
    base, _ = synth.synth_hashprint_db(107, 1, 200)
    period = np.tile(base[:25], 8)
    w = np.concatenate([base, period, base, period])
    o = np.array([0, 200, 400, 600, 800], dtype=np.int64)
    qw = np.concatenate([base[30:80], period[5:45], base[0:200]])
    qo = np.array([0, 50, 90, 290], dtype=np.int64)
    add("ties", w, o, qw, qo)

    # D: empty DB -> {"" , SIZE_MAX, 0}; empty query -> distance 0 at offset 0 of track 0
    add("empty_db", np.zeros(0, np.uint64), np.zeros(1, np.int64), base[:10].copy(), np.array([0, 10], dtype=np.int64))
    w, o = synth.synth_hashprint_db(108, 3, 64)
    add("empty_query", w, o, np.zeros(0, np.uint64), np.array([0, 0], dtype=np.int64))

    flat = {}
    for name, c in cases.items():
        for k, v in c.items():
            flat[f"{name}__{k}"] = v
    np.savez_compressed(os.path.join(OUT, "matcher.npz"), **flat)
    print("matcher.npz:", list(cases))


def gen_hashprint():
    """Stages a3..a7 + a10 from the reference headers on synthetic 22.05 kHz audio (BASELINE configs[0] shapes)."""
    R = oracle.ref()
    sr = 22050
    tracks = [synth.synth_track(1000 + i, 30.0, sr) for i in range(3)]
    mags = [nsgcq.nsgcq_magnitude(t).astype(np.float32) for t in tracks]          # [cols,121]
    # a3 by the reference
    specs = []
    for m in mags:
        s = np.ascontiguousarray(m.copy()).reshape(-1)
        R.ref_amplitude_to_db(s, m.shape[0])
        specs.append(s.reshape(m.shape))
    cols = specs[0].shape[0]
    # a10: covariance per track, summed, /(n+1) as parallel_collector.h:111 (cache.size()+1), then filters
    cov = np.zeros((2420, 2420), dtype=np.float32)
    tmp = np.zeros(2420 * 2420, dtype=np.float32)
    for s in specs:
        R.ref_calc_cov(s.reshape(-1), s.shape[0], tmp)
        cov += tmp.reshape(2420, 2420)
    cov /= np.float32(len(specs) + 1)
    filt = np.zeros(64 * 2420, dtype=np.float32)
    R.ref_calc_filters(np.ascontiguousarray(cov).reshape(-1), filt)
    filt = filt.reshape(2420, 64)     # memory of a column-major 64 x 2420
    # a4..a7 by the reference on track 0 and on a 6 s query
    q_audio, q_start = synth.synth_query(tracks[0], 2000, 6.0, sr)
    q_mag = nsgcq.nsgcq_magnitude(q_audio).astype(np.float32)
    q_spec = np.ascontiguousarray(q_mag.copy()).reshape(-1)
    R.ref_amplitude_to_db(q_spec, q_mag.shape[0])
    q_spec = q_spec.reshape(q_mag.shape)
    hp0 = oracle.ref_hashprint_from_spectrogram(specs[0], filt)
    hpq = oracle.ref_hashprint_from_spectrogram(q_spec, filt)
    y0 = np.zeros((cols - 19) * 64, dtype=np.float32)
    R.ref_project(specs[0].reshape(-1), cols, filt.reshape(-1), y0)
    hps = [oracle.ref_hashprint_from_spectrogram(s, filt) for s in specs]
    np.savez_compressed(
        os.path.join(OUT, "hashprint.npz"),
        filters=filt, mag0=mags[0], spec0=specs[0], y0_head=y0.reshape(-1, 64)[:256], hp0=hp0, hp1=hps[1], hp2=hps[2],
        q_mag=q_mag, q_spec=q_spec, hpq=hpq, q_start=np.int64(q_start),
        audio_q=q_audio, eig_cov_diag=np.diag(cov).copy())
    print("hashprint.npz: cols", cols, "words", len(hp0), "query words", len(hpq))

    # small KATs (SURVEY Appendix A probe results)
    S = np.zeros((30, 121), dtype=np.float32)
    for b in range(121):
        S[:, b] = 100.0 * b + np.arange(30)
    fr = np.zeros(2420 * 11, dtype=np.float32)
    n = R.ref_calc_frames(S.reshape(-1), 30, fr)
    assert n == 11
    ypat = np.zeros((81, 64), dtype=np.float32)   # 81 columns -> 1 word
    kat = {}
    for f in (0, 1, 31, 62, 63):
        y = ypat.copy()
        y[0, :] = -1.0
        y[0, f] = 1.0       # delta >= 0 only for filter f
        hp = np.zeros(1, dtype=np.uint64)
        R.ref_fingerprint_pack(y.reshape(-1), 81, hp)
        kat[f"only_{f}"] = hp.copy()
    hp = np.zeros(1, dtype=np.uint64)
    R.ref_fingerprint_pack(ypat.reshape(-1), 81, hp)     # all deltas exactly 0 -> all bits 1
    kat["all_zero_delta"] = hp.copy()
    # amplitude_to_db KAT
    rng = np.random.default_rng(7)
    amp = np.abs(rng.standard_normal((50, 121))).astype(np.float32) * np.float32(0.01)
    amp[3, 5] = 0.0
    amp[10, 7] = 2.5
    db = np.ascontiguousarray(amp.copy()).reshape(-1)
    R.ref_amplitude_to_db(db, 50)
    np.savez_compressed(os.path.join(OUT, "kat.npz"), frames_in=S, frames_out=fr.reshape(2420, 11), amp=amp,
                        db=db.reshape(50, 121), **kat)
    print("kat.npz written")


def gen_collector():
    """ParallelCollector::prepare + calc_hashprint + MemoryStorage::find end to end through the reference (a8, a9, a12)."""
    R = oracle.ref()
    g = np.load(os.path.join(OUT, "hashprint.npz"))
    sr = 22050
    tracks = [synth.synth_track(1000 + i, 30.0, sr) for i in range(3)]
    names = []
    for i, t in enumerate(tracks):
        m = nsgcq.nsgcq_magnitude(t).astype(np.float32)
        s = np.ascontiguousarray(m.copy()).reshape(-1)
        R.ref_amplitude_to_db(s, m.shape[0])
        nm = f"track{i}.wav".encode()
        R.ref_register_spectrogram(nm, s, m.shape[0])
        names.append(nm)
    R.ref_register_spectrogram(b"query.wav", np.ascontiguousarray(g["q_spec"]).reshape(-1), g["q_spec"].shape[0])
    col = R.ref_collector_new()
    arr = (C.c_char_p * len(names))(*names)
    prep = R.ref_collector_prepare(col, arr, len(names))
    n = R.ref_prepared_count(prep)
    db = {}
    for i in range(n):
        nm = R.ref_prepared_name(prep, i).decode()
        sz = R.ref_prepared_size(prep, i)
        db[nm] = np.ctypeslib.as_array(R.ref_prepared_words(prep, i), shape=(sz,)).copy()
    R.ref_prepared_free(prep)
    order = sorted(db)
    hpq = np.zeros(4096, dtype=np.uint64)
    nq = R.ref_collector_calc_hashprint(col, b"query.wav", hpq, 4096)
    hpq = hpq[:nq]
    words, offs = oracle.pack_db([db[k] for k in order])
    tr, cnt, off = oracle.ref_find(words, offs, hpq)
    R.ref_collector_del(col)
    np.savez_compressed(os.path.join(OUT, "collector.npz"), names=np.array(order), words=words, offs=offs, hpq=hpq,
                        res=np.array([tr, cnt, off], dtype=np.int64))
    print("collector.npz: prepared", order, "query words", nq, "find ->", (tr, cnt, off), "true start sample",
          int(g["q_start"]))


if __name__ == "__main__":
    oracle.build()
    gen_matcher()
    gen_hashprint()
    gen_collector()
