"""Deterministic synthetic inputs for tests and bench.py (there is no network for datasets).

  synth_track / synth_query  — audio: decaying harmonic notes on a semitone grid + white noise; queries are
                               pitch-shifted noisy slices of a track (BASELINE.json configs[0]).
  synth_hashprint_db / synth_hashprint_queries — matcher-only configs: iid uniform u64 words; a query is a DB slice with
                               every bit flipped with probability `flip`.
Pure numpy on the host; `device_hashprint_db` produces the same kind of data directly in HBM with torch (plumbing).
"""
from __future__ import annotations

import numpy as np


def synth_track(seed: int, seconds: float, sr: int = 44100) -> np.ndarray:
    """float32 mono audio, peak 0.5."""
    rng = np.random.default_rng(seed)
    n = int(round(seconds * sr))
    x = np.zeros(n, dtype=np.float64)
    t_pos = 0.0
    while t_pos < seconds:
        dur = rng.uniform(0.1, 0.6)
        semis = rng.integers(0, 49)                     # C3 .. C7
        f0 = 130.81 * 2.0 ** (semis / 12.0)
        i0 = int(t_pos * sr)
        i1 = min(n, i0 + int(dur * sr))
        if i1 > i0:
            tt = np.arange(i1 - i0) / sr
            env = np.exp(-tt * rng.uniform(3.0, 9.0)) * np.minimum(1.0, tt * 200.0)
            note = np.zeros(i1 - i0)
            for h in range(1, 6):
                if f0 * h < sr / 2:
                    note += (0.6 ** (h - 1)) * np.sin(2 * np.pi * f0 * h * tt + rng.uniform(0, 2 * np.pi))
            x[i0:i1] += rng.uniform(0.4, 1.0) * env * note
        t_pos += dur * rng.uniform(0.3, 0.9)            # overlap
    x += 10 ** (-40 / 20) * rng.standard_normal(n)
    x *= 0.5 / max(1e-9, np.abs(x).max())
    return x.astype(np.float32)


def synth_query(track: np.ndarray, seed: int, seconds: float, sr: int = 44100, snr_db: float = 10.0,
                max_semitones: float = 0.5):
    """(query audio float32, true start sample): a slice resampled by 2^(s/12), s ~ U[-max,max], plus noise at snr_db."""
    rng = np.random.default_rng(seed)
    n = int(round(seconds * sr))
    ratio = 2.0 ** (rng.uniform(-max_semitones, max_semitones) / 12.0)
    need = int(np.ceil(n * ratio)) + 2
    start = int(rng.integers(0, max(1, len(track) - need)))
    src = track[start:start + need].astype(np.float64)
    pos = np.arange(n) * ratio
    i = np.floor(pos).astype(np.int64)
    frac = pos - i
    i = np.clip(i, 0, len(src) - 2)
    q = src[i] * (1 - frac) + src[i + 1] * frac
    p_sig = float(np.mean(q ** 2)) + 1e-12
    q = q + np.sqrt(p_sig / (10 ** (snr_db / 10))) * rng.standard_normal(n)
    q *= 0.5 / max(1e-9, np.abs(q).max())
    return q.astype(np.float32), start


def synth_hashprint_db(seed: int, n_tracks: int, words_per_track):
    """(words uint64[total], offsets int64[n_tracks+1]); words_per_track: int or array of per-track lengths."""
    rng = np.random.default_rng(seed)
    lens = np.full(n_tracks, words_per_track, dtype=np.int64) if np.isscalar(words_per_track) \
        else np.asarray(words_per_track, dtype=np.int64)
    offs = np.zeros(n_tracks + 1, dtype=np.int64)
    np.cumsum(lens, out=offs[1:])
    words = rng.integers(0, 1 << 64, size=int(offs[-1]), dtype=np.uint64)
    return words, offs


def synth_hashprint_queries(seed: int, words: np.ndarray, offs: np.ndarray, n_queries: int, k, flip: float = 0.25):
    """Noisy DB slices. Returns (qwords, qoffs, truth[n_queries,2] = (track, offset)). k: int or array of lengths.
    Tracks shorter than the query length are never chosen as the source."""
    rng = np.random.default_rng(seed)
    ks = np.full(n_queries, k, dtype=np.int64) if np.isscalar(k) else np.asarray(k, dtype=np.int64)
    lens = np.diff(offs)
    qoffs = np.zeros(n_queries + 1, dtype=np.int64)
    np.cumsum(ks, out=qoffs[1:])
    qwords = np.zeros(int(qoffs[-1]), dtype=np.uint64)
    truth = np.zeros((n_queries, 2), dtype=np.int64)
    for q in range(n_queries):
        kq = int(ks[q])
        ok = np.nonzero(lens >= kq)[0]
        tr = int(ok[rng.integers(0, len(ok))])
        off = int(rng.integers(0, lens[tr] - kq + 1))
        sl = words[offs[tr] + off: offs[tr] + off + kq].copy()
        noise = np.zeros(kq, dtype=np.uint64)
        for b in range(64):
            noise |= (rng.random(kq) < flip).astype(np.uint64) << np.uint64(b)
        qwords[qoffs[q]:qoffs[q + 1]] = sl ^ noise
        truth[q] = (tr, off)
    return qwords, qoffs, truth


def device_hashprint_db(torch, device, seed: int, n_tracks: int, words_per_track: int):
    """Same distribution as synth_hashprint_db, generated in HBM: (int64 tensor viewed as u64 words, offsets ndarray)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    total = n_tracks * words_per_track
    words = torch.randint(-(1 << 63), (1 << 63) - 1, (total,), dtype=torch.int64, device=device, generator=g)
    offs = np.arange(n_tracks + 1, dtype=np.int64) * words_per_track
    return words, offs


def device_hashprint_queries(torch, words, offs: np.ndarray, seed: int, n_queries: int, k: int, flip: float = 0.25):
    """Noisy slices of a device-resident DB, built on the device. Returns (qwords tensor, qoffs ndarray, truth ndarray)."""
    rng = np.random.default_rng(seed)
    lens = np.diff(offs)
    ok = np.nonzero(lens >= k)[0]
    tr = ok[rng.integers(0, len(ok), size=n_queries)]
    off = (rng.random(n_queries) * (lens[tr] - k + 1)).astype(np.int64)
    start = torch.as_tensor(offs[tr] + off, device=words.device)
    idx = start[:, None] + torch.arange(k, device=words.device)[None, :]
    sl = words[idx.reshape(-1)]
    g = torch.Generator(device=words.device)
    g.manual_seed(seed + 1)
    noise = torch.zeros_like(sl)
    for b in range(64):
        bit = (torch.rand(sl.shape, device=words.device, generator=g) < flip).to(torch.int64)
        noise |= bit << b
    qwords = sl ^ noise
    qoffs = np.arange(n_queries + 1, dtype=np.int64) * k
    truth = np.stack([tr, off], axis=1).astype(np.int64)
    return qwords.contiguous(), qoffs, truth
