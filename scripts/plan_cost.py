"""Cost of CQT plan creation: hashprints of tracks whose lengths all differ (every call plans; 2.2 - 3 min, random order)
vs equal lengths. GPU box."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hpfw_b200
from hpfw_b200._lib import check
from hpfw_b200.api import stream_arg

ctx = hpfw_b200.Context(0)
g = np.load("tests/golden/hashprint.npz")
ex = hpfw_b200.HashprintExtractor(ctx); ex.set_filters(g["filters"])
base = 180 * 44100
audio = (0.1 * torch.randn(base + 441 * 2 * 128, device="cuda")).contiguous()
hp = torch.zeros(20000, dtype=torch.int64, device="cuda")
s = torch.cuda.current_stream().cuda_stream
def one(n):
    check(ctx._lib.hpfw_calc_hashprint_audio_device(ctx.handle, C.c_void_p(audio.data_ptr()), n, C.c_void_p(hp.data_ptr()), stream_arg(s)))
one(base); torch.cuda.synchronize()
def timed(label, lens):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for n in lens: one(n)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{label}: {dt / len(lens) * 1e3:.2f} ms per track")
def smooth_lengths(lo, count):
    """even N >= lo with N/2 a product of 2,3,5,7 (packed two-pass FFT path)"""
    out, h = [], lo // 2
    while len(out) < count:
        m = h
        for p in (2, 3, 5, 7):
            while m % p == 0: m //= p
        if m == 1: out.append(2 * h)
        h += 1
    return out
rng = np.random.default_rng(3)
sm = smooth_lengths(base - 2_000_000, 120)
rng.shuffle(sm)                       # a library's track lengths come in no particular order
odd = [int(base - 2_000_000 + 2 * int(v) + 1) for v in rng.integers(0, 1_000_000, size=120)]
timed("equal lengths", [base] * 32)
timed("60 distinct smooth lengths, plan cache filling", sm[:60])
timed("60 more distinct smooth lengths, steady state (every plan evicts one)", sm[60:])
timed("60 distinct non-smooth lengths (Bluestein), cache filling", odd[:60])
timed("60 more non-smooth lengths, steady state", odd[60:])
