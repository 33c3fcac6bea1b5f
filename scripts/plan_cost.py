"""Cost of CQT plan creation: hashprints of tracks whose lengths all differ (every call plans) vs equal lengths. GPU box."""
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
audio = (0.1 * torch.randn(base + 441 * 64, device="cuda")).contiguous()
hp = torch.zeros(20000, dtype=torch.int64, device="cuda")
s = torch.cuda.current_stream().cuda_stream
def one(n):
    check(ctx._lib.hpfw_calc_hashprint_audio_device(ctx.handle, C.c_void_p(audio.data_ptr()), n, C.c_void_p(hp.data_ptr()), stream_arg(s)))
one(base); torch.cuda.synchronize()
for label, lens in (("equal lengths", [base] * 32), ("32 distinct smooth lengths", [base + 441 * 2 * i for i in range(1, 33)]),
                    ("8 distinct non-smooth lengths", [base + 2 * i + 1 for i in range(1, 9)])):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for n in lens: one(n)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{label}: {dt / len(lens) * 1e3:.2f} ms per track")
