"""CQT batch timing under tuning overrides (env HPFW_CQT_*). Run on the GPU box: python scripts/cqt_tune.py [tracks]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hpfw_b200
from hpfw_b200 import _lib
ntr = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ctx = hpfw_b200.Context(0)
g = np.load("tests/golden/hashprint.npz")
ex = hpfw_b200.HashprintExtractor(ctx); ex.set_filters(g["filters"])
N = int(sys.argv[2]) if len(sys.argv) > 2 else 7938000      # samples per track (default: 3 minutes)
audio = (0.1 * torch.randn(ntr, N, device="cuda")).contiguous()
t = torch.arange(N, device="cuda", dtype=torch.float32) / 44100
audio += 0.2 * torch.sin(2 * np.pi * 440.0 * t) + 0.2 * torch.sin(2 * np.pi * 1318.5 * t)
words = ex.words(N)
hp = torch.zeros(ntr * words, dtype=torch.int64, device="cuda")
offs = np.arange(ntr + 1, dtype=np.int64) * N
s = torch.cuda.current_stream().cuda_stream
def run(): ex.calc_hashprint_batch_device(audio.data_ptr(), offs, hp.data_ptr(), s)
run(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
import time
e0.record()
h0 = time.perf_counter()
for _ in range(3): run()
host_us = (time.perf_counter() - h0) / 3 / ntr * 1e6     # host time to enqueue one track's launches (no sync)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3 / ntr
env = {k: v for k, v in os.environ.items() if k.startswith("HPFW_CQT")}
cols = ex.words(N) + 99
print(f"N={N} {env}: host enqueue {host_us:.1f} us/track; {ms*1e3:.1f} us/track ({(words)/ms/1e3:.2f} M frames/s), checksum {int(hp[:words].sum().item()) & 0xFFFFFFFF:08x}")
