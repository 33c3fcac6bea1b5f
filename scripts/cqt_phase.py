"""Per-phase clock cycles of the two main FFT passes (variant build -DHPFW_CQT_PHASECLK):
HPFW_B200_LIB=hpfw_b200/libhpfw_b200_phase.so python scripts/cqt_phase.py [tracks]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hpfw_b200
from hpfw_b200 import _lib
ntr = int(sys.argv[1]) if len(sys.argv) > 1 else 48
ctx = hpfw_b200.Context(0)
g = np.load("tests/golden/hashprint.npz")
ex = hpfw_b200.HashprintExtractor(ctx); ex.set_filters(g["filters"])
N = 7938000
audio = (0.1 * torch.randn(ntr, N, device="cuda")).contiguous()
words = ex.words(N)
hp = torch.zeros(ntr * words, dtype=torch.int64, device="cuda")
offs = np.arange(ntr + 1, dtype=np.int64) * N
s = torch.cuda.current_stream().cuda_stream
lib = _lib.load()
fn = lib.hpfw_cqt_debug_phases
fn.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
def run(): ex.calc_hashprint_batch_device(audio.data_ptr(), offs, hp.data_ptr(), s)
run(); torch.cuda.synchronize()
buf = (C.c_ulonglong * 12)()
fn(buf, 1)
run(); torch.cuda.synchronize()
fn(buf, 1)
v = np.array(list(buf), dtype=np.float64).reshape(2, 6)
names = ["first stage (own work)", "barrier after it", "middle stages", "last stage", "barrier", "store"]
for m, ctas in ((0, 1050), (1, 945)):
    tot = v[m].sum()
    print(f"pass {'AB'[m]}: {tot / (ntr * ctas):.0f} clk per CTA (thread 0)")
    for i in range(6):
        if v[m][i]: print(f"   {names[i]:28s} {v[m][i] / (ntr * ctas):8.0f} clk  {100 * v[m][i] / tot:5.1f} %")
