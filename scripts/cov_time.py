"""Device time of the covariance accumulate (hpfw_cov_add_spectrogram_device, learn.cu + cov_tc.cu) per 3-minute spectrogram.
Run on the GPU box: python scripts/cov_time.py [tracks=32]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hpfw_b200
from hpfw_b200._lib import check
from hpfw_b200.api import stream_arg
ntr = int(sys.argv[1]) if len(sys.argv) > 1 else 32
ctx = hpfw_b200.Context(0)
cols = 14510
spec = (torch.randn(ntr, cols, 121, device="cuda") * 10 - 40).contiguous()
s = torch.cuda.current_stream().cuda_stream
def run():
    check(ctx._lib.hpfw_cov_reset(ctx.handle))
    for i in range(ntr):
        check(ctx._lib.hpfw_cov_add_spectrogram_device(ctx.handle, C.c_void_p(spec[i].data_ptr()), cols, stream_arg(s)))
run(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
acc = np.zeros((2420, 2420), dtype=np.float32)
check(ctx._lib.hpfw_cov_get(ctx.handle, acc.ctypes.data_as(C.c_void_p)))
print(f"cov: {e0.elapsed_time(e1) / ntr * 1e3:.1f} us/track, checksum {float(np.abs(acc).sum()):.6e}, symmetric {bool(np.array_equal(acc, acc.T))}")
