"""Device timing of the projection kernels (CUDA-core vs tcgen05) on a batch of 3-min spectrograms. Run on the GPU box."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hpfw_b200
from hpfw_b200 import _lib
from hpfw_b200._lib import check
from hpfw_b200.api import stream_arg
ntr = int(sys.argv[1]) if len(sys.argv) > 1 else 32
ctx = hpfw_b200.Context(0)
g = np.load("tests/golden/hashprint.npz")
ex = hpfw_b200.HashprintExtractor(ctx); ex.set_filters(g["filters"])
cols = 14510
spec = torch.from_numpy(np.tile(g["spec0"], (12, 1))[:cols]).cuda()
big = spec.repeat(ntr, 1).contiguous() + 0.01 * torch.randn(ntr * cols, 121, device="cuda")
offs = np.arange(ntr + 1, dtype=np.int64) * cols
words = cols - 99
s = torch.cuda.current_stream().cuda_stream
res = {}
impls = tuple(int(x) for x in sys.argv[2].split(",")) if len(sys.argv) > 2 else (0, 2, 1)
for impl in impls:
    check(ctx._lib.hpfw_set_projection_impl(ctx.handle, impl))
    hp = torch.zeros(ntr * words, dtype=torch.int64, device="cuda")
    def run():
        check(ctx._lib.hpfw_hashprint_from_spectrogram_device(ctx.handle, C.c_void_p(big.data_ptr()), offs.ctypes.data_as(C.c_void_p), ntr, C.c_void_p(hp.data_ptr()), stream_arg(s)))
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    res[impl] = hp.cpu().numpy().view(np.uint64)
    fl = 2.0 * 64 * 2420 * (cols - 19) * ntr
    print(f"impl {impl}: {ms:.3f} ms for {ntr} tracks = {ms/ntr*1e3:.1f} us/track, {fl/ms/1e9:.1f} TFLOP/s (algorithmic)")
for impl in [i for i in impls if i != 0 and 0 in res]:
    d = int(np.unpackbits((res[impl] ^ res[0]).view(np.uint8)).sum())
    print(f"impl {impl} vs 0: {d} of {64*len(res[0])} bits differ ({100.0*d/(64*len(res[0])):.5f} %)")
for impl in [i for i in impls if i in (4, 5) and 3 in res]:
    print(f"impl {impl} vs 3: identical = {bool(np.array_equal(res[impl], res[3]))}")
