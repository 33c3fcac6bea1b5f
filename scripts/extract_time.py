"""Quick device timing of the extraction stages (CQT, projection) on 3-min tracks. Run on the GPU box."""
import ctypes as C
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hpfw_b200
from hpfw_b200 import _lib, synth
from hpfw_b200._lib import check
from hpfw_b200.api import stream_arg

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 180.0
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
ctx = hpfw_b200.Context(0)
g = np.load("tests/golden/hashprint.npz")
f = np.ascontiguousarray(g["filters"]); check(ctx._lib.hpfw_set_filters(ctx.handle, f.ctypes.data_as(C.c_void_p)))
N = int(secs * 44100)
audio = torch.from_numpy(synth.synth_track(1, min(secs, 30.0), 44100)).cuda()
audio = audio.repeat((N + len(audio) - 1) // len(audio))[:N].contiguous()
cols = ctx._lib.hpfw_cqt_cols(N); words = cols - 99
spec = torch.empty((cols, 121), dtype=torch.float32, device="cuda")
hp = torch.empty(words, dtype=torch.int64, device="cuda")
s = torch.cuda.current_stream().cuda_stream
def run():
    check(ctx._lib.hpfw_calc_hashprint_audio_device(ctx.handle, C.c_void_p(audio.data_ptr()), N, C.c_void_p(hp.data_ptr()), stream_arg(s)))
run(); torch.cuda.synchronize()
ctx.timing_enable(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
cq, ncq = ctx.timing_read(_lib.K_CQT); pj, npj = ctx.timing_read(_lib.K_PROJECT)
frames = cols - 19
print(f"N={N} cols={cols} words={words}: {ms:.3f} ms/track  ({frames/ms*1e3/1e6:.2f} M frames/s); CQT kernels {cq/reps:.3f} ms ({ncq//reps} launches), "
      f"projection {pj/reps:.3f} ms; CQT compulsory bytes {(4*N+4*121*cols)/1e6:.1f} MB -> {(4*N+4*121*cols)/(cq/reps*1e-3)/1e9:.1f} GB/s; "
      f"projection {2*64*2420*frames/(pj/reps*1e-3)/1e12:.2f} TFLOP/s")
