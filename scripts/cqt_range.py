"""One batch of CQT + projection between cudaProfilerStart / Stop, for `ncu --replay-mode app-range` (DRAM / L2 bytes of the whole
batch with the lanes running concurrently, which a per-kernel capture cannot show):
  ncu --replay-mode app-range --cache-control none --clock-control none \
      --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum python scripts/cqt_range.py 48"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hpfw_b200
ntr = int(sys.argv[1]) if len(sys.argv) > 1 else 48
N = int(sys.argv[2]) if len(sys.argv) > 2 else 7938000
ctx = hpfw_b200.Context(0)
g = np.load("tests/golden/hashprint.npz")
ex = hpfw_b200.HashprintExtractor(ctx); ex.set_filters(g["filters"])
audio = (0.1 * torch.randn(ntr, N, device="cuda")).contiguous()
words = ex.words(N)
hp = torch.zeros(ntr * words, dtype=torch.int64, device="cuda")
offs = np.arange(ntr + 1, dtype=np.int64) * N
s = torch.cuda.current_stream().cuda_stream
def run(): ex.calc_hashprint_batch_device(audio.data_ptr(), offs, hp.data_ptr(), s)
run(); run(); torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
run(); torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print(f"range: {ntr} tracks of {N} samples")
