"""Debug aid: exact copies planted at chosen offsets of one track, matched with each implementation."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hpfw_b200
from hpfw_b200 import MemoryStorage
from hpfw_b200._lib import check

ctx = hpfw_b200.Context(0)
rng = np.random.default_rng(3)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1200
k = int(sys.argv[2]) if len(sys.argv) > 2 else 50
words = rng.integers(0, 1 << 64, size=n, dtype=np.uint64)
offs = np.array([0, n], dtype=np.int64)
plant = [5, 100, 223, 224, 239, 240, 255, 256, 300, 447, 448, 470, 479, 480, 600, 700, 960, 1100]
plant = [p for p in plant if p + k <= n]
qs = [words[p:p + k].copy() for p in plant]
qw = np.concatenate(qs)
qo = np.arange(len(qs) + 1, dtype=np.int64) * k
st = MemoryStorage(ctx).build_packed(words, offs)
for impl in (0, 1, 3):
    check(ctx._lib.hpfw_set_match_impl(ctx.handle, impl))
    out = st.find_topk_packed(qw, qo, 1)
    print("impl", impl, [(int(o), int(c)) for o, c in zip(out["offset"][:, 0], out["cnt"][:, 0])])
print("plant  ", plant)
