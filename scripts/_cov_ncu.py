import sys
sys.path.insert(0, "/root/repo")
import numpy as np, hpfw_b200
ctx = hpfw_b200.Context(0); ex = hpfw_b200.HashprintExtractor(ctx)
spec = (np.random.default_rng(0).standard_normal((14510, 121)) * 3 - 40).astype(np.float32)
ex.cov_reset()
for _ in range(3): ex.cov_add_spectrogram(spec)
