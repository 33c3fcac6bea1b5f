"""Integer-pipe matcher (matcher.cu) against the tensor-core matcher (match_tc.cu) on the bench geometry: identical keys,
device time of each (CUDA events on the launching stream). Run on the GPU box:

    python scripts/match_tc_time.py [tracks=2000] [queries=128] [k=385] [track_words=14411]
"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import hpfw_b200
from hpfw_b200 import MemoryStorage, synth
from hpfw_b200._lib import check


def main():
    tracks = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    nq = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    k = int(sys.argv[3]) if len(sys.argv) > 3 else 385
    tw = int(sys.argv[4]) if len(sys.argv) > 4 else 14411
    dev = torch.device("cuda:0")
    ctx = hpfw_b200.Context(0)
    words, offs = synth.device_hashprint_db(torch, dev, 11, tracks, tw)
    st = MemoryStorage(ctx).build_device(words.data_ptr(), offs)
    qw, qo, truth = synth.device_hashprint_queries(torch, words, offs, 12, nq, k)
    keys = {}
    out = {"tracks": tracks, "queries": nq, "k": k, "track_words": tw}
    stream = torch.cuda.current_stream().cuda_stream
    wordops = ctx._lib.hpfw_db_word_ops(st._db, qo.ctypes.data_as(C.c_void_p), nq)
    impls = (3,) if os.environ.get("HPFW_TC_TIME_ONLY_F4") else (0, 1, 3)
    for impl in impls:
        check(ctx._lib.hpfw_set_match_impl(ctx.handle, impl))
        kk = torch.empty((nq, 10), dtype=torch.int64, device=dev)
        for _ in range(2):
            st.match_device(qw.data_ptr(), qo, 10, kk.data_ptr(), stream)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        a.record()
        for _ in range(reps):
            st.match_device(qw.data_ptr(), qo, 10, kk.data_ptr(), stream)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        keys[impl] = kk.cpu().numpy()
        out[f"impl{impl}_ms"] = ms
        out[f"impl{impl}_gwordops"] = wordops / ms / 1e6
        out[f"impl{impl}_qps_10k"] = nq / (ms / 1e3) * tracks / 10000.0
    if len(impls) == 1:       # HPFW_TC_TIME_ONLY_F4: timing of the default kernel only (tuning sweeps)
        print(json.dumps(out))
        return
    out["equal"] = bool(np.array_equal(keys[0], keys[1]))
    out["equal_fp4"] = bool(np.array_equal(keys[0], keys[3]))
    out["fp4_tops_equiv"] = wordops * 128 / (out["impl3_ms"] / 1e3) / 1e12
    rec = hpfw_b200.api.decode_keys(keys[1].view(np.uint64))
    out["top1_ok"] = float(np.mean((rec["track"][:, 0] == truth[:, 0]) & (rec["offset"][:, 0] == truth[:, 1])))
    if not out["equal"]:
        bad = np.argwhere(keys[0] != keys[1])
        out["first_diff"] = [int(x) for x in bad[0]]
        out["n_diff"] = int(len(bad))
        q, r = bad[0]
        out["k0"] = hex(int(keys[0][q, r]) & (2**64 - 1))
        out["k1"] = hex(int(keys[1][q, r]) & (2**64 - 1))
    out["tc_tops"] = wordops * 128 / (out["impl1_ms"] / 1e3) / 1e12
    print(json.dumps(out))


if __name__ == "__main__":
    main()
