"""Cross-stream ordering of a context's shared scratch: every entry point of one hpfw_ctx shares buffers (the
covariance accumulator, the projection's delta matrix, the matcher's best[] array ...). A call that arrives on a different
stream than the previous one must first wait, on the device, for that previous work (hpfw_ctx::order_on in common.cuh)."""
import ctypes as C

import numpy as np
import pytest

from hpfw_b200 import HashprintExtractor
from hpfw_b200._lib import check
from hpfw_b200.api import stream_arg

pytestmark = pytest.mark.gpu


def test_reset_on_context_stream_then_accumulate_on_user_streams(ctx, hashprint_golden):
    import torch
    ex = HashprintExtractor(ctx)
    spec = np.ascontiguousarray(hashprint_golden["spec0"][:600])
    ex.cov_reset()
    ex.cov_add_spectrogram(spec)
    want = ex.cov_get()
    d_spec = torch.from_numpy(spec).cuda()
    # a large unrelated kernel keeps the user streams busy so that an unordered memset would overtake
    ballast = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    for it in range(12):
        s = streams[it & 1]
        # leave garbage from the previous round in the accumulator, then: reset (context stream) + add (user stream) + get
        with torch.cuda.stream(s):
            ballast.normal_()
            check(ctx._lib.hpfw_cov_reset(ctx.handle))
            check(ctx._lib.hpfw_cov_add_spectrogram_device(ctx.handle, C.c_void_p(d_spec.data_ptr()), spec.shape[0],
                                                           stream_arg(s.cuda_stream)))
        got = ex.cov_get()                      # host entry point on the context's own stream: ordered after the add
        assert np.array_equal(got, want), it
    torch.cuda.synchronize()


def test_projection_scratch_shared_by_two_streams(ctx, hashprint_golden):
    import torch
    ex = HashprintExtractor(ctx)
    ex.set_filters(np.ascontiguousarray(hashprint_golden["filters"]))
    specs = [np.ascontiguousarray(hashprint_golden["spec0"][:1200]), np.ascontiguousarray(hashprint_golden["q_spec"])]
    want = [ex.hashprint_from_spectrogram(sp) for sp in specs]
    d_specs = [torch.from_numpy(sp).cuda() for sp in specs]
    outs = [torch.zeros(len(w), dtype=torch.int64, device="cuda") for w in want]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    for it in range(10):
        for o in outs:
            o.zero_()
        torch.cuda.synchronize()
        for i in (0, 1, 0, 1):                  # back-to-back launches on alternating streams share delta_tc / colmeta
            co = np.array([0, specs[i].shape[0]], dtype=np.int64)
            check(ctx._lib.hpfw_hashprint_from_spectrogram_device(ctx.handle, C.c_void_p(d_specs[i].data_ptr()),
                                                                  co.ctypes.data_as(C.c_void_p), 1, C.c_void_p(outs[i].data_ptr()),
                                                                  stream_arg(streams[i].cuda_stream)))
        torch.cuda.synchronize()
        for i in (0, 1):
            assert np.array_equal(outs[i].cpu().numpy().view(np.uint64), want[i]), (it, i)
