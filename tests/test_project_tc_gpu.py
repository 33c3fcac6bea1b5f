"""The tcgen05 / TMEM / TMA projection (hpfw_b200/csrc/project_tc.cu) against the CUDA-core kernel, the oracle and the
reference-generated golden hashprints. Inputs are rounded to tf32 (impl 1, 2) or fp16 (impl 3, and impl 4 / 5 = the same arithmetic with 2 / 4 tiles per CTA sharing the filter stream) — a 10-bit mantissa either way — by this path, so bits may differ
where |delta| is within that rounding noise: the bar is north_star's >= 99.9 % of bits (SURVEY probe: 99.999 % expected for
delta-first tf32) and identical top-1."""
import ctypes as C

import numpy as np
import pytest

import oracle
from hpfw_b200 import HashprintExtractor, MemoryStorage
from hpfw_b200._lib import check
from hpfw_b200.api import stream_arg

pytestmark = pytest.mark.gpu


def _bits(a, b):
    return int(np.unpackbits((a ^ b).view(np.uint8)).sum())


@pytest.fixture()
def ex(ctx, hashprint_golden):
    e = HashprintExtractor(ctx)
    e.set_filters(hashprint_golden["filters"])
    yield e
    check(ctx._lib.hpfw_set_projection_impl(ctx.handle, 4))


@pytest.mark.parametrize("impl", [2, 1, 3, 4, 5])
def test_tc_vs_cuda_core_and_reference(ctx, ex, hashprint_golden, impl):
    g = hashprint_golden
    for spec_key, hp_key in (("q_spec", "hpq"), ("spec0", "hp0")):
        check(ctx._lib.hpfw_set_projection_impl(ctx.handle, 0))
        base = ex.hashprint_from_spectrogram(g[spec_key])
        check(ctx._lib.hpfw_set_projection_impl(ctx.handle, impl))
        hp = ex.hashprint_from_spectrogram(g[spec_key])
        assert hp.shape == base.shape == g[hp_key].shape
        total = 64 * len(hp)
        assert _bits(hp, base) <= 1e-3 * total, (impl, spec_key, _bits(hp, base), total)
        assert _bits(hp, g[hp_key]) <= 1e-3 * total
        # differing bits must be tiny-margin deltas (within tf32 rounding of the inputs: 2^-11 relative per term)
        hp64, delta = oracle.hashprint_f64(g[spec_key], g["filters"])
        wrong = np.unpackbits((hp ^ hp64).view(np.uint8).reshape(-1, 8)[:, ::-1], axis=1)[:, ::-1]
        typical = np.median(np.abs(delta))
        for t, bit in np.argwhere(wrong):
            assert abs(delta[t, 63 - int(bit)]) <= 2e-2 * typical


@pytest.mark.parametrize("impl", [2, 1, 3, 4, 5])
def test_tc_bit_order_single_tap_filters(ctx, impl):
    """One non-zero tap per filter: y is a copy of one band, the expected word is known in closed form (filter f -> bit
    63-f). Spectrogram values are tf32-exact so the tensor-core path must reproduce the comparison exactly."""
    rng = np.random.default_rng(3)
    cols = 400
    spec = (rng.integers(-80 * 8, 0, size=(cols, 121)) / 8.0).astype(np.float32)     # multiples of 1/8: exact in tf32 and fp16
    filt = np.zeros((2420, 64), dtype=np.float32)
    taps = [(int(rng.integers(0, 121)), int(rng.integers(0, 20))) for _ in range(64)]
    taps[0], taps[63] = (0, 0), (120, 19)
    for f, (b, c) in enumerate(taps):
        filt[b * 20 + c, f] = 1.0
    e = HashprintExtractor(ctx)
    e.set_filters(filt)
    check(ctx._lib.hpfw_set_projection_impl(ctx.handle, impl))
    try:
        hp = e.hashprint_from_spectrogram(spec)
    finally:
        check(ctx._lib.hpfw_set_projection_impl(ctx.handle, 4))
    n = cols - 99
    exp = np.zeros(n, dtype=np.uint64)
    for f, (b, c) in enumerate(taps):
        d = spec[c:c + n, b] - spec[c + 80:c + 80 + n, b]
        exp |= (d >= 0).astype(np.uint64) << np.uint64(63 - f)
    assert np.array_equal(hp, exp)


@pytest.mark.parametrize("impl", [2, 1, 3, 4, 5])
def test_tc_batched_ragged(ctx, ex, hashprint_golden, impl):
    import torch
    g = hashprint_golden
    specs = [g["spec0"], g["q_spec"], g["spec0"][:99], g["spec0"][:100], g["spec0"][300:700], g["spec0"][:227]]
    col_offs = np.zeros(len(specs) + 1, dtype=np.int64)
    np.cumsum([s.shape[0] for s in specs], out=col_offs[1:])
    d_spec = torch.from_numpy(np.concatenate(specs, axis=0)).cuda()
    n_words = [max(s.shape[0] - 99, 0) for s in specs]
    outs = []
    for which in (0, impl):
        check(ctx._lib.hpfw_set_projection_impl(ctx.handle, which))
        d_hp = torch.zeros(sum(n_words), dtype=torch.int64, device="cuda")
        check(ctx._lib.hpfw_hashprint_from_spectrogram_device(
            ctx.handle, C.c_void_p(d_spec.data_ptr()), col_offs.ctypes.data_as(C.c_void_p), len(specs),
            C.c_void_p(d_hp.data_ptr()), stream_arg(torch.cuda.current_stream().cuda_stream)))
        torch.cuda.synchronize()
        outs.append(d_hp.cpu().numpy().view(np.uint64))
    assert _bits(outs[0], outs[1]) <= 1e-3 * 64 * len(outs[0])
    # short tracks: exactly the same word count / placement
    pos = 0
    for n in n_words:
        if n:
            assert _bits(outs[0][pos:pos + n], outs[1][pos:pos + n]) <= max(2, 2e-3 * 64 * n)
        pos += n


@pytest.mark.parametrize("impl", [1, 3, 4])
def test_tc_identical_top1(ctx, ex, hashprint_golden, collector_golden, impl):
    g, c = hashprint_golden, collector_golden
    check(ctx._lib.hpfw_set_projection_impl(ctx.handle, impl))
    hpq = ex.hashprint_from_spectrogram(g["q_spec"])
    st = MemoryStorage(ctx).build_packed(c["words"], c["offs"])
    r = st.find(hpq)
    ref = oracle.find(c["words"], c["offs"], g["hpq"])
    assert (r.track, r.offset) == (ref[0], ref[2])
