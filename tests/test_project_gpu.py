"""Parity of the CUDA projection/threshold/pack kernel (stages 2-3) with the reference-generated golden vectors and the
oracle. Floating point: tolerances are stated next to each assertion (north_star: >= 99.9 % of hashprint bits, fp32
relative tolerance for the projection itself)."""
import ctypes as C

import numpy as np
import pytest

import oracle
from hpfw_b200._lib import check
from hpfw_b200.api import stream_arg

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _cuda_core_kernel(ctx):
    """This file pins the fp32 CUDA-core kernel (impl 0); the default tcgen05 kernel is covered by test_project_tc_gpu.py."""
    check(ctx._lib.hpfw_set_projection_impl(ctx.handle, 0))
    yield
    check(ctx._lib.hpfw_set_projection_impl(ctx.handle, 3))


def _set_filters(ctx, filt):
    f = np.ascontiguousarray(filt, dtype=np.float32)
    check(ctx._lib.hpfw_set_filters(ctx.handle, f.ctypes.data_as(C.c_void_p)))


def _hashprint(ctx, spec):
    s = np.ascontiguousarray(spec, dtype=np.float32)
    cols = s.shape[0]
    hp = np.zeros(max(cols - 99, 1), dtype=np.uint64)
    n = C.c_int()
    check(ctx._lib.hpfw_hashprint_from_spectrogram(ctx.handle, s.ctypes.data_as(C.c_void_p), cols,
                                                   hp.ctypes.data_as(C.c_void_p), C.byref(n)))
    return hp[:n.value]


def _bits(a, b):
    return int(np.unpackbits((a ^ b).view(np.uint8)).sum())


def test_filters_roundtrip(ctx, hashprint_golden):
    _set_filters(ctx, hashprint_golden["filters"])
    back = np.zeros_like(hashprint_golden["filters"])
    check(ctx._lib.hpfw_get_filters(ctx.handle, back.ctypes.data_as(C.c_void_p)))
    assert np.array_equal(back, hashprint_golden["filters"])


def test_projection_values(ctx, hashprint_golden):
    g = hashprint_golden
    _set_filters(ctx, g["filters"])
    spec = np.ascontiguousarray(g["spec0"])
    cols = spec.shape[0]
    y = np.zeros((cols - 19, 64), dtype=np.float32)
    check(ctx._lib.hpfw_project(ctx.handle, spec.ctypes.data_as(C.c_void_p), cols, y.ctypes.data_as(C.c_void_p)))
    y64 = oracle.project_f64(spec, g["filters"])
    scale = np.abs(y64).max()
    # fp32 accumulation of 2420 terms: |err| <= 2e-5 * max|y| (observed ~3e-6); same bound the reference's own GEBP meets
    assert np.max(np.abs(y - y64)) <= 2e-5 * scale
    assert np.max(np.abs(y[:256] - g["y0_head"])) <= 2e-5 * scale


@pytest.mark.parametrize("spec_key,hp_key", [("spec0", "hp0"), ("q_spec", "hpq")])
def test_hashprint_bits_vs_reference(ctx, hashprint_golden, spec_key, hp_key):
    g = hashprint_golden
    _set_filters(ctx, g["filters"])
    hp = _hashprint(ctx, g[spec_key])
    ref = g[hp_key]
    assert hp.shape == ref.shape
    total = 64 * len(ref)
    diff = _bits(hp, ref)
    assert diff <= 1e-3 * total, f"{diff}/{total} bits differ from the reference"          # >= 99.9 % (north_star)
    # any differing bit must be a near-zero delta (|delta| below 1e-4 of the typical |delta|), never a real disagreement
    hp64, delta = oracle.hashprint_f64(g[spec_key], g["filters"])
    wrong = np.unpackbits((hp ^ hp64).view(np.uint8).reshape(-1, 8)[:, ::-1], axis=1)[:, ::-1]   # [n, 64], col = bit index
    # bit (63-f) = filter f
    idx = np.argwhere(wrong)
    typical = np.median(np.abs(delta))
    for t, bit in idx:
        f = 63 - int(bit)
        assert abs(delta[t, f]) <= 1e-4 * typical, (t, f, delta[t, f], typical)


def test_identical_top1_with_gpu_hashprints(ctx, hashprint_golden, collector_golden):
    """north_star: identical top-1 match when the query hashprint comes from the GPU instead of the reference."""
    from hpfw_b200 import MemoryStorage
    g, c = hashprint_golden, collector_golden
    _set_filters(ctx, g["filters"])
    hpq = _hashprint(ctx, g["q_spec"])
    st = MemoryStorage(ctx).build_packed(c["words"], c["offs"])
    r = st.find(hpq)
    ref = oracle.find(c["words"], c["offs"], g["hpq"])
    assert (r.track, r.offset) == (ref[0], ref[2])


def test_bit_order_and_ties(ctx):
    """Filter f -> bit 63-f; delta == 0 -> 1 (hashprint_handle.h:121,137-142). A filter bank with a single non-zero tap
    makes y a copy of one spectrogram band, so the expected word is known in closed form."""
    rng = np.random.default_rng(3)
    cols = 230
    spec = rng.uniform(-80, 0, size=(cols, 121)).astype(np.float32)
    filt = np.zeros((2420, 64), dtype=np.float32)
    taps = [(int(rng.integers(0, 121)), int(rng.integers(0, 20))) for _ in range(64)]
    for f, (b, c) in enumerate(taps):
        filt[b * 20 + c, f] = 1.0
    spec[100:, 7] = spec[20:cols - 80, 7]      # band 7 periodic with the lag: deltas exactly 0 there
    taps[5] = (7, 0)
    filt[:, 5] = 0
    filt[7 * 20, 5] = 1.0
    _set_filters(ctx, filt)
    hp = _hashprint(ctx, spec)
    n = cols - 99
    exp = np.zeros(n, dtype=np.uint64)
    for f, (b, c) in enumerate(taps):
        d = spec[c:c + n, b] - spec[c + 80:c + 80 + n, b]
        exp |= (d >= 0).astype(np.uint64) << np.uint64(63 - f)
    assert np.array_equal(hp, exp)
    assert np.all((hp[20:100] >> np.uint64(58)) & np.uint64(1))  # the exact-zero deltas (t in [20,100)) came out as 1


def test_batched_device_entry_equals_single(ctx, hashprint_golden):
    import torch
    g = hashprint_golden
    _set_filters(ctx, g["filters"])
    specs = [g["spec0"], g["q_spec"], g["spec0"][:99], g["spec0"][:100], g["spec0"][300:700]]   # 99 cols -> 0 words
    col_offs = np.zeros(len(specs) + 1, dtype=np.int64)
    np.cumsum([s.shape[0] for s in specs], out=col_offs[1:])
    d_spec = torch.from_numpy(np.concatenate(specs, axis=0)).cuda()
    n_words = [max(s.shape[0] - 99, 0) for s in specs]
    d_hp = torch.zeros(sum(n_words), dtype=torch.int64, device="cuda")
    check(ctx._lib.hpfw_hashprint_from_spectrogram_device(
        ctx.handle, C.c_void_p(d_spec.data_ptr()), col_offs.ctypes.data_as(C.c_void_p), len(specs),
        C.c_void_p(d_hp.data_ptr()), stream_arg(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    got = d_hp.cpu().numpy().view(np.uint64)
    pos = 0
    for s, n in zip(specs, n_words):
        if n:
            assert np.array_equal(got[pos:pos + n], _hashprint(ctx, s))
        pos += n


def test_too_short_is_an_error(ctx, hashprint_golden):
    from hpfw_b200 import HpfwError
    from hpfw_b200._lib import ERR_SHORT
    _set_filters(ctx, hashprint_golden["filters"])
    with pytest.raises(HpfwError) as e:
        _hashprint(ctx, hashprint_golden["spec0"][:99])
    assert e.value.code == ERR_SHORT
