"""Known-answer self-test of the tensor-core matcher (matcher.cu: match_tc_selftest).

The fp4 / int8 GEMM formulation of MemoryStorage::find (/root/reference/include/hpfw/audioproblems/live-song-id/
storage.h:27-64) is only exact if tcgen05.mma accumulates +-1 products exactly up to 2^18. A context proves that against the
integer-pipe kernel before the first tensor-core match; when the proof fails the route is disabled LOUDLY (HPFW_ERR_STATE
for impl 1 / 3, integer pipes for the default routing) — never a CPU fallback, never a silently different result."""
import numpy as np
import pytest

import hpfw_b200
import oracle
from hpfw_b200 import MemoryStorage, synth
from hpfw_b200._lib import ERR_STATE, check

pytestmark = pytest.mark.gpu


def test_selftest_passes_on_this_device():
    ctx = hpfw_b200.Context(0)
    try:
        for fp4 in (1, 0):
            check(ctx._lib.hpfw_match_tc_selftest(ctx.handle, fp4))
            check(ctx._lib.hpfw_match_tc_selftest(ctx.handle, fp4))      # second call: the recorded verdict
        check(ctx._lib.hpfw_set_match_impl(ctx.handle, 3))
        check(ctx._lib.hpfw_set_match_impl(ctx.handle, 1))
    finally:
        ctx.close()


def test_failed_selftest_is_loud_and_keeps_results_exact(monkeypatch, capfd):
    monkeypatch.setenv("HPFW_MATCH_TC_SELFTEST_FAIL", "1")
    ctx = hpfw_b200.Context(0)
    try:
        for impl in (3, 1):
            with pytest.raises(hpfw_b200.HpfwError) as e:
                check(ctx._lib.hpfw_set_match_impl(ctx.handle, impl))
            assert e.value.code == ERR_STATE and "self-test" in str(e.value)
        assert "FAILED its known-answer self-test" in capfd.readouterr().err
        # default routing: a batch that would go to the tensor cores runs on the integer pipes and stays bit-exact
        words, offs = synth.synth_hashprint_db(5, 20, 700)
        qw, qo, _ = synth.synth_hashprint_queries(6, words, offs, 64, 96)
        st = MemoryStorage(ctx).build_packed(words, offs)
        ctx.timing_enable(True)
        launches0 = ctx.launch_count()
        got = st.find_topk_packed(qw, qo, 3)
        tr, d, o = oracle.find_topk_batch(words, offs, qw, qo, 3, 4)
        assert np.array_equal(got["track"], tr) and np.array_equal(got["cnt"], d) and np.array_equal(got["offset"], o)
        assert ctx.launch_count() > launches0          # GPU kernels did the work
        ms, n = ctx.timing_read(hpfw_b200._lib.K_MATCH_TC)
        assert n == 0                                   # ... and none of them was match_tc_kernel
        assert ctx.timing_read(hpfw_b200._lib.K_MATCH)[1] > 0
        del st                                          # a database goes before its context
    finally:
        ctx.close()
