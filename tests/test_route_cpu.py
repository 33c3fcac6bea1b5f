"""Host logic of the matcher's kernel routing (hpfw_match_route, matcher.cu: route_queries): which queries of a batch join a
tensor-core group of up to 128 and which stay on the integer-pipe kernel. Needs no GPU."""
import ctypes as C

import numpy as np
import pytest

from hpfw_b200 import _lib


def route(lens, impl=2, fp4=1):
    L = _lib.load()
    lens = np.asarray(lens, dtype=np.int64)
    qo = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens, out=qo[1:])
    out = np.full(len(lens), -7, dtype=np.int32)
    _lib.check(L.hpfw_match_route(qo.ctypes.data_as(C.c_void_p), len(lens), impl, fp4, out.ctypes.data_as(C.c_void_p)))
    return out


def test_every_query_is_assigned_once_and_groups_hold_at_most_128():
    rng = np.random.default_rng(1)
    lens = rng.choice([0, 1, 63, 143, 385, 386, 707, 1514, 4096], size=1000)
    for impl in (0, 1, 2, 3):
        g = route(lens, impl)
        assert (g >= -1).all()
        if impl == 0:
            assert (g == -1).all()
        if impl in (1, 3):
            assert (g >= 0).all()
        ids, counts = np.unique(g[g >= 0], return_counts=True)
        assert (counts <= 128).all()
        assert np.array_equal(ids, np.arange(len(ids)))
        # groups are runs of the length-sorted order: the length ranges of two groups never interleave
        spans = sorted((lens[g == i].min(), lens[g == i].max()) for i in ids)
        for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
            assert a1 <= b0


def test_single_queries_stay_on_the_integer_pipes():
    assert route([385])[0] == -1
    assert (route([385] * 6) == -1).all()
    assert (route([385] * 14, fp4=0) == -1).all()
    assert route([])[...].size == 0


def test_a_batch_of_similar_queries_goes_to_the_tensor_cores():
    assert (route([385] * 128) == 0).all()
    g = route([385] * 300)
    assert (g >= 0).all() and len(np.unique(g)) == 3
    assert (route([385] * 12) == 0).all()            # fp4: a group pays off from ~9 equal queries
    assert (route([385] * 24, fp4=0) == 0).all()     # int8: from ~18


def test_long_stragglers_are_not_padded_onto_a_group_of_short_queries():
    lens = [63] * 128 + [1514] * 2          # 2 long queries: cheaper on the integer pipes than as a group of their own
    g = route(lens)
    assert (g[:128] == 0).all() and (g[128:] == -1).all()
    # ... but enough long queries form their own group
    lens = [63] * 128 + [1514] * 40
    g = route(lens)
    # (the group that closes on the long queries is filled up with the short ones before it: same cost, two groups)
    assert (g >= 0).all() and len(np.unique(g)) == 2 and (g[128:] == g[128]).all()


def test_short_queries_ride_along_for_free():
    # 100 long queries leave 28 free rows in their group: the 5 short ones join instead of running on the integer pipes
    g = route([385] * 100 + [63] * 5)
    assert (g == 0).all()
