"""ShardedLiveSongIdentification with one rank against the direct single-GPU calls: index() (CQT, covariance, filter
learning, hashing, DB build) and search() give the same records. The multi-rank run is scripts/sharded_liveid_check.py
(torchrun, NCCL); its host-side collectives are covered on CPU by tests/test_sharded_cpu.py."""
import numpy as np
import pytest

import hpfw_b200
from hpfw_b200 import HashprintExtractor, MemoryStorage, synth
from hpfw_b200.sharded import ShardedLiveSongIdentification

pytestmark = pytest.mark.gpu


def test_one_rank_equals_direct_path(ctx):
    sr = 44100
    tracks = [synth.synth_track(300 + i, 8.0 + 2.0 * (i % 2), sr) for i in range(5)]
    queries, truth = [], []
    for i in (0, 3, 4):
        q, _ = synth.synth_query(tracks[i], 900 + i, 6.0, sr, max_semitones=0.2)
        queries.append(q)
        truth.append(i)
    lid = ShardedLiveSongIdentification(ctx).index(tracks, names=[f"t{i}" for i in range(5)])
    res = lid.search(queries, topk=3)
    assert [int(x) for x in res["track"][:, 0]] == truth
    assert [lid.names[int(x)] for x in res["track"][:, 0]] == [f"t{i}" for i in truth]
    ex = HashprintExtractor(ctx)
    ex.set_filters(lid.filters)
    st = MemoryStorage(ctx).build([(str(i), ex.calc_hashprint(t)) for i, t in enumerate(tracks)])
    ref = st.find_topk_packed(*hpfw_b200.api.pack([ex.calc_hashprint(q) for q in queries]), 3)
    for f in ("track", "cnt", "offset"):
        assert np.array_equal(ref[f], res[f]), f
    # search-only mode with given filters (the reference's cache/filters.cereal path) builds the same DB
    lid2 = ShardedLiveSongIdentification(ctx).index(tracks, filters=lid.filters)
    res2 = lid2.search(queries, topk=3)
    for f in ("track", "cnt", "offset"):
        assert np.array_equal(res2[f], res[f]), f
