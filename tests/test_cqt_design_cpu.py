"""Host-side CQT design (band layout) of the C-ABI library vs the oracle's nsg_design. No device needed."""
import ctypes as C

import numpy as np

from oracle import nsgcq


def _design(n):
    from hpfw_b200 import _lib
    L = _lib.load()
    pos = np.zeros(121, dtype=np.int32)
    lg = np.zeros(121, dtype=np.int32)
    m = C.c_int()
    assert L.hpfw_cqt_design(n, pos.ctypes.data_as(C.c_void_p), lg.ctypes.data_as(C.c_void_p), C.byref(m)) == 0
    return pos, lg, m.value


def test_design_matches_oracle_on_survey_sizes():
    from hpfw_b200 import _lib
    L = _lib.load()
    for n, m, cols in ((7938000, 43528, 14510), (1323000, 7255, 2419), (661500, 3627, 1210), (882000, 4836, 1613),
                       (264600, 1451, 484), (132300, 725, 242), (88200, 484, 162)):
        pos, lg, mm = _design(n)
        opos, olg, om = nsgcq.nsg_design(n)
        assert mm == m == om
        assert np.array_equal(pos, opos) and np.array_equal(lg, olg)
        assert L.hpfw_cqt_cols(n) == cols == nsgcq.spectrogram_cols(n)
        assert L.hpfw_hashprint_words_for_samples(n) == max(0, cols - 99)


def test_design_matches_oracle_on_many_lengths():
    rng = np.random.default_rng(0)
    for n in list(rng.integers(60000, 20_000_000, size=300)) + [54000, 54002, 2 ** 20, 2 ** 24]:
        pos, lg, m = _design(int(n))
        opos, olg, om = nsgcq.nsg_design(int(n))
        assert m == om and np.array_equal(pos, opos) and np.array_equal(lg, olg), n
