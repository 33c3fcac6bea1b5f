"""hpfw_b200 — B200-native hashprint feature-to-match path behind hpfw's API.

Python side = a ctypes mirror of the reference's interfaces (pyhpfw.ParallelCollector, db::MemoryStorage,
LiveSongIdentification) over the C ABI in include/hpfw_b200.h. All compute is in libhpfw_b200.so (hand-written sm_100a
CUDA); importing this package without that library raises.
"""
from ._lib import HpfwError, Match, load  # noqa: F401
from .api import Context, HashprintExtractor, MemoryStorage, SearchResult  # noqa: F401

__all__ = ["Context", "HashprintExtractor", "MemoryStorage", "SearchResult", "HpfwError", "Match", "load"]
