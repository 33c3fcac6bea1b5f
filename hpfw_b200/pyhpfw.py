"""Drop-in for the reference's ctypes wrapper (modules/python/pyhpfw/pyhpfw.py:13-81): the same ParallelCollector class
and FilenameHashprintPair structure, bound to the par_collector_* symbols that libhpfw_b200.so re-exports
(include/hpfw_b200_pyhpfw.h). The only change a user of the reference makes is the library path.

Differences, all on the error side: a failed call raises HpfwError (the reference lets the C++ exception cross the C
boundary); prepare() returns (filename, hashprint) in that order as the reference's annotation says (its code builds
(hashprint, filename) tuples).
"""
from __future__ import annotations

import ctypes
from typing import List, Tuple

import numpy as np

from . import _lib


class FilenameHashprintPair(ctypes.Structure):
    _fields_ = [('filename', ctypes.c_char_p),
                ('hashprint', ctypes.POINTER(ctypes.c_uint64)),
                ('hp_size', ctypes.c_int)]


class ParallelCollector:
    def __init__(self, cache: str | None = None):
        L = self._lib = ctypes.CDLL(_lib.LIB_PATH)
        L.par_collector_new.restype = ctypes.c_void_p
        L.par_collector_prepare.restype = ctypes.POINTER(FilenameHashprintPair)
        L.par_collector_prepare.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_char_p), ctypes.c_int,
                                            ctypes.POINTER(ctypes.c_int)]
        L.par_collector_calc_hashprint.restype = ctypes.POINTER(ctypes.c_uint64)
        L.par_collector_calc_hashprint.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_int)]
        L.par_collector_del.argtypes = [ctypes.c_void_p]
        L.prepare_result_free.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.calc_hashprint_result_free.argtypes = [ctypes.POINTER(ctypes.c_uint64)]
        L.par_collector_load.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        L.par_collector_save.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        L.hpfw_last_error.restype = ctypes.c_char_p
        self._collector = L.par_collector_new()
        if cache:
            self.load(cache)

    def _fail(self, code=_lib.ERR_STATE):
        raise _lib.HpfwError(code, self._lib.hpfw_last_error().decode('utf-8', 'replace'))

    def prepare(self, filenames: List[str]) -> List[Tuple[str, np.ndarray]]:
        pyarr = [f.encode('utf-8') for f in filenames]
        arr = (ctypes.c_char_p * len(pyarr))(*pyarr)
        got = ctypes.c_int(0)
        hps = self._lib.par_collector_prepare(self._collector, arr, len(pyarr), ctypes.byref(got))
        if not hps:
            self._fail()
        out = [(hps[h].filename.decode('utf-8'),
                np.ctypeslib.as_array(hps[h].hashprint, shape=(hps[h].hp_size,)).astype(np.uint64).copy())
               for h in range(got.value)]
        self._lib.prepare_result_free(hps, got)
        return out

    def calc_hashprint(self, filename: str) -> np.ndarray:
        size = ctypes.c_int(0)
        hp = self._lib.par_collector_calc_hashprint(self._collector, filename.encode('utf-8'), ctypes.byref(size))
        if not hp:
            self._fail()
        out = np.ctypeslib.as_array(hp, shape=(size.value,)).astype(np.uint64).copy()
        self._lib.calc_hashprint_result_free(hp)
        return out

    def load(self, cache: str = ""):
        self._lib.par_collector_load(self._collector, cache.encode('utf-8'))

    def save(self, cache: str = ""):
        self._lib.par_collector_save(self._collector, cache.encode('utf-8'))

    def __del__(self):
        try:
            self._lib.par_collector_del(self._collector)
        except Exception:
            pass
