"""ctypes binding of the C ABI (include/hpfw_b200.h). Fails loudly if the CUDA library is missing: no fallback."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# HPFW_B200_LIB: A/B experiments with an alternative build of the same library (never a different implementation)
LIB_PATH = os.environ.get("HPFW_B200_LIB") or os.path.join(HERE, "libhpfw_b200.so")


class HpfwError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"hpfw_b200 error {code}: {msg}")
        self.code = code


class Match(C.Structure):
    """= hpfw_match = db::MemoryStorage::SearchResult (storage.h:11-15) with the filename as a DB index."""
    _fields_ = [("track", C.c_int64), ("cnt", C.c_uint64), ("offset", C.c_int64)]


K_MATCH, K_TOPK, K_PROJECT, K_CQT, K_OTHER, K_MATCH_TC = range(6)
OK, ERR_CUDA, ERR_ARG, ERR_LIMIT, ERR_STATE, ERR_SHORT = 0, -1, -2, -3, -4, -5

_SIGS = {
    # name: (restype, argtypes)
    "hpfw_last_error": (C.c_char_p, []),
    "hpfw_version": (C.c_char_p, []),
    "hpfw_device_count": (C.c_int, []),
    "hpfw_ctx_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "hpfw_ctx_destroy": (None, [C.c_void_p]),
    "hpfw_ctx_device": (C.c_int, [C.c_void_p]),
    "hpfw_ctx_launch_count": (C.c_uint64, [C.c_void_p]),
    "hpfw_ctx_synchronize": (C.c_int, [C.c_void_p]),
    "hpfw_ctx_timing_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "hpfw_ctx_timing_read": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.c_int]),
    "hpfw_db_build": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.POINTER(C.c_void_p)]),
    "hpfw_db_build_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p,
                                       C.POINTER(C.c_void_p)]),
    "hpfw_db_build_gather_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p,
                                              C.POINTER(C.c_void_p)]),
    "hpfw_db_destroy": (None, [C.c_void_p]),
    "hpfw_db_tracks": (C.c_int, [C.c_void_p]),
    "hpfw_db_words": (C.c_int64, [C.c_void_p]),
    "hpfw_db_find": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(Match)]),
    "hpfw_db_find_topk": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(Match)]),
    "hpfw_db_match_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "hpfw_db_find_topk_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(Match), C.c_void_p]),
    "hpfw_set_match_impl": (C.c_int, [C.c_void_p, C.c_int]),
    "hpfw_match_tc_selftest": (C.c_int, [C.c_void_p, C.c_int]),
    "hpfw_match_route": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "hpfw_topk_merge_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "hpfw_keys_decode": (None, [C.c_void_p, C.c_int, C.POINTER(Match)]),
    "hpfw_db_word_ops": (C.c_double, [C.c_void_p, C.c_void_p, C.c_int]),
    "hpfw_set_filters": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hpfw_get_filters": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hpfw_hashprint_words_for_cols": (C.c_int, [C.c_int]),
    "hpfw_hashprint_from_spectrogram": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int)]),
    "hpfw_hashprint_from_spectrogram_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                                         C.c_void_p]),
    "hpfw_set_projection_impl": (C.c_int, [C.c_void_p, C.c_int]),
    "hpfw_project": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "hpfw_cov_reset": (C.c_int, [C.c_void_p]),
    "hpfw_cov_set": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hpfw_cov_get": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hpfw_cov_get_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "hpfw_cov_set_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "hpfw_cov_add_spectrogram": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "hpfw_cov_add_spectrogram_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "hpfw_calc_filters": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hpfw_set_cqt_window": (C.c_int, [C.c_void_p, C.c_int]),
    "hpfw_cqt_cols": (C.c_int, [C.c_int64]),
    "hpfw_cqt_spectrogram": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(C.c_int)]),
    "hpfw_cqt_spectrogram_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "hpfw_cqt_magnitude": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(C.c_int)]),
    "hpfw_cqt_design": (C.c_int, [C.c_int64, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]),
    "hpfw_fft_c2c": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "hpfw_hashprint_words_for_samples": (C.c_int, [C.c_int64]),
    "hpfw_calc_hashprint_audio": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(C.c_int)]),
    "hpfw_calc_hashprint_audio_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "hpfw_calc_hashprint_audio_batch_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                                         C.c_void_p]),
    "hpfw_pcm16_to_float_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "hpfw_cqt_spectrogram_pcm16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(C.c_int)]),
    "hpfw_calc_hashprint_pcm16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(C.c_int)]),
    "hpfw_calc_hashprint_pcm16_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "hpfw_calc_hashprint_pcm16_batch_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                                         C.c_void_p]),
    "hpfw_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "hpfw_host_free": (None, [C.c_void_p]),
    "hpfw_xs_create": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t, C.POINTER(C.c_void_p)]),
    "hpfw_xs_destroy": (None, [C.c_void_p]),
    "hpfw_xs_acquire": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_void_p)]),
    "hpfw_xs_release": (C.c_int, [C.c_void_p, C.c_int]),
    "hpfw_xs_submit": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.POINTER(C.c_int)]),
    "hpfw_xs_submit_spectrogram": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "hpfw_xs_tracks": (C.c_int, [C.c_void_p]),
    "hpfw_xs_track_info": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "hpfw_xs_fetch_spectrogram": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "hpfw_xs_wait": (C.c_int, [C.c_void_p]),
    "hpfw_xs_hash_kept": (C.c_int, [C.c_void_p]),
    "hpfw_xs_drop_kept": (C.c_int, [C.c_void_p]),
    "hpfw_xs_reset": (C.c_int, [C.c_void_p]),
    "hpfw_xs_hashprints_device": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p]),
    "hpfw_xs_hashprint_host": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "hpfw_xs_hashprints_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    "hpfw_xs_build_db": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.POINTER(C.c_void_p)]),
    "hpfw_xs_match": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(Match)]),
    "hpfw_shard_plan": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "hpfw_shard_plan_weighted": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "hpfw_shard_nccl_version": (C.c_int, []),
    "hpfw_shard_unique_id": (C.c_int, [C.c_void_p]),
    "hpfw_shard_create_rank": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "hpfw_shard_create_local": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]),
    "hpfw_shard_destroy": (None, [C.c_void_p]),
    "hpfw_shard_world": (C.c_int, [C.c_void_p]),
    "hpfw_shard_rank": (C.c_int, [C.c_void_p]),
    "hpfw_shard_ctx": (C.c_void_p, [C.c_void_p, C.c_int]),
    "hpfw_shard_db": (C.c_void_p, [C.c_void_p, C.c_int]),
    "hpfw_shard_build_rank": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64]),
    "hpfw_shard_build_rank_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p]),
    "hpfw_shard_build": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "hpfw_shard_build_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "hpfw_shard_match_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "hpfw_shard_find_topk": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(Match)]),
    "hpfw_shard_find_topk_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(Match)]),
    "hpfw_shard_allreduce_cov": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hpfw_shard_allgather_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "hpfw_shard_allgatherv_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hpfw_shard_broadcast_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]),
    "hpfw_microbench_pipes": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
}

# every symbol include/hpfw_b200.h declares (tests/test_abi_cpu.py checks the header against this and the .so)
ABI_SYMBOLS = tuple(_SIGS)

_lib = None


def load():
    """dlopen the C-ABI library. Raises (never falls back) when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python -m hpfw_b200.build` (nvcc, sm_100a). "
                "hpfw_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)  # AttributeError = ABI symbol missing: fail loudly
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status: int) -> None:
    if status != OK:
        raise HpfwError(status, load().hpfw_last_error().decode("utf-8", "replace"))
