"""File writers for the bench and the tests: the reference's on-disk layouts (cereal binary archives are raw little-endian
bytes without a header) and canonical WAV files. Host-only helpers; nothing here is on the product path."""
from __future__ import annotations

from typing import Sequence

import numpy as np


def write_db_dump(path: str, names: Sequence[str], hashprints: Sequence[np.ndarray]) -> None:
    """db::MemoryStorage dump (/root/reference/include/hpfw/audioproblems/live-song-id/storage.h:67-76 = cereal of
    std::vector<FilenameFingerprintPair>): uint64 n; n x {uint64 len, name bytes, uint64 cnt, cnt x uint64 words}."""
    with open(path, "wb") as f:
        f.write(np.uint64(len(names)).tobytes())
        for name, hp in zip(names, hashprints):
            b = name.encode("utf-8")
            hp = np.ascontiguousarray(hp, dtype=np.uint64)
            f.write(np.uint64(len(b)).tobytes())
            f.write(b)
            f.write(np.uint64(len(hp)).tobytes())
            f.write(hp.tobytes())


def write_matrix_cereal(path: str, rows: int, cols: int, data_colmajor: np.ndarray) -> None:
    """cache/*.cereal and cache/spectros/<stem> (reference utils.h:77-90): int32 rows, int32 cols, column-major floats."""
    with open(path, "wb") as f:
        f.write(np.array([rows, cols], dtype=np.int32).tobytes())
        f.write(np.ascontiguousarray(data_colmajor, dtype=np.float32).tobytes())


def write_wav_pcm16(path: str, pcm: np.ndarray, sr: int) -> None:
    """Canonical 44-byte-header mono 16-bit PCM WAV."""
    x = np.ascontiguousarray(pcm, dtype=np.int16)
    hdr = (b"RIFF" + np.uint32(36 + x.nbytes).tobytes() + b"WAVEfmt " + np.uint32(16).tobytes() +
           np.uint16(1).tobytes() + np.uint16(1).tobytes() + np.uint32(sr).tobytes() + np.uint32(sr * 2).tobytes() +
           np.uint16(2).tobytes() + np.uint16(16).tobytes() + b"data" + np.uint32(x.nbytes).tobytes())
    with open(path, "wb") as f:
        f.write(hdr)
        f.write(x.tobytes())


def write_wav_f32(path: str, samples: np.ndarray, sr: int) -> None:
    """Canonical mono IEEE float32 WAV."""
    x = np.ascontiguousarray(samples, dtype=np.float32)
    hdr = (b"RIFF" + np.uint32(36 + x.nbytes).tobytes() + b"WAVEfmt " + np.uint32(16).tobytes() +
           np.uint16(3).tobytes() + np.uint16(1).tobytes() + np.uint32(sr).tobytes() + np.uint32(sr * 4).tobytes() +
           np.uint16(4).tobytes() + np.uint16(32).tobytes() + b"data" + np.uint32(x.nbytes).tobytes())
    with open(path, "wb") as f:
        f.write(hdr)
        f.write(x.tobytes())
