"""Index-time filter learning on the GPU (row a10): structured covariance vs the reference's calc_cov, subspace-iteration
filters vs a dense eigen-solve, and index() learning filters from scratch through the C++ API.

Tolerances: covariance |err| <= 2e-5 * max|cov| (fp32, different summation order from Eigen's SYRK); eigenvalues 1e-4 relative;
eigenvectors |cos| >= 0.9999 wherever the eigenvalue is separated from its neighbours by >= 1 % (the direction of an
eigenvector inside a near-degenerate cluster is not determined by the matrix to working precision — any two solvers
differ there — so those are compared as a subspace)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle
from hpfw_b200 import HashprintExtractor

pytestmark = pytest.mark.gpu


def _np_cov(spec):
    """calc_cov(calc_frames(S)^T) in float64 (hashprint_handle.h:79-102)."""
    cols = spec.shape[0]
    nf = cols - 19
    X = np.empty((nf, 2420), dtype=np.float64)
    for b in range(121):
        for c in range(20):
            X[:, b * 20 + c] = spec[c:c + nf, b]
    X -= X.mean(axis=0)
    return X.T @ X / (nf - 1)


def test_covariance_vs_float64_and_reference(ctx, hashprint_golden):
    ex = HashprintExtractor(ctx)
    spec = hashprint_golden["q_spec"]                 # 242 columns
    ex.cov_reset()
    ex.cov_add_spectrogram(spec)
    got = ex.cov_get()
    ref = _np_cov(spec)
    scale = np.abs(ref).max()
    assert np.max(np.abs(got - ref)) <= 2e-5 * scale
    assert np.array_equal(got, got.T)                 # exactly symmetric by construction
    # accumulation: a second track adds
    spec0 = hashprint_golden["spec0"]
    ex.cov_add_spectrogram(spec0)
    both = ex.cov_get()
    ref0 = _np_cov(spec0)
    assert np.max(np.abs(both - (ref + ref0))) <= 2e-5 * np.abs(ref + ref0).max()
    if oracle.ref_available():                        # the reference's own calc_cov (Eigen SYRK, fp32)
        out = np.zeros(2420 * 2420, dtype=np.float32)
        oracle.ref().ref_calc_cov(np.ascontiguousarray(spec0).reshape(-1), spec0.shape[0], out)
        ex.cov_reset()
        ex.cov_add_spectrogram(spec0)
        assert np.max(np.abs(ex.cov_get() - out.reshape(2420, 2420))) <= 5e-5 * np.abs(ref0).max()


def test_cov_set_get_roundtrip(ctx):
    ex = HashprintExtractor(ctx)
    a = np.random.default_rng(0).standard_normal((2420, 2420)).astype(np.float32)
    ex.cov_set(a)
    assert np.array_equal(ex.cov_get(), a)
    ex.cov_reset()
    assert not ex.cov_get().any()


def test_calc_filters_known_spectrum(ctx):
    """SURVEY Appendix A probe (4): for diag(1..n) the first filter picks the last index (largest eigenvalue first)."""
    ex = HashprintExtractor(ctx)
    cov = np.diag(np.arange(1, 2421, dtype=np.float32))
    f, w = ex.calc_filters(cov, install=False)
    F = f.T                                           # [64, 2420]
    for k in range(64):
        assert np.argmax(np.abs(F[k])) == 2419 - k
        assert F[k, 2419 - k] > 0.9999
    assert np.allclose(w, np.arange(2420, 2420 - 64, -1), rtol=1e-5)


def test_calc_filters_vs_dense_eigh(ctx, hashprint_golden):
    ex = HashprintExtractor(ctx)
    ex.cov_reset()
    ex.cov_add_spectrogram(hashprint_golden["spec0"])
    cov = ex.cov_get()
    f, w = ex.calc_filters(None, install=False)
    F = f.T.astype(np.float64)
    ew, ev = np.linalg.eigh(cov.astype(np.float64))
    ew, ev = ew[::-1], ev[:, ::-1]
    assert np.max(np.abs(w - ew[:64]) / ew[0]) <= 1e-4
    assert np.allclose(F @ F.T, np.eye(64), atol=1e-5)                     # orthonormal rows
    # subspace: the 64 filters span the dense solver's leading 64-dimensional eigenspace
    P = ev[:, :64]
    if (ew[63] - ew[64]) / ew[63] > 1e-3:
        assert np.linalg.norm(F - (F @ P) @ P.T) <= 1e-3 * np.sqrt(64)
    checked = 0
    for k in range(64):
        gap = min((ew[k - 1] - ew[k]) if k else np.inf, ew[k] - ew[k + 1]) / ew[k]
        if gap >= 1e-2:
            assert abs(F[k] @ ev[:, k]) >= 0.9999, (k, gap)
            checked += 1
        assert F[k, np.argmax(np.abs(F[k]))] > 0                            # sign convention
    assert checked >= 16
    # residual ||A v - w v|| small relative to the largest eigenvalue
    R = cov.astype(np.float64) @ F.T - F.T * w.astype(np.float64)
    assert np.max(np.linalg.norm(R, axis=0)) <= 1e-4 * ew[0]


def test_index_learns_filters_from_scratch_cpp(tmp_path):
    """README usage (README.md:12-33): index() with an empty cache learns the filters, search() finds the tracks."""
    from tests.test_cpp_api import _build_example, _write_wav
    from hpfw_b200 import synth
    sr = 22050
    exe = _build_example(str(tmp_path), rate=sr)
    os.makedirs(tmp_path / "original")
    os.makedirs(tmp_path / "slices")
    tracks = []
    for i in range(6):
        t = synth.synth_track(5000 + i, 30.0, sr)
        tracks.append(t)
        _write_wav(tmp_path / "original" / f"song{i}.wav", t, sr)
    for i in range(6):
        q, _ = synth.synth_query(tracks[i], 6000 + i, 6.0, sr)
        _write_wav(tmp_path / "slices" / f"q{i}_song{i}.wav", q, sr)
    p = subprocess.run([exe, "original", "slices"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    last = [ln for ln in p.stdout.splitlines() if ln.startswith("=> ")][-1].split()
    assert int(last[1]) == 0 and float(last[2]) == 1.0                      # 0 wrong, accuracy 1
    # the cache now holds what the reference would have written
    f = np.fromfile(tmp_path / "cache" / "filters.cereal", dtype=np.int32, count=2)
    c = np.fromfile(tmp_path / "cache" / "accum_cov.cereal", dtype=np.int32, count=2)
    assert tuple(f) == (64, 2420) and tuple(c) == (2420, 2420)
    F = np.fromfile(tmp_path / "cache" / "filters.cereal", dtype=np.float32, offset=8).reshape(2420, 64).T
    assert np.allclose(F @ F.T, np.eye(64), atol=1e-4)


def test_partial_accumulators_sum_to_the_whole(ctx, hashprint_golden):
    """Multi-GPU index (sharded.allreduce_covariance): per-shard accumulators, summed on the device and written back through
    hpfw_cov_get_device / hpfw_cov_set_device, equal the accumulator of all tracks; the filters learned from the sum equal
    the filters learned in one pass (same subspace: projector difference ~ 0)."""
    import ctypes as C
    import torch
    from hpfw_b200._lib import check
    from hpfw_b200.sharded import ShardedMemoryStorage
    ex = HashprintExtractor(ctx)
    g = hashprint_golden
    specs = [g["spec0"][:400], g["q_spec"], g["spec0"][300:900], g["spec0"][150:500]]
    ex.cov_reset()
    for s in specs:
        ex.cov_add_spectrogram(s)
    whole = ex.cov_get()
    parts = []
    for shard in (specs[:2], specs[2:]):
        ex.cov_reset()
        for s in shard:
            ex.cov_add_spectrogram(s)
        t = torch.empty(2420 * 2420, dtype=torch.float32, device="cuda")
        check(ctx._lib.hpfw_cov_get_device(ctx.handle, C.c_void_p(t.data_ptr()), None))
        torch.cuda.synchronize()
        parts.append(t)
    total = parts[0] + parts[1]
    check(ctx._lib.hpfw_cov_set_device(ctx.handle, C.c_void_p(total.data_ptr()), None))
    got = ex.cov_get()
    assert np.max(np.abs(got - whole)) <= 1e-5 * np.abs(whole).max()
    # world = 1: the all-reduce (hpfw_shard_allreduce_cov, in place on the accumulator) is the identity
    st = ShardedMemoryStorage(ctx, 0, 1)
    st.allreduce_covariance()
    torch.cuda.synchronize()
    st.close()
    assert np.array_equal(ex.cov_get(), got)


def test_covariance_of_a_long_track(ctx):
    """A 60,000-column spectrogram (12-minute class, the 32-point chirp-z plans' territory): the accumulated fp32
    covariance stays within 5e-5 of the float64 value (2e-5 observed; 2e-5 is the bound for 3-minute tracks) and is
    exactly symmetric."""
    rng = np.random.default_rng(0)
    cols = 60000
    spec = (rng.standard_normal((cols, 121)) * 10 - 40).astype(np.float32)
    spec += np.linspace(0, 5, cols, dtype=np.float32)[:, None]
    ex = HashprintExtractor(ctx)
    ex.cov_reset()
    ex.cov_add_spectrogram(spec)
    got = ex.cov_get()
    nf = cols - 19
    s1, s2 = np.zeros(2420), np.zeros((2420, 2420))
    for a in range(0, nf, 4096):
        b = min(nf, a + 4096)
        X = np.empty((b - a, 2420), dtype=np.float64)
        for band in range(121):
            for c in range(20):
                X[:, band * 20 + c] = spec[a + c:b + c, band]
        s1 += X.sum(0)
        s2 += X.T @ X
    ref = (s2 - np.outer(s1, s1) / nf) / (nf - 1)
    assert np.abs(got - ref).max() <= 5e-5 * np.abs(ref).max()
    assert np.array_equal(got, got.T)
