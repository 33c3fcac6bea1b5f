"""oracle/nsgcq.py — TEST INFRASTRUCTURE ONLY.

CPU restatement (numpy/scipy pocketfft, float64) of stage a1/a2 of the reference path: the whole-track
non-stationary-Gabor constant-Q transform that `spectrum::CQT::spectrogram` (include/hpfw/spectrum/cqt.h:36-84)
obtains from essentia's `NSGConstantQ`, followed by hpfw's own abs / decimate-by-3 (cqt.h:73-81) and
`amplitude_to_db` (include/hpfw/spectrum/convert.h:7-25).

PARITY UNPINNED. The arithmetic of a1 lives in essentia, a third-party dependency that is neither vendored in
/root/reference nor version-pinned there (README.md:7 "essentia", CMakeLists.txt:36 `-lessentia`) and is not installed
in this image; the reference holds no test, fixture or golden vector for it. What follows restates the published
algorithm (Holighaus/Dörfler/Velasco/Grill NSGT; Schörkhuber's cqt toolbox, which essentia's NSGConstantQ ports) with
the parameters of the reference's call site (cqt.h:54-61):

    inputSize = len(audio), gamma = 0, binsPerOctave = 24, minimumWindow = 96, window = "hann",
    minFrequency = 130.81, maxFrequency = 4186.01; essentia defaults for the rest:
    sampleRate = 44100 (NOT the template SampleRate — it is never forwarded, cqt.h:54-61), rasterize = "full",
    phaseMode = "global", normalize = "none".

Conventions we had to choose (cannot be verified against essentia here):
  * window  w_j[k] = 0.5 + 0.5 cos(2 pi k / Lg_j) for k = -floor(Lg/2) .. ceil(Lg/2)-1 (periodic Hann centred on k = 0,
    the NSG toolbox's `winfuns('hann')` form) by default; `window="symmetric"` selects the generic symmetric Hann of
    size Lg instead (see band_window) — the convention is a switch here and in the product, not a constant;
  * inverse FFT normalised by 1/M (a global scale: cancels in power_to_db except at the 1e-10 floor);
  * the global-phase rotation is omitted (only |c| is used by hpfw);
  * round() is C's half-away-from-zero;
  * the spectrogram column left unwritten by cqt.h:73-81 when M % 3 == 0 is defined as amplitude 0.
The product's CUDA CQT is checked against THIS file, with the tolerance stated in tests/test_cqt_gpu.py.
"""
from __future__ import annotations

import math

import numpy as np
from scipy import fft as sfft

NSG_SR = 44100.0          # essentia NSGConstantQ default sampleRate (see header)
F_MIN = 130.81            # cqt.h:59  C3
F_MAX = 4186.01           # cqt.h:60  C8
BINS_PER_OCTAVE = 24      # cqt.h:20,56
MIN_WINDOW = 96           # cqt.h:19,57 (HopLength is passed as minimumWindow)
DOWNSAMPLE = 3            # cqt.h:22
N_BINS = 121              # cqt.h:21


def _cround(x):
    return np.floor(np.asarray(x, dtype=np.float64) + 0.5).astype(np.int64)


def nsg_design(n_samples: int):
    """Band layout for an n_samples-long input: (pos[121], Lg[121], M). All in FFT-bin units (fftres = sr/N)."""
    b = int(math.floor(BINS_PER_OCTAVE * math.log2(F_MAX / F_MIN)))
    # libm pow per band (math.pow), so the product's host-side design (std::pow) sees bit-identical doubles
    f = np.array([F_MIN * math.pow(2.0, j / BINS_PER_OCTAVE) for j in range(b + 1)], dtype=np.float64)
    q = math.pow(2.0, 1.0 / BINS_PER_OCTAVE) - math.pow(2.0, -1.0 / BINS_PER_OCTAVE)
    bw = q * f                                     # gamma = 0
    fftres = NSG_SR / float(n_samples)
    pos = np.floor(f / fftres).astype(np.int64)
    lg = np.maximum(_cround(bw / fftres), MIN_WINDOW)
    m = int(lg[-1])                                # rasterize "full": every band is rendered at the top band's length
    assert m == int(lg.max())
    return pos, lg, m


def spectrogram_cols(n_samples: int) -> int:
    """Number of spectrogram columns the reference allocates: M / 3 + 1 (cqt.h:73)."""
    _, _, m = nsg_design(n_samples)
    return m // DOWNSAMPLE + 1


def band_window(L: int, window: str = "periodic"):
    """(k, w): tap offsets k = -floor(L/2) .. ceil(L/2)-1 relative to the band centre and the window on them.
    "periodic": 0.5 + 0.5 cos(2 pi k / L) (NSG toolbox winfuns('hann'), the default here);
    "symmetric": 0.5 - 0.5 cos(2 pi n / (L - 1)), n = k + floor(L/2) (a generic size-L "hann" window rotated onto the
    band centre). Which of the two essentia builds for "window","hann" (cqt.h:58) is unverified; the product has the same
    switch (HPFW_CQT_WINDOW / hpfw_set_cqt_window)."""
    k = np.arange(-(L // 2), -(L // 2) + L, dtype=np.int64)
    if window == "periodic":
        return k, 0.5 + 0.5 * np.cos(2.0 * np.pi * k / L)
    if window == "symmetric":
        return k, 0.5 - 0.5 * np.cos(2.0 * np.pi * (k + L // 2) / (L - 1))
    raise ValueError(f"window must be 'periodic' or 'symmetric', not {window!r}")


def nsgcq_magnitude(audio: np.ndarray, window: str = "periodic") -> np.ndarray:
    """|c_j[3 i]| as float64 [cols, 121] (time-major, i.e. the memory order of the reference's column-major
    Eigen::Matrix<float,121,Dynamic>). Column M/3 is zero when M % 3 == 0 (see header)."""
    x = np.asarray(audio, dtype=np.float64)
    n = x.shape[0]
    pos, lg, m = nsg_design(n)
    spec = sfft.fft(x)
    cols = m // DOWNSAMPLE + 1
    written = -(-m // DOWNSAMPLE)
    out = np.zeros((cols, N_BINS), dtype=np.float64)
    for j in range(N_BINS):
        L = int(lg[j])
        k, w = band_window(L, window)
        buf = np.zeros(m, dtype=np.complex128)
        buf[k % m] = spec[(int(pos[j]) + k) % n] * w
        c = sfft.ifft(buf)
        out[:written, j] = np.abs(c[::DOWNSAMPLE])
    return out


def amplitude_to_db(mag: np.ndarray) -> np.ndarray:
    """convert.h:7-25 in float64: 10 log10(max(x^2,1e-10)) - 10 log10(max(1e-10, max x^2)), floored at max-80."""
    p = np.asarray(mag, dtype=np.float64) ** 2
    mx = max(1e-10, float(p.max())) if p.size else 1e-10
    l = 10.0 * np.log10(np.maximum(p, 1e-10)) - 10.0 * math.log10(mx)
    top = float(l.max()) if l.size else 0.0
    return np.maximum(l, top - 80.0)


def spectrogram(audio: np.ndarray, window: str = "periodic") -> np.ndarray:
    """Restated `CQT::spectrogram` on an already-decoded mono buffer: float32 [cols, 121] dB."""
    return amplitude_to_db(nsgcq_magnitude(audio, window)).astype(np.float32)
