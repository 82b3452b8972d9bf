"""The restated oracle against REAL librosa, whenever librosa is importable on the box running the tests.

librosa (requirements_conda.txt:42 -> 0.11.0) is where the reference's spectral arithmetic lives
(/root/reference/spev_real_metrics.py:363 melspectrogram, :369 pyin, :370-371 rms / spectral_centroid, :730-733
mel_to_audio).  It is not installable in the build container (no network, not in the wheelhouse), so these tests
SKIP there and the oracle stays pinned by independent implementations (tests/test_oracle.py, test_oracle_pyin.py);
on any box that has librosa they run and take precedence."""
import numpy as np
import pytest

librosa = pytest.importorskip("librosa")

from oracle import librosa_restated as lr   # noqa: E402
from oracle import pyin_restated as po       # noqa: E402
from tests import synth                      # noqa: E402

SR = 22050


@pytest.fixture(scope="module")
def y():
    return synth.speechy(seed=1, n=3 * SR)


def test_mel_filter_and_melspectrogram(y):
    for fmax in (None, 8000.0):
        a = librosa.filters.mel(sr=SR, n_fft=1024, n_mels=80, fmin=0.0, fmax=fmax)
        b = lr.mel_filter(sr=SR, n_fft=1024, n_mels=80, fmin=0.0, fmax=fmax)
        assert a.dtype == b.dtype and np.array_equal(a, b)
    a = librosa.feature.melspectrogram(y=y, sr=SR, n_fft=1024, hop_length=256, n_mels=80)
    b = lr.melspectrogram(y=y, sr=SR, n_fft=1024, hop_length=256, n_mels=80)
    assert a.shape == b.shape and a.dtype == b.dtype
    assert np.abs(a - b).max() <= 1e-6 * np.abs(a).max()
    ref = np.clip(np.log(np.clip(a, 1e-5, None)), -10, 2).T                       # :364-367, :421
    assert np.abs(lr.reference_logmel(y) - ref).max() <= 5e-6


def test_stft_istft(y):
    a = librosa.stft(y, n_fft=1024, hop_length=256)
    b = lr.stft(y, n_fft=1024, hop_length=256)
    assert a.dtype == b.dtype and np.abs(a - b).max() <= 1e-6 * np.abs(a).max()
    ia = librosa.istft(a, hop_length=256, n_fft=1024)
    ib = lr.istft(a, hop_length=256, n_fft=1024)
    assert ia.shape == ib.shape and np.abs(ia - ib).max() <= 1e-6


def test_mel_to_stft_nnls_and_griffinlim(y):
    M = librosa.feature.melspectrogram(y=y[: 256 * 63], sr=SR, n_fft=1024, hop_length=256, n_mels=80)
    for T in (64, 10, 1):                       # T = 10 / 1: L-BFGS-B iterates
        a = librosa.feature.inverse.mel_to_stft(M[:, :T], sr=SR, n_fft=1024, fmin=0, fmax=8000)
        b = lr.mel_to_stft(M[:, :T], sr=SR, n_fft=1024, fmin=0, fmax=8000, lbfgs=True)
        assert np.linalg.norm(a - b) <= 1e-4 * np.linalg.norm(a), T
    S = librosa.feature.inverse.mel_to_stft(M, sr=SR, n_fft=1024, fmin=0, fmax=8000)
    # librosa draws its phases from default_rng(random_state): re-create them to share the start point
    rng = np.random.default_rng(7)
    ph = (2 * np.pi * rng.random(size=S.shape)).astype(np.float32)
    a = librosa.griffinlim(S, n_iter=8, hop_length=256, n_fft=1024, random_state=7)
    b = lr.griffinlim(S, n_iter=8, hop_length=256, n_fft=1024, init_phase=ph)
    sc_a, sc_b = lr.spectral_convergence(a, S), lr.spectral_convergence(b, S)
    assert abs(sc_a - sc_b) <= 1e-3


def test_rms_centroid_pyin(y):
    assert np.abs(librosa.feature.rms(y=y, hop_length=256) - lr.rms(y=y, hop_length=256)).max() <= 1e-6
    a = librosa.feature.spectral_centroid(y=y, sr=SR, hop_length=256)
    b = lr.spectral_centroid(y=y, sr=SR, hop_length=256)
    assert np.abs(a - b).max() <= 1e-3 * np.abs(a).max()
    f0a, va, pa = librosa.pyin(y, fmin=60, fmax=500, sr=SR, hop_length=256)
    f0b, vb, pb = po.pyin(y, fmin=60, fmax=500, sr=SR, hop_length=256)
    assert np.mean(va == vb) >= 0.99
    both = va & vb
    assert np.abs(1200 * np.log2(f0a[both] / f0b[both])).max() <= 10.0 + 1e-6     # within one 10-cent bin
    assert np.abs(pa - pb).max() <= 1e-2
