"""GPU parity of the TMA + tcgen05 (3xTF32) GEMM engine: mel projection (K2) and the
pseudo-inverse projection (K3') against the oracle and against the FFMA kernels."""
import numpy as np
import pytest
import torch

from oracle import librosa_restated as lr
from tests import synth

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


@pytest.mark.parametrize("n", [132300, 256 * 127, 256 * 128 + 17, 300, 256 * 1000])
def test_mel_project_tc_vs_oracle_and_fused(cuda, n):
    import spev_tts_b200 as sp
    y = synth.speechy(seed=2, n=n) if n > 1000 else synth.white(seed=2, n=n)
    yd = torch.from_numpy(y).to(cuda)
    power, fb = sp.stft_power_flat(yd, [n])
    assert power.shape == (1 + n // 256, 520)
    got = sp.mel_project(power).cpu().numpy()
    ref = lr.reference_logmel(y)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 1e-4, np.abs(got - ref).max()
    fused = sp.logmel(yd).cpu().numpy()
    assert np.abs(got - fused).max() <= 2e-5
    # raw mel power (mode 0)
    mp = sp.mel_project(power, log=False).cpu().numpy()
    mref = lr.melspectrogram(y=y, sr=22050, n_fft=1024, hop_length=256, n_mels=80).T
    assert np.abs(mp - mref).max() <= 2e-6 * max(1.0, np.abs(mref).max()) + 1e-4 * 0


def test_power_spectrum_pad_and_values(cuda):
    import spev_tts_b200 as sp
    y = synth.white(seed=3, n=256 * 70 + 5)
    power, _ = sp.stft_power_flat(torch.from_numpy(y).to(cuda), [len(y)])
    p = power.cpu().numpy()
    ref = (np.abs(lr.stft(y, n_fft=1024, hop_length=256)) ** 2).T
    assert np.all(p[:, 513:] == 0)
    assert np.abs(p[:, :513] - ref).max() <= 1e-5 * ref.max()


@pytest.mark.parametrize("F", [1, 63, 128, 129, 800, 2049])
def test_mel_to_mag_tc_vs_ffma_and_oracle(cuda, F):
    import spev_tts_b200 as sp
    rng = np.random.default_rng(F)
    lm = np.clip(-4 + 2 * rng.standard_normal((80, F)), -10, 2).astype(np.float32)
    ctx = sp.Context.get(cuda, sr=22050, n_mels=80, fmin=0.0, fmax=8000.0)
    try:
        S_tc = sp.mel_to_stft(np.exp(lm), sr=22050, n_fft=1024, fmin=0, fmax=8000, nnls="pinv")
        ctx.set_tensor_core(False)
        S_ff = sp.mel_to_stft(np.exp(lm), sr=22050, n_fft=1024, fmin=0, fmax=8000, nnls="pinv")
    finally:
        ctx.set_tensor_core(True)
    S_ref = lr.mel_to_stft(np.exp(lm), sr=22050, n_fft=1024, fmin=0, fmax=8000, lbfgs=False)
    assert S_tc.shape == S_ref.shape == (513, F)
    assert rel_l2(S_tc, S_ref) <= 1e-5, rel_l2(S_tc, S_ref)
    assert rel_l2(S_ff, S_ref) <= 1e-5
    assert rel_l2(S_tc, S_ff) <= 5e-6
    # is_log fused exp (Vocoder path) == exp on the host
    y1 = sp.mel_to_audio(lm, sr=22050, n_fft=1024, hop_length=256, fmin=0, fmax=8000, n_iter=0, is_log=True,
                         init_phase=np.zeros((513, F), np.float32))
    y2 = sp.mel_to_audio(np.exp(lm), sr=22050, n_fft=1024, hop_length=256, fmin=0, fmax=8000, n_iter=0,
                         init_phase=np.zeros((513, F), np.float32))
    if F > 1:
        assert rel_l2(y1, y2) <= 1e-5
