"""CPU tests of the zero-edit integration: install() patches the reference's
call sites and passes calls with unsupported parameters through to the original function."""
import sys
import types

import numpy as np
import pytest
import torch

from oracle import reference_import


@pytest.fixture()
def fake_librosa(monkeypatch):
    calls = []
    lib = types.ModuleType("librosa")
    feat = types.ModuleType("librosa.feature")
    inv = types.ModuleType("librosa.feature.inverse")
    for mod, names in ((feat, ("melspectrogram", "rms", "spectral_centroid")), (inv, ("mel_to_audio",))):
        for n in names:
            def f(*a, _n=n, **k):
                calls.append(_n)
                return ("original", _n)
            setattr(mod, n, f)
    lib.feature, feat.inverse = feat, inv

    def orig_pyin(*a, **k):
        calls.append("pyin")
        return ("original", "pyin")
    lib.pyin = orig_pyin
    monkeypatch.setitem(sys.modules, "librosa", lib)
    monkeypatch.setitem(sys.modules, "librosa.feature", feat)
    monkeypatch.setitem(sys.modules, "librosa.feature.inverse", inv)
    yield lib, calls
    import spev_tts_b200
    spev_tts_b200.uninstall()


def test_install_patches_and_passes_through(fake_librosa):
    import spev_tts_b200 as sp
    lib, calls = fake_librosa
    done = sp.install()
    assert done["librosa"] is True
    for fn in (lib.feature.melspectrogram, lib.feature.rms, lib.feature.spectral_centroid, lib.feature.inverse.mel_to_audio):
        assert hasattr(fn, "__wrapped__")
    y = np.zeros(4096, np.float32)
    # parameters outside the implemented configuration -> NotImplementedError inside the shim -> original
    assert lib.feature.melspectrogram(y=y, sr=22050, n_fft=2048, hop_length=512) == ("original", "melspectrogram")
    assert lib.feature.rms(y=y, frame_length=1024, hop_length=256) == ("original", "rms")
    assert lib.feature.spectral_centroid(y=y, sr=22050, n_fft=1024) == ("original", "spectral_centroid")
    assert lib.pyin(y, fmin=60, fmax=500, sr=22050, hop_length=128) == ("original", "pyin")
    assert calls == ["melspectrogram", "rms", "spectral_centroid", "pyin"]
    if not torch.cuda.is_available():
        # supported parameters reach the CUDA path, which fails loudly without a GPU (no silent fallback)
        with pytest.raises(RuntimeError):
            lib.feature.melspectrogram(y=y, sr=22050, n_fft=1024, hop_length=256, n_mels=80)
        with pytest.raises(RuntimeError):
            lib.pyin(y, fmin=60, fmax=500, sr=22050, hop_length=256)
        assert calls == ["melspectrogram", "rms", "spectral_centroid", "pyin"]


@pytest.mark.skipif(not reference_import.available(), reason="/root/reference not mounted")
def test_patch_model_swaps_length_regulator():
    import spev_tts_b200 as sp
    ref = reference_import.load()
    model = ref.RealMetricsFastSpeech2(vocab_size=30)
    assert type(model.length_regulator).__module__ == "spev_real_metrics"
    assert sp.patch_model(model) == 1
    assert isinstance(model.length_regulator, sp.LengthRegulator)
    orig = ref.LengthRegulator
    try:
        sp.install()
        assert ref.LengthRegulator is sp.LengthRegulator
    finally:
        sp.uninstall()
    assert ref.LengthRegulator is orig


def test_passthrough_binds_through_librosa_signatures():
    """Legitimate librosa calls the shims do not implement exactly reach the original function instead of raising
    TypeError: feature input S=, htk/norm/dtype variations, unknown keywords; spelled-out defaults and a positional
    signal are accepted."""
    import importlib
    inst = importlib.import_module("spev_tts_b200.install")
    from spev_tts_b200 import features, pitch, spectral
    seen = []

    def melspectrogram(*, y=None, sr=22050, S=None, n_fft=2048, hop_length=512, win_length=None, window="hann",
                       center=True, pad_mode="constant", power=2.0, **kwargs):
        seen.append("mel")
        return "orig"

    def rms(*, y=None, S=None, frame_length=2048, hop_length=512, center=True, pad_mode="constant", dtype=np.float32):
        seen.append("rms")
        return "orig"

    def pyin(y, *, fmin, fmax, sr=22050, frame_length=2048, hop_length=None, fill_na=np.nan):
        return "orig"
    y = np.zeros(4096, np.float32)
    w = inst._passthrough(spectral.melspectrogram, melspectrogram)
    assert w(S=np.ones((513, 4), np.float32), sr=22050) == "orig"                    # spectrogram input
    assert w(y=y, sr=22050, n_fft=1024, hop_length=256, n_mels=80, htk=True) == "orig"
    assert w(y=y, sr=22050, n_fft=1024, hop_length=256, n_mels=80, dtype=np.float64) == "orig"
    assert w(y=y, sr=22050, n_fft=1024, hop_length=256, n_mels=80, some_future_kwarg=1) == "orig"
    assert inst._passthrough(features.rms, rms)(S=np.ones((1025, 4), np.float32)) == "orig"
    assert seen == ["mel"] * 4 + ["rms"]
    b = inst._bind(spectral.melspectrogram, melspectrogram, (), dict(y=y, S=None, htk=False, norm="slaney", n_fft=1024))
    assert set(b) == {"y", "n_fft"}
    assert set(inst._bind(pitch.pyin, pyin, (y,), dict(fmin=60, fmax=500))) == {"y", "fmin", "fmax"}
    assert inst._as_float64(np.zeros(3, np.float32)).dtype == np.float64
