"""``install()``: monkey-patch the reference's call sites so that its scripts run unchanged on
the B200 path.

Patches, when the modules are importable:
  * ``librosa.feature.melspectrogram``            (used at ``spev_real_metrics.py:363``)
  * ``librosa.feature.inverse.mel_to_audio``      (used at ``:730-733``)
  * ``librosa.feature.rms`` / ``librosa.feature.spectral_centroid``   (``:370-371``, stats pass ``:314-316``)
  * ``librosa.pyin``                              (``:369``, stats pass ``:311``)
  * ``spev_real_metrics.LengthRegulator``         (``:122-146``; instantiated at ``:160``)
Calls with parameters outside the implemented configuration (n_fft != 1024, ...) are passed
through to the original function.
"""
from __future__ import annotations

import sys

from . import features, length_regulator, pitch, spectral

_installed = {}


def _passthrough(ours, theirs):
    def wrapper(*args, **kwargs):
        try:
            return ours(*args, **kwargs)
        except NotImplementedError:
            return theirs(*args, **kwargs)
    wrapper.__wrapped__ = theirs
    return wrapper


def install(verbose: bool = False) -> dict:
    done = {}
    try:
        import librosa
        import librosa.feature
        import librosa.feature.inverse
        _installed["melspectrogram"] = librosa.feature.melspectrogram
        _installed["mel_to_audio"] = librosa.feature.inverse.mel_to_audio
        librosa.feature.melspectrogram = _passthrough(spectral.melspectrogram, _installed["melspectrogram"])
        librosa.feature.inverse.mel_to_audio = _passthrough(spectral.mel_to_audio, _installed["mel_to_audio"])
        _installed["rms"] = librosa.feature.rms
        _installed["spectral_centroid"] = librosa.feature.spectral_centroid
        librosa.feature.rms = _passthrough(features.rms, _installed["rms"])
        librosa.feature.spectral_centroid = _passthrough(features.spectral_centroid, _installed["spectral_centroid"])
        if hasattr(librosa, "pyin"):
            _installed["pyin"] = librosa.pyin
            librosa.pyin = _passthrough(pitch.pyin, _installed["pyin"])
        done["librosa"] = True
    except (ImportError, AttributeError):   # absent, or a partial stub without the functions
        done["librosa"] = False
    mod = sys.modules.get("spev_real_metrics")
    if mod is not None and hasattr(mod, "LengthRegulator"):
        _installed["LengthRegulator"] = mod.LengthRegulator
        mod.LengthRegulator = length_regulator.LengthRegulator
        done["spev_real_metrics.LengthRegulator"] = True
    if verbose:
        print("spev_tts_b200.install:", done)
    return done


def uninstall() -> None:
    """Restore everything ``install()`` replaced."""
    try:
        import librosa
        for name in ("melspectrogram", "rms", "spectral_centroid"):
            if name in _installed:
                setattr(librosa.feature, name, _installed.pop(name))
        if "pyin" in _installed:
            librosa.pyin = _installed.pop("pyin")
        if "mel_to_audio" in _installed:
            librosa.feature.inverse.mel_to_audio = _installed.pop("mel_to_audio")
    except (ImportError, AttributeError):
        pass
    mod = sys.modules.get("spev_real_metrics")
    if mod is not None and "LengthRegulator" in _installed:
        mod.LengthRegulator = _installed.pop("LengthRegulator")
    _installed.clear()


def patch_model(model) -> int:
    """Swap the LengthRegulator instances of an already-built reference model
    (``RealMetricsFastSpeech2.length_regulator``, ``spev_real_metrics.py:160``)."""
    n = 0
    for name, child in list(model.named_children()):
        if type(child).__name__ == "LengthRegulator":
            setattr(model, name, length_regulator.LengthRegulator())
            n += 1
        else:
            n += patch_model(child)
    return n
