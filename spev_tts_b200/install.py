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


# librosa keyword arguments our shims do not take but whose DEFAULT value describes exactly what they compute;
# any other value (or any other unknown keyword) sends the call to the original function
_IGNORABLE_DEFAULTS = {"S": None, "htk": False, "norm": "slaney", "freq": None, "length": None}


def _same(a, b) -> bool:
    try:
        return a is b or bool(a == b)
    except Exception:
        return False


def _bind(ours, theirs, args, kwargs) -> dict:
    """Map a call written against librosa's signature onto our shim's keyword arguments.  Raises
    NotImplementedError when the call uses anything the shim does not implement exactly."""
    import inspect
    ours_params = inspect.signature(ours).parameters
    try:
        sig_t = inspect.signature(theirs)
        bound = sig_t.bind(*args, **kwargs)
    except (TypeError, ValueError):
        raise NotImplementedError("call does not bind to the original signature")
    call, extra_pos = {}, []
    for name, value in bound.arguments.items():
        kind = sig_t.parameters[name].kind
        if kind is inspect.Parameter.VAR_POSITIONAL:
            extra_pos.extend(value)
        elif kind is inspect.Parameter.VAR_KEYWORD:
            call.update(value)
        else:
            call[name] = value
    if extra_pos:                                   # librosa's first positional argument is the signal
        if len(extra_pos) > 1 or "y" in call or "M" in call:
            raise NotImplementedError("positional arguments beyond the signal")
        first = next(iter(ours_params))
        call[first] = extra_pos[0]
    out = {}
    for name, value in call.items():
        if name in ours_params:
            out[name] = value
            continue
        default = sig_t.parameters[name].default if name in sig_t.parameters else inspect.Parameter.empty
        if default is not inspect.Parameter.empty and _same(value, default):
            continue                                # spelled-out librosa default we implement implicitly
        if name in _IGNORABLE_DEFAULTS and _same(value, _IGNORABLE_DEFAULTS[name]):
            continue
        raise NotImplementedError(f"argument {name!r} is not implemented by the B200 shim")
    return out


def _passthrough(ours, theirs, post=None):
    """Wrapper installed over a librosa function.  Calls the shim when the call is one it implements exactly (bound
    through librosa's own signature, so positional ``y``, spelled-out defaults etc. all work); everything else --
    unsupported parameters (NotImplementedError), keywords the shim does not know, signature mismatches (TypeError)
    -- goes to the original function untouched.  A supported call on a host without an sm_100 GPU / without the
    built library still raises RuntimeError: this package has no CPU path and never falls back silently."""
    def wrapper(*args, **kwargs):
        try:
            call = _bind(ours, theirs, args, kwargs)
            out = ours(**call)
        except (NotImplementedError, TypeError):
            return theirs(*args, **kwargs)
        return post(out) if post is not None else out
    wrapper.__wrapped__ = theirs
    return wrapper


def _as_float64(out):
    """librosa.feature.spectral_centroid returns float64 (frequencies are float64); keep that for numpy callers."""
    import numpy as np
    return out.astype(np.float64) if isinstance(out, np.ndarray) else out


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
        librosa.feature.spectral_centroid = _passthrough(features.spectral_centroid, _installed["spectral_centroid"], _as_float64)
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
