"""Import the reference's own ``spev_real_metrics`` module  --  TEST INFRASTRUCTURE ONLY.

``/root/reference`` exists only in the build container (never on the GPU box), so this
helper is used by ``oracle/make_golden.py`` to *generate* fixtures and by CPU tests that
skip when the tree is absent.  The third-party modules the reference imports at module
scope but which are not installed (librosa, soundfile, matplotlib, phonemizer, textgrid)
are stubbed in ``sys.modules``; none of them is touched by the classes we use
(``LengthRegulator`` ``spev_real_metrics.py:122-146``, ``RealMetricsFastSpeech2``
``:148-277``).
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SPEV_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "spev_real_metrics.py"))


def _stub(name: str) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__["__stub__"] = True
    return m


def load():
    """Return the reference module (imported once, bytecode writing disabled because
    the tree is read-only)."""
    if "spev_real_metrics" in sys.modules:
        return sys.modules["spev_real_metrics"]
    if not available():
        raise FileNotFoundError(REFERENCE_ROOT)
    sys.dont_write_bytecode = True
    for name in ("librosa", "librosa.feature", "librosa.feature.inverse", "soundfile",
                 "matplotlib", "matplotlib.pyplot", "phonemizer", "textgrid"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = _stub(name)
    if getattr(sys.modules["phonemizer"], "__stub__", False):
        sys.modules["phonemizer"].phonemize = lambda *a, **k: ""
    if getattr(sys.modules["matplotlib"], "__stub__", False):
        sys.modules["matplotlib"].use = lambda *a, **k: None
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import io
        import contextlib
        with contextlib.redirect_stdout(io.StringIO()):
            mod = importlib.import_module("spev_real_metrics")
    finally:
        sys.path.remove(REFERENCE_ROOT)
    return mod
