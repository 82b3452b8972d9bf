"""Probabilistic YIN on the GPU (SURVEY 8(f) "next" row 2).

Drop-in for the reference's ``librosa.pyin(y, fmin=60, fmax=500, sr=CONFIG['sr'], hop_length=CONFIG['hop'])``
(``spev_real_metrics.py:369`` in the cache loop, ``:311`` in the statistics pass): three launches for a whole
ragged batch -- YIN difference function, trough statistics -> observation probabilities, one Viterbi decode
per utterance -- instead of seconds of CPU per utterance.
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from .batch import HOP, Context, FlatBatch, make_batch, stream_ptr
from .spectral import _ret, _to_device


class PyinContext:
    """Owns one ``spev_pyin`` (threshold / Boltzmann / transition tables for one ``(device, sr, fmin, fmax)``)."""

    _cache: dict = {}
    _lock = threading.Lock()

    def __init__(self, device: int, sr: int, fmin: float, fmax: float, hop: int = HOP):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("spev_tts_b200 needs a CUDA (sm_100) device; there is no CPU path")
        h = C.c_void_p()
        try:                                     # librosa's own table, rounded the way scipy rounds it
            import scipy.stats
            beta = np.ascontiguousarray(np.diff(scipy.stats.beta.cdf(np.linspace(0, 1, 101), 2, 18)), dtype=np.float64)
            beta_ptr = beta.ctypes.data
        except ImportError:                      # built-in closed form (same to ~1e-16 absolute)
            beta_ptr = None
        _lib.check(self.lib.spev_pyin_create(C.byref(h), int(device), int(sr), int(hop), float(fmin), float(fmax), beta_ptr),
                   "spev_pyin_create")
        self.handle, self.device, self.sr, self.hop, self.fmin = h, int(device), int(sr), int(hop), float(fmin)
        v = [C.c_int() for _ in range(4)]
        _lib.check(self.lib.spev_pyin_info(h, *[C.byref(x) for x in v]), "spev_pyin_info")
        self.n_bins, self.min_period, self.max_period, self.n_lags = (x.value for x in v)

    @classmethod
    def get(cls, device, *, sr=22050, fmin=60.0, fmax=500.0, hop=HOP) -> "PyinContext":
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError(f"spev_tts_b200 runs on CUDA devices only (got {dev})")
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        key = (idx, int(sr), float(fmin), float(fmax), int(hop))
        with cls._lock:
            ctx = cls._cache.get(key)
            if ctx is None:
                ctx = cls._cache[key] = cls(idx, sr, fmin, fmax, hop)
        return ctx

    @property
    def freqs64(self) -> np.ndarray:
        """Bin frequencies computed the way librosa computes them (numpy ``fmin * 2 ** (arange / (12 * bps))``):
        ``std::pow`` in the library differs from numpy's power in the last bit on ~6 % of the bins."""
        return self.fmin * 2.0 ** (np.arange(self.n_bins) / 120.0)

    def host_tables(self):
        """-> (dense log-transition ``[2n, 2n]``, bin frequencies ``[n]``, beta threshold weights ``[100]``)."""
        S = 2 * self.n_bins
        lt, fr, bp = np.empty((S, S)), np.empty(self.n_bins), np.empty(100)
        _lib.check(self.lib.spev_pyin_host_tables(self.handle, lt.ctypes.data, fr.ctypes.data, bp.ctypes.data),
                   "spev_pyin_host_tables")
        return lt, fr, bp

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.spev_pyin_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


def cmnd_flat(samples: torch.Tensor, batch: FlatBatch, pctx: PyinContext) -> torch.Tensor:
    """Stage 1: ``[F, n_lags]`` cumulative-mean-normalised difference (lags ``min_period..max_period``)."""
    yin = torch.empty(batch.n_frames, pctx.n_lags, dtype=torch.float32, device=samples.device)
    _lib.check(pctx.lib.spev_pyin_cmnd(pctx.handle, batch.desc, samples.data_ptr(), yin.data_ptr(),
                                       stream_ptr(samples.device)), "spev_pyin_cmnd")
    return yin


def observe(yin: torch.Tensor, pctx: PyinContext):
    """Stage 2: -> (log observation probs ``[F, n_bins]``, log unvoiced prob ``[F]``, voiced_prob ``[F]``)."""
    if not (yin.is_cuda and yin.dtype == torch.float32 and yin.is_contiguous() and yin.shape[-1] == pctx.n_lags):
        raise ValueError("yin must be a contiguous float32 CUDA tensor [F, n_lags]")
    F = yin.shape[0]
    logobs = torch.empty(F, pctx.n_bins, dtype=torch.float32, device=yin.device)
    lunv = torch.empty(F, dtype=torch.float32, device=yin.device)
    vp = torch.empty(F, dtype=torch.float32, device=yin.device)
    with torch.cuda.device(yin.device):
        _lib.check(pctx.lib.spev_pyin_observe(pctx.handle, yin.data_ptr(), F, logobs.data_ptr(), lunv.data_ptr(),
                                              vp.data_ptr(), stream_ptr(yin.device)), "spev_pyin_observe")
    return logobs, lunv, vp


def decode(logobs: torch.Tensor, log_unvoiced: torch.Tensor, frame_off, pctx: PyinContext):
    """Stage 3: Viterbi per item -> (states int32 ``[F]``, f0 ``[F]`` (NaN = unvoiced), voiced_flag bool ``[F]``)."""
    dev = logobs.device
    F = logobs.shape[0]
    fo = frame_off.to(dev, torch.int64).contiguous() if isinstance(frame_off, torch.Tensor) else \
        torch.from_numpy(np.ascontiguousarray(frame_off, dtype=np.int64)).to(dev)
    states = torch.empty(F, dtype=torch.int32, device=dev)
    f0 = torch.empty(F, dtype=torch.float32, device=dev)
    flag = torch.empty(F, dtype=torch.uint8, device=dev)
    nbytes = pctx.lib.spev_pyin_decode_workspace_bytes(pctx.handle, F)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(pctx.lib.spev_pyin_decode(pctx.handle, logobs.data_ptr(), log_unvoiced.data_ptr(), fo.data_ptr(),
                                             fo.numel() - 1, F, states.data_ptr(), f0.data_ptr(), flag.data_ptr(),
                                             ws.data_ptr(), nbytes, stream_ptr(dev)), "spev_pyin_decode")
    return states, f0, flag.bool()


def pyin_flat(samples: torch.Tensor, n_samples: Sequence[int], *, sr=22050, fmin=60.0, fmax=500.0, hop_length=HOP,
              sample_off: Optional[np.ndarray] = None, batch: Optional[FlatBatch] = None, return_states=False):
    """Ragged batch -> (f0 ``[F]``, voiced_flag ``[F]``, voiced_prob ``[F]``, frame_off ``[U+1]``).
    ``hop_length`` 256 (item i has ``1 + len_i // 256`` frames) or a multiple of it (``1 + len_i // hop_length``
    frames: the curves are computed on the hop-256 grid and every k-th frame is decoded, with the wider
    transition band that hop implies)."""
    if not (samples.is_cuda and samples.dtype == torch.float32 and samples.is_contiguous()):
        raise ValueError("samples must be a contiguous float32 CUDA tensor")
    ctx = Context.get(samples.device, sr=sr)
    pctx = PyinContext.get(samples.device, sr=sr, fmin=fmin, fmax=fmax, hop=hop_length)
    if batch is None:
        batch = make_batch(ctx, n_samples=n_samples, sample_off=sample_off)
    frame_off = batch.frame_off
    with torch.cuda.device(samples.device):
        yin = cmnd_flat(samples, batch, pctx)
        if hop_length != HOP:
            k = hop_length // HOP
            ns = np.asarray(n_samples, dtype=np.int64).reshape(-1)
            frames = 1 + ns // hop_length
            keep = np.concatenate([batch.frame_off[i] + k * np.arange(frames[i]) for i in range(len(ns))]) \
                if len(ns) else np.zeros(0, np.int64)
            yin = yin[torch.from_numpy(keep).to(samples.device)].contiguous()
            frame_off = np.concatenate([[0], np.cumsum(frames)]).astype(np.int64)
        logobs, lunv, vp = observe(yin, pctx)
        del yin
        states, f0, flag = decode(logobs, lunv, frame_off, pctx)
    if return_states:
        return f0, flag, vp, frame_off, states
    return f0, flag, vp, frame_off


def pyin(y, *, fmin, fmax, sr=22050, frame_length=2048, win_length=None, hop_length=None, n_thresholds=100,
         beta_parameters=(2, 18), boltzmann_parameter=2, resolution=0.1, max_transition_rate=35.92, switch_prob=0.01,
         no_trough_prob=0.01, fill_na=np.nan, center=True, pad_mode="constant", device=None):
    """Drop-in for ``librosa.pyin`` in the reference's configuration -> ``(f0, voiced_flag, voiced_prob)`` with
    shape ``[..., T]``; unvoiced frames hold ``fill_na``.  numpy input -> numpy float64 / bool / float64 like
    librosa; CUDA tensor input -> float32 / bool / float32 tensors on the same device (no host round trip)."""
    hop_length = frame_length // 4 if hop_length is None else hop_length
    win_length = frame_length // 2 if win_length is None else win_length
    if (frame_length, win_length, n_thresholds, tuple(beta_parameters), boltzmann_parameter, resolution,
            max_transition_rate, switch_prob, no_trough_prob, center, pad_mode) != \
            (2048, 1024, 100, (2, 18), 2, 0.1, 35.92, 0.01, 0.01, True, "constant") or hop_length not in (HOP, 2 * HOP):
        raise NotImplementedError("spev_tts_b200.pyin implements the reference's calls only: frame_length=2048, "
                                  "hop_length=256 (:369) or 512 (:311) and librosa's default pYIN parameters")
    t, was_numpy = _to_device(y, device)
    lead, n = t.shape[:-1], t.shape[-1]
    b = int(np.prod(lead)) if lead else 1
    f0, flag, vp, _, states = pyin_flat(t.reshape(-1), [n] * b, sr=sr, fmin=fmin, fmax=fmax, hop_length=hop_length,
                                        return_states=True)
    T = 1 + n // hop_length
    if was_numpy:
        # numpy in -> float64 out like librosa: f0 is read from the float64 bin table, so it is bit-identical to
        # librosa's wherever the decoded state is
        pctx = PyinContext.get(t.device, sr=sr, fmin=fmin, fmax=fmax, hop=hop_length)
        st = states.cpu().numpy().astype(np.int64)
        voiced = st < pctx.n_bins
        f0_h = pctx.freqs64[st % pctx.n_bins]
        if fill_na is not None:
            f0_h[~voiced] = fill_na
        return (f0_h.reshape(*lead, T), voiced.reshape(*lead, T),
                vp.cpu().numpy().astype(np.float64).reshape(*lead, T))
    if fill_na is None:                              # librosa: best-guess f0 of the decoded bin on unvoiced frames too
        pctx = PyinContext.get(t.device, sr=sr, fmin=fmin, fmax=fmax, hop=hop_length)
        freqs = torch.from_numpy(pctx.host_tables()[1].astype(np.float32)).to(t.device)
        f0 = freqs[states.long() % pctx.n_bins]
    elif not (isinstance(fill_na, float) and np.isnan(fill_na)):
        f0 = torch.where(flag, f0, torch.full_like(f0, float(fill_na)))
    return f0.view(*lead, T), flag.view(*lead, T), vp.view(*lead, T)
