"""Frame-level energy / brightness features and per-phoneme pooling (SURVEY 8(f) "next" row 1).

Drop-ins for the two librosa calls of the reference's cache loop that share the STFT framing
(``spev_real_metrics.py:370-371``) and for the per-phone pooling lines (``:400-417``).  pYIN (f0,
voiced probability, ``:369``), the following "next" row, lives in ``pitch.py`` / ``csrc/pyin.cu``.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .batch import HOP, Context, FlatBatch, make_batch, stream_ptr
from .spectral import _ret, _to_device


def frame_features_flat(samples: torch.Tensor, n_samples: Sequence[int], *, sr=22050,
                        sample_off: Optional[np.ndarray] = None, batch: Optional[FlatBatch] = None):
    """Ragged batch -> (rms ``[F]``, centroid ``[F]``, batch).  One launch."""
    if not (samples.is_cuda and samples.dtype == torch.float32 and samples.is_contiguous()):
        raise ValueError("samples must be a contiguous float32 CUDA tensor")
    ctx = Context.get(samples.device, sr=sr)
    if batch is None:
        batch = make_batch(ctx, n_samples=n_samples, sample_off=sample_off)
    rms = torch.empty(batch.n_frames, dtype=torch.float32, device=samples.device)
    cent = torch.empty(batch.n_frames, dtype=torch.float32, device=samples.device)
    _lib.check(ctx.lib.spev_frame_features(ctx.handle, batch.desc, samples.data_ptr(), rms.data_ptr(),
                                           cent.data_ptr(), stream_ptr(samples.device)), "spev_frame_features")
    return rms, cent, batch


def _check(frame_length, hop_length, center, pad_mode):
    if frame_length != 2048 or hop_length not in (256, 512) or not center or pad_mode != "constant":
        raise NotImplementedError("spev_tts_b200 implements the reference's calls only: 2048-sample window, "
                                  "hop 256 (or 512 = every second frame), center=True, zero padding")


def rms(*, y, frame_length=2048, hop_length=512, center=True, pad_mode="constant", device=None):
    """Drop-in for ``librosa.feature.rms(y=...)`` -> ``[..., 1, T]`` float32."""
    _check(frame_length, hop_length, center, pad_mode)
    t, was_numpy = _to_device(y, device)
    lead, n = t.shape[:-1], t.shape[-1]
    b = int(np.prod(lead)) if lead else 1
    r, _, _ = frame_features_flat(t.reshape(-1), [n] * b)
    r = r.view(*lead, 1, 1 + n // HOP)[..., :: hop_length // HOP][..., : 1 + n // hop_length]
    return _ret(r, was_numpy)


def spectral_centroid(*, y, sr=22050, n_fft=2048, hop_length=512, center=True, pad_mode="constant", device=None):
    """Drop-in for ``librosa.feature.spectral_centroid(y=...)`` -> ``[..., 1, T]`` (float32 here;
    librosa returns float64)."""
    _check(n_fft, hop_length, center, pad_mode)
    t, was_numpy = _to_device(y, device)
    lead, n = t.shape[:-1], t.shape[-1]
    b = int(np.prod(lead)) if lead else 1
    _, c, _ = frame_features_flat(t.reshape(-1), [n] * b, sr=sr)
    c = c.view(*lead, 1, 1 + n // HOP)[..., :: hop_length // HOP][..., : 1 + n // hop_length]
    return _ret(c, was_numpy)


def segment_pool(curve: torch.Tensor, frame_off, durs: torch.Tensor, phone_off, *, mu=0.0, sigma=1.0,
                 lo=-float("inf"), hi=float("inf"), log_eps: Optional[float] = None) -> torch.Tensor:
    """``clip((mean(curve[seg]) - mu) / sigma, lo, hi)`` per phoneme (``spev_real_metrics.py:400-417``).
    ``curve``: ``[F]`` float32 CUDA; ``durs``: flat int64 durations; ``frame_off`` / ``phone_off``: ``[U+1]``
    prefix offsets (array-likes or CUDA int64 tensors)."""
    dev = curve.device

    def dev64(a):
        return a.to(dev, torch.int64).contiguous() if isinstance(a, torch.Tensor) else \
            torch.from_numpy(np.ascontiguousarray(a, dtype=np.int64)).to(dev)
    fo, po, du = dev64(frame_off), dev64(phone_off), dev64(durs)
    curve = curve.to(torch.float32).contiguous()
    out = torch.empty(du.numel(), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        if log_eps is None:
            _lib.check(_lib.load().spev_segment_pool(curve.data_ptr(), fo.data_ptr(), du.data_ptr(), po.data_ptr(),
                                                     fo.numel() - 1, float(mu), float(sigma), float(lo), float(hi),
                                                     out.data_ptr(), stream_ptr(dev)), "spev_segment_pool")
        else:                                    # pool log(curve + log_eps) (``:370`` / ``:397`` take the log per frame)
            _lib.check(_lib.load().spev_segment_pool_log(curve.data_ptr(), float(log_eps), fo.data_ptr(), du.data_ptr(),
                                                         po.data_ptr(), fo.numel() - 1, float(mu), float(sigma), float(lo),
                                                         float(hi), out.data_ptr(), stream_ptr(dev)), "spev_segment_pool_log")
    return out
