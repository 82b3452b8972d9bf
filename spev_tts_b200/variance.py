"""Duration rule and bucketize+embedding lookup of the FastSpeech 2 variance adaptor.

* ``duration_rule`` -- ``spev_real_metrics.py:215``.
* ``bucketize`` / ``bucketize_embed`` -- the canonical FastSpeech 2 pitch/energy embedding the
  north-star names (``torch.bucketize`` + ``nn.Embedding``).  The in-tree reference embeds the
  curves with ``Conv1d(1,256,3)`` instead (``:163-167``) and has no bucketize call site, so the
  oracle for these is ``torch.bucketize`` + ``F.embedding`` (SURVEY a-13).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from .batch import stream_ptr


def _cuda_f32(t: torch.Tensor, what: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"spev_tts_b200: {what} must be a CUDA tensor (no CPU path)")
    return t.to(torch.float32).contiguous()


def duration_rule(log_dur: torch.Tensor, d_control: float = 1.0) -> torch.Tensor:
    """``clamp((exp(log_dur) - 1) * d_control, 0, 500).round().long()``."""
    ld = _cuda_f32(log_dur, "log_dur")
    out = torch.empty(ld.shape, dtype=torch.int64, device=ld.device)
    with torch.cuda.device(ld.device):
        _lib.check(_lib.load().spev_duration_rule(ld.data_ptr(), ld.numel(), float(d_control),
                                                  out.data_ptr(), stream_ptr(ld.device)),
                   "spev_duration_rule")
    return out


def bucketize(values: torch.Tensor, boundaries: torch.Tensor, right: bool = False) -> torch.Tensor:
    """``torch.bucketize(values, boundaries, right=right)`` -> int64 indices."""
    v = _cuda_f32(values, "values")
    b = _cuda_f32(boundaries, "boundaries")
    idx = torch.empty(v.shape, dtype=torch.int64, device=v.device)
    with torch.cuda.device(v.device):
        _lib.check(_lib.load().spev_bucketize_embed(v.data_ptr(), v.numel(), b.data_ptr(), b.numel(),
                                                    1 if right else 0, None, 0, idx.data_ptr(), None, 0,
                                                    stream_ptr(v.device)), "spev_bucketize_embed")
    return idx


def bucketize_embed(values: torch.Tensor, boundaries: torch.Tensor, table: torch.Tensor,
                    right: bool = False, accumulate_into: Optional[torch.Tensor] = None,
                    return_index: bool = False):
    """``table[bucketize(values, boundaries)]`` -> ``[..., H]`` float32.  With
    ``accumulate_into`` (``[..., H]`` float32, contiguous) the rows are added in place
    (``x = x + pitch_embedding(...)`` of canonical FastSpeech 2)."""
    v = _cuda_f32(values, "values")
    b = _cuda_f32(boundaries, "boundaries")
    tb = _cuda_f32(table, "table")
    if tb.dim() != 2 or tb.shape[0] < b.numel() + 1:
        # bucket len(boundaries) is reachable (NaN, +inf, values above the last boundary): F.embedding would raise
        raise IndexError(f"table must have at least len(boundaries)+1 = {b.numel() + 1} rows (got {tuple(tb.shape)})")
    H = tb.shape[1]
    idx = torch.empty(v.shape, dtype=torch.int64, device=v.device) if return_index else None
    if accumulate_into is not None:
        out = accumulate_into
        if not (out.is_cuda and out.dtype == torch.float32 and out.is_contiguous()
                and tuple(out.shape) == (*v.shape, H)):
            raise ValueError("accumulate_into must be a contiguous float32 CUDA tensor [..., H]")
    else:
        out = torch.empty((*v.shape, H), dtype=torch.float32, device=v.device)
    with torch.cuda.device(v.device):
        _lib.check(_lib.load().spev_bucketize_embed(v.data_ptr(), v.numel(), b.data_ptr(), b.numel(),
                                                    1 if right else 0, tb.data_ptr(), H,
                                                    idx.data_ptr() if idx is not None else None,
                                                    out.data_ptr(), 1 if accumulate_into is not None else 0,
                                                    stream_ptr(v.device)), "spev_bucketize_embed")
    return (out, idx) if return_index else out
