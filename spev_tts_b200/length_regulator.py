"""LengthRegulator drop-in (``/root/reference/spev_real_metrics.py:122-146``).

Same call surface -- ``LengthRegulator()(x, durations) -> (Tensor[B,maxF,H], LongTensor[B])`` --
but the per-(b,t) Python loop with one ``.item()`` device sync per element (6*B*T syncs per
model forward, ``:226-236``) becomes two kernel launches and exactly ONE device->host read
(``max_len``, needed to size the dense output).
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import _lib
from .batch import stream_ptr

_DUR_DTYPES = {torch.int64: 0, torch.int32: 1, torch.float32: 2, torch.float64: 3,
               torch.float16: 4, torch.bfloat16: 5}
# post-expansion clamps of the five curves, spev_real_metrics.py:239-243
# (pitch, energy, breath, rough, bright)
VARIANCE_CLAMPS = ((-3.0, 3.0), (-3.0, 3.0), (0.0, 1.0), (0.0, 2.0), (-3.0, 3.0))


class LengthPlan:
    """Result of phase 1 (sanitise + cumsum): reusable for several expands of the same batch."""

    def __init__(self, cumsum: torch.Tensor, mel_lens: torch.Tensor, max_len: int, B: int, T: int):
        self.cumsum, self.mel_lens, self.max_len, self.B, self.T = cumsum, mel_lens, max_len, B, T


_pinned_scalars: dict = {}


def _pinned_scalar(dev: torch.device) -> torch.Tensor:
    """one pinned int64 per device, reused (cudaHostAlloc costs more than the two kernels)"""
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    t = _pinned_scalars.get(key)
    if t is None:
        t = _pinned_scalars[key] = torch.empty((1,), dtype=torch.int64).pin_memory()
    return t


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(f"spev_tts_b200.LengthRegulator: {what} must be a CUDA tensor "
                           "(there is no CPU path; the reference class handles CPU tensors)")


def plan(durations: torch.Tensor) -> LengthPlan:
    """Phase 1.  ``durations``: ``[B,T]`` int64 (training) or any float/int dtype."""
    _require_cuda(durations, "durations")
    if durations.dim() != 2:
        raise ValueError("durations must be [B, T]")
    if durations.dtype not in _DUR_DTYPES:
        durations = durations.to(torch.float64 if durations.is_floating_point() else torch.int64)
    durations = durations.contiguous()
    B, T = durations.shape
    if B == 0:
        raise ValueError("max() arg is an empty sequence")   # reference: max(mel_lens), :144
    lib = _lib.load()
    dev = durations.device
    cumsum = torch.empty((B, T), dtype=torch.int32, device=dev)
    mel_lens = torch.empty((B,), dtype=torch.int64, device=dev)
    max_dev = torch.empty((1,), dtype=torch.int64, device=dev)
    max_host = _pinned_scalar(dev)
    st = torch.cuda.current_stream(dev)
    with torch.cuda.device(dev):
        _lib.check(lib.spev_lr_plan(durations.data_ptr(), _DUR_DTYPES[durations.dtype], B, T,
                                    cumsum.data_ptr(), mel_lens.data_ptr(), max_dev.data_ptr(),
                                    max_host.data_ptr(), st.cuda_stream), "spev_lr_plan")
    st.synchronize()   # the single host sync of the forward
    return LengthPlan(cumsum, mel_lens, int(max_host.item()), B, T)


def expand(x: torch.Tensor, p: LengthPlan, feats: Optional[torch.Tensor] = None,
           clamps: Optional[Sequence[Tuple[float, float]]] = None):
    """Phase 2.  ``x``: ``[B,T,H]`` any dtype (rows copied verbatim) or None;
    ``feats``: ``[n_feat,B,T]`` float32 scalar curves expanded in the same launch.
    Returns ``out [B,maxF,H]`` (and ``feats_out [n_feat,B,maxF]`` when feats is given)."""
    import ctypes as C
    lib = _lib.load()
    out = None
    dev = p.cumsum.device
    x_ptr, row_bytes, out_ptr = None, 0, None
    if x is not None:
        _require_cuda(x, "x")
        if x.dim() != 3 or x.shape[0] != p.B or x.shape[1] != p.T:
            raise ValueError(f"x must be [B={p.B}, T={p.T}, H]")
        x = x.contiguous()
        out = torch.empty((p.B, p.max_len, x.shape[2]), dtype=x.dtype, device=dev)
        x_ptr, row_bytes, out_ptr = x.data_ptr(), x.shape[2] * x.element_size(), out.data_ptr()
    n_feat, f_ptr, fo_ptr, feats_out = 0, None, None, None
    lo = hi = None
    if feats is not None:
        _require_cuda(feats, "feats")
        feats = feats.to(torch.float32).contiguous()
        n_feat = feats.shape[0]
        feats_out = torch.empty((n_feat, p.B, p.max_len), dtype=torch.float32, device=dev)
        f_ptr, fo_ptr = feats.data_ptr(), feats_out.data_ptr()
        if clamps is not None:
            lo = (C.c_float * n_feat)(*[float(c[0]) for c in clamps])
            hi = (C.c_float * n_feat)(*[float(c[1]) for c in clamps])
    with torch.cuda.device(dev):
        _lib.check(lib.spev_lr_expand_fused(x_ptr, row_bytes, f_ptr, n_feat,
                                            C.cast(lo, C.c_void_p) if lo is not None else None,
                                            C.cast(hi, C.c_void_p) if hi is not None else None,
                                            p.cumsum.data_ptr(), p.B, p.T, out_ptr, fo_ptr, p.max_len,
                                            stream_ptr(dev)), "spev_lr_expand_fused")
    if feats is None:
        return out
    return out, feats_out


class LengthRegulator(nn.Module):
    """Drop-in for the reference class (``spev_real_metrics.py:122-146``)."""

    def forward(self, x: torch.Tensor, durations: torch.Tensor):
        p = plan(durations)
        return expand(x, p), p.mel_lens


def regulate_variances(x: torch.Tensor, durations: torch.Tensor, curves: Sequence[torch.Tensor],
                       clamps: Optional[Sequence[Tuple[float, float]]] = VARIANCE_CLAMPS):
    """The six LengthRegulator calls + five clamps of ``RealMetricsFastSpeech2.forward``
    (``spev_real_metrics.py:226-243``) in two launches.  ``curves``: five ``[B,T]`` tensors
    (pitch, energy, breath, rough, bright).  Returns ``(x_expanded [B,maxF,H], mel_len [B],
    curves_expanded: list of [B,1,maxF])`` -- the shapes the reference feeds its Conv1d
    embeddings (``:245-252``)."""
    p = plan(durations)
    feats = torch.stack([c.to(torch.float32) for c in curves])
    out, fo = expand(x, p, feats, clamps)
    return out, p.mel_lens, [fo[j].unsqueeze(1) for j in range(fo.shape[0])]


def variance_adaptor(x: torch.Tensor, durations: torch.Tensor, curves: Sequence[torch.Tensor],
                     embeddings: Sequence[nn.Module], clamps: Optional[Sequence[Tuple[float, float]]] = VARIANCE_CLAMPS,
                     return_curves: bool = False):
    """``spev_real_metrics.py:226-252`` in ONE kernel after the plan: expand ``x`` and the curves by
    ``durations``, clamp, apply each curve's ``nn.Conv1d(1, H, 3, padding=1)`` embedding and sum.
    ``embeddings``: the model's ``pitch_embedding, energy_embedding, breath_embedding, rough_embedding,
    bright_embedding`` (weights are read, not copied).  Returns ``(dec_input [B,maxF,H], mel_len [B])``
    (+ the expanded clamped curves ``[n,B,maxF]`` if ``return_curves``).  Forward only (inference)."""
    import ctypes as C
    _require_cuda(x, "x")
    p = plan(durations)
    x = x.to(torch.float32).contiguous()
    B, T, H = x.shape
    feats = torch.stack([c.to(torch.float32) for c in curves]).contiguous()
    n = feats.shape[0]
    w = torch.stack([e.weight.detach().reshape(H, 3) for e in embeddings]).to(torch.float32).contiguous()
    bias = torch.stack([e.bias.detach() for e in embeddings]).to(torch.float32).contiguous()
    out = torch.empty((B, p.max_len, H), dtype=torch.float32, device=x.device)
    fo = torch.empty((n, B, p.max_len), dtype=torch.float32, device=x.device) if return_curves else None
    lo = hi = None
    if clamps is not None:
        lo = (C.c_float * n)(*[float(c[0]) for c in clamps])
        hi = (C.c_float * n)(*[float(c[1]) for c in clamps])
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().spev_variance_fuse(
            x.data_ptr(), feats.data_ptr(), n, C.cast(lo, C.c_void_p) if lo is not None else None,
            C.cast(hi, C.c_void_p) if hi is not None else None, w.data_ptr(), bias.data_ptr(), p.cumsum.data_ptr(),
            B, T, H, out.data_ptr(), fo.data_ptr() if fo is not None else None, p.max_len, stream_ptr(x.device)),
            "spev_variance_fuse")
    return (out, p.mel_lens, fo) if return_curves else (out, p.mel_lens)


def mel_mask(mel_len: torch.Tensor, max_len: int) -> torch.Tensor:
    """``spev_real_metrics.py:259``: ``arange(maxF)[None,:] >= mel_len[:,None]``."""
    return torch.arange(max_len, device=mel_len.device)[None, :] >= mel_len[:, None]
