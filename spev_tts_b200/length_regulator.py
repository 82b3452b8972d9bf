"""LengthRegulator drop-in (``/root/reference/spev_real_metrics.py:122-146``).

Same call surface -- ``LengthRegulator()(x, durations) -> (Tensor[B,maxF,H], LongTensor[B])`` --
but the per-(b,t) Python loop with one ``.item()`` device sync per element (6*B*T syncs per
model forward, ``:226-236``) becomes two kernel launches and exactly ONE device->host read
(``max_len``, needed to size the dense output; none when the caller passes ``max_len=``).

**Differentiable like the reference.**  The reference expands with ``repeat`` / ``cat`` / ``F.pad`` /
``stack`` and its Trainer back-propagates the mel loss through the module (``:544-546``, ``loss.backward()``
``:574``).  ``expand``, ``regulate_variances`` and ``variance_adaptor`` are ``torch.autograd.Function`` s whose
backward runs ``spev_lr_expand_backward`` / ``spev_variance_fuse_backward``: every gradient element has one owner
thread that adds its segment's frames in ascending order (no atomics, bit-reproducible).  Durations are integer
indices and carry no gradient, in the reference as here.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import _lib
from .batch import stream_ptr

_DUR_DTYPES = {torch.int64: 0, torch.int32: 1, torch.float32: 2, torch.float64: 3,
               torch.float16: 4, torch.bfloat16: 5}
_GRAD_DTYPES = {torch.float32: 0, torch.float64: 1, torch.float16: 2, torch.bfloat16: 3}
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


def _clamp_arrays(clamps, n):
    if clamps is None:
        return None, None
    if len(clamps) != n:
        raise ValueError(f"{len(clamps)} clamp ranges for {n} curves")
    lo = (C.c_float * n)(*[float(c[0]) for c in clamps])
    hi = (C.c_float * n)(*[float(c[1]) for c in clamps])
    return C.cast(lo, C.c_void_p), C.cast(hi, C.c_void_p)


def plan(durations: torch.Tensor, max_len: Optional[int] = None) -> LengthPlan:
    """Phase 1.  ``durations``: ``[B,T]`` int64 (training) or any float/int dtype.

    ``max_len``: the padded output length when the caller already knows it -- in training it is
    ``b['mel'].size(1)`` (``spev_real_metrics.py:544-546``: target durations sum to the mel length).  The
    forward then has NO host synchronisation and can be captured in a CUDA graph.  Rows longer than
    ``max_len`` are truncated to it (``mel_lens`` still reports their true length)."""
    _require_cuda(durations, "durations")
    if durations.dim() != 2:
        raise ValueError("durations must be [B, T]")
    if durations.dtype not in _DUR_DTYPES:
        durations = durations.to(torch.float64 if durations.is_floating_point() else torch.int64)
    durations = durations.detach().contiguous()
    B, T = durations.shape
    if B == 0:
        raise ValueError("max() arg is an empty sequence")   # reference: max(mel_lens), :144
    lib = _lib.load()
    dev = durations.device
    cumsum = torch.empty((B, T), dtype=torch.int32, device=dev)
    mel_lens = torch.empty((B,), dtype=torch.int64, device=dev)
    max_dev = torch.empty((1,), dtype=torch.int64, device=dev)
    max_host = _pinned_scalar(dev) if max_len is None else None
    st = torch.cuda.current_stream(dev)
    with torch.cuda.device(dev):
        _lib.check(lib.spev_lr_plan(durations.data_ptr(), _DUR_DTYPES[durations.dtype], B, T,
                                    cumsum.data_ptr(), mel_lens.data_ptr(), max_dev.data_ptr(),
                                    max_host.data_ptr() if max_host is not None else None, st.cuda_stream),
                   "spev_lr_plan")
    if max_len is None:
        st.synchronize()   # the single host sync of the forward
        max_len = int(max_host.item())
    elif max_len < 1:
        raise ValueError("max_len must be >= 1")
    return LengthPlan(cumsum, mel_lens, int(max_len), B, T)


def _expand_raw(x, p: LengthPlan, feats, clamps):
    """One launch of spev_lr_expand_fused on detached, contiguous tensors."""
    lib = _lib.load()
    dev = p.cumsum.device
    out = feats_out = None
    x_ptr, row_bytes, out_ptr = None, 0, None
    if x is not None:
        out = torch.empty((p.B, p.max_len, x.shape[2]), dtype=x.dtype, device=dev)
        x_ptr, row_bytes, out_ptr = x.data_ptr(), x.shape[2] * x.element_size(), out.data_ptr()
    n_feat, f_ptr, fo_ptr = 0, None, None
    lo = hi = None
    if feats is not None:
        n_feat = feats.shape[0]
        feats_out = torch.empty((n_feat, p.B, p.max_len), dtype=torch.float32, device=dev)
        f_ptr, fo_ptr = feats.data_ptr(), feats_out.data_ptr()
        lo, hi = _clamp_arrays(clamps, n_feat)
    with torch.cuda.device(dev):
        _lib.check(lib.spev_lr_expand_fused(x_ptr, row_bytes, f_ptr, n_feat, lo, hi, p.cumsum.data_ptr(), p.B, p.T,
                                            out_ptr, fo_ptr, p.max_len, stream_ptr(dev)), "spev_lr_expand_fused")
    return out, feats_out


class _ExpandFn(torch.autograd.Function):
    """``out[b,f,:] = x[b,idx(b,f),:]`` (+ the clamped scalar curves); backward = per-segment sums."""

    @staticmethod
    def forward(ctx, x, feats, p: LengthPlan, clamps):
        ctx.plan, ctx.clamps = p, clamps
        ctx.x_meta = None if x is None else (x.dtype, x.shape[2])
        ctx.save_for_backward(feats if (feats is not None and clamps is not None) else None)
        ctx.n_feat = 0 if feats is None else feats.shape[0]
        out, fo = _expand_raw(x, p, feats, clamps)
        return out, fo

    @staticmethod
    def backward(ctx, g_out, g_fo):
        p: LengthPlan = ctx.plan
        (feats,) = ctx.saved_tensors
        dev = p.cumsum.device
        lib = _lib.load()
        gx = gf = None
        go_ptr = gx_ptr = None
        dt, H = 0, 0
        if ctx.x_meta is not None and g_out is not None and ctx.needs_input_grad[0]:
            dtype, H = ctx.x_meta
            if dtype not in _GRAD_DTYPES:
                raise RuntimeError(f"spev_tts_b200.LengthRegulator: no backward for x.dtype={dtype}")
            g_out = g_out.contiguous()
            gx = torch.empty((p.B, p.T, H), dtype=dtype, device=dev)
            go_ptr, gx_ptr, dt = g_out.data_ptr(), gx.data_ptr(), _GRAD_DTYPES[dtype]
        n_feat, gfo_ptr, gf_ptr, f_ptr = 0, None, None, None
        lo = hi = None
        if ctx.n_feat and g_fo is not None and ctx.needs_input_grad[1]:
            n_feat = ctx.n_feat
            g_fo = g_fo.to(torch.float32).contiguous()
            gf = torch.empty((n_feat, p.B, p.T), dtype=torch.float32, device=dev)
            gfo_ptr, gf_ptr = g_fo.data_ptr(), gf.data_ptr()
            if ctx.clamps is not None:
                lo, hi = _clamp_arrays(ctx.clamps, n_feat)
                f_ptr = feats.data_ptr()
        if gx is not None or gf is not None:
            with torch.cuda.device(dev):
                _lib.check(lib.spev_lr_expand_backward(go_ptr, dt, H, gfo_ptr, n_feat, f_ptr, lo, hi, p.cumsum.data_ptr(),
                                                       p.B, p.T, p.max_len, gx_ptr, gf_ptr, stream_ptr(dev)),
                           "spev_lr_expand_backward")
        return gx, gf, None, None


def expand(x: Optional[torch.Tensor], p: LengthPlan, feats: Optional[torch.Tensor] = None,
           clamps: Optional[Sequence[Tuple[float, float]]] = None):
    """Phase 2.  ``x``: ``[B,T,H]`` any dtype (rows copied verbatim) or None;
    ``feats``: ``[n_feat,B,T]`` float32 scalar curves expanded in the same launch.
    Returns ``out [B,maxF,H]`` (and ``feats_out [n_feat,B,maxF]`` when feats is given).
    Differentiable in ``x`` (float32/64/16/bfloat16) and ``feats``."""
    if x is not None:
        _require_cuda(x, "x")
        if x.dim() != 3 or x.shape[0] != p.B or x.shape[1] != p.T:
            raise ValueError(f"x must be [B={p.B}, T={p.T}, H]")
        x = x.contiguous()
    if feats is not None:
        _require_cuda(feats, "feats")
        feats = feats.to(torch.float32).contiguous()
        if feats.dim() != 3 or feats.shape[1] != p.B or feats.shape[2] != p.T:
            raise ValueError(f"feats must be [n_feat, B={p.B}, T={p.T}]")
    out, fo = _ExpandFn.apply(x, feats, p, tuple(clamps) if clamps is not None else None)
    if feats is None:
        return out
    return out, fo


class LengthRegulator(nn.Module):
    """Drop-in for the reference class (``spev_real_metrics.py:122-146``), forward and backward.
    ``max_len`` (optional, not in the reference signature): see ``plan``."""

    def forward(self, x: torch.Tensor, durations: torch.Tensor, max_len: Optional[int] = None):
        p = plan(durations, max_len)
        return expand(x, p), p.mel_lens


def regulate_variances(x: torch.Tensor, durations: torch.Tensor, curves: Sequence[torch.Tensor],
                       clamps: Optional[Sequence[Tuple[float, float]]] = VARIANCE_CLAMPS,
                       max_len: Optional[int] = None):
    """The six LengthRegulator calls + five clamps of ``RealMetricsFastSpeech2.forward``
    (``spev_real_metrics.py:226-243``) in two launches.  ``curves``: five ``[B,T]`` tensors
    (pitch, energy, breath, rough, bright).  Returns ``(x_expanded [B,maxF,H], mel_len [B],
    curves_expanded: list of [B,1,maxF])`` -- the shapes the reference feeds its Conv1d
    embeddings (``:245-252``).  Differentiable in ``x`` and the curves (clamp mask like ``torch.clamp``)."""
    p = plan(durations, max_len)
    feats = torch.stack([c.to(torch.float32) for c in curves])
    out, fo = expand(x, p, feats, clamps)
    return out, p.mel_lens, [fo[j].unsqueeze(1) for j in range(fo.shape[0])]


class _VarianceAdaptorFn(torch.autograd.Function):
    """spev_variance_fuse / spev_variance_fuse_backward."""

    @staticmethod
    def forward(ctx, x, feats, w, bias, p: LengthPlan, clamps, return_curves):
        lib = _lib.load()
        B, T, H = x.shape
        n = feats.shape[0]
        out = torch.empty((B, p.max_len, H), dtype=torch.float32, device=x.device)
        fo = torch.empty((n, B, p.max_len), dtype=torch.float32, device=x.device) if return_curves else None
        lo, hi = _clamp_arrays(clamps, n)
        with torch.cuda.device(x.device):
            _lib.check(lib.spev_variance_fuse(
                x.data_ptr(), feats.data_ptr(), n, lo, hi, w.data_ptr(), bias.data_ptr(), p.cumsum.data_ptr(),
                B, T, H, out.data_ptr(), fo.data_ptr() if fo is not None else None, p.max_len, stream_ptr(x.device)),
                "spev_variance_fuse")
        ctx.plan, ctx.clamps, ctx.shape = p, clamps, (B, T, H, n)
        ctx.save_for_backward(feats, w)
        if fo is not None:
            ctx.mark_non_differentiable(fo)     # a by-product for inspection; gradients flow through `out`
            return out, fo
        return out, None

    @staticmethod
    def backward(ctx, g_out, _g_fo):
        p: LengthPlan = ctx.plan
        feats, w = ctx.saved_tensors
        B, T, H, n = ctx.shape
        dev = g_out.device
        lib = _lib.load()
        g_out = g_out.to(torch.float32).contiguous()
        need_x, need_f, need_w, need_b = ctx.needs_input_grad[:4]
        gx = torch.empty((B, T, H), dtype=torch.float32, device=dev) if need_x else None
        gf = torch.empty((n, B, T), dtype=torch.float32, device=dev) if need_f else None
        gw = torch.empty((n, H, 3), dtype=torch.float32, device=dev) if need_w else None
        gb = torch.empty((n, H), dtype=torch.float32, device=dev) if need_b else None
        ws = torch.empty(lib.spev_variance_fuse_backward_workspace_bytes(n, B, H, p.max_len), dtype=torch.uint8, device=dev)
        lo, hi = _clamp_arrays(ctx.clamps, n)
        ptr = lambda t: t.data_ptr() if t is not None else None   # noqa: E731
        with torch.cuda.device(dev):
            _lib.check(lib.spev_variance_fuse_backward(
                g_out.data_ptr(), feats.data_ptr(), n, lo, hi, w.data_ptr(), p.cumsum.data_ptr(), B, T, H, p.max_len,
                ptr(gx), ptr(gf), ptr(gw), ptr(gb), ws.data_ptr(), ws.numel(), stream_ptr(dev)),
                "spev_variance_fuse_backward")
        return gx, gf, gw, gb, None, None, None


def variance_adaptor(x: torch.Tensor, durations: torch.Tensor, curves: Sequence[torch.Tensor],
                     embeddings: Sequence[nn.Module], clamps: Optional[Sequence[Tuple[float, float]]] = VARIANCE_CLAMPS,
                     return_curves: bool = False, max_len: Optional[int] = None):
    """``spev_real_metrics.py:226-252`` in ONE kernel after the plan: expand ``x`` and the curves by
    ``durations``, clamp, apply each curve's ``nn.Conv1d(1, H, 3, padding=1)`` embedding and sum.
    ``embeddings``: the model's ``pitch_embedding, energy_embedding, breath_embedding, rough_embedding,
    bright_embedding``.  Returns ``(dec_input [B,maxF,H], mel_len [B])`` (+ the expanded clamped curves
    ``[n,B,maxF]`` if ``return_curves``).  Differentiable in ``x``, the curves and the embeddings' weights and
    biases (one fused backward kernel + a fixed-order reduction)."""
    _require_cuda(x, "x")
    p = plan(durations, max_len)
    x = x.to(torch.float32).contiguous()
    H = x.shape[2]
    feats = torch.stack([c.to(torch.float32) for c in curves]).contiguous()
    w = torch.stack([e.weight.reshape(H, 3) for e in embeddings]).to(torch.float32).contiguous()
    bias = torch.stack([e.bias for e in embeddings]).to(torch.float32).contiguous()
    out, fo = _VarianceAdaptorFn.apply(x, feats, w, bias, p, tuple(clamps) if clamps is not None else None,
                                       bool(return_curves))
    return (out, p.mel_lens, fo) if return_curves else (out, p.mel_lens)


def mel_mask(mel_len: torch.Tensor, max_len: int) -> torch.Tensor:
    """``spev_real_metrics.py:259``: ``arange(maxF)[None,:] >= mel_len[:,None]``."""
    return torch.arange(max_len, device=mel_len.device)[None, :] >= mel_len[:, None]
