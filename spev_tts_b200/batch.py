"""Device context and flat (ragged) batch descriptors for the C ABI.

A *flat batch* concatenates items (utterances or spectrograms) and describes them with prefix
offsets + per-CTA tile tables, all built on the host with numpy and uploaded in ONE
host->device copy.  Framing follows librosa ``center=True``: an item of ``N`` samples has
``T = 1 + N // hop`` frames and an ISTFT output of ``(T-1)*hop`` samples
(reference call sites: ``spev_real_metrics.py:363`` and ``:730-733``).
"""
from __future__ import annotations

import ctypes as C
import threading
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib

HOP = 256
N_FFT = 1024


def plan_tiles(counts: np.ndarray, per_tile: int):
    """Split items with ``counts[i]`` units into tiles of ``per_tile`` units.
    Returns (tile_item int32[n], tile_start int32[n]).  numpy twin of ``spev_plan_tiles``."""
    counts = np.asarray(counts, dtype=np.int64)
    nt = (counts + per_tile - 1) // per_tile
    tile_item = np.repeat(np.arange(len(counts), dtype=np.int32), nt)
    first = np.cumsum(nt) - nt
    tile_start = (np.arange(int(nt.sum()), dtype=np.int64) - np.repeat(first, nt)) * per_tile
    return tile_item, tile_start.astype(np.int32)


class Context:
    """Owns one ``spev_ctx`` (Hann window, twiddles, mel basis + pseudo-inverse for one
    ``(device, sr, n_fft, hop, win, n_mels, fmin, fmax)`` key)."""

    _cache: dict = {}
    _cache_lock = threading.Lock()

    def __init__(self, device: int, sr: int, n_fft: int, hop: int, win: int, n_mels: int,
                 fmin: float, fmax: Optional[float]):
        lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("spev_tts_b200 needs a CUDA (sm_100) device; there is no CPU path")
        self.lib = lib
        self.device = int(device)
        self.sr, self.n_fft, self.hop, self.win, self.n_mels = sr, n_fft, hop, win, n_mels
        self.fmin = float(fmin)
        self.fmax = float(fmax) if fmax is not None else 0.5 * sr
        self.tile_frames = lib.spev_tile_frames()
        self.tile_chunks = lib.spev_tile_chunks()
        h = C.c_void_p()
        _lib.check(lib.spev_create(C.byref(h), self.device, sr, n_fft, hop, win, n_mels,
                                   self.fmin, self.fmax), "spev_create")
        self.handle = h

    @classmethod
    def get(cls, device, *, sr=22050, n_fft=1024, hop=256, win=None, n_mels=80, fmin=0.0,
            fmax=None) -> "Context":
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError(f"spev_tts_b200 runs on CUDA devices only (got {dev})")
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        win = n_fft if win is None else win
        key = (idx, int(sr), int(n_fft), int(hop), int(win), int(n_mels), float(fmin),
               None if fmax is None else float(fmax))
        with cls._cache_lock:
            ctx = cls._cache.get(key)
            if ctx is None:
                ctx = cls(idx, int(sr), int(n_fft), int(hop), int(win), int(n_mels), fmin, fmax)
                cls._cache[key] = ctx
        return ctx

    def mel_basis(self) -> np.ndarray:
        out = np.empty((self.n_mels, _lib.N_BINS), dtype=np.float32)
        _lib.check(self.lib.spev_get_mel_basis(self.handle, out.ctypes.data), "spev_get_mel_basis")
        return out

    def mel_pinv(self) -> np.ndarray:
        out = np.empty((_lib.N_BINS, self.n_mels), dtype=np.float32)
        _lib.check(self.lib.spev_get_mel_pinv(self.handle, out.ctypes.data), "spev_get_mel_pinv")
        return out

    def window(self) -> np.ndarray:
        out = np.empty(self.n_fft, dtype=np.float32)
        _lib.check(self.lib.spev_get_window(self.handle, out.ctypes.data), "spev_get_window")
        return out

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.spev_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


@dataclass
class FlatBatch:
    """Host + device form of ``struct spev_batch``."""
    n_items: int
    n_frames: int
    frames: np.ndarray          # int64 [n_items]  frames per item
    sample_off: Optional[np.ndarray]   # int64 [n_items+1] or None
    frame_off: np.ndarray       # int64 [n_items+1]
    n_ftiles: int
    n_ctiles: int
    table: torch.Tensor         # device int64 buffer holding all arrays
    desc: _lib.SpevBatch
    device: torch.device

    @property
    def n_out_samples(self) -> int:
        """total ISTFT output length: sum (T_i - 1) * hop"""
        return int((self.n_frames - self.n_items) * HOP)

    def out_sample_off(self) -> np.ndarray:
        return (self.frame_off - np.arange(self.n_items + 1)) * HOP


def make_batch(ctx: Context, *, n_samples: Optional[Sequence[int]] = None,
               n_frames: Optional[Sequence[int]] = None, sample_off: Optional[np.ndarray] = None,
               with_chunks: bool = False, pinned: bool = True) -> FlatBatch:
    """Build the descriptor for items given by sample counts (waveform inputs) or frame counts
    (spectrogram inputs).  ``sample_off`` overrides the packed offsets (e.g. aligned packing:
    the gap after an item must be zero-filled, then it reads exactly like centre padding)."""
    dev = torch.device("cuda", ctx.device)
    if n_samples is not None:
        ns = np.asarray(n_samples, dtype=np.int64).reshape(-1)
        frames = 1 + ns // HOP
        if sample_off is None:
            sample_off = np.concatenate([[0], np.cumsum(ns)]).astype(np.int64)
        else:
            sample_off = np.asarray(sample_off, dtype=np.int64)
            assert sample_off.shape == (len(ns) + 1,)
    else:
        frames = np.asarray(n_frames, dtype=np.int64).reshape(-1)
        sample_off = None
    n_items = int(len(frames))
    frame_off = np.concatenate([[0], np.cumsum(frames)]).astype(np.int64)
    ft_item, ft_t0 = plan_tiles(frames, ctx.tile_frames)
    if with_chunks:
        ct_item, ct_c0 = plan_tiles(np.maximum(frames - 1, 0), ctx.tile_chunks)
    else:
        ct_item = ct_c0 = np.zeros(0, dtype=np.int32)

    # one int64 staging buffer: [sample_off | frame_off | ftile_item,ftile_t0 | ctile_item,ctile_c0]
    def as64(a32):   # pack int32 pairs into int64 words (keeps 8-byte alignment of what follows)
        n = (len(a32) + 1) // 2 * 2
        buf = np.zeros(n, dtype=np.int32)
        buf[: len(a32)] = a32
        return buf.view(np.int64)

    parts = [sample_off if sample_off is not None else np.zeros(0, np.int64), frame_off,
             as64(ft_item), as64(ft_t0), as64(ct_item), as64(ct_c0)]
    sizes = [len(p) for p in parts]
    host = torch.from_numpy(np.concatenate(parts)) if sum(sizes) else torch.zeros(0, dtype=torch.int64)
    if pinned and host.numel():
        host = host.pin_memory()
    table = host.to(dev, non_blocking=True)
    base = table.data_ptr()
    offs = np.concatenate([[0], np.cumsum(sizes)]) * 8
    d = _lib.SpevBatch()
    d.n_items = n_items
    d.n_ftiles = len(ft_item)
    d.n_ctiles = len(ct_item)
    d.n_frames = int(frame_off[-1])
    d.sample_off = base + int(offs[0]) if sample_off is not None else None
    d.frame_off = base + int(offs[1])
    d.ftile_item = base + int(offs[2])
    d.ftile_t0 = base + int(offs[3])
    d.ctile_item = base + int(offs[4]) if len(ct_item) else None
    d.ctile_c0 = base + int(offs[5]) if len(ct_item) else None
    fb = FlatBatch(n_items=n_items, n_frames=int(frame_off[-1]), frames=frames,
                   sample_off=sample_off, frame_off=frame_off, n_ftiles=len(ft_item),
                   n_ctiles=len(ct_item), table=table, desc=d, device=dev)
    fb._host = host   # keep the pinned staging buffer alive until the async copy is consumed
    return fb
