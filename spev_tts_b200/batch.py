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


TILE_DTYPE = np.dtype([("src0", "<i8"), ("lo", "<i8"), ("hi", "<i8"), ("row0", "<i8"),
                       ("n", "<i4"), ("t0", "<i4"), ("T", "<i4"), ("item", "<i4")])   # struct spev_tile
assert TILE_DTYPE.itemsize == 48


def _split(counts: np.ndarray, per_tile: int):
    """(tile_item, tile_start) for items with ``counts[i]`` units cut into ``per_tile``-unit tiles."""
    counts = np.asarray(counts, dtype=np.int64)
    nt = (counts + per_tile - 1) // per_tile
    tile_item = np.repeat(np.arange(len(counts), dtype=np.int64), nt)
    first = np.cumsum(nt) - nt
    tile_start = (np.arange(int(nt.sum()), dtype=np.int64) - np.repeat(first, nt)) * per_tile
    return tile_item, tile_start


def plan_frame_tiles(frames, sample_lo=None, n_samples=None, tile_frames: int = 32) -> np.ndarray:
    """numpy twin of ``spev_plan_frame_tiles`` (vectorised): one ``spev_tile`` per <=32-frame tile."""
    frames = np.asarray(frames, dtype=np.int64)
    fo = np.concatenate([[0], np.cumsum(frames)])[:-1]
    if sample_lo is None:
        lo = HOP * (fo - np.arange(len(frames)))
        hi = lo + (frames - 1) * HOP
    else:
        lo = np.asarray(sample_lo, dtype=np.int64)
        hi = lo + np.asarray(n_samples, dtype=np.int64)
    item, t0 = _split(frames, tile_frames)
    out = np.zeros(len(item), dtype=TILE_DTYPE)
    out["src0"] = lo[item] + HOP * t0 - N_FFT // 2
    out["lo"], out["hi"] = lo[item], hi[item]
    out["row0"] = fo[item] + t0
    out["n"] = np.minimum(tile_frames, frames[item] - t0)
    out["t0"], out["T"], out["item"] = t0, frames[item], item
    return out


def plan_chunk_tiles(frames, tile_chunks: int = 29) -> np.ndarray:
    """numpy twin of ``spev_plan_chunk_tiles``: one ``spev_tile`` per <=29-chunk ISTFT tile."""
    frames = np.asarray(frames, dtype=np.int64)
    fo = np.concatenate([[0], np.cumsum(frames)])[:-1]
    nc = np.maximum(frames - 1, 0)
    item, c0 = _split(nc, tile_chunks)
    out = np.zeros(len(item), dtype=TILE_DTYPE)
    out["src0"] = HOP * (fo[item] - item) + HOP * c0
    out["row0"] = fo[item] + c0 - 1
    out["n"] = np.minimum(tile_chunks, nc[item] - c0)
    out["t0"], out["T"], out["item"] = c0, frames[item], item
    # full tiles first, partial tiles last (same order as the C planner): tile j runs on CTA j mod grid, so the
    # cheap tiles fall into the last, incomplete round
    return out[np.argsort(out["n"] < tile_chunks, kind="stable")]


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

    def uniform_batch(self, n_items: int, n_frames: int, with_chunks: bool = True) -> "FlatBatch":
        """Descriptor of ``n_items`` equal-length spectrogram items, cached: the vocoder is called over and over with
        the same shapes, and planning + pinned staging + upload cost more than the kernels of a short utterance."""
        key = (int(n_items), int(n_frames), bool(with_chunks))
        cache = self.__dict__.setdefault("_uniform", {})
        fb = cache.get(key)
        if fb is None:
            if len(cache) >= 64:
                cache.pop(next(iter(cache)))
            fb = cache[key] = make_batch(self, n_frames=[n_frames] * n_items, with_chunks=with_chunks)
        return fb

    def side_stream(self) -> "torch.cuda.Stream":
        """A per-context helper stream (small read-backs that must not queue behind long kernel chains)."""
        st = self.__dict__.get("_side")
        if st is None:
            st = self.__dict__["_side"] = torch.cuda.Stream(torch.device("cuda", self.device))
        return st

    def set_tensor_core(self, enable: bool) -> None:
        """A/B switch: False forces the FFMA kernel for the mel->magnitude projection."""
        _lib.check(self.lib.spev_set_tensor_core(self.handle, 1 if enable else 0), "spev_set_tensor_core")

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
    sample_off: Optional[np.ndarray]   # int64 [n_items] absolute item starts (waveform batches)
    frame_off: np.ndarray       # int64 [n_items+1]
    n_ftiles: int
    n_ctiles: int
    table: torch.Tensor         # device uint8 buffer holding frame_off + tile tables
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
    (spectrogram inputs; sample tiles then describe the implicit ISTFT-output layout).
    ``sample_off[i]`` is the absolute start of item i in the flat sample buffer (default: items
    packed back to back).  Starts that are multiples of 4 samples take the 16-byte cp.async path;
    whatever lies between items is never read as signal (explicit [lo, hi) bounds per item)."""
    dev = torch.device("cuda", ctx.device)
    if n_samples is not None:
        ns = np.asarray(n_samples, dtype=np.int64).reshape(-1)
        frames = 1 + ns // HOP
        if sample_off is None:
            sample_off = np.concatenate([[0], np.cumsum(ns)])[:-1].astype(np.int64)
        else:
            sample_off = np.asarray(sample_off, dtype=np.int64).reshape(-1)[: len(ns)]
        ftiles = plan_frame_tiles(frames, sample_off, ns, ctx.tile_frames)
    else:
        frames = np.asarray(n_frames, dtype=np.int64).reshape(-1)
        sample_off = None
        ftiles = plan_frame_tiles(frames, None, None, ctx.tile_frames)
    n_items = int(len(frames))
    frame_off = np.concatenate([[0], np.cumsum(frames)]).astype(np.int64)
    ctiles = plan_chunk_tiles(frames, ctx.tile_chunks) if with_chunks else np.zeros(0, dtype=TILE_DTYPE)

    # one staging buffer: [ftiles | ctiles | frame_off]; tile tables first (16-byte aligned)
    parts = [ftiles.view(np.uint8).reshape(-1), ctiles.view(np.uint8).reshape(-1),
             frame_off.view(np.uint8).reshape(-1)]
    sizes = [p.size for p in parts]
    host = torch.from_numpy(np.concatenate(parts))
    if pinned:
        host = host.pin_memory()
    table = host.to(dev, non_blocking=True)
    base = table.data_ptr()
    assert base % 16 == 0
    d = _lib.SpevBatch()
    d.n_items = n_items
    d.n_ftiles = len(ftiles)
    d.n_ctiles = len(ctiles)
    d.n_frames = int(frame_off[-1])
    d.ftiles = base if len(ftiles) else None
    d.ctiles = base + sizes[0] if len(ctiles) else None
    d.frame_off = base + sizes[0] + sizes[1]
    fb = FlatBatch(n_items=n_items, n_frames=int(frame_off[-1]), frames=frames,
                   sample_off=sample_off, frame_off=frame_off, n_ftiles=len(ftiles),
                   n_ctiles=len(ctiles), table=table, desc=d, device=dev)
    fb._host = host   # keep the pinned staging buffer alive until the async copy is consumed
    return fb
