"""Cache record format and GPU-side batching.

* ``write_reference_cache`` / ``read_reference_cache`` -- the reference's on-disk cache:
  ``cache_stable/u_%05d.pt`` holding ``{'phs','durs','mel','pitch','energy','breath','rough','bright'}``
  plus ``metadata.json`` with ``{'files','stats','vocab'}`` (``spev_real_metrics.py:419-430``; reloaded at
  ``:291-298``), so the reference ``Trainer`` runs unchanged on a cache produced here.
* ``ResidentCache`` -- the whole corpus on one GPU as flat ragged arrays; ``collate(indices)`` returns
  exactly what ``collate_fn([dataset[i] for i in indices])`` returns (``:433-462``: ids/durs/mel/curves
  zero-padded with ``pad_sequence(batch_first=True)``, ``lens``, ``log_durs``), built by ONE kernel launch
  (``spev_collate``) with no per-item ``torch.load`` and no host->device copy per batch.
"""
from __future__ import annotations

import ctypes as C
import json
import os
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .batch import stream_ptr

CURVES = ("pitch", "energy", "breath", "rough", "bright")


def write_records(cache_dir: str, records: Sequence[dict]) -> List[str]:
    """One ``u_%05d.pt`` per record (``spev_real_metrics.py:419-425``), numbered by ``record['index']`` when present
    (the reference numbers by wav index, gaps included) else by position."""
    os.makedirs(cache_dir, exist_ok=True)
    files = []
    for i, r in enumerate(records):
        path = os.path.join(cache_dir, f"u_{int(r.get('index', i)):05d}.pt")
        torch.save({"phs": list(r["phs"]), "durs": [int(d) for d in r["durs"]],
                    "mel": torch.as_tensor(r["mel"], dtype=torch.float32).cpu().clone(),
                    **{k: np.asarray(r[k]) for k in CURVES}}, path)
        files.append(path)
    return files


def write_metadata(cache_dir: str, files: Sequence[str], stats: dict, vocab: Sequence[str]) -> None:
    """``metadata.json`` = ``{'files', 'stats', 'vocab'}`` (``:428-430``; reloaded at ``:291-298``)."""
    with open(os.path.join(cache_dir, "metadata.json"), "w") as f:
        json.dump({"files": list(files), "stats": stats, "vocab": list(vocab)}, f)


def write_reference_cache(cache_dir: str, records: Sequence[dict], stats: dict, vocab: Sequence[str]) -> List[str]:
    """records[i] = {'phs': list[str], 'durs': list[int], 'mel': Tensor[T,80], 'pitch': ndarray[P], ...}."""
    files = write_records(cache_dir, records)
    write_metadata(cache_dir, files, stats, vocab)
    return files


def read_reference_cache(cache_dir: str):
    with open(os.path.join(cache_dir, "metadata.json")) as f:
        meta = json.load(f)
    records = [torch.load(p, weights_only=False) for p in meta["files"]]
    return records, meta["stats"], meta["vocab"]


class ResidentCache:
    """Flat, GPU-resident form of the cache: mel ``[F,80]``; ids/durs ``[P]`` int64; the five curves and
    ``log_durs`` ``[P]`` float32; ``frame_off`` / ``phone_off`` ``[U+1]``."""

    def __init__(self, records: Sequence[dict], vocab: Sequence[str], stats: Optional[dict] = None, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("spev_tts_b200.ResidentCache needs a CUDA (sm_100) device; there is no CPU path")
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.vocab, self.stats = list(vocab), stats
        ph_to_idx = {p: i for i, p in enumerate(self.vocab)}
        frames = np.array([int(r["mel"].shape[0]) for r in records], dtype=np.int64)
        phones = np.array([len(r["phs"]) for r in records], dtype=np.int64)
        self.n_frames_per_item, self.n_phones_per_item = frames, phones
        self.frame_off = np.concatenate([[0], np.cumsum(frames)]).astype(np.int64)
        self.phone_off = np.concatenate([[0], np.cumsum(phones)]).astype(np.int64)
        # host-side, one-time: exactly the conversions of __getitem__ (spev_real_metrics.py:433-447)
        ids = torch.cat([torch.LongTensor([ph_to_idx.get(p, 0) for p in r["phs"]]) for r in records]) if len(records) else torch.zeros(0, dtype=torch.int64)
        durs = torch.cat([torch.LongTensor(r["durs"]) for r in records]) if len(records) else torch.zeros(0, dtype=torch.int64)
        log_durs = torch.log(torch.clamp(durs.float(), min=1) + 1)
        self.n_mels = int(records[0]["mel"].shape[1]) if len(records) else 80
        dev = self.device
        self.mel = torch.cat([torch.as_tensor(r["mel"], dtype=torch.float32) for r in records]).contiguous().to(dev)
        self.ids, self.durs, self.log_durs = ids.to(dev), durs.to(dev), log_durs.to(dev)
        self.curves: Dict[str, torch.Tensor] = {
            k: torch.cat([torch.FloatTensor(np.asarray(r[k], dtype=np.float64)) for r in records]).to(dev) for k in CURVES}
        self.d_frame_off = torch.from_numpy(self.frame_off).to(dev)
        self.d_phone_off = torch.from_numpy(self.phone_off).to(dev)

    @classmethod
    def from_flat(cls, mel: torch.Tensor, frame_off, ids: torch.Tensor, durs: torch.Tensor, phone_off,
                  curves: Dict[str, torch.Tensor], vocab: Sequence[str] = (), stats: Optional[dict] = None):
        """Adopt device-resident flat arrays directly (e.g. the ``[F,80]`` output of ``logmel_flat`` /
        ``gather_shards``) without a round trip through per-utterance records."""
        self = cls.__new__(cls)
        self.device = mel.device
        self.vocab, self.stats = list(vocab), stats
        self.frame_off = np.asarray(frame_off, dtype=np.int64)
        self.phone_off = np.asarray(phone_off, dtype=np.int64)
        self.n_frames_per_item = np.diff(self.frame_off)
        self.n_phones_per_item = np.diff(self.phone_off)
        self.n_mels = int(mel.shape[1])
        self.mel, self.ids, self.durs = mel.contiguous(), ids.contiguous(), durs.contiguous()
        self.log_durs = torch.log(torch.clamp(durs.float(), min=1) + 1)
        self.curves = {k: curves[k].to(torch.float32).contiguous() for k in CURVES}
        self.d_frame_off = torch.from_numpy(self.frame_off).to(self.device)
        self.d_phone_off = torch.from_numpy(self.phone_off).to(self.device)
        return self

    @classmethod
    def load(cls, cache_dir: str, device=None) -> "ResidentCache":
        records, stats, vocab = read_reference_cache(cache_dir)
        return cls(records, vocab, stats, device)

    def __len__(self):
        return len(self.n_frames_per_item)

    def collate(self, indices: Sequence[int]) -> Optional[dict]:
        """== ``collate_fn([dataset[i] for i in indices])`` of the reference, on the GPU."""
        if len(indices) == 0:
            return None                                         # reference: `if not batch: return None`
        idx = np.asarray(indices, dtype=np.int64)
        B = len(idx)
        t_max = int(self.n_frames_per_item[idx].max())
        p_max = int(self.n_phones_per_item[idx].max())
        dev = self.device
        sel = torch.from_numpy(idx).to(dev)
        out = {
            "ids": torch.empty((B, p_max), dtype=torch.int64, device=dev),
            "lens": torch.from_numpy(self.n_phones_per_item[idx].copy()).to(dev),
            "durs": torch.empty((B, p_max), dtype=torch.int64, device=dev),
            "mel": torch.empty((B, t_max, self.n_mels), dtype=torch.float32, device=dev),
            "log_durs": torch.empty((B, p_max), dtype=torch.float32, device=dev),
        }
        for k in CURVES:
            out[k] = torch.empty((B, p_max), dtype=torch.float32, device=dev)
        srcs = [(self.ids, out["ids"], 8, 1), (self.durs, out["durs"], 8, 1),
                (self.mel, out["mel"], 4 * self.n_mels, 0), (self.log_durs, out["log_durs"], 4, 1)]
        srcs += [(self.curves[k], out[k], 4, 1) for k in CURVES]
        arr = (_lib.SpevPadArray * len(srcs))()
        for a, (s, d, rb, pp) in zip(arr, srcs):
            a.src, a.dst, a.row_bytes, a.per_phone = s.data_ptr(), d.data_ptr(), rb, pp
        with torch.cuda.device(dev):
            _lib.check(_lib.load().spev_collate(C.cast(arr, C.c_void_p), len(srcs), self.d_frame_off.data_ptr(),
                                                self.d_phone_off.data_ptr(), sel.data_ptr(), B, t_max, p_max,
                                                stream_ptr(dev)), "spev_collate")
        return out
