"""Whole cache records on the GPU (SURVEY 8(f) rows 1-3 joined): the reference's
``RealMetricsDataset.__init__`` build (``spev_real_metrics.py:300-430``) for a corpus already in memory.

    stats = corpus_stats(waves)                                  # :305-325  (statistics pass)
    records, vocab = build_records(waves, phones, durs, stats)   # :328-417  (processing pass)
    write_reference_cache(cache_dir, records, stats, vocab)      # :419-430

Every per-frame curve (log-mel, pYIN f0 / voiced probability, RMS, spectral centroid) and every per-phoneme
pooled value is computed by the CUDA kernels of this package for the WHOLE corpus at once (a handful of
launches); the host only does what the reference does per utterance in pure Python -- the duration
re-scaling of ``:373-397`` and the vocabulary.  Text -> phonemes (espeak) and TextGrid parsing stay with
the caller: ``phones[i]`` / ``durs[i]`` are what the reference holds in ``phs`` / ``durs`` at ``:359``.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .batch import HOP, Context, make_batch, stream_ptr
from .cache import aligned_offsets
from .features import frame_features_flat, segment_pool
from .pitch import PyinContext, pyin_flat
from .spectral import logmel_flat

MIN_SAMPLES = 4000            # :309, :333


def uniform_durations(n_samples: int, n_phones: int) -> List[int]:
    """The fallback alignment of ``:357``: every phone gets ``int((len(y) / 256) / len(phs))`` frames."""
    return [int((n_samples / HOP) / n_phones)] * n_phones


def scale_durations(phs: Sequence[str], durs: Sequence[int], n_frames: int) -> Optional[Tuple[List[str], List[int]]]:
    """"Strict duration scaling" (``:373-397``): stretch the durations to ``n_frames`` in total, at least one
    frame per phone; a shortfall goes to the last phone, an excess is taken off the tail (dropping phones
    that reach zero).  None where the reference skips the utterance."""
    total = sum(durs)
    if total <= 0:
        return None
    ratio = n_frames / total
    out = [max(1, int(d * ratio)) for d in durs]
    names = list(phs)
    excess = sum(out) - n_frames
    if excess < 0:
        out[-1] -= excess
    while excess > 0 and out:
        if out[-1] > excess:
            out[-1] -= excess
            excess = 0
        else:
            excess -= out.pop()
            names.pop()
    if not out or sum(out) != n_frames:
        return None
    return names, out


def plan_records(n_samples: Sequence[int], phones: Sequence[Optional[Sequence[str]]],
                 durs: Sequence[Optional[Sequence[int]]]):
    """Host part of the processing pass: which utterances survive (``:333``, ``:359``, ``:378``, ``:396``), their
    re-scaled durations, and the vocabulary (``:328``, ``:360``, ``:428``).
    -> (kept indices, phones per kept item, durations per kept item, sorted vocab)."""
    vocab = {"<PAD>", "<UNK>", "<SIL>"}
    kept, k_phs, k_durs = [], [], []
    for i, (n, ph, du) in enumerate(zip(n_samples, phones, durs)):
        if n < MIN_SAMPLES or not ph:
            continue
        vocab.update(ph)                                      # :360 -- before the durations are validated
        fixed = scale_durations(ph, du, 1 + n // HOP)
        if fixed is None:
            continue
        kept.append(i)
        k_phs.append(fixed[0])
        k_durs.append(fixed[1])
    return kept, k_phs, k_durs, sorted(vocab)


def _upload(waves: Sequence[np.ndarray], device) -> Tuple[torch.Tensor, np.ndarray, np.ndarray]:
    lens = np.array([len(w) for w in waves], dtype=np.int64)
    starts = aligned_offsets(lens)                           # [U+1], last entry = buffer size
    # pinned staging from torch's caching host allocator (a repeat build of similar size pays no cudaHostAlloc); only
    # the <= 3 alignment samples after each item and the 4-sample tail are zeroed, not the whole buffer
    host = torch.empty(int(starts[-1]) + 4, dtype=torch.float32, pin_memory=True)
    hv = host.numpy()
    for w, s, e in zip(waves, starts[:-1], starts[1:]):
        n = len(w)
        hv[s: s + n] = w
        hv[s + n: e] = 0.0
    hv[starts[-1]:] = 0.0
    return host.to(device, non_blocking=True), lens, starts[:-1]


def corpus_stats(waves: Sequence[np.ndarray], *, sr: int = 22050, device=None) -> dict:
    """Statistics pass (``:305-325``) over ``waves`` (the caller draws the sample of <= 500 files the reference
    draws with ``random.sample``): mean / std (+1e-5) of log-f0 on voiced frames (pYIN at librosa's default hop
    512), of log(RMS + 1e-6) (hop 256) and of log(centroid + 1e-8) (librosa's default hop 512)."""
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    waves = [w for w in waves if len(w) >= MIN_SAMPLES]
    if not waves:
        raise ValueError("corpus_stats: no utterance of at least 4000 samples (the reference's np.mean of an empty list)")
    x, lens, starts = _upload(waves, dev)
    ctx = Context.get(dev, sr=sr)
    batch = make_batch(ctx, n_samples=lens, sample_off=starts)
    with torch.cuda.device(dev):
        # voiced log-f0 from the exact float64 bin table (f0 itself is only stored as float32)
        _, _, _, _, states = pyin_flat(x, lens, sr=sr, hop_length=2 * HOP, batch=batch, return_states=True)
        logf = torch.from_numpy(np.log(PyinContext.get(dev, sr=sr, hop=2 * HOP).freqs64 + 1e-8)).to(dev)
        st = states.long()
        p = logf[st[st < logf.numel()]]
        rms, cent, _ = frame_features_flat(x, lens, sr=sr, batch=batch)
        e = torch.log(rms + 1e-6).double()
        # centroid at librosa's default hop 512 = every second frame of each utterance
        keep = np.concatenate([batch.frame_off[i] + 2 * np.arange(1 + lens[i] // (2 * HOP)) for i in range(len(lens))])
        c = torch.log(cent[torch.from_numpy(keep).to(dev)].double() + 1e-8)

        def ms(t):
            return float(t.mean()), float(t.std(unbiased=False)) + 1e-5
        (pm, ps), (em, es), (cm, cs) = ms(p), ms(e), ms(c)
    return {"p_mean": pm, "p_std": ps, "e_mean": em, "e_std": es, "c_mean": cm, "c_std": cs}


def build_records(waves: Sequence[np.ndarray], phones: Sequence[Optional[Sequence[str]]],
                  durs: Sequence[Optional[Sequence[int]]], stats: dict, *, sr: int = 22050, device=None):
    """Processing pass (``:328-417``).  -> (records, vocab); ``records[k]['index']`` is the position of the
    utterance in ``waves`` (the reference numbers its files ``u_{index:05d}.pt``, gaps included).  The arrays of
    a record are views into corpus-wide host buffers (``write_reference_cache`` copies them out per file)."""
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    kept, k_phs, k_durs, vocab = plan_records([len(w) for w in waves], phones, durs)
    records: List[dict] = []
    if kept:
        x, lens, starts = _upload([waves[i] for i in kept], dev)
        ctx = Context.get(dev, sr=sr)
        batch = make_batch(ctx, n_samples=lens, sample_off=starts)
        fo = batch.frame_off
        po = np.concatenate([[0], np.cumsum([len(d) for d in k_durs])]).astype(np.int64)
        d_durs = torch.from_numpy(np.concatenate(k_durs).astype(np.int64)).to(dev)
        d_fo, d_po = torch.from_numpy(fo).to(dev), torch.from_numpy(po).to(dev)
        with torch.cuda.device(dev):
            mel, _ = logmel_flat(x, lens, sr=sr, batch=batch)
            rms, cent, _ = frame_features_flat(x, lens, sr=sr, batch=batch)
            _, _, vp, _, states = pyin_flat(x, lens, sr=sr, batch=batch, return_states=True)
            energy = segment_pool(rms, d_fo, d_durs, d_po, mu=stats["e_mean"], sigma=stats["e_std"], lo=-2.5, hi=2.5,
                                  log_eps=1e-6)
            bright = segment_pool(cent, d_fo, d_durs, d_po, mu=stats["c_mean"], sigma=stats["c_std"], lo=-2.5, hi=2.5,
                                  log_eps=1e-8)
            breath = segment_pool(vp, d_fo, d_durs, d_po, mu=1.0, sigma=-1.0, lo=0.0, hi=0.8)
            pitch = torch.empty_like(energy)
            rough = torch.empty_like(energy)
            pctx = PyinContext.get(dev, sr=sr)
            _lib.check(pctx.lib.spev_pitch_pool(pctx.handle, states.data_ptr(), d_fo.data_ptr(), d_durs.data_ptr(),
                                                d_po.data_ptr(), len(kept), float(stats["p_mean"]), float(stats["p_std"]),
                                                -2.5, 2.5, 1.5, pitch.data_ptr(), rough.data_ptr(), stream_ptr(dev)),
                       "spev_pitch_pool")
            mel_h = torch.empty(mel.shape, dtype=mel.dtype, pin_memory=True)
            mel_h.copy_(mel)                                  # pinned D2H; records below are views of this buffer
            curves = {k: v.cpu().numpy() for k, v in (("pitch", pitch), ("energy", energy), ("breath", breath),
                                                      ("rough", rough), ("bright", bright))}
        for k, i in enumerate(kept):
            rec = {"index": i, "phs": k_phs[k], "durs": k_durs[k], "mel": mel_h[fo[k]: fo[k + 1]]}
            rec.update({name: v[po[k]: po[k + 1]] for name, v in curves.items()})
            records.append(rec)
    return records, vocab


def build_cache_sharded(cache_dir: str, waves: Sequence[np.ndarray], phones, durs, stats: Optional[dict] = None, *,
                        sr: int = 22050, device=None, builder=None):
    """The whole cache build over one process per GPU (``torch.distributed`` initialised; works unsharded without
    it).  The path shards by utterance with no data-path collective: every rank builds and writes the records of
    its own shard (``cache.shard_utterances``: balanced by frames); the only exchange is an object all-gather of
    file names and vocabularies, after which rank 0 writes ``metadata.json``.
    ``stats=None``: every rank computes the statistics pass over the same utterances (deterministic, no exchange)
    -- pass the reference's <= 500-file sample result instead for large corpora.
    -> (files in wav order, stats, vocab) on every rank.  ``builder`` (tests): replaces ``build_records``."""
    import torch.distributed as dist
    from .cache import shard_utterances
    from .dataset import write_metadata, write_records
    on = dist.is_available() and dist.is_initialized()
    rank, world = (dist.get_rank(), dist.get_world_size()) if on else (0, 1)
    build = builder if builder is not None else build_records
    if stats is None:
        stats = corpus_stats(waves, sr=sr, device=device)
    lens = np.array([len(w) for w in waves], dtype=np.int64)
    mine = shard_utterances(lens, world)[rank]
    recs, vocab = build([waves[i] for i in mine], [phones[i] for i in mine], [durs[i] for i in mine], stats,
                        sr=sr, device=device)
    for r in recs:
        r["index"] = int(mine[r["index"]])                      # position in the whole corpus
    files = write_records(cache_dir, recs)
    local = ([(r["index"], f) for r, f in zip(recs, files)], list(vocab))
    gathered = [local]
    if on and world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, local)
    all_files = [f for _, f in sorted(p for part, _ in gathered for p in part)]
    all_vocab = sorted(set().union(*[set(v) for _, v in gathered]))
    if rank == 0:
        write_metadata(cache_dir, all_files, stats, all_vocab)
    if on and world > 1:
        dist.barrier()
    return all_files, stats, all_vocab
