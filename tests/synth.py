"""Deterministic synthetic inputs shared by the tests, the golden generator and bench.py
(SURVEY.md section 8(d) "Synthetic inputs").  numpy only."""
from __future__ import annotations

import numpy as np

SR = 22050
HOP = 256
N_FFT = 1024


def white(seed: int = 0, n: int = 132300, sigma: float = 0.05) -> np.ndarray:
    """cfg1 case A: 6 s of white noise, sigma 0.05 (log-mel stays inside the clamps)."""
    return (sigma * np.random.default_rng(seed).standard_normal(n)).astype(np.float32)


def speechy(seed: int = 1, n: int = 132300, sr: int = SR) -> np.ndarray:
    """cfg1 case B: harmonic stack with vibrato + 3 Hz AM + 1e-4 noise floor.  Exercises
    both clamps of the reference's log compression (about a quarter of the bins sit on
    the -10 floor, peaks reach the +2 ceiling)."""
    r = np.random.default_rng(seed)
    t = np.arange(n) / sr
    f0 = 120 + 30 * np.sin(2 * np.pi * 0.7 * t + r.uniform(0, 6))
    ph = 2 * np.pi * np.cumsum(f0) / sr
    y = np.zeros(n)
    for k in range(1, 40):
        y += k ** -1.5 * np.sin(k * ph)
    y *= 0.3 * (0.6 + 0.4 * np.sin(2 * np.pi * 3 * t))
    y += 1e-4 * r.standard_normal(n)
    return y.astype(np.float32)


def utterance_lengths(seed: int, n_utts: int, lo_s: float = 1.0, hi_s: float = 10.0,
                      sr: int = SR) -> np.ndarray:
    """cfg4: utterance lengths ~ U[lo, hi] seconds, in samples."""
    r = np.random.default_rng(seed)
    return (r.uniform(lo_s, hi_s, n_utts) * sr).astype(np.int64)


def lognormal_lengths(seed: int, n_utts: int, median_s: float = 5.0, lo_s: float = 1.0,
                      hi_s: float = 20.0, sr: int = 24000) -> np.ndarray:
    """cfg5: LibriTTS-R-like skewed lengths."""
    r = np.random.default_rng(seed)
    s = np.clip(r.lognormal(np.log(median_s), 0.6, n_utts), lo_s, hi_s)
    return (s * sr).astype(np.int64)


def cfg2_batch(seed: int = 2, B: int = 32, T: int = 200, H: int = 256):
    """cfg2: x ~ N(0,1) [B,T,H]; lens ~ U{50..T}; dur ~ U{0..20}, zero beyond lens."""
    r = np.random.default_rng(seed)
    lens = r.integers(50, T + 1, B)
    lens[0] = T
    x = r.standard_normal((B, T, H)).astype(np.float32)
    dur = r.integers(0, 21, (B, T)).astype(np.int64)
    dur[np.arange(T)[None, :] >= lens[:, None]] = 0
    return x, dur, lens.astype(np.int64)


def cfg2_features(seed: int = 2, B: int = 32, T: int = 200, n: int = 5):
    r = np.random.default_rng(seed + 1000)
    return [r.standard_normal((B, T)).astype(np.float32) for _ in range(n)]


def lr_edge_cases():
    """Edge semantics of the reference LengthRegulator (SURVEY section 7 hard part 5)."""
    r = np.random.default_rng(5)
    H = 6
    cases = {}
    x = r.standard_normal((3, 4, H)).astype(np.float32)
    cases["zero_row"] = (x, np.array([[2, 0, 1, 3], [0, 0, 0, 0], [1, 1, 1, 1]], dtype=np.int64))
    cases["gt1000"] = (x, np.array([[2, 1001, 1, 0], [1000, 0, 0, 0], [0, 0, 0, 5]], dtype=np.int64))
    cases["negative"] = (x, np.array([[2, -1, 1, 0], [-5, -5, -5, -5], [3, 0, -2, 1]], dtype=np.int64))
    cases["float"] = (x, np.array([[1.9, np.nan, 2.0, -1.0], [0.99, 0.0, np.inf, 1000.5],
                                   [1000.0, 0.5, -np.inf, 3.999]], dtype=np.float32))
    cases["all_empty"] = (x, np.zeros((3, 4), dtype=np.int64))
    x1 = r.standard_normal((1, 1, 1)).astype(np.float32)
    cases["one"] = (x1, np.array([[7]], dtype=np.int64))
    return cases


def upstream_grad(shape, seed: int) -> np.ndarray:
    """N(0,1) upstream gradient (``dL/d out``) shared between the golden generator and the GPU backward tests."""
    return np.random.default_rng(seed).standard_normal(tuple(int(s) for s in shape)).astype(np.float32)


def log_durations(seed: int = 7, B: int = 32, T: int = 200) -> np.ndarray:
    r = np.random.default_rng(seed)
    return np.clip(r.standard_normal((B, T)), -4, 4).astype(np.float32)


def bucketize_case(seed: int = 2, B: int = 32, T: int = 200, n_bins: int = 256, H: int = 256):
    r = np.random.default_rng(seed + 2000)
    bins = np.linspace(-3, 3, n_bins - 1).astype(np.float32)
    v = r.standard_normal((B, T)).astype(np.float32)
    v[0, :8] = [np.nan, np.inf, -np.inf, -3.0, 3.0, bins[10], 3.0001, bins[100]]
    m = min(T, n_bins - 1)
    v[1, :m] = bins[:m]  # values exactly on boundaries
    v[2, :m] = np.nextafter(bins[:m], np.float32(np.inf))
    v[3, :m] = np.nextafter(bins[:m], np.float32(-np.inf))
    table = r.standard_normal((n_bins, H)).astype(np.float32)
    return v, bins, table


def init_phase(shape, seed: int = 3) -> np.ndarray:
    """2*pi*U[0,1) phases shared between the oracle and the CUDA path."""
    return (2 * np.pi * np.random.default_rng(seed).random(size=shape)).astype(np.float32)


def cfg3_logmels(seed: int = 3, B: int = 16, T: int = 800):
    """cfg3 inputs: log-mels (oracle layout [B,80,T]) of `speechy` signals; built by the
    caller with the oracle.  Here: just the raw signals."""
    n = (T - 1) * HOP
    return [speechy(seed=seed * 100 + b, n=n) for b in range(B)]


def cache_records(seed: int = 8, n: int = 12, n_mels: int = 80):
    """Synthetic cache records in the reference's format (spev_real_metrics.py:419-425) + vocab/stats."""
    import torch
    r = np.random.default_rng(seed)
    vocab = sorted(["<PAD>", "<UNK>", "<SIL>"] + [chr(97 + i) for i in range(20)])
    recs = []
    for _ in range(n):
        P = int(r.integers(3, 14))
        durs = r.integers(1, 9, P)
        T = int(durs.sum())
        phs = ["<SIL>"] + [str(x) for x in r.choice(vocab + ["zz"], P - 2)] + ["<SIL>"]      # "zz": out of vocab -> id 0
        recs.append({"phs": phs, "durs": [int(d) for d in durs],
                     "mel": torch.from_numpy(np.clip(r.standard_normal((T, n_mels)) * 2 - 4, -10, 2).astype(np.float32)),
                     "pitch": np.clip(r.standard_normal(P), -2.5, 2.5), "energy": np.clip(r.standard_normal(P), -2.5, 2.5),
                     "breath": np.clip(r.random(P), 0, 0.8), "rough": np.clip(r.random(P) * 2, 0, 1.5),
                     "bright": np.clip(r.standard_normal(P), -2.5, 2.5)})
    stats = {"p_mean": 5.0, "p_std": 0.3, "e_mean": -3.0, "e_std": 1.0, "c_mean": 7.0, "c_std": 0.5}
    return recs, stats, vocab


def voiced_unvoiced(seed: int = 0, n: int = 3 * SR, sr: int = SR):
    """Speech-like pitch test signal: harmonic segments with gliding f0 (80-400 Hz), separated by noise
    bursts and silences.  -> (y float32 [n], f0_true float64 [n] with 0 where unvoiced)."""
    r = np.random.default_rng(seed)
    y = np.zeros(n)
    f_true = np.zeros(n)
    pos = 0
    kind = int(r.integers(0, 3))
    while pos < n:
        seg = int(r.uniform(0.08, 0.45) * sr)
        e = min(n, pos + seg)
        m = e - pos
        if kind == 0:                                      # voiced
            fa, fb = r.uniform(80, 400, 2)
            fb = float(np.clip(fb, fa / 1.5, fa * 1.5))
            f = np.linspace(fa, fb, m)
            ph = 2 * np.pi * np.cumsum(f) / sr
            amp = r.uniform(0.05, 0.4) * np.hanning(m + 2)[1:-1] ** 0.25
            nh = int(r.integers(2, 12))
            y[pos:e] = amp * sum(np.sin(k * ph + r.uniform(0, 6)) / k for k in range(1, nh + 1))
            f_true[pos:e] = f
        elif kind == 1:                                    # fricative-like noise
            y[pos:e] = r.uniform(0.01, 0.1) * r.standard_normal(m)
        # kind == 2: silence
        pos = e
        kind = int((kind + r.integers(1, 3)) % 3)
    y += 1e-4 * r.standard_normal(n)
    return y.astype(np.float32), f_true


def tiny_corpus(seed: int = 21, n: int = 14):
    """A miniature LJSpeech-like corpus for the cache-build tests (spev_real_metrics.py:300-430):
    list of {'name', 'y', 'text', 'intervals'}.  ``text`` drives the uniform-alignment path (:353-357),
    ``intervals`` [(t0, t1, mark)] the TextGrid path (:337-351; marks may be empty -> '<SIL>', intervals may be
    too short to get a frame).  Item 3 is shorter than 4000 samples (skipped at :333), item 5 has neither
    text nor intervals (skipped at :359), item 7 has more phones than frames (uniform duration 0 -> :378)."""
    r = np.random.default_rng(seed)
    items = []
    letters = "abcdefghijklmnop rstuv"
    for i in range(n):
        n_s = int(r.uniform(0.5, 1.8) * SR)
        if i == 3:
            n_s = 3500
        y = voiced_unvoiced(seed=100 + seed + i, n=n_s)[0]
        text = "".join(r.choice(list(letters), int(r.integers(4, 28))))
        intervals = None
        if i == 5:
            text = None
        if i == 7:
            text = "".join(r.choice(list(letters), 1 + n_s // HOP + 9))
        if i % 3 == 1 and i != 7:                            # TextGrid-aligned items: ragged durations
            t, intervals = 0.0, []
            dur_s = n_s / SR
            stretch = 3.0 if i == 10 else float(r.uniform(0.8, 1.3))   # alignments rarely match the audio length;
            while t < dur_s * stretch:                       # item 10: 3x too long, 1-2 frame phones -> tail trimming
                d = float(r.choice([0.013, 0.02, 0.03] if i == 10 else [0.004, 0.02, 0.05, 0.11, 0.3]))
                mark = "" if r.random() < 0.2 else str(r.choice(list("aeioukstn")))
                intervals.append((t, t + d, mark))
                t += d
        items.append({"name": f"utt_{i:03d}", "y": y, "text": text, "intervals": intervals})
    return items


def corpus_alignments(corpus, sr: int = SR):
    """What the reference holds in (phs, durs) at spev_real_metrics.py:359 for each item of ``tiny_corpus``:
    TextGrid intervals -> frames = int(dur * sr / 256) if > 0 (:345-349), else the text split into characters
    between two '<SIL>' with uniform durations (:353-357); (None, None) when neither exists."""
    phones, durs = [], []
    for it in corpus:
        ph, du = [], []
        if it["intervals"] is not None:
            for a, b, mark in it["intervals"]:
                frames = int((b - a) * sr / 256)
                if frames > 0:
                    ph.append(mark if mark else "<SIL>")
                    du.append(frames)
        if not ph and it["text"] is not None:
            ph = ["<SIL>"] + list(it["text"]) + ["<SIL>"]
            du = [int((len(it["y"]) / 256) / len(ph))] * len(ph)
        phones.append(ph or None)
        durs.append(du or None)
    return phones, durs


def _cfg3_oracle_job(args):
    """(worker for a spawn pool) oracle Griffin-Lim of one cfg3 item -> its spectral convergences at the n_iter list."""
    b, n_iters, sr, T = args
    from oracle import librosa_restated as lr
    lm = lr.reference_logmel(speechy(seed=300 + b, n=(T - 1) * HOP, sr=sr), sr=sr).T.copy()
    S = lr.mel_to_stft(np.exp(lm), sr=sr, n_fft=1024, fmin=0, fmax=8000, lbfgs=True)      # librosa's own solver
    ph = init_phase((513, T), seed=3000 + b)
    return b, [lr.spectral_convergence(lr.griffinlim(S, n_iter=n, hop_length=HOP, n_fft=1024, init_phase=ph), S)
               for n in n_iters]


def cfg3_oracle_sc(items, n_iters, sr=SR, T=800, procs=None):
    """Oracle spectral convergences of cfg3 items (configs[2]: [80,800] log-mels of `speechy` signals, shared initial
    phases seed 3000+b), fanned over a spawn pool (the caller may already hold a CUDA context)."""
    import multiprocessing as mp
    import os
    procs = procs or min(len(items), os.cpu_count() or 1)
    with mp.get_context("spawn").Pool(procs) as pool:
        res = dict(pool.map(_cfg3_oracle_job, [(b, tuple(n_iters), sr, T) for b in items]))
    return res
