"""CPU oracle for the SPEV-TTS spectral hot path  --  TEST INFRASTRUCTURE ONLY.

This module is the checker, never the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  Nothing under ``spev_tts_b200/`` imports ``oracle``.

What it restates
----------------
The reference (``/root/reference/spev_real_metrics.py``) performs the spectral part
of its hot path inside the third-party package **librosa** (pinned by
``requirements_conda.txt:42`` to the 0.11.0 conda build; floor ``>=0.10.0`` in
``pyproject.toml:45``).  librosa is *not vendored* in the reference tree and is not
installable in the build container, so this file restates the published librosa
0.11 algorithms, in numpy/scipy, for exactly the calls the reference makes:

* ``librosa.feature.melspectrogram(y=y, sr=22050, n_fft=1024, hop_length=256,
  n_mels=80)``                              -- call site ``spev_real_metrics.py:363``
* ``np.log(np.clip(mel, 1e-5, None))``; ``np.clip(mel, -10, 2)``; store ``mel.T``
                                            -- ``spev_real_metrics.py:364-367``, ``:421``
* ``librosa.feature.inverse.mel_to_audio(exp(mel), sr=22050, n_fft=1024,
  hop_length=256, fmin=0, fmax=8000)``      -- ``spev_real_metrics.py:725-733``

PARITY PIN STATUS
-----------------
The reference ships **no tests, golden vectors or fixtures** (SURVEY.md section 4), and
librosa itself cannot run here, so for the spectral functions parity is
**unpinned by the reference**.  The restatement is instead pinned against
independent implementations that *are* available in the container
(``tests/test_oracle.py``): ``torchaudio.functional.melscale_fbanks`` (Slaney
basis), ``torch.stft`` / ``torch.istft`` in float64 (framing, padding, window,
overlap-add, window-sum-square normalisation), ``scipy.optimize.fmin_l_bfgs_b``
(the very routine librosa's NNLS calls), ``torchaudio.functional.griffinlim`` (the
Griffin-Lim loop: momentum rule, phase normalisation, trailing ISTFT -- compared with
the re-analysis padded by reflection, as torchaudio does it) and ``torch.bucketize``.
The LengthRegulator restatement *is* pinned: ``tests/golden/lr_*.npz`` were
produced by the reference's own class (``oracle/make_golden.py`` imports
``/root/reference/spev_real_metrics.py:122-146``).

Precision conventions follow librosa: the Hann window is float64, ``window * frames``
promotes to float64, the FFT runs in float64 and is stored as complex64; power, mel
projection and log run in float32; the ISTFT overlap-add and the window-sum-square
accumulate into float32 buffers in ascending frame order.
"""
from __future__ import annotations

import numpy as np
import scipy.fft
import scipy.optimize
import scipy.signal

MAX_MEM_BLOCK = 2 ** 18  # librosa.util.MAX_MEM_BLOCK (bytes); drives NNLS column blocking


# ----------------------------------------------------------------------------------
# mel scale / filterbank  (librosa.core.convert.hz_to_mel / mel_to_hz, librosa.filters.mel)
# ----------------------------------------------------------------------------------
def hz_to_mel(frequencies, htk: bool = False):
    """Slaney (default) or HTK mel scale.  SURVEY App. A.1."""
    frequencies = np.asanyarray(frequencies, dtype=np.float64)
    if htk:
        return 2595.0 * np.log10(1.0 + frequencies / 700.0)
    f_min, f_sp = 0.0, 200.0 / 3
    mels = (frequencies - f_min) / f_sp
    min_log_hz = 1000.0
    min_log_mel = (min_log_hz - f_min) / f_sp
    logstep = np.log(6.4) / 27.0
    if frequencies.ndim:
        log_t = frequencies >= min_log_hz
        mels[log_t] = min_log_mel + np.log(frequencies[log_t] / min_log_hz) / logstep
    elif frequencies >= min_log_hz:
        mels = min_log_mel + np.log(frequencies / min_log_hz) / logstep
    return mels


def mel_to_hz(mels, htk: bool = False):
    mels = np.asanyarray(mels, dtype=np.float64)
    if htk:
        return 700.0 * (10.0 ** (mels / 2595.0) - 1.0)
    f_min, f_sp = 0.0, 200.0 / 3
    freqs = f_min + f_sp * mels
    min_log_hz = 1000.0
    min_log_mel = (min_log_hz - f_min) / f_sp
    logstep = np.log(6.4) / 27.0
    if mels.ndim:
        log_t = mels >= min_log_mel
        freqs[log_t] = min_log_hz * np.exp(logstep * (mels[log_t] - min_log_mel))
    elif mels >= min_log_mel:
        freqs = min_log_hz * np.exp(logstep * (mels - min_log_mel))
    return freqs


def mel_frequencies(n_mels: int, fmin: float, fmax: float, htk: bool = False):
    mels = np.linspace(hz_to_mel(fmin, htk=htk), hz_to_mel(fmax, htk=htk), n_mels)
    return mel_to_hz(mels, htk=htk)


def mel_filter(*, sr, n_fft, n_mels=128, fmin=0.0, fmax=None, htk=False, norm="slaney",
               dtype=np.float32):
    """librosa.filters.mel.  Rows are written into a ``dtype`` (float32) array, the
    Slaney area normalisation is then applied in that dtype.  The reference's forward
    path uses fmax=None -> sr/2 (``spev_real_metrics.py:363``), the inverse path uses
    fmin=0, fmax=8000 (``:730-733``): the two bases differ."""
    if fmax is None:
        fmax = float(sr) / 2
    n_mels = int(n_mels)
    weights = np.zeros((n_mels, int(1 + n_fft // 2)), dtype=dtype)
    fftfreqs = np.fft.rfftfreq(n=n_fft, d=1.0 / sr)
    mel_f = mel_frequencies(n_mels + 2, fmin=fmin, fmax=fmax, htk=htk)
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    if norm == "slaney":
        enorm = 2.0 / (mel_f[2: n_mels + 2] - mel_f[:n_mels])
        weights *= enorm[:, np.newaxis]
    elif norm is not None:
        raise ValueError("only norm='slaney' or None is restated")
    return weights


# ----------------------------------------------------------------------------------
# STFT / ISTFT  (librosa.core.spectrum.stft / istft / window_sumsquare)
# ----------------------------------------------------------------------------------
def _hann(win_length: int, n_fft: int):
    w = scipy.signal.get_window("hann", win_length, fftbins=True)  # periodic, float64
    if win_length < n_fft:  # util.pad_center
        lpad = (n_fft - win_length) // 2
        w = np.pad(w, (lpad, n_fft - win_length - lpad))
    return w


def _dtype_r2c(d):
    return {np.dtype(np.float32): np.complex64, np.dtype(np.float64): np.complex128}.get(
        np.dtype(d), np.complex64)


def _dtype_c2r(d):
    return {np.dtype(np.complex64): np.float32, np.dtype(np.complex128): np.float64}.get(
        np.dtype(d), np.float32)


def stft(y, *, n_fft=2048, hop_length=None, win_length=None, center=True,
         pad_mode="constant"):
    """librosa.stft for window='hann'.  ``[..., N] -> [..., 1+n_fft/2, 1+N//hop]``.
    center=True pads n_fft//2 zeros on both sides (librosa computes head and tail
    blocks separately; the result equals np.pad)."""
    if win_length is None:
        win_length = n_fft
    if hop_length is None:
        hop_length = win_length // 4
    if pad_mode != "constant":
        raise ValueError("only pad_mode='constant' is on the reference path")
    y = np.asarray(y)
    w = _hann(win_length, n_fft)
    if center:
        pad = [(0, 0)] * (y.ndim - 1) + [(n_fft // 2, n_fft // 2)]
        y = np.pad(y, pad, mode="constant")
    if y.shape[-1] < n_fft:
        raise ValueError("input too short")
    n_frames = 1 + (y.shape[-1] - n_fft) // hop_length
    out = np.empty(y.shape[:-1] + (1 + n_fft // 2, n_frames), dtype=_dtype_r2c(y.dtype))
    # block over frames to bound the float64 temporaries
    blk = max(1, (1 << 24) // (n_fft * max(1, int(np.prod(y.shape[:-1])))))
    for s in range(0, n_frames, blk):
        t = min(n_frames, s + blk)
        idx = (np.arange(s, t) * hop_length)[None, :] + np.arange(n_fft)[:, None]
        frames = y[..., idx]  # [..., n_fft, t-s]
        out[..., s:t] = scipy.fft.rfft(w[:, None] * frames, axis=-2)  # f64 FFT -> c64 store
    return out


def window_sumsquare(*, n_frames, hop_length, win_length, n_fft, dtype=np.float32):
    n = n_fft + hop_length * (n_frames - 1)
    x = np.zeros(n, dtype=dtype)
    win_sq = _hann(win_length, n_fft) ** 2  # float64
    for i in range(n_frames):  # ascending, accumulating in `dtype` (numba loop in librosa)
        s = i * hop_length
        x[s: min(n, s + n_fft)] += win_sq[: max(0, min(n_fft, n - s))]
    return x


def istft(X, *, hop_length=None, win_length=None, n_fft=None, center=True,
          dtype=None, length=None):
    """librosa.istft for window='hann'.  ``[..., 1+n_fft/2, T] -> [..., (T-1)*hop]``.
    Overlap-add accumulates float64 frames into a float32 buffer in ascending frame
    order, then divides by the float32 window-sum-square where it exceeds
    ``finfo(float32).tiny``."""
    X = np.asarray(X)
    if n_fft is None:
        n_fft = 2 * (X.shape[-2] - 1)
    if win_length is None:
        win_length = n_fft
    if hop_length is None:
        hop_length = win_length // 4
    if dtype is None:
        dtype = _dtype_c2r(X.dtype)
    w = _hann(win_length, n_fft)
    T = X.shape[-1]
    full = n_fft + hop_length * (T - 1)
    buf = np.zeros(X.shape[:-2] + (full,), dtype=dtype)
    blk = max(1, (1 << 24) // (n_fft * max(1, int(np.prod(X.shape[:-2])))))
    for s in range(0, T, blk):
        t = min(T, s + blk)
        ytmp = w[:, None] * scipy.fft.irfft(X[..., s:t], n=n_fft, axis=-2)  # float64
        for f in range(s, t):
            buf[..., f * hop_length: f * hop_length + n_fft] += ytmp[..., f - s]
    wss = window_sumsquare(n_frames=T, hop_length=hop_length, win_length=win_length,
                           n_fft=n_fft, dtype=dtype)
    start = n_fft // 2 if center else 0
    if length is None:
        out_len = full - 2 * start
    else:
        out_len = length
    y = np.zeros(X.shape[:-2] + (out_len,), dtype=dtype)
    avail = min(out_len, full - start)
    y[..., :avail] = buf[..., start: start + avail]
    wss_fix = np.zeros(out_len, dtype=dtype)
    wss_fix[:avail] = wss[start: start + avail]
    nz = wss_fix > np.finfo(dtype).tiny
    y[..., nz] /= wss_fix[nz]
    return y


# ----------------------------------------------------------------------------------
# forward path: melspectrogram + the reference's log compression
# ----------------------------------------------------------------------------------
def melspectrogram(*, y, sr=22050, n_fft=2048, hop_length=512, win_length=None,
                   center=True, pad_mode="constant", power=2.0, n_mels=128,
                   fmin=0.0, fmax=None):
    """librosa.feature.melspectrogram: |STFT|**power (float32) -> einsum with the
    float32 Slaney basis.  ``[..., N] -> [..., n_mels, T]``."""
    S = np.abs(stft(y, n_fft=n_fft, hop_length=hop_length, win_length=win_length,
                    center=center, pad_mode=pad_mode)) ** power
    basis = mel_filter(sr=sr, n_fft=n_fft, n_mels=n_mels, fmin=fmin, fmax=fmax)
    return np.einsum("...ft,mf->...mt", S, basis, optimize=True)


def reference_logmel(y, *, sr=22050, n_fft=1024, hop_length=256, n_mels=80):
    """Exactly the four statements at ``spev_real_metrics.py:363-367`` followed by the
    stored layout of ``:421`` (``mel.T`` -> ``[T, n_mels]`` float32)."""
    mel = melspectrogram(y=y, sr=sr, n_fft=n_fft, hop_length=hop_length, n_mels=n_mels)
    mel = np.log(np.clip(mel, a_min=1e-5, a_max=None))
    mel = np.clip(mel, -10.0, 2.0)
    return np.ascontiguousarray(np.swapaxes(mel.astype(np.float32), -1, -2))


# ----------------------------------------------------------------------------------
# inverse path: mel -> linear magnitude (NNLS) -> Griffin-Lim
# ----------------------------------------------------------------------------------
def _nnls_obj(x, shape, A, B):
    x = x.reshape(shape)
    diff = np.einsum("mf,...ft->...mt", A, x, optimize=True) - B
    value = (1 / B.size) * 0.5 * np.sum(diff ** 2)
    grad = (1 / B.size) * np.einsum("mf,...mt->...ft", A, diff, optimize=True)
    return value, grad.flatten()


def _nnls_lbfgs_block(A, B, x_init=None, lbfgs=True):
    if x_init is None:
        x_init = np.einsum("fm,...mt->...ft", np.linalg.pinv(A), B, optimize=True)
        np.clip(x_init, 0, None, out=x_init)
    if not lbfgs:
        return x_init
    shape = x_init.shape
    bounds = [(0, None)] * x_init.size
    x, _, _ = scipy.optimize.fmin_l_bfgs_b(_nnls_obj, x_init, args=(shape, A, B),
                                           bounds=bounds, m=A.shape[1])
    return x.reshape(shape)


def nnls(A, B, lbfgs=True):
    """librosa.util.nnls.  ``lbfgs=False`` returns the warm start
    ``clip(pinv(A) @ B, 0)`` -- which is what L-BFGS-B returns at iteration 0 for
    reference-range inputs (log-mel <= 2, T >~ 40)."""
    n_columns = MAX_MEM_BLOCK // (int(np.prod(B.shape[:-1])) * A.itemsize)
    n_columns = max(n_columns, 1)
    if B.shape[-1] <= n_columns:
        return _nnls_lbfgs_block(A, B, lbfgs=lbfgs).astype(A.dtype)
    x = np.einsum("fm,...mt->...ft", np.linalg.pinv(A), B, optimize=True)
    np.clip(x, 0, None, out=x)
    x_init = x
    if lbfgs:
        for s in range(0, x.shape[-1], n_columns):
            t = min(s + n_columns, B.shape[-1])
            x[..., s:t] = _nnls_lbfgs_block(A, B[..., s:t], x_init=x_init[..., s:t])
    return x


def mel_to_stft(M, *, sr=22050, n_fft=2048, power=2.0, fmin=0.0, fmax=None, lbfgs=True):
    M = np.asarray(M)
    basis = mel_filter(sr=sr, n_fft=n_fft, n_mels=M.shape[-2], dtype=M.dtype,
                       fmin=fmin, fmax=fmax)
    inverse = nnls(basis, M, lbfgs=lbfgs)
    return np.power(inverse, 1.0 / power, out=inverse)


def phasor(angles):
    """librosa.util.phasor: cos + j sin, evaluated in float64/complex128."""
    angles = np.asarray(angles, dtype=np.float64)
    return np.cos(angles) + 1j * np.sin(angles)


def griffinlim_step(angles, tprev, S, *, hop_length, n_fft, momentum=0.99):
    """One body of librosa.griffinlim's loop.  Returns (new_angles, rebuilt, inverse)."""
    eps = np.finfo(np.float32).tiny
    inverse = istft(angles, hop_length=hop_length, n_fft=n_fft, dtype=np.float32)
    rebuilt = stft(inverse, n_fft=n_fft, hop_length=hop_length)
    new = rebuilt.copy()
    if tprev is not None:
        new -= (momentum / (1 + momentum)) * tprev
    new /= np.abs(new) + eps
    new *= S
    return new, rebuilt, inverse


def griffinlim(S, *, n_iter=32, hop_length=None, n_fft=None, momentum=0.99,
               init_phase=None, seed=None, return_state=False):
    """librosa.griffinlim with window='hann', center=True,
    pad_mode='constant', dtype=float32.  ``init_phase`` (radians, same shape as S)
    replaces librosa's ``2*pi*default_rng().random(S.shape)`` so that a run can be
    reproduced; the reference itself seeds from OS entropy (non-deterministic)."""
    S = np.asarray(S, dtype=np.float32)
    if n_fft is None:
        n_fft = 2 * (S.shape[-2] - 1)
    if hop_length is None:
        hop_length = n_fft // 4
    if init_phase is None:
        rng = np.random.default_rng(seed)
        init_phase = 2 * np.pi * rng.random(size=S.shape)
    angles = np.empty(S.shape, dtype=np.complex64)
    angles[:] = phasor(init_phase)
    angles *= S
    tprev = None
    for _ in range(n_iter):
        angles, tprev, _ = griffinlim_step(angles, tprev, S, hop_length=hop_length,
                                           n_fft=n_fft, momentum=momentum)
    y = istft(angles, hop_length=hop_length, n_fft=n_fft, dtype=np.float32)
    if return_state:
        return y, angles, tprev
    return y


def mel_to_audio(M, *, sr=22050, n_fft=2048, hop_length=None, power=2.0, n_iter=32,
                 fmin=0.0, fmax=None, init_phase=None, seed=None, lbfgs=True):
    S = mel_to_stft(M, sr=sr, n_fft=n_fft, power=power, fmin=fmin, fmax=fmax, lbfgs=lbfgs)
    return griffinlim(S, n_iter=n_iter, hop_length=hop_length, n_fft=n_fft,
                      init_phase=init_phase, seed=seed)


def reference_vocoder_infer(logmel, *, n_iter=32, init_phase=None, seed=None, lbfgs=True):
    """``Vocoder.infer`` Griffin-Lim branch, ``spev_real_metrics.py:725-733`` with the
    module CONFIG (``:60-67``): exp -> mel_to_audio(sr=22050, n_fft=1024, hop=256,
    fmin=0, fmax=8000).  ``logmel`` is ``[..., 80, T]`` float32."""
    mel_exp = np.exp(np.asarray(logmel, dtype=np.float32))
    return mel_to_audio(mel_exp, sr=22050, n_fft=1024, hop_length=256, fmin=0, fmax=8000,
                        n_iter=n_iter, init_phase=init_phase, seed=seed, lbfgs=lbfgs)


def spectral_convergence(y, S, *, n_fft=1024, hop_length=256):
    """SC = || |STFT(y)| - S ||_F / ||S||_F  (SURVEY section 8c tolerance (ii))."""
    R = np.abs(stft(y, n_fft=n_fft, hop_length=hop_length))
    return float(np.linalg.norm(R - S) / np.linalg.norm(S))


# ----------------------------------------------------------------------------------
# LengthRegulator / duration rule / bucketize  (integer index work: bit-exact)
# ----------------------------------------------------------------------------------
def sanitize_durations(d):
    """The per-element rule of ``LengthRegulator.forward`` (``spev_real_metrics.py:
    129-133``): non-finite, negative or >1000 -> 0, then ``int()`` truncation."""
    d = np.asarray(d)
    if d.dtype.kind in "iu":
        bad = (d < 0) | (d > 1000)
        return np.where(bad, 0, d).astype(np.int64)
    d = d.astype(np.float64)
    bad = ~np.isfinite(d) | (d < 0) | (d > 1000)
    return np.trunc(np.where(bad, 0.0, d)).astype(np.int64)


def length_regulator(x, durations):
    """Vectorised restatement of ``spev_real_metrics.py:122-146``.
    x ``[B,T,H]``, durations ``[B,T]`` -> (``[B,maxF,H]``, int64 ``[B]``).  An empty row
    yields one zero frame with length 1."""
    x = np.asarray(x)
    n = sanitize_durations(durations)
    B, T = n.shape
    if B == 0:
        raise ValueError("max() arg is an empty sequence")  # reference: max(mel_lens)
    cs = np.cumsum(n, axis=1)
    tot = cs[:, -1] if T > 0 else np.zeros(B, dtype=np.int64)
    mel_lens = np.maximum(tot, 1)
    max_len = int(mel_lens.max())
    out = np.zeros((B, max_len, x.shape[2]), dtype=x.dtype)
    for b in range(B):
        if tot[b] > 0:
            idx = np.searchsorted(cs[b], np.arange(tot[b]), side="right")
            out[b, : tot[b]] = x[b, idx]
    return out, mel_lens.astype(np.int64)


def duration_rule(log_dur, d_control=1.0):
    """``spev_real_metrics.py:215``: clamp((exp(ld)-1)*d_control, 0, 500).round().long()
    in float32 with round-half-to-even."""
    ld = np.asarray(log_dur, dtype=np.float32)
    v = (np.exp(ld) - np.float32(1.0)) * np.float32(d_control)
    v = np.clip(v, np.float32(0), np.float32(500))
    return np.rint(v).astype(np.int64)


def bucketize(v, boundaries, right=False):
    """torch.bucketize semantics: right=False -> first i with boundaries[i] >= v;
    NaN and values above the last boundary -> len(boundaries) (SURVEY a-13)."""
    v = np.asarray(v)
    b = np.asarray(boundaries)
    idx = np.searchsorted(b, v, side="right" if right else "left")
    idx = np.where(np.isnan(v), len(b), idx)
    return idx.astype(np.int64)


def bucketize_embed(v, boundaries, table, right=False):
    idx = bucketize(v, boundaries, right=right)
    return idx, np.asarray(table)[idx]


# ----------------------------------------------------------------------------------
# frame-level energy / brightness and per-phoneme pooling
#   reference: spev_real_metrics.py:370-371 (rms, spectral_centroid) and :400-417 (pooling)
# ----------------------------------------------------------------------------------
def rms(*, y, frame_length=2048, hop_length=512, center=True, pad_mode="constant"):
    """librosa.feature.rms(y=...) -> ``[..., 1, T]`` float32: sqrt(mean(frame**2)) over
    centre-padded frames; squares and the mean (reduction over the frame axis) in float32."""
    y = np.asarray(y)
    if center:
        pad = [(0, 0)] * (y.ndim - 1) + [(frame_length // 2, frame_length // 2)]
        y = np.pad(y, pad, mode=pad_mode)
    n_frames = 1 + (y.shape[-1] - frame_length) // hop_length
    idx = (np.arange(n_frames) * hop_length)[None, :] + np.arange(frame_length)[:, None]
    x = y[..., idx]                                    # [..., frame_length, T]
    power = np.mean(np.abs(x).astype(np.float32) ** 2, axis=-2, keepdims=True)
    return np.sqrt(power)


def spectral_centroid(*, y, sr=22050, n_fft=2048, hop_length=512, center=True, pad_mode="constant"):
    """librosa.feature.spectral_centroid(y=...) -> ``[..., 1, T]`` float64:
    sum(freq * normalize(|STFT|, norm=1)), columns with an l1 norm below tiny left unscaled."""
    S = np.abs(stft(y, n_fft=n_fft, hop_length=hop_length, center=center, pad_mode=pad_mode))
    freq = np.fft.rfftfreq(n=n_fft, d=1.0 / sr)
    length = np.sum(np.abs(S), axis=-2, keepdims=True)
    length = np.where(length < np.finfo(S.dtype).tiny, 1.0, length).astype(S.dtype)
    Snorm = S / length
    return np.sum(freq[:, None] * Snorm, axis=-2, keepdims=True)


def phoneme_pool(curve, durs, mu, sigma, lo, hi):
    """One of the per-phone lines of ``spev_real_metrics.py:400-417``:
    ``np.clip((np.mean(curve[curr:curr+d]) - mu) / sigma, lo, hi)`` for consecutive durations."""
    out, curr = [], 0
    for d in durs:
        sl = slice(curr, curr + int(d))
        out.append(np.clip((np.mean(curve[sl]) - mu) / sigma, lo, hi))
        curr += int(d)
    return np.array(out, dtype=np.float32)
