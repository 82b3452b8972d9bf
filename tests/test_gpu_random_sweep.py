"""Seeded random sweeps against the oracle: ragged log-mel batches (lengths, packing offsets, sample rates, mel counts),
LengthRegulator forward + backward (shapes, duration patterns, dtypes), Griffin-Lim on ragged batches."""
import numpy as np
import pytest
import torch

from oracle import librosa_restated as lr
from oracle import torch_reference as tr
from tests import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", range(6))
def test_random_ragged_logmel_vs_oracle(cuda, seed):
    import spev_tts_b200 as sp
    rng = np.random.default_rng(1000 + seed)
    sr = int(rng.choice([16000, 22050, 24000]))
    n_mels = int(rng.choice([40, 64, 80, 100]))
    n = int(rng.integers(1, 9))
    lens = [int(v) for v in rng.choice([0, 1, 255, 256, 257, 1024, 4097, 9999, 30000, 70001], n)]
    gaps = rng.integers(0, 7, n)                                  # arbitrary (unaligned) gaps between the items
    starts = np.concatenate([[0], np.cumsum(np.array(lens) + gaps)])[:-1].astype(np.int64)
    buf = np.full(int(starts[-1] + lens[-1] + 8), 7.0, np.float32)   # filler that must never be read as signal
    ys = []
    for s0, ln in zip(starts, lens):
        y = (rng.choice([0.001, 0.05, 0.5]) * rng.standard_normal(ln)).astype(np.float32)
        buf[s0: s0 + ln] = y
        ys.append(y)
    out, fb = sp.logmel_flat(torch.from_numpy(buf).to(cuda), lens, sr=sr, n_mels=n_mels, sample_off=starts)
    out = out.cpu().numpy()
    assert fb.n_frames == sum(1 + ln // 256 for ln in lens)
    for i, y in enumerate(ys):
        ref = lr.reference_logmel(y, sr=sr, n_mels=n_mels)
        got = out[fb.frame_off[i]: fb.frame_off[i + 1]]
        assert got.shape == ref.shape and np.abs(got - ref).max() <= 1e-4, (seed, i, lens[i])


@pytest.mark.parametrize("seed", range(8))
def test_random_length_regulator_forward_backward(cuda, seed):
    import spev_tts_b200 as sp
    rng = np.random.default_rng(2000 + seed)
    B, T, H = int(rng.integers(1, 9)), int(rng.integers(1, 70)), int(rng.choice([1, 3, 4, 16, 100, 256]))
    kind = seed % 4
    d = rng.integers(0, [2, 6, 30, 12][kind], (B, T)).astype(np.float64)
    if kind == 1:
        d[rng.random((B, T)) < 0.1] = rng.choice([-3.0, 1001.0, np.nan, np.inf, 2.75])
    if kind == 3:
        d[rng.integers(0, B)] = 0                                  # an empty row -> one zero frame
    dt = torch.from_numpy(d) if kind == 1 else torch.from_numpy(d).long()
    dtype = torch.float64 if seed % 2 else torch.float32
    x = torch.from_numpy(rng.standard_normal((B, T, H))).to(dtype)
    xa = x.clone().to(cuda).requires_grad_(True)
    xb = x.clone().requires_grad_(True)
    oa, la = sp.LengthRegulator()(xa, dt.to(cuda))
    ob, lb = tr.LengthRegulator()(xb, dt)
    assert torch.equal(oa.detach().cpu(), ob.detach().to(dtype)) and torch.equal(la.cpu(), lb)
    g = torch.from_numpy(rng.standard_normal(tuple(ob.shape))).to(dtype)
    (oa * g.to(cuda)).sum().backward()
    if ob.requires_grad:
        (ob * g).sum().backward()
        want = xb.grad
    else:
        want = torch.zeros_like(x)                                 # all rows empty: constant zeros in the reference
    tol = 1e-12 if dtype == torch.float64 else 1e-6 * max(1.0, float(want.abs().max()))
    assert float((xa.grad.cpu() - want).abs().max()) <= tol, (seed, B, T, H)


@pytest.mark.parametrize("seed", range(3))
def test_random_ragged_griffinlim_vs_oracle(cuda, seed):
    """Ragged batch of short spectrograms, a few iterations from shared phases: per item SC delta and (few iterations:
    not yet chaotic) waveform agreement with the oracle."""
    import spev_tts_b200 as sp
    from spev_tts_b200 import _lib
    rng = np.random.default_rng(3000 + seed)
    frames = [int(v) for v in rng.choice([1, 2, 3, 5, 30, 33, 64, 97], int(rng.integers(2, 6)))]
    ctx = sp.Context.get(cuda, fmin=0.0, fmax=8000.0)
    fb = sp.make_batch(ctx, n_frames=frames, with_chunks=True)
    S_items = [np.abs(lr.stft(synth.speechy(seed=seed * 10 + i, n=max(T - 1, 1) * 256)[: (T - 1) * 256], n_fft=1024, hop_length=256))
               .astype(np.float32) if T > 1 else rng.random((513, 1)).astype(np.float32) for i, T in enumerate(frames)]
    ph_items = [synth.init_phase(S.shape, seed=seed * 100 + i) for i, S in enumerate(S_items)]
    S = torch.zeros(fb.n_frames, _lib.SPEC_LD)
    ph = torch.zeros(fb.n_frames, 513)
    for i, (Si, pi) in enumerate(zip(S_items, ph_items)):
        S[fb.frame_off[i]: fb.frame_off[i + 1], :513] = torch.from_numpy(Si.T.copy())
        ph[fb.frame_off[i]: fb.frame_off[i + 1]] = torch.from_numpy(pi.T.copy())
    n_iter = 3
    y = sp.griffinlim_flat(S.to(cuda), fb, ctx, n_iter=n_iter, init_phase=ph.to(cuda)).cpu().numpy()
    off = fb.out_sample_off()
    for i, (Si, pi, T) in enumerate(zip(S_items, ph_items, frames)):
        got = y[off[i]: off[i + 1]]
        ref = lr.griffinlim(Si, n_iter=n_iter, hop_length=256, n_fft=1024, init_phase=pi)
        assert got.shape == ref.shape == ((T - 1) * 256,)
        if T > 1:
            assert np.linalg.norm(got - ref) <= 2e-4 * max(np.linalg.norm(ref), 1e-6), (seed, i, T)
