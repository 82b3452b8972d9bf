"""BASELINE.json's FULL sizes through size-independent properties (the oracle cannot run 20 hours of audio in a test):
configs[3] -- 13,100 utterances, 6.2 M frames in one launch -- checked by exact scaling, shard invariance, batch
independence and oracle spot checks; configs[2]-sized STFT / ISTFT round trip."""
import numpy as np
import pytest
import torch

from oracle import librosa_restated as lr
from tests import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def corpus(cuda):
    from spev_tts_b200 import cache
    lens = synth.utterance_lengths(seed=4, n_utts=13100)
    starts = cache.aligned_offsets(lens)
    g = torch.Generator(device=cuda).manual_seed(4)
    x = torch.empty(int(starts[-1]), device=cuda)
    for s0 in range(0, x.numel(), 1 << 27):
        e0 = min(x.numel(), s0 + (1 << 27))
        x[s0:e0] = torch.randn(e0 - s0, generator=g, device=cuda) * 0.05
    return x, lens, starts


def test_cfg4_full_size_properties(cuda, corpus):
    import spev_tts_b200 as sp
    x, lens, starts = corpus
    ctx = sp.Context.get(cuda)
    fb = sp.make_batch(ctx, n_samples=lens, sample_off=starts)
    assert fb.n_frames == int((1 + lens // 256).sum()) == 6215191
    full, _ = sp.logmel_flat(x, lens, batch=fb)
    assert full.shape == (fb.n_frames, 80) and bool(torch.isfinite(full).all())
    assert float(full.min()) >= -10.0 and float(full.max()) <= 2.0                       # :366 clamp
    # (1) batch independence + oracle: an utterance inside the 13,100-item launch == the same utterance alone
    errs = []
    for u in (0, 1, 6550, 13099):
        y = x[int(starts[u]): int(starts[u]) + int(lens[u])].clone()
        alone, _ = sp.logmel_flat(y, [int(lens[u])])
        rows = full[int(fb.frame_off[u]): int(fb.frame_off[u + 1])]
        assert torch.equal(rows, alone), u
        errs.append(float(np.abs(rows.cpu().numpy() - lr.reference_logmel(y.cpu().numpy())).max()))
    assert max(errs) <= 1e-4, errs
    # (2) shard invariance: two halves in separate launches (other tile -> CTA assignment) == the one launch
    h = 6550
    fa = sp.make_batch(ctx, n_samples=lens[:h], sample_off=starts[:h])
    a, _ = sp.logmel_flat(x, lens[:h], batch=fa)
    fbb = sp.make_batch(ctx, n_samples=lens[h:], sample_off=starts[h:-1])
    b, _ = sp.logmel_flat(x, lens[h:], batch=fbb)
    assert torch.equal(torch.cat([a, b]), full)
    del a, b
    # (3) exact homogeneity of the power path: scaling the signal by 2 (exact in binary) scales every mel power by 4
    p1, _ = sp.logmel_flat(x, lens, batch=fb, log=False)
    x2 = x * 2.0
    p2, _ = sp.logmel_flat(x2, lens, batch=fb, log=False)
    assert torch.equal(p2, p1 * 4.0)
    # (4) and of the log path: log(4 p) - log(p) = log 4 wherever neither clamp is active (fast log: 3 ulp)
    l2, _ = sp.logmel_flat(x2, lens, batch=fb)
    free = (full > -9.9) & (l2 < 1.9)
    assert float(free.float().mean()) > 0.99
    assert float(((l2 - full)[free] - float(np.log(4.0))).abs().max()) <= 5e-6


def test_cfg3_sized_stft_istft_round_trip(cuda):
    import spev_tts_b200 as sp
    rng = np.random.default_rng(3)
    y = (0.1 * rng.standard_normal((16, 799 * 256))).astype(np.float32)
    X = sp.stft(torch.from_numpy(y).to(cuda), n_fft=1024, hop_length=256)
    assert X.shape == (16, 513, 800)
    yr = sp.istft(X, hop_length=256, n_fft=1024)
    assert yr.shape == (16, 799 * 256)
    assert float((yr.cpu() - torch.from_numpy(y)).abs().max()) <= 2e-6
