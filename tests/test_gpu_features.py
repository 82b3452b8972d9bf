"""GPU parity of the SURVEY 8(f) row-1 kernels: RMS energy, spectral centroid (2048-point STFT) and
per-phoneme pooling, against the restated librosa calls of spev_real_metrics.py:370-371, :400-417."""
import numpy as np
import pytest
import torch

from oracle import librosa_restated as lr
from tests import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [132300, 256 * 40, 255, 4000, 70001])
def test_rms_and_centroid_vs_oracle(cuda, n):
    import spev_tts_b200 as sp
    y = synth.speechy(seed=n % 97, n=n)
    r_ref = lr.rms(y=y, hop_length=256)[0]
    c_ref = lr.spectral_centroid(y=y, sr=22050, hop_length=256)[0]
    r = sp.rms(y=y, hop_length=256)
    c = sp.spectral_centroid(y=y, sr=22050, hop_length=256)
    assert r.shape == (1, 1 + n // 256) and c.shape == (1, 1 + n // 256)
    assert np.abs(r[0] - r_ref).max() <= 2e-6 * max(1.0, r_ref.max()) + 1e-7
    # the reference uses the logs (:370, :397): compare there too
    assert np.abs(np.log(r[0] + 1e-6) - np.log(r_ref + 1e-6)).max() <= 2e-5
    assert np.abs(np.log(c[0] + 1e-8) - np.log(c_ref + 1e-8)).max() <= 2e-5
    # stats pass (:316) calls spectral_centroid with librosa's default hop 512 = every second frame
    c512 = sp.spectral_centroid(y=y, sr=22050)
    c512_ref = lr.spectral_centroid(y=y, sr=22050)[0]
    assert c512.shape == (1, 1 + n // 512)
    assert np.abs(np.log(c512[0] + 1e-8) - np.log(c512_ref + 1e-8)).max() <= 2e-5


def test_silence_and_ragged_batch(cuda):
    import spev_tts_b200 as sp
    lens = [0, 100, 2048, 5000, 256 * 33 + 1, 30000]
    rng = np.random.default_rng(3)
    ys = [(0.1 * rng.standard_normal(n)).astype(np.float32) for n in lens]
    ys[2][:] = 0.0                                             # all-zero utterance: centroid 0, rms 0
    flat = torch.from_numpy(np.concatenate(ys)).to(cuda)
    r, c, fb = sp.frame_features_flat(flat, lens)
    r, c = r.cpu().numpy(), c.cpu().numpy()
    for i, y in enumerate(ys):
        sl = slice(fb.frame_off[i], fb.frame_off[i + 1])
        r_ref = lr.rms(y=y, hop_length=256)[0] if len(y) else np.zeros(1, np.float32)
        c_ref = lr.spectral_centroid(y=y, sr=22050, hop_length=256)[0] if len(y) else np.zeros(1)
        assert np.abs(r[sl] - r_ref).max() <= 1e-6, i
        assert np.abs(c[sl] - c_ref).max() <= 2e-5 * max(1.0, np.abs(c_ref).max()), i
    assert np.all(r[fb.frame_off[2]: fb.frame_off[3]] == 0) and np.all(c[fb.frame_off[2]: fb.frame_off[3]] == 0)


def test_segment_pool_vs_reference_lines(cuda):
    import spev_tts_b200 as sp
    rng = np.random.default_rng(4)
    U = 9
    frames = rng.integers(20, 400, U)
    fo = np.concatenate([[0], np.cumsum(frames)])
    curve = rng.standard_normal(fo[-1]).astype(np.float32)
    durs, po, ref = [], [0], []
    for u in range(U):
        k = int(rng.integers(1, 40))
        d = rng.multinomial(frames[u] - k, np.ones(k) / k) + 1      # >= 1 each, sums to frames[u]
        durs.append(d); po.append(po[-1] + k)
        ref.append(lr.phoneme_pool(curve[fo[u]: fo[u + 1]], d, 0.3, 1.7, -2.5, 2.5))
    durs = np.concatenate(durs).astype(np.int64)
    got = sp.segment_pool(torch.from_numpy(curve).to(cuda), fo, torch.from_numpy(durs).to(cuda), np.array(po),
                          mu=0.3, sigma=1.7, lo=-2.5, hi=2.5).cpu().numpy()
    assert np.abs(got - np.concatenate(ref)).max() <= 2e-6
    # breathiness line: clip(1 - mean(voiced_prob), 0, 0.8) == pool with mu=1, sigma=-1
    vp = rng.random(fo[-1]).astype(np.float32)
    got = sp.segment_pool(torch.from_numpy(vp).to(cuda), fo, torch.from_numpy(durs).to(cuda), np.array(po),
                          mu=1.0, sigma=-1.0, lo=0.0, hi=0.8).cpu().numpy()
    ref = np.concatenate([np.clip(1.0 - np.array([vp[fo[u]: fo[u + 1]][s: s + d].mean() for s, d in
                          zip(np.cumsum(dd) - dd, dd)]), 0.0, 0.8) for u, dd in
                          enumerate(np.split(durs, po[1:-1]))]).astype(np.float32)
    assert np.abs(got - ref).max() <= 2e-6


def test_segment_pool_log_mode(cuda):
    """spev_segment_pool_log == pooling the per-frame log, the way :370 / :397 / :404 / :408 do it in numpy."""
    import spev_tts_b200 as sp
    rng = np.random.default_rng(9)
    frames = np.array([40, 7, 123])
    fo = np.concatenate([[0], np.cumsum(frames)])
    curve = (rng.random(fo[-1]) * 0.2).astype(np.float32)
    durs = [np.array([10, 0, 30]), np.array([7]), np.array([100, 20, 3])]
    po = np.concatenate([[0], np.cumsum([len(d) for d in durs])])
    want = []
    for u, d in enumerate(durs):
        logc = np.log(curve[fo[u]: fo[u + 1]] + 1e-6)                       # float32, like np.log(rms + 1e-6)
        cur = 0
        for dd in d:
            m = np.mean(logc[cur: cur + dd]) if dd else np.float32(np.nan)     # numpy: mean of an empty slice is NaN
            want.append(np.clip((m - (-3.0)) / 1.5, -2.5, 2.5))
            cur += dd
    got = sp.segment_pool(torch.from_numpy(curve).to(cuda), fo, torch.from_numpy(np.concatenate(durs)).to(cuda), po,
                          mu=-3.0, sigma=1.5, lo=-2.5, hi=2.5, log_eps=1e-6).cpu().numpy()
    want = np.array(want, dtype=np.float64)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    assert np.abs(got[ok] - want[ok]).max() <= 2e-6          # float32 sums of up to 100 logs, sequential vs pairwise
