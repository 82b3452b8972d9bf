"""Pins for the pYIN restatement (oracle/pyin_restated.py).  The reference's librosa is not installable and
ships no pYIN vectors, so parity is UNPINNED BY THE REFERENCE; these tests pin the restatement against
independent brute-force definitions and physical known answers instead."""
import numpy as np
import pytest

from oracle import pyin_restated as po

SR = 22050


def tone(f0, n=SR, harmonics=5, seed=0, noise=1e-3):
    t = np.arange(n) / SR
    y = sum(np.sin(2 * np.pi * f0 * (h + 1) * t) / (h + 1) for h in range(harmonics))
    return (0.3 * y + noise * np.random.default_rng(seed).standard_normal(n)).astype(np.float32)


def glide(f_a, f_b, n=SR, seed=0):
    f = np.linspace(f_a, f_b, n)
    ph = 2 * np.pi * np.cumsum(f) / SR
    y = sum(np.sin((h + 1) * ph) / (h + 1) for h in range(4))
    return (0.3 * y + 1e-3 * np.random.default_rng(seed).standard_normal(n)).astype(np.float32), f


def test_config_matches_reference_call():
    cfg = po.PyinConfig()                                   # fmin=60, fmax=500, sr=22050, hop 256
    assert (cfg.min_period, cfg.max_period) == (44, 368)
    assert cfg.n_pitch_bins == 368 and cfg.transition_width == 51 and cfg.n_bins_per_semitone == 10
    assert np.isclose(cfg.beta_probs.sum(), 1.0) and len(cfg.beta_probs) == 100
    assert np.isclose(cfg.freqs[0], 60.0) and cfg.freqs[-1] <= 500.0


def test_cmnd_against_bruteforce_definition():
    rng = np.random.default_rng(3)
    cfg = po.PyinConfig()
    frames = rng.standard_normal((2048, 3))
    frames[:, 1] = np.sin(2 * np.pi * 150 * np.arange(2048) / SR)
    got = po.cmnd(frames, 2048, 1024, cfg.min_period, cfg.max_period)
    for c in range(3):
        d = po.difference_function_bruteforce(frames[:, c], 1024, cfg.max_period)
        cm = np.cumsum(d[1:]) / np.arange(1, cfg.max_period + 1)
        want = d[cfg.min_period:] / cm[cfg.min_period - 1:]
        np.testing.assert_allclose(got[:, c], want, rtol=1e-9, atol=1e-10)


def test_parabolic_interpolation_recovers_vertex():
    i = np.arange(9, dtype=np.float64)
    x = ((i - 4.3) ** 2)[:, None]
    s = po.parabolic_interpolation(x)
    assert np.isclose(s[4, 0], 0.3) and s[0, 0] == 0 and s[-1, 0] == 0
    flat = po.parabolic_interpolation(np.ones((5, 1)))
    assert np.all(flat == 0)


def test_transition_matrices():
    cfg = po.PyinConfig()
    tl = po.transition_local(cfg.n_pitch_bins, cfg.transition_width)
    np.testing.assert_allclose(tl.sum(1), 1.0, atol=1e-12)
    i, j = np.nonzero(tl)
    assert np.abs(i - j).max() == 25                         # band-limited, no wrap-around
    assert tl[0, 0] == tl[0].max() and tl[200, 200] == tl[200].max()
    np.testing.assert_allclose(tl[100, 75:126], tl[200, 175:226])   # interior rows share one shape
    sw = po.transition_loop(2, 0.99)
    np.testing.assert_allclose(sw, [[0.99, 0.01], [0.01, 0.99]], atol=1e-15)
    t = cfg.transition()
    assert t.shape == (736, 736)
    np.testing.assert_allclose(t.sum(1), 1.0, atol=1e-12)


@pytest.mark.parametrize("seed", range(6))
def test_viterbi_against_exhaustive_search(seed):
    rng = np.random.default_rng(seed)
    n, T = 4, 6
    trans = rng.random((n, n)); trans /= trans.sum(1, keepdims=True)
    prob = rng.random((n, T))
    p0 = rng.random(n); p0 /= p0.sum()
    np.testing.assert_array_equal(po.viterbi(prob, trans, p0), po.viterbi_bruteforce(prob, trans, p0))


@pytest.mark.parametrize("f0", [82.4, 110.0, 220.0, 333.0, 440.0])
def test_known_answer_tones(f0):
    f, flag, vp = po.pyin(tone(f0, n=SR // 2))
    assert len(f) == 1 + (SR // 2) // 256
    mid = slice(4, -4)                                       # edge frames see the zero padding
    assert flag[mid].all() and (vp[mid] > 0.5).all()
    cents = 1200 * np.log2(f[mid] / f0)
    assert np.abs(cents).max() <= 10.0, cents                # one 10-cent bin


def test_known_answer_glide():
    y, f_true = glide(120.0, 240.0, n=SR)
    f, flag, _ = po.pyin(y)
    idx = np.arange(len(f))[6:-6]
    # the analysed span of frame t is samples 1 .. 1024+tau of the 2048 window: centred ~(1024 - tau)/2 = 400-500
    # samples BEFORE t*hop
    want = f_true[np.clip(idx * 256 - 470, 0, len(f_true) - 1)]
    assert flag[6:-6].all()
    assert np.abs(1200 * np.log2(f[6:-6] / want)).max() <= 10.0


def test_noise_and_silence_are_unvoiced():
    rng = np.random.default_rng(0)
    f, flag, vp = po.pyin(rng.standard_normal(SR // 2).astype(np.float32) * 0.1)
    assert flag.mean() < 0.1 and np.isnan(f[~flag]).all()
    f, flag, vp = po.pyin(np.zeros(SR // 4, dtype=np.float32))
    assert not flag.any() and np.all(vp == 0) and np.isnan(f).all()


def test_voicing_switch():
    y = np.concatenate([np.zeros(SR // 4, np.float32), tone(200.0, n=SR // 2), np.zeros(SR // 4, np.float32)])
    f, flag, _ = po.pyin(y)
    t = np.arange(len(f)) * 256
    inside = (t > SR // 4 + 2048) & (t < 3 * SR // 4 - 2048)
    outside = (t < SR // 4 - 2048) | (t > 3 * SR // 4 + 2048)
    assert flag[inside].all() and not flag[outside].any()
    assert np.abs(1200 * np.log2(f[inside] / 200.0)).max() <= 10.0
