"""GPU parity of the pYIN kernels (librosa.pyin at spev_real_metrics.py:369) against
oracle/pyin_restated.py, stage by stage so that each stage is held to the tightest bar it admits:

  tables   : log-transition / bin frequencies / beta weights            -> 1e-12
  stage 1  : CMND curve (float32 direct sums vs the float64 definition)  -> 1e-3 abs, 2e-5 median
  stage 2  : observation probabilities from the SAME float32 CMND curve  -> 1e-6
  stage 3  : Viterbi on the SAME log-probabilities                       -> state paths IDENTICAL
  whole    : physical known answers + frame agreement with the oracle    -> >= 98 %
"""
import numpy as np
import pytest
import torch

from oracle import pyin_restated as po
from tests import synth

pytestmark = pytest.mark.gpu
SR = 22050
CFG = po.PyinConfig()


def _pctx(cuda):
    from spev_tts_b200.pitch import PyinContext
    return PyinContext.get(cuda)


def _signals():
    ys = [synth.voiced_unvoiced(seed=1, n=2 * SR)[0], synth.speechy(seed=5, n=SR + 77),
          synth.white(seed=2, n=SR // 2), np.zeros(3000, np.float32)]
    return ys


def test_model_tables(cuda):
    p = _pctx(cuda)
    assert (p.n_bins, p.min_period, p.max_period, p.n_lags) == (CFG.n_pitch_bins, CFG.min_period, CFG.max_period, 325)
    lt, fr, bp = p.host_tables()
    want = np.log(CFG.transition() + po.TINY64)
    np.testing.assert_allclose(lt, want, rtol=0, atol=1e-12)
    np.testing.assert_allclose(fr, CFG.freqs, rtol=1e-14)
    np.testing.assert_array_equal(p.freqs64, CFG.freqs)          # what numpy callers get: librosa's exact doubles
    np.testing.assert_array_equal(bp, CFG.beta_probs)            # the shim hands scipy's (= librosa's) table down
    # the library's built-in closed form (C callers without scipy): equal to double rounding of the CDF
    import ctypes as C
    from spev_tts_b200 import _lib
    lib, h = _lib.load(), C.c_void_p()
    _lib.check(lib.spev_pyin_create(C.byref(h), cuda.index or 0, SR, 256, 60.0, 500.0, None))
    bp2 = np.empty(100)
    _lib.check(lib.spev_pyin_host_tables(h, None, None, bp2.ctypes.data))
    lib.spev_pyin_destroy(h)
    np.testing.assert_allclose(bp2, CFG.beta_probs, rtol=0, atol=5e-15)


def _gpu_cmnd(cuda, y):
    from spev_tts_b200 import pitch as gp
    from spev_tts_b200.batch import Context, make_batch
    t = torch.from_numpy(y).to(cuda)
    fb = make_batch(Context.get(cuda), n_samples=[len(y)])
    return gp.cmnd_flat(t, fb, _pctx(cuda)), fb


def test_stage1_cmnd_vs_definition(cuda):
    for y in _signals():
        yin, _ = _gpu_cmnd(cuda, y)
        frames = po.frame_signal(y.astype(np.float64))
        want = po.cmnd(frames, 2048, 1024, CFG.min_period, CFG.max_period).T       # float64 [T, 325]
        got = yin.cpu().numpy()
        assert got.shape == want.shape
        if not y.any():
            assert np.all(got == 0)                      # the reference's |.| < 1e-6 -> 0 clean-up: silence -> 0 curve
            continue
        err = np.abs(got - want)
        assert err.max() <= 1e-3, err.max()
        assert np.median(err) <= 2e-5, np.median(err)


def test_stage2_observation_vs_oracle(cuda):
    from spev_tts_b200 import pitch as gp
    p = _pctx(cuda)
    for y in _signals()[:3]:
        yin, _ = _gpu_cmnd(cuda, y)
        # the oracle consumes the very same curve, as float64 like librosa's own CMND array (its cumulative mean
        # divides by an int64 lag vector, so the array is float64 even for float32 audio)
        y64 = yin.cpu().numpy().T.astype(np.float64)
        obs, vp = po.observation_probs(y64, po.parabolic_interpolation(y64), CFG)
        logobs, lunv, vprob = gp.observe(yin, p)
        got = np.exp(logobs.double().cpu().numpy()).T    # [n_bins, T]
        np.testing.assert_allclose(got, obs[: CFG.n_pitch_bins], rtol=1e-4, atol=1e-7)
        np.testing.assert_allclose(vprob.cpu().numpy(), vp, atol=2e-6)
        # (1 - voiced_prob) cancels to a few ulps when the voiced mass sums to 1: absolute tolerance on the probability
        np.testing.assert_allclose(np.exp(lunv.double().cpu().numpy()), obs[CFG.n_pitch_bins], rtol=1e-4, atol=1e-15)
        # the sparsity pattern (which bins carry mass) must be identical
        assert np.array_equal(got > 1e-300, obs[: CFG.n_pitch_bins] > 0)


def _oracle_decode(logobs, lunv, log_trans=None):
    """The oracle's Viterbi recursion on the same float32 log-observations.  log_trans: the library's own table
    (held to 1e-12 of the oracle's by test_model_tables) when exact ties between mirror-image unvoiced paths
    must resolve identically; None = the oracle's table."""
    lo = logobs.double().cpu().numpy()
    lu = lunv.double().cpu().numpy()
    log_prob = np.concatenate([lo, np.repeat(lu[:, None], CFG.n_pitch_bins, 1)], axis=1)     # [T, 736]
    lt = np.log(CFG.transition() + po.TINY64) if log_trans is None else log_trans
    return po.viterbi_log(log_prob, lt, np.log(CFG.p_init() + po.TINY64))


def test_stage3_viterbi_identical_paths(cuda):
    from spev_tts_b200 import pitch as gp
    p = _pctx(cuda)
    for y in _signals():
        yin, fb = _gpu_cmnd(cuda, y)
        logobs, lunv, _ = gp.observe(yin, p)
        states, f0, flag = gp.decode(logobs, lunv, fb.frame_off, p)
        want = _oracle_decode(logobs, lunv)
        np.testing.assert_array_equal(states.cpu().numpy(), want)
        f0 = f0.cpu().numpy()
        assert np.array_equal(flag.cpu().numpy(), want < CFG.n_pitch_bins)
        assert np.isnan(f0[want >= CFG.n_pitch_bins]).all()
        v = want < CFG.n_pitch_bins
        np.testing.assert_allclose(f0[v], CFG.freqs[want[v]], rtol=1e-6)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_stage3_viterbi_random_sparse_observations(cuda, seed):
    """Adversarial HMM input: isolated observation peaks that jump further than the transition band, so the
    winning predecessor is often OUTSIDE the band (the dense reference pays log(tiny) for it)."""
    from spev_tts_b200 import pitch as gp
    p = _pctx(cuda)
    rng = np.random.default_rng(seed)
    T, nb = 90, CFG.n_pitch_bins
    obs = np.zeros((T, nb))
    for t in range(T):
        if t < 40:                                      # certain-voiced frames whose single peak hops anywhere:
            obs[t, rng.integers(0, nb)] = 1.0           # unvoiced states cost log(tiny) too, so hopping wins
            continue
        for _ in range(int(rng.integers(0, 4))):
            obs[t, rng.integers(0, nb)] = rng.uniform(0.05, 0.5)
    vp = np.clip(obs.sum(1), 0, 1)
    logobs = torch.from_numpy(np.log(obs + po.TINY64).astype(np.float32)).to(cuda)
    lunv = torch.from_numpy(np.log((1 - vp) / nb + po.TINY64).astype(np.float32)).to(cuda)
    fo = np.array([0, 40, 41, T])                       # three items: 40, 1 and 49 frames
    states, _, _ = gp.decode(logobs, lunv, fo, p)
    got = states.cpu().numpy()
    lt = p.host_tables()[0]                              # uniform unvoiced stretches tie to the last bit: share the table
    for a, b in zip(fo[:-1], fo[1:]):
        np.testing.assert_array_equal(got[a:b], _oracle_decode(logobs[a:b], lunv[a:b], lt))
    jumps = np.abs(np.diff(got[:40] % nb))
    assert jumps.max() > 25                              # the out-of-band branch was really exercised


@pytest.mark.parametrize("f0", [82.4, 110.0, 220.0, 333.0, 440.0])
def test_known_answer_tones(cuda, f0):
    import spev_tts_b200 as sp
    t = np.arange(SR // 2) / SR
    y = (0.3 * sum(np.sin(2 * np.pi * f0 * (h + 1) * t) / (h + 1) for h in range(5))).astype(np.float32)
    y += 1e-3 * np.random.default_rng(0).standard_normal(len(y)).astype(np.float32)
    f, flag, vp = sp.pyin(y, fmin=60, fmax=500, sr=SR, hop_length=256)
    assert f.shape == flag.shape == vp.shape == (1 + len(y) // 256,)
    assert flag[4:-4].all() and (vp[4:-4] > 0.5).all()
    assert np.abs(1200 * np.log2(f[4:-4] / f0)).max() <= 10.0


def test_end_to_end_agreement_with_oracle(cuda):
    import spev_tts_b200 as sp
    tot = agree = 0
    for seed in (1, 2, 3):
        y, f_true = synth.voiced_unvoiced(seed=seed, n=3 * SR)
        f, flag, vp = sp.pyin(y, fmin=60, fmax=500, sr=SR, hop_length=256)
        fo, flago, vpo = po.pyin(y)
        tot += len(f)
        both = flag & flago
        same_f0 = np.zeros(len(f), bool)
        same_f0[both] = np.abs(1200 * np.log2(f[both] / fo[both])) <= 10.0 + 1e-6       # within one bin
        agree += int(np.sum((flag == flago) & (same_f0 | ~flago)))
        assert np.abs(vp - vpo).mean() <= 2e-3
        # numpy in -> float64 out, read from the same float64 bin table as librosa: bit-identical where the state is
        assert f.dtype == np.float64 and vp.dtype == np.float64 and flag.dtype == np.bool_
        assert np.mean(f[both] == fo[both]) >= 0.98
        assert np.array_equal(flag, flago) or np.mean(flag == flago) >= 0.98
        # and against the ground truth of the synthetic signal (frames well inside voiced segments)
        idx = np.clip(np.arange(len(f)) * 256 - 500, 0, len(y) - 1)     # centre of the analysed span (see test_oracle_pyin)
        inner = np.array([f_true[max(0, i - 1500): i + 1500].min() > 0 for i in idx]) & (f_true[idx] < 480) & (f_true[idx] > 64)
        assert flag[inner].mean() >= 0.97
        ok = flag & inner
        assert np.median(np.abs(1200 * np.log2(f[ok] / f_true[idx][ok]))) <= 6.0
    assert agree / tot >= 0.98, agree / tot


def test_ragged_batch_equals_single_calls(cuda):
    from spev_tts_b200 import pitch as gp
    lens = [0, 100, 2048, 256 * 33 + 1, 30000, 255]
    ys = [synth.voiced_unvoiced(seed=10 + i, n=max(n, 1))[0][:n] for i, n in enumerate(lens)]
    flat = torch.from_numpy(np.concatenate(ys)).to(cuda)
    f0, flag, vp, fo, st = gp.pyin_flat(flat, lens, return_states=True)
    assert fo[-1] == sum(1 + n // 256 for n in lens)
    st = st.cpu().numpy()
    for i, y in enumerate(ys):
        sl = slice(fo[i], fo[i + 1])
        if len(y) == 0:
            assert st[sl].tolist() and (st[sl] >= CFG.n_pitch_bins).all()        # one frame of silence: unvoiced
            continue
        _, _, vp1, _, st1 = gp.pyin_flat(torch.from_numpy(y).to(cuda), [len(y)], return_states=True)
        np.testing.assert_array_equal(st[sl], st1.cpu().numpy())
        assert torch.equal(vp[sl], vp1)


def test_unsupported_parameters_raise(cuda):
    import spev_tts_b200 as sp
    y = np.zeros(4096, np.float32)
    with pytest.raises(NotImplementedError):
        sp.pyin(y, fmin=60, fmax=500, sr=SR, hop_length=128)
    with pytest.raises(NotImplementedError):
        sp.pyin(y, fmin=60, fmax=500, sr=SR, hop_length=256, frame_length=1024)
    with pytest.raises(RuntimeError):
        sp.pyin(y, fmin=20, fmax=500, sr=SR, hop_length=256)      # max_period beyond what the kernels hold


def test_default_hop_512_of_the_statistics_pass(cuda):
    """spev_real_metrics.py:311 calls librosa.pyin WITHOUT hop_length (-> frame_length // 4 = 512): frames every
    512 samples and a 101-bin transition band."""
    import spev_tts_b200 as sp
    from spev_tts_b200 import pitch as gp
    cfg = po.PyinConfig(hop_length=512)
    assert cfg.transition_width == 101
    p = gp.PyinContext.get(cuda, hop=512)
    lt, _, _ = p.host_tables()
    np.testing.assert_allclose(lt, np.log(cfg.transition() + po.TINY64), rtol=0, atol=1e-12)
    tot = agree = 0
    for seed in (4, 5):
        y, _ = synth.voiced_unvoiced(seed=seed, n=2 * SR + 300)
        f, flag, vp = sp.pyin(y, fmin=60, fmax=500, sr=SR)                  # hop_length=None -> 512
        fo, flago, vpo = po.pyin(y, hop_length=512)
        assert f.shape == fo.shape == (1 + len(y) // 512,)
        both = flag & flago
        same = np.zeros(len(f), bool)
        same[both] = np.abs(1200 * np.log2(f[both] / fo[both])) <= 10.0 + 1e-6
        agree += int(np.sum((flag == flago) & (same | ~flago)))
        tot += len(f)
        assert np.abs(vp - vpo).mean() <= 2e-3
    assert agree / tot >= 0.98, agree / tot
    # identical paths on identical log-observations (the wide table lives in global memory: separate code path)
    y, _ = synth.voiced_unvoiced(seed=6, n=SR)
    t = torch.from_numpy(y).to(cuda)
    from spev_tts_b200.batch import Context, make_batch
    fb = make_batch(Context.get(cuda), n_samples=[len(y)])
    yin = gp.cmnd_flat(t, fb, p)[::2].contiguous()
    logobs, lunv, _ = gp.observe(yin, p)
    states, _, _ = gp.decode(logobs, lunv, np.array([0, yin.shape[0]]), p)
    lo, lu = logobs.double().cpu().numpy(), lunv.double().cpu().numpy()
    log_prob = np.concatenate([lo, np.repeat(lu[:, None], cfg.n_pitch_bins, 1)], axis=1)
    want = po.viterbi_log(log_prob, lt, np.log(cfg.p_init() + po.TINY64))
    np.testing.assert_array_equal(states.cpu().numpy(), want)


@pytest.mark.parametrize("sr", [16000, 24000])
def test_other_sample_rates(cuda, sr):
    """cfg5 runs at 24 kHz (LibriTTS-R): max_period = 400 lags -> a wider CMND block; 16 kHz -> a narrower one."""
    import spev_tts_b200 as sp
    cfg = po.PyinConfig(sr=sr)
    y, _ = synth.voiced_unvoiced(seed=7, n=int(1.5 * sr), sr=sr)
    f, flag, vp = sp.pyin(y, fmin=60, fmax=500, sr=sr, hop_length=256)
    fo, flago, vpo = po.pyin(y, sr=sr)
    assert f.shape == fo.shape
    both = flag & flago
    same = np.zeros(len(f), bool)
    same[both] = np.abs(1200 * np.log2(f[both] / fo[both])) <= 10.0 + 1e-6
    agree = np.mean((flag == flago) & (same | ~flago))
    assert agree >= 0.97, agree
    assert np.abs(vp - vpo).mean() <= 3e-3


def test_fill_na_variants_and_batched_input(cuda):
    import spev_tts_b200 as sp
    y, _ = synth.voiced_unvoiced(seed=9, n=SR)
    f_nan, flag, _ = sp.pyin(y, fmin=60, fmax=500, sr=SR, hop_length=256)
    f_zero, flag2, _ = sp.pyin(y, fmin=60, fmax=500, sr=SR, hop_length=256, fill_na=0.0)
    f_best, flag3, _ = sp.pyin(y, fmin=60, fmax=500, sr=SR, hop_length=256, fill_na=None)
    assert np.array_equal(flag, flag2) and np.array_equal(flag, flag3) and (~flag).any() and flag.any()
    assert np.isnan(f_nan[~flag]).all() and np.all(f_zero[~flag] == 0.0)
    assert np.array_equal(f_nan[flag], f_zero[flag]) and np.array_equal(f_nan[flag], f_best[flag])
    assert np.all((f_best[~flag] >= 60.0) & (f_best[~flag] <= 500.0))          # the decoded bin of the unvoiced state
    fo, _, _ = po.pyin(y, fill_na=None)
    assert np.mean(np.abs(1200 * np.log2(f_best / fo)) <= 10.0 + 1e-6) >= 0.97
    # leading batch dimensions, torch in -> torch out on the same device
    yy = torch.from_numpy(np.stack([y, y[::-1].copy()])).to(cuda)
    fb, flb, vpb = sp.pyin(yy, fmin=60, fmax=500, sr=SR, hop_length=256)
    assert fb.shape == (2, 1 + SR // 256) and fb.is_cuda and flb.dtype == torch.bool
    assert np.array_equal(flb[0].cpu().numpy(), flag)


@pytest.mark.parametrize("seed", [11, 12, 13, 14, 15])
def test_randomised_agreement_sweep(cuda, seed):
    """More signals (different voicing patterns, glide ranges, harmonic counts, levels): flags and bins vs the oracle."""
    import spev_tts_b200 as sp
    rng = np.random.default_rng(seed)
    y, _ = synth.voiced_unvoiced(seed=seed, n=int(rng.uniform(1.0, 2.0) * SR))
    y = (y * rng.uniform(0.05, 1.5)).astype(np.float32)            # level must not matter (CMND is scale-invariant)
    f, flag, vp = sp.pyin(y, fmin=60, fmax=500, sr=SR, hop_length=256)
    fo, flago, vpo = po.pyin(y)
    both = flag & flago
    assert np.mean(flag == flago) >= 0.97
    assert both.sum() == 0 or np.mean(np.abs(1200 * np.log2(f[both] / fo[both])) <= 10.0 + 1e-6) >= 0.97
    assert np.abs(vp - vpo).mean() <= 3e-3
