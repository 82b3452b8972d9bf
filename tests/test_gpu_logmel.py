"""GPU parity: fused STFT -> power -> mel -> log kernel (K1) vs the CPU oracle.
Replaces /root/reference/spev_real_metrics.py:363-367; tolerance from BASELINE north_star:
max-abs log-mel error <= 1e-4."""
import numpy as np
import pytest
import torch

from oracle import librosa_restated as lr
from tests import synth

pytestmark = pytest.mark.gpu
TOL_LOGMEL = 1e-4


def test_cfg1_white_and_speechy_vs_oracle(cuda, golden):
    import spev_tts_b200 as sp
    for name, y in (("white", synth.white(seed=0)), ("speechy", synth.speechy(seed=1))):
        ref = lr.reference_logmel(y)                       # [517, 80]
        got = sp.logmel(y)                                 # numpy in -> numpy out
        assert got.shape == ref.shape == (517, 80) and got.dtype == np.float32
        err = np.abs(got - ref).max()
        assert err <= TOL_LOGMEL, f"{name}: max-abs log-mel error {err}"
        gold = golden(f"logmel_{name}.npz")["logmel"]
        assert np.abs(got - gold).max() <= TOL_LOGMEL
    # the speechy case must exercise both clamps
    assert (ref <= -10.0).mean() > 0.1 and ref.max() == 2.0


def test_melspectrogram_dropin_signature_and_power(cuda):
    import spev_tts_b200 as sp
    y = synth.speechy(seed=4, n=50000)
    ref = lr.melspectrogram(y=y, sr=22050, n_fft=1024, hop_length=256, n_mels=80)
    got = sp.melspectrogram(y=y, sr=22050, n_fft=1024, hop_length=256, n_mels=80)
    assert got.shape == ref.shape == (80, 1 + 50000 // 256)
    rel = np.abs(got - ref) / (np.abs(ref) + 1e-6 * ref.max())
    assert rel.max() < 2e-4, rel.max()
    # torch tensor in -> torch tensor out on the device, leading dims broadcast
    yt = torch.from_numpy(np.stack([y, y[::-1].copy()]).reshape(2, 1, -1)).to(cuda)
    gt = sp.melspectrogram(y=yt, sr=22050, n_fft=1024, hop_length=256, n_mels=80)
    assert isinstance(gt, torch.Tensor) and gt.is_cuda and gt.shape == (2, 1, 80, ref.shape[1])
    assert np.allclose(gt[0, 0].cpu().numpy(), got, rtol=0, atol=0)
    with pytest.raises(NotImplementedError):
        sp.melspectrogram(y=y, sr=22050, n_fft=2048, hop_length=512)


def test_inverse_basis_and_sr24k(cuda):
    """basis keyed on (sr, fmin, fmax): forward default fmax=sr/2, cfg5 sr=24000."""
    import spev_tts_b200 as sp
    y = synth.speechy(seed=5, n=24000 * 2, sr=24000)
    for kw in (dict(sr=24000), dict(sr=22050, fmin=0.0, fmax=8000.0)):
        ref = lr.melspectrogram(y=y, n_fft=1024, hop_length=256, n_mels=80, **kw)
        got = sp.melspectrogram(y=y, n_fft=1024, hop_length=256, n_mels=80, **kw)
        rel = np.abs(got - ref) / (np.abs(ref) + 1e-6 * ref.max())
        assert rel.max() < 2e-4
        ctx = sp.Context.get(cuda, n_mels=80, fmin=kw.get("fmin", 0.0), fmax=kw.get("fmax"), sr=kw["sr"])
        ob = lr.mel_filter(n_fft=1024, n_mels=80, **kw)
        assert np.array_equal(ctx.mel_basis(), ob)


@pytest.mark.parametrize("n_mels,sr,fmin,fmax", [(128, 22050, 0.0, None), (40, 16000, 50.0, 7600.0), (96, 24000, 0.0, 8000.0), (32, 22050, 0.0, None)])
def test_other_mel_configurations(cuda, n_mels, sr, fmin, fmax):
    """The fused kernel's mel program is built at ctx creation for any (sr, n_mels, fmin, fmax)."""
    import spev_tts_b200 as sp
    y = synth.speechy(seed=n_mels, n=40000, sr=sr)
    ref = lr.melspectrogram(y=y, sr=sr, n_fft=1024, hop_length=256, n_mels=n_mels, fmin=fmin, fmax=fmax)
    got = sp.melspectrogram(y=y, sr=sr, n_fft=1024, hop_length=256, n_mels=n_mels, fmin=fmin, fmax=fmax)
    assert got.shape == ref.shape == (n_mels, 1 + 40000 // 256)
    rel = np.abs(got - ref) / (np.abs(ref) + 1e-6 * ref.max())
    assert rel.max() < 2e-4, rel.max()
    lm_ref = np.clip(np.log(np.clip(ref, 1e-5, None)), -10, 2).T
    flat, _ = sp.logmel_flat(torch.from_numpy(y).to(cuda), [len(y)], sr=sr, n_mels=n_mels, fmin=fmin, fmax=fmax)
    assert np.abs(flat.cpu().numpy() - lm_ref).max() <= TOL_LOGMEL
    # mel -> magnitude for the same basis (tcgen05 path when n_mels % 4 == 0)
    S_ref = lr.mel_to_stft(ref[:, :64], sr=sr, n_fft=1024, fmin=fmin, fmax=fmax, lbfgs=False)
    S = sp.mel_to_stft(ref[:, :64], sr=sr, n_fft=1024, fmin=fmin, fmax=fmax, nnls="pinv")
    assert np.linalg.norm(S - S_ref) / np.linalg.norm(S_ref) < 2e-5


@pytest.mark.parametrize("n_mels", [4, 12, 20, 160])
def test_few_wide_bands_and_many_narrow_bands(cuda, n_mels):
    """Round 1 rejected bases with very few, very wide bands (the per-warp padded mel program outgrew shared memory).
    The band-major program has no padding: any n_mels works, including n_mels < 16 (warps without a band) and
    n_mels > 128 (run-time band count)."""
    import spev_tts_b200 as sp
    y = synth.white(seed=1, n=20000)
    ref = lr.melspectrogram(y=y, sr=22050, n_fft=1024, hop_length=256, n_mels=n_mels)
    got = sp.melspectrogram(y=y, sr=22050, n_fft=1024, hop_length=256, n_mels=n_mels)
    assert got.shape == ref.shape
    lg, lf = np.log(np.clip(got, 1e-5, None)), np.log(np.clip(ref, 1e-5, None))
    assert np.abs(lg - lf).max() <= TOL_LOGMEL


def test_ragged_batch_edges(cuda):
    """empty / sub-hop / exact-multiple / odd lengths, aligned and unaligned packing."""
    import spev_tts_b200 as sp
    rng = np.random.default_rng(11)
    lens = [0, 1, 255, 256, 257, 1023, 1024, 8191, 8192, 33 * 256 + 5, 70001, 3, 64 * 256]
    ys = [(0.1 * rng.standard_normal(n)).astype(np.float32) for n in lens]
    flat = torch.from_numpy(np.concatenate(ys)).to(cuda)
    out, fb = sp.logmel_flat(flat, lens)
    out = out.cpu().numpy()
    assert fb.n_frames == sum(1 + n // 256 for n in lens) and out.shape == (fb.n_frames, 80)
    for i, y in enumerate(ys):
        ref = lr.reference_logmel(y)
        got = out[fb.frame_off[i]: fb.frame_off[i + 1]]
        assert got.shape == ref.shape
        assert np.abs(got - ref).max() <= TOL_LOGMEL, (i, lens[i])
    # aligned packing (every item start % 4 == 0) takes the 16-byte cp.async path; the filler
    # between items is garbage on purpose: explicit [lo, hi) bounds keep it out of the signal
    from spev_tts_b200 import cache
    starts = cache.aligned_offsets(lens)
    buf = np.full(starts[-1] + 8, 1e6, dtype=np.float32)
    for s, y in zip(starts[:-1], ys):
        buf[s: s + len(y)] = y
    out2, _ = sp.logmel_flat(torch.from_numpy(buf).to(cuda), lens, sample_off=starts)
    assert np.array_equal(out2.cpu().numpy(), out)


def test_pipelined_host_builder_matches_device_path(cuda):
    """LogMelCacheBuilder (pinned host in -> chunked H2D/kernel/D2H -> pinned host out) must give
    exactly the one-launch device result, for aligned and back-to-back packing."""
    import spev_tts_b200 as sp
    from spev_tts_b200 import cache
    lens = synth.utterance_lengths(seed=14, n_utts=300)
    rng = np.random.default_rng(14)
    ys = [(0.05 * rng.standard_normal(int(n))).astype(np.float32) for n in lens]
    flat = torch.from_numpy(np.concatenate(ys))
    ref, fb = sp.logmel_flat(flat.to(cuda), lens)
    out, fo = cache.build_logmel_cache(flat.pin_memory(), lens, device=cuda, chunk_samples=1 << 20)
    assert np.array_equal(fo, fb.frame_off) and torch.equal(out.to(cuda), ref)
    starts = cache.aligned_offsets(lens)
    buf = torch.full((int(starts[-1]),), 7.0)
    for s, y in zip(starts[:-1], ys):
        buf[s: s + len(y)] = torch.from_numpy(y)
    out2, _ = cache.build_logmel_cache(buf.pin_memory(), lens, device=cuda, chunk_samples=1 << 20,
                                       sample_off=starts)
    assert torch.equal(out2, out)


def test_pcm16_host_path_is_exact(cuda):
    """int16 PCM shipped over PCIe and widened on the device == float32(pcm / 32768) shipped directly."""
    import spev_tts_b200 as sp
    from spev_tts_b200 import cache
    lens = synth.utterance_lengths(seed=15, n_utts=120)
    rng = np.random.default_rng(15)
    starts = cache.aligned_offsets(lens)
    pcm = rng.integers(-32768, 32768, int(starts[-1]), dtype=np.int16)
    f32 = (pcm.astype(np.float32) / 32768.0)
    out_f, _ = cache.build_logmel_cache(torch.from_numpy(f32).pin_memory(), lens, device=cuda, sample_off=starts,
                                        chunk_samples=1 << 20)
    out_p, _ = cache.build_logmel_cache(torch.from_numpy(pcm).pin_memory(), lens, device=cuda, sample_off=starts,
                                        chunk_samples=1 << 20)
    assert torch.equal(out_f, out_p)
    ref = lr.reference_logmel(f32[starts[3]: starts[3] + lens[3]])
    fo = np.concatenate([[0], np.cumsum(1 + lens // 256)])
    assert np.abs(out_p[fo[3]: fo[4]].numpy() - ref).max() <= TOL_LOGMEL


def test_cfg4_scale_properties(cuda):
    """Full-size shaped run (device-generated, ~0.6 M frames here; bench.py runs all 6.2 M):
    batch result == per-utterance result (checksum of checksums), spot checks vs oracle,
    determinism."""
    import spev_tts_b200 as sp
    lens = synth.utterance_lengths(seed=4, n_utts=1310)
    g = torch.Generator(device=cuda).manual_seed(4)
    flat = torch.randn(int(lens.sum()), generator=g, device=cuda) * 0.05
    out, fb = sp.logmel_flat(flat, lens)
    out_b, _ = sp.logmel_flat(flat, lens)
    assert torch.equal(out, out_b)
    off = np.concatenate([[0], np.cumsum(lens)])
    for i in (0, 7, 500, 1309):
        y = flat[off[i]: off[i + 1]]
        single, _ = sp.logmel_flat(y.contiguous(), [int(lens[i])])
        seg = out[fb.frame_off[i]: fb.frame_off[i + 1]]
        assert torch.equal(single, seg)                      # batching does not change results
        ref = lr.reference_logmel(y.cpu().numpy())
        assert np.abs(seg.cpu().numpy() - ref).max() <= TOL_LOGMEL
    assert torch.isfinite(out).all() and out.min() >= -10 and out.max() <= 2
