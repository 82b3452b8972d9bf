"""Out-of-bounds write detection without compute-sanitizer (closed on this pool): every output
buffer handed to the C ABI sits between canary regions that must survive the launch, for odd and
ragged sizes."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
CANARY = 1.2345678e30
PAD = 4096


class Guarded:
    """float32 buffer of n elements between two PAD-element canary regions (16-byte aligned payload)."""

    def __init__(self, n, dev, dtype=torch.float32, words_per_elem=1):
        self.n = n * words_per_elem
        self.raw = torch.full((self.n + 2 * PAD,), CANARY, dtype=torch.float32, device=dev)
        self.view = self.raw[PAD: PAD + self.n]
        self.view.fill_(float("nan"))

    def ptr(self):
        return self.view.data_ptr()

    def check(self, what):
        head, tail = self.raw[:PAD], self.raw[PAD + self.n:]
        assert bool((head == CANARY).all()) and bool((tail == CANARY).all()), f"{what}: canary overwritten"
        return self.view


@pytest.mark.parametrize("lens", [[1], [255, 256, 257], [8191, 8192, 8193, 5], [33 * 256 + 5, 0, 70001]])
def test_forward_kernels_stay_in_bounds(cuda, lens):
    import spev_tts_b200 as sp
    from spev_tts_b200 import _lib
    rng = np.random.default_rng(len(lens))
    flat = torch.from_numpy(np.concatenate([(0.1 * rng.standard_normal(n)).astype(np.float32) for n in lens])).to(cuda) \
        if sum(lens) else torch.zeros(1, device=cuda)
    ctx = sp.Context.get(cuda)
    fb = sp.make_batch(ctx, n_samples=lens)
    st = torch.cuda.current_stream(cuda).cuda_stream
    F = fb.n_frames
    out = Guarded(F * 80, cuda)
    _lib.check(ctx.lib.spev_logmel(ctx.handle, fb.desc, flat.data_ptr(), out.ptr(), 1, 1e-5, -10.0, 2.0, st))
    v = out.check("spev_logmel")
    assert bool(torch.isfinite(v).all())                      # every element written
    pw = Guarded(F * 520, cuda)
    _lib.check(ctx.lib.spev_stft_power(ctx.handle, fb.desc, flat.data_ptr(), pw.ptr(), st))
    assert bool(torch.isfinite(pw.check("spev_stft_power")).all())
    mo = Guarded(F * 80, cuda)
    _lib.check(ctx.lib.spev_mel_project(ctx.handle, pw.ptr(), F, mo.ptr(), 1, 1e-5, -10.0, 2.0, st))
    assert bool(torch.isfinite(mo.check("spev_mel_project")).all())
    r, c = Guarded(F, cuda), Guarded(F, cuda)
    _lib.check(ctx.lib.spev_frame_features(ctx.handle, fb.desc, flat.data_ptr(), r.ptr(), c.ptr(), st))
    assert bool(torch.isfinite(r.check("rms")).all()) and bool(torch.isfinite(c.check("centroid")).all())


@pytest.mark.parametrize("frames", [[1], [2, 3], [29, 30, 31, 32, 33], [100, 1, 64]])
def test_griffinlim_kernels_stay_in_bounds(cuda, frames):
    import spev_tts_b200 as sp
    from spev_tts_b200 import _lib
    ctx = sp.Context.get(cuda, fmin=0.0, fmax=8000.0)
    fb = sp.make_batch(ctx, n_frames=frames, with_chunks=True)
    st = torch.cuda.current_stream(cuda).cuda_stream
    F = fb.n_frames
    g = torch.Generator(device=cuda).manual_seed(1)
    lm = (-4 + 2 * torch.randn(F, 80, generator=g, device=cuda)).clamp(-10, 2)
    S = Guarded(F * 520, cuda)
    for tc in (1, 0):                                           # tcgen05 and FFMA variants
        ctx.set_tensor_core(bool(tc))
        _lib.check(ctx.lib.spev_mel_to_mag(ctx.handle, fb.desc, lm.data_ptr(), 0, 1, S.ptr(), 520, st))
        Sv = S.check("spev_mel_to_mag")
        assert bool(torch.isfinite(Sv.view(F, 520)[:, :513]).all())
    ctx.set_tensor_core(True)
    y = Guarded(max(fb.n_out_samples, 1), cuda)
    ws_bytes = ctx.lib.spev_griffinlim_workspace_bytes(F)
    ws = Guarded((ws_bytes + 3) // 4, cuda)
    _lib.check(ctx.lib.spev_griffinlim(ctx.handle, fb.desc, S.ptr(), 520, None, 7, 3, 0.99, y.ptr(), ws.ptr(), ws_bytes, st))
    yv = y.check("spev_griffinlim y")
    ws.check("spev_griffinlim workspace")
    if fb.n_out_samples:
        assert bool(torch.isfinite(yv[: fb.n_out_samples]).all())


def test_lr_and_bucketize_stay_in_bounds(cuda):
    import spev_tts_b200 as sp
    from spev_tts_b200 import _lib
    lib = sp.load()
    st = torch.cuda.current_stream(cuda).cuda_stream
    B, T, H = 5, 37, 13
    g = torch.Generator(device=cuda).manual_seed(2)
    x = torch.randn(B, T, H, generator=g, device=cuda)
    d = torch.randint(0, 9, (B, T), generator=g, device=cuda)
    p = sp.plan(d)
    out = Guarded(B * p.max_len * H, cuda)
    fo = Guarded(5 * B * p.max_len, cuda)
    feats = torch.randn(5, B, T, generator=g, device=cuda)
    _lib.check(lib.spev_lr_expand_fused(x.data_ptr(), H * 4, feats.data_ptr(), 5, None, None, p.cumsum.data_ptr(), B, T,
                                        out.ptr(), fo.ptr(), p.max_len, st))
    assert bool(torch.isfinite(out.check("lr out")).all()) and bool(torch.isfinite(fo.check("lr feats")).all())
    n, Hh = 77, 10
    v = torch.randn(n, generator=g, device=cuda)
    bnd = torch.linspace(-3, 3, 255, device=cuda)
    tab = torch.randn(256, Hh, generator=g, device=cuda)
    emb = Guarded(n * Hh, cuda)
    _lib.check(lib.spev_bucketize_embed(v.data_ptr(), n, bnd.data_ptr(), 255, 0, tab.data_ptr(), Hh, None, emb.ptr(), 0, st))
    assert bool(torch.isfinite(emb.check("bucketize")).all())


@pytest.mark.parametrize("lens", [[1], [255, 256, 257], [8191, 8192, 8193, 5], [33 * 256 + 5, 0, 70001]])
def test_pyin_kernels_stay_in_bounds(cuda, lens):
    """Every pYIN stage writes all of its output and nothing else (guarded buffers, ragged sizes)."""
    import spev_tts_b200 as sp
    from spev_tts_b200 import _lib
    from spev_tts_b200.pitch import PyinContext
    from tests import synth
    ys = [synth.voiced_unvoiced(seed=3 + i, n=max(n, 1))[0][:n] for i, n in enumerate(lens)]
    flat = torch.from_numpy(np.concatenate(ys)).to(cuda) if sum(lens) else torch.zeros(1, device=cuda)
    ctx = sp.Context.get(cuda)
    p = PyinContext.get(cuda)
    fb = sp.make_batch(ctx, n_samples=lens)
    st = torch.cuda.current_stream(cuda).cuda_stream
    F = fb.n_frames
    yin = Guarded(F * p.n_lags, cuda)
    _lib.check(p.lib.spev_pyin_cmnd(p.handle, fb.desc, flat.data_ptr(), yin.ptr(), st))
    assert bool(torch.isfinite(yin.check("spev_pyin_cmnd")).all())
    lo, lu, vp = Guarded(F * p.n_bins, cuda), Guarded(F, cuda), Guarded(F, cuda)
    _lib.check(p.lib.spev_pyin_observe(p.handle, yin.ptr(), F, lo.ptr(), lu.ptr(), vp.ptr(), st))
    for g, name in ((lo, "logobs"), (lu, "log_unvoiced"), (vp, "voiced_prob")):
        assert bool(torch.isfinite(g.check(name)).all())
    nbytes = p.lib.spev_pyin_decode_workspace_bytes(p.handle, F)
    ws = Guarded((nbytes + 3) // 4, cuda)
    states, f0, flag = Guarded(F, cuda), Guarded(F, cuda), Guarded((F + 3) // 4, cuda)
    fo = torch.from_numpy(fb.frame_off).to(cuda)
    _lib.check(p.lib.spev_pyin_decode(p.handle, lo.ptr(), lu.ptr(), fo.data_ptr(), len(lens), F, states.ptr(), f0.ptr(),
                                      flag.ptr(), ws.ptr(), nbytes, st))
    ws.check("viterbi workspace")
    s = states.check("states").view(torch.int32)
    assert bool(((s >= 0) & (s < 2 * p.n_bins)).all())
    f0v = f0.check("f0")
    voiced = flag.check("voiced_flag").view(torch.uint8)[:F]
    assert bool(((voiced == 1) == (s < p.n_bins)).all()) and bool((torch.isnan(f0v) == (voiced == 0)).all())
    # pooling over arbitrary phone cuts
    durs = torch.tensor([max(1, int(t) // 2) for t in fb.frames for _ in range(2)], dtype=torch.int64, device=cuda)
    po = torch.arange(0, 2 * len(lens) + 1, 2, dtype=torch.int64, device=cuda)
    pitch, rough = Guarded(2 * len(lens), cuda), Guarded(2 * len(lens), cuda)
    _lib.check(p.lib.spev_pitch_pool(p.handle, states.ptr(), fo.data_ptr(), durs.data_ptr(), po.data_ptr(), len(lens),
                                     5.2, 0.3, -2.5, 2.5, 1.5, pitch.ptr(), rough.ptr(), st))
    assert bool(torch.isfinite(pitch.check("pitch")).all()) and bool(torch.isfinite(rough.check("rough")).all())
