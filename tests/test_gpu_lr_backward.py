"""GPU: the LengthRegulator / variance-adaptor drop-ins are differentiable like the reference
(/root/reference/spev_real_metrics.py:122-146 used under loss.backward(), :544-574).

Gradients are compared with (i) goldens produced by the REFERENCE'S OWN class / model modules under torch autograd
(oracle/make_golden.py), (ii) the statement-by-statement restatement oracle/torch_reference.py running on the same
device, (iii) torch.autograd.gradcheck in float64, and -- where /root/reference is mounted -- (iv) the reference's
whole RealMetricsFastSpeech2 with and without patch_model().
Tolerances: float64 exact to 1e-12; float32 segment sums <= 1e-6 relative to the gradient scale (the kernel adds
a segment's frames in ascending order; torch's repeat-backward may associate differently)."""
import numpy as np
import pytest
import torch

from oracle import reference_import
from oracle import torch_reference as tr
from tests import synth

pytestmark = pytest.mark.gpu


def _ours_grad(x, d, seed, dtype=torch.float32, cuda=None):
    import spev_tts_b200 as sp
    xt = torch.from_numpy(x).to(cuda, dtype).requires_grad_(True)
    o, _ = sp.LengthRegulator()(xt, torch.from_numpy(d).to(cuda))
    assert o.requires_grad and o.grad_fn is not None
    (o * torch.from_numpy(synth.upstream_grad(o.shape, seed)).to(cuda, dtype)).sum().backward()
    return xt.grad.cpu().numpy()


def test_backward_vs_reference_class_goldens(cuda, golden):
    g = golden("lr_backward.npz")
    x, dur, _ = synth.cfg2_batch(seed=2)
    gx = _ours_grad(x, dur, 13, cuda=cuda)
    assert gx.shape == x.shape
    ref = g["cfg2_grad_x_dec"]
    assert np.abs(gx[::3, ::7, ::5] - ref).max() <= 1e-6 * max(1.0, np.abs(ref).max())
    # expand_feat (:228-230): H = 1 through the plain class
    import spev_tts_b200 as sp
    ft = torch.from_numpy(synth.cfg2_features(seed=2)[0]).to(cuda).requires_grad_(True)
    o1, _ = sp.LengthRegulator()(ft.unsqueeze(-1), torch.from_numpy(dur).to(cuda))
    (o1 * torch.from_numpy(synth.upstream_grad(o1.shape, seed=14)).to(cuda)).sum().backward()
    assert np.abs(ft.grad.cpu().numpy() - g["cfg2_grad_feat0"]).max() <= 1e-6 * np.abs(g["cfg2_grad_feat0"]).max()
    s = golden("lr_small.npz")
    assert np.abs(_ours_grad(s["x"], s["dur"], 15, cuda=cuda) - g["small_grad_x_f32"]).max() <= 2e-6
    assert np.abs(_ours_grad(s["x"], s["dur"], 15, torch.float64, cuda) - g["small_grad_x_f64"]).max() <= 1e-12
    for name, (xe, de) in synth.lr_edge_cases().items():       # zero rows, >1000, negatives, NaN/inf durations
        got = _ours_grad(xe, de, 16, cuda=cuda)
        ref = g[f"edge_{name}_grad_x"]              # "gt1000" holds a 1000-frame segment: float32 sum order matters
        assert np.abs(got - ref).max() <= 1e-6 * max(1.0, np.abs(ref).max()), name


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_backward_half_precisions(cuda, dtype):
    import spev_tts_b200 as sp
    x, d, _ = synth.cfg2_batch(seed=4, B=3, T=60, H=24)
    d = np.minimum(d, 4)                      # segment sums of <= 4 small integers: exact in half precision
    xt = torch.from_numpy(np.round(x * 4)).to(cuda, dtype).requires_grad_(True)
    o, _ = sp.LengthRegulator()(xt, torch.from_numpy(d).to(cuda))
    gup = torch.from_numpy(np.round(synth.upstream_grad(o.shape, 3) * 2)).to(cuda, dtype)
    (o * gup).sum().backward()
    xr = xt.detach().double().cpu().requires_grad_(True)
    orf, _ = tr.LengthRegulator()(xr, torch.from_numpy(d))
    (orf * gup.double().cpu()).sum().backward()
    assert xt.grad.dtype == dtype and torch.equal(xt.grad.double().cpu(), xr.grad)


def test_gradcheck_float64(cuda):
    import spev_tts_b200 as sp
    rng = np.random.default_rng(31)
    B, T, H = 3, 9, 5
    d = torch.from_numpy(rng.integers(0, 4, (B, T))).to(cuda)
    d[1] = 0                                                     # an empty row (one zero frame)
    x = torch.from_numpy(rng.standard_normal((B, T, H))).to(cuda).requires_grad_(True)
    lr = sp.LengthRegulator()
    assert torch.autograd.gradcheck(lambda t: lr(t, d)[0], (x,), eps=1e-6, atol=1e-8)
    # and equal to the autograd of the restated reference class, element for element
    xo = x.detach().clone().requires_grad_(True)
    g = torch.from_numpy(rng.standard_normal(tuple(lr(x, d)[0].shape))).to(cuda)
    (lr(x, d)[0] * g).sum().backward()
    (tr.LengthRegulator()(xo, d)[0] * g).sum().backward()
    assert torch.allclose(x.grad, xo.grad, rtol=0, atol=1e-12)


def test_regulate_variances_backward_with_clamp_mask(cuda):
    """x and the five curves through the fused expand: gradients == autograd of the restated statements
    :226-243 (curve values outside the clamp ranges receive no gradient, like torch.clamp)."""
    import spev_tts_b200 as sp
    x, dur, _ = synth.cfg2_batch(seed=2)
    feats = [2.5 * f for f in synth.cfg2_features(seed=2)]       # scale: a good share outside [-3,3] / [0,1] / [0,2]
    dd = torch.from_numpy(dur).to(cuda)
    xa = torch.from_numpy(x).to(cuda).requires_grad_(True)
    fa = [torch.from_numpy(f).to(cuda).requires_grad_(True) for f in feats]
    xe, ml, ce = sp.regulate_variances(xa, dd, fa)
    gx = torch.from_numpy(synth.upstream_grad(xe.shape, 41)).to(cuda)
    gc = [torch.from_numpy(synth.upstream_grad(c.shape, 42 + j)).to(cuda) for j, c in enumerate(ce)]
    ((xe * gx).sum() + sum((c * g).sum() for c, g in zip(ce, gc))).backward()
    # restated reference statements on the same device
    xb = torch.from_numpy(x).to(cuda).requires_grad_(True)
    fb = [torch.from_numpy(f).to(cuda).requires_grad_(True) for f in feats]
    LR = tr.LengthRegulator()
    xr, mlr = LR(xb, dd)
    cr = [torch.clamp(LR(f.unsqueeze(-1), dd)[0].transpose(1, 2), lo, hi) for f, (lo, hi) in zip(fb, tr.VARIANCE_CLAMPS)]
    ((xr * gx).sum() + sum((c * g).sum() for c, g in zip(cr, gc))).backward()
    assert torch.equal(ml, mlr) and torch.equal(xe, xr)
    assert float((xa.grad - xb.grad).abs().max()) <= 1e-6 * float(xb.grad.abs().max())
    for j in range(5):
        assert float((fa[j].grad - fb[j].grad).abs().max()) <= 1e-6 * float(fb[j].grad.abs().max()), j
        blocked = (fb[j].detach() < tr.VARIANCE_CLAMPS[j][0]) | (fb[j].detach() > tr.VARIANCE_CLAMPS[j][1])
        assert bool(blocked.any()) and bool((fa[j].grad[blocked] == 0).all())


def _embeddings(g, cuda):
    embs = []
    for j in range(5):
        e = torch.nn.Conv1d(1, 256, kernel_size=3, padding=1)
        with torch.no_grad():
            e.weight.copy_(torch.from_numpy(g["conv_w"][j])); e.bias.copy_(torch.from_numpy(g["conv_b"][j]))
        embs.append(e.to(cuda))
    return embs


def test_fused_variance_adaptor_backward_vs_reference_model_goldens(cuda, golden):
    """variance_adaptor (one fused forward kernel, one fused backward kernel) against the autograd of the
    reference model's own modules on the same weights (:226-252)."""
    import spev_tts_b200 as sp
    g, gb = golden("variance_adaptor.npz"), golden("variance_adaptor_bwd.npz")
    embs = _embeddings(g, cuda)
    x = torch.from_numpy(g["x"]).to(cuda).requires_grad_(True)
    curves = [torch.from_numpy(c).to(cuda).requires_grad_(True) for c in g["curves"]]
    out, ml = sp.variance_adaptor(x, torch.from_numpy(g["dur"]).to(cuda), curves, embs)
    (out * torch.from_numpy(synth.upstream_grad(out.shape, seed=12)).to(cuda)).sum().backward()

    def close(a, ref, what):
        err = np.abs(a.cpu().numpy() - ref).max()
        assert err <= 2e-6 * max(1.0, np.abs(ref).max()), (what, err)
    close(x.grad, gb["grad_x"], "grad_x")
    close(torch.stack([c.grad for c in curves]), gb["grad_curves"], "grad_curves")
    close(torch.stack([e.weight.grad for e in embs]), gb["grad_w"], "grad_w")
    close(torch.stack([e.bias.grad for e in embs]), gb["grad_b"], "grad_b")


def test_fused_variance_adaptor_backward_cfg2(cuda, golden):
    """cfg2 sizes, against the unfused differentiable path (regulate_variances + torch conv1d) on the GPU; also run
    twice: the backward is bit-reproducible (no atomics)."""
    import spev_tts_b200 as sp
    embs = _embeddings(golden("variance_adaptor.npz"), cuda)
    xb, db, _ = synth.cfg2_batch(seed=2)
    feats = [2.0 * f for f in synth.cfg2_features(seed=2)]
    dd = torch.from_numpy(db).to(cuda)

    def run(fused):
        for e in embs:
            e.zero_grad()
        x = torch.from_numpy(xb).to(cuda).requires_grad_(True)
        cv = [torch.from_numpy(f).to(cuda).requires_grad_(True) for f in feats]
        if fused:
            out, _ = sp.variance_adaptor(x, dd, cv, embs)
        else:
            out, _ = tr.variance_adaptor(x, dd, cv, embs, length_regulator=sp.LengthRegulator())
        (out * torch.from_numpy(synth.upstream_grad(out.shape, 77)).to(cuda)).sum().backward()
        return [x.grad, torch.stack([c.grad for c in cv]), torch.stack([e.weight.grad.clone() for e in embs]),
                torch.stack([e.bias.grad.clone() for e in embs])]
    a, b, a2 = run(True), run(False), run(True)
    for u, v, w, name in zip(a, b, a2, ("x", "curves", "w", "b")):
        scale = float(v.abs().max())
        assert float((u.reshape(v.shape) - v).abs().max()) <= 2e-5 * scale, name   # 64 k-term float32 sums for w / b
        assert torch.equal(u, w), name


def test_training_step_through_a_surrogate_model(cuda):
    """A small encoder -> LengthRegulator -> decoder network trained for one step with the restated reference
    class and with the drop-in: every parameter receives a gradient and the gradients agree (<= 1e-6 relative)."""
    import spev_tts_b200 as sp

    class Net(torch.nn.Module):
        def __init__(self, lr):
            super().__init__()
            self.emb = torch.nn.Embedding(30, 64, padding_idx=0)
            self.enc = torch.nn.Linear(64, 64)
            self.length_regulator = lr
            self.dec = torch.nn.Linear(64, 80)

        def forward(self, ids, durs):
            h = torch.tanh(self.enc(self.emb(ids)))
            e, ml = self.length_regulator(h, durs)
            return self.dec(e), ml
    rng = np.random.default_rng(51)
    ids = torch.from_numpy(rng.integers(1, 30, (8, 40))).to(cuda)
    durs = torch.from_numpy(rng.integers(0, 9, (8, 40))).to(cuda)
    grads = []
    for lr in (tr.LengthRegulator(), sp.LengthRegulator()):
        torch.manual_seed(5)
        net = Net(lr).to(cuda)
        out, ml = net(ids, durs)
        tgt = torch.from_numpy(synth.upstream_grad(out.shape, 52)).to(cuda)
        torch.nn.functional.l1_loss(out, tgt).backward()
        assert all(p.grad is not None for p in net.parameters())
        grads.append({k: p.grad.clone() for k, p in net.named_parameters()})
    for k in grads[0]:
        scale = max(float(grads[0][k].abs().max()), 1e-12)
        assert float((grads[0][k] - grads[1][k]).abs().max()) <= 1e-6 * max(scale, 1.0), k


def test_known_max_len_needs_no_sync_and_is_graph_capturable(cuda):
    import spev_tts_b200 as sp
    x, dur, _ = synth.cfg2_batch(seed=2)
    xd, dd = torch.from_numpy(x).to(cuda), torch.from_numpy(dur).to(cuda)
    ref, ml = sp.LengthRegulator()(xd, dd)
    maxF = ref.shape[1]
    out, ml2 = sp.LengthRegulator()(xd, dd, max_len=maxF)
    assert torch.equal(out, ref) and torch.equal(ml, ml2)
    longer, _ = sp.LengthRegulator()(xd, dd, max_len=maxF + 5)       # extra zero frames
    assert torch.equal(longer[:, :maxF], ref) and float(longer[:, maxF:].abs().max()) == 0.0
    # the whole forward inside a CUDA graph (impossible with the host read of max_len)
    s = torch.cuda.Stream(cuda)
    with torch.cuda.stream(s):
        sp.regulate_variances(xd, dd, [xd[..., 0]] * 5, max_len=maxF)    # warm-up allocations
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            o, m, cv = sp.regulate_variances(xd, dd, [xd[..., 0]] * 5, max_len=maxF)
        o.zero_()
        g.replay()
    torch.cuda.synchronize(cuda)
    assert torch.equal(o, ref) and torch.equal(m, ml)


def test_clamp_propagates_nan_like_torch(cuda):
    import spev_tts_b200 as sp
    x, dur, _ = synth.cfg2_batch(seed=2, B=2, T=60, H=8)
    f = synth.cfg2_features(seed=2, B=2, T=60)[0]
    f[0, 3] = np.nan
    dur[0, 3] = 2
    _, _, cv = sp.regulate_variances(torch.from_numpy(x).to(cuda), torch.from_numpy(dur).to(cuda),
                                     [torch.from_numpy(f).to(cuda)] * 5)
    ref = torch.clamp(tr.LengthRegulator()(torch.from_numpy(f).unsqueeze(-1), torch.from_numpy(dur))[0].transpose(1, 2), -3.0, 3.0)
    assert bool(torch.isnan(cv[0]).any()) and torch.equal(torch.isnan(cv[0].cpu()), torch.isnan(ref))
    assert np.array_equal(cv[0].cpu().numpy(), ref.numpy(), equal_nan=True)


@pytest.mark.skipif(not reference_import.available(), reason="/root/reference is not mounted on this box")
def test_reference_model_grads_with_and_without_patch_model(cuda):
    """The reference's RealMetricsFastSpeech2 forward + loss.backward() (:544-574) with its own LengthRegulator and
    after spev_tts_b200.patch_model(): no parameter loses its gradient and all gradients agree."""
    import spev_tts_b200 as sp
    ref = reference_import.load()
    rng = np.random.default_rng(61)
    B, T = 4, 24
    ids = torch.from_numpy(rng.integers(1, 40, (B, T))).to(cuda)
    lens = torch.full((B,), T, device=cuda)
    durs = torch.from_numpy(rng.integers(0, 6, (B, T))).to(cuda)
    curves = [torch.from_numpy(rng.standard_normal((B, T)).astype(np.float32)).to(cuda) for _ in range(5)]
    grads = []
    for patched in (False, True):
        torch.manual_seed(7)
        model = ref.RealMetricsFastSpeech2(vocab_size=40).to(cuda).eval()    # eval: dropout off, same seed
        if patched:
            assert sp.patch_model(model) == 1
        out = model(ids, lens, target_durations=durs, target_pitch=curves[0], target_energy=curves[1],
                    target_breath=curves[2].abs(), target_rough=curves[3].abs(), target_bright=curves[4])
        tgt = torch.from_numpy(synth.upstream_grad(out["mel_pred"].shape, 62)).to(cuda)
        loss = torch.nn.functional.l1_loss(out["mel_pred"], tgt) + out["log_duration_pred"].pow(2).mean()
        loss.backward()
        grads.append({k: (None if p.grad is None else p.grad.clone()) for k, p in model.named_parameters()})
    for k, g0 in grads[0].items():
        g1 = grads[1][k]
        assert (g0 is None) == (g1 is None), k
        if g0 is not None:
            assert float((g0 - g1).abs().max()) <= 1e-6 * max(1.0, float(g0.abs().max())), k
