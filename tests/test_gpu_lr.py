"""GPU parity (bit-exact): LengthRegulator, duration rule, bucketize+embedding vs golden
vectors produced by the REFERENCE's own class (/root/reference/spev_real_metrics.py:122-146,
:215) and torch.bucketize."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import librosa_restated as lr
from tests import synth

pytestmark = pytest.mark.gpu


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_cfg2_bit_exact_vs_reference_golden(cuda, golden):
    import spev_tts_b200 as sp
    x, dur, lens = synth.cfg2_batch(seed=2)
    g = golden("lr_cfg2.npz")
    out, mel_lens = sp.LengthRegulator()(torch.from_numpy(x).to(cuda), torch.from_numpy(dur).to(cuda))
    assert mel_lens.dtype == torch.int64 and mel_lens.device == out.device
    assert np.array_equal(mel_lens.cpu().numpy(), g["mel_lens"])
    o = out.cpu().numpy()
    assert tuple(o.shape) == tuple(g["shape"])
    assert sha(o) == str(g["sha256"])
    assert np.array_equal(o[3, ::7, :16], g["row3"])
    # and against the numpy restatement (which also runs on the GPU box)
    ro, rl = lr.length_regulator(x, dur)
    assert np.array_equal(o, ro) and np.array_equal(rl, g["mel_lens"])


def test_fused_variance_expand(cuda, golden):
    import spev_tts_b200 as sp
    x, dur, _ = synth.cfg2_batch(seed=2)
    feats = synth.cfg2_features(seed=2)
    g = golden("lr_cfg2_feats.npz")["feats"]                     # reference class, H=1 calls
    xe, ml, curves = sp.regulate_variances(torch.from_numpy(x).to(cuda), torch.from_numpy(dur).to(cuda),
                                           [torch.from_numpy(f).to(cuda) for f in feats], clamps=None)
    for j in range(5):
        assert curves[j].shape == (32, 1, g.shape[2])
        assert np.array_equal(curves[j][:, 0].cpu().numpy(), g[j])
    # with the reference's post-clamps (spev_real_metrics.py:239-243)
    _, _, cl = sp.regulate_variances(torch.from_numpy(x).to(cuda), torch.from_numpy(dur).to(cuda),
                                     [torch.from_numpy(f).to(cuda) for f in feats])
    for j, (lo, hi) in enumerate(sp.VARIANCE_CLAMPS):
        assert np.array_equal(cl[j][:, 0].cpu().numpy(), np.clip(g[j], np.float32(lo), np.float32(hi)))
    # H=1 through the plain class == expand_feat of the reference
    o1, _ = sp.LengthRegulator()(torch.from_numpy(feats[0]).to(cuda).unsqueeze(-1), torch.from_numpy(dur).to(cuda))
    assert np.array_equal(o1[..., 0].cpu().numpy(), g[0])
    mask = sp.mel_mask(ml, xe.shape[1])
    assert mask.shape == (32, xe.shape[1]) and bool((mask.sum(1) == xe.shape[1] - ml).all())


def test_small_and_edge_cases(cuda, golden):
    import spev_tts_b200 as sp
    g = golden("lr_small.npz")
    o, l = sp.LengthRegulator()(torch.from_numpy(g["x"]).to(cuda), torch.from_numpy(g["dur"]).to(cuda))
    assert np.array_equal(o.cpu().numpy(), g["out"]) and np.array_equal(l.cpu().numpy(), g["mel_lens"])
    e = golden("lr_edge.npz")
    for name in synth.lr_edge_cases():
        o, l = sp.LengthRegulator()(torch.from_numpy(e[name + "_x"]).to(cuda),
                                    torch.from_numpy(e[name + "_dur"]).to(cuda))
        assert np.array_equal(l.cpu().numpy(), e[name + "_lens"]), name
        assert np.array_equal(o.cpu().numpy(), e[name + "_out"]), name
    with pytest.raises(ValueError):
        sp.LengthRegulator()(torch.zeros(0, 4, 8, device=cuda), torch.zeros(0, 4, dtype=torch.int64, device=cuda))
    with pytest.raises(RuntimeError):
        sp.LengthRegulator()(torch.zeros(1, 4, 8), torch.ones(1, 4, dtype=torch.int64))   # CPU tensors: loud


@pytest.mark.parametrize("ddt", [torch.int64, torch.int32, torch.float32, torch.float64, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("xdt,H", [(torch.float32, 256), (torch.float16, 7), (torch.float64, 3), (torch.uint8, 5)])
def test_dtypes_and_row_sizes(cuda, ddt, xdt, H):
    import spev_tts_b200 as sp
    rng = np.random.default_rng(9)
    B, T = 7, 67
    d = rng.integers(-2, 12, (B, T)).astype(np.float64)
    if ddt.is_floating_point:
        d += rng.choice([0.0, 0.25, 0.5, 0.75], (B, T))
        d[0, 3] = np.nan; d[1, 4] = np.inf; d[2, 5] = 1000.5; d[3, 6] = 1000.0
    dt = torch.from_numpy(d).to(ddt)
    x = torch.from_numpy(rng.integers(0, 255, (B, T, H))).to(xdt)
    ro, rl = lr.length_regulator(x.numpy() if xdt != torch.bfloat16 else x.float().numpy(), dt.double().numpy())
    o, l = sp.LengthRegulator()(x.to(cuda), dt.to(cuda))
    assert np.array_equal(l.cpu().numpy(), rl)
    assert np.array_equal(o.cpu().numpy(), ro)


def test_long_rows_and_large_batch(cuda):
    import spev_tts_b200 as sp
    rng = np.random.default_rng(10)
    B, T, H = 512, 300, 64
    x = rng.standard_normal((B, T, H)).astype(np.float32)
    d = rng.integers(0, 9, (B, T)).astype(np.int64)
    ro, rl = lr.length_regulator(x, d)
    o, l = sp.LengthRegulator()(torch.from_numpy(x).to(cuda), torch.from_numpy(d).to(cuda))
    assert np.array_equal(l.cpu().numpy(), rl) and np.array_equal(o.cpu().numpy(), ro)


def test_duration_rule_vs_reference_formula(cuda, golden):
    import spev_tts_b200 as sp
    g = golden("duration_rule.npz")
    ld = torch.from_numpy(g["log_dur"]).to(cuda)
    for dc in (1.0, 0.5, 1.7):
        got = sp.duration_rule(ld, dc)
        assert got.dtype == torch.int64
        assert np.array_equal(got.cpu().numpy(), g[f"d_{dc}"]), dc
    # round-half-to-even and the clamp, on exact inputs: log(1+v) is not exact, so feed the
    # rule through its own inverse only for the clamp ends
    big = torch.tensor([50.0, -50.0, 0.0], device=cuda)
    assert sp.duration_rule(big).tolist() == [500, 0, 0]
    assert np.array_equal(lr.duration_rule(g["log_dur"], 1.0), g["d_1.0"])


def test_bucketize_embed_vs_torch(cuda, golden):
    import spev_tts_b200 as sp
    v, bins, table = synth.bucketize_case(seed=2)
    g = golden("bucketize.npz")
    vt, bt, tt = (torch.from_numpy(a).to(cuda) for a in (v, bins, table))
    assert np.array_equal(sp.bucketize(vt, bt).cpu().numpy(), g["idx"])
    assert np.array_equal(sp.bucketize(vt, bt, right=True).cpu().numpy(), g["idx_right"])
    emb, idx = sp.bucketize_embed(vt, bt, tt, return_index=True)
    assert np.array_equal(idx.cpu().numpy(), g["idx"])
    e = emb.cpu().numpy()
    assert e.shape == (32, 200, 256) and sha(e) == str(g["emb_sha256"])
    assert np.array_equal(e, table[g["idx"]])
    # known answers: [nan, inf, -inf, -3, 3, bins[10], 3.0001] -> [255,255,0,0,254,10,255]
    assert g["idx"][0, :7].tolist() == [255, 255, 0, 0, 254, 10, 255]
    # accumulate: x + table[idx], H not a multiple of 4 exercises the scalar path
    acc = torch.ones(32, 200, 256, device=cuda)
    sp.bucketize_embed(vt, bt, tt, accumulate_into=acc)
    assert np.array_equal(acc.cpu().numpy(), 1.0 + table[g["idx"]])
    t3 = torch.from_numpy(table[:, :7].copy()).to(cuda)
    assert np.array_equal(sp.bucketize_embed(vt, bt, t3).cpu().numpy(), table[:, :7][g["idx"]])


def test_fused_variance_adaptor_vs_reference_modules(cuda, golden):
    """expand + clamp + five Conv1d(1,256,3) embeddings + sum in one kernel == the reference model's own
    modules on the same weights (spev_real_metrics.py:226-252); float32 sums: tolerance 2e-6 abs."""
    import spev_tts_b200 as sp
    g = golden("variance_adaptor.npz")
    embs = []
    for j in range(5):
        e = torch.nn.Conv1d(1, 256, kernel_size=3, padding=1)
        with torch.no_grad():
            e.weight.copy_(torch.from_numpy(g["conv_w"][j])); e.bias.copy_(torch.from_numpy(g["conv_b"][j]))
        embs.append(e.to(cuda))
    x = torch.from_numpy(g["x"]).to(cuda)
    d = torch.from_numpy(g["dur"]).to(cuda)
    curves = [torch.from_numpy(c).to(cuda) for c in g["curves"]]
    out, mel_len, cv = sp.variance_adaptor(x, d, curves, embs, return_curves=True)
    assert np.array_equal(mel_len.cpu().numpy(), g["mel_len"])
    assert out.shape == g["dec_input"].shape
    assert np.array_equal(cv.cpu().numpy(), g["curves_expanded"])                 # expanded + clamped curves: exact
    err = np.abs(out.detach().cpu().numpy() - g["dec_input"]).max()
    assert err <= 2e-6 * max(1.0, np.abs(g["dec_input"]).max()), err
    # larger shape against torch's own conv1d on the GPU (cfg2 sizes)
    xb, db, _ = synth.cfg2_batch(seed=2)
    feats = [torch.from_numpy(f).to(cuda) * 2 for f in synth.cfg2_features(seed=2)]
    xe, ml, ce = sp.regulate_variances(torch.from_numpy(xb).to(cuda), torch.from_numpy(db).to(cuda), feats)
    ref = xe.transpose(1, 2)
    with torch.no_grad():
        for e, c in zip(embs, ce):
            ref = ref + e(c)
    ref = ref.transpose(1, 2)
    out2, ml2 = sp.variance_adaptor(torch.from_numpy(xb).to(cuda), torch.from_numpy(db).to(cuda), feats, embs)
    assert torch.equal(ml, ml2) and float((out2.detach() - ref).abs().max()) < 1e-5
