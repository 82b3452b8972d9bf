"""Generate ``tests/golden/*.npz``  --  TEST INFRASTRUCTURE ONLY.

Run in the build container (where ``/root/reference`` is mounted):

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden

* ``lr_*.npz``      -- produced by the REFERENCE'S OWN ``LengthRegulator``
                       (``/root/reference/spev_real_metrics.py:122-146``) and the reference's
                       duration rule (``:215``, evaluated with torch exactly as written there).
                       These pin the LengthRegulator restatement and the CUDA kernels.
* ``bucketize.npz`` -- ``torch.bucketize`` + ``F.embedding`` known answers (no reference
                       implementation exists in-tree, SURVEY a-13).
* ``logmel_*.npz``, ``gl_*.npz`` -- produced by ``oracle.librosa_restated`` (the reference's
                       librosa is not installable: parity for these is unpinned by the
                       reference; the fixtures freeze the restatement so that it cannot
                       drift silently, and ``tests/test_oracle.py`` checks it against
                       torch / torchaudio / scipy independently).

* ``collate.npz``   -- produced by the REFERENCE'S OWN ``RealMetricsDataset`` (constructor early-return
                       path ``:291-298``, ``__getitem__`` ``:433-447``) and ``collate_fn`` (``:449-462``) on a
                       cache written by ``spev_tts_b200.dataset.write_reference_cache``.

* ``variance_adaptor.npz`` -- the block ``:226-252`` executed with the reference model's own
                       ``length_regulator`` and ``*_embedding`` Conv1d modules (weights stored in the fixture).

Inputs are regenerated from seeds by ``tests/synth.py``; large outputs are stored as
SHA-256 digests plus a decimated slice so that the fixtures stay small.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import librosa_restated as lr  # noqa: E402
from oracle import reference_import  # noqa: E402
from tests import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run_reference_cache_build(ref, corpus, workdir):
    """Execute the REFERENCE'S OWN ``RealMetricsDataset.__init__`` build path (``spev_real_metrics.py:300-430``)
    on an in-memory corpus.  Its third-party calls are bound to the restated oracle (librosa.load / pyin /
    feature.*), an identity ``phonemize`` and a JSON-backed ``textgrid`` -- everything else (statistics,
    duration re-scaling, per-phone pooling, clipping, file naming, metadata) is the reference's code."""
    import contextlib
    import io
    import json
    import types
    from oracle import pyin_restated as po
    data_dir, tg_dir, cache_dir = (os.path.join(workdir, d) for d in ("data", "tg", "cache"))
    os.makedirs(data_dir), os.makedirs(tg_dir)
    for it in corpus:
        np.asarray(it["y"], dtype=np.float32).tofile(os.path.join(data_dir, it["name"] + ".wav"))
        if it["text"] is not None:
            with open(os.path.join(data_dir, it["name"] + ".txt"), "w") as f:
                f.write(it["text"])
        if it["intervals"] is not None:
            with open(os.path.join(tg_dir, it["name"] + ".TextGrid"), "w") as f:
                json.dump(it["intervals"], f)

    class _Tier(list):
        name = "phones"

    class _TextGrid:
        @staticmethod
        def fromFile(path):
            with open(path) as f:
                return [_Tier(types.SimpleNamespace(minTime=a, maxTime=b, mark=m) for a, b, m in json.load(f))]

    L = ref.librosa
    saved = {k: getattr(L, k, None) for k in ("load", "pyin", "feature")}
    saved_mod = {k: getattr(ref, k, None) for k in ("phonemize", "textgrid", "TEXTGRID_AVAILABLE", "tqdm")}
    L.load = lambda path, sr=None: (np.fromfile(path, dtype=np.float32), sr)
    L.pyin = lambda y, fmin, fmax, sr=22050, hop_length=None, **k: po.pyin(
        y, fmin=fmin, fmax=fmax, sr=sr, hop_length=hop_length or 512)
    L.feature = types.SimpleNamespace(rms=lr.rms, spectral_centroid=lr.spectral_centroid,
                                      melspectrogram=lr.melspectrogram)
    ref.phonemize = lambda text, **k: text
    ref.textgrid = types.SimpleNamespace(TextGrid=_TextGrid)
    ref.TEXTGRID_AVAILABLE = True
    ref.tqdm = lambda it, **k: it
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            ds = ref.RealMetricsDataset(data_dir, textgrid_dir=tg_dir, cache_dir=cache_dir, force_rebuild=True)
    finally:
        for k, v in saved.items():
            setattr(L, k, v) if v is not None else (hasattr(L, k) and delattr(L, k))
        for k, v in saved_mod.items():
            setattr(ref, k, v)
    return ds


def cache_build_golden(ref) -> None:
    import tempfile
    corpus = synth.tiny_corpus(seed=21)
    with tempfile.TemporaryDirectory() as tmp:
        ds = run_reference_cache_build(ref, corpus, tmp)
        assert len(ds) >= 8, len(ds)
        gold = {"stats_keys": np.array(sorted(ds.stats)), "stats": np.array([ds.stats[k] for k in sorted(ds.stats)]),
                "vocab": np.array(ds.vocab), "n_records": np.array(len(ds))}
        idx = []
        for k, path in enumerate(ds.metadata):
            u = torch.load(path, weights_only=False)
            i = int(os.path.basename(path)[2:7])
            idx.append(i)
            gold[f"r{k}_phs"] = np.array(u["phs"])
            gold[f"r{k}_durs"] = np.array(u["durs"], dtype=np.int64)
            gold[f"r{k}_mel_shape"] = np.array(u["mel"].shape)
            gold[f"r{k}_mel_dec"] = u["mel"].numpy()[::16, ::8].copy()
            for c in ("pitch", "energy", "breath", "rough", "bright"):
                gold[f"r{k}_{c}"] = np.asarray(u[c])
        gold["index"] = np.array(idx)
    np.savez_compressed(os.path.join(OUT, "cache_build.npz"), **gold)


def main() -> None:
    os.makedirs(OUT, exist_ok=True)
    ref = reference_import.load()
    LR = ref.LengthRegulator()

    # ---- cfg2: B=32, T<=200, H=256, int64 durations (reference class, CPU) -------------
    x, dur, lens = synth.cfg2_batch(seed=2)
    out, mel_lens = LR(torch.from_numpy(x), torch.from_numpy(dur))
    out = out.numpy()
    np.savez_compressed(os.path.join(OUT, "lr_cfg2.npz"), mel_lens=mel_lens.numpy(),
                        shape=np.array(out.shape), sha256=np.array(sha(out)),
                        row3=out[3, ::7, :16].copy(), lens=lens)
    # the five scalar curves go through the same class with H=1 (``:228-236``)
    feats = synth.cfg2_features(seed=2)
    fo = [LR(torch.from_numpy(f).unsqueeze(-1), torch.from_numpy(dur))[0].numpy()[..., 0]
          for f in feats]
    np.savez_compressed(os.path.join(OUT, "lr_cfg2_feats.npz"), feats=np.stack(fo))

    # ---- small dense case, stored in full ---------------------------------------------
    rng = np.random.default_rng(22)
    xs = rng.standard_normal((5, 17, 12)).astype(np.float32)
    ds = rng.integers(0, 6, (5, 17)).astype(np.int64)
    ds[2] = 0
    o, l = LR(torch.from_numpy(xs), torch.from_numpy(ds))
    np.savez_compressed(os.path.join(OUT, "lr_small.npz"), x=xs, dur=ds, out=o.numpy(),
                        mel_lens=l.numpy())

    # ---- edge cases (SURVEY App. B last row) -------------------------------------------
    cases = {}
    for name, (xe, de) in synth.lr_edge_cases().items():
        o, l = LR(torch.from_numpy(xe), torch.from_numpy(de))
        cases[name + "_x"] = xe
        cases[name + "_dur"] = de
        cases[name + "_out"] = o.numpy()
        cases[name + "_lens"] = l.numpy()
    np.savez_compressed(os.path.join(OUT, "lr_edge.npz"), **cases)

    # ---- inference duration rule, exactly ``spev_real_metrics.py:215`` -----------------
    ld = synth.log_durations(seed=7)
    rule = {}
    for dc in (1.0, 0.5, 1.7):
        t = torch.clamp((torch.exp(torch.from_numpy(ld)) - 1) * dc, min=0, max=500).round().long()
        rule[f"d_{dc}"] = t.numpy()
    half = np.array([0.5, 1.5, 2.5, 3.5, 499.5, 500.5, 1e9, -3.0], dtype=np.float32)
    rule["half_in"] = half
    rule["half_out"] = torch.clamp(torch.from_numpy(half), min=0, max=500).round().long().numpy()
    np.savez_compressed(os.path.join(OUT, "duration_rule.npz"), log_dur=ld, **rule)

    # ---- bucketize + embedding (torch CPU ops are the oracle) --------------------------
    v, bins, table = synth.bucketize_case(seed=2)
    idx = torch.bucketize(torch.from_numpy(v), torch.from_numpy(bins))
    idx_r = torch.bucketize(torch.from_numpy(v), torch.from_numpy(bins), right=True)
    emb = torch.nn.functional.embedding(idx, torch.from_numpy(table)).numpy()
    np.savez_compressed(os.path.join(OUT, "bucketize.npz"), idx=idx.numpy(),
                        idx_right=idx_r.numpy(), emb_sha256=np.array(sha(emb)),
                        emb_head=emb[:4, :8].copy())

    # ---- restated librosa path (unpinned by the reference; frozen here) ----------------
    for name, y in (("white", synth.white(seed=0)), ("speechy", synth.speechy(seed=1))):
        lm = lr.reference_logmel(y)
        np.savez_compressed(os.path.join(OUT, f"logmel_{name}.npz"), logmel=lm)
    lm = lr.reference_logmel(synth.speechy(seed=3, n=256 * 63))
    S = lr.mel_to_stft(np.exp(lm.T), sr=22050, n_fft=1024, fmin=0, fmax=8000)
    ph = synth.init_phase(S.shape, seed=3)
    y8, ang8, tprev8 = lr.griffinlim(S, n_iter=8, hop_length=256, n_fft=1024, init_phase=ph,
                                     return_state=True)
    np.savez_compressed(os.path.join(OUT, "gl_small.npz"), logmel=lm, S=S, y8=y8,
                        sc8=np.array(lr.spectral_convergence(y8, S)))
    # ---- cache consumer: the reference's own Dataset.__getitem__ + collate_fn (:433-462) ----------
    import tempfile
    from spev_tts_b200.dataset import write_reference_cache
    recs, stats, vocab = synth.cache_records(seed=8)
    with tempfile.TemporaryDirectory() as tmp:
        write_reference_cache(tmp, recs, stats, vocab)
        # the reference's own constructor takes its early-return path (:291-298: > 10 records and a
        # metadata.json) -- i.e. it accepts the cache we wrote
        ds = ref.RealMetricsDataset("unused_data_dir", cache_dir=tmp, force_rebuild=False)
        assert len(ds) == len(recs) and ds.vocab == vocab and ds.stats == stats
        gold = {}
        for name, idx in (("a", [0, 5, 2, 7]), ("b", list(range(len(recs)))), ("c", [3])):
            batch = ref.collate_fn([ds[i] for i in idx])
            gold[name + "_idx"] = np.array(idx)
            for k, v in batch.items():
                gold[f"{name}_{k}"] = v.numpy()
    np.savez_compressed(os.path.join(OUT, "collate.npz"), **gold)

    # ---- variance adaptor block, run through the reference model's OWN modules (:226-252) -------------
    torch.manual_seed(11)
    model = ref.RealMetricsFastSpeech2(vocab_size=40).eval()
    rng = np.random.default_rng(11)
    B, T, H = 3, 15, 256
    xv = torch.from_numpy(rng.standard_normal((B, T, H)).astype(np.float32))
    dv = torch.from_numpy(rng.integers(0, 7, (B, T)).astype(np.int64))
    dv[1, 9:] = 0                                                   # a shorter row -> padded frames
    curves = [torch.from_numpy((3 * rng.standard_normal((B, T))).astype(np.float32)) for _ in range(5)]
    with torch.no_grad():                                           # the statements of :226-252, verbatim
        x_expanded, mel_len = model.length_regulator(xv, dv)

        def expand_feat(f, d):
            expanded, _ = model.length_regulator(f.unsqueeze(-1), d)
            return expanded.transpose(1, 2)
        pitch, energy, breath, rough, bright = [expand_feat(c, dv) for c in curves]
        pitch = torch.clamp(pitch, -3.0, 3.0)
        energy = torch.clamp(energy, -3.0, 3.0)
        breath = torch.clamp(breath, 0.0, 1.0)
        rough = torch.clamp(rough, 0.0, 2.0)
        bright = torch.clamp(bright, -3.0, 3.0)
        dec_input = x_expanded.transpose(1, 2)
        dec_input = dec_input + model.pitch_embedding(pitch) + model.energy_embedding(energy) + \
            model.breath_embedding(breath) + model.rough_embedding(rough) + model.bright_embedding(bright)
        dec_input = dec_input.transpose(1, 2)
    embs = [model.pitch_embedding, model.energy_embedding, model.breath_embedding, model.rough_embedding,
            model.bright_embedding]
    np.savez_compressed(os.path.join(OUT, "variance_adaptor.npz"), x=xv.numpy(), dur=dv.numpy(),
                        curves=np.stack([c.numpy() for c in curves]),
                        conv_w=np.stack([e.weight.detach().numpy() for e in embs]),      # [5, 256, 1, 3]
                        conv_b=np.stack([e.bias.detach().numpy() for e in embs]),
                        dec_input=dec_input.numpy(), mel_len=mel_len.numpy(),
                        curves_expanded=np.stack([t[:, 0].numpy() for t in (pitch, energy, breath, rough, bright)]))
    # ---- backward of the variance adaptor block: the reference model's own modules under autograd -------
    xg = xv.clone().requires_grad_(True)
    cg = [c.clone().requires_grad_(True) for c in curves]
    for e in embs:
        e.zero_grad()
    x_expanded, _ = model.length_regulator(xg, dv)
    ex = [model.length_regulator(c.unsqueeze(-1), dv)[0].transpose(1, 2) for c in cg]
    ex = [torch.clamp(e, lo, hi) for e, (lo, hi) in zip(ex, ((-3.0, 3.0), (-3.0, 3.0), (0.0, 1.0), (0.0, 2.0), (-3.0, 3.0)))]
    di = x_expanded.transpose(1, 2)
    di = di + model.pitch_embedding(ex[0]) + model.energy_embedding(ex[1]) + model.breath_embedding(ex[2]) + \
        model.rough_embedding(ex[3]) + model.bright_embedding(ex[4])
    di = di.transpose(1, 2)
    gup = torch.from_numpy(synth.upstream_grad(di.shape, seed=12))
    (di * gup).sum().backward()
    np.savez_compressed(os.path.join(OUT, "variance_adaptor_bwd.npz"),
                        grad_x=xg.grad.numpy(), grad_curves=np.stack([c.grad.numpy() for c in cg]),
                        grad_w=np.stack([e.weight.grad.numpy() for e in embs]),
                        grad_b=np.stack([e.bias.grad.numpy() for e in embs]))

    # ---- LengthRegulator backward: the REFERENCE'S OWN class under torch autograd (:122-146, used under
    #      loss.backward() at :574) -------------------------------------------------------------------------
    bw = {}
    x, dur, _ = synth.cfg2_batch(seed=2)
    xt = torch.from_numpy(x).requires_grad_(True)
    o, _ = LR(xt, torch.from_numpy(dur))
    (o * torch.from_numpy(synth.upstream_grad(o.shape, seed=13))).sum().backward()
    bw["cfg2_grad_x_dec"] = xt.grad.numpy()[::3, ::7, ::5].copy()
    bw["cfg2_grad_x_sha256_f32"] = np.array(sha(xt.grad.numpy()))
    ft = torch.from_numpy(synth.cfg2_features(seed=2)[0]).requires_grad_(True)          # expand_feat, H = 1 (:228-230)
    o1, _ = LR(ft.unsqueeze(-1), torch.from_numpy(dur))
    (o1 * torch.from_numpy(synth.upstream_grad(o1.shape, seed=14))).sum().backward()
    bw["cfg2_grad_feat0"] = ft.grad.numpy()
    small = np.load(os.path.join(OUT, "lr_small.npz"))
    for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        xs_t = torch.from_numpy(small["x"]).to(dt).requires_grad_(True)
        o, _ = LR(xs_t, torch.from_numpy(small["dur"]))
        (o * torch.from_numpy(synth.upstream_grad(o.shape, seed=15)).to(dt)).sum().backward()
        bw[f"small_grad_x_{tag}"] = xs_t.grad.numpy()
    for name, (xe, de) in synth.lr_edge_cases().items():
        xe_t = torch.from_numpy(xe).requires_grad_(True)
        o, _ = LR(xe_t, torch.from_numpy(de))
        if o.requires_grad:      # an all-empty batch yields constant zeros in the reference (no grad_fn)
            (o * torch.from_numpy(synth.upstream_grad(o.shape, seed=16))).sum().backward()
        bw[f"edge_{name}_grad_x"] = xe_t.grad.numpy() if xe_t.grad is not None else np.zeros_like(xe)
    np.savez_compressed(os.path.join(OUT, "lr_backward.npz"), **bw)

    cache_build_golden(ref)
    print("golden fixtures written to", OUT)
    for f in sorted(os.listdir(OUT)):
        print(f"  {f:28s} {os.path.getsize(os.path.join(OUT, f)):>9d} B")


if __name__ == "__main__":
    main()
