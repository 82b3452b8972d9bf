"""Whole cache records built on the GPU (spev_tts_b200.records) against the fixture produced by the REFERENCE'S
OWN RealMetricsDataset constructor (oracle/make_golden.py: its librosa calls bound to the restated oracle;
statistics, duration scaling, pooling, clipping, file naming = the reference's code, spev_real_metrics.py:300-430)."""
import os

import numpy as np
import pytest
import torch

from tests import synth

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "cache_build.npz"))
STATS = dict(zip(GOLD["stats_keys"].tolist(), GOLD["stats"].tolist()))


def test_statistics_pass(cuda):
    import spev_tts_b200 as sp
    corpus = synth.tiny_corpus(seed=21)
    stats = sp.corpus_stats([it["y"] for it in corpus], device=cuda)
    assert sorted(stats) == sorted(STATS)
    for k in ("e_mean", "e_std", "c_mean", "c_std"):
        assert abs(stats[k] - STATS[k]) <= 2e-4 * max(1.0, abs(STATS[k])), (k, stats[k], STATS[k])
    # log-f0 statistics ride on the pYIN state paths (>= 98 % frame agreement with the oracle)
    assert abs(stats["p_mean"] - STATS["p_mean"]) <= 5e-3 and abs(stats["p_std"] - STATS["p_std"]) <= 5e-3


def test_processing_pass_records(cuda, tmp_path):
    import spev_tts_b200 as sp
    corpus = synth.tiny_corpus(seed=21)
    phones, durs = synth.corpus_alignments(corpus)
    records, vocab = sp.build_records([it["y"] for it in corpus], phones, durs, STATS, device=cuda)
    assert vocab == GOLD["vocab"].tolist()
    assert [r["index"] for r in records] == GOLD["index"].tolist()
    n_ph = n_pitch_ok = n_rough_ok = n_breath_ok = 0
    for k, r in enumerate(records):
        assert r["phs"] == GOLD[f"r{k}_phs"].tolist() and r["durs"] == GOLD[f"r{k}_durs"].tolist()
        assert tuple(r["mel"].shape) == tuple(GOLD[f"r{k}_mel_shape"])
        assert np.abs(r["mel"].numpy()[::16, ::8] - GOLD[f"r{k}_mel_dec"]).max() <= 1e-4
        # energy / brightness: float32 curve -> mean -> normalise -> clip
        assert np.abs(r["energy"] - GOLD[f"r{k}_energy"]).max() <= 2e-4
        assert np.abs(r["bright"] - GOLD[f"r{k}_bright"]).max() <= 2e-4
        # pitch / roughness / breathiness ride on pYIN: equal wherever the state paths agree
        n_ph += len(r["phs"])
        n_pitch_ok += int(np.sum(np.abs(r["pitch"] - GOLD[f"r{k}_pitch"]) <= 2e-3))
        n_rough_ok += int(np.sum(np.abs(r["rough"] - GOLD[f"r{k}_rough"]) <= 2e-3))
        n_breath_ok += int(np.sum(np.abs(r["breath"] - GOLD[f"r{k}_breath"]) <= 5e-3))
        for c in ("pitch", "energy", "bright"):
            assert np.all(np.abs(r[c]) <= 2.5)
        assert np.all((r["breath"] >= 0) & (r["breath"] <= 0.8)) and np.all((r["rough"] >= 0) & (r["rough"] <= 1.5))
    assert n_pitch_ok / n_ph >= 0.97 and n_rough_ok / n_ph >= 0.97 and n_breath_ok / n_ph >= 0.97, \
        (n_pitch_ok / n_ph, n_rough_ok / n_ph, n_breath_ok / n_ph)
    # the cache written from these records uses the reference's file naming (wav index, gaps included) and reloads
    files = sp.write_reference_cache(str(tmp_path), records, STATS, vocab)
    assert [os.path.basename(f) for f in files] == [f"u_{i:05d}.pt" for i in GOLD["index"]]
    recs2, stats2, vocab2 = sp.read_reference_cache(str(tmp_path))
    assert stats2 == STATS and vocab2 == vocab and len(recs2) == len(records)
    rc = sp.ResidentCache(recs2, vocab2, stats2, device=cuda)
    batch = rc.collate([0, 3, 5])
    assert batch["mel"].shape[0] == 3 and batch["pitch"].shape == batch["ids"].shape


def test_pitch_pool_kernel_vs_numpy(cuda):
    """spev_pitch_pool against the literal numpy lines :399-414 on random state paths."""
    import spev_tts_b200 as sp
    from spev_tts_b200 import _lib
    from spev_tts_b200.batch import stream_ptr
    from spev_tts_b200.pitch import PyinContext
    p = PyinContext.get(cuda)
    rng = np.random.default_rng(5)
    U = 7
    frames = rng.integers(10, 300, U)
    fo = np.concatenate([[0], np.cumsum(frames)]).astype(np.int64)
    states = rng.integers(0, 2 * p.n_bins, fo[-1]).astype(np.int32)
    states[fo[2]: fo[3]] = p.n_bins + 5                        # an all-unvoiced utterance
    durs, po, want_p, want_r = [], [0], [], []
    freqs = p.host_tables()[1]
    pm, ps = 5.2, 0.31
    for u in range(U):
        k = int(rng.integers(1, 40))
        cuts = np.sort(rng.integers(0, frames[u] + 1, k - 1))
        d = np.diff(np.concatenate([[0], cuts, [frames[u]]]))  # zero-length phones included
        st = states[fo[u]: fo[u + 1]]
        f0 = np.where(st < p.n_bins, freqs[st % p.n_bins], np.nan)
        f0_log = np.log(np.nan_to_num(f0, nan=1e-8) + 1e-8)
        curr = 0
        for dd in d:
            seg = f0_log[curr: curr + dd]
            v = seg[seg > -5]
            want_p.append(np.clip((np.mean(v) - pm) / ps if v.size else 0, -2.5, 2.5))
            want_r.append(np.clip(np.std(v) if v.size else 0, 0.0, 1.5))
            curr += dd
        durs.append(d)
        po.append(po[-1] + k)
    d_durs = torch.from_numpy(np.concatenate(durs).astype(np.int64)).to(cuda)
    d_fo, d_po = torch.from_numpy(fo).to(cuda), torch.from_numpy(np.array(po, dtype=np.int64)).to(cuda)
    pitch = torch.empty(po[-1], dtype=torch.float32, device=cuda)
    rough = torch.empty_like(pitch)
    _lib.check(p.lib.spev_pitch_pool(p.handle, torch.from_numpy(states).to(cuda).data_ptr(), d_fo.data_ptr(), d_durs.data_ptr(),
                                     d_po.data_ptr(), U, pm, ps, -2.5, 2.5, 1.5, pitch.data_ptr(), rough.data_ptr(),
                                     stream_ptr(cuda)))
    np.testing.assert_allclose(pitch.cpu().numpy(), np.array(want_p), atol=1e-6)
    np.testing.assert_allclose(rough.cpu().numpy(), np.array(want_r), atol=1e-6)
