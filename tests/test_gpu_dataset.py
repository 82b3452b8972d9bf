"""GPU parity (bit-exact) of ResidentCache.collate against goldens produced by the reference's own
Dataset.__getitem__ + collate_fn (/root/reference/spev_real_metrics.py:433-462)."""
import numpy as np
import pytest
import torch

from tests import synth

pytestmark = pytest.mark.gpu


def test_collate_bit_exact_vs_reference_golden(cuda, golden, tmp_path):
    import spev_tts_b200 as sp
    recs, stats, vocab = synth.cache_records(seed=8)
    sp.write_reference_cache(str(tmp_path), recs, stats, vocab)
    cache = sp.ResidentCache.load(str(tmp_path), device=cuda)
    assert len(cache) == len(recs)
    g = golden("collate.npz")
    for name in ("a", "b", "c"):
        batch = cache.collate(g[f"{name}_idx"])
        assert set(batch) == {"ids", "lens", "durs", "mel", "log_durs", "pitch", "energy", "breath", "rough", "bright"}
        for k, v in batch.items():
            ref = g[f"{name}_{k}"]
            assert v.is_cuda and tuple(v.shape) == ref.shape and str(v.dtype).split(".")[-1] == str(ref.dtype), (name, k)
            assert np.array_equal(v.cpu().numpy(), ref), (name, k)
    assert cache.collate([]) is None


def test_collate_large_random_batches(cuda):
    import spev_tts_b200 as sp
    recs, stats, vocab = synth.cache_records(seed=21, n=300)
    cache = sp.ResidentCache(recs, vocab, stats, device=cuda)
    rng = np.random.default_rng(0)
    for _ in range(5):
        idx = rng.choice(300, 64, replace=False)
        b = cache.collate(idx)
        tmax = max(recs[i]["mel"].shape[0] for i in idx)
        assert b["mel"].shape == (64, tmax, 80)
        for j, i in enumerate(idx[:8]):
            T, P = recs[i]["mel"].shape[0], len(recs[i]["phs"])
            assert torch.equal(b["mel"][j, :T].cpu(), recs[i]["mel"]) and bool((b["mel"][j, T:] == 0).all())
            assert b["durs"][j, :P].tolist() == recs[i]["durs"] and bool((b["durs"][j, P:] == 0).all())
            assert int(b["lens"][j]) == P
            assert np.array_equal(b["energy"][j, :P].cpu().numpy(), np.asarray(recs[i]["energy"], np.float32))
