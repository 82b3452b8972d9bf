"""CPU tests of the cache record format: the files we write are the reference's
format (spev_real_metrics.py:419-430) and, where /root/reference is mounted, its own Dataset loads them."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import reference_import
from tests import synth


def test_write_read_roundtrip(tmp_path):
    from spev_tts_b200.dataset import read_reference_cache, write_reference_cache
    recs, stats, vocab = synth.cache_records(seed=8)
    files = write_reference_cache(str(tmp_path), recs, stats, vocab)
    assert [os.path.basename(f) for f in files[:2]] == ["u_00000.pt", "u_00001.pt"]
    meta = json.load(open(tmp_path / "metadata.json"))
    assert set(meta) == {"files", "stats", "vocab"} and meta["files"] == files and meta["vocab"] == vocab
    back, st, vc = read_reference_cache(str(tmp_path))
    assert st == stats and vc == vocab and len(back) == len(recs)
    for a, b in zip(recs, back):
        assert set(b) == {"phs", "durs", "mel", "pitch", "energy", "breath", "rough", "bright"}
        assert b["phs"] == a["phs"] and b["durs"] == a["durs"] and torch.equal(b["mel"], a["mel"])
        assert b["mel"].shape[1] == 80 and b["mel"].dtype == torch.float32            # [T, 80] == mel.T (:421)
        assert np.array_equal(b["pitch"], a["pitch"])


@pytest.mark.skipif(not reference_import.available(), reason="/root/reference not mounted")
def test_reference_dataset_accepts_our_cache_and_matches_golden(tmp_path, golden):
    from spev_tts_b200.dataset import write_reference_cache
    ref = reference_import.load()
    recs, stats, vocab = synth.cache_records(seed=8)
    write_reference_cache(str(tmp_path), recs, stats, vocab)
    ds = ref.RealMetricsDataset("unused", cache_dir=str(tmp_path), force_rebuild=False)   # early-return path :291-298
    assert len(ds) == len(recs)
    g = golden("collate.npz")
    batch = ref.collate_fn([ds[i] for i in g["a_idx"]])
    for k, v in batch.items():
        assert np.array_equal(v.numpy(), g[f"a_{k}"]), k
