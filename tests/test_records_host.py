"""Host logic of the cache build (spev_real_metrics.py:328-397) against the REFERENCE'S OWN constructor, executed
by oracle/make_golden.py on tests.synth.tiny_corpus (fixture tests/golden/cache_build.npz)."""
import os

import numpy as np

from tests import synth

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "cache_build.npz"))


def test_plan_records_equals_reference_build():
    import spev_tts_b200 as sp
    corpus = synth.tiny_corpus(seed=21)
    phones, durs = synth.corpus_alignments(corpus)
    kept, k_phs, k_durs, vocab = sp.plan_records([len(it["y"]) for it in corpus], phones, durs)
    assert kept == GOLD["index"].tolist()                      # items 3 (short), 5 (no text), 7 (zero durations) dropped
    assert vocab == GOLD["vocab"].tolist()
    dropped_phones = 0
    for k, i in enumerate(kept):
        assert k_phs[k] == GOLD[f"r{k}_phs"].tolist(), i
        assert k_durs[k] == GOLD[f"r{k}_durs"].tolist(), i
        assert sum(k_durs[k]) == 1 + len(corpus[i]["y"]) // 256 == GOLD[f"r{k}_mel_shape"][0]
        dropped_phones += len(phones[i]) - len(k_phs[k])
    assert dropped_phones > 0                                  # the tail-trimming branch (:385-394) was exercised


def test_scale_durations_branches():
    import spev_tts_b200 as sp
    assert sp.scale_durations(list("abc"), [0, 0, 0], 10) is None              # :378
    assert sp.scale_durations(list("abc"), [2, 2, 2], 6) == (list("abc"), [2, 2, 2])
    assert sp.scale_durations(list("abc"), [1, 1, 1], 10) == (list("abc"), [3, 3, 4])     # shortfall -> last phone
    # max(1, .) inflates the sum: the excess is taken off the tail, phones that reach zero are dropped
    ph, du = sp.scale_durations(list("abcd"), [10, 1, 1, 1], 3)
    assert sum(du) == 3 and len(ph) == len(du) and all(d >= 1 for d in du)
    assert sp.scale_durations(list("abcdef"), [1] * 6, 2) == (list("ab"), [1, 1])
    assert sp.uniform_durations(13000, 10) == [5] * 10 and sp.uniform_durations(13000, 60) == [0] * 60
