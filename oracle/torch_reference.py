"""Autograd-capable torch restatement of the reference's LengthRegulator and of the variance-adaptor block
--  TEST INFRASTRUCTURE ONLY (the checker for the CUDA backward kernels, never the product).

``/root/reference`` is not present on the GPU box, so the ``-m gpu`` tests cannot instantiate the reference's
classes there.  This module restates, statement by statement, the two pieces of
``/root/reference/spev_real_metrics.py`` that the product's backward must reproduce:

* ``LengthRegulator.forward``  (``:122-146``): per-(b,t) ``.item()``, validation, ``repeat`` / ``cat`` /
  ``F.pad`` / ``stack`` -- all differentiable torch ops, which is what the reference's Trainer
  back-propagates through (``:544-546``, ``loss.backward()`` ``:574``);
* the variance-adaptor statements of ``RealMetricsFastSpeech2.forward`` (``:226-252``).

PIN STATUS: pinned by the reference.  ``tests/test_oracle_torch_reference.py`` (CPU, runs where
``/root/reference`` is mounted) checks outputs AND gradients of both restatements against the reference's own
class / model modules, and ``oracle/make_golden.py`` stores gradients produced by the reference's own class under
``tests/golden/lr_backward.npz`` / ``variance_adaptor.npz`` for the GPU tests.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


class LengthRegulator(nn.Module):
    """``spev_real_metrics.py:122-146``."""

    def forward(self, x, durations):
        rows, mel_lens = [], []
        for b in range(x.size(0)):                                   # :126
            pieces = []
            for t in range(x.size(1)):                               # :128
                d = durations[b, t].item()                           # :129
                if not np.isfinite(d) or d < 0 or d > 1000:          # :131
                    d = 0
                n = int(d)                                           # :133
                if n > 0:
                    pieces.append(x[b, t:t + 1].repeat(n, 1))        # :135
            if not pieces:                                           # :137-139: empty row -> one zero frame
                rows.append(torch.zeros(1, x.size(2), device=x.device))
                mel_lens.append(1)
            else:                                                    # :140-142
                rows.append(torch.cat(pieces, dim=0))
                mel_lens.append(rows[-1].size(0))
        max_len = max(mel_lens)                                      # :144
        stacked = torch.stack([F.pad(o, (0, 0, 0, max_len - o.size(0))) for o in rows])   # :145
        return stacked, torch.LongTensor(mel_lens).to(x.device)      # :146


VARIANCE_CLAMPS = ((-3.0, 3.0), (-3.0, 3.0), (0.0, 1.0), (0.0, 2.0), (-3.0, 3.0))   # :239-243


def variance_adaptor(x, durations, curves, embeddings, length_regulator=None, clamps=VARIANCE_CLAMPS):
    """``spev_real_metrics.py:226-252``: six LengthRegulator calls, five clamps, five Conv1d(1,H,3,padding=1)
    embeddings added to the expanded hidden states.  ``curves``: (pitch, energy, breath, rough, bright) ``[B,T]``;
    ``embeddings``: the five Conv1d modules in the same order.  Returns ``(dec_input [B,maxF,H], mel_len)``."""
    lr = length_regulator if length_regulator is not None else LengthRegulator()
    x_expanded, mel_len = lr(x, durations)                           # :226

    def expand_feat(f, d):                                           # :228-230
        expanded, _ = lr(f.unsqueeze(-1), d)
        return expanded.transpose(1, 2)
    ex = [expand_feat(c, durations) for c in curves]                 # :232-236
    if clamps is not None:
        ex = [torch.clamp(e, lo, hi) for e, (lo, hi) in zip(ex, clamps)]   # :239-243
    dec_input = x_expanded.transpose(1, 2)                           # :245
    for emb, e in zip(embeddings, ex):                               # :246-251 (left-to-right sum)
        dec_input = dec_input + emb(e)
    return dec_input.transpose(1, 2), mel_len                        # :252
