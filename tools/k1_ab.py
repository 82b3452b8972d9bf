"""A/B of the fused log-mel kernel variants (gpurun scratch tool): python tools/k1_ab.py [n_utts]
variant 0 = tile lock-step k_stft_mel<0>, variant 1 = decoupled warps k_stft_mel_ws."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import spev_tts_b200 as sp
from spev_tts_b200 import _lib, cache
from tests import synth

n_utts = int(sys.argv[1]) if len(sys.argv) > 1 else 13100
dev = torch.device("cuda:0")
lens = synth.utterance_lengths(seed=4, n_utts=n_utts)
starts = cache.aligned_offsets(lens)
g = torch.Generator(device=dev).manual_seed(4)
x = torch.empty(int(starts[-1]), device=dev)
for s0 in range(0, x.numel(), 1 << 27):
    e0 = min(x.numel(), s0 + (1 << 27))
    x[s0:e0] = torch.randn(e0 - s0, generator=g, device=dev) * 0.05
ctx = sp.Context.get(dev)
fb = sp.make_batch(ctx, n_samples=lens, sample_off=starts)
outs = []
for variant in (0, 1, 0, 1):
    _lib.check(ctx.lib.spev_set_logmel_variant(ctx.handle, variant))
    out = torch.zeros((fb.n_frames, 80), device=dev)
    for _ in range(3):
        sp.logmel_flat(x, lens, out=out, batch=fb)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        sp.logmel_flat(x, lens, out=out, batch=fb)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(f"variant {variant}: {ms:.3f} ms  {fb.n_frames / ms / 1e3:.1f} M frames/s  roofline {1344 * fb.n_frames / ms / 1e6 / 6542.1:.4f}")
    outs.append(out)
print("bit-identical:", bool(torch.equal(outs[0], outs[1])), "max abs diff", float((outs[0] - outs[1]).abs().max()))
