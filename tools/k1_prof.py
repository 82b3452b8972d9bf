"""ncu target: each fused log-mel variant twice on a small cfg4-shaped shard (python tools/k1_prof.py [n_utts])."""
import sys
import torch
sys.path.insert(0, ".")
import spev_tts_b200 as sp
from spev_tts_b200 import _lib, cache
from tests import synth

n_utts = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
dev = torch.device("cuda:0")
lens = synth.utterance_lengths(seed=4, n_utts=n_utts)
starts = cache.aligned_offsets(lens)
x = torch.randn(int(starts[-1]), device=dev) * 0.05
ctx = sp.Context.get(dev)
fb = sp.make_batch(ctx, n_samples=lens, sample_off=starts)
out = torch.zeros((fb.n_frames, 80), device=dev)
for variant in (0, 1):
    _lib.check(ctx.lib.spev_set_logmel_variant(ctx.handle, variant))
    for _ in range(2):
        sp.logmel_flat(x, lens, out=out, batch=fb)
torch.cuda.synchronize()
print("frames", fb.n_frames)
