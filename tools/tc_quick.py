"""tcgen05 mel GEMM: time + agreement with the fused FFMA log-mel on the cfg4 corpus (gpurun scratch tool)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import spev_tts_b200 as sp
from spev_tts_b200 import _lib, cache
from tests import synth
dev = torch.device("cuda:0")
n_utts = int(sys.argv[1]) if len(sys.argv) > 1 else 13100
lens = synth.utterance_lengths(seed=4, n_utts=n_utts)
starts = cache.aligned_offsets(lens)
g = torch.Generator(device=dev).manual_seed(4)
x = torch.empty(int(starts[-1]), device=dev)
for s0 in range(0, x.numel(), 1 << 27):
    e0 = min(x.numel(), s0 + (1 << 27))
    x[s0:e0] = torch.randn(e0 - s0, generator=g, device=dev) * 0.05
ctx = sp.Context.get(dev)
fb = sp.make_batch(ctx, n_samples=lens, sample_off=starts)
F = fb.n_frames
ref = torch.empty((F, 80), device=dev)
sp.logmel_flat(x, lens, out=ref, batch=fb)
P = torch.empty((F, _lib.SPEC_LD), device=dev)
out = torch.empty((F, 80), device=dev)
st = torch.cuda.current_stream(dev).cuda_stream
_lib.check(ctx.lib.spev_stft_power(ctx.handle, fb.desc, x.data_ptr(), P.data_ptr(), st))
def go():
    _lib.check(ctx.lib.spev_mel_project(ctx.handle, P.data_ptr(), F, out.data_ptr(), 1, 1e-5, -10.0, 2.0, st))
for _ in range(3): go()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): go()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
print(f"mel_project {ms:.3f} ms  {F * 2400 / ms / 1e6:.0f} GB/s  max|diff| vs fused {float((out - ref).abs().max()):.2e}")
