"""Small run of every kernel for compute-sanitizer (memcheck / racecheck): tiny, odd-sized inputs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import spev_tts_b200 as sp
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
lens = [0, 1, 255, 300, 8191, 8192 + 257, 33 * 256 + 5, 3]
ys = [(0.1 * rng.standard_normal(n)).astype(np.float32) for n in lens]
flat = torch.from_numpy(np.concatenate(ys)).to(dev)
out, fb = sp.logmel_flat(flat, lens)                                  # scalar staging path
from spev_tts_b200 import cache
starts = cache.aligned_offsets(lens)
buf = torch.zeros(int(starts[-1]) + 8, device=dev)
for s, y in zip(starts[:-1], ys):
    buf[s: s + len(y)] = torch.from_numpy(y).to(dev)
out2, _ = sp.logmel_flat(buf, lens, sample_off=starts)                  # 16-byte staging path
assert torch.equal(out, out2)
p, _ = sp.stft_power_flat(buf, lens) if False else sp.stft_power_flat(flat, lens)
m = sp.mel_project(p)                                                   # tcgen05 GEMM
assert float((m - out).abs().max()) < 1e-4
r, c, _ = sp.frame_features_flat(flat, lens)
pool = sp.segment_pool(r, fb.frame_off, torch.tensor([1] * int(fb.n_frames), device=dev), fb.frame_off)
lm = torch.from_numpy(np.clip(-4 + 2 * rng.standard_normal((2, 80, 37)), -10, 2).astype(np.float32)).to(dev)
y = sp.mel_to_audio(lm, sr=22050, n_fft=1024, hop_length=256, fmin=0, fmax=8000, n_iter=3, is_log=True)
S = sp.mel_to_stft(torch.exp(lm), sr=22050, n_fft=1024, fmin=0, fmax=8000)
X = sp.stft(y, n_fft=1024, hop_length=256)
yi = sp.istft(X, hop_length=256, n_fft=1024)
x = torch.randn(3, 11, 20, device=dev)
d = torch.randint(0, 7, (3, 11), device=dev)
o, ml = sp.LengthRegulator()(x, d)
o2, ml2, cv = sp.regulate_variances(x, d, [torch.randn(3, 11, device=dev) for _ in range(5)])
dr = sp.duration_rule(torch.randn(3, 11, device=dev))
e = sp.bucketize_embed(torch.randn(5, 9, device=dev), torch.linspace(-3, 3, 255, device=dev), torch.randn(256, 12, device=dev))
pcm = torch.randint(-30000, 30000, (1001,), dtype=torch.int16)
o3, _ = cache.build_logmel_cache(pcm.pin_memory(), [1001], device=dev)
torch.cuda.synchronize()
print("sanitize_smoke ok", out.shape, y.shape, o.shape)
