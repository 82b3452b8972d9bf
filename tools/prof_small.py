"""ncu target (gpurun scratch tool): one or two launches of every small / auxiliary kernel on its benchmark shape --
LengthRegulator plan / expand / backward (cfg2 and B=512), bucketize+embed, fused variance adaptor forward / backward,
tcgen05 GEMMs (mel projection on 262 k frames, mel->magnitude), Griffin-Lim kernels (cfg3, 3 iterations)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import spev_tts_b200 as sp
from spev_tts_b200 import _lib
from tests import synth

dev = torch.device("cuda:0")
x, dur, _ = synth.cfg2_batch(seed=2)
feats = synth.cfg2_features(seed=2)
xd, dd = torch.from_numpy(x).to(dev), torch.from_numpy(dur).to(dev)
fd = [torch.from_numpy(f).to(dev) for f in feats]
for _ in range(2):                                   # cfg2: plan + fused expand, and the backward
    xg = xd.clone().requires_grad_(True)
    fg = [f.clone().requires_grad_(True) for f in fd]
    o, ml, cv = sp.regulate_variances(xg, dd, fg)
    (o.sum() + sum(c.sum() for c in cv)).backward()
rng = np.random.default_rng(12)                      # B = 512
xb = torch.from_numpy(rng.standard_normal((512, 200, 256)).astype(np.float32)).to(dev).requires_grad_(True)
db = torch.from_numpy(rng.integers(0, 21, (512, 200))).to(dev)
ob, _ = sp.LengthRegulator()(xb, db)
ob.sum().backward()
v, bins, table = synth.bucketize_case(seed=2)        # bucketize + embed, phone level and frame level
vt, bt, tt = (torch.from_numpy(a).to(dev) for a in (v, bins, table))
sp.bucketize_embed(vt, bt, tt)
sp.bucketize_embed(torch.randn(32, 2000, device=dev), bt, tt)
embs = [torch.nn.Conv1d(1, 256, 3, padding=1).to(dev) for _ in range(5)]
for _ in range(2):                                   # fused variance adaptor forward + backward
    xg = xd.clone().requires_grad_(True)
    fg = [f.clone().requires_grad_(True) for f in fd]
    out, _ = sp.variance_adaptor(xg, dd, fg, embs)
    out.sum().backward()
ctx = sp.Context.get(dev)                            # tcgen05 mel projection (A/B path) on 262,144 frames
F = 1 << 18
power = torch.rand(F, _lib.SPEC_LD, device=dev)
lm = sp.mel_project(power)
ctx8 = sp.Context.get(dev, fmin=0.0, fmax=8000.0)    # mel -> magnitude (tcgen05 and FFMA) + Griffin-Lim, cfg3
fb = ctx8.uniform_batch(16, 800)
lmel = (-4 + 2 * torch.randn(16 * 800, 80, device=dev)).clamp(-10, 2)
S = sp.mel_to_mag_flat(lmel.view(-1), fb, ctx8, layout=0, is_log=True)
S1 = sp.mel_to_mag_flat(lmel.view(16, 800, 80).transpose(1, 2).contiguous().view(-1), fb, ctx8, layout=1, is_log=True)
y = sp.griffinlim_flat(S, fb, ctx8, n_iter=3, seed=1)
torch.cuda.synchronize()
print("ok", float(y.abs().mean()), float(lm.mean()))
