"""Where the wall time of Vocoder.infer goes on cfg3 (gpurun scratch tool)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import spev_tts_b200 as sp
from spev_tts_b200 import spectral
from tests import synth
dev = torch.device("cuda:0")
B, T = 16, 800
ys = np.stack([synth.speechy(seed=300 + b, n=(T - 1) * 256) for b in range(B)])
lm = sp.logmel(torch.from_numpy(ys).to(dev)).transpose(1, 2).contiguous().cpu().pin_memory()
voc = sp.Vocoder(n_iter=60, device=dev)
for nn in ("librosa", "pinv"):
    voc.nnls = nn
    voc.infer(lm); voc.infer(lm)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(10):
        w = voc.infer(lm)
    print(f"Vocoder.infer nnls={nn}: {(time.perf_counter() - t) * 100:.3f} ms per call")
# stages
import cProfile, pstats
voc.nnls = "librosa"
pr = cProfile.Profile(); pr.enable()
for _ in range(10):
    voc.infer(lm)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
