"""Timing probe for the pYIN stages (gpurun scratch tool): python tools/pyin_probe.py [n_utts]"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from spev_tts_b200 import pitch as gp
from spev_tts_b200.batch import Context, make_batch
from tests import synth

n_utts = int(sys.argv[1]) if len(sys.argv) > 1 else 592
dev = torch.device("cuda:0")
lens = synth.utterance_lengths(4, n_utts)
ys = [synth.voiced_unvoiced(seed=i % 16, n=int(lens[:16].max()))[0] for i in range(16)]
flat = np.concatenate([ys[i % 16][: lens[i]] for i in range(n_utts)])
x = torch.from_numpy(flat).to(dev)
fb = make_batch(Context.get(dev), n_samples=lens)
p = gp.PyinContext.get(dev)
def t(fn, n=5):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, r
ms1, yin = t(lambda: gp.cmnd_flat(x, fb, p))
ms2, (lo, lu, vp) = t(lambda: gp.observe(yin, p))
ms3, _ = t(lambda: gp.decode(lo, lu, fb.frame_off, p))
F = fb.n_frames
print(f"utts {n_utts} frames {F}: cmnd {ms1:.2f} ms ({F/ms1/1e3:.1f} M frames/s)  observe {ms2:.2f} ms ({F/ms2/1e3:.1f} M/s)  "
      f"viterbi {ms3:.2f} ms ({F/ms3/1e3:.1f} M/s)  total {ms1+ms2+ms3:.2f} ms -> {F/(ms1+ms2+ms3)/1e3:.2f} M frames/s, "
      f"{lens.sum()/22050/(ms1+ms2+ms3)*1e3:.0f} audio-s/s")
