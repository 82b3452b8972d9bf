"""Full-size pYIN run (cfg4: 13,100 utterances, ~6.2 M frames, ~27 GB of intermediates) with size-independent checks:
states in range, flags/f0 consistent, and spot utterances identical to their stand-alone decode."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
import spev_tts_b200 as sp
from spev_tts_b200 import pitch as gp
from tests import synth

dev = torch.device("cuda:0")
n_utts = int(sys.argv[1]) if len(sys.argv) > 1 else 13100
lens = synth.utterance_lengths(4, 13100)[:n_utts]
base = [synth.voiced_unvoiced(seed=i, n=int(lens.max()))[0] for i in range(16)]
starts = sp.cache.aligned_offsets(lens)
host = np.zeros(int(starts[-1]) + 4, np.float32)
for i, (n, s) in enumerate(zip(lens, starts[:-1])):
    host[s: s + n] = base[i % 16][:n]
x = torch.from_numpy(host).to(dev)
torch.cuda.synchronize()
t = time.perf_counter()
f0, flag, vp, fo, states = gp.pyin_flat(x, lens, sample_off=starts[:-1], return_states=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t
F = int(fo[-1])
print(f"{n_utts} utterances, {F} frames, {lens.sum() / 22050 / 3600:.1f} h of audio in {dt:.2f} s wall "
      f"({F / dt / 1e6:.1f} M frames/s, {lens.sum() / 22050 / dt:.0f} audio-s/s), peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
st = states
nb = gp.PyinContext.get(dev).n_bins
assert bool(((st >= 0) & (st < 2 * nb)).all())
assert bool(((st < nb) == flag).all()) and bool((torch.isnan(f0) == ~flag).all())
assert bool(((vp >= 0) & (vp <= 1)).all())
for u in (0, 1, n_utts // 3, n_utts // 2, n_utts - 1):
    y = x[int(starts[u]): int(starts[u]) + int(lens[u])].contiguous()
    _, _, vp1, _, st1 = gp.pyin_flat(y, [int(lens[u])], return_states=True)
    assert torch.equal(st1, st[int(fo[u]): int(fo[u + 1])]) and torch.equal(vp1, vp[int(fo[u]): int(fo[u + 1])]), u
print("full-size checks ok; voiced fraction", float(flag.float().mean()))
