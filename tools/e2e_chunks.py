"""e2e cache build (pinned host in/out) for several chunk sizes / buffer counts (gpurun scratch tool)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from spev_tts_b200 import cache
from tests import synth
dev = torch.device("cuda:0")
lens = synth.utterance_lengths(seed=4, n_utts=13100)
starts = cache.aligned_offsets(lens)
total = int(starts[-1])
host = torch.empty(total, dtype=torch.float32, pin_memory=True)
host.normal_(0, 0.05)
F = int((1 + lens // 256).sum())
out_host = torch.empty((F, 80), dtype=torch.float32, pin_memory=True)
for cs, nb in ((1 << 26, 2), (1 << 26, 3), (1 << 25, 3), (1 << 25, 4), (1 << 24, 4), (1 << 26, 2)):
    b = cache.LogMelCacheBuilder(dev, chunk_samples=cs, n_buffers=nb)
    plan = cache.plan_chunks(lens, cs, starts)
    b.build(host, lens, out_host=out_host, plan=plan); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        b.build(host, lens, out_host=out_host, plan=plan)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"chunk {cs >> 20:3d} Mi samples x {nb} buffers: {ms:7.2f} ms  {F / ms / 1e3:6.1f} M frames/s  ({len(plan.chunks)} chunks)")
