"""Wall-clock timeline of one Vocoder.infer call on cfg3 (gpurun scratch tool)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import spev_tts_b200 as sp
from spev_tts_b200 import spectral as S_, _lib
from spev_tts_b200.batch import Context
from tests import synth
dev = torch.device("cuda:0")
B, T = 16, 800
ys = np.stack([synth.speechy(seed=300 + b, n=(T - 1) * 256) for b in range(B)])
lm = sp.logmel(torch.from_numpy(ys).to(dev)).transpose(1, 2).contiguous().cpu().pin_memory()
voc = sp.Vocoder(n_iter=60, device=dev)
for _ in range(3): voc.infer(lm)
def now(): return time.perf_counter()
for rep in range(3):
    torch.cuda.synchronize()
    t0 = now()
    t, was_numpy = S_._to_device(lm, dev)
    ctx = Context.get(t.device, sr=22050, n_mels=80, fmin=0, fmax=8000)
    batch = ctx.uniform_batch(B, T, with_chunks=True)
    tm = S_.items_to_rows(t.reshape(B, 80, T)).view(-1)
    Sm = S_.mel_to_mag_flat(tm, batch, ctx, layout=0, is_log=True)
    t1 = now()
    n = S_.nnls_refine(Sm, tm.view(B * T, 80), ctx, B, T, is_log=True)
    t2 = now()
    y = S_.griffinlim_flat(Sm, batch, ctx, n_iter=60, seed=1)
    t3 = now()
    out = torch.empty(y.shape, dtype=y.dtype, pin_memory=True)
    t4 = now()
    out.copy_(y, non_blocking=True)
    torch.cuda.current_stream(dev).synchronize()
    t5 = now()
    print(f"prep+mag enqueue {1e3*(t1-t0):.3f}  nnls_refine (sync) {1e3*(t2-t1):.3f}  GL enqueue {1e3*(t3-t2):.3f}  pinned alloc {1e3*(t4-t3):.3f}  D2H+sync {1e3*(t5-t4):.3f}  total {1e3*(t5-t0):.3f} ms")
