"""Log-mel training-cache builder: the B200 replacement of the mel lines of the reference's
cache loop (``/root/reference/spev_real_metrics.py:330-426``; mel statements ``:363-367``,
stored layout ``'mel': mel.T`` ``:421``).

* ``shard_utterances``  -- length-balanced (greedy LPT) assignment of utterances to ranks.  The
  path is embarrassingly data-parallel: no collective runs inside or between kernels.
* ``build_logmel_cache`` -- host buffers in, host buffers out: chunked H2D copy / fused kernel /
  D2H copy on three streams with double buffering (this is the end-to-end path a caller with
  wavs in host memory uses; ``bench.py`` times it as ``e2e``).
* ``gather_shards``      -- the only cross-GPU step: variable-length gather of the ``[F_r, 80]``
  shards (NCCL on GPUs over NVLink; gloo in the CPU tests).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib, spectral
from .batch import HOP, Context, make_batch


# ---------------------------------------------------------------------------------------------
# host placement: pinned staging buffers should live on the NUMA node the GPU hangs off
# ---------------------------------------------------------------------------------------------
def _parse_cpulist(text: str):
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.extend(range(int(a), int(b or a) + 1))
    return cpus


def bind_to_gpu_numa_node(device) -> dict:
    """Pin the calling process to the CPUs of the NUMA node that `device`'s PCIe root belongs to
    (Linux sysfs).  Call it BEFORE allocating pinned host buffers: first-touch then places them on
    that node, so that H2D/D2H copies of several ranks do not cross sockets.  Best effort: returns
    what it did and never raises."""
    import os
    info = {"bound": False}
    try:
        dev = torch.device(device)
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        props = torch.cuda.get_device_properties(idx)
        bdf = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        info.update(pci=bdf, numa_node=node)
        if node < 0:
            return info
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set(_parse_cpulist(f.read())) & set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            info.update(bound=True, n_cpus=len(cpus))
    except Exception as e:   # sysfs layout / permissions differ: keep going unbound
        info["error"] = repr(e)
    return info


# ---------------------------------------------------------------------------------------------
# sharding (host logic; CPU-testable)
# ---------------------------------------------------------------------------------------------
def frames_of(n_samples) -> np.ndarray:
    return 1 + np.asarray(n_samples, dtype=np.int64) // HOP


def shard_utterances(n_samples: Sequence[int], world_size: int) -> List[np.ndarray]:
    """Greedy longest-processing-time assignment balancing the number of frames per rank.
    Returns, per rank, the sorted utterance indices it owns.  Deterministic."""
    fr = frames_of(n_samples)
    order = np.argsort(-fr, kind="stable")
    loads = np.zeros(world_size, dtype=np.int64)
    owner = np.empty(len(fr), dtype=np.int64)
    for i in order:
        r = int(np.argmin(loads))
        owner[i] = r
        loads[r] += fr[i]
    return [np.sort(np.nonzero(owner == r)[0]) for r in range(world_size)]


def aligned_offsets(n_samples: Sequence[int], align: int = 4) -> np.ndarray:
    """Item starts ``[U+1]`` (last entry = buffer size) with every start a multiple of ``align``
    samples, so that the kernels take the 16-byte cp.async staging path.  The <= align-1 filler
    samples between items are never read as signal."""
    ns = np.asarray(n_samples, dtype=np.int64)
    padded = (ns + align - 1) // align * align
    return np.concatenate([[0], np.cumsum(padded)]).astype(np.int64)


@dataclass
class CachePlan:
    """Chunks of consecutive utterances sized to ~``chunk_samples`` samples."""
    n_samples: np.ndarray
    sample_off: np.ndarray      # [U+1] item starts in the host buffer (last = end of buffer)
    frame_off: np.ndarray       # [U+1]
    chunks: List[Tuple[int, int]]   # (first utt, last utt exclusive)


def plan_chunks(n_samples: Sequence[int], chunk_samples: int = 1 << 26,
                sample_off: Optional[np.ndarray] = None) -> CachePlan:
    ns = np.asarray(n_samples, dtype=np.int64)
    so = np.concatenate([[0], np.cumsum(ns)]) if sample_off is None else np.asarray(sample_off, dtype=np.int64)
    assert so.shape == (len(ns) + 1,) and np.all(so[1:] - so[:-1] >= ns)
    fo = np.concatenate([[0], np.cumsum(frames_of(ns))])
    chunks, a = [], 0
    while a < len(ns):
        b = int(np.searchsorted(so, so[a] + chunk_samples, side="right")) - 1
        b = max(b, a + 1)
        b = min(b, len(ns))
        chunks.append((a, b))
        a = b
    return CachePlan(ns, so, fo, chunks)


# ---------------------------------------------------------------------------------------------
# end-to-end builder: pinned host -> device -> pinned host, pipelined
# ---------------------------------------------------------------------------------------------
class LogMelCacheBuilder:
    """Reusable pipelined builder (device staging buffers and streams are allocated once)."""

    def __init__(self, device, *, sr=22050, n_mels=80, chunk_samples: int = 1 << 26, n_buffers: int = 2):
        self.device = torch.device(device)
        self.sr, self.n_mels = sr, n_mels
        self.chunk_samples = chunk_samples
        self.n_buffers = n_buffers
        self.ctx = Context.get(self.device, sr=sr, n_mels=n_mels)
        self.copy_in = torch.cuda.Stream(self.device)
        self.compute = torch.cuda.Stream(self.device)
        self.copy_out = torch.cuda.Stream(self.device)
        self._pcm: List[Optional[torch.Tensor]] = [None] * n_buffers
        self._in: List[Optional[torch.Tensor]] = [None] * n_buffers
        self._out: List[Optional[torch.Tensor]] = [None] * n_buffers
        # per-buffer "free again" events live across build() calls: a second build() enqueued without a host
        # synchronise must not start copying into a staging buffer the previous build's kernels still read
        self._in_free: List[Optional[torch.cuda.Event]] = [None] * n_buffers    # compute finished reading buffer i
        self._out_free: List[Optional[torch.cuda.Event]] = [None] * n_buffers   # D2H finished reading buffer i
        self.launches = 0

    def _buf(self, pool, i, n, dtype=torch.float32):
        if pool[i] is None or pool[i].numel() < n:
            pool[i] = torch.empty(max(n, 1), dtype=dtype, device=self.device)
        return pool[i]

    def build(self, samples_host: torch.Tensor, n_samples: Sequence[int],
              out_host: Optional[torch.Tensor] = None, plan: Optional[CachePlan] = None,
              sample_off: Optional[np.ndarray] = None):
        """``samples_host``: flat host tensor (pinned for full copy speed) holding the utterances at
        ``sample_off`` (default: packed back to back; ``aligned_offsets`` gives the fast layout),
        either float32 or **int16 PCM** (shipped as 2 bytes/sample and widened to ``pcm/32768`` on the
        device -- the exact values a 16-bit wav decodes to); returns ``(out_host [F, n_mels] pinned,
        frame_off)``."""
        pcm = samples_host.dtype == torch.int16
        if not pcm and samples_host.dtype != torch.float32:
            raise ValueError("samples_host must be float32 or int16 PCM")
        if plan is None:
            plan = plan_chunks(n_samples, self.chunk_samples, sample_off)
        F = int(plan.frame_off[-1])
        if out_host is None:
            out_host = torch.empty((F, self.n_mels), dtype=torch.float32).pin_memory()
        nb = self.n_buffers
        in_free, out_free = self._in_free, self._out_free
        max_in = max(int(plan.sample_off[b] - plan.sample_off[a]) for a, b in plan.chunks)
        max_out = max(int(plan.frame_off[b] - plan.frame_off[a]) for a, b in plan.chunks) * self.n_mels
        for i in range(nb):
            self._buf(self._in, i, max_in)
            self._buf(self._out, i, max_out)
            if pcm:
                self._buf(self._pcm, i, max_in, torch.int16)
        # descriptors are tiny; build them all up front so the loop only enqueues
        batches = [make_batch(self.ctx, n_samples=plan.n_samples[a:b],
                              sample_off=plan.sample_off[a:b] - plan.sample_off[a]) for a, b in plan.chunks]
        tables_up = torch.cuda.Event()
        tables_up.record(torch.cuda.current_stream(self.device))   # descriptor uploads and staging allocations ran here
        self.compute.wait_event(tables_up)
        self.copy_in.wait_event(tables_up)                         # (a re-used allocation may still be in use upstream)
        self.copy_out.wait_event(tables_up)
        for ci, (a, b) in enumerate(plan.chunks):
            i = ci % nb
            s0, s1 = int(plan.sample_off[a]), int(plan.sample_off[b])
            f0, f1 = int(plan.frame_off[a]), int(plan.frame_off[b])
            d_in = self._in[i][: s1 - s0]
            d_out = self._out[i][: (f1 - f0) * self.n_mels].view(f1 - f0, self.n_mels)
            with torch.cuda.stream(self.copy_in):
                if in_free[i] is not None:          # also set by a previous build()
                    self.copy_in.wait_event(in_free[i])
                (self._pcm[i][: s1 - s0] if pcm else d_in).copy_(samples_host[s0:s1], non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(self.copy_in)
            with torch.cuda.stream(self.compute):
                self.compute.wait_event(ready)
                if out_free[i] is not None:
                    self.compute.wait_event(out_free[i])
                if pcm:
                    _lib.check(self.ctx.lib.spev_pcm16_to_f32(self._pcm[i].data_ptr(), s1 - s0, d_in.data_ptr(),
                                                              self.compute.cuda_stream), "spev_pcm16_to_f32")
                    self.launches += 1
                spectral.logmel_flat(d_in, plan.n_samples[a:b], sr=self.sr, n_mels=self.n_mels,
                                     out=d_out, batch=batches[ci])
                self.launches += 1
                in_free[i] = torch.cuda.Event()
                in_free[i].record(self.compute)
                done = torch.cuda.Event()
                done.record(self.compute)
            with torch.cuda.stream(self.copy_out):
                self.copy_out.wait_event(done)
                out_host[f0:f1].copy_(d_out, non_blocking=True)
                out_free[i] = torch.cuda.Event()
                out_free[i].record(self.copy_out)
        fin = torch.cuda.Event()
        fin.record(self.copy_out)
        torch.cuda.current_stream(self.device).wait_event(fin)
        self._keep = batches
        return out_host, plan.frame_off


def build_logmel_cache(samples_host: torch.Tensor, n_samples: Sequence[int], device=None,
                       sample_off: Optional[np.ndarray] = None, **kw):
    """One-shot convenience wrapper around ``LogMelCacheBuilder`` (synchronises before returning)."""
    device = spectral.default_device() if device is None else torch.device(device)
    out, fo = LogMelCacheBuilder(device, **kw).build(samples_host, n_samples, sample_off=sample_off)
    torch.cuda.synchronize(device)
    return out, fo


# ---------------------------------------------------------------------------------------------
# the one cross-rank step: gather the shards
# ---------------------------------------------------------------------------------------------
def gather_shards(local: torch.Tensor, dst: int = 0, group=None):
    """Variable-length gather of ``[F_r, n_mels]`` shards to rank ``dst``.
    Returns ``(list of per-rank tensors on dst | None elsewhere, frames_per_rank)``.
    Uses an all_gather of the row counts followed by point-to-point transfers (NCCL has no
    gatherv); works on the NCCL (GPU, NVLink) and gloo (CPU tests) backends."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    if rank == dst:
        parts = [local if r == dst else torch.empty((counts[r], local.shape[1]), dtype=local.dtype,
                                                    device=local.device) for r in range(world)]
        ops = [dist.P2POp(dist.irecv, parts[r], r, group) for r in range(world) if r != dst and counts[r]]
        if ops:
            for q in dist.batch_isend_irecv(ops):   # one grouped NCCL launch: transfers run concurrently
                q.wait()
        return parts, counts
    if local.shape[0]:
        for q in dist.batch_isend_irecv([dist.P2POp(dist.isend, local.contiguous(), dst, group)]):
            q.wait()
    return None, counts


def assemble(parts: Sequence[torch.Tensor], shards: Sequence[np.ndarray], n_samples: Sequence[int]):
    """Re-order gathered shards into corpus order.  Returns ``(cache [F, n_mels], frame_off)``
    such that utterance ``u`` is ``cache[frame_off[u]:frame_off[u+1]]`` -- i.e. the tensor each
    ``cache_stable/u_%05d.pt`` would hold under ``'mel'`` (``spev_real_metrics.py:419-425``)."""
    fr = frames_of(n_samples)
    fo = np.concatenate([[0], np.cumsum(fr)])
    out = torch.empty((int(fo[-1]), parts[0].shape[1]), dtype=parts[0].dtype, device=parts[0].device)
    for part, idx in zip(parts, shards):
        if len(idx) == 0:
            continue
        loc = np.concatenate([[0], np.cumsum(fr[idx])])
        dest = np.repeat(fo[idx] - loc[:-1], fr[idx]) + np.arange(loc[-1])
        out[torch.from_numpy(dest).to(out.device)] = part
    return out, fo
