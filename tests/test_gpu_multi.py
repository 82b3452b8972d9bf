"""Multi-GPU end-to-end correctness (needs >= 2 GPUs; skipped on a 1-GPU box): utterance sharding,
per-rank fused kernel, NCCL gather of the variable-length shards and re-assembly must reproduce the
single-GPU cache bit for bit (the path has no data-path collective, so there is nothing to reorder)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import synth

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import spev_tts_b200 as sp
    from spev_tts_b200 import cache
    lens = synth.utterance_lengths(seed=33, n_utts=200)
    rng = np.random.default_rng(33)
    ys = [(0.05 * rng.standard_normal(int(n))).astype(np.float32) for n in lens]     # same corpus on every rank
    shards = cache.shard_utterances(lens, world)
    mine = shards[rank]
    starts = cache.aligned_offsets(lens[mine])
    buf = np.zeros(int(starts[-1]), np.float32)
    for s, u in zip(starts[:-1], mine):
        buf[s: s + lens[u]] = ys[u]
    local, _ = sp.logmel_flat(torch.from_numpy(buf).to(dev), lens[mine], sample_off=starts)
    parts, counts = cache.gather_shards(local, dst=0)
    ok = True
    if rank == 0:
        full, fo = cache.assemble(parts, shards, lens)
        flat = torch.from_numpy(np.concatenate(ys)).to(dev)
        ref, fb = sp.logmel_flat(flat, lens)
        ok = bool(torch.equal(full, ref)) and np.array_equal(fo, fb.frame_off) and counts == [int((1 + lens[s] // 256).sum()) for s in shards]
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_sharded_build_gather_assemble_equals_single_gpu():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res == {0: True, 1: True}
