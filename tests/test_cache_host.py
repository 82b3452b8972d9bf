"""CPU tests of the multi-GPU host logic: utterance sharding, chunk planning and the
variable-length shard gather (world_size 2 over gloo)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import synth


def test_shard_utterances_partition_and_balance():
    from spev_tts_b200 import cache
    lens = synth.utterance_lengths(seed=4, n_utts=13100)
    for world in (1, 2, 4, 8):
        shards = cache.shard_utterances(lens, world)
        allidx = np.sort(np.concatenate(shards))
        assert np.array_equal(allidx, np.arange(len(lens)))           # exact partition
        loads = np.array([cache.frames_of(lens[s]).sum() for s in shards])
        assert loads.max() - loads.min() <= cache.frames_of(lens).max()   # LPT bound
        assert loads.max() / loads.mean() < 1.001
        again = cache.shard_utterances(lens, world)
        assert all(np.array_equal(a, b) for a, b in zip(shards, again))   # deterministic


def test_plan_chunks_covers_everything():
    from spev_tts_b200 import cache
    lens = synth.utterance_lengths(seed=5, n_utts=500)
    plan = cache.plan_chunks(lens, chunk_samples=1 << 22)
    assert plan.chunks[0][0] == 0 and plan.chunks[-1][1] == 500
    for (a, b), (c, d) in zip(plan.chunks[:-1], plan.chunks[1:]):
        assert b == c and b > a
    sizes = [plan.sample_off[b] - plan.sample_off[a] for a, b in plan.chunks]
    assert max(sizes) <= (1 << 22) + lens.max()
    assert plan.frame_off[-1] == (1 + lens // 256).sum()
    one = cache.plan_chunks([10 ** 9], chunk_samples=1 << 20)        # a single oversize utterance
    assert one.chunks == [(0, 1)]


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _gather_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from spev_tts_b200 import cache
    lens = synth.utterance_lengths(seed=6, n_utts=37)
    shards = cache.shard_utterances(lens, world)
    fr = cache.frames_of(lens)
    fo = np.concatenate([[0], np.cumsum(fr)])
    # stand-in "cache rows": row value encodes (utterance, frame) so assembly can be verified
    mine = shards[rank]
    rows = [torch.stack([torch.full((5,), float(u)), torch.arange(5.0) * 0 + 0]) for u in mine]   # dummy
    local = torch.cat([torch.stack([torch.tensor([float(u), float(t), 0., 0.]) for t in range(fr[u])])
                       for u in mine]) if len(mine) else torch.zeros((0, 4))
    parts, counts = cache.gather_shards(local, dst=0)
    ok = True
    if rank == 0:
        assert counts == [int(fr[s].sum()) for s in shards]
        full, fo2 = cache.assemble(parts, shards, lens)
        ok = np.array_equal(fo2, fo)
        for u in range(len(lens)):
            seg = full[fo[u]: fo[u + 1]]
            ok &= bool((seg[:, 0] == u).all()) and bool((seg[:, 1] == torch.arange(float(fr[u]))).all())
    else:
        assert parts is None
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_shards_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == {0: True, 1: True}


def test_cpulist_parser():
    from spev_tts_b200 import cache
    assert cache._parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert cache._parse_cpulist("") == []
