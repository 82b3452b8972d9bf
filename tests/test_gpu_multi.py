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


def _pipeline_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import spev_tts_b200 as sp
    from spev_tts_b200 import cache
    lens = synth.utterance_lengths(seed=34, n_utts=300)
    rng = np.random.default_rng(34)
    ys = [(0.05 * rng.standard_normal(int(n))).astype(np.float32) for n in lens]     # same corpus on every rank
    plan = cache.plan_shards(lens, world, n_chunks=3)
    ok = True
    ref = fo = None
    if rank == 0:
        ref, fb = sp.logmel_flat(torch.from_numpy(np.concatenate(ys)).to(dev), lens)   # one GPU, whole corpus
        fo = fb.frame_off
    for transport in ("nccl", "p2p"):
        bld = cache.ShardedCacheBuilder(plan, rank, dev, dst=0, transport=transport, reserve_sms=16)
        buf = np.zeros(int(bld.sample_off[-1]), np.float32)
        for s0, u in zip(bld.sample_off[:-1], plan.shards[rank]):
            buf[s0: s0 + lens[u]] = ys[u]
        samples = torch.from_numpy(buf).to(dev)
        out = bld.alloc_out()
        for overlap in (False, True, True):                      # the last one back to back with the previous step
            if rank == 0:
                out.fill_(float("nan"))
            bld.build(samples, out, gather=True, overlap=overlap)
            torch.cuda.synchronize(dev)
            dist.barrier()
            if rank == 0:
                gc = cache.GatheredCache(out, plan)
                full, fo2 = gc.corpus_order()
                ok = ok and bool(torch.equal(full, ref)) and np.array_equal(fo2, fo)
                ok = ok and all(bool(torch.equal(gc.utterance(u), ref[fo[u]: fo[u + 1]])) for u in (0, 7, 150, 299))
            dist.barrier()
        del out, bld
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_sharded_builder_both_transports_equal_single_gpu():
    """ShardedCacheBuilder.build (kernel per chunk + gather into rank 0), NCCL and p2p (symmetric-memory window, copy-engine
    pushes), serial and overlapped: the root's gathered cache == the single-GPU cache of the whole corpus, bit for bit."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_pipeline_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res == {0: True, 1: True}


def _records_worker(rank, world, port, tmp, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import spev_tts_b200 as sp
    corpus = synth.tiny_corpus(seed=21)
    waves = [it["y"] for it in corpus]
    phones, durs = synth.corpus_alignments(corpus)
    stats = {"p_mean": 5.4, "p_std": 0.35, "e_mean": -4.6, "e_std": 2.9, "c_mean": 8.0, "c_std": 0.88}
    files, _, vocab = sp.build_cache_sharded(os.path.join(tmp, "sharded"), waves, phones, durs, stats, device=dev)
    ok = True
    if rank == 0:
        recs, vocab1 = sp.build_records(waves, phones, durs, stats, device=dev)          # one process, one GPU
        got, _, vocab2 = sp.read_reference_cache(os.path.join(tmp, "sharded"))
        ok = vocab == vocab1 == vocab2 and len(got) == len(recs)
        ok = ok and [os.path.basename(f) for f in files] == [f"u_{r['index']:05d}.pt" for r in recs]
        for a, b in zip(got, recs):
            ok = ok and a["phs"] == b["phs"] and a["durs"] == b["durs"] and torch.equal(a["mel"], b["mel"])
            ok = ok and all(np.array_equal(a[k], b[k]) for k in ("pitch", "energy", "breath", "rough", "bright"))
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_sharded_record_build_equals_single_gpu(tmp_path):
    """Whole cache records (log-mel, pYIN, pooling) built by two ranks == built by one, bit for bit."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_records_worker, args=(r, world, port, str(tmp_path), q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res == {0: True, 1: True}
