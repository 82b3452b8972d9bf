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


def test_plan_shards_layout():
    from spev_tts_b200 import cache
    lens = synth.utterance_lengths(seed=4, n_utts=1310)
    for world in (1, 2, 8):
        p = cache.plan_shards(lens, world, n_chunks=4)
        cover = np.zeros(p.n_rows, dtype=np.int64)
        for u in range(len(lens)):
            cover[p.utt_row[u]: p.utt_row[u] + p.frames[u]] += 1
        assert (cover == 1).all() and p.n_rows == cache.frames_of(lens).sum()
        for r in range(world):                       # chunks tile the rank's block of rows, in order
            rows = [p.chunk_rows(r, k) for k in range(len(p.chunks[r]))]
            assert rows[0][0] == p.row_off[r] and rows[-1][1] == p.row_off[r + 1]
            assert all(a[1] == b[0] for a, b in zip(rows[:-1], rows[1:]))
            sizes = np.array([hi - lo for lo, hi in rows])
            assert sizes.max() - sizes.min() <= 2 * cache.frames_of(lens).max()
    one = cache.plan_shards([3000], 4, n_chunks=4)   # more ranks than utterances: empty shards are fine
    assert one.n_rows == 12 and [len(s) for s in one.shards].count(0) == 3


def _pipeline_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from spev_tts_b200 import cache
    lens = synth.utterance_lengths(seed=7, n_utts=41)
    plan = cache.plan_shards(lens, world, n_chunks=3)
    fr = plan.frames

    def kernel(samples, out_rows, a, b):             # stand-in rows: (utterance, frame) so placement can be verified
        idx = plan.shards[rank][a:b]
        rows = [torch.tensor([[float(u), float(t), 0.0, 0.0] for t in range(fr[u])]) for u in idx]
        out_rows.copy_(torch.cat(rows) if rows else torch.zeros((0, 4)))
    ok = True
    for overlap in (False, True):
        bld = cache.ShardedCacheBuilder(plan, rank, "cpu", n_mels=4, kernel=kernel)
        out = torch.full((plan.n_rows if rank == 0 else bld.n_rows_local, 4), -1.0)
        bld.build(None, out, gather=True, overlap=overlap)
        if rank == 0:
            g = cache.GatheredCache(out, plan)
            for u in range(len(lens)):
                seg = g.utterance(u)
                ok &= seg.shape[0] == fr[u] and bool((seg[:, 0] == u).all()) and bool((seg[:, 1] == torch.arange(float(fr[u]))).all())
        ok &= bld.launches == (len(plan.chunks[rank]) if overlap else 1)
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_builder_pipeline_world2_gloo():
    """ShardedCacheBuilder.build over gloo (host logic of the chunked compute/gather overlap): every utterance's rows
    land where the plan says, serial and overlapped."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_pipeline_worker, args=(r, 2, port, q)) for r in range(2)]
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


def _fake_builder(waves, phones, durs, stats, *, sr=22050, device=None):
    """CPU stand-in for build_records (which needs a GPU): drops items shorter than 4000 samples like the real one,
    returns records whose content encodes the waveform so that the merge can be checked."""
    recs, vocab = [], {"<PAD>", "<UNK>", "<SIL>"}
    for i, (w, ph, du) in enumerate(zip(waves, phones, durs)):
        if len(w) < 4000:
            continue
        vocab.update(ph)
        T = 1 + len(w) // 256
        recs.append({"index": i, "phs": list(ph), "durs": [T], "mel": torch.full((T, 80), float(w[0])),
                     **{k: np.array([float(w[0])]) for k in ("pitch", "energy", "breath", "rough", "bright")}})
    return recs, sorted(vocab)


def _sharded_build_worker(rank, world, port, tmp, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import spev_tts_b200 as sp
    lens = [5000, 3000, 9000, 4100, 12000, 4000, 2000, 7000, 8000]
    waves = [np.full(n, float(i), np.float32) for i, n in enumerate(lens)]
    phones = [["<SIL>", chr(97 + i)] for i in range(len(lens))]
    durs = [[1, 1]] * len(lens)
    stats = {"p_mean": 5.0, "p_std": 0.3, "e_mean": -3.0, "e_std": 1.0, "c_mean": 7.0, "c_std": 0.5}
    files, st, vocab = sp.build_cache_sharded(tmp, waves, phones, durs, stats, builder=_fake_builder)
    q.put((rank, [os.path.basename(f) for f in files], vocab))
    dist.destroy_process_group()


def test_sharded_cache_build_merges_ranks(tmp_path):
    """world_size 2 over gloo: every rank writes its shard's files, rank 0 the merged metadata (wav order, gaps)."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sharded_build_worker, args=(r, world, port, str(tmp_path), q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [f"u_{i:05d}.pt" for i in (0, 2, 3, 4, 5, 7, 8)]            # items 1 and 6 are shorter than 4000 samples
    want_vocab = sorted({"<PAD>", "<UNK>", "<SIL>"} | {chr(97 + i) for i in (0, 2, 3, 4, 5, 7, 8)})
    for _, files, vocab in res:
        assert files == want and vocab == want_vocab
    import spev_tts_b200 as sp
    recs, stats, vocab = sp.read_reference_cache(str(tmp_path))
    assert len(recs) == 7 and vocab == want_vocab and stats["p_mean"] == 5.0
    for r, i in zip(recs, (0, 2, 3, 4, 5, 7, 8)):
        assert float(r["mel"][0, 0]) == float(i) and r["phs"] == ["<SIL>", chr(97 + i)]   # file i holds utterance i
