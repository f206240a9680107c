"""world_size-2 gloo test (CPU) of the multi-GPU host logic: shard planning with halo,
site ownership, gather to rank 0.  Stage 1 is played by the ORACLE here (the checker,
no GPU): each rank reports the sites that START in its own slice, exactly what
ipcr_gpu_find_sites_device does on a shard."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, seq, orients, lens, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    import torch.distributed as dist

    from ipcr_b200 import sharding
    from ipcr_b200.device import SITE_DTYPE
    from oracle import oracle as O

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    offs = sharding.contig_offsets(lens)
    sites = []
    for (s0, s1, c, rec_off, own) in sharding.shard_segments(lens, world, rank):
        piece = seq[s0:s1]
        for qi, (s, rc) in enumerate(orients):
            for pos, mm, idx in O.find_matches(piece, s, 2, 0, 0 if rc else 3, left_tw=3 if rc else 0):
                if pos < own:  # only sites that start inside the owned part
                    mask = sum(1 << j for j in idx)
                    sites.append((rec_off + pos, mask, c, qi, mm, 0))
    arr = np.array(sites, dtype=SITE_DTYPE) if sites else np.zeros(0, dtype=SITE_DTYPE)
    recs = [(0, False)] * len(lens)
    for (s0, s1, c, rec_off, own) in sharding.shard_segments(lens, world, rank):
        has_n = b"N" in seq[s0:s1]
        recs[c] = (max(recs[c][0], rec_off + (s1 - s0)), recs[c][1] or has_n)
    allsites, merged = sharding.gather_sites(dist, arr, recs, len(lens), rank, world)
    q.put((rank, sorted(map(tuple, allsites.tolist())), merged))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_whole_scan(oracle):
    import random

    import ipcr_test_helpers as H

    rng = random.Random(8)
    lens = [7000, 1200, 9000, 300]
    total = sum(lens)
    g = H.random_genome(rng, total, n_frac=0.002)
    a, b = "AGAGTTTGATCMTGGCTCAG", "TACGGYTACCTTGTTAYGACTT"
    offs = [0, 7000, 8200, 17200]
    for k, pos in enumerate([100, 3000, 8300, 8600 + 20, 12000, 16000]):
        H.plant_amplicon(rng, g, pos, a, b, 150, k % 3, (k + 1) % 3, 3, revcomp_strand=k % 2 == 1)
    # a site straddling the rank cut (total // 2 = 8750 lies in contig 2)
    g[8750 - 9:8750 + 11] = a.replace("M", "C").encode()
    seq = bytes(g)
    orients = [(a, False), (b, False), (H.revcomp(a), True), (H.revcomp(b), True)]
    want = []
    for c in range(len(lens)):
        rec = seq[offs[c]:offs[c] + lens[c]]
        for qi, (s, rc) in enumerate(orients):
            for pos, mm, idx in oracle.find_matches(rec, s, 2, 0, 0 if rc else 3, left_tw=3 if rc else 0):
                want.append((pos, sum(1 << j for j in idx), c, qi, mm, 0))
    assert len(want) >= 12
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, seq, orients, lens, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in results) == [0, 1]
    for _rank, got, merged in results:  # every rank ends up with the same global view
        assert got == sorted(want)
        assert [m[0] for m in merged] == lens
        assert [m[1] for m in merged] == [b"N" in seq[offs[c]:offs[c] + lens[c]] for c in range(4)]


def test_gather_overflow_round(monkeypatch):
    """more sites than SITE_CAP on one rank: the second padded round carries the rest"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_overflow_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for n, ok in res:
        assert n == 50 + 7 and ok


def _overflow_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from ipcr_b200 import sharding
    from ipcr_b200.device import SITE_DTYPE

    sharding.SITE_CAP = 16
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 50 if rank == 0 else 7
    arr = np.zeros(n, dtype=SITE_DTYPE)
    arr["pos"] = np.arange(n) + 1000 * rank
    arr["orient"] = rank
    allsites, merged = sharding.gather_sites(dist, arr, [(10 + rank, rank == 1)], 1, rank, world)
    want = sorted([(i, 0) for i in range(50)] + [(1000 + i, 1) for i in range(7)])
    got = sorted((int(s["pos"]), int(s["orient"])) for s in allsites)
    q.put((len(allsites), got == want and merged == [(11, True)]))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_segments_cover_every_start_once():
    from ipcr_b200 import sharding

    lens = [1000, 50, 3000, 7, 1200]
    total = sum(lens)
    for world in (1, 2, 3, 4, 8):
        owned = np.zeros(total, dtype=np.int32)
        for rank in range(world):
            for (s0, s1, c, rec_off, own) in sharding.shard_segments(lens, world, rank):
                offs = sharding.contig_offsets(lens)
                assert s0 - offs[c] == rec_off and s1 <= offs[c + 1]
                assert own <= s1 - s0
                # halo never crosses into another record and is at most HALO long
                assert (s1 - s0) - own <= sharding.HALO
                owned[s0:s0 + own] += 1
        assert (owned == 1).all()
