"""Multi-GPU sharding of the scan: cut the concatenated records into
per-rank slices with a primer-length halo, run stage 1 per rank, gather the sparse site
lists on rank 0 and pair once.  No data-path collective: only the site lists (a few KB)
and per-record flags travel.  Host logic only; works with any torch.distributed backend
(NCCL on the GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

HALO = 64  # >= IPCR_GPU_MAX_PRIMER_LEN - 1


def contig_offsets(lens: Sequence[int]) -> List[int]:
    offs = [0]
    for ln in lens:
        offs.append(offs[-1] + ln)
    return offs


def rank_slice(total: int, world: int, rank: int, halo: int = HALO) -> Tuple[int, int, int]:
    """(lo, hi, piece_end): the rank owns [lo, hi) and additionally sees [hi, piece_end)."""
    lo = total * rank // world
    hi = total * (rank + 1) // world
    return lo, hi, (min(total, hi + halo) if world > 1 else total)


def shard_segments(lens: Sequence[int], world: int, rank: int, halo: int = HALO):
    """Segments of this rank's slice: (slice_begin, slice_end, rec_idx, rec_off, own_len) in
    concatenated coordinates.  A site belongs to the rank whose [lo, hi) holds its start;
    the halo only ever extends inside the same record."""
    offs = contig_offsets(lens)
    total = offs[-1]
    lo, hi, piece_end = rank_slice(total, world, rank, halo)
    out = []
    for c in range(len(lens)):
        s0, s1 = max(offs[c], lo), min(offs[c + 1], piece_end)
        own1 = min(offs[c + 1], hi)
        if s0 >= s1 or own1 <= s0:
            continue
        out.append((s0, s1, c, s0 - offs[c], own1 - s0))
    return out


def gather_sites(dist, sites, records, n_recs: int, rank: int, world: int, device=None):
    """All ranks call this with their numpy site array (ipcr_site records) and their view of
    the records [(len, has_reset_byte)].  Rank 0 gets (all sites, merged records); others
    (None, merged records)."""
    import numpy as np
    import torch

    from .device import SITE_DTYPE

    kw = {"device": device} if device is not None else {}
    cnt = torch.tensor([len(sites)], dtype=torch.int64, **kw)
    cnts = [torch.zeros(1, dtype=torch.int64, **kw) for _ in range(world)]
    dist.all_gather(cnts, cnt)
    m = max(int(c.item()) for c in cnts)
    pad = torch.zeros(max(1, m) * SITE_DTYPE.itemsize, dtype=torch.uint8, **kw)
    if len(sites):
        raw = torch.from_numpy(np.ascontiguousarray(sites).view(np.uint8).copy())
        pad[: raw.numel()].copy_(raw)
    allp = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
    dist.gather(pad, allp, dst=0)
    flags = torch.tensor([[int(r[0]), int(r[1])] for r in records], dtype=torch.int64, **kw).reshape(n_recs, 2)
    dist.all_reduce(flags, op=dist.ReduceOp.MAX)
    merged = [(int(flags[i, 0].item()), bool(flags[i, 1].item())) for i in range(n_recs)]
    if rank != 0:
        return None, merged
    parts = [allp[r][: int(cnts[r].item()) * SITE_DTYPE.itemsize].cpu().numpy().view(SITE_DTYPE)
             for r in range(world)]
    allsites = np.concatenate(parts) if parts else np.zeros(0, dtype=SITE_DTYPE)
    return allsites, merged
