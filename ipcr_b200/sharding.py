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


SITE_CAP = 4096  # sites per rank that travel in the single fixed-size collective (96 KB)


def gather_sites(dist, sites, records, n_recs: int, rank: int, world: int, device=None):
    """All ranks call this with their numpy site array (ipcr_site records) and their view of
    the records [(len, has_reset_byte)].  EVERY rank gets (all sites, merged records) from ONE
    all_gather of a fixed-size block [count | per-record len, flag | up to SITE_CAP sites];
    ranks with more sites than SITE_CAP send the rest in a second, padded round (rare)."""
    import numpy as np
    import torch

    from .device import SITE_DTYPE

    kw = {"device": device} if device is not None else {}
    isz = SITE_DTYPE.itemsize
    hdr_words = 1 + 2 * n_recs
    hdr_bytes = 8 * hdr_words
    block = hdr_bytes + SITE_CAP * isz
    hdr = np.zeros(hdr_words, dtype=np.int64)
    hdr[0] = len(sites)
    if n_recs:
        hdr[1:] = np.asarray(records, dtype=np.int64).reshape(-1)[: 2 * n_recs]
    mine = np.zeros(block, dtype=np.uint8)
    mine[:hdr_bytes] = hdr.view(np.uint8)
    raw = np.ascontiguousarray(sites).view(np.uint8).reshape(-1)
    n_first = min(len(sites), SITE_CAP)
    mine[hdr_bytes:hdr_bytes + n_first * isz] = raw[: n_first * isz]
    send = torch.from_numpy(mine)
    if device is not None:
        send = send.to(device)
    recv = torch.empty(world * block, dtype=torch.uint8, **kw)
    dist.all_gather_into_tensor(recv, send)
    allb = recv.cpu().numpy().reshape(world, block)
    heads = np.ascontiguousarray(allb[:, :hdr_bytes]).view(np.int64).reshape(world, hdr_words)
    counts = [int(c) for c in heads[:, 0]]
    merged_len = heads[:, 1::2].max(axis=0).tolist()
    merged_flag = (heads[:, 2::2] != 0).any(axis=0).tolist()
    parts = []
    for r in range(world):
        k = min(counts[r], SITE_CAP)
        if k:
            parts.append(np.ascontiguousarray(allb[r, hdr_bytes:hdr_bytes + k * isz]).view(SITE_DTYPE))
    extra = max(0, max(counts) - SITE_CAP)
    if extra:  # overflow round, padded to the largest remainder
        rest = np.zeros(extra * isz, dtype=np.uint8)
        if len(sites) > SITE_CAP:
            rest[: (len(sites) - SITE_CAP) * isz] = raw[SITE_CAP * isz:]
        send2 = torch.from_numpy(rest)
        if device is not None:
            send2 = send2.to(device)
        recv2 = torch.empty(world * extra * isz, dtype=torch.uint8, **kw)
        dist.all_gather_into_tensor(recv2, send2)
        all2 = recv2.cpu().numpy().reshape(world, extra * isz)
        for r in range(world):
            k = max(0, counts[r] - SITE_CAP)
            if k:
                parts.append(all2[r, : k * isz].copy().view(SITE_DTYPE))
    merged = list(zip(merged_len, merged_flag))
    allsites = np.concatenate(parts) if parts else np.zeros(0, dtype=SITE_DTYPE)
    return allsites, merged
