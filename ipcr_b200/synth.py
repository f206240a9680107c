"""Deterministic synthetic workloads: panels, contig layouts and
planted amplicons.  Input generation only -- nothing here is on the scan path."""
from __future__ import annotations

import random
from dataclasses import dataclass
from typing import List, Sequence, Tuple

from .engine import Pair

IUPAC_AMBIG = "RYSWKMBDHVN"
_MASKS = {"A": "A", "C": "C", "G": "G", "T": "T", "R": "AG", "Y": "CT", "S": "CG", "W": "AT", "K": "GT",
          "M": "AC", "B": "CGT", "D": "AGT", "H": "ACT", "V": "ACG", "N": "ACGT"}
_COMP = dict(zip("ACGTRYSWKMBVDHN", "TGCAYRSWMKVBHDN"))


def revcomp(s: str) -> str:
    return "".join(_COMP[c] for c in reversed(s))


def multiplex_panel(n_pairs: int, seed: int = 7, self_pairs: bool = True) -> Tuple[List[Pair], List[Pair]]:
    """n random 18-25-nt pairs, 10 % with 1-2 IUPAC codes (config 5).  Returns (user pairs,
    expanded pairs); the expansion appends A:self / B:self once per unique primer like
    ipcr-multiplex (internal/common/primers.go:41-74)."""
    rng = random.Random(seed)

    def mk(n):
        s = [rng.choice("ACGT") for _ in range(n)]
        if rng.random() < 0.1:
            for _ in range(rng.randint(1, 2)):
                s[rng.randrange(n)] = rng.choice(IUPAC_AMBIG)
        return "".join(s)

    base = [Pair(f"p{i:03d}", mk(rng.randint(18, 25)), mk(rng.randint(18, 25))) for i in range(n_pairs)]
    out = list(base)
    if self_pairs:
        seen_a, seen_b = set(), set()
        for p in base:
            if p.forward not in seen_a:
                seen_a.add(p.forward)
                out.append(Pair(p.id + "+A:self", p.forward, p.forward))
            if p.reverse not in seen_b:
                seen_b.add(p.reverse)
                out.append(Pair(p.id + "+B:self", p.reverse, p.reverse))
    return base, out


def pair_16s(self_pairs: bool = True) -> Tuple[List[Pair], List[Pair]]:
    """27F / 1492R (README.md:46) with the default --self expansion
    (internal/common/primers.go:11-38)."""
    a, b = "AGAGTTTGATCMTGGCTCAG", "TACGGYTACCTTGTTAYGACTT"
    base = [Pair("16S", a, b)]
    out = list(base)
    if self_pairs:
        out += [Pair("16S+A:self", a, a), Pair("16S+B:self", b, b)]
    return base, out


def contig_lengths(total: int, n_contigs: int) -> List[int]:
    """n contig lengths summing to `total`, spread between total/60 and total/12 (24 contigs of
    50-250 Mbp at 3 Gbp), deterministic: len_k = lo + (hi-lo) * u_k**g with
    the exponent g chosen so that the sum is exact."""
    if n_contigs == 1:
        return [total]
    lo, hi = total / (2.5 * n_contigs), total / (0.5 * n_contigs)
    u = [((k * 37) % n_contigs) / (n_contigs - 1) for k in range(n_contigs)]
    g_lo, g_hi = 0.05, 50.0
    for _ in range(80):
        g = (g_lo + g_hi) / 2
        if sum(lo + (hi - lo) * x ** g for x in u) > total:
            g_lo = g
        else:
            g_hi = g
    lens = [int(lo + (hi - lo) * x ** g) // 1000 * 1000 for x in u]
    lens[lens.index(max(lens))] += total - sum(lens)
    return lens


@dataclass
class Planted:
    pair_idx: int
    contig: int
    start: int       # contig-local
    length: int
    revcomp_strand: bool
    left: bytes      # bytes written at start
    right: bytes     # bytes written at start + length - len(right)


def _concrete(rng: random.Random, primer: str) -> str:
    return "".join(rng.choice(_MASKS[c]) for c in primer)


def _mutate(rng: random.Random, s: str, k: int, lo: int, hi: int) -> str:
    b = list(s)
    for i in rng.sample(range(lo, hi), min(k, max(0, hi - lo))):
        b[i] = rng.choice([c for c in "ACGT" if c != b[i]])
    return "".join(b)


def plan_amplicons(pairs: Sequence[Pair], contig_lens: Sequence[int], max_mm: int, tw: int, seed: int = 11,
                   per_pair=(1, 3), len_range=(80, 1900)) -> List[Planted]:
    """1-3 amplicons per pair on both strands with 0..max_mm substitutions outside the 3'
    window of each primer.  Positions never overlap."""
    rng = random.Random(seed)
    out: List[Planted] = []
    used = {}
    for i, p in enumerate(pairs):
        for _ in range(rng.randint(*per_pair)):
            for _try in range(100):
                c = rng.randrange(len(contig_lens))
                length = rng.randint(max(len_range[0], len(p.forward) + len(p.reverse)), len_range[1])
                if contig_lens[c] < length + 4000:
                    continue
                start = rng.randrange(2000, contig_lens[c] - length - 2000)
                slot = (c, start // 4000)
                if slot in used or (c, start // 4000 + 1) in used:
                    continue
                used[slot] = used[(c, start // 4000 + 1)] = True
                break
            else:
                continue
            rs = rng.random() < 0.5
            a, b = (p.reverse, p.forward) if rs else (p.forward, p.reverse)
            left = _mutate(rng, _concrete(rng, a), rng.randint(0, max_mm), 0, max(0, len(a) - tw))
            right = revcomp(_mutate(rng, _concrete(rng, b), rng.randint(0, max_mm), 0, max(0, len(b) - tw)))
            out.append(Planted(i, c, start, length, rs, left.encode(), right.encode()))
    return out


def apply_planted(buf, contig_offsets: Sequence[int], planted: Sequence[Planted]) -> None:
    """Write planted amplicons into `buf`: a bytearray/numpy array, or a torch uint8 tensor
    (host or device) holding the concatenated contigs."""
    is_torch = hasattr(buf, "data_ptr")
    if is_torch:
        import torch
    for pl in planted:
        base = contig_offsets[pl.contig] + pl.start
        for off, data in ((0, pl.left), (pl.length - len(pl.right), pl.right)):
            if is_torch:
                t = torch.frombuffer(bytearray(data), dtype=torch.uint8)
                buf[base + off:base + off + len(data)].copy_(t)
            else:
                buf[base + off:base + off + len(data)] = data
