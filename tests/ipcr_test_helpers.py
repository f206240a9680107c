"""Shared test helpers: deterministic fixtures and input generators."""
from __future__ import annotations

import json
import os
import random

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")

IUPAC = "ACGTRYSWKMBDHVN"
COMP = dict(zip("ACGTRYSWKMBVDHN", "TGCAYRSWMKVBHDN"))


def load_reference_cases():
    with open(os.path.join(GOLDEN, "reference_cases.json")) as f:
        return json.load(f)


def revcomp(s: str) -> str:
    return "".join(COMP[c] for c in reversed(s))


# ---- port of the reference's deterministic benchmark fixture -----------------
# core/engine/performance_benchmark_test.go:20-106
def bench_dna(n: int, seed: int = 0x5EED1234) -> bytearray:
    out = bytearray(n)
    x = seed & 0xFFFFFFFF
    alpha = b"ACGT"
    for i in range(n):
        x = (x * 1664525 + 1013904223) & 0xFFFFFFFF
        out[i] = alpha[(x >> 30) & 3]
    return out


def bench_primer(idx: int, n: int) -> str:
    buf = bytearray(n)
    x = (0x9E3779B9 ^ (idx * 0x45D9F3B)) & 0xFFFFFFFF
    alpha = b"ACGT"
    for i in range(n):
        x = (x * 1103515245 + 12345 + ((i * 97) & 0xFFFFFFFF)) & 0xFFFFFFFF
        buf[i] = alpha[(x >> 29) & 3]
    buf[0] = alpha[idx & 3]
    buf[1] = alpha[(idx + 1) & 3]
    buf[2] = alpha[(idx + 2) & 3]
    buf[n - 1] = alpha[(idx + 3) & 3]
    return buf.decode()


def _different_base(b: int) -> int:
    return {ord("A"): ord("C"), ord("C"): ord("G"), ord("G"): ord("T")}.get(b, ord("A"))


def make_engine_bench_fixture(pair_count, genome_len, mutate_forward, reference_n):
    product_len, primer_len = 180, 20
    pair_count = max(1, pair_count)
    needed = 256 + pair_count * 256 + product_len
    genome_len = max(genome_len, needed)
    seq = bench_dna(genome_len)
    pairs = []
    for i in range(pair_count):
        fwd = bench_primer(i * 2, primer_len)
        rev = bench_primer(i * 2 + 1, primer_len)
        pairs.append((fwd, rev, 128, product_len + 32))
        start = 128 + i * 256
        planted = bytearray(fwd.encode())
        if mutate_forward:
            planted[10] = _different_base(planted[10])
        if reference_n:
            planted[11] = ord("N")
        seq[start:start + len(planted)] = planted
        rc = revcomp(rev).encode()
        seq[start + product_len - len(rev):start + product_len] = rc
    return bytes(seq), pairs


# ---- random inputs -----------------------------------------------------------
def random_genome(rng: random.Random, n: int, n_frac=0.0, lower_frac=0.0, other_frac=0.0) -> bytearray:
    g = bytearray(rng.choice(b"ACGT") for _ in range(n))
    for i in range(n):
        r = rng.random()
        if r < n_frac:
            g[i] = ord("N")
        elif r < n_frac + lower_frac:
            g[i] = g[i] + 32
        elif r < n_frac + lower_frac + other_frac:
            g[i] = rng.choice(b"RYKMSWBDHVU*-x")
    return g


def random_primer(rng: random.Random, n: int, iupac_frac=0.0) -> str:
    return "".join(
        rng.choice(IUPAC[4:]) if rng.random() < iupac_frac else rng.choice("ACGT") for _ in range(n)
    )


def mutate(rng: random.Random, s: str, k: int, lo: int, hi: int) -> str:
    """Substitute k distinct positions in [lo, hi) with a different concrete base."""
    b = list(s)
    idx = rng.sample(range(lo, hi), min(k, max(0, hi - lo)))
    for i in idx:
        b[i] = rng.choice([c for c in "ACGT" if c != b[i]])
    return "".join(b)


def concrete(rng: random.Random, primer: str) -> str:
    """A concrete ACGT instance of an IUPAC primer."""
    masks = {"A": "A", "C": "C", "G": "G", "T": "T", "R": "AG", "Y": "CT", "S": "CG", "W": "AT",
             "K": "GT", "M": "AC", "B": "CGT", "D": "AGT", "H": "ACT", "V": "ACG", "N": "ACGT"}
    return "".join(rng.choice(masks[c]) for c in primer)


def plant_amplicon(rng, genome: bytearray, pos: int, fwd: str, rev: str, length: int,
                   fwd_mm=0, rev_mm=0, tw=3, revcomp_strand=False):
    """Write fwd ... rc(rev) (or rev ... rc(fwd)) into genome at pos with mismatches
    outside the 3' terminal window."""
    a, b = (rev, fwd) if revcomp_strand else (fwd, rev)
    left = mutate(rng, concrete(rng, a), fwd_mm, 0, max(0, len(a) - tw))
    right_primer = mutate(rng, concrete(rng, b), rev_mm, 0, max(0, len(b) - tw))
    right = revcomp(right_primer)
    genome[pos:pos + len(left)] = left.encode()
    genome[pos + length - len(right):pos + length] = right.encode()


def multiset(products):
    from collections import Counter

    return Counter(products)


def shared_seed_case(rng, n_pairs=12, genome_len=6000):
    """Pairs whose primers share their 3' 12-mer (only the 5' parts differ): with a seed of <= 12
    columns every forward orientation has the same seed k-mers, so one dirty seed window yields
    more than 16 index hits (the overflow path of the dirty-window probe).  Returns (pairs, seq)
    with sites planted so that an N falls into the seed window of each."""
    core_a, core_b = "GATCCTGGCTCA", "CCTTGTTACGAC"
    pairs = []
    for _ in range(n_pairs):
        pa = "".join(rng.choice("ACGT") for _ in range(8)) + core_a
        pb = "".join(rng.choice("ACGT") for _ in range(9)) + core_b
        pairs.append((pa, pb, 0, 0))
    g = random_genome(rng, genome_len)
    for i, (pa, pb, _x, _y) in enumerate(pairs):
        start = 100 + i * 450
        site_a = bytearray(pa.encode())
        site_a[len(pa) - 8] = ord("N")            # inside the 3' 12-mer, outside the 3-base window
        rb = bytearray(revcomp(pb).encode())
        rb[7] = ord("N")                          # rc(B): protected window is its LEFT end
        g[start:start + len(site_a)] = site_a
        e = start + 200
        g[e - len(rb):e] = rb
    return pairs, bytes(g)
