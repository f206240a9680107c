"""Parity at BASELINE.json's full sizes through size-independent properties (the oracle
cannot brute-force 3 Gbp in seconds):
  * every planted amplicon comes back with its exact coordinates, type and mismatch counts;
  * both filter kernels (table-driven / NVRTC-specialised) and the sharded path agree;
  * the scan is idempotent;
  * a sampled window of the genome, cut out and given to the oracle, yields exactly the
    products the full scan reports inside that window.
"""
import random

import numpy as np
import pytest

import ipcr_test_helpers as H
import ref_host as R
from ipcr_b200 import Config, Engine, Pair, sharding, synth
from ipcr_b200 import device as D
from ipcr_b200 import hostlib as HL

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

SEED = 991


def _genome(sc, total, seed):
    buf = torch.empty(total, dtype=torch.uint8, device="cuda")
    D.synth_fill(sc, buf, 0, seed)
    return buf


def _expected(oracle, pairs, lens, planted):
    """(rec, pair id, type, start, end, fwd_mm, rev_mm) of every planted amplicon."""
    out = set()
    for pl in planted:
        p = pairs[pl.pair_idx]
        a, b = (p.reverse, p.forward) if pl.revcomp_strand else (p.forward, p.reverse)
        fwd_mm = oracle.mismatch_count(pl.left, a.encode())
        rev_mm = oracle.mismatch_count(synth.revcomp(pl.right.decode()).encode(), b.encode())
        out.add((pl.contig, p.id, "revcomp" if pl.revcomp_strand else "forward", pl.start, pl.start + pl.length,
                 fwd_mm, rev_mm))
    return out


def _key(p):
    return (int(p.sequence_id), p.experiment_id, p.type, p.start, p.end, p.fwd_mm, p.rev_mm)


def _scan(cp, sc, buf, lens, jit):
    cp.set_jit(jit)
    offs = sharding.contig_offsets(lens)
    segs = [(buf[offs[c]:offs[c + 1]], c) for c in range(len(lens))]
    return D.scan_device(cp, sc, segs, n_recs=len(lens))


def _check_window(oracle, cfgd, all_pairs, buf, lens, prods, rng, n_windows=3, width=400_000):
    """Cut windows out of the device genome and replay them on the CPU oracle."""
    offs = sharding.contig_offsets(lens)
    raw = [(p.forward, p.reverse, p.min_product, p.max_product) for p in all_pairs]
    ocfg = oracle.make_cfg(**{k: v for k, v in cfgd.items()})
    for _ in range(n_windows):
        c = rng.randrange(len(lens))
        w0 = rng.randrange(0, max(1, lens[c] - width))
        w1 = min(lens[c], w0 + width)
        piece = bytes(buf[offs[c] + w0:offs[c] + w1].cpu().numpy().tobytes())
        want = {(all_pairs[t[0]].id, t[1], t[2] + w0, t[3] + w0, t[5], t[6], t[7], t[8])
                for t in oracle.simulate_compiled(ocfg, raw, piece)}
        got = {(p.experiment_id, p.type, p.start, p.end, p.fwd_mm, p.rev_mm, tuple(p.fwd_mismatch_idx),
                tuple(p.rev_mismatch_idx)) for p in prods
               if int(p.sequence_id) == c and p.start >= w0 and p.end <= w1 and p.start < p.end}
        # products fully inside the window must match (sites near the window edge may pair outside)
        inner = {g for g in got if g[2] >= w0 + 64 and g[3] <= w1 - 64}
        assert inner <= want
        assert {w for w in want if w[2] >= w0 + 64 and w[3] <= w1 - 64} <= got


def test_config3_100mbp_multicontig(oracle):
    """BASELINE config 3 shape: 100 Mbp = 50 x 2 Mbp contigs, 27F/1492R + self, mm 2."""
    rng = random.Random(SEED)
    lens = [2_000_000] * 50
    user, allp = synth.pair_16s()
    cfgd = dict(max_mm=2, terminal_window=3, max_len=2000, hit_cap=10000, seed_len=12)
    eng = Engine(Config(**cfgd))
    cp = eng.compile_panel(allp)
    sc = eng.new_simulation_scratch(cp)
    buf = _genome(sc, sum(lens), SEED)
    planted = synth.plan_amplicons(user * 40, lens, 2, 3, seed=SEED, per_pair=(1, 1), len_range=(200, 1900))
    for pl in planted:
        pl.pair_idx = 0
    synth.apply_planted(buf, sharding.contig_offsets(lens), planted)
    torch.cuda.synchronize()
    exp = _expected(oracle, user, lens, planted)
    results = {}
    for jit in (0, 2, 3):
        prods = _scan(cp, sc, buf, lens, jit)
        assert sc.timing()["used_jit"] == {0: 0, 2: 1, 3: 2}[jit]
        results[jit] = [(_key(p), p.fwd_mismatch_idx, p.rev_mismatch_idx) for p in prods]
        assert exp <= {_key(p) for p in prods}
    assert results[0] == results[2] == results[3]  # same products, same order, all three filters
    again = _scan(cp, sc, buf, lens, 2)
    assert [(_key(p), p.fwd_mismatch_idx, p.rev_mismatch_idx) for p in again] == results[2]  # idempotent
    _check_window(oracle, cfgd, allp, buf, lens, again, rng)


def test_config4_500mbp_outer_pairs(oracle):
    """BASELINE config 4 shape: 500 Mbp = 10 x 50 Mbp, outer 27F/1492R, mm 1 (input >> L2)."""
    rng = random.Random(SEED + 1)
    lens = [50_000_000] * 10
    user, allp = synth.pair_16s()
    cfgd = dict(max_mm=1, terminal_window=3, max_len=2000, hit_cap=10000, seed_len=12)
    eng = Engine(Config(**cfgd))
    cp = eng.compile_panel(allp)
    sc = eng.new_simulation_scratch(cp)
    buf = _genome(sc, sum(lens), SEED + 1)
    planted = synth.plan_amplicons(user * 30, lens, 1, 3, seed=SEED + 1, per_pair=(1, 1), len_range=(1200, 1700))
    for pl in planted:
        pl.pair_idx = 0
    synth.apply_planted(buf, sharding.contig_offsets(lens), planted)
    torch.cuda.synchronize()
    exp = _expected(oracle, user, lens, planted)
    prods = _scan(cp, sc, buf, lens, 2)
    assert sc.timing()["used_jit"] == 1
    assert exp <= {_key(p) for p in prods}
    _check_window(oracle, cfgd, allp, buf, lens, prods, rng, n_windows=2)
    seeded = _scan(cp, sc, buf, lens, 3)  # the seed index reproduces the bit-parallel scan exactly
    assert sc.timing()["used_jit"] == 2
    assert [(_key(p), p.fwd_mismatch_idx, p.rev_mismatch_idx) for p in seeded] == \
           [(_key(p), p.fwd_mismatch_idx, p.rev_mismatch_idx) for p in prods]
    cp.set_jit(2)
    # shard-and-gather (4 shards) reproduces the single scan exactly
    offs = sharding.contig_offsets(lens)
    sites, recs = [], [(0, False)] * len(lens)
    for r in range(4):
        segs = [(buf[s0:s1], c, ro, own) for (s0, s1, c, ro, own) in sharding.shard_segments(lens, 4, r)]
        s_r, rec_r = D.find_sites_device(cp, sc, segs, len(lens), force_raw=False)
        sites.append(s_r)
        recs = [(max(a[0], b[0]), a[1] or b[1]) for a, b in zip(recs, rec_r)]
    sharded = D.pair_sites(cp, sc, np.concatenate(sites), [(lens[i], recs[i][1]) for i in range(len(lens))])
    assert [(_key(p), p.fwd_mismatch_idx, p.rev_mismatch_idx) for p in sharded] == \
           [(_key(p), p.fwd_mismatch_idx, p.rev_mismatch_idx) for p in prods]


def test_config5_3gbp_multiplex_mm3(oracle):
    """BASELINE config 5: 3 Gbp (24 contigs of 50-250 Mbp), 100-pair panel + self, mm 3."""
    rng = random.Random(SEED + 2)
    total = 3_000_000_000
    lens = synth.contig_lengths(total, 24)
    assert sum(lens) == total and min(lens) >= 45_000_000 and max(lens) <= 260_000_000
    user, allp = synth.multiplex_panel(100, seed=SEED)
    cfgd = dict(max_mm=3, terminal_window=3, max_len=2000, hit_cap=10000, seed_len=12)
    eng = Engine(Config(**cfgd))
    cp = eng.compile_panel(allp)
    sc = eng.new_simulation_scratch(cp)
    buf = _genome(sc, total, SEED + 2)
    planted = synth.plan_amplicons(user, lens, 3, 3, seed=SEED + 2)
    synth.apply_planted(buf, sharding.contig_offsets(lens), planted)
    torch.cuda.synchronize()
    exp = _expected(oracle, user, lens, planted)
    prods = _scan(cp, sc, buf, lens, 2)
    t = sc.timing()
    assert t["used_jit"] == 1
    got = {_key(p) for p in prods}
    assert exp <= got, sorted(exp - got)[:5]
    # positions exceed int32 inside the concatenated buffer and every product respects its bounds
    assert all(0 <= p.start < p.end <= lens[int(p.sequence_id)] and p.length == p.end - p.start <= 2000 for p in prods)
    assert all(p.fwd_mm <= 3 and p.rev_mm <= 3 and len(p.fwd_mismatch_idx) == p.fwd_mm for p in prods)
    _check_window(oracle, cfgd, allp, buf, lens, prods, rng, n_windows=2, width=150_000)
    # mm 3 x 400 orientations saturates the shared-memory bitmap: the seed index then probes its
    # hash table for every position -- slower, still exact
    seeded = _scan(cp, sc, buf, lens, 3)
    assert sc.timing()["used_jit"] == 2
    assert [(_key(p), p.fwd_mismatch_idx, p.rev_mismatch_idx) for p in seeded] == \
           [(_key(p), p.fwd_mismatch_idx, p.rev_mismatch_idx) for p in prods]


# ---- full equality at configs 3 and 4: whole-genome sorted TSV text, GPU host mirror vs the oracle-driven
# restatement of the reference host (the oracle walks 100-500 Mbp in seconds for a 4-orientation panel) ----
def _np_genome(seed, n, lower_blocks=3, n_runs=2):
    """uniform ACGT with a few soft-masked (lowercase) blocks and N runs, as a bytearray"""
    r = np.random.default_rng(seed)
    g = np.frombuffer(b"ACGT", dtype=np.uint8)[r.integers(0, 4, n)].copy()
    for _ in range(lower_blocks):
        a = int(r.integers(0, max(1, n - 5000)))
        g[a:a + 3000] |= 0x20
    for _ in range(n_runs):
        a = int(r.integers(0, max(1, n - 500)))
        g[a:a + int(r.integers(1, 200))] = ord("N")
    return bytearray(g.tobytes())


def _write_fasta(path, recs, width=80):
    with open(path, "wb") as f:
        for rid, seq in recs:
            f.write(b">" + rid.encode() + b" synthetic\n")
            a = np.frombuffer(seq, dtype=np.uint8)
            full = len(a) // width * width
            if full:
                body = np.empty((full // width, width + 1), dtype=np.uint8)
                body[:, :width] = a[:full].reshape(-1, width)
                body[:, width] = 10
                f.write(body.tobytes())
            if full < len(a):
                f.write(a[full:].tobytes() + b"\n")
    return str(path)


def test_config3_full_text_equality_with_probe(tmp_path, oracle):
    """BASELINE config 3 end to end: 100 Mbp in 50 contigs, 27F/1492R + TaqMan-style probe on both
    strands, --mismatches 2: the sorted ipcr-probe TSV of the GPU path (native FASTA feeder, device
    scan, ipcr_gpu_probe_annotate) equals the reference host restated around the CPU oracle, row for row."""
    rng = random.Random(SEED + 30)
    a, b = "AGAGTTTGATCMTGGCTCAG", "TACGGYTACCTTGTTAYGACTT"
    probe = "ACTGGCTCACCTTGCGAGTCATGC"
    recs = []
    for c in range(50):
        n = 2_000_000
        g = _np_genome(SEED + 100 + c, n)
        for i in range(2):
            pos = 50_000 + i * 900_000 + rng.randint(0, 1000)
            H.plant_amplicon(rng, g, pos, a, b, rng.randint(300, 1800), rng.randint(0, 2), rng.randint(0, 2), 3,
                             revcomp_strand=(c + i) % 2 == 1)
            pc = H.mutate(rng, H.concrete(rng, probe), rng.randint(0, 1), 0, len(probe))
            pc = pc if (c + i) % 3 else H.revcomp(pc)
            g[pos + 60:pos + 60 + len(pc)] = pc.encode()
        recs.append((f"contig{c}", bytes(g)))
    fa = _write_fasta(tmp_path / "c3.fa", recs)
    del recs
    for pmm, req in ((0, True), (1, False)):
        od = dict(mismatches=2, sort=True)
        got = HL.run_probe(HL.Options(**od), [fa], a, b, probe, "taq", pmm, req)
        want = R.run(oracle, "probe", [fa], od, fwd=a, rev=b, probe=probe, probe_name="taq", probe_max_mm=pmm,
                     require_probe=req)
        assert got == want
        assert len(got.splitlines()) > 30 and "true" in got


def test_config4_full_text_equality_nested(tmp_path, oracle):
    """BASELINE config 4 end to end: 500 Mbp in 10 contigs, outer 27F/1492R + an inner pair planted
    inside the outer amplicons, --mismatches 1: sorted ipcr-nested TSV, GPU path (ipcr_gpu_nested:
    every amplicon scanned with the inner panel on the device) against the oracle-driven reference host."""
    rng = random.Random(SEED + 40)
    a, b = "AGAGTTTGATCMTGGCTCAG", "TACGGYTACCTTGTTAYGACTT"
    ia, ib = "GCAACGAGCGCAACCC", "ACGGGCGGTGTGTRC"
    recs = []
    for c in range(10):
        n = 50_000_000
        g = _np_genome(SEED + 200 + c, n, lower_blocks=5, n_runs=0)
        for i in range(4):
            pos = 1_000_000 + i * 11_000_000 + rng.randint(0, 5000)
            H.plant_amplicon(rng, g, pos, a, b, rng.randint(1200, 1700), rng.randint(0, 1), rng.randint(0, 1), 3,
                             revcomp_strand=(c + i) % 2 == 1)
            H.plant_amplicon(rng, g, pos + 100, ia, ib, rng.randint(60, 150), 0, rng.randint(0, 1), 3,
                             revcomp_strand=i % 4 == 1)
        recs.append((f"chr{c}", bytes(g)))
    fa = _write_fasta(tmp_path / "c4.fa", recs)
    del recs
    od = dict(mismatches=1, sort=True)
    got = HL.run_nested(HL.Options(**od), [fa], a, b, ia, ib, require_inner=False)
    want = R.run(oracle, "nested", [fa], od, fwd=a, rev=b, inner_fwd=ia, inner_rev=ib, require_inner=False)
    assert got == want
    assert len(got.splitlines()) > 40 and "\ttrue\t" in got


def test_config5_3gbp_multiplex_mm2_headline(oracle):
    """The bench workload itself (bench.py: same seeds, same planting, --mismatches 2): the default
    filter's product list equals the bit-parallel filter's, in order, holds every planted amplicon, and
    8 windows replayed on the CPU oracle give exactly the products reported inside them."""
    BSEED = 20251018
    rng = random.Random(SEED + 5)
    total = 3_000_000_000
    lens = synth.contig_lengths(total, 24)
    user, allp = synth.multiplex_panel(100, seed=BSEED)
    cfgd = dict(max_mm=2, terminal_window=3, max_len=2000, hit_cap=10000, seed_len=12)
    eng = Engine(Config(**cfgd))
    cp = eng.compile_panel(allp)
    sc = eng.new_simulation_scratch(cp)
    buf = _genome(sc, total, BSEED)
    planted = synth.plan_amplicons(user, lens, 2, 3, seed=BSEED + 1)
    synth.apply_planted(buf, sharding.contig_offsets(lens), planted)
    torch.cuda.synchronize()
    exp = _expected(oracle, user, lens, planted)
    cp.set_jit(1)  # automatic = what the bench runs
    prods = _scan(cp, sc, buf, lens, 1)
    assert sc.timing()["used_jit"] == 2
    got = {_key(p) for p in prods}
    assert exp <= got, sorted(exp - got)[:5]
    rows = [(_key(p), p.fwd_mismatch_idx, p.rev_mismatch_idx) for p in prods]
    bitpar = _scan(cp, sc, buf, lens, 2)
    assert sc.timing()["used_jit"] == 1
    assert [(_key(p), p.fwd_mismatch_idx, p.rev_mismatch_idx) for p in bitpar] == rows
    _check_window(oracle, cfgd, allp, buf, lens, prods, rng, n_windows=8, width=200_000)
    # windows AROUND planted amplicons, so that the replay actually sees products
    offs = sharding.contig_offsets(lens)
    raw = [(p.forward, p.reverse, p.min_product, p.max_product) for p in allp]
    ocfg = oracle.make_cfg(**cfgd)
    for pl in rng.sample(planted, 8):
        w0 = max(0, pl.start - 50_000)
        w1 = min(lens[pl.contig], pl.start + pl.length + 50_000)
        piece = bytes(buf[offs[pl.contig] + w0:offs[pl.contig] + w1].cpu().numpy().tobytes())
        want = {(allp[t[0]].id, t[1], t[2] + w0, t[3] + w0, t[5], t[6], t[7], t[8])
                for t in oracle.simulate_compiled(ocfg, raw, piece)}
        inside = {(p.experiment_id, p.type, p.start, p.end, p.fwd_mm, p.rev_mm, tuple(p.fwd_mismatch_idx),
                   tuple(p.rev_mismatch_idx)) for p in prods
                  if int(p.sequence_id) == pl.contig and p.start >= w0 + 64 and p.end <= w1 - 64 and p.start < p.end}
        assert inside and inside == {w for w in want if w[2] >= w0 + 64 and w[3] <= w1 - 64}
