"""CPU-side checks of the CUDA library: it loads, exports every symbol the
header declares, fails loudly without a GPU, and its host-side panel compiler
(unique orientations, `Have`, filter plans) agrees with the oracle.  No kernel
runs here."""
import os
import random
import re

import pytest

import ipcr_test_helpers as H
from ipcr_b200 import Config, Engine, Pair, _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    L = _capi.lib()
    header = open(os.path.join(ROOT, "include", "ipcr_gpu.h")).read()
    declared = set(re.findall(r"\b(ipcr_gpu_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_capi.SYMBOLS)
    for name in declared:
        assert hasattr(L, name), name
    assert L.ipcr_gpu_abi_version() == 2


def test_struct_sizes_match_header():
    import ctypes as C
    assert C.sizeof(_capi.ProductC) == 64
    assert C.sizeof(_capi.SiteC) == 24
    assert C.sizeof(_capi.RecordInfoC) == 16
    assert C.sizeof(_capi.SegmentC) == 40
    assert C.sizeof(_capi.PairC) == 48
    assert C.sizeof(_capi.Cfg) == 40


def test_fails_loudly_without_gpu():
    L = _capi.lib()
    if L.ipcr_gpu_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(_capi.GpuError, match="no CPU fallback"):
        Engine(Config()).simulate_batch("s", b"ACGTACGTACGT", [Pair("x", "ACG", "ACG")])
    err = _capi.errbuf()
    assert L.ipcr_gpu_init(0, err, len(err)) != 0 and b"no CPU fallback" in err.value


def test_panel_rejects_what_the_reference_panics_on():
    e = Engine(Config())
    for bad in ("ACGX", "acgt", "ACGU"):  # core/primer/rc_test.go:37-53
        with pytest.raises(ValueError, match="reverse-complement"):
            e.compile_panel([Pair("x", bad, "ACGT")])
    with pytest.raises(ValueError, match="max_mm"):
        Engine(Config(max_mm=-1)).compile_panel([Pair("x", "ACGT", "ACGT")])


def test_unique_orientations_are_deduplicated():
    # `--self` expands one pair into three (internal/common/primers.go:11-38) but only 4 strings
    a, b = "AGAGTTTGATCMTGGCTCAG", "TACGGYTACCTTGTTAYGACTT"
    pairs = [Pair("m", a, b), Pair("m+A:self", a, a), Pair("m+B:self", b, b)]
    cp = Engine(Config(max_mm=2, terminal_window=3)).compile_panel(pairs)
    assert cp.num_orientations == 4
    got = {cp.orientation(i) for i in range(4)}
    assert got == {(a, "A"), (b, "A"), (H.revcomp(a), "a"), (H.revcomp(b), "a")}


def test_have_matches_reference_seeding(oracle):
    cases = H.load_reference_cases()
    for c in cases["seed_structure"] + [k for k in cases["engine_known_answers"] if "have" in k["expect"]]:
        cfgd = c["cfg"]
        cp = Engine(Config(**{k: v for k, v in cfgd.items()})).compile_panel(
            [Pair(f"p{i}", p[0], p[1], p[2], p[3]) for i, p in enumerate(c["pairs"])])
        panel = oracle.Panel(oracle.make_cfg(**cfgd), [tuple(p) for p in c["pairs"]])
        for i in range(len(c["pairs"])):
            for w in "ABab":
                assert cp.have(i, w) == panel.have(i, w), (c["source"], i, w)


def test_have_matches_oracle_randomized(oracle):
    rng = random.Random(7)
    for _ in range(200):
        la, lb = rng.randint(1, 40), rng.randint(1, 40)
        a = H.random_primer(rng, la, rng.choice([0, 0.1, 0.5, 1.0]))
        b = H.random_primer(rng, lb, rng.choice([0, 0.1, 0.5]))
        cfgd = dict(max_mm=rng.randint(0, 4), terminal_window=rng.choice([0, 1, 3, 5]),
                    seed_len=rng.choice([-1, 0, 4, 8, 12, 16, 33]))
        cp = Engine(Config(**cfgd)).compile_panel([Pair("x", a, b)])
        panel = oracle.Panel(oracle.make_cfg(**cfgd), [(a, b)])
        for w in "ABab":
            assert cp.have(0, w) == panel.have(0, w), (a, b, cfgd, w)


def _true_sites(oracle, seq, orient, rc_type, mm, tw, raw):
    if rc_type:
        return {m[0] for m in oracle.find_matches(seq, orient, mm, 0, 0, left_tw=0 if raw else tw)}
    return {m[0] for m in oracle.find_matches(seq, orient, mm, 0, tw)}


def test_filter_plan_never_loses_a_true_site(oracle):
    """The plan is only a candidate generator; it must flag every start the
    reference accepts (same contract as the reference's seeds, approx_seed_oracle_test.go)."""
    rng = random.Random(99)
    checked = 0
    for it in range(150):
        n = rng.randint(120, 300)
        g = H.random_genome(rng, n, n_frac=rng.choice([0, 0.03]), lower_frac=rng.choice([0, 0.03]))
        mm, tw = rng.randint(0, 3), rng.choice([0, 1, 3, 4])
        la, lb = rng.randint(3, 40), rng.randint(3, 40)
        a = H.random_primer(rng, la, rng.choice([0, 0.15]))
        b = H.random_primer(rng, lb, rng.choice([0, 0.15]))
        for _ in range(3):
            H.plant_amplicon(rng, g, rng.randint(0, n - la - lb - 10), a, b, la + lb + rng.randint(0, 10),
                             rng.randint(0, mm), rng.randint(0, mm), tw, revcomp_strand=rng.random() < 0.5)
        seq = bytes(g)
        cp = Engine(Config(max_mm=mm, terminal_window=tw, hit_cap=rng.choice([0, 5]))).compile_panel(
            [Pair("x", a, b)])
        for q in range(cp.num_orientations):
            s, t = cp.orientation(q)
            for raw in (False, True):
                truth = _true_sites(oracle, seq, s, t == "a", mm, tw, raw)
                for pos in truth:
                    assert cp.plan_eval(q, seq, pos, raw), (s, t, mm, tw, raw, pos)
                    checked += 1
    assert checked > 500


def test_filter_plan_is_selective_on_random_dna(oracle):
    rng = random.Random(5)
    g = bytes(H.random_genome(rng, 20000))
    cp = Engine(Config(max_mm=2, terminal_window=3)).compile_panel(
        [Pair("16S", "AGAGTTTGATCMTGGCTCAG", "TACGGYTACCTTGTTAYGACTT")])
    for q in range(cp.num_orientations):
        hits = sum(cp.plan_eval(q, g, p) for p in range(len(g) - 30))
        assert hits <= 4, (q, hits)


def test_seed_index_never_loses_a_true_site_and_never_repeats_one(oracle):
    """The seed index (csrc/seed.hpp) is the other candidate generator: one k-mer per genome
    position against every enumerated seed variant of the panel.  Same contract as the
    reference's seeds (core/engine/approx_seed_oracle_test.go): every start the brute force
    accepts is a candidate -- and no (orientation, start) comes twice, because a duplicate
    candidate would become a duplicate site."""
    rng = random.Random(4242)
    checked = 0
    for it in range(120):
        n = rng.randint(120, 320)
        g = H.random_genome(rng, n, n_frac=rng.choice([0, 0.02, 0.1]), lower_frac=rng.choice([0, 0.03]))
        mm, tw = rng.randint(0, 3), rng.choice([0, 1, 3, 4])
        pairs = []
        for k in range(rng.randint(1, 3)):
            la, lb = rng.randint(3, 40), rng.randint(3, 40)
            a = H.random_primer(rng, la, rng.choice([0, 0.15]))
            b = H.random_primer(rng, lb, rng.choice([0, 0.15]))
            pairs.append(Pair(f"p{k}", a, b))
            for _ in range(2):
                H.plant_amplicon(rng, g, rng.randint(0, max(0, n - la - lb - 10)), a, b, la + lb + rng.randint(0, 10),
                                 rng.randint(0, mm), rng.randint(0, mm), tw, revcomp_strand=rng.random() < 0.5)
        seq = bytes(g)
        cp = Engine(Config(max_mm=mm, terminal_window=tw, hit_cap=rng.choice([0, 5]))).compile_panel(pairs)
        for raw in (False, True):
            cands, info = cp.seed_eval(seq, raw)
            assert len(cands) == len(set(cands)), "duplicate candidate"
            cset = set(cands)
            assert 1 <= info["S"] <= 16
            for q in range(cp.num_orientations):
                s, t = cp.orientation(q)
                for pos in _true_sites(oracle, seq, s, t == "a", mm, tw, raw):
                    assert (q, pos) in cset, (s, t, mm, tw, raw, pos, info)
                    checked += 1
    assert checked > 800


def test_seed_index_complete_on_a_multiplex_panel(oracle):
    """Panel scale: 40 pairs + self pairs, planted amplicons with mismatches, N and lowercase in
    the genome; every site the brute force accepts for any of the ~160 orientations is a
    candidate of the index, for both plan variants."""
    from ipcr_b200 import synth
    rng = random.Random(99)
    user, allp = synth.multiplex_panel(40, seed=3)
    for mm in (1, 2):
        lens = [600000]
        g = H.random_genome(rng, lens[0], n_frac=0.001, lower_frac=0.001)
        for pl in synth.plan_amplicons(user, lens, mm, 3, seed=5 + mm, per_pair=(1, 1), len_range=(60, 400)):
            g[pl.start:pl.start + len(pl.left)] = pl.left
            e = pl.start + pl.length
            g[e - len(pl.right):e] = pl.right
        seq = bytes(g)
        cp = Engine(Config(max_mm=mm, terminal_window=3, max_len=2000, hit_cap=10000)).compile_panel(allp)
        n_true = 0
        for raw in (False, True):
            cands, info = cp.seed_eval(seq, raw)
            assert len(cands) == len(set(cands))
            cset = set(cands)
            for q in range(cp.num_orientations):
                s, t = cp.orientation(q)
                for pos in _true_sites(oracle, seq, s, t == "a", mm, 3, raw):
                    assert (q, pos) in cset, (s, t, mm, raw, pos, info)
                    n_true += 1
        assert n_true >= 2 * 2 * 36  # both primer sites of (nearly) every planted amplicon, both variants


def test_seed_index_high_mismatch_budgets(oracle):
    """max_mm 4..5 on N-rich sequences: dirty windows with more than four invalid bytes leave the
    replacement lookups for the verify-everything path; completeness must survive."""
    rng = random.Random(515)
    checked = 0
    for it in range(25):
        n = rng.randint(150, 260)
        g = H.random_genome(rng, n, n_frac=rng.choice([0.05, 0.15, 0.3]))
        mm, tw = rng.choice([4, 5]), rng.choice([0, 2, 3])
        la, lb = rng.randint(8, 30), rng.randint(8, 30)
        a, b = H.random_primer(rng, la, rng.choice([0, 0.15])), H.random_primer(rng, lb, 0)
        for _ in range(3):
            H.plant_amplicon(rng, g, rng.randint(0, n - la - lb - 10), a, b, la + lb + rng.randint(0, 10),
                             rng.randint(0, mm), rng.randint(0, mm), tw, revcomp_strand=rng.random() < 0.5)
        seq = bytes(g)
        cp = Engine(Config(max_mm=mm, terminal_window=tw)).compile_panel([Pair("x", a, b)])
        for raw in (False, True):
            cands, _info = cp.seed_eval(seq, raw)
            assert len(cands) == len(set(cands))
            cset = set(cands)
            for q in range(cp.num_orientations):
                s, t = cp.orientation(q)
                for pos in _true_sites(oracle, seq, s, t == "a", mm, tw, raw):
                    assert (q, pos) in cset, (s, t, mm, tw, raw, pos)
                    checked += 1
    assert checked > 100


def test_seed_index_dirty_window_overflow(oracle, monkeypatch):
    """More than 16 orientations behind one dirty seed window: the probe gives up on the window
    and every orientation is verified there -- still complete, still duplicate-free."""
    monkeypatch.setenv("IPCR_SEED_LEN", "10")
    rng = random.Random(12)
    raw_pairs, seq = H.shared_seed_case(rng)
    pairs = [Pair(f"p{i}", p[0], p[1]) for i, p in enumerate(raw_pairs)]
    cp = Engine(Config(max_mm=2, terminal_window=3, max_len=400)).compile_panel(pairs)
    n_true = 0
    for raw in (False, True):
        cands, info = cp.seed_eval(seq, raw)
        assert info["S"] == 10
        assert len(cands) == len(set(cands))
        cset = set(cands)
        for q in range(cp.num_orientations):
            s, t = cp.orientation(q)
            for pos in _true_sites(oracle, seq, s, t == "a", 2, 3, raw):
                assert (q, pos) in cset
                n_true += 1
    assert n_true >= 2 * 2 * 12
    # the shared window really overflows: all 48 orientations are candidates at one k-mer start
    per_start = {}
    for q, pos in cp.seed_eval(seq)[0]:
        per_start.setdefault(pos, set()).add(q)
    assert max(len(v) for v in per_start.values()) >= 12


def test_seed_index_shape_on_the_benchmark_panel():
    from ipcr_b200 import synth
    _, allp = synth.multiplex_panel(100, seed=7)
    cp = Engine(Config(max_mm=2, terminal_window=3, max_len=2000, hit_cap=10000)).compile_panel(allp)
    rng = random.Random(3)
    g = bytes(H.random_genome(rng, 50000))
    cands, info = cp.seed_eval(g)
    assert info["S"] >= 12 and info["bloom"] and info["bloom_fill"] < 0.35 and info["bloom_k"] == 2
    assert info["stride"] == 1  # 220 k first-level keys: a strided first level would double them (measured: 4.0 vs 2.3 ms)
    # mm 1: 20 times fewer keys.  The first level looks at every 4th position only (its keys at 4 window offsets)
    # and takes the second bitmap bit, because a false positive there stalls a whole warp
    info1 = Engine(Config(max_mm=1, terminal_window=3)).compile_panel(allp).seed_eval(g)[1]
    assert info1["stride"] == 4 and info1["bloom_k"] == 2 and info1["S"] - info1["S1"] >= 3 and info1["bloom_fill"] < 0.1
    # one pair, no mismatches: a handful of keys, one bit is enough
    _, one = synth.multiplex_panel(1, seed=7)
    info0 = Engine(Config(max_mm=0, terminal_window=3)).compile_panel(one).seed_eval(g)[1]
    assert info0["stride"] == 4 and info0["bloom_k"] == 1
    assert info["entries"] < 1 << 20
    assert len(cands) < 400  # ~ entries / 4^S per position


def test_primer_length_limit_is_an_error_not_a_truncation():
    """The GPU engine carries mismatch positions as 64-bit masks: primers up to 64 nt compile, longer ones
    are refused at panel creation with a message naming the limit (the reference stores up to 65 535 nt,
    core/engine/compiled.go:43-48; documented difference)."""
    eng = Engine(Config(max_mm=1, terminal_window=3))
    ok = eng.compile_panel([Pair("edge", "ACGT" * 16, "TTGACCGTAAGC")])
    assert ok.num_orientations == 4
    with pytest.raises(Exception) as ei:
        eng.compile_panel([Pair("long", "ACGT" * 16 + "A", "TTGACCGTAAGC")])
    assert "longer than 64" in str(ei.value)
    with pytest.raises(Exception) as ei:
        eng.compile_panel([Pair("long", "TTGACCGTAAGC", "ACGT" * 17)])
    assert "longer than 64" in str(ei.value)


def test_pigeonhole_first_level_never_loses_a_true_site(oracle, monkeypatch):
    """The pigeonhole first level of the seed index (two exact 8-mer tables for the low / high half of a 16-base
    window, each within max_mm / 2 mismatches; chosen by the cost model when a panel's keys would saturate the
    shared-memory bitmap, e.g. mm 3 x 400 orientations) is forced on small panels here: every start the brute
    force accepts is still a candidate, none comes twice."""
    monkeypatch.setenv("IPCR_SEED_PIGEON", "1")
    rng = random.Random(777)
    checked = used = 0
    for it in range(80):
        n = rng.randint(150, 360)
        g = H.random_genome(rng, n, n_frac=rng.choice([0, 0.02]), lower_frac=rng.choice([0, 0.03]))
        mm, tw = rng.randint(0, 3), rng.choice([0, 1, 3, 4])
        pairs = []
        for k in range(rng.randint(1, 3)):
            la, lb = rng.randint(16, 34), rng.randint(16, 34)
            a = H.random_primer(rng, la, rng.choice([0, 0.15]))
            b = H.random_primer(rng, lb, rng.choice([0, 0.15]))
            pairs.append(Pair(f"p{k}", a, b))
            for _ in range(3):
                H.plant_amplicon(rng, g, rng.randint(0, max(0, n - la - lb - 10)), a, b, la + lb + rng.randint(0, 10),
                                 rng.randint(0, mm), rng.randint(0, mm), tw, revcomp_strand=rng.random() < 0.5)
        seq = bytes(g)
        cp = Engine(Config(max_mm=mm, terminal_window=tw, hit_cap=rng.choice([0, 5]))).compile_panel(pairs)
        for raw in (False, True):
            cands, info = cp.seed_eval(seq, raw)
            if info["S"] != 16:
                continue
            assert info["bloom_k"] == 3  # the debug tap reports the pigeonhole tables as k = 3
            used += 1
            assert len(cands) == len(set(cands)), "duplicate candidate"
            cset = set(cands)
            for q in range(cp.num_orientations):
                s, t = cp.orientation(q)
                for pos in _true_sites(oracle, seq, s, t == "a", mm, tw, raw):
                    assert (q, pos) in cset, (s, t, mm, tw, raw, pos, info)
                    checked += 1
    assert used > 60 and checked > 600


@pytest.mark.parametrize("stride", [2, 4])
def test_strided_first_level_never_loses_a_true_site(oracle, monkeypatch, stride):
    """Anchor stride of the seed index's first level (small panels: one shared-memory lookup per 2 or 4 positions;
    the bitmap then holds each seed's first-level keys at `stride` window offsets): with the stride forced, every
    start the brute force accepts is still a candidate, none comes twice, and the candidate set is the one the
    unstrided index gives (the first level is only a filter: the second level decides)."""
    rng = random.Random(4242 + stride)
    checked = used = 0
    for it in range(60):
        n = rng.randint(150, 360)
        g = H.random_genome(rng, n, n_frac=rng.choice([0, 0.02]), lower_frac=rng.choice([0, 0.03]))
        mm, tw = rng.randint(0, 3), rng.choice([0, 1, 3, 4])
        pairs = []
        for k in range(rng.randint(1, 3)):
            la, lb = rng.randint(12, 34), rng.randint(12, 34)
            a = H.random_primer(rng, la, rng.choice([0, 0.15]))
            b = H.random_primer(rng, lb, rng.choice([0, 0.15]))
            pairs.append(Pair(f"p{k}", a, b))
            for _ in range(3):
                H.plant_amplicon(rng, g, rng.randint(0, max(0, n - la - lb - 10)), a, b, la + lb + rng.randint(0, 10),
                                 rng.randint(0, mm), rng.randint(0, mm), tw, revcomp_strand=rng.random() < 0.5)
        seq = bytes(g)
        cfg = Config(max_mm=mm, terminal_window=tw, hit_cap=rng.choice([0, 5]))
        plain = Engine(cfg).compile_panel(pairs)
        cp = Engine(cfg).compile_panel(pairs)
        for raw in (False, True):
            monkeypatch.setenv("IPCR_SEED_STRIDE_FORCE", "1")  # the index is built at its first use
            want, info1 = plain.seed_eval(seq, raw)
            monkeypatch.setenv("IPCR_SEED_STRIDE_FORCE", str(stride))
            cands, info = cp.seed_eval(seq, raw)
            assert info1["stride"] == 1
            if not info["bloom"] or info["stride"] != stride:
                continue  # a seed too short to leave `stride - 1` bases beside a 3-base key
            used += 1
            assert info["S"] - info["S1"] >= stride - 1
            assert len(cands) == len(set(cands)), "duplicate candidate"
            assert sorted(cands) == sorted(want)
            cset = set(cands)
            for q in range(cp.num_orientations):
                s, t = cp.orientation(q)
                for pos in _true_sites(oracle, seq, s, t == "a", mm, tw, raw):
                    assert (q, pos) in cset, (s, t, mm, tw, raw, pos, info)
                    checked += 1
    assert used > 60 and checked > 400
