"""Pins the CPU oracle (oracle/) against every known-answer test the reference
holds for the scan path (tests/golden/reference_cases.json cites file:line).
CPU only."""
import random

import pytest

import ipcr_test_helpers as H

CASES = H.load_reference_cases()


def _cfg(O, d):
    return O.make_cfg(**d)


def test_base_match_table(oracle):
    for g, p, want in CASES["base_match"]["cases"]:
        assert oracle.base_match(g, p) is want


def test_iupac_mask_snapshot(oracle):
    for c, m in CASES["iupac_mask"]["cases"].items():
        assert oracle.iupac_mask(c) == m, c


def test_base_match_genome_must_be_uppercase_acgt(oracle):
    # core/primer/iupac.go:62-67: genome side is strict
    for g in "acgtNnRUu*-":
        for p in "ACGTN":
            assert not oracle.base_match(g, p)


def test_mismatch_count(oracle):
    for g, p, want in CASES["mismatch_count"]["cases"]:
        assert oracle.mismatch_count(g, p) == want
    with pytest.raises(ValueError):
        oracle.mismatch_count("AAA", "AA")


def test_revcomp(oracle):
    for src, want in CASES["revcomp"]["cases"]:
        assert oracle.revcomp(src) == want.encode()
    for bad in CASES["revcomp"]["invalid"]:
        with pytest.raises(ValueError):
            oracle.revcomp(bad)


def test_find_matches_known_answers(oracle):
    fm = CASES["find_matches"]
    for c in fm["cases"]:
        hits = oracle.find_matches(fm["seq"], c["primer"], c["max_mm"], 0, c["tw"])
        assert len(hits) == c["count"], c["name"]
        if c["count"] > 0 and c["first"] != -1:
            assert hits[0][0] == c["first"], c["name"]


@pytest.mark.parametrize("case", CASES["engine_known_answers"], ids=lambda c: c["name"])
@pytest.mark.parametrize("path", ["bruteforce", "compiled"])
def test_engine_known_answers(oracle, case, path):
    cfg = _cfg(oracle, case["cfg"])
    pairs = [tuple(p) for p in case["pairs"]]
    if path == "bruteforce":
        got = oracle.simulate_bruteforce(cfg, pairs, case["seq"])
    else:
        got = oracle.simulate_compiled(cfg, pairs, case["seq"])
    e = case["expect"]
    if "count" in e:
        assert len(got) == e["count"]
    if "min_count" in e:
        assert len(got) >= e["min_count"]
    if "first" in e:
        f = got[0]
        d = dict(start=f[2], end=f[3], length=f[4], fwd_mm=f[5], rev_mm=f[6], fwd_idx=list(f[7]),
                 rev_idx=list(f[8]))
        for k, v in e["first"].items():
            assert d[k] == v, (k, d)
    if "length_range" in e:
        lo, hi = e["length_range"]
        assert all(lo <= p[4] <= hi for p in got)
    if "has_type" in e:
        assert any(p[1] == e["has_type"] for p in got)
    if e.get("wrap"):
        p = got[0]
        assert p[2] > p[3] and p[4] == len(case["seq"]) - p[2] + p[3]
    if "have" in e and path == "compiled":
        panel = oracle.Panel(cfg, pairs)
        for pi, which, want in e["have"]:
            assert panel.have(pi, which) is want


@pytest.mark.parametrize("case", CASES["seed_structure"], ids=lambda c: c["source"])
def test_seed_structure(oracle, case):
    cfg = _cfg(oracle, case["cfg"])
    pairs = [tuple(p) for p in case["pairs"]]
    panel = oracle.Panel(cfg, pairs)
    if "num_patterns" in case:
        assert panel.num_patterns == case["num_patterns"]
    if "all_have" in case:
        for i in range(len(pairs)):
            for w in "ABab":
                assert panel.have(i, w) is case["all_have"]
    for pi, which, want in case.get("have", []):
        assert panel.have(pi, which) is want, (pi, which)


def _equiv(oracle, cfgd, pairs, seq):
    cfg = _cfg(oracle, cfgd)
    fast = oracle.simulate_compiled(cfg, pairs, seq)
    brute = oracle.simulate_bruteforce(cfg, pairs, seq)
    assert H.multiset(fast) == H.multiset(brute)
    return fast


def test_seeded_equals_bruteforce_on_reference_cases(oracle):
    eq = CASES["equivalence"]
    n = 0
    for c in eq["cases"]:
        for cfgd in eq["configs"]:
            _equiv(oracle, cfgd, [tuple(p) for p in c["pairs"]], c["seq"])
            n += 1
    assert n == 81
    for c in CASES["equivalence_extra"]:
        seqs = c.get("seqs") or [c["seq"]]
        for s in seqs:
            for cfgd in c["configs"]:
                _equiv(oracle, cfgd, [tuple(p) for p in c["pairs"]], s)


def test_survey_derived_expectations(oracle):
    for c in CASES["survey_derived"]["cases"]:
        got = oracle.simulate_bruteforce(_cfg(oracle, c["cfg"]), [tuple(p) for p in c["pairs"]], c["seq"])
        want = [tuple([p[0], p[1], p[2], p[3], p[4], p[5], p[6], tuple(p[7]), tuple(p[8])]) for p in c["products"]]
        assert got == want


@pytest.mark.parametrize("case", CASES["bench_fixture"]["cases"], ids=lambda c: str(c["pair_count"]))
def test_bench_fixture_planted_products(oracle, case):
    seq, pairs = H.make_engine_bench_fixture(case["pair_count"], case["genome_len"],
                                             case["mutate_forward"], case["reference_n"])
    cfg = _cfg(oracle, case["cfg"])
    got = _equiv(oracle, case["cfg"], pairs, seq)
    panel = oracle.Panel(cfg, pairs)
    if case["all_seeded"]:
        assert all(panel.have(i, w) for i in range(len(pairs)) for w in "ABab")
    # every planted pair yields its 180-bp forward product at 128 + 256*i
    for i in range(len(pairs)):
        assert any(p[0] == i and p[1] == "forward" and p[2] == 128 + 256 * i and p[4] == 180 for p in got), i


def test_hit_cap_keeps_first_matches(oracle):
    # core/engine/hit_collect_test.go:29-42 (collector) / match.go:46-48,84-86
    hits = oracle.find_matches("AAAAAA", "AA", 0, 1, 0)
    assert hits == [(0, 0, ())]
    hits = oracle.find_matches("AAAAAA", "AA", 0, 3, 0)
    assert [h[0] for h in hits] == [0, 1, 2]


def test_seeded_equals_bruteforce_randomized(oracle):
    rng = random.Random(1234)
    for it in range(300):
        n = rng.randint(80, 400)
        g = H.random_genome(rng, n, n_frac=rng.choice([0, 0, 0.02]), lower_frac=rng.choice([0, 0.05]),
                            other_frac=rng.choice([0, 0.01]))
        npairs = rng.randint(1, 3)
        pairs = []
        tw = rng.choice([0, 1, 3])
        mm = rng.choice([0, 1, 2, 3])
        for _ in range(npairs):
            la, lb = rng.randint(4, 24), rng.randint(4, 24)
            a = H.random_primer(rng, la, rng.choice([0, 0.1]))
            b = H.random_primer(rng, lb, rng.choice([0, 0.1]))
            length = rng.randint(la + lb, la + lb + 20)
            pos = rng.randint(0, n - length)
            H.plant_amplicon(rng, g, pos, a, b, length, rng.randint(0, mm), rng.randint(0, mm), tw,
                             revcomp_strand=rng.random() < 0.5)
            pairs.append((a, b, rng.choice([0, 10]), rng.choice([0, 200])))
        cfgd = dict(max_mm=mm, terminal_window=tw, min_len=rng.choice([0, 1, 20]),
                    max_len=rng.choice([0, 100, 500]), seed_len=rng.choice([0, 4, 12]),
                    circular=rng.random() < 0.3, hit_cap=0)
        _equiv(oracle, cfgd, pairs, bytes(g))
