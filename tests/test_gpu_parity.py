"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle.

Bar: bit-exact -- same products, same mismatch counts and positions, and the
same EMISSION ORDER as the reference's production path (oracle layer 2,
compiled.go:162-267); additionally equal as a multiset to the reference's own
brute-force oracle (bruteforce.go:11-38) whenever the hit cap does not bind.
"""
import os
import random

import numpy as np
import pytest

import ipcr_test_helpers as H
from ipcr_b200 import Config, Engine, Pair, SimulationScratch, _capi
from ipcr_b200 import device as D

pytestmark = pytest.mark.gpu

CASES = H.load_reference_cases()


def _tuple(p):
    return (p.experiment_id, p.type, p.start, p.end, p.length, p.fwd_mm, p.rev_mm,
            tuple(p.fwd_mismatch_idx), tuple(p.rev_mismatch_idx))


def _otuple(pairs, t):
    return (pairs[t[0]].id,) + tuple(t[1:])


def _mk_pairs(raw):
    return [Pair(f"p{i}", p[0], p[1], p[2] if len(p) > 2 else 0, p[3] if len(p) > 3 else 0)
            for i, p in enumerate(raw)]


def _check(oracle, cfgd, raw_pairs, seq, jit=0, cap_binds=False):
    pairs = _mk_pairs(raw_pairs)
    eng = Engine(Config(**cfgd))
    cp = eng.compile_panel(pairs)
    ocfg = oracle.make_cfg(**cfgd)
    want = [_otuple(pairs, t) for t in oracle.simulate_compiled(ocfg, [tuple(p) for p in raw_pairs], seq)]
    # every case runs through the requested bit-parallel filter AND through the seed index (3); the
    # seed index through all three stage-2 forms: the asynchronous chain in bucket mode (default), the
    # chain's general (bitonic) mode, and the multi-kernel path
    for mode in dict.fromkeys((jit, 3)):
        cp.set_jit(mode)
        for env in ((None,) if mode != 3 else (None, "IPCR_NO_BUCKETS", "IPCR_NO_CHAIN")):
            if env:
                os.environ[env] = "1"
            try:
                got = [_tuple(p) for p in eng.simulate_compiled("seq", seq, cp)]
            finally:
                if env:
                    del os.environ[env]
            assert got == want, (mode, env, cfgd, raw_pairs, bytes(seq)[:80])
    # small panels take the strided first level of the seed index (anchor stride 4) by default; the same case
    # through the stride-1 and stride-2 kernels (the index is built at a panel's first scan, so: fresh panels)
    for stride in ("1", "2"):
        os.environ["IPCR_SEED_STRIDE_FORCE"] = stride
        try:
            cp2 = eng.compile_panel(pairs)
            cp2.set_jit(3)
            got2 = [_tuple(p) for p in eng.simulate_compiled("seq", seq, cp2)]
        finally:
            del os.environ["IPCR_SEED_STRIDE_FORCE"]
        assert got2 == want, ("stride", stride, cfgd, raw_pairs, bytes(seq)[:80])
    if not cap_binds:
        brute = [_otuple(pairs, t) for t in oracle.simulate_bruteforce(ocfg, [tuple(p) for p in raw_pairs], seq)]
        assert H.multiset(got) == H.multiset(brute)
    return got


def test_device_is_blackwell():
    L = _capi.lib()
    assert L.ipcr_gpu_device_count() >= 1
    err = _capi.errbuf()
    assert L.ipcr_gpu_init(0, err, len(err)) == 0, err.value


@pytest.mark.parametrize("case", CASES["engine_known_answers"], ids=lambda c: c["name"])
def test_reference_known_answers(oracle, case):
    got = _check(oracle, case["cfg"], case["pairs"], case["seq"].encode())
    e = case["expect"]
    if "count" in e:
        assert len(got) == e["count"]
    if "min_count" in e:
        assert len(got) >= e["min_count"]
    if "first" in e:
        f = got[0]
        d = dict(start=f[2], end=f[3], length=f[4], fwd_mm=f[5], rev_mm=f[6], fwd_idx=list(f[7]), rev_idx=list(f[8]))
        for k, v in e["first"].items():
            assert d[k] == v
    if e.get("wrap"):
        assert got[0][2] > got[0][3]


def test_reference_equivalence_matrix(oracle):
    eq = CASES["equivalence"]
    n = 0
    for c in eq["cases"]:
        for cfgd in eq["configs"]:
            _check(oracle, cfgd, c["pairs"], c["seq"].encode())
            n += 1
    assert n == 81
    for c in CASES["equivalence_extra"]:
        for s in c.get("seqs") or [c["seq"]]:
            for cfgd in c["configs"]:
                _check(oracle, cfgd, c["pairs"], s.encode())


@pytest.mark.parametrize("case", CASES["bench_fixture"]["cases"], ids=lambda c: str(c["pair_count"]))
def test_reference_bench_fixture(oracle, case):
    seq, pairs = H.make_engine_bench_fixture(case["pair_count"], case["genome_len"], case["mutate_forward"],
                                             case["reference_n"])
    got = _check(oracle, case["cfg"], pairs, seq)
    for i in range(len(pairs)):
        assert any(p[0] == f"p{i}" and p[1] == "forward" and p[2] == 128 + 256 * i and p[4] == 180 for p in got)


def test_survey_derived_rows(oracle):
    for c in CASES["survey_derived"]["cases"]:
        got = _check(oracle, c["cfg"], c["pairs"], c["seq"].encode())
        want = [("p%d" % p[0], p[1], p[2], p[3], p[4], p[5], p[6], tuple(p[7]), tuple(p[8])) for p in c["products"]]
        # SURVEY lists brute-force order; the production path emits the halo-found site
        # last (compiled.go:211-232), which _check already pinned against oracle layer 2
        assert sorted(got) == sorted(want)


def test_empty_and_short_inputs(oracle):
    cfg = dict(max_mm=1, terminal_window=3)
    for seq in (b"", b"A", b"ACGTACGTACGT"):
        # README.md:46 primers are longer than the record => header only
        _check(oracle, cfg, [("AGAGTTTGATCMTGGCTCAG", "TACGGYTACCTTGTTAYGACTT")], seq)
    eng = Engine(Config())
    assert eng.simulate_batch("s", b"ACGT", []) == []


def test_pack_planes_match_host_packing():
    rng = random.Random(3)
    seq = bytes(H.random_genome(rng, 70001, n_frac=0.01, lower_frac=0.02, other_frac=0.01))
    eng = Engine(Config())
    cp = eng.compile_panel([Pair("x", "ACGTACGT", "TTGACC")])
    cp.set_jit(0)  # the bit-parallel filter reads bit-planes
    sc = SimulationScratch(cp)
    g = D.Genome.upload(sc, [(seq, 0)])
    W, nt, planes = g.planes()
    assert nt >= 1 and W >= 32
    a = np.frombuffer(seq, dtype=np.uint8)
    total = len(seq)
    rows = W + 31
    # position of (tile, row, lane, bit)
    t, r, l, i = np.meshgrid(np.arange(nt), np.arange(rows), np.arange(32), np.arange(32), indexing="ij")
    P = ((t * 1024 + l * 32 + i) * W + r).astype(np.int64)
    valid = P < total
    byte = np.where(valid, a[np.minimum(P, total - 1)], 0)
    for k, base in enumerate(b"ACGT"):
        want = ((byte == base) & valid).astype(np.uint64)
        word = (want << np.arange(32, dtype=np.uint64)).sum(axis=3).astype(np.uint32)
        assert np.array_equal(word, planes[..., k]), "ACGT"[k]
    assert g.records() == [(total, True)]
    g2 = D.Genome.upload(sc, [(b"ACGTacgtACGT", 0), (b"ACGTNACGT", 1), (b"ACGTACGT", 2)])
    assert g2.records() == [(12, False), (9, True), (8, False)]


def test_pack_codes_match_host_packing():
    """k_pack2 (seed-index format): 2-bit codes, hard-mismatch bitmap and reset flags against a
    numpy restatement, incl. lowercase, N, other bytes and the ragged tail / padding."""
    rng = random.Random(5)
    for n in (70001, 4096, 33, 1):
        seq = bytes(H.random_genome(rng, n, n_frac=0.01, lower_frac=0.02, other_frac=0.01))
        eng = Engine(Config())
        cp = eng.compile_panel([Pair("x", "ACGTACGT", "TTGACC")])
        cp.set_jit(3)
        sc = SimulationScratch(cp)
        g = D.Genome.upload(sc, [(seq, 0)])
        codes, inval = g.codes()
        ng = len(inval)
        assert ng == (n + 4095) // 4096 * 128 + 1 and len(codes) == 2 * ng
        a = np.zeros(ng * 32, dtype=np.uint8)
        a[:n] = np.frombuffer(seq, dtype=np.uint8)
        valid = np.isin(a, np.frombuffer(b"ACGT", dtype=np.uint8))
        want_inv = (~valid).reshape(ng, 32).astype(np.uint64)
        want_inv = (want_inv << np.arange(32, dtype=np.uint64)).sum(axis=1).astype(np.uint32)
        assert np.array_equal(inval, want_inv)
        c2 = ((a >> 1) & 3).reshape(2 * ng, 16).astype(np.uint64)
        want_codes = (c2 << (2 * np.arange(16, dtype=np.uint64))).sum(axis=1).astype(np.uint32)
        # codes of invalid bytes are don't-care for the scan but deterministic: (byte >> 1) & 3
        assert np.array_equal(codes, want_codes)
        has_reset = any(chr(c) not in "ACGTacgt" for c in seq)
        assert g.records() == [(n, has_reset)]
        W, nt, planes = g.planes()
        assert nt == 0  # this handle holds codes, not planes


def test_dirty_input_falls_back_to_bit_parallel(oracle, monkeypatch):
    """An input with too many non-ACGT seed windows leaves the seed index for the bit-parallel
    filter inside the same call (threshold lowered here); the result is the same."""
    rng = random.Random(31)
    a, b = "AGAGTTTGATCMTGGCTCAG", "TACGGYTACCTTGTTAYGACTT"
    g = H.random_genome(rng, 60000, n_frac=0.02)
    for k in range(6):
        H.plant_amplicon(rng, g, 500 + k * 9000, a, b, 700, rng.randint(0, 2), rng.randint(0, 2), 3,
                         revcomp_strand=k % 2 == 1)
    seq = bytes(g)
    cfgd = dict(max_mm=2, terminal_window=3, max_len=2000, hit_cap=0)
    pairs = [Pair("16S", a, b), Pair("16S+A:self", a, a)]
    raw = [(p.forward, p.reverse, 0, 0) for p in pairs]
    want = [_otuple(pairs, t) for t in oracle.simulate_compiled(oracle.make_cfg(**cfgd), raw, seq)]
    assert want
    eng = Engine(Config(**cfgd))
    cp = eng.compile_panel(pairs)
    sc = eng.new_simulation_scratch(cp)
    got = [_tuple(p) for p in eng.simulate_compiled_with_scratch("s", seq, cp, sc)]
    assert got == want and sc.timing()["used_jit"] == 2  # seed index handles a few thousand dirty windows
    monkeypatch.setenv("IPCR_SEED_DIRTY_MAX", "100")
    got = [_tuple(p) for p in eng.simulate_compiled_with_scratch("s", seq, cp, sc)]
    assert got == want and sc.timing()["used_jit"] in (0, 1)
    monkeypatch.delenv("IPCR_SEED_DIRTY_MAX")
    got = [_tuple(p) for p in eng.simulate_compiled_with_scratch("s", seq, cp, sc)]
    assert got == want and sc.timing()["used_jit"] == 2


def test_candidates_cover_all_true_sites(oracle):
    rng = random.Random(11)
    n = 200000
    g = H.random_genome(rng, n, n_frac=0.001)
    a, b = "AGAGTTTGATCMTGGCTCAG", "TACGGYTACCTTGTTAYGACTT"
    for k in range(20):
        H.plant_amplicon(rng, g, 1000 + k * 9000, a, b, 1500, rng.randint(0, 2), rng.randint(0, 2), 3,
                         revcomp_strand=k % 2 == 1)
    seq = bytes(g)
    eng = Engine(Config(max_mm=2, terminal_window=3, max_len=2000))
    cp = eng.compile_panel([Pair("16S", a, b)])
    cp.set_jit(0)
    sc = SimulationScratch(cp)
    prods = eng.simulate_compiled_with_scratch("s", seq, cp, sc)
    assert len(prods) >= 20
    cands = set(sc.candidates())
    for q in range(cp.num_orientations):
        s, t = cp.orientation(q)
        truth = oracle.find_matches(seq, s, 2, 0, 0 if t == "a" else 3, left_tw=3 if t == "a" else 0)
        assert truth
        for pos, _mm, _idx in truth:
            assert (q, pos) in cands
    # and the filter is selective: far fewer candidates than positions
    assert len(cands) < n // 100


def _random_case(rng, big=False):
    n = rng.randint(2000, 60000) if big else rng.randint(80, 700)
    g = H.random_genome(rng, n, n_frac=rng.choice([0, 0, 0.01, 0.05]), lower_frac=rng.choice([0, 0.03]),
                        other_frac=rng.choice([0, 0.005]))
    mm, tw = rng.choice([0, 1, 2, 3]), rng.choice([0, 1, 3])
    pairs = []
    for _ in range(rng.randint(1, 4)):
        la, lb = rng.randint(4, 28), rng.randint(4, 28)
        a = H.random_primer(rng, la, rng.choice([0, 0.1]))
        b = H.random_primer(rng, lb, rng.choice([0, 0.1]))
        for _k in range(rng.randint(1, 3)):
            length = rng.randint(la + lb, la + lb + 40)
            H.plant_amplicon(rng, g, rng.randint(0, n - length), a, b, length, rng.randint(0, mm), rng.randint(0, mm),
                             tw, revcomp_strand=rng.random() < 0.5)
        pairs.append((a, b, rng.choice([0, 0, 12]), rng.choice([0, 0, 150])))
        if rng.random() < 0.3:
            pairs.append((a, a, 0, 0))  # self pair (internal/common/primers.go:11-38)
    cfgd = dict(max_mm=mm, terminal_window=tw, min_len=rng.choice([0, 1, 30]), max_len=rng.choice([0, 120, 2000]),
                seed_len=rng.choice([0, 4, 12]), circular=rng.random() < 0.3, hit_cap=0)
    return cfgd, pairs, bytes(g)


def test_randomized_parity_small(oracle):
    rng = random.Random(2024)
    for _ in range(120):
        cfgd, pairs, seq = _random_case(rng)
        _check(oracle, cfgd, pairs, seq)


def test_randomized_parity_medium(oracle):
    rng = random.Random(77)
    for _ in range(12):
        cfgd, pairs, seq = _random_case(rng, big=True)
        _check(oracle, cfgd, pairs, seq)


def test_jit_kernel_parity_randomized(oracle):
    """The NVRTC-specialised scan kernel (forced, jit=2) against the oracle."""
    rng = random.Random(4242)
    for it in range(10):
        cfgd, pairs, seq = _random_case(rng, big=it % 2 == 0)
        _check(oracle, cfgd, pairs, seq, jit=2)
    # many orientations -> several groups, chunked rows
    rng = random.Random(9)
    n = 400000
    g = H.random_genome(rng, n, n_frac=0.0002)
    pairs = []
    for i in range(24):
        a = H.random_primer(rng, rng.randint(18, 25), 0.05)
        b = H.random_primer(rng, rng.randint(18, 25), 0.05)
        H.plant_amplicon(rng, g, 2000 + i * 15000, a, b, rng.randint(80, 1500), rng.randint(0, 3), rng.randint(0, 3), 3,
                         revcomp_strand=i % 2 == 1)
        pairs.append((a, b, 0, 0))
    got = _check(oracle, dict(max_mm=3, terminal_window=3, max_len=2000, hit_cap=10000), pairs, bytes(g), jit=2)
    assert len({p[0] for p in got}) == 24


def test_hit_cap_semantics(oracle):
    # dense short primers so that the cap binds; covers seeded, forced-fallback (reset byte)
    # and unseeded-orientation paths (compiled.go:185-190,238-257)
    rng = random.Random(5)
    for it in range(60):
        n = rng.randint(60, 400)
        g = H.random_genome(rng, n, n_frac=rng.choice([0, 0.02]))
        a = H.random_primer(rng, rng.randint(3, 5), rng.choice([0, 0.3]))
        b = H.random_primer(rng, rng.randint(3, 5), rng.choice([0, 0.3]))
        cfgd = dict(max_mm=rng.choice([0, 1, 2]), terminal_window=rng.choice([0, 1, 2]), min_len=0,
                    max_len=rng.choice([0, 40]), seed_len=rng.choice([-1, 0, 3]), hit_cap=rng.choice([1, 2, 5, 30]),
                    circular=rng.random() < 0.3)
        _check(oracle, cfgd, [(a, b, 0, 0)], bytes(g), cap_binds=True)


def test_dense_products(oracle):
    seq = b"ACGT" * 500
    got = _check(oracle, dict(max_mm=0, terminal_window=0, max_len=60), [("ACG", "ACG")], seq)
    assert len(got) > 5000


def _n_run_genome(rng, a, b):
    """9 kbp with four N runs; between consecutive runs an amplicon whose forward site starts k
    bases INSIDE the run before it (its 5' end is N) and whose reverse site ends k bases inside
    the next run, k = 1, 2, 3."""
    runs = ((700, 64), (2000, 200), (4000, 1000), (7000, 33))
    g = H.random_genome(rng, 9000)
    for start, ln in runs:
        g[start:start + ln] = b"N" * ln
    ca, cb = a.replace("M", "C"), H.revcomp(b.replace("Y", "T"))
    for k in (1, 2, 3):
        (s0, l0), (s1, _l1) = runs[k - 1], runs[k]
        e0 = s0 + l0
        g[e0:e0 + len(ca) - k] = ca[k:].encode()        # forward primer, first k bases fall into the run
        g[s1 - (len(cb) - k):s1] = cb[:len(cb) - k].encode()  # rc(reverse) site, last k bases fall into the run
    return bytes(g)


def test_long_n_runs(oracle):
    """Assembly gaps: primer sites overlapping a run boundary by 1..3 bases (genome N = one
    mismatch each, outside the 3' window), runs longer than the scan kernel's 64-position view."""
    rng = random.Random(8)
    a, b = "AGAGTTTGATCMTGGCTCAG", "TACGGYTACCTTGTTAYGACTT"
    seq = _n_run_genome(rng, a, b)
    for mm in (1, 2, 3):
        cfgd = dict(max_mm=mm, terminal_window=3, max_len=3000, hit_cap=0)
        got = _check(oracle, cfgd, [(a, b, 0, 0), (a, a, 0, 0), (b, b, 0, 0)], seq)
        assert len(got) == mm  # the amplicon with k N's per site needs max_mm >= k
        _check(oracle, dict(cfgd, terminal_window=0, hit_cap=100), [(a, b, 0, 0)], seq, cap_binds=True)


def test_both_bitmap_widths(oracle, monkeypatch):
    """The shared-memory bitmap sets one or two bits per k-mer (chosen by the cost model from its
    fill); force each on the same inputs."""
    rng = random.Random(606)
    cases = [_random_case(rng, big=True) for _ in range(3)]
    for k in ("1", "2"):
        monkeypatch.setenv("IPCR_SEED_BLOOM_K", k)
        for cfgd, pairs, seq in cases:
            _check(oracle, cfgd, pairs, seq, jit=3)
    monkeypatch.delenv("IPCR_SEED_BLOOM_K")


def test_high_mismatch_budgets_on_n_rich_input(oracle):
    """max_mm 4..5 with many N: dirty windows holding more than four invalid bytes (no replacement
    lookups, every orientation verified) and huge variant sets."""
    rng = random.Random(515)
    for it in range(6):
        n = rng.randint(300, 900)
        g = H.random_genome(rng, n, n_frac=rng.choice([0.05, 0.15, 0.3]))
        mm, tw = rng.choice([4, 5]), rng.choice([0, 2, 3])
        la, lb = rng.randint(8, 30), rng.randint(8, 30)
        a, b = H.random_primer(rng, la, rng.choice([0, 0.15])), H.random_primer(rng, lb, 0)
        for _ in range(3):
            H.plant_amplicon(rng, g, rng.randint(0, n - la - lb - 10), a, b, la + lb + rng.randint(0, 10),
                             rng.randint(0, mm), rng.randint(0, mm), tw, revcomp_strand=rng.random() < 0.5)
        cfgd = dict(max_mm=mm, terminal_window=tw, max_len=200, hit_cap=rng.choice([0, 0, 7]))
        _check(oracle, cfgd, [(a, b, 0, 0)], bytes(g), cap_binds=cfgd["hit_cap"] > 0)


def test_dirty_window_overflow_path(oracle, monkeypatch):
    """> 16 orientations share the k-mers of one dirty seed window (k_seed_dirty_probe gives up
    on it and k_seed_dirty_expand verifies every orientation there)."""
    monkeypatch.setenv("IPCR_SEED_LEN", "10")
    rng = random.Random(12)
    raw_pairs, seq = H.shared_seed_case(rng)
    got = _check(oracle, dict(max_mm=2, terminal_window=3, max_len=400, hit_cap=0), raw_pairs, seq)
    assert len(got) >= 12
    _check(oracle, dict(max_mm=2, terminal_window=3, max_len=400, hit_cap=5), raw_pairs, seq, cap_binds=True)


def test_many_sites_take_the_multi_block_sort(oracle):
    """> 8192 sites: the device-wide radix sort / multi-level scan instead of the single-block
    kernels used for small site lists."""
    seq = (b"ACGT" * 9000) + b"TTTTT" + (b"GGCA" * 3000)
    got = _check(oracle, dict(max_mm=0, terminal_window=0, max_len=16), [("ACG", "ACG"), ("GCA", "GGC")], seq)
    assert len(got) > 20000


def test_need_sites_strings():
    seq = b"TTTACGTACAAAAGGTACCTTTGGGACGTATAAAAGGTACCAAA"
    eng = Engine(Config(max_mm=1, terminal_window=0, min_len=6, max_len=80, need_sites=True))
    prods = eng.simulate("seq", seq, Pair("j", "ACGTAC", "GGTACC", 6, 80))
    assert prods
    for p in prods:
        assert p.fwd_site == seq[p.start:p.start + len(p.fwd_primer)].decode()
        assert len(p.rev_site) == len(p.rev_primer)


def test_scratch_reuse_and_batch(oracle):
    raw = [("ACGTAC", "GGTACC", 6, 60)]
    pairs = _mk_pairs(raw)
    cfgd = dict(max_mm=1, terminal_window=0, min_len=6, max_len=60, seed_len=4)
    eng = Engine(Config(**cfgd))
    cp = eng.compile_panel(pairs)
    sc = eng.new_simulation_scratch(cp)
    seqs = [b"TTTACGTACAAAAGGTACCTTT", b"TTTACGTATAAAAGGTACCTTT", b"TTTGGTACCAAAAGTACGTTTT",
            b"TTTACGTACAAAACCCCCCCTTT", b"", b"ACG"]
    ocfg = oracle.make_cfg(**cfgd)
    per_rec = []
    for s in seqs:
        got = [_tuple(p) for p in eng.simulate_compiled_with_scratch("seq", s, cp, sc)]
        want = [_otuple(pairs, t) for t in oracle.simulate_compiled(ocfg, raw, s)]
        assert got == want
        per_rec.append(want)
    batch = D.scan_batch(cp, sc, seqs, seq_ids=[f"r{i}" for i in range(len(seqs))])
    flat = [(f"r{i}",) + w for i, ws in enumerate(per_rec) for w in ws]
    assert [(p.sequence_id,) + _tuple(p) for p in batch] == flat


def test_split_phase_shards_equal_single_scan(oracle):
    rng = random.Random(31)
    n = 300000
    g = H.random_genome(rng, n, n_frac=0.0005)
    a, b = "AGAGTTTGATCMTGGCTCAG", "TACGGYTACCTTGTTAYGACTT"
    for k in range(12):
        H.plant_amplicon(rng, g, 500 + k * 24000, a, b, 1400 + k, rng.randint(0, 2), rng.randint(0, 2), 3,
                         revcomp_strand=k % 3 == 0)
    # one amplicon straddling every shard boundary
    shards = 4
    cut = [n * k // shards for k in range(shards + 1)]
    for c in cut[1:-1]:
        H.plant_amplicon(rng, g, c - 700, a, b, 1400, 1, 1, 3)
        g[c - 10:c + 10] = (a.replace("M", "A")).encode()  # a site straddling the cut itself
    seq = bytes(g)
    pairs = [Pair("16S", a, b), Pair("16S+A:self", a, a), Pair("16S+B:self", b, b)]
    for cfgd in (dict(max_mm=2, terminal_window=3, max_len=2000, hit_cap=0),
                 dict(max_mm=2, terminal_window=3, max_len=2000, hit_cap=10000),
                 dict(max_mm=2, terminal_window=3, max_len=0, circular=True)):
        eng = Engine(Config(**cfgd))
        cp = eng.compile_panel(pairs)
        sc = eng.new_simulation_scratch(cp)
        whole = [_tuple(p) for p in eng.simulate_compiled_with_scratch("0", seq, cp, sc)]
        assert whole
        halo = 64
        sites = []
        for k in range(shards):
            lo, hi = cut[k], cut[k + 1]
            piece = seq[lo:min(n, hi + halo)]
            gk = D.Genome.upload(sc, [(piece, 0, lo, hi - lo)])
            sites.append(D.find_sites(cp, sc, gk))
            gk.close()
        allsites = np.concatenate(sites)
        has_reset = b"N" in seq
        got = [_tuple(p) for p in D.pair_sites(cp, sc, allsites, [(n, has_reset)])]
        assert got == whole
        want = [_otuple(pairs, t) for t in oracle.simulate_compiled(
            oracle.make_cfg(**cfgd), [(p.forward, p.reverse, 0, 0) for p in pairs], seq)]
        assert whole == want


def test_streamed_batch_equals_plain(oracle, monkeypatch):
    """ipcr_gpu_scan_batch overlaps H2D with the scan slab by slab on large inputs; force
    the slab path on a small multi-record input and compare with the oracle, including the
    raw-redo rule when only a later slab holds a reset byte."""
    rng = random.Random(77)
    a, b = "AGAGTTTGATCMTGGCTCAG", "TACGGYTACCTTGTTAYGACTT"
    pairs = [Pair("16S", a, b), Pair("16S+A:self", a, a), Pair("16S+B:self", b, b)]
    raw = [(p.forward, p.reverse, 0, 0) for p in pairs]
    lens = [90000, 500, 140000, 70000]
    recs = []
    for ri, n in enumerate(lens):
        g = H.random_genome(rng, n)
        for k in range(max(1, n // 20000)):
            H.plant_amplicon(rng, g, 200 + k * 18000 if n > 20000 else 10, a, b, 300 + k, rng.randint(0, 2),
                             rng.randint(0, 2), 3, revcomp_strand=k % 2 == 1)
        recs.append(g)
    total = sum(lens)
    for slabs in (3, 5):
        for c in [(total * k // slabs) & ~15 for k in range(1, slabs)]:  # a site across every slab cut
            off = 0
            for ri, n in enumerate(lens):
                if off + 30 <= c < off + n - 30:
                    recs[ri][c - off - 9:c - off + 11] = a.replace("M", "C").encode()
                off += n
    recs[3][60000] = ord("N")  # reset byte only in the LAST slab
    seqs = [bytes(g) for g in recs]
    for cfgd in (dict(max_mm=2, terminal_window=3, max_len=2000, hit_cap=0),
                 dict(max_mm=2, terminal_window=3, max_len=2000, hit_cap=10000),
                 dict(max_mm=1, terminal_window=0, max_len=0, hit_cap=3, circular=True)):
        eng = Engine(Config(**cfgd))
        cp = eng.compile_panel(pairs)
        sc = eng.new_simulation_scratch(cp)
        want = []
        for ri, sq in enumerate(seqs):
            want += [(f"r{ri}",) + _otuple(pairs, t) for t in oracle.simulate_compiled(oracle.make_cfg(**cfgd), raw, sq)]
        for slabs in ("0", "3", "5"):
            monkeypatch.setenv("IPCR_STREAM_SLABS", slabs)
            got = [(p.sequence_id,) + _tuple(p) for p in D.scan_batch(cp, sc, seqs, seq_ids=[f"r{i}" for i in range(4)])]
            assert got == want, (cfgd, slabs)
        # the records are Python bytes = pageable host memory: the same through the multi-threaded pinned staging ring
        # (by default only inputs >= 64 Mbp take it), with the slab count left to the library and forced
        monkeypatch.setenv("IPCR_STAGED_H2D_MIN", "1")
        for slabs in (None, "2", "5"):
            if slabs is None:
                monkeypatch.delenv("IPCR_STREAM_SLABS")
            else:
                monkeypatch.setenv("IPCR_STREAM_SLABS", slabs)
            got = [(p.sequence_id,) + _tuple(p) for p in D.scan_batch(cp, sc, seqs, seq_ids=[f"r{i}" for i in range(4)])]
            assert got == want, (cfgd, "staged", slabs)
        monkeypatch.delenv("IPCR_STAGED_H2D_MIN")
        monkeypatch.delenv("IPCR_STREAM_SLABS")


def test_ecoli_sized_circular_config2(oracle):
    """BASELINE config 2: 27F/1492R --circular --mismatches 2 on a 4.6 Mbp synthetic genome
    with planted ~1.5 kb amplicons, one spanning the origin."""
    rng = random.Random(46)
    n = 4_600_000
    a, b = "AGAGTTTGATCMTGGCTCAG", "TACGGYTACCTTGTTAYGACTT"
    g = bytearray(np.frombuffer(b"ACGT", dtype=np.uint8)[np.random.default_rng(46).integers(0, 4, n)].tobytes())
    for k in range(6):
        H.plant_amplicon(rng, g, 100_000 + k * 700_000, a, b, 1500 + k, k % 3, (k + 1) % 3, 3, revcomp_strand=k % 2 == 1)
    # origin-spanning: forward site near the end, reverse site near the start
    fa = H.mutate(rng, H.concrete(rng, a), 1, 0, len(a) - 3)
    g[n - 800:n - 800 + len(a)] = fa.encode()
    rb = H.revcomp(H.mutate(rng, H.concrete(rng, b), 2, 0, len(b) - 3))
    g[700 - len(b):700] = rb.encode()
    # decoys: mismatch inside the 3' window, and 3 mismatches
    g[2_000_000:2_000_000 + len(a)] = (H.concrete(rng, a)[:-1] + "T").encode()
    g[2_100_000:2_100_000 + len(a)] = H.mutate(rng, H.concrete(rng, a), 3, 0, len(a) - 3).encode()
    seq = bytes(g)
    pairs = [("AGAGTTTGATCMTGGCTCAG", "TACGGYTACCTTGTTAYGACTT", 0, 0), (a, a, 0, 0), (b, b, 0, 0)]
    cfgd = dict(max_mm=2, terminal_window=3, max_len=2000, hit_cap=10000, circular=True, seed_len=12)
    for jit in (0, 2):
        got = _check(oracle, cfgd, pairs, seq, jit=jit)
        assert sum(1 for p in got if p[0] == "p0") >= 7
        assert any(p[2] > p[3] for p in got)  # the wrap product
