"""End-to-end text parity of the four tools' hot flow (ipcr, ipcr-probe, ipcr-nested,
ipcr-multiplex): FASTA on disk -> sorted TSV, GPU engine + device visitors
(ipcr_b200.hostlib) against the oracle-driven restatement of the reference host (ref_host)."""
import random

import pytest

import ipcr_test_helpers as H
import ref_host as R
from ipcr_b200 import Pair
from ipcr_b200 import hostlib as HL


def _opts(**kw):
    return HL.Options(**kw)


# ---- CPU: host policy pieces against the reference's own unit vectors ----------------------
def test_split_chunk_suffix():
    # internal/common/ids.go:11-27 (+ its test table)
    assert HL.split_chunk_suffix("chr1:100-200") == ("chr1", 100, True)
    assert HL.split_chunk_suffix("chr1") == ("chr1", 0, False)
    assert HL.split_chunk_suffix("a:b:5-9") == ("a:b", 5, True)
    assert HL.split_chunk_suffix("x:") == ("x:", 0, False)
    assert HL.split_chunk_suffix("x:12") == ("x:12", 0, False)
    assert HL.split_chunk_suffix("x:1_0-3") == ("x:1_0-3", 0, False)
    for sid in ("chr1:100-200", "s", "a:-5-3", "q:+7-9", "w: 1-2"):
        assert HL.split_chunk_suffix(sid) == R.split_chunk_suffix(sid)


def test_rolling_chunks_match_reference_window():
    rng = random.Random(1)
    for _ in range(200):
        n = rng.randint(0, 120)
        seq = bytes(rng.choice(b"ACGT") for _ in range(n))
        cs, ov = rng.randint(-2, 40), rng.randint(-3, 45)
        assert list(HL.rolling_chunks("r", seq, cs, ov)) == R.stream_chunks("r", seq, cs, ov)


def test_validate_chunking_and_windows():
    # internal/runutil/runutil.go:10-59
    assert HL.effective_terminal_window(0) == 0 and HL.effective_terminal_window(-1) == 0
    assert HL.effective_terminal_window(3) == 3
    assert HL.validate_chunking(False, 0, 2000, 20)[:2] == (0, 0)
    assert HL.validate_chunking(True, 5000, 2000, 20)[:2] == (0, 0)
    assert HL.validate_chunking(False, 5000, 0, 20)[:2] == (0, 0)
    assert HL.validate_chunking(False, 2000, 2000, 20)[:2] == (0, 0)
    assert HL.validate_chunking(False, 5000, 2000, 20)[:2] == (5000, 2000)


def test_self_pair_expansion():
    ps = [Pair("x", "acgt", "GGCC", 5, 50), Pair("y", "ACGT", "TTAA")]
    assert [(p.id, p.forward, p.reverse) for p in HL.add_self_pairs(ps)] == \
        [(t[0], t[1], t[2]) for t in R.add_self_pairs([("x", "acgt", "GGCC", 5, 50), ("y", "ACGT", "TTAA", 0, 0)])]
    assert [(p.id, p.forward, p.reverse) for p in HL.add_self_pairs_unique(ps)] == \
        [(t[0], t[1], t[2]) for t in R.add_self_pairs_unique([("x", "acgt", "GGCC", 5, 50), ("y", "ACGT", "TTAA", 0, 0)])]


# ---- GPU: whole-flow text parity -------------------------------------------------------------
def _write(tmp_path, name, recs, width=60):
    p = tmp_path / name
    with open(p, "wb") as f:
        for rid, seq in recs:
            f.write(b">" + rid.encode() + b" desc\n")
            for i in range(0, len(seq), width):
                f.write(seq[i:i + width] + b"\n")
    return str(p)


@pytest.mark.gpu
def test_reference_integration_chunked_equals_unchunked(tmp_path, oracle):
    # internal/integration/integration_test.go:125-225: ACGTAC/ACGTAC, --max-length 8 on (ACGT)x7,
    # --chunk-size 16 == no chunking under --sort; SURVEY 8c: 36 rows
    fa = tmp_path / "chunk.fa"
    fa.write_bytes(b">s\nACGTACGTACGTACGTACGTACGTACGT\n")
    outs = {}
    for chunk in (0, 16):
        o = _opts(max_length=8, sort=True, header=False, chunk_size=chunk)
        outs[chunk] = HL.run_ipcr(o, [str(fa)], fwd="ACGTAC", rev="ACGTAC")
        want = R.run(oracle, "ipcr", [str(fa)], dict(max_length=8, sort=True, header=False, chunk_size=chunk),
                     fwd="ACGTAC", rev="ACGTAC")
        assert outs[chunk] == want
    assert outs[0] == outs[16]
    assert len(outs[0].splitlines()) == 36


@pytest.mark.gpu
def test_config1_demo_inputs(tmp_path, oracle):
    # BASELINE config 1: 27F/1492R --mismatches 1 on demo.fa => header only; and the bundled TSV primers
    fa = tmp_path / "demo.fa"
    fa.write_bytes(b">s\nACGTACGTACGT\n")
    o = _opts(mismatches=1, sort=True)
    got = HL.run_ipcr(o, [str(fa)], fwd="AGAGTTTGATCMTGGCTCAG", rev="TACGGYTACCTTGTTAYGACTT")
    assert got == HL.TSV_HEADER + "\n"
    got = HL.run_ipcr(o, [str(fa)], pairs=[Pair("X", "AGRGTTYGATYMTGGCTCAG", "RGYTACCTTGTTACGACTT", 0, 2000)])
    assert got == HL.TSV_HEADER + "\n"
    # core/engine/engine_test.go:11-35 through the whole flow: 12 rows x 3 experiments
    got = HL.run_ipcr(_opts(sort=True, terminal_window=0), [str(fa)], fwd="ACG", rev="ACG")
    want = R.run(oracle, "ipcr", [str(fa)], dict(sort=True, terminal_window=0), fwd="ACG", rev="ACG")
    assert got == want and len(got.splitlines()) == 1 + 36


def _genome_with_amplicons(rng, n, a, b, k, probe=None, inner=None, n_frac=0.001):
    g = H.random_genome(rng, n, n_frac=n_frac, lower_frac=0.1)
    for i in range(k):
        pos = 200 + i * (n - 2500) // max(1, k)
        length = rng.randint(300, 1800)
        H.plant_amplicon(rng, g, pos, a, b, length, rng.randint(0, 2), rng.randint(0, 2), 3, revcomp_strand=i % 2 == 1)
        if probe:
            pc = H.mutate(rng, H.concrete(rng, probe), rng.randint(0, 1), 0, len(probe))
            pc = pc if i % 3 else H.revcomp(pc)
            g[pos + 60:pos + 60 + len(pc)] = pc.encode()
        if inner:
            ia, ib = inner
            H.plant_amplicon(rng, g, pos + 100, ia, ib, rng.randint(60, 150), 0, rng.randint(0, 1), 3,
                             revcomp_strand=i % 4 == 1)
    return bytes(g)


@pytest.mark.gpu
@pytest.mark.parametrize("circular", [False, True])
def test_ipcr_text_parity(tmp_path, oracle, circular):
    rng = random.Random(5)
    a, b = "AGAGTTTGATCMTGGCTCAG", "TACGGYTACCTTGTTAYGACTT"
    recs = [(f"chr{i}", _genome_with_amplicons(rng, rng.randint(20000, 60000), a, b, 4)) for i in range(4)]
    fa = _write(tmp_path, "g.fa", recs)
    for sort in (True, False):
        for chunk in (0, 9000):
            od = dict(mismatches=2, sort=sort, circular=circular, chunk_size=chunk)
            got = HL.run_ipcr(_opts(**od), [fa], fwd=a, rev=b)
            want = R.run(oracle, "ipcr", [fa], od, fwd=a, rev=b)
            assert got == want
            assert len(got.splitlines()) > 16


@pytest.mark.gpu
def test_probe_text_parity(tmp_path, oracle):
    # BASELINE config 3 flow: primers + TaqMan-style probe, --mismatches 2, --probe-max-mm 0/1
    rng = random.Random(6)
    a, b = "AGAGTTTGATCMTGGCTCAG", "TACGGYTACCTTGTTAYGACTT"
    probe = "ACTGGCTCACCTTGCGAGTCATGC"
    recs = [(f"contig{i}", _genome_with_amplicons(rng, 40000, a, b, 5, probe=probe)) for i in range(3)]
    fa = _write(tmp_path, "p.fa", recs)
    for pmm, req, prb in ((0, True, probe), (1, True, probe), (1, False, probe), (0, False, "ACTGGYTCACCNTGCGAGTCATGC")):
        od = dict(mismatches=2, sort=True)
        got = HL.run_probe(_opts(**od), [fa], a, b, prb, "taq", pmm, req)
        want = R.run(oracle, "probe", [fa], od, fwd=a, rev=b, probe=prb, probe_name="taq", probe_max_mm=pmm,
                     require_probe=req)
        assert got == want
        assert "true" in got


@pytest.mark.gpu
def test_nested_text_parity(tmp_path, oracle):
    # BASELINE config 4 flow: outer 27F-1492R + inner pair, --mismatches 1
    rng = random.Random(7)
    a, b = "AGAGTTTGATCMTGGCTCAG", "TACGGYTACCTTGTTAYGACTT"
    ia, ib = "GCAACGAGCGCAACCC", "ACGGGCGGTGTGTRC"
    recs = [(f"c{i}", _genome_with_amplicons(rng, 50000, a, b, 5, inner=(ia, ib), n_frac=0)) for i in range(3)]
    fa = _write(tmp_path, "n.fa", recs)
    for req in (False, True):
        od = dict(mismatches=1, sort=True)
        got = HL.run_nested(_opts(**od), [fa], a, b, ia, ib, require_inner=req)
        want = R.run(oracle, "nested", [fa], od, fwd=a, rev=b, inner_fwd=ia, inner_rev=ib, require_inner=req)
        assert got == want
        assert "\ttrue\t" in got
    # internal/visitors/nested_test.go:9-37 through the device path
    fa2 = tmp_path / "amp.fa"
    fa2.write_bytes(b">amp\nAAAACACACACGGGACACACTTTACCCC\n")
    od = dict(mismatches=0, terminal_window=0, sort=True, self_pairs=False, max_length=0)
    got = HL.run_nested(_opts(**od), [str(fa2)], "AAAA", "GGGG", "AAA", "CCC")
    want = R.run(oracle, "nested", [str(fa2)], od, fwd="AAAA", rev="GGGG", inner_fwd="AAA", inner_rev="CCC")
    assert got == want


@pytest.mark.gpu
def test_multiplex_text_parity(tmp_path, oracle):
    from ipcr_b200 import synth

    rng = random.Random(8)
    user, _all = synth.multiplex_panel(12, seed=3, self_pairs=False)
    n = 120000
    g = H.random_genome(rng, n)
    for i, p in enumerate(user):
        H.plant_amplicon(rng, g, 1000 + i * 9000, p.forward, p.reverse, rng.randint(100, 1500), rng.randint(0, 3),
                         rng.randint(0, 3), 3, revcomp_strand=i % 2 == 1)
    fa = _write(tmp_path, "m.fa", [("chrM", bytes(g))])
    od = dict(mismatches=3, sort=True)
    got = HL.run_multiplex(_opts(**od), [fa], user)
    want = R.run(oracle, "multiplex", [fa], od, pairs=[(p.id, p.forward, p.reverse, 0, 0) for p in user])
    assert got == want
    assert len(got.splitlines()) >= 13
