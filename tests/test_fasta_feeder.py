"""The native FASTA feeder against a line-by-line restatement of the reference reader
(core/fasta/scan.go:10-69, normalize.go:5-14, stream.go:125-131, path_ctx.go:40-81).
Parsing needs no GPU."""
import gzip
import os
import random

import pytest

from ipcr_b200.fasta import FastaFile

SPACE = b" \t\n\v\f\r"


def ref_reader(data: bytes):
    """streamWholeRecords over scanFASTALines (lines <= 1 MiB)."""
    recs, rid, seq = [], "", bytearray()

    def flush():
        if rid != "":
            recs.append((rid, bytes(seq)))

    pos, n = 0, len(data)
    while pos < n:
        nl = data.find(b"\n", pos)
        frag = data[pos:] if nl < 0 else data[pos:nl + 1]
        pos = n if nl < 0 else nl + 1
        if frag[:1] == b">":
            flush()
            hdr = frag[1:].strip(SPACE)
            for i, c in enumerate(hdr):
                if c in b" \t":
                    hdr = hdr[:i]
                    break
            rid = hdr.decode("latin-1")
            seq = bytearray()
        elif rid != "":
            line = frag.strip(SPACE)
            seq += line.upper()  # bytes.upper() maps a-z only (normalize.go:5-14)
    flush()
    return recs


CASES = {
    "demo": b">s\nACGTACGTACGT\n",
    "no trailing newline": b">s desc here\nACGT\nacgt",
    "crlf + lowercase + blank lines": b">chr1 x\r\nacgtn\r\n\r\nNNNN\r\n>chr2\tz\r\nAC GT\r\n",
    "leading junk before first header": b"ACGT\nGGGG\n>a\nAC\n",
    "empty id drops the record": b">\nACGT\n> \nGG\n>ok\nTT\n",
    "gt not at line start": b">a\nAC>GT\n >b\nTT\n",
    "header only + empty record": b">a\n>b\n\n>c",
    "whitespace inside and around": b">a\n  AC GT\t\n\x0bAC\x0c\n",
    "iupac digits stars kept": b">a\nRYKM*-12\nxyz\n",
    "empty file": b"",
    "only newline": b"\n\n",
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_feeder_matches_reference_reader(tmp_path, name):
    data = CASES[name]
    p = tmp_path / "x.fa"
    p.write_bytes(data)
    f = FastaFile(str(p), threads=3)
    assert f.records() == ref_reader(data)
    assert f.total_bases == sum(len(s) for _i, s in ref_reader(data))


@pytest.mark.skipif(not os.path.exists("/root/reference/demo.fa"), reason="reference tree not mounted")
def test_bundled_demo_files_of_the_reference():
    # BASELINE config 1 inputs: demo.fa = 12-bp record `s`, demo.fa.gz = a different 8-bp `s`
    assert FastaFile("/root/reference/demo.fa").records() == [("s", b"ACGTACGTACGT")]
    assert FastaFile("/root/reference/demo.fa.gz").records() == [("s", b"ACGTACGT")]


def test_gzip_by_magic_and_suffix(tmp_path):
    data = b">s\nACGTACGT\n>t\nacgtnn\n"
    p1 = tmp_path / "demo.fa.gz"
    p1.write_bytes(gzip.compress(data))
    p2 = tmp_path / "renamed.fasta"  # gzip magic without the suffix (open.go:39-49)
    p2.write_bytes(gzip.compress(data))
    for p in (p1, p2):
        assert FastaFile(str(p)).records() == [("s", b"ACGTACGT"), ("t", b"ACGTNN")]


def test_missing_file_raises():
    with pytest.raises(OSError):
        FastaFile("/nonexistent/file.fa")


def test_multithreaded_large_file(tmp_path):
    rng = random.Random(4)
    parts = []
    for r in range(40):
        parts.append(f">rec{r} len\n".encode() if r % 7 else b">\n")
        n = rng.randint(1000, 600_000)
        body = bytes(rng.choice(b"ACGTacgtNn") for _ in range(2000)) * (n // 2000 + 1)
        body = body[:n]
        w = rng.choice([60, 80, 61, 4096])
        eol = rng.choice([b"\n", b"\r\n"])
        parts.append(eol.join(body[i:i + w] for i in range(0, n, w)) + eol)
    data = b"".join(parts)
    assert len(data) > (8 << 20)
    p = tmp_path / "big.fa"
    p.write_bytes(data)
    want = ref_reader(data)
    for threads in (1, 2, 5, 16):
        assert FastaFile(str(p), threads=threads).records() == want


@pytest.mark.gpu
def test_scan_fasta_matches_per_record_oracle(tmp_path, oracle, monkeypatch):
    import ipcr_test_helpers as H
    from ipcr_b200 import Config, Engine, Pair

    rng = random.Random(21)
    a, b = "AGAGTTTGATCMTGGCTCAG", "TACGGYTACCTTGTTAYGACTT"
    pairs = [Pair("16S", a, b), Pair("16S+A:self", a, a), Pair("16S+B:self", b, b)]
    recs = []
    text = bytearray()
    for r in range(6):
        n = rng.randint(3000, 90000)
        g = H.random_genome(rng, n, n_frac=0.001 if r % 2 else 0, lower_frac=0.3)
        H.plant_amplicon(rng, g, 100, a, b, 1500, 1, 2, 3, revcomp_strand=r % 2 == 1)
        text += f">contig{r} some description\n".encode()
        for i in range(0, n, 70):
            text += bytes(g[i:i + 70]) + b"\n"
        recs.append((f"contig{r}", bytes(g).upper()))
    p = tmp_path / "g.fa"
    p.write_bytes(bytes(text))
    cfgd = dict(max_mm=2, terminal_window=3, max_len=2000, hit_cap=10000)
    eng = Engine(Config(**cfgd))
    cp = eng.compile_panel(pairs)
    sc = eng.new_simulation_scratch(cp)
    f = FastaFile(str(p), pin=True)
    assert f.records() == recs
    raw = [(q.forward, q.reverse, 0, 0) for q in pairs]
    want = []
    for rid, sq in recs:
        for t in oracle.simulate_compiled(oracle.make_cfg(**cfgd), raw, sq):
            want.append((rid, pairs[t[0]].id) + tuple(t[1:]))
    for slabs in ("0", "3"):
        monkeypatch.setenv("IPCR_STREAM_SLABS", slabs)
        got = [(q.sequence_id, q.experiment_id, q.type, q.start, q.end, q.length, q.fwd_mm, q.rev_mm,
                q.fwd_mismatch_idx, q.rev_mismatch_idx) for q in f.scan(cp, sc)]
        assert got == want and len(got) >= 6


def test_host_packer_matches_its_definition():
    """ipcr_gpu_pack_host: code = (byte >> 1) & 3 (A 0, C 1, T 2, G 3), 16 bases per word; invalid bit =
    not an uppercase A/C/G/T (core/primer/iupac.go:62-67), 32 per word; padding counts as invalid."""
    import ctypes as C

    import numpy as np

    from ipcr_b200 import _capi

    L = _capi.lib()
    rng = np.random.default_rng(5)
    for n in (0, 1, 31, 32, 33, 4095, 4096, 4097, 70001):
        a = np.frombuffer(b"ACGTNacgtRYKM*-", dtype=np.uint8)[rng.integers(0, 15, n)].copy()
        mostly = rng.random(n) < 0.9
        a[mostly] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, int(mostly.sum()))]
        cw, iw = C.c_uint64(), C.c_uint64()
        L.ipcr_gpu_packed_words(n, C.byref(cw), C.byref(iw))
        assert cw.value == 2 * iw.value and iw.value * 32 >= n + 32
        codes = np.full(cw.value, 0xDEADBEEF, dtype=np.uint32)
        inval = np.full(iw.value, 0xDEADBEEF, dtype=np.uint32)
        buf = a.tobytes()
        for threads in (1, 3):
            assert L.ipcr_gpu_pack_host(buf, n, codes.ctypes.data, inval.ctypes.data, threads) == 0
            padded = np.zeros(iw.value * 32, dtype=np.uint8)
            padded[:n] = a
            want_c = ((padded >> 1) & 3).astype(np.uint64).reshape(-1, 16)
            want_codes = (want_c << (2 * np.arange(16, dtype=np.uint64))).sum(axis=1).astype(np.uint32)
            bad = ~np.isin(padded, np.frombuffer(b"ACGT", dtype=np.uint8))
            want_inval = (bad.reshape(-1, 32).astype(np.uint64) << np.arange(32, dtype=np.uint64)).sum(axis=1).astype(np.uint32)
            assert np.array_equal(codes, want_codes) and np.array_equal(inval, want_inval), n


def test_feeder_fast_paths_and_fused_pack(tmp_path):
    """A 9 MB file (several threads): long clean stretches (the SSE2 counting path), CRLF and padded lines, '>'
    inside lines, blank lines, lowercase, N runs; records against the restatement of core/fasta, and -- the
    packing needs pinned memory, i.e. a CUDA driver -- the fused packed output against ipcr_gpu_pack_host of
    the concatenated records."""
    import ctypes as C

    import numpy as np

    from ipcr_b200 import _capi

    rng = np.random.default_rng(11)
    parts, want = [], []
    for r in range(7):
        n = int(rng.integers(200_000, 2_500_000))
        seq = np.frombuffer(b"ACGTacgtNRY", dtype=np.uint8)[rng.choice(11, n, p=[.22, .22, .22, .22, .02, .02, .02, .02, .02, .01, .01])]
        seq = seq.tobytes()
        width = [60, 80, 70, 61, 80, 1000, 80][r]
        lines = [seq[i:i + width] for i in range(0, n, width)]
        for k in range(0, len(lines), 997):  # sprinkle the odd cases
            lines[k] = [b"  " + lines[k] + b" \t", lines[k] + b"\r", b"\t" + lines[k], lines[k][:5] + b">" + lines[k][5:]][k % 4]
        body = b"\n".join(lines)
        if r == 2:
            body = body.replace(b"\n", b"\r\n")
        parts.append(b">rec%d some description\n" % r + body + (b"\n\n" if r % 2 else b"\n"))
    data = b"".join(parts)
    assert len(data) > 8 << 20
    path = tmp_path / "big.fa"
    path.write_bytes(data)
    ref = ref_reader(data)
    L = _capi.lib()
    for threads in (1, 3, 8):
        ff = FastaFile(str(path), threads=threads)
        assert ff.records() == ref
        ff.close()
    cat = b"".join(s_ for _id, s_ in ref)
    cw, iw = C.c_uint64(), C.c_uint64()
    L.ipcr_gpu_packed_words(len(cat), C.byref(cw), C.byref(iw))
    wc = np.zeros(cw.value, dtype=np.uint32)
    wi = np.zeros(iw.value, dtype=np.uint32)
    assert L.ipcr_gpu_pack_host(cat, len(cat), wc.ctypes.data, wi.ctypes.data, 2) == 0
    for threads in (1, 5, 8):
        for only in (False, True):
            ff = FastaFile(str(path), threads=threads, pack=not only, pack_only=only)
            pc, pi, ng = C.POINTER(C.c_uint32)(), C.POINTER(C.c_uint32)(), C.c_uint64()
            assert L.ipcr_gpu_fasta_packed(ff._h, C.byref(pc), C.byref(pi), C.byref(ng)) == 0
            assert ng.value == iw.value
            got_c = np.ctypeslib.as_array(pc, shape=(2 * ng.value,)).copy()
            got_i = np.ctypeslib.as_array(pi, shape=(ng.value,)).copy()
            assert np.array_equal(got_c, wc) and np.array_equal(got_i, wi), (threads, only)
            ff.close()


def _packed_arrays(ff):
    """(codes, bitmap) of a packed handle; also checks the handle's run list (what ipcr_gpu_scan_fasta ships instead of
    the bitmap -- the one-trip feeder derives it from the blocks that wrote a non-zero bitmap word) against the runs
    of the bitmap itself."""
    import ctypes as C

    import numpy as np

    from ipcr_b200 import _capi

    L = _capi.lib()
    pc, pi, ng = C.POINTER(C.c_uint32)(), C.POINTER(C.c_uint32)(), C.c_uint64()
    assert L.ipcr_gpu_fasta_packed(ff._h, C.byref(pc), C.byref(pi), C.byref(ng)) == 0
    codes = np.ctypeslib.as_array(pc, shape=(2 * ng.value,)).copy()
    inval = np.ctypeslib.as_array(pi, shape=(ng.value,)).copy()
    pr, nr = C.POINTER(C.c_uint64)(), C.c_uint64()
    if L.ipcr_gpu_fasta_runs(ff._h, C.byref(pr), C.byref(nr)) == 0:
        got = [int(pr[i]) for i in range(2 * nr.value)]
        cap = max(16, 2 * nr.value + 16)
        buf = (C.c_uint64 * (2 * cap))()
        n = L.ipcr_gpu_packed_runs(inval.ctypes.data_as(C.POINTER(C.c_uint32)), ff.total_bases, buf, cap, 1)
        assert n == nr.value and got == [int(buf[i]) for i in range(2 * n)], (got[:8], n)
    return codes, inval


def test_feeder_pack_only_equals_pack(tmp_path, monkeypatch):
    """IPCR_FASTA_PACK_ONLY: the feeder never writes the ASCII records -- every thread normalises into a staging
    buffer and packs from there, the 32-base groups shared by two threads' output ranges (and the ragged last
    group) are assembled in edge slots.  Codes and bitmap must equal the host packer's on the reference reader's
    records, for any number of threads, tiny ranges (several thread boundaries inside ONE group, threads that only
    see headers), ragged totals, lines longer than the staging buffer; the records read back are the reader's with
    every non-ACGT byte as 'N'."""
    import ctypes as C

    import numpy as np

    from ipcr_b200 import _capi

    L = _capi.lib()
    monkeypatch.setenv("IPCR_FEEDER_MT_MIN", "1")
    rng = np.random.default_rng(5)
    alphabet = np.frombuffer(b"ACGTacgtNRYn-", dtype=np.uint8)

    def body(n, width):
        seq = alphabet[rng.integers(0, len(alphabet), n)].tobytes()
        return b"\n".join(seq[i:i + width] for i in range(0, n, width))

    def n_run_genome(n, runs, width):
        seq = bytearray(np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n)].tobytes())
        for _ in range(runs):
            a, k = int(rng.integers(0, n - 1)), int(rng.integers(1, 300))
            k = min(k, n - a)
            seq[a:a + k] = b"N" * k
        seq = bytes(seq)
        return b"\n".join(seq[i:i + width] for i in range(0, n, width))

    files = {
        "tiny": b">a\nACGT\n>b x\nAC\n\n>c\n" + b"G" * 70 + b"\n",
        "headers_only_middle": b">a\n" + body(300, 60) + b"\n" + b"".join(b">e%d\n" % i for i in range(40)) + b">z\n" + body(97, 13) + b"\n",
        "dropped_ids": b"ACGT\n> \nACGTACGT\n>k\n" + body(1000, 80) + b"\n>\nTTTT\n>m\n" + body(33, 80) + b"\n",
        "one_long_line": b">chr\n" + body(200_000, 200_000) + b"\n>t\nacgtn\n",
        "crlf_padded": (b">r1 d\n" + body(5000, 61) + b"\n>r2\n  " + body(777, 50) + b" \n").replace(b"\n", b"\r\n"),
        "aligned_total": b">p\n" + body(4096, 64) + b"\n",
        # a clean genome with a few N runs (1 .. 300 bases, many across block boundaries): the run list then comes from
        # the few blocks that wrote a non-zero bitmap word, and runs that continue into the next block must be merged
        "n_runs": b">g\n" + n_run_genome(200_000, 25, 80) + b"\n>h\n" + n_run_genome(30_000, 6, 61) + b"\n",
        "empty": b"",
        "no_records": b"ACGT\nACGT\n",
    }
    for name, data in files.items():
        path = tmp_path / (name + ".fa")
        path.write_bytes(data)
        ref = ref_reader(data)
        cat = b"".join(s_ for _id, s_ in ref)
        cw, iw = C.c_uint64(), C.c_uint64()
        L.ipcr_gpu_packed_words(len(cat), C.byref(cw), C.byref(iw))
        wc = np.zeros(cw.value, dtype=np.uint32)
        wi = np.zeros(iw.value, dtype=np.uint32)
        assert L.ipcr_gpu_pack_host(cat, len(cat), wc.ctypes.data, wi.ctypes.data, 1) == 0
        as_n = [(i, bytes(c if c in b"ACGT" else ord("N") for c in s_)) for i, s_ in ref]
        # pack-only runs in ONE trip by default (blocks claimed in order, offsets through a chain of published
        # prefixes); tiny blocks put many block boundaries inside one group.  IPCR_FEEDER_TWO_PASS: the older form
        for mode in ("two_pass", "1048576", "37", "1000", "1"):
            if mode == "two_pass":
                monkeypatch.setenv("IPCR_FEEDER_TWO_PASS", "1")
            else:
                monkeypatch.delenv("IPCR_FEEDER_TWO_PASS", raising=False)
                monkeypatch.setenv("IPCR_FEEDER_BLOCK", mode)
            for threads in (1, 2, 3, 7, 16, 31):
                only = FastaFile(str(path), threads=threads, pack_only=True)
                assert only.packed and len(only) == len(ref) and only.total_bases == len(cat), (name, mode, threads)
                assert only.ids() == [i for i, _ in ref]
                got_c, got_i = _packed_arrays(only)
                assert np.array_equal(got_c, wc) and np.array_equal(got_i, wi), (name, mode, threads)
                assert only.records() == as_n, (name, mode, threads)
                only.close()
        monkeypatch.delenv("IPCR_FEEDER_BLOCK", raising=False)
        for threads in (1, 3, 16):
            full = FastaFile(str(path), threads=threads, pack=True)
            got_c, got_i = _packed_arrays(full)
            assert np.array_equal(got_c, wc) and np.array_equal(got_i, wi), (name, threads)
            assert full.records() == ref
            full.close()


def test_feeder_pack_only_fuzz(tmp_path, monkeypatch):
    """Byte soup against the restatement of the reference reader: random mixes of sequence runs, newlines, CRLF,
    blanks and tabs at line ends and inside lines, '>' at line starts and inside lines, empty ids, no final newline --
    through the one-trip pack-only parse (32-byte clean-chunk fast path, line rules for everything else) with
    block sizes from 1 byte up, and through the two-pass form."""
    import ctypes as C

    import numpy as np

    from ipcr_b200 import _capi

    L = _capi.lib()
    monkeypatch.setenv("IPCR_FEEDER_MT_MIN", "1")
    rng = random.Random(2024)
    tokens = [b"\n", b"\n", b"\r\n", b" ", b"\t", b">", b">id ", b"> \n", b">x\ty\n", b"\n\n", b"\x0b", b"N", b"acgt", b"-"]
    for it in range(120):
        parts = []
        for _ in range(rng.randint(1, 60)):
            if rng.random() < 0.6:
                parts.append(bytes(rng.choice(b"ACGTacgtNn") for _ in range(rng.choice([1, 5, 31, 32, 33, 64, 80, 200]))))
                parts.append(b"\n" if rng.random() < 0.8 else b"")
            else:
                parts.append(rng.choice(tokens))
        data = b"".join(parts)
        if rng.random() < 0.5:
            data = b">first\n" + data
        path = tmp_path / f"fuzz{it}.fa"
        path.write_bytes(data)
        ref = ref_reader(data)
        cat = b"".join(s_ for _id, s_ in ref)
        cw, iw = C.c_uint64(), C.c_uint64()
        L.ipcr_gpu_packed_words(len(cat), C.byref(cw), C.byref(iw))
        wc = np.zeros(cw.value, dtype=np.uint32)
        wi = np.zeros(iw.value, dtype=np.uint32)
        assert L.ipcr_gpu_pack_host(cat, len(cat), wc.ctypes.data, wi.ctypes.data, 1) == 0
        for mode in ("two_pass", "1048576", str(rng.randint(1, 200)), "64"):
            if mode == "two_pass":
                monkeypatch.setenv("IPCR_FEEDER_TWO_PASS", "1")
            else:
                monkeypatch.delenv("IPCR_FEEDER_TWO_PASS", raising=False)
                monkeypatch.setenv("IPCR_FEEDER_BLOCK", mode)
            only = FastaFile(str(path), threads=rng.choice([1, 2, 5]), pack_only=True)
            assert [(i, n) for i, n in zip(only.ids(), [len(s_) for _i, s_ in ref])] == [(i, len(s_)) for i, s_ in ref], (it, mode)
            assert only.total_bases == len(cat), (it, mode, data)
            got_c, got_i = _packed_arrays(only)
            assert np.array_equal(got_c, wc) and np.array_equal(got_i, wi), (it, mode, data)
            only.close()
        plain = FastaFile(str(path), threads=2)
        assert plain.records() == ref, (it, data)
        plain.close()


def test_invalid_run_list_round_trips():
    """ipcr_gpu_packed_runs: the hard-mismatch bitmap as [start, end) runs (what crosses PCIe instead of the bitmap);
    expanding the runs gives the bitmap back, padding included; too many runs are refused."""
    import ctypes as C

    import numpy as np

    from ipcr_b200 import _capi

    L = _capi.lib()
    rng = np.random.default_rng(3)
    for n in (0, 5, 4096, 70001, 3_000_000):
        a = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n)].copy()
        for _ in range(min(40, n // 50)):  # N runs of mixed lengths, some adjacent to word boundaries
            at = int(rng.integers(0, n))
            a[at:at + int(rng.choice([1, 2, 31, 32, 33, 500, 5000]))] = ord("N")
        cw, iw = C.c_uint64(), C.c_uint64()
        L.ipcr_gpu_packed_words(n, C.byref(cw), C.byref(iw))
        codes = np.zeros(cw.value, dtype=np.uint32)
        inval = np.zeros(iw.value, dtype=np.uint32)
        assert L.ipcr_gpu_pack_host(a.tobytes(), n, codes.ctypes.data, inval.ctypes.data, 2) == 0
        runs = (C.c_uint64 * 400)()
        for threads in (1, 4):
            k = L.ipcr_gpu_packed_runs(inval.ctypes.data, n, runs, 199, threads)
            assert k >= 1  # at least the padding behind the last base
            bits = np.zeros(iw.value * 32, dtype=np.uint8)
            prev_end = -1
            for i in range(k):
                lo, hi = runs[2 * i], runs[2 * i + 1]
                assert prev_end < lo < hi <= iw.value * 32  # disjoint, ascending, not touching
                bits[lo:hi] = 1
                prev_end = hi
            back = (bits.reshape(-1, 32).astype(np.uint64) << np.arange(32, dtype=np.uint64)).sum(axis=1).astype(np.uint32)
            assert np.array_equal(back, inval), (n, threads)
        if n >= 70001:
            assert L.ipcr_gpu_packed_runs(inval.ctypes.data, n, runs, 3, 1) == -1
