"""Multi-piece paths of the scan, all through the C ABI and all against the CPU oracle:
  * the device-resident split phase (ipcr_gpu_find_sites_block -> blocks -> ipcr_gpu_pair_blocks),
    which is what a multi-process multi-GPU run exchanges with ONE all-gather;
  * several devices inside one process (ipcr_gpu_init(0) + ipcr_gpu_scan_batch), skipped on a
    single-GPU box;
  * the reference's worker protocol: many host threads, one scratch each, one shared panel
    (/root/reference/internal/pipeline/pipeline.go:62-125,
    internal/pipeline/pipeline_engine_contract_test.go:213-325).
"""
import random
import threading

import pytest

import ipcr_test_helpers as H
from ipcr_b200 import Config, Engine, Pair, _capi
from ipcr_b200 import device as D
from ipcr_b200.fasta import FastaFile

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

A27F, B1492R = "AGAGTTTGATCMTGGCTCAG", "TACGGYTACCTTGTTAYGACTT"
PAIRS = [Pair("16S", A27F, B1492R), Pair("16S+A:self", A27F, A27F), Pair("16S+B:self", B1492R, B1492R)]
RAW = [(p.forward, p.reverse, 0, 0) for p in PAIRS]


def _tuple(p):
    return (p.experiment_id, p.type, p.start, p.end, p.length, p.fwd_mm, p.rev_mm,
            tuple(p.fwd_mismatch_idx), tuple(p.rev_mismatch_idx))


def _otuple(t):
    return (PAIRS[t[0]].id,) + tuple(t[1:])


def _records(rng, lens, with_n=True):
    recs = []
    for n in lens:
        g = H.random_genome(rng, n)
        for k in range(max(1, n // 20000)):
            H.plant_amplicon(rng, g, 200 + k * 18000 if n > 20000 else 10, A27F, B1492R, 300 + k, rng.randint(0, 2),
                             rng.randint(0, 2), 3, revcomp_strand=k % 2 == 1)
        recs.append(g)
    if with_n:
        recs[-1][len(recs[-1]) // 2] = ord("N")  # a reset byte only in the last piece
    return recs


def _plant_across_cuts(recs, lens, pieces, align=16):
    total = sum(lens)
    for c in [(total * k // pieces) & ~(align - 1) for k in range(1, pieces)]:
        off = 0
        for ri, n in enumerate(lens):
            if off + 30 <= c < off + n - 30:
                recs[ri][c - off - 9:c - off + 11] = A27F.replace("M", "C").encode()
            off += n


CFGS = (dict(max_mm=2, terminal_window=3, max_len=2000, hit_cap=0),
        dict(max_mm=2, terminal_window=3, max_len=2000, hit_cap=10000),
        dict(max_mm=1, terminal_window=0, max_len=0, hit_cap=3, circular=True))


def _oracle_rows(oracle, cfgd, seqs):
    want = []
    for ri, sq in enumerate(seqs):
        want += [(f"r{ri}",) + _otuple(t) for t in oracle.simulate_compiled(oracle.make_cfg(**cfgd), RAW, sq)]
    return want


def test_block_split_phase_equals_oracle(oracle):
    """4 shards on one device: every shard leaves a site block, stage 2 pairs over the blocks.  The
    reset byte sits in the last shard only, so with a hit cap the first round (plain plan variant)
    must be refused (return code 2) and the second round run raw."""
    rng = random.Random(5)
    lens = [90000, 500, 140000, 70000]
    recs = _records(rng, lens)
    shards = 4
    _plant_across_cuts(recs, lens, shards)
    seqs = [bytes(g) for g in recs]
    cat = b"".join(seqs)
    total = len(cat)
    dev = torch.frombuffer(bytearray(cat), dtype=torch.uint8).cuda()
    offs = [0]
    for n in lens:
        offs.append(offs[-1] + n)
    for cfgd in CFGS:
        eng = Engine(Config(**cfgd))
        cp = eng.compile_panel(PAIRS)
        sc = eng.new_simulation_scratch(cp)
        want = _oracle_rows(oracle, cfgd, seqs)
        rounds = 0
        while True:
            rounds += 1
            assert rounds <= 4
            cap = D.scratch_site_cap(sc)
            bb = D.site_block_bytes(len(lens), cap)
            allb = torch.zeros(shards * bb, dtype=torch.uint8, device="cuda")
            for k in range(shards):
                lo, hi = total * k // shards, total * (k + 1) // shards
                end = min(total, hi + 64)
                segs = []
                for c in range(len(lens)):
                    s0, s1, own1 = max(offs[c], lo), min(offs[c + 1], end), min(offs[c + 1], hi)
                    if s0 < s1 and own1 > s0:
                        segs.append((dev[s0:s1], c, s0 - offs[c], own1 - s0))
                D.find_sites_block(cp, sc, segs, len(lens), allb[k * bb:(k + 1) * bb], cap)
            got = D.pair_blocks(cp, sc, allb, shards, cap, lens, seq_ids=[f"r{i}" for i in range(len(lens))])
            if got is not None:
                break
        if cfgd["hit_cap"] > 0 and cfgd["terminal_window"] > 0:
            assert rounds == 2  # the raw-rc redo
        assert [(p.sequence_id,) + _tuple(p) for p in got] == want, cfgd


def test_multi_device_scan_batch_equals_single(oracle, monkeypatch):
    L = _capi.lib()
    if L.ipcr_gpu_device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    err = _capi.errbuf()
    rng = random.Random(9)
    lens = [90000, 500, 140000, 70000, 3]
    n_dev = L.ipcr_gpu_device_count()
    recs = _records(rng, lens)
    _plant_across_cuts(recs, lens, n_dev)
    seqs = [bytes(g) for g in recs]
    ids = [f"r{i}" for i in range(len(lens))]
    monkeypatch.setenv("IPCR_MULTI_MIN_BP", "1")
    try:
        for cfgd in CFGS:
            want = _oracle_rows(oracle, cfgd, seqs)
            eng = Engine(Config(**cfgd))
            cp = eng.compile_panel(PAIRS)
            sc = eng.new_simulation_scratch(cp)
            assert L.ipcr_gpu_init(1, err, len(err)) == 0, err.value
            single = [(p.sequence_id,) + _tuple(p) for p in D.scan_batch(cp, sc, seqs, seq_ids=ids)]
            assert single == want
            assert L.ipcr_gpu_init(0, err, len(err)) == 0, err.value
            assert L.ipcr_gpu_num_devices() == n_dev
            for _ in range(2):  # second call: warm peers, sticky raw variant
                multi = [(p.sequence_id,) + _tuple(p) for p in D.scan_batch(cp, sc, seqs, seq_ids=ids)]
                assert multi == want, cfgd
    finally:
        L.ipcr_gpu_init(1, err, len(err))


def test_concurrent_workers_share_one_panel(oracle):
    """8 host threads, own scratch each, one shared panel, different records; the first calls race
    the lazy seed-index build and the per-device table upload.  Results equal the serial ones."""
    rng = random.Random(21)
    n_workers, per_worker = 8, 3
    seqs = []
    for w in range(n_workers * per_worker):
        n = rng.choice([3000, 40000, 120000])
        g = H.random_genome(rng, n, n_frac=0.0005 if w % 2 else 0.0)
        for k in range(1 + n // 40000):
            H.plant_amplicon(rng, g, 100 + k * 30000, A27F, B1492R, 400 + w, rng.randint(0, 2), rng.randint(0, 2), 3,
                             revcomp_strand=(w + k) % 2 == 1)
        seqs.append(bytes(g))
    cfgd = dict(max_mm=2, terminal_window=3, max_len=2000, hit_cap=10000)
    want = [[_otuple(t) for t in oracle.simulate_compiled(oracle.make_cfg(**cfgd), RAW, sq)] for sq in seqs]
    for mode in (3, 2):  # seed index (asynchronous chain), NVRTC-specialised bit-parallel
        eng = Engine(Config(**cfgd))
        cp = eng.compile_panel(PAIRS)  # fresh panel: nothing built yet
        cp.set_jit(mode)
        got = [None] * len(seqs)
        errors = []
        gate = threading.Barrier(n_workers)

        def worker(w):
            try:
                sc = eng.new_simulation_scratch(cp)
                gate.wait()
                for i in range(w, len(seqs), n_workers):
                    got[i] = [_tuple(p) for p in eng.simulate_compiled_with_scratch(f"s{i}", seqs[i], cp, sc)]
            except Exception as ex:  # noqa: BLE001
                errors.append(repr(ex))

        th = [threading.Thread(target=worker, args=(w,)) for w in range(n_workers)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        assert not errors, errors
        assert got == want, mode


def test_packed_feeder_scan_equals_ascii_and_oracle(oracle, monkeypatch):
    """The feeder's packed output (2-bit codes + hard-mismatch bitmap, 0.375 B/bp over PCIe) through
    ipcr_gpu_scan_fasta: no pack kernel, verify on the codes, 1..5 streamed slabs -- same products, same
    order as the ASCII path and the oracle.  Records hold N runs and soft-masked (lowercase) stretches,
    which the reader upper-cases (core/fasta/normalize.go:5-14) before anything is packed."""
    rng = random.Random(13)
    lens = [90000, 500, 140000, 70000, 3]
    recs = _records(rng, lens)
    _plant_across_cuts(recs, lens, 4)
    recs[0][5000:5400] = bytes(recs[0][5000:5400]).lower()
    recs[2][100:160] = b"N" * 60
    seqs = [bytes(g) for g in recs]
    upper = [sq.upper() for sq in seqs]
    ids = [f"r{i}" for i in range(len(lens))]
    for cfgd in CFGS:
        want = _oracle_rows(oracle, cfgd, upper)
        eng = Engine(Config(**cfgd))
        cp = eng.compile_panel(PAIRS)
        sc = eng.new_simulation_scratch(cp)
        packed = FastaFile.from_memory(seqs, ids=ids, pack=True)
        plain = FastaFile.from_memory(seqs, ids=ids, pack=False)
        assert packed.packed and not plain.packed
        assert packed.records() == list(zip(ids, upper))
        for slabs in ("1", "3", "5"):
            monkeypatch.setenv("IPCR_STREAM_SLABS", slabs)
            got = [(p.sequence_id,) + _tuple(p) for p in packed.scan(cp, sc)]
            assert got == want, (cfgd, slabs)
        # the bitmap travels as a run list by default (rebuilt on the device); the same with the bitmap itself
        monkeypatch.setenv("IPCR_NO_RUNS", "1")
        assert [(p.sequence_id,) + _tuple(p) for p in packed.scan(cp, sc)] == want, cfgd
        # pageable packed records go through the multi-threaded pinned staging ring (inputs >= 64 Mbp by default)
        pageable = FastaFile.from_memory(seqs, ids=ids, pin=False, pack=True)
        monkeypatch.setenv("IPCR_STAGED_H2D_MIN", "1")
        for slabs in ("1", "4"):
            monkeypatch.setenv("IPCR_STREAM_SLABS", slabs)
            assert [(p.sequence_id,) + _tuple(p) for p in pageable.scan(cp, sc)] == want, (cfgd, "staged+bitmap", slabs)
        monkeypatch.delenv("IPCR_NO_RUNS")
        assert [(p.sequence_id,) + _tuple(p) for p in pageable.scan(cp, sc)] == want, (cfgd, "staged")
        monkeypatch.delenv("IPCR_STAGED_H2D_MIN")
        pageable.close()
        monkeypatch.delenv("IPCR_STREAM_SLABS")
        assert [(p.sequence_id,) + _tuple(p) for p in plain.scan(cp, sc)] == want
        packed.close()
        plain.close()


def test_multi_device_packed_scan_fasta_equals_oracle(oracle, monkeypatch):
    """ipcr_gpu_scan_fasta over the packed form after ipcr_gpu_init(0): every device copies its own
    4096-base-aligned slice of the codes (the bitmap from the clipped run list, or its slice), the site blocks
    meet on the calling scratch's device.  Same rows, same order as one device and the oracle; binding
    sites are planted across every slice boundary, N runs straddle one."""
    L = _capi.lib()
    if L.ipcr_gpu_device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    err = _capi.errbuf()
    rng = random.Random(21)
    lens = [90000, 500, 140000, 70000, 3]
    n_dev = L.ipcr_gpu_device_count()
    recs = _records(rng, lens)
    _plant_across_cuts(recs, lens, n_dev, align=4096)
    cut1 = (sum(lens) // n_dev) & ~4095  # an N run over the first boundary (inside record 0 or 2)
    off = 0
    for ri, n in enumerate(lens):
        if off + 200 <= cut1 < off + n - 200:
            recs[ri][cut1 - off + 40:cut1 - off + 90] = b"N" * 50
            recs[ri][cut1 - off - 120:cut1 - off - 60] = b"N" * 60
        off += n
    recs[0][5000:5400] = bytes(recs[0][5000:5400]).lower()
    seqs = [bytes(g) for g in recs]
    upper = [sq.upper() for sq in seqs]
    ids = [f"r{i}" for i in range(len(lens))]
    monkeypatch.setenv("IPCR_MULTI_MIN_BP", "1")
    try:
        for cfgd in CFGS:
            want = _oracle_rows(oracle, cfgd, upper)
            eng = Engine(Config(**cfgd))
            cp = eng.compile_panel(PAIRS)
            sc = eng.new_simulation_scratch(cp)
            packed = FastaFile.from_memory(seqs, ids=ids, pack=True)
            assert L.ipcr_gpu_init(0, err, len(err)) == 0, err.value
            for runs in (True, False, True):  # third call: warm peers, sticky raw variant
                if runs:
                    monkeypatch.delenv("IPCR_NO_RUNS", raising=False)
                else:
                    monkeypatch.setenv("IPCR_NO_RUNS", "1")
                got = [(p.sequence_id,) + _tuple(p) for p in packed.scan(cp, sc)]
                assert got == want, (cfgd, runs)
            assert L.ipcr_gpu_init(1, err, len(err)) == 0, err.value
            assert [(p.sequence_id,) + _tuple(p) for p in packed.scan(cp, sc)] == want
            packed.close()
    finally:
        monkeypatch.delenv("IPCR_NO_RUNS", raising=False)
        L.ipcr_gpu_init(1, err, len(err))


def test_block_split_phase_packed_equals_oracle(oracle):
    """Shards that receive the PACKED form (what a multi-process run copies to each GPU): stage 1 from
    device-resident codes + bitmap into site blocks, stage 2 over the blocks."""
    rng = random.Random(17)
    lens = [90000, 500, 140000, 70000]
    recs = _records(rng, lens)
    shards = 3
    _plant_across_cuts(recs, lens, shards)
    seqs = [bytes(g) for g in recs]
    cat = b"".join(seqs)
    total = len(cat)
    offs = [0]
    for n in lens:
        offs.append(offs[-1] + n)
    reset = [any(c not in b"ACGTacgt" for c in sq) for sq in seqs]
    for cfgd in CFGS:
        eng = Engine(Config(**cfgd))
        cp = eng.compile_panel(PAIRS)
        sc = eng.new_simulation_scratch(cp)
        want = _oracle_rows(oracle, cfgd, seqs)
        pieces = []
        for k in range(shards):
            lo, hi = (total * k // shards) & ~4095, ((total * (k + 1) // shards) & ~4095) if k + 1 < shards else total
            end = min(total, hi + 64)
            codes, inval = D.pack_host(cat[lo:end])
            segs = []
            for c in range(len(lens)):
                s0, s1, own1 = max(offs[c], lo), min(offs[c + 1], end), min(offs[c + 1], hi)
                if s0 < s1 and own1 > s0:
                    segs.append((s0 - lo, s1 - s0, own1 - s0, s0 - offs[c], c, reset[c]))
            d_inval = inval.cuda()
            got_runs = D.packed_runs(inval, end - lo, max_runs=4096)
            if got_runs is not None and k % 2 == 0:  # every other shard: the bitmap rebuilt on the device from its runs
                rebuilt = torch.full_like(d_inval, 0x5A5A5A5A)
                D.inval_from_runs(sc, rebuilt, end - lo, *got_runs)
                torch.cuda.synchronize()
                assert torch.equal(rebuilt.cpu(), inval)
                d_inval = rebuilt
            pieces.append((codes.cuda(), d_inval, end - lo, segs))
        while True:
            cap = D.scratch_site_cap(sc)
            bb = D.site_block_bytes(len(lens), cap)
            allb = torch.zeros(shards * bb, dtype=torch.uint8, device="cuda")
            for k, (codes, inval, n, segs) in enumerate(pieces):
                D.find_sites_block_packed(cp, sc, codes, inval, n, segs, len(lens), allb[k * bb:(k + 1) * bb], cap)
            got = D.pair_blocks(cp, sc, allb, shards, cap, lens, seq_ids=[f"r{i}" for i in range(len(lens))])
            if got is not None:
                break
        assert [(p.sequence_id,) + _tuple(p) for p in got] == want, cfgd


def test_device_sort_equals_less_product(oracle):
    """ipcr_gpu_sort_products against common.LessProduct (hostlib.less_key restates internal/common/sort.go:34-74)
    on a dense product list: several records whose IDs sort differently from their file order, chunk-style
    offsets, circular wrap products, two pairs sharing one experiment ID (ties the device reports)."""
    import ctypes as C

    from ipcr_b200 import hostlib as HL
    from ipcr_b200.engine import Engine as E

    rng = random.Random(33)
    # the two "alpha" pairs differ in two forward columns: a site one mismatch away from both ties on every
    # integer key and differs only in its mismatch list -- the tie-break the device leaves to the caller
    pairs = [Pair("zeta", "GATTAC", "TTGACC"), Pair("alpha", "ACGTTG", "CCATGG"), Pair("alpha", "ACCTTC", "CCATGG"),
             Pair("mid", "TTGACC", "GATTAC")]
    ids = ["chr10", "chr2", "chrX:100-900", "chr1", "b"]
    lens = [30000, 21000, 800, 25000, 12000]
    seqs = [bytes(H.random_genome(rng, n)) for n in lens]
    eng = Engine(Config(max_mm=1, terminal_window=0, max_len=600, hit_cap=0, circular=True))
    cp = eng.compile_panel(pairs)
    sc = eng.new_simulation_scratch(cp)
    arr = (_capi.SeqC * len(seqs))()
    keep = []
    for i, sq in enumerate(seqs):
        b = C.create_string_buffer(sq, len(sq))
        keep.append(b)
        arr[i] = _capi.SeqC(C.cast(b, C.c_void_p).value, len(sq))
    out = C.POINTER(_capi.ProductC)()
    n = C.c_uint64()
    err = _capi.errbuf()
    _capi.check(_capi.lib().ipcr_gpu_scan_batch(cp._h, sc._h, arr, len(seqs), C.byref(out), C.byref(n), err, len(err)), err)
    n = n.value
    assert n > 5000
    raw = (_capi.ProductC * n)()
    C.memmove(raw, out, n * C.sizeof(_capi.ProductC))
    prods = [E._product(raw[i], ids[raw[i].rec_idx], cp, None) for i in range(n)]
    for p in prods:
        p.source_file = "f.fa"
    want = sorted(prods, key=HL.less_key)
    # ranks of the strings, offsets of the chunk-style ID
    split = [HL.split_chunk_suffix(x) for x in ids]
    bases = [b if ok else x for (b, _o, ok), x in zip(split, ids)]
    offs = [o if ok else 0 for (_b, o, ok) in split]
    ub, up = sorted(set(bases)), sorted(set(p.id for p in pairs))
    tie = D.sort_products(sc, raw, n, [ub.index(b) for b in bases], [up.index(p.id) for p in pairs], offs)
    got = [E._product(raw[i], ids[raw[i].rec_idx], cp, None) for i in range(n)]
    for p in got:
        p.source_file = "f.fa"
    # settle the reported ties like the caller has to
    i = 0
    while i < n:
        j = i + 1
        while j < n and tie[j]:
            j += 1
        if j - i > 1:
            got[i:j] = sorted(got[i:j], key=HL.less_key)
        i = j
    key = lambda p: (p.sequence_id, p.experiment_id, p.type, p.start, p.end, p.length, p.fwd_mm, p.rev_mm,
                     tuple(p.fwd_mismatch_idx), tuple(p.rev_mismatch_idx))
    assert [key(p) for p in got] == [key(p) for p in want]
    assert any(tie)
