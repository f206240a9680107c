"""Host-side glue around the scan, mirroring the reference packages that sit either side of
the hot path.  Kept deliberately small: policy, ordering and formatting only --
every match, annotation and inner scan is computed on the GPU.

    common   internal/common/{sort,ids,primers}.go, internal/runutil/{runutil,lru_set}.go
    fasta    core/fasta/path_ctx.go (rolling chunks over the native feeder's records)
    pipeline internal/pipeline/pipeline.go:33-194 (chunk ids, global coordinates, LRU dedupe)
    visitors internal/visitors/{pass,probe,nested}.go on the device
    output   internal/output/rows.go:22-30, probeoutput/text.go:11-28, nestedoutput/text.go:10-30
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from dataclasses import dataclass
from typing import Callable, List, Sequence, Tuple

from . import _capi
from .engine import Config, Engine, Pair, Product, SimulationScratch
from .fasta import FastaFile

# ---------------------------------------------------------------- common ----------------


def split_chunk_suffix(sid: str) -> Tuple[str, int, bool]:
    """common.SplitChunkSuffix (internal/common/ids.go:11-27)."""
    colon = sid.rfind(":")
    if colon == -1 or colon == len(sid) - 1:
        return sid, 0, False
    suffix = sid[colon + 1:]
    dash = suffix.find("-")
    if dash == -1:
        return sid, 0, False
    try:
        start = _atoi(suffix[:dash])
    except ValueError:
        return sid, 0, False
    return sid[:colon], start, True


def _atoi(s: str) -> int:
    """strconv.Atoi: optional sign followed by ASCII digits only."""
    body = s[1:] if s[:1] in "+-" else s
    if not body or not all("0" <= c <= "9" for c in body):
        raise ValueError(s)
    return int(s)


def ints_csv(a: Sequence[int]) -> str:
    return ",".join(str(v) for v in a)


def less_key(p: Product):
    """common.LessProduct (internal/common/sort.go:34-74) as a sort key."""
    base, off, ok = split_chunk_suffix(p.sequence_id)
    if not ok:
        base, off = p.sequence_id, 0
    return (p.source_file, base, p.start + off, p.end + off, p.length, p.type, p.experiment_id, p.fwd_mm,
            p.rev_mm, ints_csv(p.fwd_mismatch_idx), ints_csv(p.rev_mismatch_idx), p.sequence_id, p.seq)


def add_self_pairs(pairs: Sequence[Pair]) -> List[Pair]:
    """common.AddSelfPairs (internal/common/primers.go:11-38)."""
    out = list(pairs)
    for p in pairs:
        if p.forward != "":
            out.append(Pair(p.id + "+A:self", p.forward.upper(), p.forward.upper()))
        if p.reverse != "":
            out.append(Pair(p.id + "+B:self", p.reverse.upper(), p.reverse.upper()))
    return out


def add_self_pairs_unique(pairs: Sequence[Pair]) -> List[Pair]:
    """common.AddSelfPairsUnique (internal/common/primers.go:41-74)."""
    out = list(pairs)
    seen_a, seen_b = set(), set()
    for p in pairs:
        f = p.forward.strip().upper()
        if f and f not in seen_a:
            seen_a.add(f)
            out.append(Pair(p.id + "+A:self", f, f))
        r = p.reverse.strip().upper()
        if r and r not in seen_b:
            seen_b.add(r)
            out.append(Pair(p.id + "+B:self", r, r))
    return out


def effective_terminal_window(v: int) -> int:  # runutil.go:10-15
    return 0 if v < 1 else v


def validate_chunking(circular: bool, chunk_size: int, max_len: int, max_primer_len: int):
    """runutil.ValidateChunking (runutil.go:39-59) -> (chunk_size, overlap, warnings)."""
    if chunk_size <= 0:
        return 0, 0, []
    if circular:
        return 0, 0, ["chunking disabled for circular templates"]
    if max_len <= 0:
        return 0, 0, ["chunking disabled: a finite effective max product length is required to compute safe overlap"]
    if chunk_size <= max_len:
        return 0, 0, [f"chunk-size ({chunk_size}) <= effective max product length ({max_len}): disabling chunking"]
    return chunk_size, max_len, []


def effective_max_product_len(global_max: int, pairs: Sequence[Pair]) -> int:
    """appcore.effectiveMaxProductLen (internal/appcore/core.go:46-66)."""
    eff, unbounded = global_max, False
    for p in pairs:
        if p.max_product > 0:
            eff = max(eff, p.max_product)
            continue
        if global_max <= 0:
            unbounded = True
    return 0 if unbounded else eff


class LRUSet:
    """runutil.LRUSet (lru_set.go:8-38)."""

    def __init__(self, cap: int):
        self.cap = cap if cap > 0 else 200_000
        self.d: "OrderedDict" = OrderedDict()

    def add(self, k) -> bool:
        if k in self.d:
            self.d.move_to_end(k)
            return True
        self.d[k] = None
        if len(self.d) > self.cap:
            self.d.popitem(last=False)
        return False


# ---------------------------------------------------------------- fasta chunks ----------


def rolling_chunks(rid: str, seq: bytes, chunk_size: int, overlap: int):
    """core/fasta/path_ctx.go:19-179: whole record, or rolling overlapped chunks named
    "id:start-end"."""
    if overlap < 0:
        overlap = 0
    step = chunk_size - overlap
    if chunk_size <= 0 or overlap >= chunk_size or step <= 0 or len(seq) <= chunk_size:
        yield rid, seq
        return
    start, last_end = 0, 0
    while len(seq) - start > chunk_size:
        yield f"{rid}:{start}-{start + chunk_size}", seq[start:start + chunk_size]
        last_end = start + chunk_size
        start += step
    if last_end < len(seq):
        yield f"{rid}:{start}-{len(seq)}", seq[start:]


# ---------------------------------------------------------------- visitors ---------------


def normalize_oligo(s: str) -> str:
    """oligo.Normalize (core/oligo/validate.go:37-46)."""
    return "".join(ch.upper() for ch in s if not ch.isspace() and ch not in "'\"")


def validate_oligo(raw: str) -> str:
    s = normalize_oligo(raw)
    if s == "":
        raise ValueError("empty oligo")
    for i, ch in enumerate(s):
        if ch not in "ACGTRYSWKMBDHVN":
            raise ValueError(f"invalid base {ch!r} at {i + 1}; allowed: A C G T R Y S W K M B D H V N")
    return s


@dataclass
class AnnotatedProduct:
    """probeoutput.AnnotatedProduct (internal/probeoutput/types.go:10-22)."""
    product: Product
    probe_name: str = ""
    probe_seq: str = ""
    probe_found: bool = False
    probe_strand: str = ""
    probe_pos: int = 0
    probe_mm: int = 0
    probe_site: str = ""


@dataclass
class NestedProduct:
    """nestedoutput.NestedProduct (internal/nestedoutput/types.go:6-19)."""
    product: Product
    inner_found: bool = False
    inner_pair_id: str = ""
    inner_start: int = 0
    inner_end: int = 0
    inner_length: int = 0
    inner_type: str = ""
    inner_fwd_mm: int = 0
    inner_rev_mm: int = 0


class PassThrough:
    need_seq = False

    def visit_batch(self, scratch, raw, n, products):
        return list(products)


def _amplicons(scratch: SimulationScratch, raw, n: int) -> List[bytes]:
    L = _capi.lib()
    offs = (C.c_uint64 * (n + 1))()
    need = C.c_uint64()
    err = _capi.errbuf()
    _capi.check(L.ipcr_gpu_amplicons(scratch._h, raw, n, None, 0, offs, C.byref(need), err, len(err)), err)
    buf = C.create_string_buffer(max(1, need.value))
    _capi.check(L.ipcr_gpu_amplicons(scratch._h, raw, n, C.cast(buf, C.c_void_p), need.value, offs,
                                     C.byref(need), err, len(err)), err)
    return [buf.raw[offs[i]:offs[i + 1]] for i in range(n)]


class Probe:
    """visitors.Probe (internal/visitors/probe.go:10-32); BestHit runs on the GPU."""
    need_seq = True

    def __init__(self, name: str, seq: str, max_mm: int, require: bool):
        self.name, self.seq, self.max_mm, self.require = name, seq, max_mm, require

    def visit_batch(self, scratch, raw, n, products):
        out = []
        if n == 0:
            return out
        probe = "" if self.seq.strip() == "" else validate_oligo(self.seq)
        hits = (_capi.ProbeHitC * n)()
        err = _capi.errbuf()
        pb = probe.encode()
        _capi.check(_capi.lib().ipcr_gpu_probe_annotate(scratch._h, raw, n, pb, len(pb), self.max_mm, hits, err,
                                                        len(err)), err)
        for p, h in zip(products, hits):
            if self.require and not h.found:
                continue
            ap = AnnotatedProduct(p, self.name, self.seq.upper())
            if h.found:
                amp = p.seq.upper()
                ap.probe_found, ap.probe_strand = True, chr(h.strand)
                ap.probe_pos, ap.probe_mm = h.pos, h.mm
                ap.probe_site = amp[h.pos:h.pos + len(probe)]
            out.append(ap)
        return out


class Nested:
    """visitors.Nested (internal/visitors/nested.go:11-66): ONE batched inner scan per record
    instead of one engine per amplicon."""
    need_seq = True

    def __init__(self, inner_pairs: Sequence[Pair], engine_cfg: Config, require_inner: bool):
        self.inner_pairs = list(inner_pairs)
        self.cfg = engine_cfg
        self.require_inner = require_inner
        self._cp = Engine(engine_cfg).compile_panel(self.inner_pairs)
        self._sc = SimulationScratch(self._cp) if self.inner_pairs else None
        self._ids = (C.c_char_p * max(1, len(self.inner_pairs)))(*[p.id.encode() for p in self.inner_pairs])

    def visit_batch(self, scratch, raw, n, products):
        out = []
        if n == 0:
            return out
        best = (_capi.ProductC * n)()
        if self.inner_pairs:
            err = _capi.errbuf()
            _capi.check(_capi.lib().ipcr_gpu_nested(scratch._h, raw, n, self._cp._h, self._sc._h, self._ids, best,
                                                    err, len(err)), err)
        for i, p in enumerate(products):
            b = best[i]
            if not self.inner_pairs or b.length < 0:
                if self.require_inner:
                    continue
                out.append(NestedProduct(p))
                continue
            out.append(NestedProduct(p, True, self.inner_pairs[b.pair_idx].id, b.start, b.end, b.length,
                                     "forward" if b.type == 0 else "revcomp", b.fwd_mm, b.rev_mm))
        return out


# ---------------------------------------------------------------- pipeline ---------------


@dataclass
class PipelineConfig:
    """pipeline.Config (internal/pipeline/pipeline.go:15-22)."""
    chunk_size: int = 0
    overlap: int = 0
    circular: bool = False
    need_seq: bool = False
    dedup_cap: int = 0


def for_each_product(cfg: PipelineConfig, seq_files: Sequence[str], pairs: Sequence[Pair], eng: Engine,
                     visitor, emit: Callable[[object], None]) -> int:
    """pipeline.ForEachProduct (pipeline.go:33-194) with a single worker: compile the panel once,
    one scratch, scan per record/chunk, translate chunk-local coordinates, LRU-dedupe, visit."""
    cp = eng.compile_panel(pairs)
    if cp._h is None:
        return 0
    scratch = eng.new_simulation_scratch(cp)
    seen = LRUSet(cfg.dedup_cap)
    need_seq = cfg.need_seq or getattr(visitor, "need_seq", False)
    total = 0
    L = _capi.lib()
    for fa in seq_files:
        ff = FastaFile(fa)
        for ri in range(len(ff)):
            rid, seq = ff.record(ri)
            for cid, cseq in rolling_chunks(rid, seq, cfg.chunk_size, cfg.overlap):
                out = C.POINTER(_capi.ProductC)()
                n_out = C.c_uint64()
                err = _capi.errbuf()
                _capi.check(L.ipcr_gpu_scan(cp._h, scratch._h, cseq, len(cseq), C.byref(out), C.byref(n_out), err,
                                            len(err)), err)
                n = n_out.value
                if n == 0:
                    continue
                # the scratch's product buffer is reused by the visitors' own scans: copy first
                raw = (_capi.ProductC * n)()
                C.memmove(raw, out, n * C.sizeof(_capi.ProductC))
                amps = _amplicons(scratch, raw, n) if need_seq else None
                keep_idx, prods = [], []
                base, off, ok = split_chunk_suffix(cid)
                if not ok:
                    base, off = cid, 0
                for i in range(n):
                    p = Engine._product(raw[i], cid, cp, cseq if cp.cfg.need_sites else None)
                    if amps is not None:
                        p.seq = amps[i].decode("latin-1")
                    p.source_file = fa
                    key = (base, fa, p.start + off, p.end + off, p.type, p.experiment_id)
                    if seen.add(key):
                        continue
                    if ok:
                        p.sequence_id, p.start, p.end = base, p.start + off, p.end + off
                    keep_idx.append(i)
                    prods.append(p)
                if not prods:
                    continue
                if len(keep_idx) != n:
                    sub = (_capi.ProductC * len(keep_idx))()
                    for k, i in enumerate(keep_idx):
                        sub[k] = raw[i]
                    raw = sub
                for item in visitor.visit_batch(scratch, raw, len(prods), prods):
                    emit(item)
                    total += 1
        ff.close()
    return total


# ---------------------------------------------------------------- output -----------------

TSV_HEADER = ("source_file\tsequence_id\texperiment_id\tstart\tend\tlength\ttype\tfwd_mm\trev_mm\t"
              "fwd_mm_i\trev_mm_i")
TSV_HEADER_PROBE = TSV_HEADER + "\tprobe_name\tprobe_seq\tprobe_found\tprobe_strand\tprobe_pos\tprobe_mm\tprobe_site"
TSV_HEADER_NESTED = ("source_file\tsequence_id\touter_experiment_id\touter_start\touter_end\touter_length\t"
                     "outer_type\tinner_experiment_id\tinner_found\tinner_start\tinner_end\tinner_length\t"
                     "inner_type\tinner_fwd_mm\tinner_rev_mm")


def format_base_row(p: Product) -> str:  # output/rows.go:22-30
    return "\t".join([p.source_file, p.sequence_id, p.experiment_id, str(p.start), str(p.end), str(p.length),
                      p.type, str(p.fwd_mm), str(p.rev_mm), ints_csv(p.fwd_mismatch_idx),
                      ints_csv(p.rev_mismatch_idx)])


def format_probe_row(ap: AnnotatedProduct) -> str:  # probeoutput/text.go:11-28
    pos = str(ap.probe_pos) if ap.probe_found else ""
    mm = str(ap.probe_mm) if ap.probe_found else ""
    return "\t".join([format_base_row(ap.product), ap.probe_name, ap.probe_seq,
                      "true" if ap.probe_found else "false", ap.probe_strand, pos, mm, ap.probe_site])


def format_nested_row(np_: NestedProduct) -> str:  # nestedoutput/text.go:10-30
    p = np_.product
    f = np_.inner_found
    return "\t".join([p.source_file, p.sequence_id, p.experiment_id, str(p.start), str(p.end), str(p.length),
                      p.type, np_.inner_pair_id, "true" if f else "false",
                      str(np_.inner_start) if f else "", str(np_.inner_end) if f else "",
                      str(np_.inner_length) if f else "", np_.inner_type if f else "",
                      str(np_.inner_fwd_mm) if f else "", str(np_.inner_rev_mm) if f else ""])


# ---------------------------------------------------------------- apps -------------------


@dataclass
class Options:
    """the flags that reach the hot path (internal/clibase/common.go:61-110 defaults)."""
    mismatches: int = 0
    terminal_window: int = 3
    min_length: int = 0
    max_length: int = 2000
    hit_cap: int = 10000
    seed_length: int = 12
    circular: bool = False
    self_pairs: bool = True
    chunk_size: int = 0
    dedupe_cap: int = 200000
    sort: bool = False
    header: bool = True


def _run(o: Options, files, pairs, visitor, header: str, fmt: Callable, key: Callable) -> str:
    tw = effective_terminal_window(o.terminal_window)
    max_plen = max([len(p.forward) for p in pairs] + [len(p.reverse) for p in pairs] + [0])
    eff = effective_max_product_len(o.max_length, pairs)
    if eff > 0 and eff < max_plen:
        raise ValueError(f"effective maximum product length ({eff}) is smaller than the longest primer length ({max_plen})")
    chunk, overlap, _warn = validate_chunking(o.circular, o.chunk_size, eff, max_plen)
    eng = Engine(Config(max_mm=o.mismatches, terminal_window=tw, min_len=o.min_length, max_len=o.max_length,
                        hit_cap=o.hit_cap, seed_len=o.seed_length, circular=o.circular))
    items = []
    for_each_product(PipelineConfig(chunk, overlap, o.circular, getattr(visitor, "need_seq", False), o.dedupe_cap),
                     files, pairs, eng, visitor, items.append)
    if o.sort:
        items.sort(key=key)  # list.sort is stable like sort.SliceStable
    lines = [header] if o.header else []
    lines += [fmt(x) for x in items]
    return "".join(l + "\n" for l in lines)


def run_ipcr(o: Options, files, fwd=None, rev=None, pairs=None) -> str:
    """cmd/ipcr (internal/app/app.go:90-113), text output."""
    ps = list(pairs) if pairs is not None else [Pair("manual", fwd, rev, o.min_length, o.max_length)]
    if o.self_pairs:
        ps = add_self_pairs(ps)
    return _run(o, files, ps, PassThrough(), TSV_HEADER, format_base_row, less_key)


def run_multiplex(o: Options, files, pairs) -> str:
    """cmd/ipcr-multiplex with a TSV panel (internal/multiplexapp/app.go:195-272)."""
    ps = add_self_pairs_unique(pairs) if o.self_pairs else list(pairs)
    return _run(o, files, ps, PassThrough(), TSV_HEADER, format_base_row, less_key)


def run_probe(o: Options, files, fwd, rev, probe, probe_name="probe", probe_max_mm=0, require_probe=True) -> str:
    """cmd/ipcr-probe (internal/probeapp/app.go:89-116)."""
    ps = [Pair("manual", fwd, rev, o.min_length, o.max_length)]
    if o.self_pairs:
        ps = add_self_pairs(ps)
    v = Probe(probe_name, probe.upper(), probe_max_mm, require_probe)
    return _run(o, files, ps, v, TSV_HEADER_PROBE, format_probe_row, lambda ap: less_key(ap.product))


def run_nested(o: Options, files, fwd, rev, inner_fwd, inner_rev, require_inner=False) -> str:
    """cmd/ipcr-nested (internal/nestedapp/app.go:93-167)."""
    outer = [Pair("outer", fwd.upper(), rev.upper(), o.min_length, o.max_length)]
    inner = [Pair("inner", inner_fwd.upper(), inner_rev.upper())]
    if o.self_pairs:
        outer, inner = add_self_pairs(outer), add_self_pairs(inner)
    tw = effective_terminal_window(o.terminal_window)
    v = Nested(inner, Config(max_mm=o.mismatches, terminal_window=tw, seed_len=o.seed_length, circular=False),
               require_inner)
    return _run(o, files, outer, v, TSV_HEADER_NESTED, format_nested_row, lambda x: less_key(x.product))
