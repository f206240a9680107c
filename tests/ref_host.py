"""Test-side restatement of the reference's host flow AROUND the engine, with the CPU oracle
as the engine: reader -> (chunks) -> engine -> collector (global coords, dedupe) -> visitor ->
writer under --sort.  Independent of ipcr_b200.hostlib on purpose.  Citations are file:line
under /root/reference."""
from __future__ import annotations

from collections import OrderedDict

SPACE = b" \t\n\v\f\r"
COMP = dict(zip("ACGTRYSWKMBVDHN", "TGCAYRSWMKVBHDN"))


def read_fasta(data: bytes):
    """core/fasta/scan.go:10-69 + path_ctx.go:40-81 + normalize.go + stream.go:125-131"""
    recs, rid, seq = [], "", bytearray()
    pos, n = 0, len(data)
    while pos < n:
        nl = data.find(b"\n", pos)
        frag = data[pos:] if nl < 0 else data[pos:nl + 1]
        pos = n if nl < 0 else nl + 1
        if frag[:1] == b">":
            if rid != "":
                recs.append((rid, bytes(seq)))
            hdr = frag[1:].strip(SPACE)
            for i, c in enumerate(hdr):
                if c in b" \t":
                    hdr = hdr[:i]
                    break
            rid, seq = hdr.decode("latin-1"), bytearray()
        elif rid != "":
            seq += frag.strip(SPACE).upper()  # bytes.upper() maps a-z only: normalize.go:5-14
    if rid != "":
        recs.append((rid, bytes(seq)))
    return recs


def stream_chunks(rid, seq, chunk_size, overlap):
    """core/fasta/path_ctx.go:83-179 (the rolling window, fed the whole record at once)"""
    if overlap < 0:
        overlap = 0
    if chunk_size <= 0 or overlap >= chunk_size:
        return [(rid, seq)]
    step = chunk_size - overlap
    out, window, window_start, last_end, emitted = [], bytearray(seq), 0, 0, False
    while len(window) > chunk_size:
        out.append((f"{rid}:{window_start}-{window_start + chunk_size}", bytes(window[:chunk_size])))
        last_end, emitted = window_start + chunk_size, True
        window = window[step:] if step < len(window) else bytearray()
        window_start += step
    if not emitted:
        return [(rid, seq)]
    if last_end < len(seq):
        out.append((f"{rid}:{window_start}-{len(seq)}", bytes(window)))
    return out


def split_chunk_suffix(sid):
    """internal/common/ids.go:11-27"""
    colon = sid.rfind(":")
    if colon == -1 or colon == len(sid) - 1:
        return sid, 0, False
    suf = sid[colon + 1:]
    dash = suf.find("-")
    if dash == -1:
        return sid, 0, False
    num = suf[:dash]
    body = num[1:] if num[:1] in "+-" else num
    if not body or not body.isascii() or not body.isdigit():
        return sid, 0, False
    return sid[:colon], int(num), True


def revcomp(s: str) -> str:
    return "".join(COMP[c] for c in reversed(s))


def add_self_pairs(pairs):
    """internal/common/primers.go:11-38; pairs are (id, fwd, rev, min, max)"""
    out = list(pairs)
    for (pid, f, r, _mn, _mx) in pairs:
        if f:
            out.append((pid + "+A:self", f.upper(), f.upper(), 0, 0))
        if r:
            out.append((pid + "+B:self", r.upper(), r.upper(), 0, 0))
    return out


def add_self_pairs_unique(pairs):
    """internal/common/primers.go:41-74"""
    out, sa, sb = list(pairs), set(), set()
    for (pid, f, r, _mn, _mx) in pairs:
        fu, ru = f.strip().upper(), r.strip().upper()
        if fu and fu not in sa:
            sa.add(fu)
            out.append((pid + "+A:self", fu, fu, 0, 0))
        if ru and ru not in sb:
            sb.add(ru)
            out.append((pid + "+B:self", ru, ru, 0, 0))
    return out


class Prod(dict):
    __getattr__ = dict.__getitem__


def engine_products(O, cfgd, pairs, seq_id, seq):
    cfg = O.make_cfg(**cfgd)
    out = []
    for t in O.simulate_compiled(cfg, [(p[1], p[2], p[3], p[4]) for p in pairs], seq):
        out.append(Prod(experiment_id=pairs[t[0]][0], sequence_id=seq_id, start=t[2], end=t[3], length=t[4],
                        type=t[1], fwd_mm=t[5], rev_mm=t[6], fwd_idx=t[7], rev_idx=t[8], seq="", source_file=""))
    return out


def pipeline(O, files, pairs, cfgd, chunk_size, overlap, need_seq, dedup_cap=200000):
    """internal/pipeline/pipeline.go:33-194, single worker"""
    seen, res = OrderedDict(), []
    for fa in files:
        for rid, seq in read_fasta(open(fa, "rb").read()):
            for cid, cseq in stream_chunks(rid, seq, chunk_size, overlap):
                for p in engine_products(O, cfgd, pairs, cid, cseq):
                    if need_seq:
                        if cfgd.get("circular") and p.start > p.end:
                            p["seq"] = (cseq[p.start:] + cseq[:p.end]).decode("latin-1")
                        else:
                            p["seq"] = cseq[p.start:p.end].decode("latin-1")
                    p["source_file"] = fa
                    base, off, ok = split_chunk_suffix(p.sequence_id)
                    if not ok:
                        base, off = p.sequence_id, 0
                    key = (base, fa, p.start + off, p.end + off, p.type, p.experiment_id)
                    if key in seen:
                        seen.move_to_end(key)
                        continue
                    seen[key] = None
                    if len(seen) > dedup_cap:
                        seen.popitem(last=False)
                    if ok:
                        p["sequence_id"], p["start"], p["end"] = base, p.start + off, p.end + off
                    res.append(p)
    return res


def csv(a):
    return ",".join(str(v) for v in a)


def less_key(p):
    """internal/common/sort.go:34-74"""
    base, off, ok = split_chunk_suffix(p.sequence_id)
    if not ok:
        base, off = p.sequence_id, 0
    return (p.source_file, base, p.start + off, p.end + off, p.length, p.type, p.experiment_id, p.fwd_mm, p.rev_mm,
            csv(p.fwd_idx), csv(p.rev_idx), p.sequence_id, p.seq)


def base_row(p):
    """internal/output/rows.go:22-30"""
    return "\t".join([p.source_file, p.sequence_id, p.experiment_id, str(p.start), str(p.end), str(p.length), p.type,
                      str(p.fwd_mm), str(p.rev_mm), csv(p.fwd_idx), csv(p.rev_idx)])


def best_hit(O, amplicon: str, probe: str, max_mm: int):
    """core/oligo/oligo.go:19-77 -> (found, strand, pos, mm, site)"""
    amp = amplicon.upper()
    if probe.strip() == "":
        return (False, "", 0, 0, "")
    prb = "".join(ch.upper() for ch in probe if not ch.isspace() and ch not in "'\"")
    rc = revcomp(prb)
    if max_mm == 0 and all(c in "ACGT" for c in prb):
        i = amp.find(prb)
        if i >= 0:
            return (True, "+", i, 0, amp[i:i + len(prb)])
        i = amp.find(rc)
        if i >= 0:
            return (True, "-", i, 0, amp[i:i + len(rc)])
        return (False, "", 0, 0, "")
    best = None
    for strand, pat in (("+", prb), ("-", rc)):
        hits = O.find_matches(amp.encode("latin-1"), pat, max_mm, 0, 0)
        if not hits:
            continue
        loc = hits[0]
        for h in hits[1:]:
            if h[1] < loc[1] or (h[1] == loc[1] and h[0] < loc[0]):
                loc = h
        cand = (True, strand, loc[0], loc[1], amp[loc[0]:loc[0] + len(pat)])
        if best is None or cand[3] < best[3] or (cand[3] == best[3] and cand[2] < best[2]):
            best = cand
    return best or (False, "", 0, 0, "")


def nested_visit(O, p, inner_pairs, cfgd):
    """internal/visitors/nested.go:17-66 -> None or the chosen inner product"""
    hits = engine_products(O, cfgd, inner_pairs, "amplicon", p.seq.encode("latin-1"))
    if not hits:
        return None
    hits.sort(key=lambda h: (h.fwd_mm + h.rev_mm, -h.length, h.start, h.end, h.experiment_id))
    return hits[0]


HDR = "source_file\tsequence_id\texperiment_id\tstart\tend\tlength\ttype\tfwd_mm\trev_mm\tfwd_mm_i\trev_mm_i"
HDR_PROBE = HDR + "\tprobe_name\tprobe_seq\tprobe_found\tprobe_strand\tprobe_pos\tprobe_mm\tprobe_site"
HDR_NESTED = ("source_file\tsequence_id\touter_experiment_id\touter_start\touter_end\touter_length\touter_type\t"
              "inner_experiment_id\tinner_found\tinner_start\tinner_end\tinner_length\tinner_type\tinner_fwd_mm\tinner_rev_mm")


def _cfg(o):
    tw = 0 if o["terminal_window"] < 1 else o["terminal_window"]
    return dict(max_mm=o["mismatches"], terminal_window=tw, min_len=o["min_length"], max_len=o["max_length"],
                hit_cap=o["hit_cap"], seed_len=o["seed_length"], circular=o["circular"])


def _chunking(o, pairs):
    """internal/appcore/core.go:46-98 + internal/runutil/runutil.go:20-59"""
    eff, unbounded = o["max_length"], False
    for p in pairs:
        if p[4] > 0:
            eff = max(eff, p[4])
        elif o["max_length"] <= 0:
            unbounded = True
    if unbounded:
        eff = 0
    cs = o["chunk_size"]
    if cs <= 0 or o["circular"] or eff <= 0 or cs <= eff:
        return 0, 0
    return cs, eff


DEFAULTS = dict(mismatches=0, terminal_window=3, min_length=0, max_length=2000, hit_cap=10000, seed_length=12,
                circular=False, self_pairs=True, chunk_size=0, dedupe_cap=200000, sort=False, header=True)


def run(O, mode, files, opts, **kw):
    o = dict(DEFAULTS)
    o.update(opts)
    if mode in ("ipcr", "probe"):
        pairs = kw.get("pairs") or [("manual", kw["fwd"], kw["rev"], o["min_length"], o["max_length"])]
        if o["self_pairs"]:
            pairs = add_self_pairs(pairs)
    elif mode == "multiplex":
        pairs = add_self_pairs_unique(kw["pairs"]) if o["self_pairs"] else list(kw["pairs"])
    else:
        pairs = [("outer", kw["fwd"].upper(), kw["rev"].upper(), o["min_length"], o["max_length"])]
        inner = [("inner", kw["inner_fwd"].upper(), kw["inner_rev"].upper(), 0, 0)]
        if o["self_pairs"]:
            pairs, inner = add_self_pairs(pairs), add_self_pairs(inner)
    cfgd = _cfg(o)
    cs, ov = _chunking(o, pairs)
    need_seq = mode in ("probe", "nested")
    prods = pipeline(O, files, pairs, cfgd, cs, ov, need_seq, o["dedupe_cap"])
    rows = []
    if mode in ("ipcr", "multiplex"):
        items = [(p, base_row(p)) for p in prods]
        hdr = HDR
    elif mode == "probe":
        hdr, items = HDR_PROBE, []
        for p in prods:
            f, strand, pos, mm, site = best_hit(O, p.seq, kw["probe"].upper(), kw.get("probe_max_mm", 0))
            if kw.get("require_probe", True) and not f:
                continue
            items.append((p, "\t".join([base_row(p), kw.get("probe_name", "probe"), kw["probe"].upper(),
                                        "true" if f else "false", strand, str(pos) if f else "", str(mm) if f else "",
                                        site])))
    else:
        hdr, items = HDR_NESTED, []
        icfg = dict(max_mm=cfgd["max_mm"], terminal_window=cfgd["terminal_window"], seed_len=cfgd["seed_len"],
                    circular=False)
        for p in prods:
            h = nested_visit(O, p, inner, icfg)
            if h is None and kw.get("require_inner", False):
                continue
            f = h is not None
            items.append((p, "\t".join([p.source_file, p.sequence_id, p.experiment_id, str(p.start), str(p.end),
                                        str(p.length), p.type, h.experiment_id if f else "", "true" if f else "false",
                                        str(h.start) if f else "", str(h.end) if f else "", str(h.length) if f else "",
                                        h.type if f else "", str(h.fwd_mm) if f else "", str(h.rev_mm) if f else ""])))
    if o["sort"]:
        items.sort(key=lambda it: less_key(it[0]))
    lines = ([hdr] if o["header"] else []) + [r for _p, r in items]
    return "".join(l + "\n" for l in lines)
