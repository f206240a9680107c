"""ctypes front-end for the CPU oracle (oracle/libipcr_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / ``--impl reference`` legs of bench.py.  The product package
(ipcr_b200/) never imports this module.

Product tuples are
``(pair_idx, type, start, end, length, fwd_mm, rev_mm, fwd_idx, rev_idx)``
with ``type`` in {"forward", "revcomp"} and the idx fields as tuples of ints,
i.e. the columns the reference prints (internal/output/rows.go:22-30).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Iterable, List, Sequence, Tuple

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libipcr_oracle.so")


class Cfg(C.Structure):
    _fields_ = [
        ("max_mm", C.c_int64),
        ("terminal_window", C.c_int64),
        ("min_len", C.c_int64),
        ("max_len", C.c_int64),
        ("hit_cap", C.c_int64),
        ("seed_len", C.c_int64),
        ("circular", C.c_int32),
        ("need_sites", C.c_int32),
    ]


class Pair(C.Structure):
    _fields_ = [
        ("fwd", C.c_char_p),
        ("fwd_len", C.c_int64),
        ("rev", C.c_char_p),
        ("rev_len", C.c_int64),
        ("min_product", C.c_int64),
        ("max_product", C.c_int64),
    ]


class Match(C.Structure):
    _fields_ = [
        ("pos", C.c_int64),
        ("mm", C.c_int32),
        ("len", C.c_int32),
        ("idx_off", C.c_int64),
        ("idx_n", C.c_int32),
        ("pad", C.c_int32),
    ]


class MatchList(C.Structure):
    _fields_ = [
        ("v", C.POINTER(Match)),
        ("n", C.c_size_t),
        ("cap", C.c_size_t),
        ("pool", C.POINTER(C.c_int32)),
        ("pn", C.c_size_t),
        ("pcap", C.c_size_t),
    ]


class Product(C.Structure):
    _fields_ = [
        ("pair_idx", C.c_int32),
        ("type", C.c_int32),
        ("start", C.c_int64),
        ("end", C.c_int64),
        ("length", C.c_int64),
        ("fwd_mm", C.c_int32),
        ("rev_mm", C.c_int32),
        ("fwd_idx_off", C.c_int64),
        ("rev_idx_off", C.c_int64),
        ("fwd_idx_n", C.c_int32),
        ("rev_idx_n", C.c_int32),
    ]


class Products(C.Structure):
    _fields_ = [
        ("v", C.POINTER(Product)),
        ("n", C.c_size_t),
        ("cap", C.c_size_t),
        ("pool", C.POINTER(C.c_int32)),
        ("pn", C.c_size_t),
        ("pcap", C.c_size_t),
    ]


_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (oracle/Makefile). Returns the .so path."""
    srcs = [os.path.join(_HERE, f) for f in ("ipcr_oracle.c", "ipcr_seeded.c", "ipcr_oracle.h")]
    stale = force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs if os.path.exists(s)
    )
    if stale:
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []))
    return _LIB_PATH


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        build()
    L = C.CDLL(_LIB_PATH)
    L.orc_iupac_mask.restype = C.c_uint8
    L.orc_iupac_mask.argtypes = [C.c_uint8]
    L.orc_base_match.restype = C.c_int
    L.orc_base_match.argtypes = [C.c_uint8, C.c_uint8]
    L.orc_mismatch_count.restype = C.c_int
    L.orc_mismatch_count.argtypes = [C.c_char_p, C.c_char_p, C.c_int64]
    L.orc_revcomp.restype = C.c_int64
    L.orc_revcomp.argtypes = [C.c_char_p, C.c_int64, C.c_char_p]
    L.orc_matchlist_init.argtypes = [C.POINTER(MatchList)]
    L.orc_matchlist_free.argtypes = [C.POINTER(MatchList)]
    L.orc_products_init.argtypes = [C.POINTER(Products)]
    L.orc_products_free.argtypes = [C.POINTER(Products)]
    L.orc_find_matches.argtypes = [
        C.c_char_p, C.c_int64, C.c_char_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
        C.POINTER(MatchList),
    ]
    L.orc_filter_left_tw.argtypes = [C.POINTER(MatchList), C.c_int64]
    L.orc_simulate_bruteforce.restype = C.c_int
    L.orc_simulate_bruteforce.argtypes = [
        C.POINTER(Cfg), C.POINTER(Pair), C.c_int64, C.c_char_p, C.c_int64, C.POINTER(Products),
    ]
    L.orc_compile_panel.restype = C.c_void_p
    L.orc_compile_panel.argtypes = [C.POINTER(Cfg), C.POINTER(Pair), C.c_int64]
    L.orc_panel_free.argtypes = [C.c_void_p]
    L.orc_panel_have.restype = C.c_int
    L.orc_panel_have.argtypes = [C.c_void_p, C.c_int64, C.c_int]
    L.orc_panel_num_patterns.restype = C.c_int64
    L.orc_panel_num_patterns.argtypes = [C.c_void_p]
    L.orc_panel_num_nodes.restype = C.c_int64
    L.orc_panel_num_nodes.argtypes = [C.c_void_p]
    L.orc_panel_seed_span.restype = C.c_int
    L.orc_panel_seed_span.argtypes = [
        C.c_void_p, C.c_int64, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
    ]
    L.orc_simulate_compiled.argtypes = [C.c_void_p, C.c_char_p, C.c_int64, C.POINTER(Products)]
    L.orc_scan_threaded.restype = C.c_int64
    L.orc_scan_threaded.argtypes = [
        C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_int64), C.c_int64, C.c_int64,
        C.c_int64, C.c_int,
    ]
    _lib = L
    return L


TYPE_NAMES = ("forward", "revcomp")
ProductTuple = Tuple[int, str, int, int, int, int, int, Tuple[int, ...], Tuple[int, ...]]


def make_cfg(max_mm=0, terminal_window=0, min_len=0, max_len=0, hit_cap=0, seed_len=0,
             circular=False, need_sites=False) -> Cfg:
    return Cfg(max_mm, terminal_window, min_len, max_len, hit_cap, seed_len,
               1 if circular else 0, 1 if need_sites else 0)


def _b(s) -> bytes:
    return s if isinstance(s, (bytes, bytearray)) else s.encode("latin-1")


def make_pairs(pairs: Sequence[Sequence]) -> Tuple[C.Array, list]:
    """pairs: iterable of (fwd, rev[, min_product, max_product])."""
    arr = (Pair * max(1, len(pairs)))()
    keep = []
    for i, p in enumerate(pairs):
        f, r = _b(p[0]), _b(p[1])
        keep.append((f, r))
        arr[i] = Pair(f, len(f), r, len(r), int(p[2]) if len(p) > 2 else 0,
                      int(p[3]) if len(p) > 3 else 0)
    return arr, keep


def _products_to_list(ps: Products) -> List[ProductTuple]:
    out = []
    for i in range(ps.n):
        p = ps.v[i]
        fi = tuple(ps.pool[p.fwd_idx_off + k] for k in range(p.fwd_idx_n))
        ri = tuple(ps.pool[p.rev_idx_off + k] for k in range(p.rev_idx_n))
        out.append((p.pair_idx, TYPE_NAMES[p.type], p.start, p.end, p.length, p.fwd_mm,
                    p.rev_mm, fi, ri))
    return out


def base_match(g: str, p: str) -> bool:
    return bool(lib().orc_base_match(ord(g), ord(p)))


def iupac_mask(c: str) -> int:
    return int(lib().orc_iupac_mask(ord(c)))


def mismatch_count(g, p) -> int:
    g, p = _b(g), _b(p)
    if len(g) != len(p):
        raise ValueError("MismatchCount: length mismatch")
    return int(lib().orc_mismatch_count(g, p, len(g)))


def revcomp(s) -> bytes:
    s = _b(s)
    out = C.create_string_buffer(max(1, len(s)))
    bad = lib().orc_revcomp(s, len(s), out)
    if bad:
        raise ValueError(f"invalid reverse-complement base {chr(s[bad - 1])!r} at position {bad}")
    return out.raw[: len(s)]


def find_matches(seq, primer, max_mm, cap_hits, terminal_window, left_tw=0):
    """primer.FindMatches (+ optional filterLeftTW). Returns [(pos, mm, idx tuple)]."""
    seq, primer = _b(seq), _b(primer)
    ml = MatchList()
    L = lib()
    L.orc_matchlist_init(C.byref(ml))
    L.orc_find_matches(seq, len(seq), primer, len(primer), max_mm, cap_hits, terminal_window,
                       C.byref(ml))
    if left_tw:
        L.orc_filter_left_tw(C.byref(ml), left_tw)
    out = []
    for i in range(ml.n):
        m = ml.v[i]
        out.append((m.pos, m.mm, tuple(ml.pool[m.idx_off + k] for k in range(m.idx_n))))
    L.orc_matchlist_free(C.byref(ml))
    return out


def simulate_bruteforce(cfg: Cfg, pairs, seq) -> List[ProductTuple]:
    seq = _b(seq)
    arr, _keep = make_pairs(pairs)
    ps = Products()
    L = lib()
    L.orc_products_init(C.byref(ps))
    rc = L.orc_simulate_bruteforce(C.byref(cfg), arr, len(pairs), seq, len(seq), C.byref(ps))
    if rc != 0:
        L.orc_products_free(C.byref(ps))
        raise ValueError("invalid primer base for reverse-complement")
    out = _products_to_list(ps)
    L.orc_products_free(C.byref(ps))
    return out


class Panel:
    """engine.CompiledPanel of the seeded production path (compiled.go:77-136)."""

    def __init__(self, cfg: Cfg, pairs):
        self.cfg = cfg
        self.pairs = list(pairs)
        self._arr, self._keep = make_pairs(pairs)
        self._h = lib().orc_compile_panel(C.byref(cfg), self._arr, len(self.pairs))
        if not self._h:
            raise ValueError("invalid primer base for reverse-complement")

    def close(self):
        if self._h:
            lib().orc_panel_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def have(self, pair_idx: int, which: str) -> bool:
        return bool(lib().orc_panel_have(self._h, pair_idx, ord(which)))

    def seed_span(self, pair_idx: int, which: str):
        off, ln = C.c_int64(), C.c_int64()
        ok = lib().orc_panel_seed_span(self._h, pair_idx, ord(which), C.byref(off), C.byref(ln))
        return (off.value, ln.value) if ok else None

    @property
    def num_patterns(self) -> int:
        return int(lib().orc_panel_num_patterns(self._h))

    @property
    def num_nodes(self) -> int:
        return int(lib().orc_panel_num_nodes(self._h))

    def simulate(self, seq) -> List[ProductTuple]:
        seq = _b(seq)
        ps = Products()
        L = lib()
        L.orc_products_init(C.byref(ps))
        L.orc_simulate_compiled(self._h, seq, len(seq), C.byref(ps))
        out = _products_to_list(ps)
        L.orc_products_free(C.byref(ps))
        return out

    def scan_threaded(self, seqs: Iterable[bytes], chunk_size=0, overlap=0, threads=1) -> int:
        seqs = [_b(s) for s in seqs]
        n = len(seqs)
        ptrs = (C.c_char_p * max(1, n))(*seqs)
        lens = (C.c_int64 * max(1, n))(*[len(s) for s in seqs])
        return int(lib().orc_scan_threaded(self._h, ptrs, lens, n, chunk_size, overlap, threads))


def simulate_compiled(cfg: Cfg, pairs, seq) -> List[ProductTuple]:
    p = Panel(cfg, pairs)
    try:
        return p.simulate(seq)
    finally:
        p.close()
