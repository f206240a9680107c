"""Host-side mirror of the reference engine interface, backed by the CUDA library.

Names and call protocol follow /root/reference/core/engine:
    Config                      engine.go:10-19
    Pair                        core/primer/pair.go:4-10
    Product                     product.go:4-35 (printable fields)
    Engine.compile_panel        compiled.go:96-136   (CompilePanel)
    Engine.new_simulation_scratch  hit_collect.go:19-34
    Engine.simulate_batch       engine.go:49-51      (SimulateBatch)
    Engine.simulate_compiled[_with_scratch]  compiled.go:141-155
    Engine.for_each_compiled_product        compiled.go:162-267
so that it satisfies the same four pipeline interfaces
(internal/pipeline/sim.go:11-39).  CamelCase aliases are provided so parity
tests read like the reference's Go tests.  Everything below the marshalling
happens on the GPU through include/ipcr_gpu.h; nothing here computes matches.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

from . import _capi
from ._capi import GpuError

_COMP = {ord(a): ord(b) for a, b in zip("ACGTRYSWKMBVDHN", "TGCAYRSWMKVBHDN")}


def rev_comp(seq: bytes) -> bytes:
    """primer.RevComp (core/primer/rc.go:28-56): fails loudly on unknown bases."""
    out = bytearray(len(seq))
    n = len(seq)
    for i in range(n):
        b = seq[n - 1 - i]
        c = _COMP.get(b)
        if c is None:
            raise ValueError(
                f"invalid reverse-complement base {chr(b)!r} at position {n - i}; "
                "expected normalized uppercase IUPAC DNA")
        out[i] = c
    return bytes(out)


@dataclass
class Config:
    """engine.Config (core/engine/engine.go:10-19)."""
    max_mm: int = 0
    terminal_window: int = 0
    min_len: int = 0
    max_len: int = 0
    hit_cap: int = 0
    need_sites: bool = False
    seed_len: int = 0
    circular: bool = False


@dataclass
class Pair:
    """primer.Pair (core/primer/pair.go:4-10)."""
    id: str
    forward: str
    reverse: str
    min_product: int = 0
    max_product: int = 0


@dataclass
class Product:
    """engine.Product (core/engine/product.go:4-35)."""
    experiment_id: str = ""
    sequence_id: str = ""
    start: int = 0
    end: int = 0
    length: int = 0
    type: str = ""
    fwd_mm: int = 0
    rev_mm: int = 0
    fwd_mismatch_idx: Tuple[int, ...] = ()
    rev_mismatch_idx: Tuple[int, ...] = ()
    fwd_primer: str = ""
    rev_primer: str = ""
    fwd_site: str = ""
    rev_site: str = ""
    seq: str = ""
    source_file: str = ""

    def key(self):
        """The tuple the reference's oracle tests compare (approx_seed_oracle_test.go:12-42)."""
        return (self.sequence_id, self.experiment_id, self.start, self.end, self.length, self.type,
                self.fwd_mm, self.rev_mm, self.fwd_mismatch_idx, self.rev_mismatch_idx)


TYPE_NAMES = ("forward", "revcomp")


class CompiledPanel:
    """engine.CompiledPanel (compiled.go:77-91): immutable, shareable across workers."""

    def __init__(self, pairs: Sequence[Pair], cfg: Config):
        self.pairs: List[Pair] = [Pair(p.id, p.forward, p.reverse, p.min_product, p.max_product)
                                  for p in pairs]  # snapshot (compiled.go:98)
        self.cfg = Config(**vars(cfg))
        self._h = None
        if not self.pairs:
            return
        L = _capi.lib()
        c = _capi.Cfg(cfg.max_mm, cfg.terminal_window, cfg.min_len, cfg.max_len, cfg.hit_cap,
                      cfg.seed_len, 1 if cfg.circular else 0, 1 if cfg.need_sites else 0)
        arr = (_capi.PairC * len(self.pairs))()
        self._keep = []
        for i, p in enumerate(self.pairs):
            f, r = p.forward.encode("latin-1"), p.reverse.encode("latin-1")
            self._keep.append((f, r))
            arr[i] = _capi.PairC(f, len(f), r, len(r), p.min_product, p.max_product)
        err = _capi.errbuf()
        h = L.ipcr_gpu_panel_create(C.byref(c), arr, len(self.pairs), err, len(err))
        if not h:
            # the reference panics on invalid primers (rc.go:28-34); we raise
            raise ValueError(err.value.decode("utf-8", "replace"))
        self._h = h

    def have(self, pair_idx: int, which: str) -> bool:
        """compiledHas (compiled.go:35-37)."""
        if self._h is None:
            return False
        return bool(_capi.lib().ipcr_gpu_panel_have(self._h, pair_idx, ord(which)))

    @property
    def num_orientations(self) -> int:
        return int(_capi.lib().ipcr_gpu_panel_num_orientations(self._h)) if self._h else 0

    FILTER_BITPARALLEL_TABLE, FILTER_AUTO, FILTER_BITPARALLEL_JIT, FILTER_SEED_INDEX = 0, 1, 2, 3

    def set_filter(self, which: int) -> None:
        """Candidate filter: FILTER_AUTO (default: seed index with automatic fallback), FILTER_SEED_INDEX,
        FILTER_BITPARALLEL_JIT, FILTER_BITPARALLEL_TABLE.  All produce the same products."""
        self.set_jit(which)

    def set_jit(self, mode: int) -> None:
        if self._h:
            _capi.lib().ipcr_gpu_panel_set_jit(self._h, mode)

    def orientation(self, idx: int) -> Tuple[str, str]:
        buf = C.create_string_buffer(128)
        t = C.c_int()
        if _capi.lib().ipcr_gpu_debug_orientation(self._h, idx, buf, len(buf), C.byref(t)):
            raise IndexError(idx)
        return buf.value.decode(), chr(t.value)

    def plan_eval(self, orient: int, seq: bytes, start: int, raw_mode: bool = False) -> bool:
        return _capi.lib().ipcr_gpu_debug_plan_eval(self._h, orient, seq, len(seq), start,
                                                   1 if raw_mode else 0) == 1

    def seed_eval(self, seq: bytes, raw_mode: bool = False):
        """Host evaluation of the panel's seed index on `seq` (test tap, no GPU): returns
        ([(orientation, start)...], info) -- the candidates the seed filter hands to verify."""
        L = _capi.lib()
        n = C.c_uint64()
        info = (C.c_int32 * 8)()
        L.ipcr_gpu_debug_seed_eval(self._h, 1 if raw_mode else 0, seq, len(seq), None, 0, C.byref(n), info)
        buf = (C.c_uint64 * max(1, n.value))()
        if L.ipcr_gpu_debug_seed_eval(self._h, 1 if raw_mode else 0, seq, len(seq), buf, n.value, C.byref(n), info):
            raise RuntimeError("seed index evaluation failed")
        cands = [(int(buf[i]) >> 44, int(buf[i]) & ((1 << 44) - 1)) for i in range(n.value)]
        return cands, dict(S=info[0], entries=info[1], table_bits=info[2], bloom=bool(info[3]),
                           bloom_fill=info[4] / 1e6, bloom_k=info[5] & 0xFF, stride=info[5] >> 8, S1=info[6], keys1=info[7])

    def jit_source(self, raw_mode: bool = False) -> str:
        need = C.c_size_t()
        L = _capi.lib()
        L.ipcr_gpu_debug_jit_source(self._h, 1 if raw_mode else 0, None, 0, C.byref(need))
        buf = C.create_string_buffer(max(1, need.value))
        L.ipcr_gpu_debug_jit_source(self._h, 1 if raw_mode else 0, buf, len(buf), C.byref(need))
        return buf.value.decode()

    def jit_compile(self, raw_mode: bool = False) -> bytes:
        """NVRTC-compile the specialised kernel (no GPU needed); returns the sm_100a cubin."""
        L = _capi.lib()
        size = C.c_size_t()
        err = _capi.errbuf()
        buf = C.create_string_buffer(64 << 20)
        _capi.check(L.ipcr_gpu_debug_jit_compile(self._h, 1 if raw_mode else 0, buf, len(buf), C.byref(size),
                                                 err, len(err)), err)
        return buf.raw[: size.value]

    def close(self):
        if self._h is not None:
            _capi.lib().ipcr_gpu_panel_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SimulationScratch:
    """engine.SimulationScratch (hit_collect.go:11-34): one per worker; owns a CUDA
    stream plus device and pinned buffers."""

    def __init__(self, cp: CompiledPanel):
        self.panel = cp
        self._h = None
        if cp._h is None:
            return
        err = _capi.errbuf()
        h = _capi.lib().ipcr_gpu_scratch_create(cp._h, err, len(err))
        if not h:
            raise GpuError(err.value.decode("utf-8", "replace"))
        self._h = h

    def timing(self) -> dict:
        t = _capi.TimingC()
        if self._h is None or _capi.lib().ipcr_gpu_last_timing(self._h, C.byref(t)):
            return {}
        return {k: getattr(t, k) for k, _ in _capi.TimingC._fields_}

    def candidates(self) -> List[Tuple[int, int]]:
        n = C.c_uint64()
        L = _capi.lib()
        L.ipcr_gpu_debug_candidates(self._h, None, 0, C.byref(n))
        buf = (C.c_uint64 * max(1, n.value))()
        L.ipcr_gpu_debug_candidates(self._h, buf, n.value, C.byref(n))
        return [(int(buf[i]) >> 44, int(buf[i]) & ((1 << 44) - 1)) for i in range(n.value)]

    def close(self):
        if self._h is not None:
            _capi.lib().ipcr_gpu_scratch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _as_buffer(seq) -> Tuple[int, int, object]:
    """(address, length, keepalive) of a bytes-like or an object exposing data_ptr()/numel()."""
    if hasattr(seq, "data_ptr"):
        return int(seq.data_ptr()), int(seq.numel()), seq
    if isinstance(seq, bytes):
        return C.cast(C.c_char_p(seq), C.c_void_p).value or 0, len(seq), seq
    mv = memoryview(seq)
    if mv.readonly:
        b = bytes(mv)
        return C.cast(C.c_char_p(b), C.c_void_p).value or 0, len(b), b
    arr = (C.c_ubyte * len(mv)).from_buffer(mv)
    return C.addressof(arr), len(mv), arr


class Engine:
    """engine.Engine (engine.go:22-27) on the GPU."""

    def __init__(self, cfg: Optional[Config] = None):
        self.cfg = cfg or Config()

    # -- construction helpers ------------------------------------------------
    @staticmethod
    def New(cfg: Config) -> "Engine":  # engine.New (engine.go:27)
        return Engine(cfg)

    def set_hit_cap(self, n: int) -> None:  # engine.go:30
        self.cfg.hit_cap = n

    def compile_panel(self, pairs: Sequence[Pair]) -> CompiledPanel:
        return CompiledPanel(pairs, self.cfg)

    def new_simulation_scratch(self, cp: CompiledPanel) -> SimulationScratch:
        return SimulationScratch(cp)

    # -- the four simulator interfaces (internal/pipeline/sim.go:11-39) ------
    def simulate(self, seq_id: str, seq, pair: Pair) -> List[Product]:  # engine.go:33-36
        return self.simulate_batch(seq_id, seq, [pair])

    def simulate_batch(self, seq_id: str, seq, pairs: Sequence[Pair]) -> List[Product]:
        return self.simulate_compiled(seq_id, seq, self.compile_panel(pairs))

    def simulate_compiled(self, seq_id: str, seq, cp: CompiledPanel) -> List[Product]:
        return self.simulate_compiled_with_scratch(seq_id, seq, cp, None)

    def simulate_compiled_with_scratch(self, seq_id: str, seq, cp: CompiledPanel,
                                       scratch: Optional[SimulationScratch]) -> List[Product]:
        out: List[Product] = []

        def emit(p: Product):
            out.append(p)
            return None

        self.for_each_compiled_product(seq_id, seq, cp, scratch, emit)
        return out

    def for_each_compiled_product(self, seq_id: str, seq, cp: Optional[CompiledPanel],
                                  scratch: Optional[SimulationScratch],
                                  emit: Optional[Callable[[Product], Optional[Exception]]]):
        """ForEachCompiledProduct (compiled.go:162-267).  `emit` returns None to go
        on or an exception/True-ish error value to stop; that value is returned."""
        if cp is None or not cp.pairs or emit is None or cp._h is None:
            return None
        own = scratch is None
        if own:
            scratch = SimulationScratch(cp)
        try:
            addr, n, keep = _as_buffer(seq)
            L = _capi.lib()
            out = C.POINTER(_capi.ProductC)()
            n_out = C.c_uint64()
            err = _capi.errbuf()
            rc = L.ipcr_gpu_scan(cp._h, scratch._h, addr, n, C.byref(out), C.byref(n_out), err, len(err))
            _capi.check(rc, err)
            del keep
            need_sites = cp.cfg.need_sites
            seq_bytes = None
            if need_sites:
                seq_bytes = bytes(seq) if not isinstance(seq, bytes) else seq
            for i in range(n_out.value):
                e = emit(self._product(out[i], seq_id, cp, seq_bytes))
                if e is not None:
                    return e
            return None
        finally:
            if own:
                scratch.close()

    @staticmethod
    def _product(pc, seq_id: str, cp: CompiledPanel, seq_bytes: Optional[bytes]) -> Product:
        pair = cp.pairs[pc.pair_idx]
        fwd_is_a = pc.type == 0
        p = Product(
            experiment_id=pair.id, sequence_id=seq_id, start=pc.start, end=pc.end, length=pc.length,
            type=TYPE_NAMES[pc.type], fwd_mm=pc.fwd_mm, rev_mm=pc.rev_mm,
            fwd_mismatch_idx=_capi.mask_to_idx(pc.fwd_mm_mask),
            rev_mismatch_idx=_capi.mask_to_idx(pc.rev_mm_mask, descending=True),
            fwd_primer=pair.forward if fwd_is_a else pair.reverse,
            rev_primer=pair.reverse if fwd_is_a else pair.forward,
        )
        if seq_bytes is not None:  # NeedSites (engine.go:177-185, 242-250)
            flen = len(p.fwd_primer)
            rlen = len(p.rev_primer)
            if p.start + flen <= len(seq_bytes):
                p.fwd_site = seq_bytes[p.start:p.start + flen].decode("latin-1")
            b_start = p.end - rlen
            if p.end <= len(seq_bytes):
                p.rev_site = rev_comp(seq_bytes[b_start:p.end]).decode("latin-1")
        return p

    # Go-style aliases
    SetHitCap = set_hit_cap
    CompilePanel = compile_panel
    NewSimulationScratch = new_simulation_scratch
    Simulate = simulate
    SimulateBatch = simulate_batch
    SimulateCompiled = simulate_compiled
    SimulateCompiledWithScratch = simulate_compiled_with_scratch
    ForEachCompiledProduct = for_each_compiled_product


def New(cfg: Config) -> Engine:
    return Engine(cfg)
