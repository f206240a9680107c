"""Native FASTA feeder (include/ipcr_gpu.h, csrc/fasta_host.hpp): the replacement for the
reference's core/fasta reader in front of the scan.  Marshalling only."""
from __future__ import annotations

import ctypes as C
from typing import List, Tuple

from . import _capi
from .engine import CompiledPanel, Engine, Product, SimulationScratch


class FastaFile:
    """All records of a FASTA file, normalised like fasta.StreamChunksPathCtx with chunking off
    (core/fasta/path_ctx.go:40-81)."""

    PIN, PACK, PACK_ONLY = 1, 2, 4  # IPCR_FASTA_* flags

    def __init__(self, path: str, threads: int = 0, pin: bool = False, pack: bool = False, pack_only: bool = False):
        """pack_only: keep ONLY the packed form (the feeder never writes the ASCII records; `records()` then
        rebuilds them from the codes with every non-ACGT byte as 'N')."""
        err = _capi.errbuf()
        flags = (self.PIN if pin else 0) | (self.PACK if pack else 0) | (self.PACK_ONLY if pack_only else 0)
        h = _capi.lib().ipcr_gpu_fasta_open(path.encode(), threads, flags, err, len(err))
        if not h:
            raise OSError(err.value.decode("utf-8", "replace"))
        self._h = h
        self.path = path

    @classmethod
    def from_memory(cls, seqs, ids=None, threads: int = 0, pin: bool = True, pack: bool = True):
        """The feeder's output for records that already sit in host memory (bytes-like, or tensors
        exposing data_ptr()/numel()): copied, upper-cased, optionally packed."""
        from .engine import _as_buffer

        arr = (_capi.SeqC * max(1, len(seqs)))()
        keep = []
        for i, s_ in enumerate(seqs):
            addr, n, k = _as_buffer(s_)
            keep.append(k)
            arr[i] = _capi.SeqC(addr, n)
        idarr = None
        if ids is not None:
            idarr = (C.c_char_p * max(1, len(ids)))(*[x.encode() for x in ids])
        err = _capi.errbuf()
        flags = (cls.PIN if pin else 0) | (cls.PACK if pack else 0)
        h = _capi.lib().ipcr_gpu_fasta_from_memory(arr, idarr, len(seqs), threads, flags, err, len(err))
        if not h:
            raise OSError(err.value.decode("utf-8", "replace"))
        self = cls.__new__(cls)
        self._h = h
        self.path = "<memory>"
        return self

    @property
    def packed(self) -> bool:
        return bool(_capi.lib().ipcr_gpu_fasta_is_packed(self._h))

    def scan_raw(self, cp: CompiledPanel, scratch: SimulationScratch):
        """ipcr_gpu_scan_fasta without building Product objects: (ctypes product pointer, count)."""
        out = C.POINTER(_capi.ProductC)()
        n = C.c_uint64()
        err = _capi.errbuf()
        _capi.check(_capi.lib().ipcr_gpu_scan_fasta(cp._h, scratch._h, self._h, C.byref(out), C.byref(n), err,
                                                    len(err)), err)
        return out, n.value

    def __len__(self) -> int:
        return int(_capi.lib().ipcr_gpu_fasta_num_records(self._h))

    @property
    def total_bases(self) -> int:
        return int(_capi.lib().ipcr_gpu_fasta_total_bases(self._h))

    def record(self, i: int) -> Tuple[str, bytes]:
        rid, seq, n = C.c_char_p(), C.c_void_p(), C.c_uint64()
        if _capi.lib().ipcr_gpu_fasta_record(self._h, i, C.byref(rid), C.byref(seq), C.byref(n)):
            raise IndexError(i)
        return rid.value.decode("latin-1"), C.string_at(seq.value, n.value) if n.value else b""

    def records(self) -> List[Tuple[str, bytes]]:
        return [self.record(i) for i in range(len(self))]

    def ids(self) -> List[str]:
        L = _capi.lib()
        out = []
        rid = C.c_char_p()
        for i in range(len(self)):
            L.ipcr_gpu_fasta_record(self._h, i, C.byref(rid), None, None)
            out.append(rid.value.decode("latin-1"))
        return out

    def scan(self, cp: CompiledPanel, scratch: SimulationScratch, source_file: str = "") -> List[Product]:
        """ForEachCompiledProduct over every record (record-major, reference emission order)."""
        out = C.POINTER(_capi.ProductC)()
        n = C.c_uint64()
        err = _capi.errbuf()
        _capi.check(_capi.lib().ipcr_gpu_scan_fasta(cp._h, scratch._h, self._h, C.byref(out), C.byref(n), err,
                                                    len(err)), err)
        ids = self.ids()
        res = []
        for i in range(n.value):
            p = Engine._product(out[i], ids[out[i].rec_idx], cp, None)
            p.source_file = source_file or self.path
            res.append(p)
        return res

    def close(self):
        if self._h is not None:
            _capi.lib().ipcr_gpu_fasta_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
