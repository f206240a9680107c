"""Thin wrappers over the resident-genome and split-phase entry points of the C ABI
(include/ipcr_gpu.h): pack a genome once, scan it with many panels, or find sites
on per-GPU shards and pair them once.  Marshalling only."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

from . import _capi
from .engine import CompiledPanel, Engine, Product, SimulationScratch, _as_buffer


class Genome:
    """A packed genome resident in HBM: the raw bytes plus whichever packed format the panel's filter
    reads (2-bit codes + hard-mismatch bitmap for the seed index, striped one-hot bit-planes for the
    bit-parallel scan); the other format is packed on demand."""

    def __init__(self, handle, n_recs: int, keep=None):
        self._h = handle
        self.n_recs = n_recs
        self._keep = keep

    @staticmethod
    def _segments(segments: Sequence[Tuple]):
        """segments: (buffer, rec_idx[, rec_off[, own_len]]) ; buffer is bytes-like or a
        tensor exposing data_ptr()/numel()."""
        arr = (_capi.SegmentC * max(1, len(segments)))()
        keep = []
        n_recs = 0
        for i, s in enumerate(segments):
            addr, n, k = _as_buffer(s[0])
            keep.append(k)
            rec = int(s[1])
            rec_off = int(s[2]) if len(s) > 2 else 0
            own = int(s[3]) if len(s) > 3 else n
            arr[i] = _capi.SegmentC(addr, n, own, rec_off, rec, 0)
            n_recs = max(n_recs, rec + 1)
        return arr, keep, n_recs

    @classmethod
    def upload(cls, scratch: SimulationScratch, segments: Sequence[Tuple], n_recs: Optional[int] = None):
        arr, keep, nr = cls._segments(segments)
        err = _capi.errbuf()
        h = _capi.lib().ipcr_gpu_genome_upload(scratch._h, arr, len(segments), n_recs or nr, err, len(err))
        if not h:
            raise _capi.GpuError(err.value.decode("utf-8", "replace"))
        return cls(h, n_recs or nr)

    @classmethod
    def adopt(cls, scratch: SimulationScratch, segments: Sequence[Tuple], n_recs: Optional[int] = None):
        """Segment buffers are DEVICE memory (e.g. torch cuda uint8 tensors, views of one
        allocation in ascending order); they must outlive the genome."""
        arr, keep, nr = cls._segments(segments)
        err = _capi.errbuf()
        h = _capi.lib().ipcr_gpu_genome_adopt(scratch._h, arr, len(segments), n_recs or nr, err, len(err))
        if not h:
            raise _capi.GpuError(err.value.decode("utf-8", "replace"))
        return cls(h, n_recs or nr, keep)

    @property
    def bases(self) -> int:
        return int(_capi.lib().ipcr_gpu_genome_bases(self._h))

    def records(self):
        arr = (_capi.RecordInfoC * max(1, self.n_recs))()
        _capi.lib().ipcr_gpu_genome_records(self._h, arr, self.n_recs)
        return [(int(arr[i].len), bool(arr[i].has_reset_byte)) for i in range(self.n_recs)]

    def planes(self):
        """(W, n_tiles, numpy uint32 array [tile, row, lane, plane]) -- debug tap."""
        import numpy as np

        L = _capi.lib()
        nw, W, nt = C.c_uint64(), C.c_uint32(), C.c_uint64()
        L.ipcr_gpu_debug_planes(self._h, None, 0, C.byref(nw), C.byref(W), C.byref(nt))
        buf = np.zeros(max(1, nw.value), dtype=np.uint32)
        rc = L.ipcr_gpu_debug_planes(self._h, buf.ctypes.data_as(C.POINTER(C.c_uint32)), nw.value,
                                     C.byref(nw), C.byref(W), C.byref(nt))
        if rc:
            raise _capi.GpuError("debug_planes failed")
        return W.value, nt.value, buf[: nw.value].reshape(nt.value, W.value + 31, 32, 4)

    def codes(self):
        """(codes uint32[2 * n_groups], inval uint32[n_groups]) of the seed-index format -- debug tap."""
        import numpy as np

        L = _capi.lib()
        ng = C.c_uint64()
        L.ipcr_gpu_debug_codes(self._h, None, None, 0, C.byref(ng))
        codes = np.zeros(max(1, 2 * ng.value), dtype=np.uint32)
        inval = np.zeros(max(1, ng.value), dtype=np.uint32)
        if L.ipcr_gpu_debug_codes(self._h, codes.ctypes.data_as(C.POINTER(C.c_uint32)),
                                  inval.ctypes.data_as(C.POINTER(C.c_uint32)), ng.value, C.byref(ng)):
            raise _capi.GpuError("debug_codes failed")
        return codes[: 2 * ng.value], inval[: ng.value]

    def close(self):
        if self._h is not None:
            _capi.lib().ipcr_gpu_genome_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _products(out, n, cp: CompiledPanel, seq_ids: Optional[Sequence[str]] = None) -> List[Product]:
    res = []
    for i in range(n):
        pc = out[i]
        sid = seq_ids[pc.rec_idx] if seq_ids else str(pc.rec_idx)
        res.append(Engine._product(pc, sid, cp, None))
    return res


def scan_genome(cp: CompiledPanel, scratch: SimulationScratch, genome: Genome,
                seq_ids: Optional[Sequence[str]] = None, raw: bool = False):
    """ForEachCompiledProduct over every record of a resident genome (record-major)."""
    L = _capi.lib()
    out = C.POINTER(_capi.ProductC)()
    n = C.c_uint64()
    err = _capi.errbuf()
    _capi.check(L.ipcr_gpu_scan_genome(cp._h, scratch._h, genome._h, C.byref(out), C.byref(n), err, len(err)), err)
    if raw:
        return out, n.value
    return _products(out, n.value, cp, seq_ids)


def scan_batch(cp: CompiledPanel, scratch: SimulationScratch, seqs: Sequence, seq_ids=None) -> List[Product]:
    L = _capi.lib()
    arr = (_capi.SeqC * max(1, len(seqs)))()
    keep = []
    for i, s in enumerate(seqs):
        addr, n, k = _as_buffer(s)
        keep.append(k)
        arr[i] = _capi.SeqC(addr, n)
    out = C.POINTER(_capi.ProductC)()
    n = C.c_uint64()
    err = _capi.errbuf()
    _capi.check(L.ipcr_gpu_scan_batch(cp._h, scratch._h, arr, len(seqs), C.byref(out), C.byref(n), err, len(err)), err)
    return _products(out, n.value, cp, seq_ids)


def scan_device(cp: CompiledPanel, scratch: SimulationScratch, segments: Sequence[Tuple], n_recs=None,
                seq_ids=None, raw: bool = False):
    """pack + scan of records that already live in device memory (one timed chain)."""
    arr, keep, nr = Genome._segments(segments)
    L = _capi.lib()
    out = C.POINTER(_capi.ProductC)()
    n = C.c_uint64()
    err = _capi.errbuf()
    _capi.check(L.ipcr_gpu_scan_device(cp._h, scratch._h, arr, len(segments), n_recs or nr, C.byref(out),
                                       C.byref(n), err, len(err)), err)
    if raw:
        return out, n.value
    return _products(out, n.value, cp, seq_ids)


def find_sites(cp: CompiledPanel, scratch: SimulationScratch, genome: Genome):
    """Stage 1 on one device.  Returns a numpy structured array of ipcr_site records."""
    import numpy as np

    L = _capi.lib()
    out = C.POINTER(_capi.SiteC)()
    n = C.c_uint64()
    err = _capi.errbuf()
    _capi.check(L.ipcr_gpu_find_sites(cp._h, scratch._h, genome._h, C.byref(out), C.byref(n), err, len(err)), err)
    if n.value == 0:
        return np.zeros(0, dtype=SITE_DTYPE)
    raw = C.string_at(out, n.value * C.sizeof(_capi.SiteC))
    return np.frombuffer(raw, dtype=SITE_DTYPE).copy()


def find_sites_device(cp: CompiledPanel, scratch: SimulationScratch, segments: Sequence[Tuple], n_recs: int,
                      force_raw: bool = True):
    """Stage 1 on device-resident shards (pack + filter + verify).  Returns (sites, records)
    where records[i] = (len seen by this shard, has_reset_byte)."""
    import numpy as np

    arr, keep, _nr = Genome._segments(segments)
    L = _capi.lib()
    out = C.POINTER(_capi.SiteC)()
    n = C.c_uint64()
    recs = (_capi.RecordInfoC * max(1, n_recs))()
    err = _capi.errbuf()
    _capi.check(L.ipcr_gpu_find_sites_device(cp._h, scratch._h, arr, len(segments), n_recs,
                                             1 if force_raw else 0, C.byref(out), C.byref(n), recs, err,
                                             len(err)), err)
    records = [(int(recs[i].len), bool(recs[i].has_reset_byte)) for i in range(n_recs)]
    if n.value == 0:
        return np.zeros(0, dtype=SITE_DTYPE), records
    raw = C.string_at(out, n.value * C.sizeof(_capi.SiteC))
    return np.frombuffer(raw, dtype=SITE_DTYPE).copy(), records


def pair_sites(cp: CompiledPanel, scratch: SimulationScratch, sites, records: Sequence[Tuple[int, bool]],
               seq_ids=None, raw: bool = False):
    """Stage 2: sites gathered from any number of shards + whole-record (len, has_reset_byte).
    raw=True returns (ctypes product pointer, count) without building Product objects."""
    import numpy as np

    L = _capi.lib()
    sites = np.ascontiguousarray(sites, dtype=SITE_DTYPE)
    recs = (_capi.RecordInfoC * max(1, len(records)))()
    for i, (ln, rs) in enumerate(records):
        recs[i].len = ln
        recs[i].has_reset_byte = 1 if rs else 0
    out = C.POINTER(_capi.ProductC)()
    n = C.c_uint64()
    err = _capi.errbuf()
    ptr = sites.ctypes.data_as(C.POINTER(_capi.SiteC))
    _capi.check(L.ipcr_gpu_pair_sites(cp._h, scratch._h, ptr, len(sites), recs, len(records), C.byref(out),
                                      C.byref(n), err, len(err)), err)
    if raw:
        return out, n.value
    return _products(out, n.value, cp, seq_ids)


def site_block_bytes(n_recs: int, site_cap: int) -> int:
    return int(_capi.lib().ipcr_gpu_site_block_bytes(n_recs, site_cap))


def scratch_site_cap(scratch: SimulationScratch) -> int:
    return int(_capi.lib().ipcr_gpu_scratch_site_cap(scratch._h))


def scratch_stream(scratch: SimulationScratch) -> int:
    """The scratch's cudaStream_t as an integer (torch.cuda.ExternalStream(...) orders NCCL against it)."""
    return int(_capi.lib().ipcr_gpu_scratch_stream(scratch._h) or 0)


def find_sites_block(cp: CompiledPanel, scratch: SimulationScratch, segments: Sequence[Tuple], n_recs: int,
                     block, site_cap: int, force_raw: bool = False) -> None:
    """Stage 1 of device-resident shards into `block` (device uint8 tensor of site_block_bytes);
    asynchronous on the scratch's stream: nothing is waited for."""
    arr, keep, _nr = Genome._segments(segments)
    err = _capi.errbuf()
    _capi.check(_capi.lib().ipcr_gpu_find_sites_block(cp._h, scratch._h, arr, len(segments), n_recs,
                                                      1 if force_raw else 0, _ptr(block), site_cap,
                                                      err, len(err)), err)


def _ptr(x) -> int:
    """device address of a tensor, or the integer address itself"""
    return int(x.data_ptr()) if hasattr(x, "data_ptr") else int(x)


class PeerExchange:
    """The peer-memory exchange of site blocks (include/ipcr_gpu.h, "peer exchange"): every rank exports its
    double-buffered block through CUDA IPC, the handles travel once through torch.distributed, and from then on
    stage 2 reads the peers' blocks straight over NVLink -- no collective on the data path."""

    def __init__(self, dist, scratch: SimulationScratch, n_recs: int, rank: int, world: int, device):
        """Collective: every rank must call it.  Raises on EVERY rank if any rank cannot export or map a
        block (each phase ends with an agreement), so a caller can fall back to a collective library."""
        import torch

        self.scratch, self.n_recs, self.rank, self.world = scratch, n_recs, rank, world
        self.site_cap = scratch_site_cap(scratch)
        L = _capi.lib()
        handle = C.create_string_buffer(64)
        err = _capi.errbuf()

        def agree(ok: bool, what: str):
            t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            if int(t[0]) == 0:
                raise _capi.GpuError(f"peer exchange unavailable: {what} failed on "
                                     + ("this rank: " + err.value.decode("utf-8", "replace") if not ok else "another rank"))

        agree(L.ipcr_gpu_peer_create(scratch._h, n_recs, self.site_cap, handle, err, len(err)) == 0, "exporting the block")
        mine = torch.frombuffer(bytearray(handle.raw), dtype=torch.uint8).to(device)
        allh = torch.empty(world * 64, dtype=torch.uint8, device=device)
        dist.all_gather_into_tensor(allh, mine)
        raw = bytes(allh.cpu().numpy().tobytes())
        agree(L.ipcr_gpu_peer_open(scratch._h, raw, world, rank, err, len(err)) == 0, "mapping the peers' blocks")
        # nobody starts a step before every rank has mapped every block (the agreement is a barrier)

    def begin_step(self) -> int:
        p = _capi.lib().ipcr_gpu_peer_begin_step(self.scratch._h)
        if not p:
            raise _capi.GpuError("peer exchange is not set up")
        return int(p)

    def publish(self) -> None:
        err = _capi.errbuf()
        _capi.check(_capi.lib().ipcr_gpu_peer_publish(self.scratch._h, err, len(err)), err)

    def pair(self, cp: CompiledPanel, rec_lens: Sequence[int]):
        """(ctypes product pointer, count), or None when the step has to run again on every rank."""
        L = _capi.lib()
        lens = (C.c_uint64 * max(1, len(rec_lens)))(*rec_lens)
        out = C.POINTER(_capi.ProductC)()
        n = C.c_uint64()
        err = _capi.errbuf()
        rc = L.ipcr_gpu_pair_peers(cp._h, self.scratch._h, lens, len(rec_lens), C.byref(out), C.byref(n), err, len(err))
        if rc == 2:
            return None
        _capi.check(rc, err)
        return out, n.value

    def close(self) -> None:
        _capi.lib().ipcr_gpu_peer_close(self.scratch._h)


def pack_host(buf, threads: int = 0):
    """ipcr_gpu_pack_host on a host buffer (bytes-like or a CPU uint8 tensor): returns (codes, inval) as
    pinned CPU int32 tensors when torch is importable, else numpy uint32 arrays."""
    L = _capi.lib()
    addr, n, keep = _as_buffer(buf)
    cw, iw = C.c_uint64(), C.c_uint64()
    L.ipcr_gpu_packed_words(n, C.byref(cw), C.byref(iw))
    try:
        import torch

        pin = torch.cuda.is_available()
        codes = torch.empty(cw.value, dtype=torch.int32, pin_memory=pin)
        inval = torch.empty(iw.value, dtype=torch.int32, pin_memory=pin)
        pc, pi = codes.data_ptr(), inval.data_ptr()
    except ImportError:  # pragma: no cover
        import numpy as np

        codes = np.empty(cw.value, dtype=np.uint32)
        inval = np.empty(iw.value, dtype=np.uint32)
        pc, pi = codes.ctypes.data, inval.ctypes.data
    if L.ipcr_gpu_pack_host(addr, n, pc, pi, threads):
        raise _capi.GpuError("ipcr_gpu_pack_host failed")
    return codes, inval


def packed_runs(inval, n_bases: int, max_runs: int = 1 << 20, threads: int = 0):
    """Run list of a host bitmap (ipcr_gpu_packed_runs): ctypes uint64 array of 2 * n entries and n, or None."""
    runs = (C.c_uint64 * (2 * max_runs + 2))()
    n = _capi.lib().ipcr_gpu_packed_runs(_ptr(inval), n_bases, runs, max_runs, threads)
    return (runs, int(n)) if n >= 0 else None


def inval_from_runs(scratch: SimulationScratch, d_inval, n_bases: int, runs, n_runs: int) -> None:
    """Rebuild the device bitmap from a run list, on the scratch's stream."""
    err = _capi.errbuf()
    _capi.check(_capi.lib().ipcr_gpu_inval_from_runs(scratch._h, _ptr(d_inval), n_bases, runs, n_runs, err, len(err)), err)


def find_sites_block_packed(cp: CompiledPanel, scratch: SimulationScratch, codes, inval, n_bases: int,
                            segments: Sequence[Tuple], n_recs: int, block, site_cap: int, force_raw: bool = False):
    """Stage 1 from a packed DEVICE buffer.  segments: (off, len, own_len, rec_off, rec_idx, has_reset_byte)."""
    arr = (_capi.PackedSegmentC * max(1, len(segments)))()
    for i, sg in enumerate(segments):
        arr[i] = _capi.PackedSegmentC(int(sg[0]), int(sg[1]), int(sg[2]), int(sg[3]), int(sg[4]), 1 if sg[5] else 0)
    err = _capi.errbuf()
    _capi.check(_capi.lib().ipcr_gpu_find_sites_block_packed(
        cp._h, scratch._h, int(codes.data_ptr()), int(inval.data_ptr()), n_bases, arr, len(segments), n_recs,
        1 if force_raw else 0, _ptr(block), site_cap, err, len(err)), err)


def pair_blocks(cp: CompiledPanel, scratch: SimulationScratch, blocks, n_blocks: int, site_cap: int,
                rec_lens: Sequence[int], seq_ids=None, raw: bool = False):
    """Stage 2 over gathered blocks (device uint8 tensor).  Returns None when stage 1 has to run
    again on every rank (capacity grown / raw plan variant needed), else the products."""
    L = _capi.lib()
    lens = (C.c_uint64 * max(1, len(rec_lens)))(*rec_lens)
    out = C.POINTER(_capi.ProductC)()
    n = C.c_uint64()
    err = _capi.errbuf()
    rc = L.ipcr_gpu_pair_blocks(cp._h, scratch._h, int(blocks.data_ptr()), n_blocks, site_cap, lens, len(rec_lens),
                                C.byref(out), C.byref(n), err, len(err))
    if rc == 2:
        return None
    _capi.check(rc, err)
    if raw:
        return out, n.value
    return _products(out, n.value, cp, seq_ids)


def sort_products(scratch: SimulationScratch, raw, n: int, rec_rank: Sequence[int], pair_rank: Sequence[int],
                  rec_off: Optional[Sequence[int]] = None):
    """`--sort` on the device: orders the ctypes product array `raw` (n entries, in place) like
    common.LessProduct given the string ranks; returns the tie flags (bytes)."""
    L = _capi.lib()
    rr = (C.c_uint32 * max(1, len(rec_rank)))(*rec_rank)
    pr = (C.c_uint32 * max(1, len(pair_rank)))(*pair_rank)
    ro = (C.c_int64 * max(1, len(rec_rank)))(*rec_off) if rec_off is not None else None
    tie = (C.c_uint8 * max(1, n))()
    err = _capi.errbuf()
    _capi.check(L.ipcr_gpu_sort_products(scratch._h, raw, n, rr, ro, len(rec_rank), pr, len(pair_rank), tie, err,
                                         len(err)), err)
    return bytes(tie[:n])


def synth_fill(scratch: SimulationScratch, tensor, first_pos: int, seed: int) -> None:
    """Fill a device uint8 tensor with the deterministic synthetic genome (bench input)."""
    err = _capi.errbuf()
    _capi.check(_capi.lib().ipcr_gpu_synth_fill(scratch._h, int(tensor.data_ptr()), int(tensor.numel()),
                                                first_pos, seed, err, len(err)), err)


def synth_bases_host(first_pos: int, n: int, seed: int) -> bytes:
    """Host restatement of k_synth_fill (csrc/kernels.cuh) for planting / checking inputs."""
    import numpy as np

    pos = np.arange(first_pos, first_pos + n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = np.uint64(seed) ^ ((pos >> np.uint64(5)) * np.uint64(0xD6E8FEB86659FD93))
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    two = ((x >> ((pos & np.uint64(31)) * np.uint64(2))) & np.uint64(3)).astype(np.uint8)
    return np.frombuffer(b"ACGT", dtype=np.uint8)[two].tobytes()


try:
    import numpy as _np

    SITE_DTYPE = _np.dtype([("pos", "<u8"), ("mm_mask", "<u8"), ("rec_idx", "<u4"), ("orient", "<u2"),
                            ("mm", "u1"), ("flags", "u1")])
    assert SITE_DTYPE.itemsize == 24
except ImportError:  # pragma: no cover
    SITE_DTYPE = None
