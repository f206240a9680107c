"""ctypes binding of the C ABI declared in include/ipcr_gpu.h.

This is the same surface a cgo shim binds.  Loading fails
loudly when the CUDA library has not been built; there is no Python or CPU
implementation of the scan behind it.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libipcr_b200.so")


class Cfg(C.Structure):
    _fields_ = [
        ("max_mm", C.c_int32),
        ("terminal_window", C.c_int32),
        ("min_len", C.c_int64),
        ("max_len", C.c_int64),
        ("hit_cap", C.c_int64),
        ("seed_len", C.c_int32),
        ("circular", C.c_uint8),
        ("need_sites", C.c_uint8),
        ("reserved", C.c_uint8 * 2),
    ]


class PairC(C.Structure):
    _fields_ = [
        ("fwd", C.c_char_p),
        ("fwd_len", C.c_uint32),
        ("rev", C.c_char_p),
        ("rev_len", C.c_uint32),
        ("min_product", C.c_int64),
        ("max_product", C.c_int64),
    ]


class SeqC(C.Structure):
    _fields_ = [("seq", C.c_void_p), ("len", C.c_uint64)]


class SegmentC(C.Structure):
    _fields_ = [
        ("seq", C.c_void_p),
        ("len", C.c_uint64),
        ("own_len", C.c_uint64),
        ("rec_off", C.c_uint64),
        ("rec_idx", C.c_uint32),
        ("reserved", C.c_uint32),
    ]


class PackedSegmentC(C.Structure):
    _fields_ = [
        ("off", C.c_uint64),
        ("len", C.c_uint64),
        ("own_len", C.c_uint64),
        ("rec_off", C.c_uint64),
        ("rec_idx", C.c_uint32),
        ("has_reset_byte", C.c_uint32),
    ]


class ProductC(C.Structure):
    _fields_ = [
        ("start", C.c_int64),
        ("end", C.c_int64),
        ("length", C.c_int64),
        ("fwd_mm_mask", C.c_uint64),
        ("rev_mm_mask", C.c_uint64),
        ("pair_idx", C.c_uint32),
        ("rec_idx", C.c_uint32),
        ("fwd_mm", C.c_uint16),
        ("rev_mm", C.c_uint16),
        ("type", C.c_uint8),
        ("reserved", C.c_uint8 * 11),
    ]


class SiteC(C.Structure):
    _fields_ = [
        ("pos", C.c_uint64),
        ("mm_mask", C.c_uint64),
        ("rec_idx", C.c_uint32),
        ("orient", C.c_uint16),
        ("mm", C.c_uint8),
        ("flags", C.c_uint8),
    ]


class RecordInfoC(C.Structure):
    _fields_ = [("len", C.c_uint64), ("has_reset_byte", C.c_uint8), ("reserved", C.c_uint8 * 7)]


class ProbeHitC(C.Structure):
    _fields_ = [("pos", C.c_int64), ("mm", C.c_uint16), ("found", C.c_uint8), ("strand", C.c_uint8),
                ("reserved", C.c_uint8 * 4)]


class TimingC(C.Structure):
    _fields_ = [
        ("h2d_ms", C.c_float), ("pack_ms", C.c_float), ("filter_ms", C.c_float),
        ("verify_ms", C.c_float), ("pair_ms", C.c_float), ("d2h_ms", C.c_float),
        ("total_ms", C.c_float),
        ("candidates", C.c_uint64), ("sites", C.c_uint64), ("products", C.c_uint64),
        ("kernel_launches", C.c_uint32), ("used_jit", C.c_uint32),
        ("plan_variant", C.c_uint32), ("reserved", C.c_uint32),
    ]


assert C.sizeof(ProductC) == 64 and C.sizeof(SiteC) == 24

# every entry point include/ipcr_gpu.h declares
SYMBOLS = [
    "ipcr_gpu_abi_version", "ipcr_gpu_device_count", "ipcr_gpu_init", "ipcr_gpu_set_device", "ipcr_gpu_num_devices",
    "ipcr_gpu_site_block_bytes", "ipcr_gpu_scratch_site_cap", "ipcr_gpu_scratch_stream",
    "ipcr_gpu_find_sites_block", "ipcr_gpu_pair_blocks", "ipcr_gpu_find_sites_block_packed",
    "ipcr_gpu_sort_products", "ipcr_gpu_fasta_from_memory",
    "ipcr_gpu_peer_create", "ipcr_gpu_peer_open", "ipcr_gpu_peer_begin_step", "ipcr_gpu_peer_publish",
    "ipcr_gpu_pair_peers", "ipcr_gpu_peer_close", "ipcr_gpu_fasta_is_packed", "ipcr_gpu_fasta_packed", "ipcr_gpu_packed_words", "ipcr_gpu_pack_host",
    "ipcr_gpu_packed_runs", "ipcr_gpu_inval_from_runs",
    "ipcr_gpu_panel_create", "ipcr_gpu_panel_destroy", "ipcr_gpu_panel_num_pairs",
    "ipcr_gpu_panel_num_orientations", "ipcr_gpu_panel_have", "ipcr_gpu_panel_set_jit", "ipcr_gpu_panel_set_filter",
    "ipcr_gpu_scratch_create", "ipcr_gpu_scratch_destroy",
    "ipcr_gpu_scan", "ipcr_gpu_scan_batch", "ipcr_gpu_scan_device",
    "ipcr_gpu_genome_upload", "ipcr_gpu_genome_adopt", "ipcr_gpu_genome_destroy",
    "ipcr_gpu_genome_bases", "ipcr_gpu_genome_records", "ipcr_gpu_scan_genome",
    "ipcr_gpu_find_sites", "ipcr_gpu_find_sites_device", "ipcr_gpu_pair_sites",
    "ipcr_gpu_fasta_open", "ipcr_gpu_fasta_close", "ipcr_gpu_fasta_num_records", "ipcr_gpu_fasta_runs",
    "ipcr_gpu_fasta_total_bases", "ipcr_gpu_fasta_record", "ipcr_gpu_scan_fasta",
    "ipcr_gpu_probe_annotate", "ipcr_gpu_amplicons", "ipcr_gpu_nested",
    "ipcr_gpu_product_mismatch_idx", "ipcr_gpu_last_timing", "ipcr_gpu_synth_fill",
    "ipcr_gpu_debug_planes", "ipcr_gpu_debug_candidates", "ipcr_gpu_debug_plan_eval",
    "ipcr_gpu_debug_orientation", "ipcr_gpu_debug_jit_source", "ipcr_gpu_debug_jit_compile",
    "ipcr_gpu_debug_seed_eval", "ipcr_gpu_debug_codes",
]

_lib = None


def lib():
    """Loads libipcr_b200.so (raises if it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m ipcr_b200.build` "
            "(the ipcr GPU engine has no Python/CPU fallback)")
    L = C.CDLL(LIB_PATH)
    P, V, U64, SZ = C.POINTER, C.c_void_p, C.c_uint64, C.c_size_t
    ERR = (C.c_char_p, SZ)

    def sig(name, res, *args):
        f = getattr(L, name)
        f.restype = res
        f.argtypes = list(args)

    sig("ipcr_gpu_abi_version", C.c_int)
    sig("ipcr_gpu_device_count", C.c_int)
    sig("ipcr_gpu_init", C.c_int, C.c_int, *ERR)
    sig("ipcr_gpu_set_device", C.c_int, C.c_int, *ERR)
    sig("ipcr_gpu_num_devices", C.c_int)
    sig("ipcr_gpu_site_block_bytes", U64, SZ, U64)
    sig("ipcr_gpu_scratch_site_cap", U64, V)
    sig("ipcr_gpu_scratch_stream", V, V)
    sig("ipcr_gpu_find_sites_block", C.c_int, V, V, P(SegmentC), SZ, SZ, C.c_int, V, U64, *ERR)
    sig("ipcr_gpu_pair_blocks", C.c_int, V, V, V, C.c_uint32, U64, P(U64), SZ, P(P(ProductC)), P(U64), *ERR)
    sig("ipcr_gpu_panel_create", V, P(Cfg), P(PairC), SZ, *ERR)
    sig("ipcr_gpu_panel_destroy", None, V)
    sig("ipcr_gpu_panel_num_pairs", SZ, V)
    sig("ipcr_gpu_panel_num_orientations", SZ, V)
    sig("ipcr_gpu_panel_have", C.c_int, V, SZ, C.c_int)
    sig("ipcr_gpu_panel_set_jit", None, V, C.c_int)
    sig("ipcr_gpu_panel_set_filter", None, V, C.c_int)
    sig("ipcr_gpu_scratch_create", V, V, *ERR)
    sig("ipcr_gpu_scratch_destroy", None, V)
    sig("ipcr_gpu_scan", C.c_int, V, V, V, U64, P(P(ProductC)), P(U64), *ERR)
    sig("ipcr_gpu_scan_batch", C.c_int, V, V, P(SeqC), SZ, P(P(ProductC)), P(U64), *ERR)
    sig("ipcr_gpu_scan_device", C.c_int, V, V, P(SegmentC), SZ, SZ, P(P(ProductC)), P(U64), *ERR)
    sig("ipcr_gpu_genome_upload", V, V, P(SegmentC), SZ, SZ, *ERR)
    sig("ipcr_gpu_genome_adopt", V, V, P(SegmentC), SZ, SZ, *ERR)
    sig("ipcr_gpu_genome_destroy", None, V)
    sig("ipcr_gpu_genome_bases", U64, V)
    sig("ipcr_gpu_genome_records", C.c_int, V, P(RecordInfoC), SZ)
    sig("ipcr_gpu_scan_genome", C.c_int, V, V, V, P(P(ProductC)), P(U64), *ERR)
    sig("ipcr_gpu_find_sites", C.c_int, V, V, V, P(P(SiteC)), P(U64), *ERR)
    sig("ipcr_gpu_find_sites_device", C.c_int, V, V, P(SegmentC), SZ, SZ, C.c_int, P(P(SiteC)), P(U64),
        P(RecordInfoC), *ERR)
    sig("ipcr_gpu_pair_sites", C.c_int, V, V, P(SiteC), U64, P(RecordInfoC), SZ, P(P(ProductC)), P(U64), *ERR)
    sig("ipcr_gpu_fasta_open", V, C.c_char_p, C.c_int, C.c_int, *ERR)
    sig("ipcr_gpu_fasta_runs", C.c_int, V, P(P(U64)), P(U64))
    sig("ipcr_gpu_fasta_from_memory", V, P(SeqC), P(C.c_char_p), SZ, C.c_int, C.c_int, *ERR)
    sig("ipcr_gpu_fasta_is_packed", C.c_int, V)
    sig("ipcr_gpu_fasta_packed", C.c_int, V, P(P(C.c_uint32)), P(P(C.c_uint32)), P(U64))
    sig("ipcr_gpu_packed_words", None, U64, P(U64), P(U64))
    sig("ipcr_gpu_pack_host", C.c_int, V, U64, V, V, C.c_int)
    sig("ipcr_gpu_packed_runs", C.c_int64, V, U64, P(U64), U64, C.c_int)
    sig("ipcr_gpu_inval_from_runs", C.c_int, V, V, U64, P(U64), U64, *ERR)
    sig("ipcr_gpu_find_sites_block_packed", C.c_int, V, V, V, V, U64, P(PackedSegmentC), SZ, SZ, C.c_int, V, U64, *ERR)
    sig("ipcr_gpu_peer_create", C.c_int, V, SZ, U64, V, *ERR)
    sig("ipcr_gpu_peer_open", C.c_int, V, V, C.c_uint32, C.c_uint32, *ERR)
    sig("ipcr_gpu_peer_begin_step", V, V)
    sig("ipcr_gpu_peer_publish", C.c_int, V, *ERR)
    sig("ipcr_gpu_pair_peers", C.c_int, V, V, P(U64), SZ, P(P(ProductC)), P(U64), *ERR)
    sig("ipcr_gpu_peer_close", None, V)
    sig("ipcr_gpu_sort_products", C.c_int, V, P(ProductC), U64, P(C.c_uint32), P(C.c_int64), SZ, P(C.c_uint32), SZ,
        P(C.c_uint8), *ERR)
    sig("ipcr_gpu_fasta_close", None, V)
    sig("ipcr_gpu_fasta_num_records", SZ, V)
    sig("ipcr_gpu_fasta_total_bases", U64, V)
    sig("ipcr_gpu_fasta_record", C.c_int, V, SZ, P(C.c_char_p), P(V), P(U64))
    sig("ipcr_gpu_scan_fasta", C.c_int, V, V, V, P(P(ProductC)), P(U64), *ERR)
    sig("ipcr_gpu_probe_annotate", C.c_int, V, P(ProductC), U64, C.c_char_p, C.c_uint32, C.c_int32, P(ProbeHitC), *ERR)
    sig("ipcr_gpu_amplicons", C.c_int, V, P(ProductC), U64, V, U64, P(U64), P(U64), *ERR)
    sig("ipcr_gpu_nested", C.c_int, V, P(ProductC), U64, V, V, P(C.c_char_p), P(ProductC), *ERR)
    sig("ipcr_gpu_product_mismatch_idx", C.c_int, P(ProductC), C.c_int, P(C.c_int32), C.c_int)
    sig("ipcr_gpu_last_timing", C.c_int, V, P(TimingC))
    sig("ipcr_gpu_synth_fill", C.c_int, V, V, U64, U64, U64, *ERR)
    sig("ipcr_gpu_debug_planes", C.c_int, V, P(C.c_uint32), U64, P(U64), P(C.c_uint32), P(U64))
    sig("ipcr_gpu_debug_codes", C.c_int, V, P(C.c_uint32), P(C.c_uint32), U64, P(U64))
    sig("ipcr_gpu_debug_candidates", C.c_int, V, P(U64), U64, P(U64))
    sig("ipcr_gpu_debug_plan_eval", C.c_int, V, C.c_uint32, C.c_char_p, U64, C.c_int64, C.c_int)
    sig("ipcr_gpu_debug_seed_eval", C.c_int, V, C.c_int, C.c_char_p, U64, P(U64), U64, P(U64), P(C.c_int32))
    sig("ipcr_gpu_debug_orientation", C.c_int, V, C.c_uint32, C.c_char_p, SZ, P(C.c_int))
    sig("ipcr_gpu_debug_jit_source", C.c_int, V, C.c_int, C.c_char_p, SZ, P(SZ))
    sig("ipcr_gpu_debug_jit_compile", C.c_int, V, C.c_int, C.c_char_p, SZ, P(SZ), *ERR)
    _lib = L
    return L


class GpuError(RuntimeError):
    """A failed C-ABI call (CUDA fault, bad argument, no device...)."""


def errbuf():
    return C.create_string_buffer(1024)


def check(rc, buf):
    if rc != 0:
        raise GpuError(buf.value.decode("utf-8", "replace") or "ipcr_gpu call failed")


def mask_to_idx(mask: int, descending: bool = False):
    idx = [j for j in range(64) if (mask >> j) & 1]
    return tuple(reversed(idx)) if descending else tuple(idx)
