#!/usr/bin/env python
"""bench.py -- headline benchmark of the ipcr scan hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched with torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json config 5, mm = 2 headline): a 3 Gbp synthetic genome in 24
contigs of 50-250 Mbp, a 100-pair multiplex panel (18-25-nt primers, 10 % with IUPAC
codes; + unique self pairs => 300 pairs / 400 unique orientations), --mismatches 2,
--terminal-window 3, --max-length 2000, --hit-cap 10000, 1-3 planted amplicons per pair.

One step = one pass of the hot path (ForEachCompiledProduct over every record):
  value : ASCII genome already resident in HBM -> pack -> filter -> verify -> pair -> products
          to the host; device time from CUDA events on the scratch's own stream.
  e2e   : the same through the reference-facing call with HOST buffers (pinned), H2D and
          D2H inside the timed region, wall clock.
N > 1, --scaling strong (default): ONE 3 Gbp genome is cut into N shards with a 64-base halo;
every rank runs stage 1 on its shard into a fixed-size site block in device memory, ONE NCCL
all-gather moves the blocks (sites never visit the host), stage 2 pairs once; time = max over
ranks of the device time (CUDA events around stage 1 -> all-gather -> stage 2 -> products).
e2e at N > 1: every rank additionally copies its shard from its own pinned host buffer first.
N > 1, --scaling weak: every rank scans its OWN 3 Gbp genome (independent records, no
collective); value = bases of all ranks / max-over-ranks device time.
--impl reference times the CPU restatement of the reference's seeded Aho-Corasick engine
(oracle/, Go toolchain absent) on all host cores over a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "Gbp scanned/sec x primer pairs (mm=2)"
UNIT = "Gbp*pairs/s"
SEED = 20251018


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--genome-mbp", type=float, default=3000.0)
    ap.add_argument("--contigs", type=int, default=24)
    ap.add_argument("--pairs", type=int, default=100)
    ap.add_argument("--mm", type=int, default=2)
    ap.add_argument("--tw", type=int, default=3)
    ap.add_argument("--max-len", type=int, default=2000)
    ap.add_argument("--hit-cap", type=int, default=10000)
    ap.add_argument("--cpu-sample-mbp", type=float, default=0.0, help="0 = calibrate to ~15 s")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--filter", default="auto", choices=["auto", "seed", "bitparallel", "generic"],
                    help="candidate filter: auto = the library's own choice (seed index when selective)")
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="N > 1: strong = ONE genome sharded over the GPUs, site blocks all-gathered on the device, "
                         "paired once (north_star); weak = one whole genome per GPU (independent records, no collective)")
    ap.add_argument("--n-per-mbp", type=float, default=0.0,
                    help="scatter this many N bytes per Mbp over the genome (robustness runs; default off)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--in-process-devices", type=int, default=1,
                    help="N = 1 only: also time the packed end-to-end call spread over K devices of this ONE process "
                         "(ipcr_gpu_init(K); what a single Go host process would see)")
    ap.add_argument("--no-bitparallel", action="store_true", help="skip the bit-parallel comparison pass")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1, sharded: how the site blocks meet -- peer = CUDA IPC peer memory read by the pairing kernel "
                         "itself (falls back to nccl if the mapping fails), nccl = one all-gather")
    ap.add_argument("--no-e2e-fasta", action="store_true", help="skip the FASTA-on-disk -> sorted TSV measurement (N = 1)")
    ap.add_argument("--no-n-variant", action="store_true", help="skip the second value on a genome with scattered N (N = 1)")
    return ap.parse_args()


def workload_config(a, n_gpus, user_pairs, all_pairs, n_orient):
    if n_gpus == 1:
        sharding = "none"
    elif a.scaling == "weak":
        sharding = (f"{n_gpus} genomes of {a.genome_mbp:g} Mbp, one per GPU (independent records, "
                    "no data-path collective)")
    else:
        sharding = (f"{n_gpus} shards of ONE genome, 64-base halo; stage 1 per rank -> one NCCL all-gather of fixed-size "
                    "site blocks (device-resident) -> stage 2 once")
    return {
        "workload": f"C5 multiplex: {a.genome_mbp:g} Mbp synthetic genome ({a.contigs} contigs), "
                    f"{len(user_pairs)}-pair panel (+self => {len(all_pairs)} pairs, {n_orient} unique orientations), "
                    f"--mismatches {a.mm} --terminal-window {a.tw} --max-length {a.max_len} --hit-cap {a.hit_cap}",
        "genome_bp": int(a.genome_mbp * 1e6),
        "genome_bp_all_gpus": int(a.genome_mbp * 1e6) * (n_gpus if a.scaling == "weak" else 1),
        "pairs": len(user_pairs),
        "pairs_with_self": len(all_pairs),
        "unique_orientations": n_orient,
        "max_mm": a.mm,
        "scattered_n_per_mbp": a.n_per_mbp,
        "sharding": sharding,
        "l2": "inputs (1 B/bp ASCII + 0.375-0.5 B/bp packed, per GPU) are >> the 126 MB L2, no flush needed",
    }


# --------------------------------------------------------------------------------------
# CPU reference arm / baseline (oracle layer 2 = the reference's seeded production path)
# --------------------------------------------------------------------------------------
def cpu_reference_run(a, steps, warmup, sample_mbp=None):
    from ipcr_b200 import synth
    from ipcr_b200.device import synth_bases_host
    from oracle import oracle as O

    O.build()
    user_pairs, all_pairs = synth.multiplex_panel(a.pairs, seed=SEED)
    cores = os.cpu_count() or 1
    cfg = O.make_cfg(max_mm=a.mm, terminal_window=a.tw, max_len=a.max_len, hit_cap=a.hit_cap, seed_len=12)
    t0 = time.perf_counter()
    panel = O.Panel(cfg, [(p.forward, p.reverse, 0, 0) for p in all_pairs])
    compile_s = time.perf_counter() - t0
    total = int(a.genome_mbp * 1e6)
    lens = synth.contig_lengths(total, a.contigs)
    planted = synth.plan_amplicons(user_pairs, lens, a.mm, a.tw, seed=SEED + 1)
    offs = [0]
    for ln in lens:
        offs.append(offs[-1] + ln)

    def sample(n_bp):
        n_bp = min(n_bp, lens[0])
        buf = bytearray(synth_bases_host(0, n_bp, SEED))
        for pl in planted:  # amplicons that fall inside the sample (contig 0 prefix)
            if pl.contig == 0 and pl.start + pl.length <= n_bp:
                synth.apply_planted(buf, [0], [pl])
        return bytes(buf)

    chunk = 1_000_000  # best case for the reference: --chunk-size well above max-length, all cores busy
    if sample_mbp is None or sample_mbp <= 0:
        probe = sample(2_000_000)
        t0 = time.perf_counter()
        panel.scan_threaded([probe], chunk_size=chunk, overlap=a.max_len, threads=cores)
        dt = max(1e-3, time.perf_counter() - t0)
        sample_bp = int(min(lens[0], max(2_000_000, 2_000_000 * (15.0 / dt) / max(1, steps + warmup))))
    else:
        sample_bp = int(sample_mbp * 1e6)
    seq = sample(sample_bp)
    for _ in range(warmup):
        panel.scan_threaded([seq], chunk_size=chunk, overlap=a.max_len, threads=cores)
    t0 = time.perf_counter()
    n_prod = 0
    for _ in range(steps):
        n_prod = panel.scan_threaded([seq], chunk_size=chunk, overlap=a.max_len, threads=cores)
    dt = (time.perf_counter() - t0) / max(1, steps)
    gbp_s = len(seq) / 1e9 / dt
    # the reference's DEFAULT flags: --chunk-size 0 => one worker per RECORD (internal/pipeline/pipeline.go:164-173),
    # so a genome keeps at most n_contigs cores busy and finishes with its longest contig.  Bounded sample: every
    # contig scaled by the same factor (same imbalance), bases from the same generator.
    default_flags = None
    try:
        scale = min(1.0, max(len(seq), 8_000_000) / float(total))
        pieces = [bytes(synth_bases_host(offs[c], max(1000, int(lens[c] * scale)), SEED)) for c in range(len(lens))]
        t0 = time.perf_counter()
        panel.scan_threaded(pieces, chunk_size=0, overlap=0, threads=cores)
        dt0 = max(1e-6, time.perf_counter() - t0)
        g0 = sum(len(x) for x in pieces) / 1e9 / dt0
        default_flags = {"value": g0 * len(user_pairs), "gbp_per_s": g0, "unit": UNIT, "cores": min(cores, len(lens)),
                         "kind": "port",
                         "sample": f"all {len(lens)} contigs scaled to {sum(len(x) for x in pieces) / 1e6:.1f} Mbp, one worker per "
                                   f"record, no chunking (the reference's default flags); {cores} threads available"}
    except Exception as ex:  # never sink the main number
        default_flags = {"value": None, "error": str(ex)}
    return {
        "default_flags": default_flags,
        "sample_mbp": len(seq) / 1e6,
        "value": gbp_s * len(user_pairs),
        "gbp_per_s": gbp_s,
        "unit": UNIT,
        "cores": cores,
        "kind": "port",
        "sample": f"{len(seq) / 1e6:.1f} Mbp prefix of contig 0 of the same genome, full {len(all_pairs)}-pair panel, "
                  f"seeded AC engine restated in C (oracle/ipcr_seeded.c), {cores} threads over 1 Mbp chunks "
                  f"(--chunk-size 1000000, overlap {a.max_len}); panel compile {compile_s:.1f} s not timed; "
                  f"{n_prod} products",
        "ms_per_step": dt * 1e3,
        "patterns": panel.num_patterns,
        "automaton_nodes": panel.num_nodes,
    }, user_pairs, all_pairs


def unique_orientations(pairs):
    """A, B, rc(A), rc(B) de-duplicated over the panel, like CompilePanel on the GPU side."""
    from ipcr_b200 import synth
    seen = set()
    for p in pairs:
        for s_ in (p.forward, p.reverse):
            seen.add((s_, False))
            seen.add((synth.revcomp(s_), True))
    return len(seen)


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    res, user_pairs, all_pairs = cpu_reference_run(a, a.steps, a.warmup, a.cpu_sample_mbp or None)
    line = {
        "impl": "reference",
        "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
        "scaling": a.scaling, "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(a, a.gpus, user_pairs, all_pairs, unique_orientations(all_pairs)),
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "cpu_baseline_default_flags": res["default_flags"],
        "reference_sample_mbp": res["sample_mbp"],
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gbp_per_s": res["gbp_per_s"],
        "note": "Go toolchain absent in this image: this is the C restatement of the reference's seeded engine; each step "
                "scans a bounded PREFIX of contig 0 of the same genome (reference_sample_mbp) with the full panel and the "
                "value is that rate -- a streaming automaton's rate does not depend on how much of the genome it is shown",
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region.  The region is tens of milliseconds
    long, so NVML is polled from a thread every ~2 ms; `nvidia-smi -lms` is the fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, gpu_index=0):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index
        self.nvml = None
        self.samples = []   # (sm_mhz, reasons bitmask)
        self.max_mhz = None
        self.run = False

    def _visible_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        if vis:
            ids = [x.strip() for x in vis.split(",") if x.strip()]
            if self.gpu < len(ids) and ids[self.gpu].isdigit():
                return int(ids[self.gpu])
        return self.gpu

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.handle = None
            try:  # the CUDA device's own PCI address (CUDA and NVML may order devices differently)
                import torch
                pr = torch.cuda.get_device_properties(self.gpu)
                bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
                self.handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
            except Exception:
                self.handle = None
            if self.handle is None:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._visible_index())
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.run = True
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while self.run:
            try:
                mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                try:
                    why = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:
                    why = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.samples.append((mhz, why))
            except Exception:
                pass
            time.sleep(0.002)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.run = False
            self.t.join(timeout=1.0)
            sm = sorted(x[0] for x in self.samples)
            bits = 0
            for _m, w in self.samples:
                bits |= w
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz,
                    "reasons": sorted(name for bit, name in self.REASONS if bits & bit), "samples": len(sm),
                    "source": "nvml, polled every ~2 ms inside the timed region"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 100"}


# dram bytes (read + write) of one launch from the committed `ncu --set full` capture of the
# bench command (profiles/README.md says which file); None until captured
NCU_TRAFFIC = {"k_seed_scan": 1.1927e9,    # profiles/r02_ncu_full_k_seed_scan.csv (3 Gbp, 400 orientations): 1.1882 GB read + 4.5 MB written
               "ipcr_scan_jit": 1.635e9}   # profiles/r01_ncu_full_scan_jit_and_pack_raw.csv (3 Gbp, one group)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def fasta_to_tsv(a, cp, sc, buf, lens, offs, user_pairs, total):
    """Writes the synthetic genome as an 80-column FASTA (untimed), then times FASTA -> sorted TSV."""
    import numpy as np
    import torch

    from ipcr_b200 import hostlib as HL
    from ipcr_b200.fasta import FastaFile

    d = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else "/tmp"
    path = os.path.join(d, f"ipcr_bench_{os.getpid()}.fa")
    try:
        with open(path, "wb") as f:
            for c, n in enumerate(lens):
                f.write(b">chr%d synthetic\n" % (c + 1))
                seq = buf[offs[c]:offs[c + 1]].cpu().numpy()
                full = n // 80 * 80
                body = np.empty((full // 80, 81), dtype=np.uint8)
                body[:, :80] = seq[:full].reshape(-1, 80)
                body[:, 80] = 10
                f.write(body.tobytes())
                if full < n:
                    f.write(seq[full:].tobytes() + b"\n")
        size = os.path.getsize(path)
        threads = os.cpu_count() or 1
        best = None
        for _ in range(2):  # second pass: page cache and pinned pools warm
            t0 = time.perf_counter()
            ff = FastaFile(path, threads=0, pin=False, pack_only=True)
            t1 = time.perf_counter()
            prods = ff.scan(cp, sc)
            t2 = time.perf_counter()
            prods.sort(key=HL.less_key)
            text = "\n".join([HL.TSV_HEADER] + [HL.format_base_row(p) for p in prods]) + "\n"
            t3 = time.perf_counter()
            scan_t = sc.timing()
            ff.close()
            cur = {"total_s": t3 - t0, "open_parse_pack_s": t1 - t0, "scan_s": t2 - t1, "sort_format_s": t3 - t2,
                   "scan_device_ms": scan_t["total_ms"], "rows": len(prods), "tsv_bytes": len(text)}
            if best is None or cur["total_s"] < best["total_s"]:
                best = cur
        g = total / 1e9
        best.update({"value": g / best["total_s"] * len(user_pairs), "unit": UNIT, "gbp_per_s": g / best["total_s"],
                     "fasta_bytes": size, "feeder_threads": threads, "file_system": d,
                     "api": "ipcr_gpu_fasta_open(path, IPCR_FASTA_PACK_ONLY) -> ipcr_gpu_scan_fasta -> --sort order -> TSV rows "
                            "(page cache warm; best of 2)"})
        return best
    finally:
        try:
            os.remove(path)
        except OSError:
            pass


# --------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------
def run_ours(a):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus and world > 1:
        a.gpus = world
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        torch.cuda.set_device(0)
    dev = torch.device("cuda", torch.cuda.current_device())

    from ipcr_b200 import Config, Engine, _capi, sharding, synth
    from ipcr_b200 import device as D

    L = _capi.lib()
    err = _capi.errbuf()
    # one process per GPU: the thread is bound to its device by torch; one device per call
    if L.ipcr_gpu_init(1, err, len(err)) != 0:
        raise RuntimeError(err.value.decode())

    user_pairs, all_pairs = synth.multiplex_panel(a.pairs, seed=SEED)
    eng = Engine(Config(max_mm=a.mm, terminal_window=a.tw, max_len=a.max_len, hit_cap=a.hit_cap, seed_len=12))
    cp = eng.compile_panel(all_pairs)
    cp.set_jit({"auto": 1, "seed": 3, "bitparallel": 2, "generic": 0}[a.filter])
    sc = eng.new_simulation_scratch(cp)

    total = int(a.genome_mbp * 1e6)
    lens = synth.contig_lengths(total, a.contigs)
    offs = [0]
    for ln in lens:
        offs.append(offs[-1] + ln)
    planted = synth.plan_amplicons(user_pairs, lens, a.mm, a.tw, seed=SEED + 1)
    weak = world > 1 and a.scaling == "weak"
    shards, shard_rank = (1, 0) if weak else (world, rank)
    # this rank's slice [lo, hi) of the concatenated genome (+ halo); weak: its own whole genome
    lo, hi, piece_end = sharding.rank_slice(total, shards, shard_rank)
    buf = torch.empty(piece_end - lo, dtype=torch.uint8, device=dev)
    D.synth_fill(sc, buf, lo, SEED + (7919 * rank if weak else 0))
    for pl in planted:  # write the (clipped) planted primers that fall into this slice
        base = offs[pl.contig] + pl.start
        for off, data in ((0, pl.left), (pl.length - len(pl.right), pl.right)):
            s0, s1 = base + off, base + off + len(data)
            c0, c1 = max(s0, lo), min(s1, piece_end)
            if c0 < c1:
                t = torch.frombuffer(bytearray(data[c0 - s0:c1 - s0]), dtype=torch.uint8)
                buf[c0 - lo:c1 - lo].copy_(t)
    if a.n_per_mbp > 0:  # scattered ambiguity codes (input generation, untimed); the same positions on every rank
        gen = torch.Generator()
        gen.manual_seed(SEED + 17 + (7919 * rank if weak else 0))
        k = int(a.n_per_mbp * total / 1e6)
        idx = torch.randint(0, total, (k,), generator=gen)
        idx = idx[(idx >= lo) & (idx < piece_end)] - lo
        buf[idx.to(dev)] = ord("N")
    torch.cuda.synchronize()

    segments = [(buf[s0 - lo:s1 - lo], c, rec_off, own)
                for (s0, s1, c, rec_off, own) in sharding.shard_segments(lens, shards, shard_rank)]
    n_recs = len(lens)

    def step_single():
        _out, n = D.scan_device(cp, sc, segments, n_recs=n_recs, raw=True)
        t = sc.timing()
        return n, t["total_ms"], t

    # ---- sharded: stage 1 -> all-gather of site blocks -> stage 2, all on the scratch's stream ----
    ext = torch.cuda.ExternalStream(D.scratch_stream(sc), device=dev)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    blocks = {"cap": 0, "mine": None, "all": None}
    host_slice = {}

    # the exchange: peer memory (the pairing kernel reads every rank's block in place) or one NCCL all-gather
    peer = {"x": None, "why": None}
    last_raw = {"p": None}
    if world > 1 and not weak and a.exchange == "peer":
        try:  # collective and all-or-nothing: it raises on every rank if any rank cannot export / map
            peer["x"] = D.PeerExchange(dist, sc, n_recs, rank, world, dev)
        except Exception as ex:  # e.g. CUDA IPC not permitted in this container: one NCCL all-gather instead
            peer["x"], peer["why"] = None, str(ex)

    def feed_packed():
        """H2D of this rank's packed shard: the codes, and the bitmap as its run list when it has one."""
        host_slice["dcodes"].copy_(host_slice["codes"], non_blocking=True)
        if host_slice.get("runs") is not None:
            runs, n_runs = host_slice["runs"]
            D.inval_from_runs(sc, host_slice["dinval"], host_slice["n"], runs, n_runs)
        else:
            host_slice["dinval"].copy_(host_slice["inval"], non_blocking=True)

    def step_peer(from_host=False):
        launches = 0
        while True:
            if D.scratch_site_cap(sc) != peer["x"].site_cap:  # a block overflowed: bigger blocks, new mappings
                peer["x"].close()
                dist.barrier()  # every import is closed before any rank frees the block it exported
                peer["x"] = D.PeerExchange(dist, sc, n_recs, rank, world, dev)
            with torch.cuda.stream(ext):
                evs[0].record()
                blk = peer["x"].begin_step()
                if from_host:
                    feed_packed()
                    D.find_sites_block_packed(cp, sc, host_slice["dcodes"], host_slice["dinval"], host_slice["n"],
                                              host_slice["segs"], n_recs, blk, peer["x"].site_cap)
                else:
                    D.find_sites_block(cp, sc, segments, n_recs, blk, peer["x"].site_cap)
                peer["x"].publish()  # nothing to launch: the verify kernel's last block already published
                launches += sc.timing()["kernel_launches"]
                evs[1].record()
                evs[2].record()
                res = peer["x"].pair(cp, lens)
                evs[3].record()
            if res is not None:
                break
        evs[3].synchronize()
        t = dict(sc.timing())
        t["kernel_launches"] = launches + 1
        t["gather_ms"] = 0.0
        t["find_total_ms"] = evs[0].elapsed_time(evs[1])
        last_raw["p"] = res[0]
        return res[1], evs[0].elapsed_time(evs[3]), t

    def step_multi(from_host=False):
        if peer["x"] is not None:
            return step_peer(from_host)
        launches = 0
        while True:
            site_cap = D.scratch_site_cap(sc)
            if blocks["cap"] != site_cap:
                bb = D.site_block_bytes(n_recs, site_cap)
                blocks.update(cap=site_cap, mine=torch.empty(bb, dtype=torch.uint8, device=dev),
                              all=torch.empty(world * bb, dtype=torch.uint8, device=dev))
            with torch.cuda.stream(ext):
                h0 = time.perf_counter()
                evs[0].record()
                if from_host:  # this rank's shard, packed, from its own pinned host buffers
                    feed_packed()
                    D.find_sites_block_packed(cp, sc, host_slice["dcodes"], host_slice["dinval"], host_slice["n"],
                                              host_slice["segs"], n_recs, blocks["mine"], site_cap)
                else:
                    D.find_sites_block(cp, sc, segments, n_recs, blocks["mine"], site_cap)
                launches += sc.timing()["kernel_launches"]
                evs[1].record()
                h1 = time.perf_counter()
                dist.all_gather_into_tensor(blocks["all"], blocks["mine"])
                evs[2].record()
                h2 = time.perf_counter()
                # every rank pairs: all see the same blocks, so all take the same retry decisions
                res = D.pair_blocks(cp, sc, blocks["all"], world, site_cap, lens, raw=True)
                evs[3].record()
                h3 = time.perf_counter()
                if os.environ.get("IPCR_BENCH_TRACE") and rank == 0:
                    print(f"[trace] host ms: find {1e3*(h1-h0):.3f} gather {1e3*(h2-h1):.3f} pair {1e3*(h3-h2):.3f}", file=sys.stderr)
            if res is not None:
                break
        evs[3].synchronize()
        t = dict(sc.timing())
        t["kernel_launches"] = launches + 1
        t["gather_ms"] = evs[1].elapsed_time(evs[2])
        t["find_total_ms"] = evs[0].elapsed_time(evs[1])
        last_raw["p"] = res[0]
        return res[1], evs[0].elapsed_time(evs[3]), t

    step = step_single if world == 1 or weak else step_multi

    for _ in range(max(3, a.warmup)):
        n_prod, _ms, tm = step()
    if peer["x"] is not None:  # the two exchanges must hand stage 2 the same sites: identical product bytes
        import ctypes
        n_a, _ms, _t = step()
        bytes_peer = ctypes.string_at(last_raw["p"], n_a * 64)
        px, peer["x"] = peer["x"], None
        n_b, _ms, _t = step()
        bytes_nccl = ctypes.string_at(last_raw["p"], n_b * 64)
        peer["x"] = px
        if bytes_peer != bytes_nccl:
            raise RuntimeError(f"peer exchange and NCCL all-gather disagree ({n_a} vs {n_b} products)")
    sampler = ClockSampler(torch.cuda.current_device())
    if rank == 0:
        sampler.start()  # NVML start-up takes tens of ms: before the barrier, or the other ranks wait for it in step 1
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    wall0 = time.perf_counter()
    dev_ms_sum, launches, filt_ms, last = 0.0, 0, 0.0, None
    for _ in range(a.steps):
        n_prod, ms, tm = step()
        dev_ms_sum += ms
        launches += tm["kernel_launches"]
        filt_ms += tm["filter_ms"]
        last = tm
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - wall0) * 1e3 / a.steps
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = dev_ms_sum / a.steps
    if world > 1:
        t = torch.tensor([ms_per_step, wall_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_per_step, wall_ms = float(t[0]), float(t[1])
        lt = torch.tensor([launches, n_prod if weak else 0], dtype=torch.int64, device=dev)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt[0].item())
        if weak:
            n_prod = int(lt[1].item())
    # ---- end to end through the reference-facing call with host buffers: every rank feeds its own
    # genome from pinned host memory (N = 1, or weak scaling); wall clock, max over ranks ----
    e2e, e2e_stage = None, None
    n_units = world if weak else 1
    if not a.no_e2e and (world == 1 or weak):
        host = torch.empty(total, dtype=torch.uint8, pin_memory=True)
        host.copy_(buf)
        torch.cuda.synchronize()
        recs = [host[offs[c]:offs[c + 1]] for c in range(n_recs)]
        # (a) plain reference-facing call: ASCII records, 1 B/bp over PCIe
        D.scan_batch(cp, sc, recs)  # warm (buffer growth)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(max(1, a.e2e_steps)):
            prods = D.scan_batch(cp, sc, recs)
        ascii_ms = (time.perf_counter() - t0) * 1e3 / max(1, a.e2e_steps)
        ascii_stage = {k: sc.timing()[k] for k in ("h2d_ms", "filter_ms", "pair_ms", "total_ms")}
        assert len(prods) == n_prod or weak
        # (a') the same call with PAGEABLE host records -- what a Go []byte or a Python bytes object is: the library
        # stages them through its own pinned ring with several host threads (N = 1 only: it is a host-side figure)
        pageable_ms, pageable_stage = None, None
        if world == 1:
            host_pg = torch.empty(total, dtype=torch.uint8)  # not pinned
            host_pg.copy_(host)
            recs_pg = [host_pg[offs[c]:offs[c + 1]] for c in range(n_recs)]
            D.scan_batch(cp, sc, recs_pg)  # warm (the ring is allocated once per scratch)
            t0 = time.perf_counter()
            for _ in range(max(1, a.e2e_steps)):
                prods_pg = D.scan_batch(cp, sc, recs_pg)
            pageable_ms = (time.perf_counter() - t0) * 1e3 / max(1, a.e2e_steps)
            pageable_stage = {k: sc.timing()[k] for k in ("h2d_ms", "filter_ms", "pair_ms", "total_ms")}
            assert len(prods_pg) == n_prod
            del host_pg, recs_pg
        # (b) the native feeder's packed output (0.375 B/bp, produced once when the FASTA is read): the
        # records are handed to the feeder untimed, every timed step ships the packed form and scans it
        from ipcr_b200.fasta import FastaFile
        tp0 = time.perf_counter()
        ff = FastaFile.from_memory(recs, pin=True, pack=True)
        feeder_pack_s = time.perf_counter() - tp0
        ff.scan_raw(cp, sc)  # warm
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        d2h = 0
        for _ in range(max(1, a.e2e_steps)):
            _out, n_e = ff.scan_raw(cp, sc)
            d2h = n_e * 64
        e2e_ms = (time.perf_counter() - t0) * 1e3 / max(1, a.e2e_steps)
        te = sc.timing()
        if not weak and n_e != n_prod:
            raise RuntimeError(f"packed e2e path returned {n_e} products, device-resident path {n_prod}")
        one_process = None
        if world == 1 and a.in_process_devices > 1:
            from ipcr_b200 import _capi
            L, err = _capi.lib(), _capi.errbuf()
            if L.ipcr_gpu_init(a.in_process_devices, err, len(err)) != 0:
                raise RuntimeError(err.value.decode())
            try:
                for _ in range(2):
                    ff.scan_raw(cp, sc)  # warm: peer scratches, tables on every device
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(max(1, a.e2e_steps)):
                    _out, n_m = ff.scan_raw(cp, sc)
                mp_ms = (time.perf_counter() - t0) * 1e3 / max(1, a.e2e_steps)
                tm = sc.timing()
                if n_m != n_prod:
                    raise RuntimeError(f"in-process multi-device e2e returned {n_m} products, one device {n_prod}")
                one_process = {"devices": int(L.ipcr_gpu_num_devices()), "ms_per_step": mp_ms,
                               "value": total / 1e9 / (mp_ms / 1e3) * len(user_pairs),
                               "stage_ms": {k: tm[k] for k in ("h2d_ms", "filter_ms", "verify_ms", "pair_ms", "total_ms")},
                               "api": "the same ipcr_gpu_scan_fasta call after ipcr_gpu_init(K): every device copies its "
                                      "slice of the codes over its own PCIe link, blocks meet by NVLink peer copies"}
            finally:
                L.ipcr_gpu_init(1, err, len(err))
        ff.close()
        if world > 1:
            t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_ms = float(t[0])
            t = torch.tensor([d2h], dtype=torch.int64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            d2h = int(t[0])
        e2e_stage = {k: te[k] for k in ("h2d_ms", "pack_ms", "filter_ms", "verify_ms", "pair_ms", "total_ms")}
        g = n_units * total / 1e9
        # 2-bit codes (8 bytes per 32 bases); the hard-mismatch bitmap of this N-free genome travels as a run list
        # of a few bytes (12 bytes per 32 bases when --n-per-mbp scatters too many runs for that)
        packed_bytes = n_units * ((total + 4095) // 4096 * 4096 // 32 + 1) * (8 if a.n_per_mbp == 0 else 12)
        e2e = {"value": g / (e2e_ms / 1e3) * len(user_pairs), "unit": UNIT, "ms_per_step": e2e_ms,
               "gbp_per_s": g / (e2e_ms / 1e3), "h2d_bytes_per_step": packed_bytes, "d2h_bytes_per_step": d2h,
               "api": "ipcr_gpu_scan_fasta(feeder handle holding the packed records in pinned host memory) -> products "
                      "(C ABI): H2D of the 2-bit codes in slabs under the scan (the hard-mismatch bitmap as a run list, rebuilt on "
                      "the device), verify on the codes",
               "feeder_pack_s_untimed": feeder_pack_s,
               "ascii": {"value": g / (ascii_ms / 1e3) * len(user_pairs), "ms_per_step": ascii_ms,
                         "h2d_bytes_per_step": n_units * total, "stage_ms": ascii_stage,
                         "api": "ipcr_gpu_scan_batch(ASCII host records, pinned) -> products: 1 B/bp over PCIe",
                         "pageable_ms_per_step": pageable_ms, "pageable_stage_ms": pageable_stage}}
        if one_process is not None:
            e2e["one_process"] = one_process
        del host, recs
    elif not a.no_e2e:
        # sharded: every rank feeds ITS shard from its own pinned host buffer; wall clock between barriers
        hs = torch.empty(piece_end - lo, dtype=torch.uint8, pin_memory=True)
        hs.copy_(buf)
        torch.cuda.synchronize()
        # this rank's shard as the feeder leaves it: packed, pinned (untimed); record flags from the bytes
        hc, hi_ = D.pack_host(hs)
        psegs = []
        for (s0, s1, c, rec_off, own) in sharding.shard_segments(lens, shards, shard_rank):
            seg = hs[s0 - lo:s1 - lo]
            bad = bool(((seg != 65) & (seg != 67) & (seg != 71) & (seg != 84) & (seg != 97) & (seg != 99)
                        & (seg != 103) & (seg != 116)).any())
            psegs.append((s0 - lo, s1 - s0, own, rec_off, c, bad))
        host_slice.update(t=hs, codes=hc, inval=hi_, dcodes=torch.empty_like(hc, device=dev),
                          dinval=torch.empty_like(hi_, device=dev), segs=psegs, n=piece_end - lo,
                          runs=D.packed_runs(hi_, piece_end - lo, max_runs=max(4096, hi_.numel() // 64)))
        step_multi(from_host=True)  # warm
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        n_e2e = 0
        for _ in range(max(1, a.e2e_steps)):
            n_e2e, _ms, te = step_multi(from_host=True)
        torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / max(1, a.e2e_steps)
        t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t[0])
        t = torch.tensor([host_slice["codes"].numel() * 4 + (host_slice["runs"][1] * 16 if host_slice.get("runs") is not None
                                                              else host_slice["inval"].numel() * 4)], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        h2d = int(t[0])
        e2e_stage = {k: te[k] for k in ("pack_ms", "filter_ms", "verify_ms", "pair_ms", "gather_ms", "find_total_ms")}
        g = total / 1e9
        e2e = {"value": g / (e2e_ms / 1e3) * len(user_pairs), "unit": UNIT, "ms_per_step": e2e_ms,
               "gbp_per_s": g / (e2e_ms / 1e3), "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": n_e2e * 64 * world,
               "api": "per rank: packed shard (2-bit codes + the bitmap's run list, pinned host) -> H2D -> ipcr_gpu_find_sites_block_packed "
                      "-> NCCL all-gather of site blocks -> ipcr_gpu_pair_blocks -> products on the host of every rank (C ABI)"}
        if n_e2e != n_prod:
            raise RuntimeError(f"packed sharded e2e returned {n_e2e} products, device-resident path {n_prod}")
        host_slice.clear()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    n_orient = cp.num_orientations
    gbp = n_units * total / 1e9
    gbp_s = gbp / (ms_per_step / 1e3)
    value = gbp_s * len(user_pairs)
    peak, peak_src = measured_peak()
    n_groups = (n_orient + 3) // 4
    my_bp = piece_end - lo
    dur_ms = (filt_ms / a.steps) / max(1, n_groups)
    achieved = (my_bp * 0.5 / 1e9) / (dur_ms / 1e3) if dur_ms > 0 else 0.0
    path = int(last["used_jit"]) if last else 0
    roofline = {
        "bound": "hbm", "kernel": "ipcr_scan_jit (one launch per group of 4 orientations)",
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak else None,
        "traffic": NCU_TRAFFIC.get("ipcr_scan_jit") if my_bp == 3_000_000_000 else None, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": my_bp * 0.5, "avg_launch_ms": dur_ms, "launches_per_step": n_groups,
        "note": "0.5 B/bp x bases one launch streams; a 400-orientation panel needs 100 such passes, so the "
                "whole scan is LOP3-issue bound, each pass is HBM-streaming (see DESIGN.md)",
    }
    if path == 2:  # seed index: ONE pass over 2-bit codes + hard-mismatch bitmap serves the whole panel
        dur_ms = filt_ms / a.steps
        achieved = (my_bp * 0.375 / 1e9) / (dur_ms / 1e3) if dur_ms > 0 else 0.0
        sm_hz = 1.965e9
        lds_peak = 148 * 32 * sm_hz / 3.53  # random 32-lane access: E[max bank load] = 3.53 cycles per warp lookup
        pack_ms = last["pack_ms"]
        try:  # anchor stride of the seed index's first level (1 for the 100-pair panel, 2 or 4 for small panels)
            stride = max(1, int(cp.seed_eval(b"ACGT" * 16)[1].get("stride", 1)))
        except Exception:
            stride = 1
        roofline = {
            "bound": "hbm", "kernel": "k_seed_scan: one launch for all orientations",
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak else None,
            "traffic": NCU_TRAFFIC.get("k_seed_scan") if my_bp == 3_000_000_000 and a.mm == 2 and a.pairs == 100 else None, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": my_bp * 0.375, "avg_launch_ms": dur_ms, "launches_per_step": 1,
            "note": "0.25 B/bp codes + 0.125 B/bp bitmap, read once for the WHOLE panel (the bit-parallel filter "
                    "re-reads 0.5 B/bp per group of 4 orientations).  Not HBM-bound by design: the binding "
                    "resource is one random shared-memory lookup per genome position (see DESIGN.md)",
            "shared_memory_lookups": {"achieved_per_s": my_bp / stride / (dur_ms / 1e3) if dur_ms > 0 else None,
                                      "peak_per_s_random_banks": lds_peak, "positions_per_lookup": stride,
                                      "frac": (my_bp / stride / (dur_ms / 1e3)) / lds_peak if dur_ms > 0 else None},
            "pack_kernel": {"kernel": "k_pack2", "bound": "hbm", "bytes": my_bp * 1.375, "ms": pack_ms,
                            "achieved": my_bp * 1.375 / 1e9 / (pack_ms / 1e3) if pack_ms > 0 else None,
                            "frac": (my_bp * 1.375 / 1e9 / (pack_ms / 1e3)) / peak if pack_ms > 0 and peak else None},
        }
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(3, a.warmup),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": workload_config(a, world, user_pairs, all_pairs, n_orient),
        "gbp_per_s": gbp_s, "wall_ms_per_step": wall_ms, "products": n_prod, "planted_amplicons": len(planted) * n_units,
        "stage_ms": {k: last[k] for k in ("pack_ms", "filter_ms", "verify_ms", "pair_ms", "d2h_ms")} if last else None,
        "candidates": last["candidates"] if last else None, "sites": last["sites"] if last else None,
        "rank0_ms": ({k: last.get(k) for k in ("find_total_ms", "gather_ms")} if last and world > 1 else None),
        "exchange": (None if world == 1 or weak else
                     ("peer memory: every rank's pairing kernel reads the other ranks' site blocks in place (CUDA IPC over "
                      "NVLink, release/acquire on a step counter); no collective on the data path" if peer["x"] is not None
                      else "one NCCL all-gather of the site blocks" + (f" (peer mapping failed: {peer['why']})" if peer["why"] else ""))),
        "filter_path": {0: "table-driven bit-parallel", 1: "NVRTC bit-parallel", 2: "seed index"}.get(path),
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
    }
    if n_prod < len(planted) * n_units:
        line["warning"] = f"only {n_prod} products for {len(planted) * n_units} planted amplicons"

    line["e2e"] = e2e
    if e2e_stage is not None:
        line["e2e_stage_ms"] = e2e_stage

    # ---- FASTA on disk -> sorted TSV (SURVEY 8d "end to end"): native feeder (mmap, parse, upper-case, pack),
    # packed H2D in slabs under the scan, --sort order, TSV rows.  Page cache warm; N = 1 ----
    if world == 1 and not a.no_e2e_fasta:
        try:
            line["e2e_fasta"] = fasta_to_tsv(a, cp, sc, buf, lens, offs, user_pairs, total)
        except Exception as ex:
            line["e2e_fasta"] = {"error": str(ex)}

    # ---- second value: the same genome with N scattered over it (real assemblies hold N: with --hit-cap > 0 the
    # reference then caps rc orientations before the 3' window filter, which needs the raw-rc index) ----
    if world == 1 and not a.no_n_variant and a.n_per_mbp == 0:
        try:
            gen = torch.Generator()
            gen.manual_seed(SEED + 17)
            idx = torch.randint(0, total, (int(10 * total / 1e6),), generator=gen).to(dev)
            keep = buf[idx].clone()
            buf[idx] = ord("N")
            for _ in range(3):
                step_single()
            ms_n, k = 0.0, max(3, a.steps)
            for _ in range(k):
                n_n, ms, tmn = step_single()
                ms_n += ms
            buf[idx] = keep
            g_n = total / 1e9 / (ms_n / k / 1e3)
            line["with_scattered_n"] = {
                "n_per_mbp": 10, "ms_per_step": ms_n / k, "gbp_per_s": g_n, "value": g_n * len(user_pairs), "unit": UNIT,
                "products": n_n, "plan_variant": int(tmn["plan_variant"]),
                "stage_ms": {k2: tmn[k2] for k2 in ("pack_ms", "filter_ms", "verify_ms", "pair_ms", "d2h_ms")},
                "note": "plan_variant 1 = raw-rc seed index (hit-cap parity on records holding a reset byte)"}
            step_single()  # back on the N-free genome (and its plan variant) for what follows
        except Exception as ex:
            line["with_scattered_n"] = {"error": str(ex)}

    # ---- the bit-parallel filter on the same input, for the per-pass HBM roofline (N = 1) ----
    if world == 1 and path == 2 and not a.no_bitparallel:
        try:
            cp.set_jit(2)
            for _ in range(2):
                step_single()
            f_ms, tot_ms, k = 0.0, 0.0, 2
            for _ in range(k):
                _n, ms, tmb = step_single()
                f_ms += tmb["filter_ms"]
                tot_ms += ms
            per_launch = f_ms / k / max(1, n_groups)
            ach = (my_bp * 0.5 / 1e9) / (per_launch / 1e3) if per_launch > 0 else 0.0
            line["bitparallel"] = {
                "kernel": "ipcr_scan_jit (NVRTC, one launch per group of 4 orientations)", "ms_per_step": tot_ms / k,
                "filter_ms": f_ms / k, "launches_per_step": n_groups, "avg_launch_ms": per_launch,
                "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                             "frac": ach / peak if peak else None, "algorithmic_bytes_per_launch": my_bp * 0.5},
                "note": "same products; the literal AND/popcount design of north_star, kept as the fallback for inputs "
                        "too dirty for the seed index",
            }
            cp.set_jit({"auto": 1, "seed": 3, "bitparallel": 2, "generic": 0}[a.filter])
        except Exception as ex:  # the comparison pass must never sink the main line
            line["bitparallel"] = {"error": str(ex)}
            cp.set_jit({"auto": 1, "seed": 3, "bitparallel": 2, "generic": 0}[a.filter])

    if world == 1 and not a.no_cpu_baseline:
        try:
            res, _u, _p = cpu_reference_run(a, 1, 0, a.cpu_sample_mbp or None)
            line["cpu_baseline"] = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}
            line["cpu_baseline"]["gbp_per_s"] = res["gbp_per_s"]
            line["cpu_baseline_default_flags"] = res["default_flags"]
            line["reference_sample_mbp"] = res["sample_mbp"]
        except Exception as ex:  # the baseline must never sink the GPU line
            line["cpu_baseline"] = {"value": None, "error": str(ex)}
    sys.stdout.flush()
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    # NCCL prints its version banner on stdout when NCCL_DEBUG is set; keep stdout for the JSON line
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ.pop("NCCL_DEBUG")
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
