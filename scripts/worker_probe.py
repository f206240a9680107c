#!/usr/bin/env python
"""Does the reference's worker model (several scratches scanning different records concurrently,
internal/pipeline/pipeline.go:66-125) buy overlap on one GPU?  Wall clock per full pass."""
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from ipcr_b200 import Config, Engine, sharding, synth  # noqa: E402
from ipcr_b200 import device as D  # noqa: E402

total = 3_000_000_000
lens = synth.contig_lengths(total, 24)
user, allp = synth.multiplex_panel(100, seed=20251018)
eng = Engine(Config(max_mm=2, terminal_window=3, max_len=2000, hit_cap=10000, seed_len=12))
cp = eng.compile_panel(allp)
scs = [eng.new_simulation_scratch(cp) for _ in range(4)]
buf = torch.empty(total, dtype=torch.uint8, device="cuda")
D.synth_fill(scs[0], buf, 0, 1)
torch.cuda.synchronize()
offs = sharding.contig_offsets(lens)
segs = [(buf[offs[c]:offs[c + 1]], c, 0, lens[c]) for c in range(24)]


def split(w):
    # contiguous runs of contigs (adopted segments of one call are scanned as ONE span of the buffer,
    # so the groups must not interleave), balanced by length
    groups, c = [], 0
    for k in range(w):
        target = total * (k + 1) // w
        g = []
        while c < 24 and (offs[c + 1] <= target or not g or k == w - 1):
            g.append(segs[c])
            c += 1
            if k < w - 1 and offs[c] >= target:
                break
        groups.append(g)
    return groups


def one_pass(groups):
    out = [0] * len(groups)

    def run(k):
        _o, n = D.scan_device(cp, scs[k], groups[k], n_recs=24, raw=True)
        out[k] = n

    if len(groups) == 1:
        run(0)
    else:
        th = [threading.Thread(target=run, args=(k,)) for k in range(len(groups))]
        for t in th:
            t.start()
        for t in th:
            t.join()
    return sum(out)


for w in (1, 2, 3, 4):
    groups = split(w)
    for _ in range(3):
        n = one_pass(groups)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        n = one_pass(groups)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 100
    print(f"bases={[sum(int(x[3]) for x in g) for g in groups]} workers={w} wall_ms_per_pass={ms:.3f} products={n} device_ms_each={[round(s.timing()['total_ms'], 3) for s in scs[:w]]}", flush=True)
