"""Quick device-side timing probe (not the bench): synthetic genome, one panel, stage timings."""
import argparse
import sys
import os
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ipcr_b200 import Config, Engine, Pair, SimulationScratch
from ipcr_b200 import device as D


def panel_pairs(n_pairs, seed=7, self_pairs=True):
    import random
    rng = random.Random(seed)
    pairs = []
    if n_pairs == 1:
        base = [("16S", "AGAGTTTGATCMTGGCTCAG", "TACGGYTACCTTGTTAYGACTT")]
    else:
        base = []
        for i in range(n_pairs):
            la, lb = rng.randint(18, 25), rng.randint(18, 25)
            def mk(n):
                s = [rng.choice("ACGT") for _ in range(n)]
                if rng.random() < 0.1:
                    for _ in range(rng.randint(1, 2)):
                        s[rng.randrange(n)] = rng.choice("RYSWKMBDHVN")
                return "".join(s)
            base.append((f"p{i}", mk(la), mk(lb)))
    for (i, a, b) in base:
        pairs.append(Pair(i, a, b))
    if self_pairs:
        seenA, seenB = set(), set()
        for (i, a, b) in base:
            if a not in seenA:
                seenA.add(a); pairs.append(Pair(i + "+A:self", a, a))
            if b not in seenB:
                seenB.add(b); pairs.append(Pair(i + "+B:self", b, b))
    return pairs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mbp", type=float, default=500)
    ap.add_argument("--pairs", type=int, default=1)
    ap.add_argument("--mm", type=int, default=2)
    ap.add_argument("--tw", type=int, default=3)
    ap.add_argument("--jit", type=int, default=1)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--hit-cap", type=int, default=10000)
    a = ap.parse_args()
    n = int(a.mbp * 1e6)
    pairs = panel_pairs(a.pairs)
    eng = Engine(Config(max_mm=a.mm, terminal_window=a.tw, max_len=2000, hit_cap=a.hit_cap))
    t0 = time.time()
    cp = eng.compile_panel(pairs)
    cp.set_jit(a.jit)
    sc = SimulationScratch(cp)
    print(f"panel: {len(pairs)} pairs, {cp.num_orientations} unique orientations, compile {time.time()-t0:.2f}s")
    buf = torch.empty(n, dtype=torch.uint8, device="cuda")
    D.synth_fill(sc, buf, 0, 1234)
    torch.cuda.synchronize()
    for it in range(a.iters):
        t0 = time.time()
        g = D.Genome.adopt(sc, [(buf, 0)])
        t1 = time.time()
        out, np_ = D.scan_genome(cp, sc, g, raw=True)
        t2 = time.time()
        tm = sc.timing()
        print(f"iter {it}: adopt+pack {1e3*(t1-t0):.2f} ms, scan wall {1e3*(t2-t1):.2f} ms, products {np_}, "
              f"filter {tm['filter_ms']:.3f} ms verify {tm['verify_ms']:.3f} pair {tm['pair_ms']:.3f} "
              f"cands {tm['candidates']} sites {tm['sites']} launches {tm['kernel_launches']} jit {tm['used_jit']}")
        if tm["filter_ms"] > 0:
            gbps = n / tm["filter_ms"] / 1e6
            print(f"   filter: {gbps:.1f} Gbp/s = {gbps*0.5:.1f} GB/s algorithmic ({gbps*0.5/6540.2*100:.1f}% of measured HBM copy)")
        g.close()


if __name__ == "__main__":
    main()
