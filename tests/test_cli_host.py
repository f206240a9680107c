"""The C++ host (ipcr_b200/ipcr_gpu_cli: flags -> FASTA feeder -> GPU scan -> device visitors ->
--sort -> TSV) against the oracle-driven restatement of the reference host (ref_host.py)."""
import os
import random
import subprocess

import pytest

import ipcr_test_helpers as H
import ref_host as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "ipcr_b200", "ipcr_gpu_cli")


def run_cli(*args, expect_rc=0):
    r = subprocess.run([CLI] + [str(a) for a in args], capture_output=True, text=True, timeout=600)
    assert r.returncode == expect_rc, (r.returncode, r.stderr[-500:])
    return r.stdout


def test_cli_is_built_and_rejects_bad_usage():
    assert os.path.exists(CLI), "run `python -m ipcr_b200.build`"
    # internal/clibase/common.go:124-174
    for args, msg in ((["ipcr", "x.fa"], "provide --primers or --forward/--reverse"),
                      (["ipcr", "-f", "ACGT", "x.fa"], "must be supplied together"),
                      (["ipcr", "-f", "ACGT", "-r", "ACGX", "x.fa"], "invalid primer base"),
                      (["ipcr", "-f", "ACGT", "-r", "ACGT"], "at least one sequence file"),
                      (["ipcr", "-f", "ACGT", "-r", "ACGT", "--bogus", "x.fa"], "not defined")):
        r = subprocess.run([CLI] + args, capture_output=True, text=True)
        assert r.returncode == 2 and msg in r.stderr, (args, r.stderr)


def _write(tmp_path, name, recs, width=70):
    p = tmp_path / name
    with open(p, "wb") as f:
        for rid, seq in recs:
            f.write(b">" + rid.encode() + b" x\n")
            for i in range(0, len(seq), width):
                f.write(seq[i:i + width] + b"\n")
    return str(p)


def _genome(rng, n, a, b, k, probe=None, inner=None, n_frac=0.0005):
    g = H.random_genome(rng, n, n_frac=n_frac, lower_frac=0.05)
    for i in range(k):
        pos = 150 + i * (n - 2600) // max(1, k)
        H.plant_amplicon(rng, g, pos, a, b, rng.randint(300, 1800), rng.randint(0, 2), rng.randint(0, 2), 3,
                         revcomp_strand=i % 2 == 1)
        if probe:
            pc = H.mutate(rng, H.concrete(rng, probe), rng.randint(0, 1), 0, len(probe))
            g[pos + 70:pos + 70 + len(pc)] = (pc if i % 3 else H.revcomp(pc)).encode()
        if inner:
            H.plant_amplicon(rng, g, pos + 110, inner[0], inner[1], rng.randint(60, 150), 0, rng.randint(0, 1), 3)
    return bytes(g)


A, B = "AGAGTTTGATCMTGGCTCAG", "TACGGYTACCTTGTTAYGACTT"


@pytest.mark.gpu
def test_cli_ipcr_modes(tmp_path, oracle):
    rng = random.Random(31)
    fa = _write(tmp_path, "g.fa", [(f"chr{i}", _genome(rng, rng.randint(20000, 50000), A, B, 4)) for i in range(3)])
    fa2 = _write(tmp_path, "h.fa", [("plasmid", _genome(rng, 9000, A, B, 2))])
    for extra, od in ((["--sort"], dict(sort=True)),
                      (["--sort", "--chunk-size", "7000"], dict(sort=True, chunk_size=7000)),
                      (["--sort", "--circular", "--no-header"], dict(sort=True, circular=True, header=False)),
                      (["--sort", "--self=false", "--terminal-window", "0", "--hit-cap", "0"],
                       dict(sort=True, self_pairs=False, terminal_window=0, hit_cap=0)),
                      ([], dict())):
        od = dict(mismatches=2, **od)
        got = run_cli("ipcr", "-f", A, "-r", B.lower(), "-m", 2, *extra, fa, fa2)
        want = R.run(oracle, "ipcr", [fa, fa2], od, fwd=A, rev=B)
        assert got == want, extra
    # config 1: demo input => header only, exit code from --no-match-exit-code
    demo = tmp_path / "demo.fa"
    demo.write_bytes(b">s\nACGTACGTACGT\n")
    out = run_cli("ipcr", "-f", A, "-r", B, "-m", 1, "--no-match-exit-code", 4, demo, expect_rc=4)
    assert out == R.HDR + "\n"


@pytest.mark.gpu
def test_cli_primer_tsv_and_multiplex(tmp_path, oracle):
    from ipcr_b200 import synth

    rng = random.Random(32)
    user, _ = synth.multiplex_panel(8, seed=5, self_pairs=False)
    g = H.random_genome(rng, 90000)
    for i, p in enumerate(user):
        H.plant_amplicon(rng, g, 800 + i * 10000, p.forward, p.reverse, rng.randint(100, 1500), rng.randint(0, 2),
                         rng.randint(0, 2), 3, revcomp_strand=i % 2 == 1)
    fa = _write(tmp_path, "m.fa", [("chrM", bytes(g))])
    tsv = tmp_path / "panel.tsv"
    tsv.write_text("# id fwd rev min max\n" + "".join(f"{p.id}\t{p.forward}\t{p.reverse}\t0\t1800\n" for p in user))
    pairs = [(p.id, p.forward, p.reverse, 0, 1800) for p in user]
    od = dict(mismatches=2, sort=True)
    assert run_cli("multiplex", "--primers", tsv, "-m", 2, "--sort", fa) == \
        R.run(oracle, "multiplex", [fa], od, pairs=pairs)
    assert run_cli("ipcr", "--primers", tsv, "-m", 2, "--sort", fa) == R.run(oracle, "ipcr", [fa], od, pairs=pairs)


@pytest.mark.gpu
def test_cli_probe_and_nested(tmp_path, oracle):
    rng = random.Random(33)
    probe = "ACTGGCTCACCTTGCGAGTCATGC"
    ia, ib = "GCAACGAGCGCAACCC", "ACGGGCGGTGTGTRC"
    fa = _write(tmp_path, "p.fa", [(f"c{i}", _genome(rng, 40000, A, B, 5, probe=probe, inner=(ia, ib), n_frac=0))
                                   for i in range(3)])
    od = dict(mismatches=2, sort=True)
    for pmm, req in ((0, True), (1, False)):
        got = run_cli("probe", "-f", A, "-r", B, "-m", 2, "--sort", "--probe", probe, "--probe-name", "taq",
                      "--probe-max-mm", pmm, f"--require-probe={'true' if req else 'false'}", fa)
        want = R.run(oracle, "probe", [fa], od, fwd=A, rev=B, probe=probe, probe_name="taq", probe_max_mm=pmm,
                     require_probe=req)
        assert got == want
    od = dict(mismatches=1, sort=True)
    for req in (False, True):
        got = run_cli("nested", "-f", A, "-r", B, "-m", 1, "--sort", "--inner-forward", ia, "--inner-reverse", ib,
                      *(["--require-inner"] if req else []), fa)
        want = R.run(oracle, "nested", [fa], od, fwd=A, rev=B, inner_fwd=ia, inner_rev=ib, require_inner=req)
        assert got == want
