#!/bin/bash
# baseline/run_go_reference.sh -- the reference's own Go binary as the CPU baseline.
#
# Iff a Go toolchain (>= 1.22, go.mod:3) is on PATH: build the UNMODIFIED reference out of a scratch copy of
# $IPCR_REFERENCE (default /root/reference; its own `make build`, Makefile:33-39) and time
#     ipcr --sort --threads $(nproc)                            (the defaults: one worker per record)
#     ipcr --sort --threads $(nproc) --chunk-size 4000000        (its best case on few long records)
#     ipcr --sort --threads $(nproc) --chunk-size 4000000 --hit-cap 0
# on the FASTA / primer files given.  Otherwise it says so and exits 0: this image and the GPU boxes carry no Go
# toolchain and there is no network, which is why bench.py times the C restatement of the same engine
# (oracle/ipcr_seeded.c) instead and labels it "port".
#
# usage: baseline/run_go_reference.sh genome.fa primers.tsv [mismatches]   (primers.tsv: id fwd rev [min max])
set -u
FASTA=${1:-}
PRIMERS=${2:-}
MM=${3:-2}
REF=${IPCR_REFERENCE:-/root/reference}
if ! command -v go >/dev/null 2>&1; then
    echo "skipped: no Go toolchain on PATH (go version: command not found); see bench.py --impl reference for the C port"
    exit 0
fi
if [ -z "$FASTA" ] || [ -z "$PRIMERS" ] || [ ! -d "$REF" ]; then
    echo "usage: $0 genome.fa primers.tsv [mismatches]   (reference sources expected at $REF)" >&2
    exit 2
fi
WORK=$(mktemp -d)
trap 'rm -rf "$WORK"' EXIT
cp -r "$REF" "$WORK/ref" || exit 1
( cd "$WORK/ref" && make build ) >"$WORK/build.log" 2>&1 || { echo "reference build failed:"; tail -20 "$WORK/build.log"; exit 1; }
BIN=$(ls "$WORK"/ref/bin/ipcr* 2>/dev/null | head -1)
[ -x "$BIN" ] || { echo "no ipcr binary under bin/ after make build"; exit 1; }
CORES=$(nproc)
BASES=$(grep -v '^>' "$FASTA" | tr -d '\n\r \t' | wc -c)
echo "go: $(go version)   cores: $CORES   genome: $BASES bp   mismatches: $MM"
for FLAGS in "" "--chunk-size 4000000" "--chunk-size 4000000 --hit-cap 0"; do
    START=$(date +%s.%N)
    "$BIN" --primers "$PRIMERS" --mismatches "$MM" --sort --threads "$CORES" $FLAGS "$FASTA" >"$WORK/out.tsv" || { echo "run failed ($FLAGS)"; continue; }
    END=$(date +%s.%N)
    python3 - "$START" "$END" "$BASES" "$(wc -l <"$WORK/out.tsv")" "$FLAGS" <<'PY'
import sys
t0, t1, bases, rows, flags = float(sys.argv[1]), float(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
dt = t1 - t0
print(f"flags [{flags or 'defaults'}]: {dt:.2f} s, {bases / 1e9 / dt:.4f} Gbp/s, {rows - 1} rows (FASTA on disk -> sorted TSV)")
PY
done
