#!/bin/bash
# round 2: strided first level with the T second-level lookups of an anchor issued together: large panels
mkdir -p gpurun_out
out=gpurun_out/r02_seed_stride4.txt
: > $out
FLAGS="--steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-e2e-fasta --no-n-variant --no-bitparallel"
run() {
  env IPCR_SEED_DEBUG=1 "$@" timeout 300 python bench.py $FLAGS $EXTRA > gpurun_out/stride.json 2>gpurun_out/stride.err
  python - "$* $EXTRA" >> $out <<'PY'
import sys,json
ix=[l.strip() for l in open('gpurun_out/stride.err') if l.startswith('seed index')]
for l in open('gpurun_out/stride.json'):
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[1], '|', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['stage_ms'].items()}, d['candidates'], d['sites'], d['products'], '|', ix[0] if ix else '')
PY
}
for P in 100 50 25; do
  EXTRA="--pairs $P"
  run IPCR_SEED_STRIDE=1
  run IPCR_SEED_STRIDE=2
  run IPCR_SEED_STRIDE=4
done
for P in 10 3 1; do
  EXTRA="--pairs $P"
  run IPCR_SEED_STRIDE=2
  run IPCR_SEED_STRIDE=4
done
EXTRA="--pairs 1 --mm 1" run IPCR_NOOP=1
EXTRA="--pairs 1 --mm 3" run IPCR_NOOP=1
EXTRA="--pairs 10 --mm 3" run IPCR_NOOP=1
EXTRA="--pairs 10 --mm 3" run IPCR_SEED_STRIDE=1
cat $out
