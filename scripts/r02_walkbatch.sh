#!/bin/bash
# round 2: second-level lookups of a lane's passes issued NB at a time (IPCR_SEED_VAR 2: NB = 2, 8: NB = 3), headline panel
mkdir -p gpurun_out
out=gpurun_out/r02_seed_walkbatch.txt
: > $out
FLAGS="--steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-e2e-fasta --no-n-variant --no-bitparallel"
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
EXTRA=""
run IPCR_SEED_VAR=0
run IPCR_SEED_VAR=2
run IPCR_SEED_VAR=8
EXTRA="--pairs 10 --mm 3"
run IPCR_SEED_VAR=0
run IPCR_SEED_VAR=2
run IPCR_SEED_VAR=8
EXTRA="--pairs 50"
run IPCR_SEED_VAR=0 IPCR_SEED_STRIDE=1
run IPCR_SEED_VAR=2 IPCR_SEED_STRIDE=1
cat $out
