#!/bin/bash
# round 2: strided first level after the cost model was refitted (expected-max walk term): the model's choice per
# panel size against the forced strides, 3 Gbp, mm 2 (+ the small BASELINE config shapes)
mkdir -p gpurun_out
out=gpurun_out/r02_seed_stride3.txt
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
for P in 1 3 10 25 50; do
  EXTRA="--pairs $P"
  run IPCR_NOOP=1
  run IPCR_SEED_STRIDE=1
  [ $P -ge 10 ] && run IPCR_SEED_STRIDE=2
done
EXTRA="--pairs 1 --mm 3" run IPCR_NOOP=1
EXTRA="--pairs 1 --mm 1" run IPCR_NOOP=1
EXTRA="" run IPCR_NOOP=1
EXTRA="--genome-mbp 500 --contigs 10 --pairs 1 --mm 1" run IPCR_NOOP=1
EXTRA="--genome-mbp 4.6 --contigs 1 --pairs 1 --mm 2" run IPCR_NOOP=1
cat $out
