#!/bin/bash
# sweep of the seed-scan kernel's knobs on the benchmark workload (run under gpurun)
mkdir -p gpurun_out
out=gpurun_out/seed_sweep.txt
: > $out
run() {
  env "$@" timeout 200 python bench.py --filter seed --steps 3 --warmup 3 --no-cpu-baseline --no-e2e $EXTRA 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$* $EXTRA', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['stage_ms'].items()}, d['candidates'], d['products'])
" >> $out
}
EXTRA="--pairs 1" run IPCR_SEED_BLOOM_K=1
EXTRA="--pairs 1" run IPCR_SEED_BLOOM_K=2
EXTRA=""
for S in 14 15 16; do for K in 1 2; do run IPCR_SEED_LEN=$S IPCR_SEED_BLOOM_K=$K; done; done
cat $out
