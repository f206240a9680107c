#!/bin/bash
# sweep of the seed-scan kernel's knobs on the benchmark workload (run under gpurun)
mkdir -p gpurun_out
out=gpurun_out/seed_sweep.txt
: > $out
run() {
  env "$@" timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-bitparallel $EXTRA 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$* $EXTRA', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['stage_ms'].items()}, d['candidates'], d['products'])
" >> $out
}
EXTRA=""
for B in 1 2 3 4; do run IPCR_SEED_BATCH=$B; done
EXTRA="--mm 3" run IPCR_SEED_BATCH=1
EXTRA="--mm 3" run IPCR_SEED_BATCH=4
cat $out
