#!/bin/bash
# round 2: sweep of the two-level seed index (S, S1, bitmap bits) on the benchmark workload (run under gpurun)
mkdir -p gpurun_out
out=gpurun_out/r02_seed_sweep.txt
: > $out
run() {
  env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-bitparallel $EXTRA 2>>gpurun_out/sweep.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$* $EXTRA', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['stage_ms'].items()}, d['candidates'], d['sites'], d['products'])
" >> $out
}
EXTRA=""
run IPCR_NOOP=1
for S1 in 10 11 12 13 14 16; do for K in 1 2; do run IPCR_SEED_LEN=16 IPCR_SEED_LEN1=$S1 IPCR_SEED_BLOOM_K=$K; done; done
for S1 in 11 12 13 15; do run IPCR_SEED_LEN=15 IPCR_SEED_LEN1=$S1 IPCR_SEED_BLOOM_K=2; done
EXTRA="--n-per-mbp 10" run IPCR_NOOP=1
EXTRA="--mm 3" run IPCR_NOOP=1
EXTRA="--mm 3" run IPCR_SEED_LEN=16 IPCR_SEED_LEN1=12 IPCR_SEED_BLOOM_K=1 IPCR_SEED_BLOOM=1
EXTRA="--mm 3" run IPCR_SEED_LEN=16 IPCR_SEED_LEN1=13 IPCR_SEED_BLOOM_K=1 IPCR_SEED_BLOOM=1
EXTRA="--pairs 1" run IPCR_NOOP=1
cat $out
