#!/bin/bash
# round 2: split first-level key (low A + high B bases of the 16-base seed window) vs the plain high-12 key
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r02_pytest_split.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_split.log
out=gpurun_out/r02_seed_split.txt
: > $out
run() {
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-bitparallel --no-e2e-fasta --no-n-variant $EXTRA 2>>gpurun_out/r02_split.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$* $EXTRA', 'step', round(d['ms_per_step'],4), 'filter', round(d['stage_ms']['filter_ms'],4), 'cand', d['candidates'], 'products', d['products'])
" >> $out
}
EXTRA=""
run IPCR_NOOP=1
run IPCR_SEED_NO_SPLIT=1
for ab in 6,6 5,7 7,5 6,7 7,6 5,6 6,5 7,7 8,6 6,8; do run IPCR_SEED_SPLIT=$ab; done
run IPCR_SEED_SPLIT=6,6 IPCR_SEED_BLOOM_K=1
EXTRA="--n-per-mbp 10" run IPCR_NOOP=1
EXTRA="--n-per-mbp 10" run IPCR_SEED_NO_SPLIT=1
EXTRA="--pairs 1" run IPCR_NOOP=1
EXTRA="--pairs 1" run IPCR_SEED_NO_SPLIT=1
cat $out
