#!/bin/bash
# round 2: tuning switches of k_seed_scan (IPCR_SEED_VAR bits: 1 shuffle look-ahead, 2 mad address, 4 dynamic chunks)
mkdir -p gpurun_out
out=gpurun_out/r02_seed_variants.txt
: > $out
for extra in "" "--genome-mbp 375"; do
for v in 0 1 4 5 0 5; do
  IPCR_SEED_VAR=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-bitparallel --no-e2e-fasta --no-n-variant $extra 2>>gpurun_out/r02_var.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('IPCR_SEED_VAR=$v $extra', 'step', round(d['ms_per_step'],4), 'filter', round(d['stage_ms']['filter_ms'],4), 'products', d['products'])
" >> $out
done
done
cat $out
