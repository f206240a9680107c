#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_375.csv python bench.py --steps 2 --warmup 3 --genome-mbp 375 --no-cpu-baseline --no-e2e --no-bitparallel > gpurun_out/ncu.log 2>&1
grep -v "^==" gpurun_out/r02_launches_375.csv | python -c "
import csv,sys
for r in csv.DictReader(sys.stdin):
    print(r['ID'], r['Kernel Name'][:40], r['Block Size'], r['Grid Size'], r['Metric Value'])
" | tail -14
